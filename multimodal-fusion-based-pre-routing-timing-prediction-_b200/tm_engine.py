"""Fused design step: U-Net + level-wise GNN + mask fusion + head, forward and backward.

``DesignStep.run(batch)`` performs what one optimisation step of the reference does to one design
with all sampled endpoints in a batch (src/train.py:465 CNN forward, :490-511 level loop,
:513-522 MSE on arrival time, :552-553 backward) as a straight sequence of libtm_b200 launches
with no autograd graph, then leaves the gradients in ``param.grad`` (like ``optim.zero_grad();
loss.backward()``).  It is numerically the same computation the ``nn.Module`` surface
(``model.py`` / ``Unet.py``) performs through autograd; tests check both against the oracle.

Data parallelism: designs are independent samples, so ranks own disjoint designs
and the only exchange is one gradient all-reduce per step.  ``DesignStep`` posts it in three
buckets on a side stream as soon as each group of gradients exists (head+fusion -> GNN ->
U-Net), so the NCCL transfers over NVLink overlap the remaining backward kernels.
"""
import os

import numpy as np
import torch

import tm_dp
import tm_lib
import tm_ops
import tm_unet
from tm_graph import ConeGraph, MaskCSR, TimingGraph
from tm_lib import call, stream
from tm_ops import D, RELU, gemm_nn, gemm_tn, transpose


def build_models(map_size, pooling="max", seed=0, device="cuda"):
    """Reference wiring (train.py:56-81 with the head sized as model.py:267,280 produce it)."""
    import model as M
    import Unet as U
    torch.manual_seed(seed)
    gnn = M.PathConv(out_feat_dim=128, hidden_feat_dim=128, cell_feat_dim=36, net_feat_dim=2)
    fcn = torch.nn.Linear(map_size * map_size, 128)
    torch.nn.init.xavier_uniform_(fcn.weight, gain=torch.nn.init.calculate_gain("relu"))
    mdl = M.PathModel(gnn, None, fcn, None, None, M.MLP(128 + 128 + 32, 2 * (128 + 128 + 32), 1))
    cnn = U.UNet(pooling)
    if os.environ.get("TM_UNET_MATH"):                 # e.g. "bf16": image branch on the TMA-fed bf16 tensor-core path
        cnn.math = os.environ["TM_UNET_MATH"]
    return mdl.to(device).train(), cnn.to(device).train()


class HostDesign:
    """A design in pinned host memory: what a data loader hands over each step."""

    FIELDS = ("net_src", "net_dst", "cell_src", "cell_dst", "cell_feat", "net_feat", "pis", "endpoints",
              "endpoint_level", "arrival_time", "mask_indptr", "mask_cols", "image")

    def __init__(self, d, pin=True):
        self.n, self.map_size = d.n, d.map_size
        t = {"net_src": d.net_src, "net_dst": d.net_dst, "cell_src": d.cell_src, "cell_dst": d.cell_dst,
             "cell_feat": d.cell_feat, "net_feat": d.net_feat, "pis": d.pis, "endpoints": d.endpoints.astype(np.int32),
             "endpoint_level": d.level[d.endpoints].astype(np.float32), "arrival_time": d.arrival_time,
             "mask_indptr": d.mask_indptr, "mask_cols": d.mask_cols, "image": d.image}
        self.t = {}
        for k, v in t.items():
            x = torch.from_numpy(np.ascontiguousarray(v))
            self.t[k] = x.pin_memory() if pin and torch.cuda.is_available() else x

    def nbytes(self, per_step_only=False):
        keys = ("cell_feat", "net_feat", "endpoints", "endpoint_level", "arrival_time", "mask_indptr",
                "mask_cols", "image") if per_step_only else self.FIELDS
        return int(sum(self.t[k].numel() * self.t[k].element_size() for k in keys))


class DesignBatch:
    """Device-resident batch of one design."""

    def __init__(self, graph, mask_csr, endpoints, endpoint_level, arrival_time, image, cell_feat=None, net_feat=None,
                 rows=None, dynamic=False):
        self.graph, self.mask_csr = graph, mask_csr
        # per-batch feature tensors (a prefetched batch must not disturb the one in flight through
        # the shared graph object); default: whatever the graph carries
        self.cell_feat = cell_feat if cell_feat is not None else graph.ndata["cell_feat"]
        self.net_feat = net_feat if net_feat is not None else graph.ndata["net_feat"]
        self.endpoints = endpoints                 # int32 (T,)
        self.endpoint_level = endpoint_level       # float32 (T,)
        self.arrival_time = arrival_time           # float32 (T,)
        self.image = image                         # (C,H,W)
        # rows: the mask row (path id) of every endpoint -- th.index_select(path_masks, 0, paths), train.py:500;
        # None: endpoint t owns row t (one design's endpoints in one batch, test.py:176)
        self.mask_rows = mask_csr.select(rows) if rows is not None else mask_csr.select_all()
        # dynamic: the endpoint batch (endpoints / levels / labels / mask rows) is refreshed IN PLACE between steps, so
        # the step re-runs the on-device row selection every time (captured into its CUDA graph): one capture per
        # design serves every shuffled batch of the reference's DataLoader (train.py:470-486)
        self.dynamic = bool(dynamic)
        self._cone = None

    def cone(self):
        """``tm_graph.ConeGraph`` of this batch's endpoints (cached): the sub-netlist that can reach an endpoint, on
        which ``DesignStep(..., prune=True)`` runs -- same predictions, loss and gradients over a fraction of the pins.
        None for a dynamic batch, whose endpoints change between replays of one captured step (the step then runs on
        the whole netlist).  The endpoints of a non-dynamic batch must not be modified in place."""
        if self.dynamic:
            return None
        if self._cone is None:
            self._cone = ConeGraph(self.graph, self.endpoints)
        return self._cone

    def set_endpoints(self, endpoints, endpoint_level, arrival_time, rows):
        """Refresh the endpoint batch in place (same batch size): what changes between two DataLoader batches."""
        self.endpoints.copy_(endpoints, non_blocking=True)
        self.endpoint_level.copy_(endpoint_level, non_blocking=True)
        self.arrival_time.copy_(arrival_time, non_blocking=True)
        self.mask_rows.rows.copy_(rows, non_blocking=True)
        self._cone = None

    @staticmethod
    def from_host(h, device, graph=None):
        """Upload a HostDesign.  With ``graph`` (a TimingGraph already on the device, schedule
        cached) only the per-step tensors move: features, endpoints, masks, labels, image."""
        dv = {k: v.to(device, non_blocking=True) for k, v in h.t.items()
              if graph is None or k not in ("net_src", "net_dst", "cell_src", "cell_dst", "pis")}
        if graph is None:
            graph = TimingGraph(h.n, (dv["net_src"], dv["net_dst"]), (dv["cell_src"], dv["cell_dst"]), pis=dv["pis"])
            graph.ndata["cell_feat"], graph.ndata["net_feat"] = dv["cell_feat"], dv["net_feat"]
        mask = MaskCSR(dv["mask_indptr"], dv["mask_cols"], h.map_size * h.map_size)
        return DesignBatch(graph, mask, dv["endpoints"], dv["endpoint_level"], dv["arrival_time"], dv["image"],
                           cell_feat=dv["cell_feat"], net_feat=dv["net_feat"])

    @staticmethod
    def from_synth(d, device):
        return DesignBatch.from_host(HostDesign(d, pin=False), device)


class PreparedDesign:
    VALUE_FIELDS = ("cell_feat", "net_feat", "image", "arrival_time")

    def __init__(self, step, host, device, graph=None, pool=None):
        self.batch = DesignBatch.from_host(host, device, graph=graph)
        self.static = {"cell_feat": self.batch.cell_feat, "net_feat": self.batch.net_feat,
                       "image": self.batch.image, "arrival_time": self.batch.arrival_time}
        torch.cuda.synchronize()
        self.replay = step.capture(self.batch, pool=pool)

    def upload(self, host, stream=None):
        """Host -> device copy of the per-step values into the graph's static inputs."""
        ctx = torch.cuda.stream(stream) if stream is not None else torch.cuda.stream(torch.cuda.current_stream())
        with ctx:
            for k in self.VALUE_FIELDS:
                self.static[k].copy_(host.t[k], non_blocking=True)

    def nbytes(self):
        return int(sum(t.numel() * t.element_size() for t in self.static.values()))

    def step(self, host=None):
        if host is not None:
            self.upload(host)
        return self.replay()


class DesignStep:
    def __init__(self, model, cnn, process_group=None, world_size=1, prune=None):
        """``prune`` (default: environment ``TM_CONE=1``, else off): run the netlist branch on the sub-netlist that can
        reach the batch's endpoints (``tm_graph.ConeGraph``) instead of the whole netlist.  Predictions, loss and every
        gradient are unchanged (``tests/test_gpu_parity.py::test_design_step_on_cone_subnetlist``); pins that no
        endpoint of the batch depends on are simply not evaluated.  Off by default: the headline step evaluates every
        pin, as the reference does."""
        self.model, self.cnn = model, cnn
        self.prune = (os.environ.get("TM_CONE", "0") == "1") if prune is None else bool(prune)
        self.pg, self.world = process_group, world_size
        self.comm_stream = torch.cuda.Stream() if world_size > 1 else None
        # The image branch (U-Net) and the netlist branch (GNN) are independent until the fusion, and
        # both are chains of small latency-bound kernels: they run on two streams, forward and backward.
        self.image_stream = torch.cuda.Stream() if torch.cuda.is_available() else None
        self.overlap = True
        self._buckets = {}            # name -> tm_dp.FlatBucket (persistent flat gradient buffers)
        self._graphs = []             # captured CUDA graphs (released by close())
        g = model.gnn
        sd = dict(g.named_parameters())
        self.gnn_params = [sd[k] for k in tm_ops.GNN_PARAM_NAMES]
        self.cnn_names = [k for k, _ in cnn.named_parameters()]
        self.cnn_params = dict(cnn.named_parameters())

    # ---------------------------------------------------------------- forward pieces
    def _head_image_side(self, b, feat, X):
        """The head's inputs that do not depend on the propagation: mask fusion of the feature map and the level
        embedding (``mlp_alpha``), written into columns D.. of the head matrix X.  Runs on the image stream."""
        m = self.model
        T = int(b.endpoints.numel())
        width = X.shape[1]
        wt = tm_ops.fusion_forward(b.mask_rows, feat, m.fcn.weight.detach(), m.fcn.bias.detach(), X[:, D:], width)
        a0, a2 = m.mlp_alpha.layers[0], m.mlp_alpha.layers[2]
        lv = b.endpoint_level.reshape(T, 1)
        ha = tm_ops.mlp2_forward(lv, 1, None, T, a0.weight.detach(), a0.bias.detach(), a2.weight.detach(),
                                 a2.bias.detach(), X[:, 2 * D:], width)
        return dict(X=X, wt=wt, ha=ha, lv=lv, width=width)

    def _head_netlist_side(self, H, b, hs, endpoints=None):
        """Endpoint rows of H into columns 0..D of X, then ``mlp_fuse``."""
        m = self.model
        T = int(b.endpoints.numel())
        X, width = hs["X"], hs["width"]
        call("tm_gather_cols", T, D, H, D, b.endpoints if endpoints is None else endpoints, X, width, 0, stream())
        f0, f2 = m.mlp_fuse.layers[0], m.mlp_fuse.layers[2]
        pred = torch.empty(T, f2.weight.shape[0], dtype=torch.float32, device=H.device)
        hs["hf"] = tm_ops.mlp2_forward(X, width, None, T, f0.weight.detach(), f0.bias.detach(), f2.weight.detach(),
                                       f2.bias.detach(), pred, pred.shape[1])
        return pred

    def _head_forward(self, H, b, feat):
        T = int(b.endpoints.numel())
        X = torch.empty(T, D + D + self.model.global_dim, dtype=torch.float32, device=H.device)
        hs = self._head_image_side(b, feat, X)
        return self._head_netlist_side(H, b, hs), hs

    def forward(self, b):
        """Inference: predictions for the batch's endpoints (validate(), train.py:137-291)."""
        fmap, _ = tm_unet.unet_forward(self.cnn, b.image, need_bwd=False, update_stats=self.cnn.training)
        sched = b.graph.schedule()
        H, _ = tm_ops.gnn_forward(sched, b.cell_feat, b.net_feat,
                                  [p.detach() for p in self.gnn_params], save=False)
        pred, _ = self._head_forward(H, b, fmap.reshape(-1))
        return pred.squeeze(-1)

    # ---------------------------------------------------------------- full step
    def run(self, b, grad_scale=1.0):
        """Forward + backward.  Returns (loss (1,) device tensor, pred (T,)); fills ``.grad``."""
        m, cnn = self.model, self.cnn
        dev = b.image.device
        main = torch.cuda.current_stream()
        if getattr(b, "dynamic", False):
            b.mask_rows.rebuild()                            # this batch's mask-row selection (7 small kernels, no host sync)
        side = self.image_stream if (self.overlap and self.image_stream is not None) else main
        side.wait_stream(main)                               # fork: inputs / parameters are ready on `main`
        # the netlist branch runs on the sub-netlist that can reach this batch's endpoints (tm_graph.ConeGraph)
        cone = b.cone() if self.prune else None
        sched = (cone.graph if cone is not None else b.graph).schedule()
        endpoints = cone.endpoints if cone is not None else b.endpoints
        x_rows = (cone.cell_x_rows, cone.net_x_rows) if cone is not None else None
        gp = [p.detach() for p in self.gnn_params]
        # the longer chain is enqueued first so the host's launch time for the other one overlaps it
        # (per-level kernels while the image stream runs next to them: measured 5.07 ms/step against 5.49 ms
        #  with the persistent forward, which owns all SMs for 0.9 ms and serialises the U-Net behind it)
        T = int(b.endpoints.numel())
        X = torch.empty(T, D + D + m.global_dim, dtype=torch.float32, device=dev)     # head input [H rows | fused map | level]
        G = torch.empty(sched.n, D, dtype=torch.float32, device=dev)                 # dLoss/dH, seeded at the endpoints
        H, saved = tm_ops.gnn_forward(sched, b.cell_feat, b.net_feat, gp, save=True, x_rows=x_rows,
                                      impl=16 if side is not main and os.environ.get("TM_GNN_IMPL") is None else None)
        img_cap = int(os.environ.get("TM_IMAGE_SM_CAP", "0")) if side is not main else 0
        with torch.cuda.stream(side):
            old_cap = tm_lib.lib().tm_tc_set_grid_cap(img_cap)
            fmap, ust = tm_unet.unet_forward(cnn, b.image, need_bwd=True, update_stats=cnn.training)
            tm_lib.lib().tm_tc_set_grid_cap(old_cap)
            feat = fmap.reshape(-1)
            # everything of the head that does not read H runs here, off the netlist branch's critical path
            # (the image stream is idle long before the 101 levels finish): mask fusion, level embedding, G = 0
            hs = self._head_image_side(b, feat, X)
            G.zero_()
        main.wait_stream(side)                               # join: the head needs the fused feature map
        pred = self._head_netlist_side(H, b, hs, endpoints)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        gpred = torch.empty(T, 1, dtype=torch.float32, device=dev)
        call("tm_mse", T, pred, b.arrival_time, loss, gpred, float(grad_scale), stream())

        # ---- head backward: mlp_fuse on the netlist stream (its dX seeds the backward sweep) ...
        width = hs["width"]
        f0, f2 = m.mlp_fuse.layers[0], m.mlp_fuse.layers[2]
        dw1, db1, dw2, db2, dX = tm_ops.mlp2_backward(X, width, None, T, f0.weight.detach(), f2.weight.detach(),
                                                      hs["hf"], gpred, 1, need_dx=True,
                                                      wgrad_stream=side if side is not main else None)
        side.wait_stream(main)                               # dX is ready
        # ... the level embedding and the mask fusion (whose dF starts the U-Net backward) on the image stream
        a0, a2 = m.mlp_alpha.layers[0], m.mlp_alpha.layers[2]
        with torch.cuda.stream(side):
            da1, dab1, da2, dab2, _ = tm_ops.mlp2_backward(hs["lv"], 1, None, T, a0.weight.detach(), a2.weight.detach(),
                                                           hs["ha"], dX[:, 2 * D:], width, b1=a0.bias.detach())
            dF, dfw, dfb = tm_ops.fusion_backward(b.mask_rows, feat, hs["wt"], dX[:, D:], width)
            head = [(f0.weight, dw1), (f0.bias, db1), (f2.weight, dw2), (f2.bias, db2), (a0.weight, da1),
                    (a0.bias, dab1), (a2.weight, da2), (a2.bias, dab2), (m.fcn.weight, dfw), (m.fcn.bias, dfb)]
            self._assign(head)
            self._post_allreduce("head", [p for p, _ in head])

        # ---- GNN backward (main stream) next to the U-Net backward (image stream)
        call("tm_scatter_add_cols", T, D, dX, width, 0, endpoints, G, D, stream())
        ggrads = tm_ops.gnn_backward(sched, saved, gp, G)
        pairs = list(zip(self.gnn_params, ggrads))
        self._assign(pairs)
        with torch.cuda.stream(side):
            old_cap = tm_lib.lib().tm_tc_set_grid_cap(img_cap)
            ug = tm_unet.unet_backward(cnn, ust, dF.reshape(fmap.shape))
            tm_lib.lib().tm_tc_set_grid_cap(old_cap)
            upairs = [(self.cnn_params[k], ug[k]) for k in self.cnn_names]
            self._assign(upairs)
        main.wait_stream(side)                               # join: every gradient exists on `main`
        # The GNN and U-Net gradients go out as ONE exchange after the join (2.5 MB): posted separately they were two
        # latency-bound all-reduces queued behind each other on the communication stream at the very end of the step
        # (the U-Net's could not start before the GNN's, which waits for the last weight-gradient kernel)
        self._post_allreduce("tail", [p for p, _ in pairs] + [p for p, _ in upairs])
        self._wait_allreduce()
        # a tensor-core kernel whose barrier timed out leaves garbage: the step's loss becomes NaN (loud in any
        # training loop, also under graph replay); tm_lib.check_err_flags() names the cause on the host
        call("tm_poison_on_error", loss, tm_lib.err_flag(dev), stream())
        return loss, pred.squeeze(-1)

    # ---------------------------------------------------------------- CUDA graph
    def capture(self, b, warmup=2, pool=None):
        """Capture ``run(b)`` (about 500 launches on two streams) into one CUDA graph.  The batch's
        tensors become the graph's static inputs: refresh them in place (``tensor.copy_``) and call
        the returned function, which replays the graph and returns the same (loss, pred) tensors;
        ``param.grad`` tensors are rewritten in place by every replay.  On the data-parallel path the
        bucketed NCCL all-reduces (head + fusion early, GNN + U-Net after the join) are captured INSIDE the graph (persistent flat buffers,
        ``tm_dp.FlatBucket``), where they overlap the backward as in the eager schedule; with
        ``TM_DP_GRAPH=0`` the graph holds the rank's compute only and one exchange is posted after each
        replay.  ``pool``: a ``torch.cuda.graph_pool_handle()`` shared by graphs that are replayed one
        after the other (several prepared designs of one rank share their activation memory).
        Call ``close()`` before ``torch.distributed.destroy_process_group()``."""
        if self.world <= 1 or os.environ.get("TM_DP_GRAPH", "1") != "0":
            return self._capture_local(b, warmup, pool)
        world, self.world = self.world, 1                  # no collectives inside the captured region
        try:
            replay_local = self._capture_local(b, warmup, pool)
        finally:
            self.world = world
        params = [p for p in list(self.model.parameters()) + list(self.cnn.parameters())]

        def replay():
            out = replay_local()
            self._post_allreduce("all", [p for p in params if p.grad is not None])
            self._wait_allreduce()
            return out
        replay.graph = replay_local.graph
        return replay

    def _capture_local(self, b, warmup=2, pool=None):
        cur = torch.cuda.current_stream()
        s = torch.cuda.Stream(priority=int(os.environ.get("TM_MAIN_PRIORITY", "0")))   # (a higher priority for the netlist branch was measured: no effect)
        s.wait_stream(cur)
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self.run(b)
        cur.wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, pool=pool):
            loss, pred = self.run(b)
        self._graphs.append(g)
        grads = [(p, p.grad) for p in list(self.model.parameters()) + list(self.cnn.parameters()) if p.grad is not None]

        def replay():
            g.replay()
            for p, gr in grads:                      # several captured steps may coexist: re-point .grad
                p.grad = gr
            return loss, pred
        replay.graph = g
        return replay

    def prepare(self, host, device, graph=None, pool=None):
        """A design made resident for repeated steps: uploads everything once, builds the structure
        that depends only on the netlist and the endpoint batch (level schedule, CSRs, mask runs / CSC)
        and captures the step as a CUDA graph.  Returns a ``PreparedDesign``; its ``step(host)``
        refreshes the per-step VALUES (features, image, labels) from pinned host memory and replays.
        The structure (edges, endpoints, masks) must be the one given here."""
        return PreparedDesign(self, host, device, graph, pool)

    @staticmethod
    def _assign(pairs):
        for p, g in pairs:
            p.grad = g.reshape(p.shape)

    # ---------------------------------------------------------------- data parallel
    def _post_allreduce(self, name, params):
        if self.world <= 1:
            return
        bucket = self._buckets.get(name)
        if bucket is None:
            bucket = self._buckets[name] = tm_dp.FlatBucket(params, self.world, self.pg)
        bucket.post(self.comm_stream)

    def _wait_allreduce(self):
        if self.world > 1 and self.comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self.comm_stream)

    def close(self):
        """Release the captured graphs (they hold NCCL kernels of this step's communicator): call before
        ``torch.distributed.destroy_process_group()``."""
        torch.cuda.synchronize()
        for g in self._graphs:
            g.reset()
        self._graphs = []
        torch.cuda.synchronize()

    def adam_step(self, state, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        """Fused Adam over every parameter that has a gradient (train.py:431-435,555): ONE launch for all ~70
        tensors (``tm_adam_multi``).  The pointer / chunk tables live on the device and are rebuilt only when a
        ``.grad`` tensor moved (never under CUDA-graph replay, where gradients are rewritten in place)."""
        import numpy as np
        state["step"] = state.get("step", 0) + 1
        ps = [p for p in list(self.model.parameters()) + list(self.cnn.parameters()) if p.grad is not None]
        if not ps:
            return
        for p in ps:
            if id(p) not in state:
                state[id(p)] = (torch.zeros_like(p), torch.zeros_like(p))
            if not p.grad.is_contiguous():
                p.grad = p.grad.contiguous()
        key = tuple((p.data_ptr(), p.grad.data_ptr()) for p in ps)
        tab = state.get("_table")
        if tab is None or tab[0] != key:
            chunk = int(tm_lib.lib().tm_adam_chunk())
            rows = np.zeros((len(ps), 5), dtype=np.int64)
            ct, co = [], []
            for i, p in enumerate(ps):
                m, v = state[id(p)]
                rows[i] = (p.data_ptr(), p.grad.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel())
                for off in range(0, p.numel(), chunk):
                    ct.append(i)
                    co.append(off)
            dev = ps[0].device
            tab = state["_table"] = (key, torch.from_numpy(rows).to(dev), torch.tensor(ct, dtype=torch.int32, device=dev),
                                     torch.tensor(co, dtype=torch.int32, device=dev))
        call("tm_adam_multi", int(tab[2].numel()), tab[1], tab[2], tab[3], lr, betas[0], betas[1], eps, weight_decay,
             state["step"], 1.0, stream())
