"""BASELINE.json configurations 2, 3 and 4 at FULL size (the headline bench runs config 2; the
other configs are parity-test cases).

* config 2 (100k cells, 331 819 pins, 101 levels): forward propagation H against the oracle on the
  whole design (the oracle's forward takes a few seconds), schedule bit-exact.
* config 3 (300k cells, ~1M pins, 101 levels): the oracle would take minutes, so the check is the
  size-independent property of the recurrence itself -- EVERY pin must satisfy its own level equation
  given the H rows of its predecessors (src/model.py:88-116,138-153) -- evaluated in fp64 on a random
  sample of pins of every kind, plus bit-determinism of a second run and the schedule invariants.
* config 4 (U-Net, batch 32 of 512x512, bf16 tensor-core operands): batch 2 against the oracle at
  the bf16 bar (rtol 2e-2) on the OUTPUT; batch 32 forward+backward for shapes, finiteness and
  determinism.  Parameter gradients of this network are not a 2e-2 quantity at 8-bit operands for
  ANY implementation: 18 layers of max-pool arg-max and ReLU gates re-route the upstream gradient
  whenever an activation moves across a tie, so they are checked by direction and size (cosine
  >= 0.85, norm within 25 %); `profiles/diag_c4.py` prints the per-mode figures (3xTF32: 3e-3
  relative L2, two-term bf16: 1e-2, single TF32: 1e-1, bf16: 3e-1).  The fp32-class default mode is
  held to rtol 1e-3 on every gradient in test_gpu_parity.py::test_unet_vs_oracle.
"""
import numpy as np
import pytest
import torch

import tm_synth
from conftest import assert_close, design_to_oracle
from oracle import restate

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def mods(pkg):
    import tm_graph
    import tm_ops
    import tm_unet
    return {"graph": tm_graph, "ops": tm_ops, "unet": tm_unet}


def _gnn(seed):
    import model as M
    torch.manual_seed(seed)
    return M.PathConv(out_feat_dim=128, hidden_feat_dim=128, cell_feat_dim=36, net_feat_dim=2)


def _graph(mods, d):
    g = mods["graph"].TimingGraph(d.n, (torch.from_numpy(d.net_src), torch.from_numpy(d.net_dst)),
                                  (torch.from_numpy(d.cell_src), torch.from_numpy(d.cell_dst)), pis=torch.from_numpy(d.pis))
    g.ndata["cell_feat"] = torch.from_numpy(d.cell_feat)
    g.ndata["net_feat"] = torch.from_numpy(d.net_feat)
    return g.to(DEV)


def _check_schedule(sched, d):
    lvl = sched.level.cpu().numpy()
    assert np.array_equal(lvl, d.level), "pin -> level map differs from the generator's longest-path levels"
    for src, dst in ((d.net_src, d.net_dst), (d.cell_src, d.cell_dst)):
        assert bool((lvl[src] < lvl[dst]).all()), "an edge does not go up in level"
    order = sched.order.cpu().numpy()
    ptr = sched.level_ptr.cpu().numpy() if hasattr(sched, "level_ptr") else None
    assert np.array_equal(np.sort(order), np.arange(d.n)), "the schedule is not a permutation of the pins"
    assert bool((np.diff(lvl[order]) >= 0).all()), "pins are not ordered by level"
    if ptr is not None:
        assert ptr[0] == 0 and ptr[-1] == d.n


def test_config2_forward_vs_oracle(mods):
    d = tm_synth.make_design(seed=0, **tm_synth.CONFIGS["c2"])
    gnn = _gnn(0)
    sd = {"gnn." + k: v.detach().clone() for k, v in gnn.state_dict().items()}
    od = design_to_oracle(d)
    with torch.no_grad():
        Href = restate.gnn_propagate(sd, "gnn", d.n, od["levels"], od["net_csr"], od["cell_csr"], od["cell_feat"], od["net_feat"])
    g = _graph(mods, d)
    _check_schedule(g.schedule(), d)
    with torch.no_grad():
        H = gnn.to(DEV).propagate(g)
    assert_close(H, Href, 1e-3, 1e-4, "H (config 2, 331 819 pins)")


def _recurrence_residual(d, gnn, H, S, pins):
    """max |h[v] - f(H[preds])| / scale over the sampled pins, in fp64."""
    sd = {k: v.detach().double().cpu() for k, v in gnn.state_dict().items()}
    H = H.double().cpu()
    S = S.double().cpu()
    order = np.argsort(d.net_dst, kind="stable"); nptr = np.searchsorted(d.net_dst[order], np.arange(d.n + 1)); nsrc = d.net_src[order]
    order = np.argsort(d.cell_dst, kind="stable"); cptr = np.searchsorted(d.cell_dst[order], np.arange(d.n + 1)); csrc = d.cell_src[order]
    W1, b1 = sd["fc_cell_neigh.layers.0.weight"], sd["fc_cell_neigh.layers.0.bias"]
    W2, b2 = sd["fc_cell_neigh.layers.2.weight"], sd["fc_cell_neigh.layers.2.bias"]
    worst = 0.0
    for v in pins:
        lv = int(d.level[v])
        if lv == 0:
            want = torch.relu(S[v])
        elif lv & 1:
            src = nsrc[nptr[v]:nptr[v + 1]]
            a = H[src].mean(0) if src.size else torch.zeros(128, dtype=torch.float64)
            want = torch.relu(S[v] + a)
        else:
            src = csrc[cptr[v]:cptr[v + 1]]
            if src.size:
                m = H[src]
                a = (m * torch.softmax(m, 0)).sum(0)
            else:
                a = torch.zeros(128, dtype=torch.float64)
            want = torch.relu(S[v] + W2 @ torch.relu(W1 @ a + b1) + b2)
        worst = max(worst, float((H[v] - want).abs().max()) / max(1.0, float(want.abs().max())))
    return worst


def test_config3_million_pin_recurrence(mods):
    ops = mods["ops"]
    d = tm_synth.make_design(seed=0, n_endpoints=64, **tm_synth.CONFIGS["c3"])
    assert d.n > 900_000 and d.num_levels == 101
    gnn = _gnn(3).to(DEV)
    g = _graph(mods, d)
    sched = g.schedule()
    _check_schedule(sched, d)
    params = [dict(gnn.named_parameters())[k].detach() for k in ops.GNN_PARAM_NAMES]
    with torch.no_grad():
        H, _ = ops.gnn_forward(sched, g.ndata["cell_feat"], g.ndata["net_feat"], params, save=False)
        H2, _ = ops.gnn_forward(sched, g.ndata["cell_feat"], g.ndata["net_feat"], params, save=False)
    assert torch.equal(H, H2), "propagation is not bit-deterministic"
    assert bool(torch.isfinite(H).all())
    # hoisted self terms in fp64 on the host for the sampled pins only
    rng = np.random.default_rng(0)
    pins = np.concatenate([rng.choice(np.flatnonzero(d.level == 0), 50, replace=False),
                           rng.choice(np.flatnonzero((d.level & 1) == 1), 400, replace=False),
                           rng.choice(np.flatnonzero((d.level > 0) & ((d.level & 1) == 0)), 400, replace=False)])
    cpu = {k: v.detach().double().cpu() for k, v in gnn.state_dict().items()}
    S = torch.zeros(d.n, 128, dtype=torch.float64)
    cf, nf = torch.from_numpy(d.cell_feat).double(), torch.from_numpy(d.net_feat).double()
    for v in pins:
        if d.level[v] & 1:
            S[v] = cpu["fc_net_self.layers.2.weight"] @ torch.relu(cpu["fc_net_self.layers.0.weight"] @ nf[v] + cpu["fc_net_self.layers.0.bias"]) \
                + cpu["fc_net_self.layers.2.bias"]
        else:
            S[v] = cpu["fc_cell_self.layers.2.weight"] @ torch.relu(cpu["fc_cell_self.layers.0.weight"] @ cf[v] + cpu["fc_cell_self.layers.0.bias"]) \
                + cpu["fc_cell_self.layers.2.bias"]
    res = _recurrence_residual(d, gnn, H, S, pins)
    assert res < 1e-4, f"a pin violates its level equation: relative residual {res:.3e}"


def test_config4_unet_bf16(mods):
    import Unet as U
    unet = mods["unet"]
    torch.manual_seed(4)
    net = U.UNet("max").train()
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    # ---- batch 2 of 512x512 against the oracle at the bf16 bar
    x = torch.rand(2, 3, 512, 512)
    P = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in sd.items()}
    ref, _ = restate.unet_forward(P, x, "max")
    g = torch.randn_like(ref)
    names = [k for k, _ in net.named_parameters()]
    gref = torch.autograd.grad(ref, [P[k] for k in names], g)
    net = net.to(DEV)
    old = unet.MATH
    try:
        unet.MATH = "bf16"
        out = net(x.to(DEV))
        assert_close(out, ref, 2e-2, 2e-2, "unet out (bf16 operands)")
        out.backward(g.to(DEV))
        for k, r in zip(names, gref):
            a = dict(net.named_parameters())[k].grad.double().cpu().reshape(-1)
            r = r.double().reshape(-1)
            cos = float((a @ r) / (a.norm() * r.norm() + 1e-300))
            ratio = float(a.norm() / (r.norm() + 1e-300))
            assert cos >= 0.85 and 0.75 <= ratio <= 1.25, f"{k}: cosine {cos:.3f}, norm ratio {ratio:.3f}"
        # ---- the BASELINE shape: batch 32 of 512x512, forward + backward
        net.zero_grad()
        torch.manual_seed(5)
        xb = torch.rand(32, 3, 512, 512, device=DEV)
        o1 = net(xb)
        assert o1.shape == (32, 1, 256, 256) and bool(torch.isfinite(o1).all())
        o1.sum().backward()
        assert all(p.grad is not None and bool(torch.isfinite(p.grad).all()) for p in net.parameters())
        with torch.no_grad():
            o2 = net(xb)
        assert torch.equal(o1, o2), "U-Net forward is not bit-deterministic"
    finally:
        unet.MATH = old
