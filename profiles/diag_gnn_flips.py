"""How many gradient elements of the GPU propagation miss rtol 1e-3 against the fp64 oracle, per back end and
arithmetic mode, and where they sit (ReLU-gate evidence).  python profiles/diag_gnn_flips.py"""
import importlib, json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "multimodal-fusion-based-pre-routing-timing-prediction-_b200"
for p in (ROOT, os.path.join(ROOT, PKG), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
importlib.import_module(PKG)
from conftest import design_to_oracle
from oracle import restate
import model as M
import tm_graph, tm_lib, tm_ops, tm_synth

for cfg, seed in (("tiny", 5), ("c1", 1)):
    d = tm_synth.make_design(seed=seed, **tm_synth.CONFIGS[cfg])
    torch.manual_seed(seed)
    gnn = M.PathConv(out_feat_dim=128, hidden_feat_dim=128, cell_feat_dim=36, net_feat_dim=2)
    od = design_to_oracle(d)
    torch.manual_seed(seed)
    Gout = torch.zeros(d.n, 128)
    ep = torch.from_numpy(d.endpoints)
    Gout[ep] = torch.randn(ep.numel(), 128)
    Gout += 0.01 * torch.randn(d.n, 128)
    names = ["gnn." + k for k in tm_ops.GNN_PARAM_NAMES]
    sd = {"gnn." + k: v.detach().double().clone().requires_grad_(True) for k, v in gnn.state_dict().items()}
    H64 = restate.gnn_propagate(sd, "gnn", d.n, od["levels"], od["net_csr"], od["cell_csr"], od["cell_feat"].double(), od["net_feat"].double())
    g64 = torch.autograd.grad(H64, [sd[k] for k in names], Gout.double())
    g = tm_graph.TimingGraph(d.n, (d.net_src, d.net_dst), (d.cell_src, d.cell_dst), pis=d.pis)
    g.ndata["cell_feat"], g.ndata["net_feat"] = torch.from_numpy(d.cell_feat), torch.from_numpy(d.net_feat)
    g = g.to("cuda")
    gg = gnn.to("cuda")
    for math in ("tf32x3", "fp32"):
        for impl in (0, 3):
            tm_ops.MATH = math
            tm_lib.lib().tm_gnn_set_impl(impl)
            gg.zero_grad()
            H = gg.propagate(g)
            H.backward(Gout.cuda())
            herr = float((H.detach().double().cpu() - H64.detach()).abs().max() / H64.detach().abs().max())
            rows = {}
            for k, r in zip(tm_ops.GNN_PARAM_NAMES, g64):
                a = dict(gg.named_parameters())[k].grad.double().cpu()
                scale = float(r.abs().max())
                err = (a - r).abs()
                tol = 1e-4 * scale + 1e-3 * r.abs()
                rows[k] = [int((err > tol).sum()), round(float((err / tol).max()), 2)]
            print(json.dumps(dict(cfg=cfg, math=math, impl=impl, H_rel=herr, bad_and_worst=rows)))
