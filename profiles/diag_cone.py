"""How much of the netlist is in the backward cone of one endpoint batch?  Fraction of all-zero rows of dLoss/dz per
pin class and per level after one backward sweep of config 2 (1 350 endpoints).  Usage: python profiles/diag_cone.py"""
import importlib, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_NAME = "multimodal-fusion-based-pre-routing-timing-prediction-_b200"
for p in (ROOT, os.path.join(ROOT, PKG_NAME), os.path.join(ROOT, "profiles")):
    sys.path.insert(0, p)
importlib.import_module(PKG_NAME)
import tm_engine, tm_ops, tm_synth  # noqa: E402

dev = torch.device("cuda", 0)
for cfg in ("c2",):
    d = tm_synth.make_design(seed=0, **tm_synth.CONFIGS[cfg])
    model, cnn = tm_engine.build_models(d.map_size, seed=0, device=dev)
    batch = tm_engine.DesignBatch.from_host(tm_engine.HostDesign(d, pin=True), dev)
    sched = batch.graph.schedule()
    gp = [p.detach() for p in tm_engine.DesignStep(model, cnn).gnn_params]
    H, saved = tm_ops.gnn_forward(sched, batch.cell_feat, batch.net_feat, gp, save=True)
    G = torch.zeros(sched.n, 128, device=dev)
    G[batch.endpoints.long()] = torch.randn(batch.endpoints.numel(), 128, device=dev)
    tm_ops.gnn_backward(sched, saved, gp, G)
    torch.cuda.synchronize()
    nz = (G != 0).any(dim=1)
    lv = sched.level.long()
    out = dict(cfg=cfg, n=sched.n, endpoints=int(batch.endpoints.numel()), active_rows=float(nz.float().mean()),
               active_cell_class=float(nz[sched.cell_class.long()].float().mean()),
               active_net_class=float(nz[sched.net_class.long()].float().mean()))
    per = []
    for l in range(0, sched.num_levels, 10):
        m = lv == l
        per.append((l, int(m.sum()), round(float(nz[m].float().mean()), 3)))
    out["per_level(level, pins, active)"] = per
    # structural cone: reachability from the endpoints (ignores ReLU zeros)
    print(json.dumps(out))
