"""Where does the config-2 design step spend its time?  Captures the step as CUDA graphs in five forms and times
each replay: full two-stream step, one stream (the sum of all kernels, warm), the netlist branch alone (U-Net
outputs cached), the image branch alone (propagation outputs cached), and the head alone (both cached).
Usage: python profiles/diag_step_branches.py   -> JSON lines."""
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_NAME = "multimodal-fusion-based-pre-routing-timing-prediction-_b200"
for p in (ROOT, os.path.join(ROOT, PKG_NAME)):
    sys.path.insert(0, p)
importlib.import_module(PKG_NAME)
import tm_engine  # noqa: E402
import tm_ops  # noqa: E402
import tm_synth  # noqa: E402
import tm_unet  # noqa: E402


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(iters):
        fn()
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / iters


def main():
    dev = torch.device("cuda", 0)
    d = tm_synth.make_design(seed=0, **tm_synth.CONFIGS[os.environ.get("TM_DIAG_CFG", "c2")])
    model, cnn = tm_engine.build_models(d.map_size, seed=0, device=dev)
    host = tm_engine.HostDesign(d, pin=True)
    batch = tm_engine.DesignBatch.from_host(host, dev)
    step = tm_engine.DesignStep(model, cnn)
    for _ in range(3):
        step.run(batch)
    torch.cuda.synchronize()
    out = {}
    out["full"] = timeit(step.capture(batch))
    step.overlap = False
    out["one_stream"] = timeit(step.capture(batch))
    step.overlap = True

    real = dict(uf=tm_unet.unet_forward, ub=tm_unet.unet_backward, gf=tm_ops.gnn_forward, gb=tm_ops.gnn_backward)
    cache = {}

    def cached(name):
        def fn(*a, **k):
            if name not in cache:
                cache[name] = real[name](*a, **k)
            return cache[name]
        return fn

    def variant(names):
        cache.clear()
        tm_unet.unet_forward, tm_unet.unet_backward = (cached("uf"), cached("ub")) if "u" in names else (real["uf"], real["ub"])
        tm_ops.gnn_forward, tm_ops.gnn_backward = (cached("gf"), cached("gb")) if "g" in names else (real["gf"], real["gb"])
        step.run(batch)                 # fills the caches eagerly
        torch.cuda.synchronize()
        t = timeit(step.capture(batch))
        tm_unet.unet_forward, tm_unet.unet_backward = real["uf"], real["ub"]
        tm_ops.gnn_forward, tm_ops.gnn_backward = real["gf"], real["gb"]
        return t

    out["netlist_branch_only(unet cached)"] = variant("u")
    out["image_branch_only(gnn cached)"] = variant("g")
    out["head_only(both cached)"] = variant("ug")
    print(json.dumps({k: round(v, 4) for k, v in out.items()}), flush=True)
    step.close()


if __name__ == "__main__":
    main()
