"""Host enqueue time vs device time of one design step (config 2)."""
import importlib, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "multimodal-fusion-based-pre-routing-timing-prediction-_b200"
sys.path.insert(0, os.path.join(ROOT, PKG))
importlib.import_module(PKG)
import tm_engine, tm_synth
d = tm_synth.make_design(seed=0, **tm_synth.CONFIGS["c2"])
model, cnn = tm_engine.build_models(d.map_size, seed=0, device="cuda")
batch = tm_engine.DesignBatch.from_host(tm_engine.HostDesign(d, pin=True), "cuda")
step = tm_engine.DesignStep(model, cnn)
for _ in range(3): step.run(batch)
torch.cuda.synchronize()
for overlap in (True, False):
    step.overlap = overlap
    step.run(batch); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10): step.run(batch)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"overlap={overlap}: host enqueue {1e3*(t1-t0)/10:.2f} ms/step, total {1e3*(t2-t0)/10:.2f} ms/step")
