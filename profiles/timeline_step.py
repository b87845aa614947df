"""Kernel timeline of ONE replay of the captured config-2 design step (torch.profiler / CUPTI): start, duration and
stream of every kernel, warm and with the real two-stream concurrency.  Writes gpurun_out/timeline_step.tsv and
prints per-stream summaries.  Usage: python profiles/timeline_step.py [one_stream]"""
import importlib
import os
import sys
from collections import defaultdict

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_NAME = "multimodal-fusion-based-pre-routing-timing-prediction-_b200"
for p in (ROOT, os.path.join(ROOT, PKG_NAME)):
    sys.path.insert(0, p)
importlib.import_module(PKG_NAME)
import tm_engine  # noqa: E402
import tm_synth  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    d = tm_synth.make_design(seed=0, **tm_synth.CONFIGS["c2"])
    model, cnn = tm_engine.build_models(d.map_size, seed=0, device=dev)
    batch = tm_engine.DesignBatch.from_host(tm_engine.HostDesign(d, pin=True), dev)
    step = tm_engine.DesignStep(model, cnn)
    step.overlap = "one_stream" not in sys.argv
    for _ in range(3):
        step.run(batch)
    replay = step.capture(batch)
    for _ in range(5):
        replay()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        replay()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ks = []
    for e in evs:
        tr = e.time_range
        ks.append((tr.start, tr.end - tr.start, getattr(e, "device_index", 0), e.name))
    ks.sort()
    if not ks:
        print("no CUDA events captured")
        return
    t0 = ks[0][0]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    # stream ids are not exposed on FunctionEvent in every version: recover them from the chrome trace
    trace_path = os.path.join(ROOT, "gpurun_out", "timeline_step_trace.json")
    prof.export_chrome_trace(trace_path)
    import json
    tr = json.load(open(trace_path))
    rows = []
    for ev in tr["traceEvents"]:
        if ev.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "ts" in ev:
            rows.append((float(ev["ts"]), float(ev["dur"]), ev.get("args", {}).get("stream", -1), ev["name"]))
    rows.sort()
    t0 = rows[0][0]
    with open(os.path.join(ROOT, "gpurun_out", "timeline_step.tsv"), "w") as f:
        f.write("start_us\tdur_us\tstream\tname\n")
        for ts, dur, st, name in rows:
            f.write(f"{ts - t0:.1f}\t{dur:.1f}\t{st}\t{name[:140]}\n")
    os.remove(trace_path)
    end = max(ts + dur for ts, dur, _, _ in rows) - t0
    print(f"kernels {len(rows)}  span {end:.1f} us")
    per = defaultdict(lambda: [0, 0.0])
    for ts, dur, st, name in rows:
        per[st][0] += 1
        per[st][1] += dur
    for st, (n, busy) in per.items():
        print(f"stream {st}: {n} kernels, busy {busy:.1f} us")
    agg = defaultdict(lambda: [0, 0.0])
    for ts, dur, st, name in rows:
        key = (st, name.split("(")[0][:90])
        agg[key][0] += 1
        agg[key][1] += dur
    for (st, name), (n, tot) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
        print(f"{tot:9.1f} us  n={n:4d}  avg {tot / n:7.1f}  s{st}  {name}")
    step.close()


if __name__ == "__main__":
    main()
