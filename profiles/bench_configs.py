"""Device timings of BASELINE.json configs 3, 4 and 5 (config 2 is bench.py's headline line).

  config 3: GNN-only level-wise propagation on the ~1M-pin graph: forward / backward pass time,
            achieved algorithmic GB/s against the measured HBM copy peak.
  config 4: U-Net alone, batch 32 of 512x512, bf16 tensor-core operands: forward and forward+backward
            time, algorithmic TFLOP/s (593.8 GFLOP fwd, x3 fwd+bwd) against the measured
            dense bf16 peak.
  config 5: 64 config-2 designs (seeds 0..63) stepped one after the other on ONE GPU (the
            denominator of the data-parallel scaling runs): designs/s.

  config 1: the reference's CPU-runnable inference case (~5k cells, 64x64 image): GPU vs the CPU oracle.

Prints one JSON object per config.  CUDA events on the launching stream, 3 warm-ups.
Usage: python profiles/bench_configs.py [c3] [c4] [c5] [--c5-designs N]
"""
import importlib
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "multimodal-fusion-based-pre-routing-timing-prediction-_b200"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, PKG))
importlib.import_module(PKG)
import tm_engine  # noqa: E402
import tm_lib  # noqa: E402
import tm_ops  # noqa: E402
import tm_synth  # noqa: E402
import tm_unet  # noqa: E402

DEV = torch.device("cuda", 0)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    j = json.load(open(p)) if os.path.isfile(p) else {}
    return float(j.get("hbm_gbs", 6544.0)), float(j.get("bf16_tflops", 1639.0))


def timed(fn, reps=5, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def config1(cpu=True):
    """BASELINE config 1 (the reference's own CPU-runnable case): inference on a ~5k-cell design with a 64x64 image.
    GPU: DesignStep.forward (schedule cached), CUDA events.  CPU: the oracle port, no_grad, all host threads."""
    import numpy as np
    d = tm_synth.make_design(seed=0, **tm_synth.CONFIGS["c1"])
    model, cnn = tm_engine.build_models(d.map_size, seed=0, device=DEV)
    batch = tm_engine.DesignBatch.from_synth(d, DEV)
    step = tm_engine.DesignStep(model, cnn)
    with torch.no_grad():
        t_gpu = timed(lambda: step.forward(batch), reps=10)
        pred = step.forward(batch).cpu()
    out = {"config": "c1: inference, %d pins, %d levels, %dx%d image, %d endpoints" % (d.n, d.num_levels, 2 * d.map_size, 2 * d.map_size, d.endpoints.size),
           "gpu_ms": t_gpu, "gpu_designs_per_s": 1e3 / t_gpu}
    if cpu:
        from oracle import levelize, restate
        torch.set_num_threads(os.cpu_count() or 1)
        t = torch.from_numpy
        ni, ns = levelize.in_csr(d.n, d.net_src, d.net_dst)
        ci, cs = levelize.in_csr(d.n, d.cell_src, d.cell_dst)
        od = dict(n=d.n, levels=[t(x.astype(np.int64)) for x in d.level_lists()],
                  net_csr=(t(ni).long(), t(ns).long()), cell_csr=(t(ci).long(), t(cs).long()),
                  cell_feat=t(d.cell_feat), net_feat=t(d.net_feat), image=t(d.image), endpoints=t(d.endpoints),
                  endpoint_level=t(d.level[d.endpoints].astype(np.int64)), mask_indptr=t(d.mask_indptr).long(),
                  mask_cols=t(d.mask_cols).long(), arrival_time=t(d.arrival_time))
        sd_m = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        sd_c = {k: v.detach().cpu() for k, v in cnn.state_dict().items()}
        with torch.no_grad():
            ref = restate.design_step(sd_m, sd_c, od, with_grad=False)
            best = 1e9
            for _ in range(3):
                t0 = time.perf_counter()
                restate.design_step(sd_m, sd_c, od, with_grad=False)
                best = min(best, time.perf_counter() - t0)
        err = float((pred - ref["pred"]).abs().max() / ref["pred"].abs().max())
        out.update({"cpu_oracle_ms(best of 3, %d threads)" % (os.cpu_count() or 1): best * 1e3,
                    "max_rel_err_vs_oracle": err})
    return out


def config3():
    d = tm_synth.make_design(seed=0, n_endpoints=64, **tm_synth.CONFIGS["c3"])
    import model as M
    from tm_graph import TimingGraph
    torch.manual_seed(3)
    gnn = M.PathConv(out_feat_dim=128, hidden_feat_dim=128, cell_feat_dim=36, net_feat_dim=2).to(DEV)
    t = torch.from_numpy
    g = TimingGraph(d.n, (t(d.net_src), t(d.net_dst)), (t(d.cell_src), t(d.cell_dst)), pis=t(d.pis))
    g.ndata["cell_feat"], g.ndata["net_feat"] = t(d.cell_feat), t(d.net_feat)
    g = g.to(DEV)
    t0 = time.perf_counter()
    sched = g.schedule()
    torch.cuda.synchronize()
    t_sched = time.perf_counter() - t0
    gp = [dict(gnn.named_parameters())[k].detach() for k in tm_ops.GNN_PARAM_NAMES]
    cf, nf = g.ndata["cell_feat"], g.ndata["net_feat"]
    H, saved = tm_ops.gnn_forward(sched, cf, nf, gp, save=True)
    S = torch.rand(sched.n, 128, device=DEV)
    w1t, w2t = tm_ops.transpose(gp[8]), tm_ops.transpose(gp[10])
    nb = tm_lib.ws_bytes("tm_gnn_ws_bytes")
    ws = tm_lib.workspace(nb, DEV)

    def prop():
        tm_lib.call("tm_gnn_forward", sched.struct, 0, sched.num_levels, H, S, w1t, gp[9], w2t, gp[11],
                    saved["A"], saved["LSE"], saved["HID"], ws, nb, tm_lib.stream())
    G = torch.zeros(sched.n, 128, device=DEV)
    GA, GH, GZ = torch.empty_like(saved["A"]), torch.empty_like(saved["HID"]), torch.empty_like(saved["A"])

    def bwd():
        tm_lib.call("tm_gnn_backward", sched.struct, H, G, gp[8], gp[10], saved["A"], saved["LSE"], saved["HID"],
                    GA, GH, GZ, ws, nb, tm_lib.stream())
    tf, tb = timed(prop), timed(bwd)
    t_full_f = timed(lambda: tm_ops.gnn_forward(sched, cf, nf, gp, save=True))
    t_full_b = timed(lambda: tm_ops.gnn_backward(sched, saved, gp, G))
    hbm, _ = peaks()
    bf, bb = sched.algorithmic_bytes_fwd(), sched.algorithmic_bytes_bwd()
    return {"config": "c3: GNN-only propagation, %d pins, %d levels" % (d.n, sched.num_levels),
            "schedule_build_ms(first call, incl. CSR + levelize)": t_sched * 1e3,
            "propagate_fwd_ms": tf, "propagate_bwd_ms": tb, "fwd_GBps": bf / tf / 1e6, "bwd_GBps": bb / tb / 1e6,
            "fwd_frac_of_hbm_peak": bf / tf / 1e6 / hbm, "bwd_frac_of_hbm_peak": bb / tb / 1e6 / hbm, "hbm_peak_GBps": hbm,
            "gnn_fwd_total_ms(with hoisted MLPs)": t_full_f, "gnn_bwd_total_ms(with weight grads)": t_full_b,
            "pins_per_s_fwd_bwd": d.n / ((t_full_f + t_full_b) * 1e-3)}


def config4(batch=32, size=512, modes=("bf16",)):
    import Unet as U
    out = []
    for mode in modes:
        torch.manual_seed(4)
        net = U.UNet("max").train().to(DEV)
        x = torch.rand(batch, 3, size, size, device=DEV)
        old = tm_unet.MATH
        tm_unet.MATH = mode
        try:
            st = {}

            def fwd():
                st["o"], st["s"] = tm_unet.unet_forward(net, x, need_bwd=True, update_stats=False)
            t_f = timed(fwd, reps=3, warm=3)
            g = torch.ones_like(st["o"])

            def both():
                fwd()
                tm_unet.unet_backward(net, st["s"], g)
            t_fb = timed(both, reps=3, warm=2)
        finally:
            tm_unet.MATH = old
        flop_f = 593.8e9 * batch / 32 * (size / 512) ** 2
        _, tpk = peaks()
        out.append({"config": "c4: U-Net batch %d x %dx%d, operands %s" % (batch, size, size, mode),
                    "fwd_ms": t_f, "fwd_bwd_ms": t_fb, "fwd_TFLOPs": flop_f / t_f / 1e9, "fwd_bwd_TFLOPs": 3 * flop_f / t_fb / 1e9,
                    "fwd_frac_of_bf16_peak": flop_f / t_f / 1e9 / tpk, "fwd_bwd_frac_of_bf16_peak": 3 * flop_f / t_fb / 1e9 / tpk,
                    "bf16_peak_TFLOPs": tpk, "images_per_s_fwd_bwd": batch / (t_fb * 1e-3)})
        del net, x, st
        torch.cuda.empty_cache()
    return out


def layoutnet(size=512, cpu=True):
    """LayoutNet (model.py:216-247), the reference's DEFAULT image branch (train.py:70 without --unet): 2 x 512 x 512
    input, forward + backward, device time and algorithmic TFLOP/s (2*R*S*Cin*Cout*Hout*Wout per convolution:
    9x9 2->32 @512, 7x7 32->64 @256, 9x9 64->32 @128, 7x7 32->1 @128 = 21.4 GFLOP forward), with the oracle port on
    the host cores beside it."""
    import model as M
    from oracle import restate
    torch.manual_seed(1)
    net = M.LayoutNet("max").to(DEV).train()
    x = torch.rand(1, 2, size, size, device=DEV)
    sc = (size / 512) ** 2
    flop_f = sc * 2.0 * (81 * 2 * 32 * 512 * 512 + 49 * 32 * 64 * 256 * 256 + 81 * 64 * 32 * 128 * 128 + 49 * 32 * 1 * 128 * 128)
    out = {"config": "LayoutNet %dx%dx2, batch 1 (reference default CNN)" % (size, size), "fwd_GFLOP": flop_f / 1e9}
    res = {}
    for math in ("tf32x3", "bf16"):
        old = tm_unet.MATH
        try:
            tm_unet.MATH = math
            with torch.no_grad():
                t_f = timed(lambda: net(x), reps=5, warm=3)

            def both():
                net.zero_grad()
                net(x).sum().backward()
            t_fb = timed(both, reps=5, warm=3)
        finally:
            tm_unet.MATH = old
        _, tpk = peaks()
        res[math] = {"fwd_ms": t_f, "fwd_bwd_ms": t_fb, "fwd_TFLOPs": flop_f / t_f / 1e9, "fwd_bwd_TFLOPs": 3 * flop_f / t_fb / 1e9,
                     "fwd_bwd_frac_of_bf16_peak": 3 * flop_f / t_fb / 1e9 / tpk}
    out["modes"] = res
    if cpu:
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in net.state_dict().items()}
        xc = x.cpu()

        def cpu_both():
            y = restate.layoutnet_forward(sd, xc, "max")
            torch.autograd.grad(y.sum(), list(sd.values()))
        cpu_both()
        t0 = time.perf_counter()
        for _ in range(3):
            cpu_both()
        out["cpu_port_fwd_bwd_ms"] = (time.perf_counter() - t0) / 3 * 1e3
        out["cpu_threads"] = threads
    return out


def config5(n_designs=64, distinct=8):
    """64 design steps on one GPU.  `distinct` different designs are made resident (structure +
    captured step); the 64 steps cycle through them (the generator is the same for all seeds, so the
    other 56 have the same size distribution)."""
    d0 = tm_synth.make_design(seed=0, **tm_synth.CONFIGS["c2"])
    model, cnn = tm_engine.build_models(d0.map_size, seed=0, device=DEV)
    step = tm_engine.DesignStep(model, cnn)
    preps = []
    for s in range(distinct):
        d = d0 if s == 0 else tm_synth.make_design(seed=s, **tm_synth.CONFIGS["c2"])
        preps.append(step.prepare(tm_engine.HostDesign(d, pin=False), DEV))
    for p in preps:
        p.step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n_designs):
        preps[i % distinct].step()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    return {"config": "c5 denominator: %d config-2 design steps on 1 GPU (%d distinct designs resident)" % (n_designs, distinct),
            "ms_total": ms, "designs_per_s": n_designs / (ms * 1e-3)}


def cpu_baselines():
    """the CPU side of configs 3 and 4 (the oracle port on all host threads): config 3 GNN-only
    forward+backward on the full ~1M-pin graph; config 4 on a reduced batch (2 images), scaled linearly to 32."""
    import numpy as np
    from oracle import levelize, restate
    import model as M
    import Unet as U
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    out = []
    d = tm_synth.make_design(seed=0, n_endpoints=64, **tm_synth.CONFIGS["c3"])
    torch.manual_seed(3)
    gnn = M.PathConv(out_feat_dim=128, hidden_feat_dim=128, cell_feat_dim=36, net_feat_dim=2)
    t = torch.from_numpy
    ni, ns = levelize.in_csr(d.n, d.net_src, d.net_dst)
    ci, cs = levelize.in_csr(d.n, d.cell_src, d.cell_dst)
    levels = [t(x.astype(np.int64)) for x in d.level_lists()]
    P = {"gnn." + k: v.detach().clone().requires_grad_(True) for k, v in gnn.state_dict().items() if "drive" not in k and "attn" not in k}
    best = 1e9
    for _ in range(2):
        t0 = time.perf_counter()
        H = restate.gnn_propagate(P, "gnn", d.n, levels, (t(ni).long(), t(ns).long()), (t(ci).long(), t(cs).long()),
                                  t(d.cell_feat), t(d.net_feat))
        H[t(d.endpoints).long()].square().sum().backward()
        best = min(best, time.perf_counter() - t0)
    out.append({"config": "c3 CPU: GNN-only forward+backward, %d pins" % d.n, "cpu_s": best, "threads": threads,
                "pins_per_s": d.n / best})
    torch.manual_seed(4)
    net = U.UNet("max").train()
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    x = torch.rand(2, 3, 512, 512)
    best = 1e9
    for _ in range(2):
        Pc = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone()) for k, v in sd.items()}
        t0 = time.perf_counter()
        o, _ = restate.unet_forward(Pc, x, "max")
        o.sum().backward()
        best = min(best, time.perf_counter() - t0)
    out.append({"config": "c4 CPU: U-Net forward+backward, batch 2 x 512x512 fp32 (reduced batch)", "cpu_s": best, "threads": threads,
                "images_per_s": 2 / best, "batch32_s(scaled x16)": best * 16})
    return out


def profile_c4(mode="bf16", batch=32, size=512):
    """One forward+backward between cudaProfilerStart/Stop (for `ncu --profile-from-start off`)."""
    import Unet as U
    torch.manual_seed(4)
    net = U.UNet("max").train().to(DEV)
    x = torch.rand(batch, 3, size, size, device=DEV)
    tm_unet.MATH = mode
    for i in range(2):
        if i == 1:
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
        o, s = tm_unet.unet_forward(net, x, need_bwd=True, update_stats=False)
        tm_unet.unet_backward(net, s, torch.ones_like(o))
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    if "--profile-c4" in sys.argv:
        profile_c4(mode=os.environ.get("TM_C4_MODE", "bf16"))
        sys.exit(0)
    want = [a for a in sys.argv[1:] if not a.startswith("--")] or ["c1", "c3", "c4", "c5"]
    if "c1" in want:
        print(json.dumps(config1()), flush=True)
    if "c3" in want:
        print(json.dumps(config3()), flush=True)
    if "c4" in want:
        for r in config4(modes=("bf16", "tf32x3") if "--c4-all" in sys.argv else ("bf16",)):
            print(json.dumps(r), flush=True)
    if "layoutnet" in want:
        print(json.dumps(layoutnet()), flush=True)
    if "c5" in want:
        print(json.dumps(config5()), flush=True)
    if "cpu" in want:
        for r in cpu_baselines():
            print(json.dumps(r), flush=True)
