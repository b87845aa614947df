"""Bottleneck bisection of the streamed tcgen05 GEMM (needs `make INSTRUMENT=1`): TM_TC_DEBUG bits switch off the
epilogue stores (1), the global loads (2), the lo-plane staging (4), the MMAs (8).  One process per setting."""
import importlib, json, os, subprocess, sys
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (ROOT, os.path.join(ROOT, "multimodal-fusion-based-pre-routing-timing-prediction-_b200"), os.path.join(ROOT, "profiles")):
        sys.path.insert(0, p)
    importlib.import_module("multimodal-fusion-based-pre-routing-timing-prediction-_b200")
    import tm_ops
    from dev_gnn_persist import timeit
    M, N, K = [int(x) for x in sys.argv[2:5]]
    math = sys.argv[5]
    A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda"); C = torch.empty(M, N, device="cuda")
    t = timeit(lambda: tm_ops.gemm_nn(M, N, K, A, K, W, K, C, N, b_is_nk=True, math=math))
    print(json.dumps(dict(M=M, N=N, K=K, math=math, dbg=int(os.environ.get("TM_TC_DEBUG", "0")), ms=round(t, 4))), flush=True)
else:
    for shape in (("229819", "128", "256"), ("102000", "256", "128")):
        for math in ("tf32x3", "tf32"):
            for dbg in (0, 1, 2, 4, 8, 3, 6, 7, 9, 15):
                env = dict(os.environ, TM_TC_DEBUG=str(dbg))
                subprocess.run([sys.executable, __file__, "child", *shape, math], env=env)
