"""Effective bandwidth of the streamed dense contractions of the netlist branch by arithmetic mode (pipeline depth
differs: tf32x3 keeps 2 k-blocks in flight, tf32 5).  Usage: python profiles/diag_gemm_modes.py -> JSON lines."""
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
import tm_ops  # noqa: E402
from dev_gnn_persist import timeit  # noqa: E402

dev = torch.device("cuda")
torch.manual_seed(0)
for (M, N, K) in ((229819, 128, 256), (102000, 256, 128), (102000, 256, 36)):
    A = torch.randn(M, K, device=dev)
    W = torch.randn(N, K, device=dev) * 0.05
    C = torch.empty(M, N, device=dev)
    ref = (A[:4096].double() @ W.double().t()).float()
    for math in ("tf32x3", "tf32", "tc6", "tc3", "bf16"):
        t = timeit(lambda: tm_ops.gemm_nn(M, N, K, A, K, W, K, C, N, b_is_nk=True, math=math))
        err = float((C[:4096] - ref).abs().max() / ref.abs().max())
        mb = (M * K + M * N) * 4 / 1e6
        print(json.dumps(dict(kind="nn", M=M, N=N, K=K, math=math, ms=round(t, 4), GBps=round(mb / t, 1), err=err)), flush=True)
# weight gradient: C[M,N] = A[R,M]^T B[R,N]
for (M, N, R) in ((128, 256, 100000), (256, 128, 100000), (256, 36, 102000)):
    A = torch.randn(R, M, device=dev)
    B = torch.randn(R, N, device=dev)
    C = torch.empty(M, N, device=dev)
    for math in ("tf32x3", "tf32", "tc6", "tc3", "bf16"):
        old = tm_ops.MATH
        tm_ops.MATH = math
        try:
            t = timeit(lambda: tm_ops.gemm_tn(M, N, R, A, M, B, N, C, N))
        finally:
            tm_ops.MATH = old
        mb = (R * M + R * N) * 4 / 1e6
        print(json.dumps(dict(kind="tn", M=M, N=N, R=R, math=math, ms=round(t, 4), GBps=round(mb / t, 1))), flush=True)
