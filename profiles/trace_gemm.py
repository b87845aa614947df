"""TM_TC_TRACE=1 python profiles/trace_gemm.py : clock64 timeline of CTA 0 of one big tcgen05 GEMM."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
importlib.import_module("multimodal-fusion-based-pre-routing-timing-prediction-_b200")
import tm_ops
M, N, K = 229819, 128, 256
A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda"); b = torch.randn(N, device="cuda")
C = torch.empty(M, N, device="cuda")
tm_ops.gemm_nn(M, N, K, A, K, W, K, C, N, bias=b, b_is_nk=True)
torch.cuda.synchronize()
