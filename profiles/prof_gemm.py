"""Stand-alone launch of the hoisted-MLP layer-2 GEMM (M=229k, N=128, K=256) and its weight gradient,
for `ncu --set full` (run 3 warm launches, profile between cudaProfilerStart/Stop)."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
importlib.import_module("multimodal-fusion-based-pre-routing-timing-prediction-_b200")
import tm_ops
M, N, K = 229819, 128, 256
torch.manual_seed(0)
A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda"); b = torch.randn(N, device="cuda")
C = torch.empty(M, N, device="cuda"); G = torch.randn(M, N, device="cuda"); dW = torch.empty(N, K, device="cuda")
def run():
    tm_ops.gemm_nn(M, N, K, A, K, W, K, C, N, bias=b, b_is_nk=True)
    tm_ops.gemm_tn(N, K, M, G, N, A, K, dW, K)
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); tm_ops.gemm_nn(M, N, K, A, K, W, K, C, N, bias=b, b_is_nk=True); e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1); print(f"gemm_nn {t*1e3:.1f} us  {2*M*N*K/t/1e9:.1f} TFLOP/s  {(M*K+M*N)*4/t/1e6:.0f} GB/s")
e0.record(); tm_ops.gemm_tn(N, K, M, G, N, A, K, dW, K); e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1); print(f"gemm_tn {t*1e3:.1f} us  {2*M*N*K/t/1e9:.1f} TFLOP/s  {(M*K+M*N)*4/t/1e6:.0f} GB/s")
torch.cuda.profiler.start(); run(); torch.cuda.synchronize(); torch.cuda.profiler.stop()
