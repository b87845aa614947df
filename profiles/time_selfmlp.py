"""Timing probe of the fused net-pin MLP kernels (tm_selfmlp.cu) at config-2 size."""
import importlib, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, ROOT + "/multimodal-fusion-based-pre-routing-timing-prediction-_b200", ROOT + "/profiles"):
    sys.path.insert(0, p)
importlib.import_module("multimodal-fusion-based-pre-routing-timing-prediction-_b200")
import tm_lib
from dev_gnn_persist import timeit
M = 229819
X = torch.randn(M, 2, device="cuda"); W1 = torch.randn(256, 2, device="cuda"); b1 = torch.randn(256, device="cuda")
W2 = torch.randn(128, 256, device="cuda") * 0.1; b2 = torch.randn(128, device="cuda")
out = torch.empty(M + 1000, 128, device="cuda"); rows = torch.sort(torch.randperm(M + 1000, device="cuda")[:M]).values.int()
nb = tm_lib.ws_bytes("tm_selfmlp_ws_bytes"); ws = tm_lib.workspace(nb, "cuda")
t = timeit(lambda: tm_lib.call("tm_selfmlp_gen_forward", M, X, 2, None, 2, W1, b1, W2, b2, out, 128, rows, ws, nb, tm_lib.stream()))
G = torch.randn(M + 1000, 128, device="cuda") * 1e-3
gmax = G.abs().max().reshape(1)
nb2 = tm_lib.ws_bytes("tm_selfmlp_wgrad2_ws_bytes"); ws2 = tm_lib.workspace(nb2, "cuda"); dw = torch.empty(128, 256, device="cuda")
t2 = timeit(lambda: tm_lib.call("tm_selfmlp_gen_wgrad2", M, G, 128, rows, X, 2, None, 2, W1, b1, gmax, dw, ws2, nb2, tm_lib.stream()))
print(os.environ.get("TM_SELFMLP_DEBUG", "0"), "forward", round(t * 1e3, 1), "us; wgrad2", round(t2 * 1e3, 1), "us")
