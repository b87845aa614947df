"""One-off edge-size sweep of the fused self-term MLP kernels (tm_selfmlp.cu) against fp64: M = 1 .. a few tiles,
with and without row lists.  Prints the worst relative error per kernel; exits non-zero on a miss."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, ROOT + "/multimodal-fusion-based-pre-routing-timing-prediction-_b200"):
    sys.path.insert(0, p)
importlib.import_module("multimodal-fusion-based-pre-routing-timing-prediction-_b200")
import tm_lib as lib

DEV = "cuda"
worst = {}


def rel(a, b):
    b = b.double()
    return float((a.double() - b).abs().max() / max(float(b.abs().max()), 1e-30))


def note(k, e):
    worst[k] = max(worst.get(k, 0.0), e)


for M in (1, 2, 7, 31, 63, 64, 65, 127, 128, 129, 255, 300, 1025, 9473, 18945):
    for gather in (False, True):
        torch.manual_seed(M * 2 + gather)
        n = M + 13
        X2 = torch.randn(n, 2, device=DEV); X36 = torch.randn(n, 36, device=DEV)
        W1 = torch.randn(256, 2, device=DEV) * 0.5; W136 = torch.randn(256, 36, device=DEV) * 0.3
        b1 = torch.randn(256, device=DEV) * 0.3
        W2 = torch.randn(128, 256, device=DEV) * 0.1; b2 = torch.randn(128, device=DEV) * 0.1
        G = torch.randn(n, 128, device=DEV) * 1e-3
        H = torch.randn(n, 256, device=DEV).relu_()
        r = torch.randperm(n, device=DEV)[:M].int().contiguous() if gather else None
        sel = (lambda t: t[r.long()]) if gather else (lambda t: t[:M])
        st = lib.stream()
        # forward
        out = torch.zeros(n, 128, device=DEV)
        nb = lib.ws_bytes("tm_selfmlp_ws_bytes")
        lib.call("tm_selfmlp_gen_forward", M, X2, 2, r, 2, W1, b1, W2, b2, out, 128, r, lib.workspace(nb, DEV), nb, st)
        ref = (sel(X2).double() @ W1.double().t() + b1.double()).relu() @ W2.double().t() + b2.double()
        note("forward", rel(sel(out), ref))
        # wgrad2
        db = torch.empty(128, device=DEV); gmax = torch.empty(1, device=DEV)
        nbc = lib.ws_bytes("tm_colsum_ws", M, 128)
        lib.call("tm_colsum_absmax", M, 128, G, 128, r, db, gmax, lib.workspace(nbc, DEV), nbc, st)
        nb = lib.ws_bytes("tm_selfmlp_wgrad2_ws_bytes"); dw = torch.empty(128, 256, device=DEV)
        lib.call("tm_selfmlp_gen_wgrad2", M, G, 128, r, X2, 2, r, 2, W1, b1, gmax, dw, lib.workspace(nb, DEV), nb, st)
        h = (sel(X2).double() @ W1.double().t() + b1.double()).relu()
        note("wgrad2", rel(dw, sel(G).double().t() @ h)); note("colsum", rel(db, sel(G).double().sum(0)))
        # bwd1
        nb = lib.ws_bytes("tm_selfmlp_bwd1_ws_bytes"); dW1 = torch.empty(256, 2, device=DEV); db1 = torch.empty(256, device=DEV)
        lib.call("tm_selfmlp_gen_bwd1", M, G, 128, r, X2, 2, r, 2, W1, b1, W2, dW1, db1, lib.workspace(nb, DEV), nb, st)
        pre = sel(X2).double() @ W1.double().t() + b1.double()
        dh = (sel(G).double() @ W2.double()) * (pre > 0)
        note("bwd1.db1", rel(db1, dh.sum(0))); note("bwd1.dW1", rel(dW1, dh.t() @ sel(X2).double()))
        # rows_dh
        nb = lib.ws_bytes("tm_selfmlp_rows_dh_ws_bytes"); DH = torch.zeros(n, 256, device=DEV)
        lib.call("tm_selfmlp_rows_dh", M, G, 128, r, W2, H, 256, r, DH, 256, lib.workspace(nb, DEV), nb, st)
        note("rows_dh", rel(sel(DH), (sel(G).double() @ W2.double()) * (sel(H) > 0)))
        # lin1
        nb = lib.ws_bytes("tm_selfmlp_lin1_ws_bytes"); HH = torch.zeros(M, 256, device=DEV)
        rmx = torch.empty(M, device=DEV)
        lib.call("tm_selfmlp_lin1_relu", M, X36, 36, r, 36, W136, b1, HH, 256, rmx, lib.workspace(nb, DEV), nb, st)
        note("lin1", rel(HH, (sel(X36).double() @ W136.double().t() + b1.double()).relu()))
        nb = lib.ws_bytes("tm_selfmlp_ws_bytes"); o2 = torch.zeros(n, 128, device=DEV)
        lib.call("tm_selfmlp_rows_forward", M, HH, 256, None, rmx, W2, b2, o2, 128, r, lib.workspace(nb, DEV), nb, st)
        note("rows_forward", rel(sel(o2), HH.double() @ W2.double().t() + b2.double()))
torch.cuda.synchronize()
print({k: f"{v:.2e}" for k, v in worst.items()})
bad = {k: v for k, v in worst.items() if v > (2e-3 if k.startswith("bwd1") else 1e-4)}
sys.exit(1 if bad else 0)
