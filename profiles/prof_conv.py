"""One TMA convolution layer of the bf16 U-Net in isolation (for `ncu --set full --profile-from-start off`) and its
device time / algorithmic TFLOP/s / HBM GB/s.  usage: python profiles/prof_conv.py B H W Cin Cout [reps]"""
import importlib, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "multimodal-fusion-based-pre-routing-timing-prediction-_b200"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, PKG))
importlib.import_module(PKG)
import tm_lib as L  # noqa: E402

B, H, W, Cin, Cout = [int(v) for v in sys.argv[1:6]]
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 10
dev = "cuda"
torch.manual_seed(0)
npix = B * H * W
xb = torch.randn(npix, Cin, device=dev).bfloat16()
gb = torch.randn(npix, Cout, device=dev).bfloat16()
w = torch.randn(Cout, Cin, 3, 3, device=dev) * 0.1
P = L.ws_bytes("tm_conv3x3_bf16_pack", W, Cin, Cout)
wq = torch.empty(9, P * Cout, P * Cin, dtype=torch.bfloat16, device=dev)
L.call("tm_conv3x3_pack_bf16", Cout, Cin, w, wq, P, Cin, 0, L.stream())
y = torch.empty(npix, Cout, device=dev)
dw = torch.empty(Cout, Cin, 3, 3, device=dev)
nb = L.ws_bytes("tm_conv3x3_bf16_wgrad_ws", B, H, W, Cin, Cout)
ws = L.workspace(nb, dev)
err = L.err_flag(dev)


def fprop():
    L.call("tm_conv3x3_bf16", B, H, W, Cin, Cout, P, xb, wq, None, y, Cout, 0, None, err, L.stream())


def wgrad():
    L.call("tm_conv3x3_bf16_wgrad", B, H, W, Cin, Cin, Cout, xb, gb, dw, ws, nb, err, L.stream())


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


tf, tw = timed(fprop), timed(wgrad)
torch.cuda.profiler.start()
fprop()
wgrad()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
flop = 2.0 * 9 * Cin * Cout * npix
print(json.dumps({"layer": f"conv3x3 {Cin}->{Cout}, {B}x{H}x{W}, pixels/row P={P}", "fprop_ms": tf, "wgrad_ms": tw,
                  "fprop_TFLOPs": flop / tf / 1e9, "wgrad_TFLOPs": flop / tw / 1e9,
                  "fprop_hbm_GBps(bf16 in + fp32 out)": npix * (2 * Cin + 4 * Cout) / tf / 1e6,
                  "wgrad_hbm_GBps(bf16 x + bf16 dy)": npix * 2 * (Cin + Cout) / tw / 1e6, "err": int(err.item())}))
