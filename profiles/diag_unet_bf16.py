"""bf16 image branch (TMA-fed tcgen05 convolutions) against the bf16-rounded oracle and the fp32 oracle:
per-tensor error statistics of the output and of every parameter gradient.  python profiles/diag_unet_bf16.py [H] [B]"""
import importlib, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "multimodal-fusion-based-pre-routing-timing-prediction-_b200"
for p in (ROOT, os.path.join(ROOT, PKG)):
    sys.path.insert(0, p)
importlib.import_module(PKG)
import Unet as U
import tm_unet
from oracle import restate

H = int(sys.argv[1]) if len(sys.argv) > 1 else 512
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.manual_seed(4)
net = U.UNet("max").train()
sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
x = torch.rand(B, 3, H, H)
names = [k for k, _ in net.named_parameters()]


def oracle(rounding):
    P = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone()) for k, v in sd.items()}
    out, _ = restate.unet_forward(P, x, "max", rounding=rounding)
    return out, P


torch.manual_seed(9)
o32, P32 = oracle(None)
g = torch.randn_like(o32)
g32 = torch.autograd.grad(o32, [P32[k] for k in names], g)
ob, Pb = oracle("bf16")
gb = torch.autograd.grad(ob, [Pb[k] for k in names], g)
net = net.to("cuda")
tm_unet.MATH = "bf16"
out = net(x.cuda())
out.backward(g.cuda())


def stats(a, r):
    a, r = a.detach().double().cpu().reshape(-1), r.detach().double().reshape(-1)
    scale = float(r.abs().max())
    err = (a - r).abs()
    return dict(max_over_scale=float(err.max() / scale), rel_l2=float((a - r).norm() / r.norm()),
                frac_gt_2e2=float((err > 2e-2 * r.abs() + 2e-3 * scale).double().mean()), n=a.numel())


print(json.dumps(dict(t="out", vs_bf16=stats(out, ob), vs_fp32=stats(out, o32), oracle_bf16_vs_fp32=stats(ob, o32))))
for k, rb, r32 in zip(names, gb, g32):
    a = dict(net.named_parameters())[k].grad
    print(json.dumps(dict(t=k, vs_bf16=stats(a, rb), vs_fp32_rel_l2=stats(a, r32)["rel_l2"])))
