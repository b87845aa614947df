"""Layer-by-layer comparison of the bf16 image branch with the bf16-rounded oracle (forward intermediates)."""
import importlib, json, os, sys
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "multimodal-fusion-based-pre-routing-timing-prediction-_b200"
for p in (ROOT, os.path.join(ROOT, PKG)):
    sys.path.insert(0, p)
importlib.import_module(PKG)
import Unet as U
import tm_unet
from oracle import restate

H = int(sys.argv[1]) if len(sys.argv) > 1 else 128
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.manual_seed(4)
net = U.UNet("max").train()
sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
x = torch.rand(B, 3, H, H)
bf = restate._bf16_rn


def rel(a, r):
    a, r = a.double().cpu(), r.double()
    return float((a - r).abs().max() / r.abs().max())


def nhwc(t, C, h):
    return t.reshape(B, h, h, C).permute(0, 3, 1, 2)


net_d = U.UNet("max").train()
net_d.load_state_dict(sd)
net_d = net_d.cuda()
tm_unet.MATH = "bf16"
tm_unet._enter(net_d)
with torch.no_grad():
    out, st = tm_unet.unet_forward(net_d, x.cuda(), need_bwd=True, update_stats=False)
torch.cuda.synchronize()
# oracle, block enc0 by hand
p = "inc.double_conv"
r1 = F.conv2d(bf(x), bf(sd[p + ".0.weight"]), None, padding=1)
e = st["enc0"]
print(json.dumps(dict(tma=bool(e.get("tma")), r1=rel(nhwc(e["r1"], 16, H), r1))))
a1 = F.relu(F.batch_norm(r1, None, None, sd[p + ".1.weight"], sd[p + ".1.bias"], training=True))
if e.get("a1b") is not None:
    print(json.dumps(dict(a1b_vs_bf_a1=rel(nhwc(e["a1b"].float(), 16, H), bf(a1)))))
    nd = (nhwc(e["a1b"].float(), 16, H).cpu() != bf(a1)).float().mean()
    print(json.dumps(dict(a1b_frac_diff=float(nd))))
r2 = F.conv2d(bf(a1), bf(sd[p + ".3.weight"]), None, padding=1)
print(json.dumps(dict(r2=rel(nhwc(e["r2"], 16, H), r2))))
# product's r2 from the PRODUCT's a1b (isolates the second convolution)
if e.get("a1b") is not None:
    r2p = F.conv2d(nhwc(e["a1b"].float(), 16, H).cpu(), bf(sd[p + ".3.weight"]), None, padding=1)
    print(json.dumps(dict(r2_given_product_a1b=rel(nhwc(e["r2"], 16, H), r2p))))


# ---- the whole chain, block by block (oracle intermediates captured by re-running its pieces)
def l2(a, r):
    a, r = a.double().cpu(), r.double()
    return float((a - r).norm() / r.norm())


def dc(pfx, xin, C):
    r1 = F.conv2d(bf(xin), bf(sd[pfx + ".0.weight"]), None, padding=1)
    a1 = F.relu(F.batch_norm(r1, None, None, sd[pfx + ".1.weight"], sd[pfx + ".1.bias"], training=True))
    r2 = F.conv2d(bf(a1), bf(sd[pfx + ".3.weight"]), None, padding=1)
    o = F.relu(F.batch_norm(r2, None, None, sd[pfx + ".4.weight"], sd[pfx + ".4.bias"], training=True))
    return r1, r2, o


chans = [16, 32, 64, 128]
names = ["inc.double_conv", "down1.maxpool_conv.1.double_conv", "down2.maxpool_conv.1.double_conv", "down3.maxpool_conv.1.double_conv"]
cur = x
skips = []
for i in range(4):
    if i > 0:
        cur = F.max_pool2d(cur, 2)
    h = H >> i
    r1, r2, o = dc(names[i], cur, chans[i])
    e = st[f"enc{i}"]
    po = e["out"][:, :chans[i]] if e["ldo"] != chans[i] else e["out"]
    print(json.dumps(dict(block=f"enc{i}", tma=bool(e.get("tma")), r1=[rel(nhwc(e["r1"], chans[i], h), r1), l2(nhwc(e["r1"], chans[i], h), r1)],
                          r2=[rel(nhwc(e["r2"], chans[i], h), r2), l2(nhwc(e["r2"], chans[i], h), r2)],
                          out=[rel(nhwc(po.contiguous(), chans[i], h), o), l2(nhwc(po.contiguous(), chans[i], h), o)])))
    skips.append(o)
    cur = o
y = cur
for j, nm in enumerate(["up1", "up2", "up3"]):
    i = 2 - j
    h = H >> i
    up = F.conv_transpose2d(bf(y), bf(sd[nm + ".up.weight"]), sd[nm + ".up.bias"], stride=2)
    catb = st["cat"][i]
    pu = nhwc(catb[:, chans[i]:].contiguous(), chans[i], h)
    r1, r2, o = dc(nm + ".conv.double_conv", torch.cat([skips[i], up], 1), chans[i])
    e = st[f"dec{j}"]
    print(json.dumps(dict(block=f"dec{j}", tma=bool(e.get("tma")), up=[rel(pu, up), l2(pu, up)],
                          r1=[rel(nhwc(e["r1"], chans[i], h), r1), l2(nhwc(e["r1"], chans[i], h), r1)],
                          r2=[rel(nhwc(e["r2"], chans[i], h), r2), l2(nhwc(e["r2"], chans[i], h), r2)],
                          out=[rel(nhwc(e["out"], chans[i], h), o), l2(nhwc(e["out"], chans[i], h), o)])))
    y = o
