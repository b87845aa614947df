"""Image branch on libtm_b200: U-Net (src/Unet.py:85-119) and LayoutNet (src/model.py:216-247).

Activations are NHWC fp32 on the device; skip connections and transposed-conv outputs are written
straight into the two halves of a shared concat buffer (``torch.cat`` of Unet.py:67 costs
nothing).  BatchNorm always uses batch statistics -- the reference never calls ``.eval()``
(train.py:436-437) -- and updates the running statistics exactly like ``nn.BatchNorm2d``
(momentum 0.1, unbiased variance).  Every contraction is a ``tm_conv2d_nhwc`` /
``tm_convt2x2_nhwc`` implicit GEMM; nothing here calls cuDNN or ATen convolution.
"""
import os

import torch

import tm_lib
from tm_lib import call, stream

RELU = 2


def _empty(*shape, dev):
    return torch.empty(*shape, dtype=torch.float32, device=dev)


class _Ws:
    """One growing scratch buffer shared by the split-reduction kernels of a pass."""

    def __init__(self, dev):
        self.dev, self.buf = dev, None

    def get(self, nbytes):
        if self.buf is None or self.buf.numel() < nbytes:
            self.buf = tm_lib.workspace(nbytes, self.dev)
        return self.buf


# arithmetic of the convolutions: "tc3" (split-bf16 x3 on tcgen05, fp32-class), "bf16" (tcgen05,
# plain bf16 operands) or "fp32" (CUDA-core implicit GEMM).  Per network: net.math overrides.
MATH = None            # None -> follow tm_ops.MATH
_CUR = None            # math mode of the network currently being run (net.math)


def _prec(math):
    import tm_ops
    return tm_ops._precision(math or _CUR or MATH or tm_ops.MATH)


def _enter(net):
    global _CUR
    _CUR = getattr(net, "math", None)


USE_TMA = os.environ.get("TM_CONV_TMA", "1") != "0"   # bf16 mode: TMA-fed 3x3 convolutions on bf16 activations


def _cpad(c):
    return max(16, (c + 15) // 16 * 16)


def _tma_ok(B, H, W, cin, cout):
    """bf16 mode and a shape the TMA kernels take (H, W powers of two; channels multiples of 16)."""
    return (USE_TMA and _prec(None) == 0 and cout <= 128 and _cpad(cin) <= 128
            and tm_lib.ws_bytes("tm_conv3x3_bf16_supported", B, H, W, _cpad(cin), cout) == 1
            and (cin % 16 != 0 or tm_lib.ws_bytes("tm_conv3x3_bf16_supported", B, H, W, cout, cin) == 1))


def _to_bf16(x, ldx, npix, C):
    """fp32 rows (stride ldx) -> compact bf16 [npix][pad16(C)]: the TMA operand of the next convolution."""
    out = torch.empty(npix, _cpad(C), dtype=torch.bfloat16, device=x.device)
    call("tm_to_bf16_rows", npix, C, x, ldx, out, _cpad(C), stream())
    return out


def _pack_bf16(w, W, need_dgrad):
    """-> ((forward operand, P), (data-gradient operand, P) or None) for images W wide."""
    cout, cin = w.shape[0], w.shape[1]
    dev = w.device
    wc = w.detach().float().contiguous()

    def one(nc, kc, dgrad):
        kp = _cpad(kc)
        P = tm_lib.ws_bytes("tm_conv3x3_bf16_pack", W, kp, nc)
        q = torch.empty(9, P * nc, P * kp, dtype=torch.bfloat16, device=dev)
        call("tm_conv3x3_pack_bf16", cout, cin, wc, q, P, kp, dgrad, stream())
        return q, P
    return one(cout, cin, 0), (one(cin, cout, 1) if need_dgrad else None)


FUSE_BN_STATS = os.environ.get("TM_FUSE_BN_STATS", "1") != "0"   # BN statistics in the convolution's epilogue


def _conv_tma(xb, B, H, W, cinp, cout, wq, y, ldy, want_stats=False):
    """-> (stats buffer, number of partials) when the epilogue also took the batch-norm statistics, else None."""
    q, P = wq
    stats = None
    if want_stats and FUSE_BN_STATS and P * cout <= 128:
        nb = tm_lib.ws_bytes("tm_conv3x3_bf16_stats_bytes", cout, P)
        stats = (torch.empty(nb, dtype=torch.uint8, device=y.device), nb // (16 * cout))
    call("tm_conv3x3_bf16", B, H, W, cinp, cout, P, xb, q, None, y, ldy, 0, stats[0] if stats else None,
         tm_lib.err_flag(y.device), stream())
    return stats


def _wgrad_tma(ws, xb, dyb, B, H, W, cin, cout):
    dev = dyb.device
    dw = _empty(cout, cin, 3, 3, dev=dev)
    nb = tm_lib.ws_bytes("tm_conv3x3_bf16_wgrad_ws", B, H, W, _cpad(cin), cout)
    call("tm_conv3x3_bf16_wgrad", B, H, W, _cpad(cin), cin, cout, xb, dyb, dw, ws.get(nb), nb, tm_lib.err_flag(dev), stream())
    return dw


def _conv(x, ldx, B, H, W, cin, cout, k, wf, bias, y, ldy, flags=0, math=None):
    prec = _prec(math)
    if prec is None:
        call("tm_conv2d_nhwc", B, H, W, cin, cout, k, x, ldx, wf, bias, y, ldy, flags, stream())
    else:
        nb = tm_lib.ws_bytes("tm_tc_conv2d_splitk_ws", B, H, W, cin, cout, k, prec)   # 0: the layer fills the GPU unsplit
        call("tm_tc_conv2d_nhwc_splitk", B, H, W, cin, cout, k, x, ldx, wf, bias, y, ldy, flags, prec,
             tm_lib.workspace(nb, y.device) if nb else None, nb, tm_lib.err_flag(y.device), stream())


def _conv_wgrad(ws, x, ldx, dy, lddy, B, H, W, cin, cout, k, want_bias, math=None):
    dev = dy.device
    dwf = _empty(k * k * cin, cout, dev=dev)
    dbias = _empty(cout, dev=dev) if want_bias else None
    prec = _prec(math)
    if prec is None:
        nb = tm_lib.ws_bytes("tm_conv2d_wgrad_ws", B, H, W, cin, cout, k)
        call("tm_conv2d_wgrad_nhwc", B, H, W, cin, cout, k, x, ldx, dy, lddy, dwf, dbias, ws.get(nb), nb, stream())
    else:
        nb = tm_lib.ws_bytes("tm_tc_conv2d_wgrad_ws", B, H, W, cin, cout, k)
        call("tm_tc_conv2d_wgrad_nhwc", B, H, W, cin, cout, k, x, ldx, dy, lddy, dwf, prec, ws.get(nb), nb,
             tm_lib.err_flag(dev), stream())
        if want_bias:
            nbc = tm_lib.ws_bytes("tm_colsum_ws", B * H * W, cout)
            call("tm_colsum", B * H * W, cout, dy, lddy, None, dbias, 0, tm_lib.workspace(nbc, dev), nbc, stream())
    dw = _empty(cout, cin, k, k, dev=dev)
    call("tm_conv_unpack_wgrad", cout, cin, k, dwf, dw, stream())
    return dw, dbias


def _pack(w, need_bwd=True):
    cout, cin, k, _ = w.shape
    dev = w.device
    w = w.detach().float().contiguous()
    wf = _empty(k * k * cin, cout, dev=dev)
    wb = _empty(k * k * cout, cin, dev=dev) if need_bwd else None
    call("tm_conv_pack_weight", cout, cin, k, w, wf, wb, stream())
    return wf, wb


EVAL = False          # set by unet_forward: nn.Module.eval() -> BatchNorm uses its running statistics


def _bn_relu_fwd(ws, x, ldx, npix, C, bn, y, ldy, update_stats, yb=None, stats=None):
    dev = x.device
    mean, invstd = _empty(C, dev=dev), _empty(C, dev=dev)
    if EVAL:
        if bn.running_mean is None:
            raise NotImplementedError("eval-mode BatchNorm without running statistics (track_running_stats=False)")
        call("tm_bn_relu_eval", npix, C, x, ldx, bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var,
             float(bn.eps), y, ldy, mean, invstd, yb, stream())
        return mean, invstd
    nb = tm_lib.ws_bytes("tm_bn_ws", npix, C)
    rm = bn.running_mean if update_stats else None
    rv = bn.running_var if update_stats else None
    call("tm_bn_relu_forward", npix, C, x, ldx, bn.weight.detach(), bn.bias.detach(), rm, rv,
         float(bn.momentum), float(bn.eps), y, ldy, mean, invstd, yb, stats[0] if stats else None,
         stats[1] if stats else 0, ws.get(nb), nb, stream())
    if update_stats and bn.num_batches_tracked is not None:
        bn.num_batches_tracked += 1
    return mean, invstd


def _bn_relu_bwd(ws, x, ldx, dy, lddy, npix, C, gamma, beta, mean, invstd, dx, lddx, dxb=None):
    """dx (fp32, may be None) and / or dxb (compact bf16) receive the gradient w.r.t. the BN input.  The ReLU mask
    is rebuilt from x and the saved statistics (bit-identical to the forward), so the activation is not read."""
    dev = x.device
    dg, db = _empty(C, dev=dev), _empty(C, dev=dev)
    nb = tm_lib.ws_bytes("tm_bn_ws", npix, C)
    call("tm_bn_relu_backward", npix, C, x, ldx, None, 0, dy, lddy, gamma, beta, mean, invstd, dx, lddx, dg, db, dxb,
         ws.get(nb), nb, stream())
    return dg, db


# --------------------------------------------------------------------------------------------
# U-Net
# --------------------------------------------------------------------------------------------
def _double_conv_fwd(ws, st, name, mods, x, ldx, B, H, W, cin, cout, out, ldo, update_stats, need_bwd):
    """conv3x3 -> BN -> ReLU -> conv3x3 -> BN -> ReLU (Unet.py:15-22); final activation -> out."""
    conv1, bn1, conv2, bn2 = mods
    dev = x.device
    npix = B * H * W
    cmid = conv1.weight.shape[0]
    r1, r2 = _empty(npix, cmid, dev=dev), _empty(npix, cout, dev=dev)
    if _tma_ok(B, H, W, cin, cmid) and _tma_ok(B, H, W, cmid, cout):
        # bf16 mode: TMA-fed tcgen05 convolutions on compact bf16 copies of the activations (kept for wgrad)
        xb = _to_bf16(x, ldx, npix, cin)
        wq1, wd1 = _pack_bf16(conv1.weight, W, need_bwd and cin % 16 == 0)
        wq2, wd2 = _pack_bf16(conv2.weight, W, need_bwd)
        s1 = _conv_tma(xb, B, H, W, _cpad(cin), cmid, wq1, r1, cmid, want_stats=True)
        a1b = torch.empty(npix, cmid, dtype=torch.bfloat16, device=dev)
        # the mid activation exists only as the bf16 operand of conv2 (the backward rebuilds its ReLU mask from r1)
        m1, i1 = _bn_relu_fwd(ws, r1, cmid, npix, cmid, bn1, None, 0, update_stats, yb=a1b, stats=s1)
        s2 = _conv_tma(a1b, B, H, W, cmid, cout, wq2, r2, cout, want_stats=True)
        m2, i2 = _bn_relu_fwd(ws, r2, cout, npix, cout, bn2, out, ldo, update_stats, stats=s2)
        st[name] = dict(tma=True, xb=xb if need_bwd else None, a1b=a1b if need_bwd else None, wd1=wd1, wd2=wd2,
                        x=x, ldx=ldx, B=B, H=H, W=W, cin=cin, cmid=cmid, cout=cout, r1=r1, r2=r2, out=out,
                        ldo=ldo, m1=m1, i1=i1, m2=m2, i2=i2, g1=bn1.weight.detach(), g2=bn2.weight.detach(),
                        be1=bn1.bias.detach(), be2=bn2.bias.detach())
        return
    a1 = _empty(npix, cmid, dev=dev)
    wf1, wb1 = _pack(conv1.weight, need_bwd)
    wf2, wb2 = _pack(conv2.weight, need_bwd)
    _conv(x, ldx, B, H, W, cin, cmid, 3, wf1, None, r1, cmid)
    m1, i1 = _bn_relu_fwd(ws, r1, cmid, npix, cmid, bn1, a1, cmid, update_stats)
    _conv(a1, cmid, B, H, W, cmid, cout, 3, wf2, None, r2, cout)
    m2, i2 = _bn_relu_fwd(ws, r2, cout, npix, cout, bn2, out, ldo, update_stats)
    st[name] = dict(x=x, ldx=ldx, B=B, H=H, W=W, cin=cin, cmid=cmid, cout=cout, r1=r1, a1=a1, r2=r2, out=out,
                    ldo=ldo, m1=m1, i1=i1, m2=m2, i2=i2, wb1=wb1, wb2=wb2, g1=bn1.weight.detach(),
                    g2=bn2.weight.detach(), be1=bn1.bias.detach(), be2=bn2.bias.detach())


def _double_conv_bwd(ws, s, dout, lddo, grads, prefix, dx, lddx):
    """dout: gradient w.r.t. the block output (strided).  Writes dx (if not None)."""
    B, H, W, cin, cmid, cout = s["B"], s["H"], s["W"], s["cin"], s["cmid"], s["cout"]
    npix = B * H * W
    dev = dout.device
    da1 = _empty(npix, cmid, dev=dev)
    if s.get("tma"):
        # the batch-norm backward emits its result directly as the bf16 operand shared by the weight and the
        # data gradient; the fp32 copy is never needed
        dr2b = torch.empty(npix, cout, dtype=torch.bfloat16, device=dev)
        dg2, db2 = _bn_relu_bwd(ws, s["r2"], cout, dout, lddo, npix, cout, s["g2"], s["be2"], s["m2"], s["i2"],
                                None, 0, dxb=dr2b)
        dw2 = _wgrad_tma(ws, s["a1b"], dr2b, B, H, W, cmid, cout)
        _conv_tma(dr2b, B, H, W, cout, cmid, s["wd2"], da1, cmid)
        dr1b = torch.empty(npix, cmid, dtype=torch.bfloat16, device=dev)
        dg1, db1 = _bn_relu_bwd(ws, s["r1"], cmid, da1, cmid, npix, cmid, s["g1"], s["be1"], s["m1"], s["i1"],
                                None, 0, dxb=dr1b)
        dw1 = _wgrad_tma(ws, s["xb"], dr1b, B, H, W, cin, cmid)
        if dx is not None:
            _conv_tma(dr1b, B, H, W, cmid, cin, s["wd1"], dx, lddx)
    else:
        dr2 = _empty(npix, cout, dev=dev)
        dg2, db2 = _bn_relu_bwd(ws, s["r2"], cout, dout, lddo, npix, cout, s["g2"], s["be2"], s["m2"], s["i2"],
                                dr2, cout)
        dr1 = _empty(npix, cmid, dev=dev)
        dw2, _ = _conv_wgrad(ws, s["a1"], cmid, dr2, cout, B, H, W, cmid, cout, 3, False)
        _conv(dr2, cout, B, H, W, cout, cmid, 3, s["wb2"], None, da1, cmid)
        dg1, db1 = _bn_relu_bwd(ws, s["r1"], cmid, da1, cmid, npix, cmid, s["g1"], s["be1"], s["m1"], s["i1"],
                                dr1, cmid)
        dw1, _ = _conv_wgrad(ws, s["x"], s["ldx"], dr1, cmid, B, H, W, cin, cmid, 3, False)
        if dx is not None:
            _conv(dr1, cmid, B, H, W, cmid, cin, 3, s["wb1"], None, dx, lddx)
    grads[prefix + ".0.weight"] = dw1
    grads[prefix + ".1.weight"], grads[prefix + ".1.bias"] = dg1, db1
    grads[prefix + ".3.weight"] = dw2
    grads[prefix + ".4.weight"], grads[prefix + ".4.bias"] = dg2, db2


def _dc_mods(dc):
    seq = dc.double_conv
    return seq[0], seq[1], seq[3], seq[4]


def unet_forward(net, x, need_bwd=True, update_stats=True):
    """``net``: the UNet module (parameter container).  x: (B,3,H,W) or (3,H,W) on a CUDA device.
    Returns (out (B,1,H/2,W/2), state)."""
    _enter(net)
    global EVAL
    EVAL = not net.training
    if EVAL and need_bwd:
        raise NotImplementedError("UNet in eval mode is inference only here (BatchNorm backward with frozen statistics is not "
                                  "built): call it under torch.no_grad(), or keep the module in train mode like the reference")
    if x.dim() == 3:                                   # train.py:465 passes (C,H,W)
        x = x.unsqueeze(0)
    x = x.detach().float().contiguous()
    tm_lib.require_cuda(x, "UNet input")
    B, C, H, W = x.shape
    if C != 3 or H % 8 or W % 8:
        raise RuntimeError("UNet expects (B,3,H,W) with H and W multiples of 8")
    dev = x.device
    mode = 0 if net.pooling == "max" else 1
    ws = _Ws(dev)
    st = {"ws": ws, "mode": mode, "B": B, "H": H, "W": W}
    x0 = _empty(B * H * W, 3, dev=dev)
    call("tm_nchw_to_nhwc", B, 3, H, W, x, x0, 3, stream())
    chans = [16, 32, 64, 128]
    Hs = [H, H // 2, H // 4, H // 8]
    Ws_ = [W, W // 2, W // 4, W // 8]
    cat = [_empty(B * Hs[i] * Ws_[i], 2 * chans[i], dev=dev) for i in range(3)]     # [skip | up]
    x4 = _empty(B * Hs[3] * Ws_[3], 128, dev=dev)
    enc = [net.inc, net.down1.maxpool_conv[1], net.down2.maxpool_conv[1], net.down3.maxpool_conv[1]]
    cur, ld, cin = x0, 3, 3
    st["pool"] = []
    for i in range(4):
        if i > 0:                                      # Down = pool -> DoubleConv (Unet.py:33-36)
            pooled = _empty(B * Hs[i] * Ws_[i], chans[i - 1], dev=dev)
            idx = torch.empty(B * Hs[i] * Ws_[i] * chans[i - 1], dtype=torch.uint8, device=dev) if mode == 0 else None
            call("tm_pool2x2_forward", B, Hs[i - 1], Ws_[i - 1], chans[i - 1], mode, cur, ld, pooled, chans[i - 1],
                 idx, 0, stream())
            st["pool"].append(dict(idx=idx, src_ld=ld, C=chans[i - 1], H=Hs[i - 1], W=Ws_[i - 1]))
            cur, ld, cin = pooled, chans[i - 1], chans[i - 1]
        out, ldo = (cat[i], 2 * chans[i]) if i < 3 else (x4, 128)
        _double_conv_fwd(ws, st, f"enc{i}", _dc_mods(enc[i]), cur, ld, B, Hs[i], Ws_[i], cin, chans[i], out, ldo,
                         update_stats, need_bwd)
        cur, ld = out, ldo
    # decoder: Up = ConvTranspose2d(k2,s2) -> cat([skip, up]) -> DoubleConv (Unet.py:53-68)
    ups = [net.up1, net.up2, net.up3]
    y, ldy, cy = x4, 128, 128
    for j, up in enumerate(ups):
        i = 2 - j                                      # resolution index of the skip
        cout_t = cy // 2
        half = cat[i][:, chans[i]:]                    # second half of the concat buffer
        wup = up.up.weight.detach().float().contiguous()
        if (USE_TMA and _prec(None) == 0
                and tm_lib.ws_bytes("tm_convt2x2_bf16_supported", B, Hs[i + 1], Ws_[i + 1], cy, cout_t) == 1):
            # bf16 mode: one-tap TMA convolution whose epilogue scatters the 2x2 output windows
            yb = _to_bf16(y, ldy, B * Hs[i + 1] * Ws_[i + 1], cy)
            wfq = torch.empty(4 * cout_t, cy, dtype=torch.bfloat16, device=dev)
            wdq = torch.empty(4, cy, cout_t, dtype=torch.bfloat16, device=dev) if need_bwd else None
            call("tm_convt2x2_pack_bf16", cy, cout_t, wup, wfq, wdq, stream())
            call("tm_convt2x2_bf16", B, Hs[i + 1], Ws_[i + 1], cy, cout_t, yb, wfq, up.up.bias.detach(), half,
                 2 * chans[i], tm_lib.err_flag(dev), stream())
            st[f"up{j}"] = dict(x=y, ldx=ldy, cin=cy, cout=cout_t, wdq=wdq, xb=yb if need_bwd else None, i=i)
        else:
            wt, wtT = _empty(cy, 4 * cout_t, dev=dev), _empty(4 * cout_t, cy, dev=dev)
            call("tm_convt_pack_weight", cy, cout_t, wup, wt, wtT, stream())
            call("tm_convt2x2_nhwc", B, Hs[i + 1], Ws_[i + 1], cy, cout_t, y, ldy, wt, up.up.bias.detach(), half,
                 2 * chans[i], Hs[i], Ws_[i], 0, 0, stream())
            st[f"up{j}"] = dict(x=y, ldx=ldy, cin=cy, cout=cout_t, wtT=wtT, i=i)
        out = _empty(B * Hs[i] * Ws_[i], chans[i], dev=dev)
        _double_conv_fwd(ws, st, f"dec{j}", _dc_mods(up.conv), cat[i], 2 * chans[i], B, Hs[i], Ws_[i], 2 * chans[i],
                         chans[i], out, chans[i], update_stats, need_bwd)
        y, ldy, cy = out, chans[i], chans[i]
    # OutConv: 1x1 conv (bias) -> pool -> ReLU (Unet.py:74-78)
    oc = net.outc.conv[0]
    wo = oc.weight.detach().float().reshape(-1).contiguous()         # (1,16,1,1) -> [16]
    o_raw = _empty(B * H * W, 1, dev=dev)
    call("tm_conv1x1_c1_forward", B * H * W, 16, y, ldy, wo, oc.bias.detach(), o_raw, 1, stream())
    out = _empty(B, 1, H // 2, W // 2, dev=dev)
    oidx = torch.empty(B * (H // 2) * (W // 2), dtype=torch.uint8, device=dev) if mode == 0 else None
    call("tm_pool2x2_forward", B, H, W, 1, mode, o_raw, 1, out, 1, oidx, RELU, stream())
    st.update(y3=y, ld3=ldy, wo=wo, out=out, oidx=oidx, cat=cat, chans=chans, Hs=Hs, Ws=Ws_)
    return out, st


def unet_backward(net, st, gout):
    """gout: (B,1,H/2,W/2).  Returns {state_dict-style name: gradient} for every parameter."""
    _enter(net)
    B, H, W, mode = st["B"], st["H"], st["W"], st["mode"]
    ws, chans, Hs, Ws_, cat = st["ws"], st["chans"], st["Hs"], st["Ws"], st["cat"]
    dev = gout.device
    gout = gout.detach().float().contiguous()
    grads = {}
    d_oraw = _empty(B * H * W, 1, dev=dev)
    call("tm_pool2x2_backward", B, H, W, 1, mode, gout, 1, st["out"], 1, st["oidx"], d_oraw, 1, RELU, None, 0, stream())
    dw, db = _empty(1, 16, 1, 1, dev=dev), _empty(1, dev=dev)
    nb = tm_lib.ws_bytes("tm_conv1x1_c1_wgrad_ws", 16)
    call("tm_conv1x1_c1_wgrad", B * H * W, 16, st["y3"], st["ld3"], d_oraw, 1, dw, db, ws.get(nb), nb, stream())
    grads["outc.conv.0.weight"], grads["outc.conv.0.bias"] = dw, db
    dy = _empty(B * H * W, 16, dev=dev)
    call("tm_conv1x1_c1_dgrad", B * H * W, 16, d_oraw, 1, st["wo"], dy, 16, stream())
    lddy = 16
    dskip = [None, None, None]
    names = ["up1", "up2", "up3"]
    for j in (2, 1, 0):
        i = 2 - j
        dcat = _empty(B * Hs[i] * Ws_[i], 2 * chans[i], dev=dev)
        _double_conv_bwd(ws, st[f"dec{j}"], dy, lddy, grads, f"{names[j]}.conv.double_conv", dcat, 2 * chans[i])
        dskip[i] = dcat                                  # first half = gradient of the skip tensor
        u = st[f"up{j}"]
        half = dcat[:, chans[i]:]
        dy = _empty(B * Hs[i + 1] * Ws_[i + 1], u["cin"], dev=dev)
        if u.get("wdq") is not None:
            # bf16 mode: weight and data gradient on the TMA path, both from one bf16 copy of the gradient half
            npo = B * Hs[i] * Ws_[i]
            halfb = _to_bf16(half, 2 * chans[i], npo, u["cout"])
            dwu, dbt = _empty(u["cin"], u["cout"], 2, 2, dev=dev), _empty(u["cout"], dev=dev)
            nb = tm_lib.ws_bytes("tm_convt2x2_bf16_wgrad_ws", B, Hs[i + 1], Ws_[i + 1], u["cin"], u["cout"])
            call("tm_convt2x2_bf16_wgrad", B, Hs[i + 1], Ws_[i + 1], u["cin"], u["cout"], u["xb"], halfb, dwu, ws.get(nb), nb,
                 tm_lib.err_flag(dev), stream())
            nbc = tm_lib.ws_bytes("tm_colsum_ws", npo, u["cout"])
            call("tm_colsum", npo, u["cout"], half, 2 * chans[i], None, dbt, 0, tm_lib.workspace(nbc, dev), nbc, stream())
            call("tm_convt2x2_bf16_dgrad", B, Hs[i + 1], Ws_[i + 1], u["cin"], u["cout"], halfb, u["wdq"], dy, u["cin"],
                 tm_lib.err_flag(dev), stream())
        else:
            nb = tm_lib.ws_bytes("tm_convt2x2_wgrad_ws", B, Hs[i + 1], Ws_[i + 1], u["cin"], u["cout"])
            dwt, dbt = _empty(u["cin"], 4 * u["cout"], dev=dev), _empty(u["cout"], dev=dev)
            call("tm_convt2x2_wgrad_nhwc", B, Hs[i + 1], Ws_[i + 1], u["cin"], u["cout"], u["x"], u["ldx"], half,
                 2 * chans[i], Hs[i], Ws_[i], 0, 0, dwt, dbt, ws.get(nb), nb, stream())
            dwu = _empty(u["cin"], u["cout"], 2, 2, dev=dev)
            call("tm_convt_unpack_wgrad", u["cin"], u["cout"], dwt, dwu, stream())
            call("tm_convt2x2_dgrad_nhwc", B, Hs[i + 1], Ws_[i + 1], u["cin"], u["cout"], half, 2 * chans[i], Hs[i],
                 Ws_[i], 0, 0, u["wtT"], dy, u["cin"], stream())
        grads[f"{names[j]}.up.weight"], grads[f"{names[j]}.up.bias"] = dwu, dbt
        lddy = u["cin"]
    # encoder, deepest first; dy is the gradient of x4
    enc_names = ["inc.double_conv", "down1.maxpool_conv.1.double_conv", "down2.maxpool_conv.1.double_conv",
                 "down3.maxpool_conv.1.double_conv"]
    for i in (3, 2, 1, 0):
        s = st[f"enc{i}"]
        if i > 0:
            dpool = _empty(B * Hs[i] * Ws_[i], chans[i - 1], dev=dev)
            _double_conv_bwd(ws, s, dy, lddy, grads, enc_names[i], dpool, chans[i - 1])
            p = st["pool"][i - 1]
            dprev = _empty(B * Hs[i - 1] * Ws_[i - 1], chans[i - 1], dev=dev)
            # pooling backward and the gradient arriving through the skip connection (first half of dcat) in one pass
            call("tm_pool2x2_backward", B, Hs[i - 1], Ws_[i - 1], chans[i - 1], mode, dpool, chans[i - 1], None, 0,
                 p["idx"], dprev, chans[i - 1], 0, dskip[i - 1], 2 * chans[i - 1], stream())
            dy, lddy = dprev, chans[i - 1]
        else:
            _double_conv_bwd(ws, s, dy, lddy, grads, enc_names[0], None, 0)
    return grads


UNET_PARAM_ORDER = None


def unet_param_names(net):
    return [k for k, _ in net.named_parameters()]


class UNetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, net, x, need, *params):
        # `need` is decided by the caller: grad mode is always off inside Function.forward
        out, st = unet_forward(net, x, need_bwd=need, update_stats=net.training)
        ctx.net, ctx.st = net, st
        ctx.names = unet_param_names(net)
        return out

    @staticmethod
    def backward(ctx, gout):
        grads = unet_backward(ctx.net, ctx.st, gout)
        return (None, None, None) + tuple(grads[k] for k in ctx.names)


# --------------------------------------------------------------------------------------------
# LayoutNet (model.py:216-247): conv9 -> ReLU -> pool -> conv7 -> ReLU -> pool -> conv9 -> ReLU
#                               -> conv7 -> LeakyReLU(0.1)
# --------------------------------------------------------------------------------------------
LAYOUT_SPEC = [(0, 2, 32, 9, True), (3, 32, 64, 7, True), (6, 64, 32, 9, False), (8, 32, 1, 7, False)]


def layoutnet_forward(net, x, need_bwd=True):
    _enter(net)
    squeeze = x.dim() == 3
    if squeeze:
        x = x.unsqueeze(0)
    x = x.detach().float().contiguous()
    tm_lib.require_cuda(x, "LayoutNet input")
    B, C, H, W = x.shape
    dev = x.device
    mode = 0 if net.pooling == "max" else 1
    ws = _Ws(dev)
    cur = _empty(B * H * W, C, dev=dev)
    call("tm_nchw_to_nhwc", B, C, H, W, x, cur, C, stream())
    st = dict(ws=ws, mode=mode, B=B, layers=[], squeeze=squeeze)
    h, w = H, W
    for li, (idx, cin, cout, k, pool) in enumerate(LAYOUT_SPEC):
        conv = net.encode[idx]
        wf, wb = _pack(conv.weight, need_bwd)
        last = li == len(LAYOUT_SPEC) - 1
        y = _empty(B * h * w, cout, dev=dev)
        _conv(cur, cin, B, h, w, cin, cout, k, wf, conv.bias.detach(), y, cout, 0 if last else RELU)
        rec = dict(x=cur, y=y, cin=cin, cout=cout, k=k, H=h, W=w, wb=wb, pool=None, name=f"encode.{idx}")
        if last:
            act = _empty(B * h * w, cout, dev=dev)
            call("tm_leaky_relu_forward", B * h * w * cout, y, 0.1, act, stream())
            rec["act"] = act
            cur = act
        elif pool:
            pooled = _empty(B * (h // 2) * (w // 2), cout, dev=dev)
            pidx = torch.empty(B * (h // 2) * (w // 2) * cout, dtype=torch.uint8, device=dev) if mode == 0 else None
            call("tm_pool2x2_forward", B, h, w, cout, mode, y, cout, pooled, cout, pidx, 0, stream())
            rec["pool"] = pidx if mode == 0 else True
            cur, h, w = pooled, h // 2, w // 2
        else:
            cur = y
        st["layers"].append(rec)
    out = cur.reshape(B, 1, h, w)                         # C == 1: NHWC == NCHW
    st["out_hw"] = (h, w)
    return (out[0] if squeeze else out), st


def layoutnet_backward(net, st, gout):
    _enter(net)
    B, mode, ws = st["B"], st["mode"], st["ws"]
    dev = gout.device
    g = gout.detach().float().contiguous().reshape(-1, 1)
    grads = {}
    n_layers = len(st["layers"])
    for li in range(n_layers - 1, -1, -1):
        r = st["layers"][li]
        h, w, cin, cout, k = r["H"], r["W"], r["cin"], r["cout"], r["k"]
        npix = B * h * w
        if li == n_layers - 1:
            dy = _empty(npix, cout, dev=dev)
            call("tm_leaky_relu_backward", npix * cout, r["act"], g, 0.1, dy, stream())
        else:
            if r["pool"] is not None:
                dyp = _empty(npix, cout, dev=dev)
                call("tm_pool2x2_backward", B, h, w, cout, mode, g, cout, None, 0,
                     r["pool"] if mode == 0 else None, dyp, cout, 0, None, 0, stream())
                g = dyp
            dy = _empty(npix, cout, dev=dev)
            call("tm_leaky_relu_backward", npix * cout, r["y"], g, 0.0, dy, stream())   # ReLU
        dw, db = _conv_wgrad(ws, r["x"], cin, dy, cout, B, h, w, cin, cout, k, True)
        grads[r["name"] + ".weight"], grads[r["name"] + ".bias"] = dw, db
        if li > 0:
            dx = _empty(npix, cin, dev=dev)
            _conv(dy, cout, B, h, w, cout, cin, k, r["wb"], None, dx, cin)
            g = dx
    return grads


class LayoutNetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, net, x, need, *params):
        out, st = layoutnet_forward(net, x, need_bwd=need)
        ctx.net, ctx.st = net, st
        ctx.names = [k for k, _ in net.named_parameters()]
        return out

    @staticmethod
    def backward(ctx, gout):
        grads = layoutnet_backward(ctx.net, ctx.st, gout)
        return (None, None, None) + tuple(grads[k] for k in ctx.names)
