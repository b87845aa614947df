"""Floating-point side of the oracle: a pure-PyTorch CPU restatement.

Oracle / test infrastructure only (see oracle/__init__.py).  Every function is
written against a ``state_dict``-style mapping that uses the REFERENCE's
parameter names, so weights can be exchanged with the
reference classes, with the product modules and with the golden fixtures.

Follows, line by line in meaning (not in code):

* ``MLP``                     src/model.py:10-24
* ``PathConv.forward`` + UDFs src/model.py:88-116,138-153,158-213
* ``PathModel.forward``       src/model.py:269-292
* ``LayoutNet``               src/model.py:216-247
* ``UNet`` and its blocks     src/Unet.py:8-119
* the per-design step         src/train.py:465,490-522,552-553

Pinned by ``tests/test_oracle_pinning.py`` against the reference's own modules
run in the dev container and by the committed fixtures under ``tests/golden``.
"""
import torch
import torch.nn.functional as F


# ---------------------------------------------------------------------------
# MLP (model.py:10-24): Linear -> LeakyReLU(negative_slope) -> ... -> Linear
# ---------------------------------------------------------------------------
def mlp(sd, prefix, x, negative_slope=0.0):
    idx = sorted({int(k[len(prefix) + 8:].split(".")[0]) for k in sd
                  if k.startswith(prefix + ".layers.") and k.endswith(".weight")})
    for j, i in enumerate(idx):
        x = F.linear(x, sd[f"{prefix}.layers.{i}.weight"], sd[f"{prefix}.layers.{i}.bias"])
        if j < len(idx) - 1:
            x = F.leaky_relu(x, negative_slope)
    return x


# ---------------------------------------------------------------------------
# PathConv (model.py:158-213) over all levels
# ---------------------------------------------------------------------------
def _edge_slots(indptr, nodes):
    start = indptr[nodes]
    deg = indptr[nodes + 1] - start
    rows = torch.repeat_interleave(torch.arange(nodes.numel()), deg)
    first = torch.cumsum(deg, 0) - deg
    slots = torch.repeat_interleave(start, deg) + (torch.arange(int(deg.sum())) - first[rows])
    return rows, slots, deg


def gnn_level(sd, prefix, h, level_id, nodes, net_csr, cell_csr, cell_feat, net_feat, gate=None):
    """One ``PathConv.forward`` call: returns the new ``h`` (out of place).
    ``gate`` (bool (n, D), optional): teacher-forced output gates -- ``relu(z)`` becomes ``z * gate[v]``, i.e. the
    piecewise-linear branch ANOTHER evaluation took (its ``H > 0``), so that gradients can be compared element by
    element even when the two evaluations disagree about the sign of a pre-activation that is ~0."""
    nodes = torch.as_tensor(nodes, dtype=torch.int64)
    if nodes.numel() == 0:                                   # DGL pull on [] is a no-op
        return h
    D = h.shape[1]
    if level_id % 2 == 1:                                    # model.py:185-187 net level
        indptr, src = net_csr
        rows, slots, deg = _edge_slots(indptr, nodes)
        agg = torch.zeros(nodes.numel(), D, dtype=h.dtype).index_add(0, rows, h[src[slots]])
        agg = agg / deg.clamp(min=1).to(h.dtype)[:, None]    # builtin mean; 0 for no in-edges
        new = mlp(sd, f"{prefix}.fc_net_self", net_feat[nodes]) + agg      # model.py:103-108
    elif level_id == 0:                                      # model.py:148-153,200-204
        new = mlp(sd, f"{prefix}.fc_cell_self", cell_feat[nodes])
    else:                                                    # model.py:113-116,138-146
        indptr, src = cell_csr
        rows, slots, deg = _edge_slots(indptr, nodes)
        m = h[src[slots]]
        mx = torch.full((nodes.numel(), D), -float("inf"), dtype=h.dtype)
        mx = mx.scatter_reduce(0, rows[:, None].expand(-1, D), m.detach(), "amax")
        e = torch.exp(m - mx[rows])
        s = torch.zeros(nodes.numel(), D, dtype=h.dtype).index_add(0, rows, e)
        w = e / s[rows]                                      # softmax over in-edges, per channel
        agg = torch.zeros(nodes.numel(), D, dtype=h.dtype).index_add(0, rows, m * w)
        new = mlp(sd, f"{prefix}.fc_cell_self", cell_feat[nodes]) + \
            mlp(sd, f"{prefix}.fc_cell_neigh", agg)
    act = F.relu(new) if gate is None else new * gate[nodes].to(new.dtype)
    return h.index_copy(0, nodes, act)                       # model.py:207-208


def gnn_propagate(sd, prefix, n, levels, net_csr, cell_csr, cell_feat, net_feat, out_dim=128, gate=None):
    """All levels in order from h = 0 (train.py:342,490-503).  Returns H (n, out_dim)."""
    h = torch.zeros(n, out_dim, dtype=cell_feat.dtype)
    for lid, nodes in enumerate(levels):
        h = gnn_level(sd, prefix, h, lid, nodes, net_csr, cell_csr, cell_feat, net_feat, gate=gate)
    return h


# ---------------------------------------------------------------------------
# PathModel head (model.py:269-292) and the mask fusion (train.py:500-501)
# ---------------------------------------------------------------------------
def dense_mask_rows(mask_indptr, mask_cols, rows, width, dtype=torch.float32):
    out = torch.zeros(len(rows), width, dtype=dtype)
    for i, r in enumerate(rows):
        out[i, mask_cols[mask_indptr[r]:mask_indptr[r + 1]].long()] = 1
    return out


def head_level(sd, h_rows, path_map, level_id, global_dim=32):
    """``PathModel.forward`` after the GNN call, for the endpoints of one level."""
    parts = []
    if h_rows is not None:
        parts.append(h_rows)
    if path_map is not None:
        parts.append(F.linear(path_map, sd["fcn.weight"], sd["fcn.bias"]))      # model.py:272
    lvl = torch.tensor([float(level_id)], dtype=torch.float32)
    g = mlp(sd, "mlp_alpha", lvl).expand(parts[0].shape[0], global_dim)         # model.py:280
    parts.append(g)
    return mlp(sd, "mlp_fuse", torch.cat(parts, 1)).squeeze(-1)                 # model.py:290-292


# ---------------------------------------------------------------------------
# UNet (Unet.py) and LayoutNet (model.py:216-247)
# ---------------------------------------------------------------------------
def _pool(x, pooling):
    return F.max_pool2d(x, 2) if pooling == "max" else F.avg_pool2d(x, 2)


def _bf16_rn(t):
    """Round to bf16 (nearest even) and back: what the product's bf16 mode does to a tensor-core operand."""
    return t.to(torch.bfloat16).to(t.dtype)


class _ContractBf16(torch.autograd.Function):
    """A convolution / transposed convolution whose THREE contractions (forward, data gradient, weight
    gradient) see bf16-rounded operands and accumulate in the tensor's own precision -- the rounding points of
    the product's bf16 image branch (tm_unet.py): activations and weights are rounded when they become a
    tensor-core operand, the incoming gradient is rounded ONCE and that copy feeds both gradients, the bias
    gradient is summed from the unrounded gradient.  Everything else (BN, ReLU, pooling, 1x1 OutConv) stays as is."""

    @staticmethod
    def forward(ctx, x, w, b, kind):
        xb, wb = _bf16_rn(x.detach()), _bf16_rn(w.detach())
        ctx.save_for_backward(xb, wb)
        ctx.kind, ctx.has_b = kind, b is not None
        if kind == "conv3x3":
            return F.conv2d(xb, wb, b, padding=1)
        return F.conv_transpose2d(xb, wb, b, stride=2)

    @staticmethod
    def backward(ctx, g):
        xb, wb = ctx.saved_tensors
        gb = _bf16_rn(g)
        with torch.enable_grad():
            x_, w_ = xb.detach().requires_grad_(True), wb.detach().requires_grad_(True)
            y = F.conv2d(x_, w_, None, padding=1) if ctx.kind == "conv3x3" else F.conv_transpose2d(x_, w_, None, stride=2)
            dx, dw = torch.autograd.grad(y, (x_, w_), gb)
        db = g.sum((0, 2, 3)) if ctx.has_b else None
        return dx, dw, db, None


def _conv3x3(x, w, rounding):
    return F.conv2d(x, w, None, padding=1) if rounding is None else _ContractBf16.apply(x, w, None, "conv3x3")


def _force(y, forced, key):
    """Teacher forcing: give ``y`` the VALUE of ``forced[key]`` (another implementation's result for the same
    tensor) while gradients keep flowing through ``y``.  With every contraction output forced, the ReLU gates,
    max-pool winners and batch statistics downstream are the other implementation's, so the two backward passes
    differentiate the same piecewise-linear function and can be compared element by element."""
    if forced is None or key not in forced:
        return y
    return y + (forced[key].to(y.dtype) - y).detach()


def _double_conv(sd, p, x, stats, momentum=0.1, eps=1e-5, rounding=None, forced=None):
    for ic, ib in ((0, 1), (3, 4)):                          # Unet.py:15-22
        x = _force(_conv3x3(x, sd[f"{p}.{ic}.weight"], rounding), forced, f"{p}.{ic}")
        rm = sd[f"{p}.{ib}.running_mean"].detach().clone()
        rv = sd[f"{p}.{ib}.running_var"].detach().clone()
        x = F.batch_norm(x, rm, rv, sd[f"{p}.{ib}.weight"], sd[f"{p}.{ib}.bias"],
                         training=True, momentum=momentum, eps=eps)
        stats[f"{p}.{ib}.running_mean"], stats[f"{p}.{ib}.running_var"] = rm, rv
        x = F.relu(x)
    return x


def unet_forward(sd, x, pooling="max", rounding=None, forced=None):
    """Train-mode UNet forward (the reference never calls ``.eval()``, train.py:436-437).

    Returns ``(out, new_running_stats)``; ``x`` may be (C,H,W) or (B,C,H,W).
    ``rounding="bf16"``: NOT the reference's arithmetic -- the same network with the operands of every 3x3
    convolution and transposed convolution rounded to bf16 (see ``_ContractBf16``), i.e. the oracle of the
    product's bf16 image branch (BASELINE config 4), so that branch can be held to rtol 2e-2 per element on
    the output AND on every gradient; its distance to the unrounded oracle is reported separately.
    ``forced``: {"<block>.double_conv.<0|3>" | "up<k>.up": tensor (B,C,H,W)} -- see ``_force``.
    """
    if x.dim() == 3:
        x = x.unsqueeze(0)
    st = {}
    r = rounding
    x1 = _double_conv(sd, "inc.double_conv", x, st, rounding=r, forced=forced)
    x2 = _double_conv(sd, "down1.maxpool_conv.1.double_conv", _pool(x1, pooling), st, rounding=r, forced=forced)
    x3 = _double_conv(sd, "down2.maxpool_conv.1.double_conv", _pool(x2, pooling), st, rounding=r, forced=forced)
    x4 = _double_conv(sd, "down3.maxpool_conv.1.double_conv", _pool(x3, pooling), st, rounding=r, forced=forced)
    y = x4
    for name, skip in (("up1", x3), ("up2", x2), ("up3", x1)):
        if r is None:
            y = F.conv_transpose2d(y, sd[f"{name}.up.weight"], sd[f"{name}.up.bias"], stride=2)
        else:
            y = _ContractBf16.apply(y, sd[f"{name}.up.weight"], sd[f"{name}.up.bias"], "convt2x2")
        y = _force(y, forced, f"{name}.up")
        dy, dx = skip.shape[2] - y.shape[2], skip.shape[3] - y.shape[3]
        y = F.pad(y, [dx // 2, dx - dx // 2, dy // 2, dy - dy // 2])            # Unet.py:59-63
        y = _double_conv(sd, f"{name}.conv.double_conv", torch.cat([skip, y], 1), st, rounding=r, forced=forced)
    y = F.conv2d(y, sd["outc.conv.0.weight"], sd["outc.conv.0.bias"])           # Unet.py:74-78
    return F.relu(_pool(y, pooling)), st


def layoutnet_forward(sd, x, pooling="max"):
    y = F.relu(F.conv2d(x, sd["encode.0.weight"], sd["encode.0.bias"], padding=4))
    y = _pool(y, pooling)
    y = F.relu(F.conv2d(y, sd["encode.3.weight"], sd["encode.3.bias"], padding=3))
    y = _pool(y, pooling)
    y = F.relu(F.conv2d(y, sd["encode.6.weight"], sd["encode.6.bias"], padding=4))
    y = F.conv2d(y, sd["encode.8.weight"], sd["encode.8.bias"], padding=3)
    return F.leaky_relu(y, 0.1)


# ---------------------------------------------------------------------------
# one design step (train.py:465,490-522,552-553): predictions, loss, gradients
# ---------------------------------------------------------------------------
def design_step(sd_model, sd_cnn, d, pooling="max", with_grad=True, cnn="unet", unet_rounding=None, unet_forced=None,
                gnn_gate=None):
    """``d`` is a dict of CPU tensors:
    n, levels (list of int64 tensors), net_csr, cell_csr (indptr, src int64),
    cell_feat, net_feat, image (C,H,W), endpoints (int64, grouped by level in
    ascending level order), endpoint_level (int64), mask_indptr, mask_cols,
    arrival_time (per endpoint).
    """
    P = {k: (v.detach().clone().requires_grad_(with_grad) if v.is_floating_point() else v)
         for k, v in sd_model.items()}
    C = {k: (v.detach().clone().requires_grad_(with_grad)
             if v.is_floating_point() and "running" not in k else v)
         for k, v in sd_cnn.items()}
    if cnn == "unet":
        fmap, stats = unet_forward(C, d["image"], pooling, rounding=unet_rounding, forced=unet_forced)
    else:
        fmap, stats = layoutnet_forward(C, d["image"], pooling), {}
    feat = fmap.reshape(1, -1)                                                  # train.py:465
    H = gnn_propagate(P, "gnn", d["n"], d["levels"], d["net_csr"], d["cell_csr"],
                      d["cell_feat"], d["net_feat"], gate=gnn_gate)             # gnn_gate: teacher-forced output gates
    preds = []
    ep, el = d["endpoints"], d["endpoint_level"]
    for lid in torch.unique(el).tolist():                                        # ascending levels
        sel = torch.nonzero(el == lid).squeeze(1)
        dm = dense_mask_rows(d["mask_indptr"], d["mask_cols"], sel.tolist(), feat.shape[1])
        preds.append(head_level(P, H[ep[sel]], dm * feat, lid))                 # train.py:500-503
    pred = torch.cat(preds)
    loss = F.mse_loss(pred, d["arrival_time"])                                  # train.py:520-522
    out = {"pred": pred.detach(), "loss": loss.detach(), "H": H.detach(),
           "feat_map": fmap.detach(), "bn_stats": stats}
    if with_grad:
        names = [k for k, v in P.items() if v.requires_grad] + \
                ["cnn." + k for k, v in C.items() if v.requires_grad]
        tens = [v for v in P.values() if v.requires_grad] + [v for v in C.values() if v.requires_grad]
        gr = torch.autograd.grad(loss, tens, allow_unused=True)
        out["grads"] = {k: g for k, g in zip(names, gr)}
    return out
