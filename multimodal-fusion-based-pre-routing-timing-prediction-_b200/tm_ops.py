"""Host-side operators over ``libtm_b200``: GNN propagation, mask fusion, MLPs.

Two layers:

* plain functions (``gnn_forward`` / ``gnn_backward`` / ``fusion_*`` / ``mlp2_*``) that launch the
  CUDA kernels on the current stream and return raw tensors -- used by the fused design step
  (``tm_engine.py``) with no autograd bookkeeping;
* ``torch.autograd.Function`` wrappers (``GnnPropagateFn``, ``MaskFusionFn``, ``LinearFn``) over the
  same functions -- used by the drop-in ``nn.Module`` surface (``model.py``) so that the reference's
  ``train.py`` loop (``loss.backward(retain_graph=True)``, anomaly mode) works unchanged.
  Saved buffers are never freed or overwritten in ``backward``.

Reference semantics: src/model.py:10-24 (MLP), :88-116,:138-213 (PathConv), :269-292 (PathModel),
src/train.py:500-501 (mask fusion).
"""
import os

import torch

import tm_lib
from tm_lib import call, stream

D = 128
BIAS, RELU, MASK, ACCUM = 1, 2, 4, 8

# Arithmetic of the dense contractions (MLPs, weight gradients):
#   "tf32x3" tcgen05 kind::tf32, x = hi + lo with hi = the fp32 word itself (the tensor core reads its
#          upper 19 bits) and lo = x - trunc(x); hi*hi + hi*lo + lo*hi, fp32 accumulate in TMEM:
#          ~21-bit products, fp32-class accuracy -- the default for the rtol 1e-3 path;
#   "tf32" single-pass TF32 (11-bit operands);
#   "tc6"  bf16 operands split into three terms (24 bits), 6 MMAs;  "tc3": two terms, 3 MMAs;
#   "bf16" tcgen05 with plain bf16 operands (rtol 2e-2 class);
#   "fp32" the CUDA-core FFMA kernels (tm_gemm_nn / tm_gemm_tn).
MATH = os.environ.get("TM_MATH", "tf32x3")
GEN_HIDDEN = os.environ.get("TM_GEN_HIDDEN", "1") != "0"  # generate Linear(<=2, hid) hidden layers instead of storing them
FUSED_SELF_MLP = os.environ.get("TM_FUSED_SELF_MLP", "1") != "0"   # tm_selfmlp.cu for the 2 -> 256 -> 128 net-pin MLP
FUSE_RUNS = os.environ.get("TM_FUSE_RUNS", "1") != "0"    # run-length / prefix-table mask fusion forward
TC_MIN_N = 4           # Linear(hid, 1) (the head's last layer): a 128 x 32 tensor-core tile would be 97 % padding
TC_MIN_K = 16          # contractions shorter than this stay on the CUDA-core kernel (K = 1, 2: pure bandwidth)


def _precision(math=None):
    m = math or MATH
    if m not in ("tf32x3", "tf32", "tc6", "tc3", "bf16", "fp32"):
        raise ValueError(f"unknown math mode {m!r}")
    return {"tf32x3": 3, "tf32": 4, "tc6": 2, "tc3": 1, "bf16": 0, "fp32": None}[m]


def _colsum(R, C, X, ld, rows, out, accumulate):
    nb = tm_lib.ws_bytes("tm_colsum_ws", R, C)
    call("tm_colsum", R, C, X, ld, rows, out, accumulate, tm_lib.workspace(nb, out.device), nb, stream())


class _Aux:
    """Helper stream for bandwidth-bound side work (bias-gradient column sums) that is independent of
    the tensor-core GEMM enqueued next to it: the persistent GEMM CTAs fill the shared memory of every
    SM but leave thread / register room, so a plain streaming kernel co-resides with them."""
    stream = None
    pending = []

    @classmethod
    def run(cls, fn, *tensors):
        if not torch.cuda.is_available():
            return fn()
        if cls.stream is None:
            cls.stream = torch.cuda.Stream()
        cur = torch.cuda.current_stream()
        if torch.cuda.is_current_stream_capturing() or cur == cls.stream:
            return fn()
        # every tensor passed in stays referenced by the caller until its aux_join(), so no
        # record_stream (which would stall the caching allocator's reuse of these large buffers)
        cls.stream.wait_stream(cur)
        with torch.cuda.stream(cls.stream):
            fn()
        cls.pending.append(cur)

    @classmethod
    def join(cls):
        """Make the current stream wait for everything posted from it."""
        if cls.stream is not None and cls.pending:
            torch.cuda.current_stream().wait_stream(cls.stream)
            cls.pending = []


def aux_join():
    _Aux.join()


def _f32c(t):
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


def _rowmajor(t):
    """2-D tensor usable as a strided row-major operand (unit inner stride)."""
    if t.dtype != torch.float32 or t.stride(-1) != 1:
        t = t.float().contiguous()
    return t


# --------------------------------------------------------------------------------------------
# thin wrappers
# --------------------------------------------------------------------------------------------
def gemm_nn(M, N, K, A, lda, B, ldb, C, ldc, a_rows=None, c_rows=None, bias=None, mask=None, ldmask=0, flags=0,
            b_is_nk=False, math=None):
    """C = epi(A @ B).  B is [K,N] row-major, or [N,K] (an nn.Linear weight as stored) with b_is_nk."""
    if bias is not None:
        flags |= BIAS
    if mask is not None:
        flags |= MASK
    prec = _precision(math)
    if N == 1 and b_is_nk and mask is None and not (flags & ACCUM) and M > 0:
        call("tm_rowdot", M, K, A, lda, a_rows, B, bias, C, ldc, c_rows, 1 if flags & RELU else 0, stream())
        return
    if prec is None or K < TC_MIN_K or N < TC_MIN_N:
        if b_is_nk:
            B, ldb = transpose(B[:, :K] if B.shape[1] != K else B), N
        call("tm_gemm_nn", M, N, K, A, lda, a_rows, B, ldb, C, ldc, c_rows, bias, mask, ldmask, flags, stream())
    else:
        if not b_is_nk and B.dim() == 2 and B.is_contiguous() and B.shape == (K, N) and B.numel() <= (1 << 22):
            # a [K,N] weight: hand the tensor cores the K-major [N,K] form (one tiny transpose
            # instead of an MN-major operand re-staged by every CTA)
            B, ldb, b_is_nk = transpose(B), K, True
        call("tm_tc_gemm_nn", M, N, K, A, lda, a_rows, B, ldb, 1 if b_is_nk else 0, C, ldc, c_rows, bias, mask,
             ldmask, flags, prec, tm_lib.err_flag(C.device), stream())


def gemm_tn(M, N, R, A, lda, B, ldb, C, ldc, a_rows=None, b_rows=None, colsum_a=None, colsum_b=None,
            accumulate=0, math=None):
    """C (+)= A^T @ B over the R rows; optional column sums of A / B (bias gradients)."""
    prec = _precision(math)
    if prec is None or (M < TC_MIN_K and N < TC_MIN_K):
        nb = tm_lib.ws_bytes("tm_gemm_tn_ws", M, N, R)
        ws = tm_lib.workspace(nb, C.device)
        call("tm_gemm_tn", M, N, R, A, lda, a_rows, B, ldb, b_rows, C, ldc, colsum_a, colsum_b, accumulate,
             ws, nb, stream())
        return
    nb = tm_lib.ws_bytes("tm_tc_gemm_tn_ws", M, N, R)
    ws = tm_lib.workspace(nb, C.device)
    if colsum_a is not None:
        _Aux.run(lambda: _colsum(R, M, A, lda, a_rows, colsum_a, accumulate), A, colsum_a, a_rows)
    if colsum_b is not None:
        _Aux.run(lambda: _colsum(R, N, B, ldb, b_rows, colsum_b, accumulate), B, colsum_b, b_rows)
    call("tm_tc_gemm_tn", M, N, R, A, lda, a_rows, B, ldb, b_rows, C, ldc, accumulate, prec, ws, nb,
         tm_lib.err_flag(C.device), stream())


def transpose(w):
    """(rows, cols) -> (cols, rows), contiguous."""
    w = _f32c(w)
    out = torch.empty(w.shape[1], w.shape[0], dtype=torch.float32, device=w.device)
    call("tm_transpose", w.shape[0], w.shape[1], w, out, stream())
    return out


def colsum(X, R, C, ld, out=None, accumulate=0):
    if out is None:
        out = torch.empty(C, dtype=torch.float32, device=X.device)
    _colsum(R, C, X, ld, None, out, accumulate)
    return out


# --------------------------------------------------------------------------------------------
# two-layer MLP (model.py:10-24 with sizes (in, hid, out)):  y = W2 relu(W1 x + b1) + b2
# --------------------------------------------------------------------------------------------
def _recurrence_math():
    """Arithmetic of the hoisted self-term MLPs, whose output S seeds the 101-level recurrence.  Round 1 kept
    24-bit products ("tc6", six bf16 MMAs) here; measured on B200 (``profiles/diag_gemm_modes.py``) the default
    3xTF32 contraction has the same error on these shapes (max 3.0e-6 against 2.4e-6 of the row scale: the fp32
    accumulation dominates both) and every parity test -- including the full-size config 2 / 3 gradient checks -- is
    unchanged, while the step is 0.09 ms shorter.  ``TM_RECURRENCE_MATH=tc6`` restores the round-1 choice."""
    forced = os.environ.get("TM_RECURRENCE_MATH")
    if forced:
        return None if forced == "default" else forced
    return None


def mlp2_forward(x, ldx, rows, n_rows, w1, b1, w2, b2, out, ldo, out_rows=None, math=None):
    """x rows (optionally gathered by ``rows``) -> out rows (optionally scattered by ``out_rows``).
    Returns the hidden activations (n_rows, hid) needed by ``mlp2_backward``."""
    hid, kin = w1.shape
    nout = w2.shape[0]
    prec = _precision(math)
    if GEN_HIDDEN and kin <= 2 and prec is not None and n_rows > 0 and hid >= TC_MIN_K:
        # Linear(<=2, hid) first layer: the hidden activations cost 2 FMAs each to recompute, so they are
        # generated inside the GEMM's operand loader and never stored (returns None; mlp2_backward
        # regenerates them the same way)
        if FUSED_SELF_MLP and hid == 256 and nout == 128 and prec == 3 and ldo % 4 == 0 and out.data_ptr() % 16 == 0:
            # the reference's sizes: fused kernel, fp16 two-term split operands generated in place (tm_selfmlp.cu)
            nb = tm_lib.ws_bytes("tm_selfmlp_ws_bytes")
            call("tm_selfmlp_gen_forward", n_rows, x, ldx, rows, kin, _f32c(w1), _f32c(b1), _f32c(w2), _f32c(b2), out, ldo,
                 out_rows, tm_lib.workspace(nb, out.device), nb, stream())
            return None
        call("tm_tc_mlp2_smallk_forward", n_rows, hid, nout, x, ldx, rows, kin, _f32c(w1), _f32c(b1), _f32c(w2), _f32c(b2),
             out, ldo, out_rows, prec, tm_lib.err_flag(out.device), stream())
        return None
    h = torch.empty(n_rows, hid, dtype=torch.float32, device=out.device)
    if (FUSED_SELF_MLP and hid == 256 and prec == 3 and kin % 4 == 0 and kin <= 48 and ldx % 4 == 0 and x.data_ptr() % 16 == 0
            and n_rows > 0):
        # fc_cell_self's first layer (36 inputs): output-bound, fp16 two-term split on tcgen05 (tm_selfmlp.cu)
        nb = tm_lib.ws_bytes("tm_selfmlp_lin1_ws_bytes")
        fuse2 = nout == 128 and ldo % 4 == 0 and out.data_ptr() % 16 == 0
        rowmax = torch.empty(n_rows, dtype=torch.float32, device=out.device) if fuse2 else None
        call("tm_selfmlp_lin1_relu", n_rows, x, ldx, rows, kin, _f32c(w1), _f32c(b1), h, hid, rowmax,
             tm_lib.workspace(nb, out.device), nb, stream())
        if fuse2:                                   # second layer on the same machinery, per-row scale from rowmax
            nb = tm_lib.ws_bytes("tm_selfmlp_ws_bytes")
            call("tm_selfmlp_rows_forward", n_rows, h, hid, None, rowmax, _f32c(w2), _f32c(b2), out, ldo, out_rows,
                 tm_lib.workspace(nb, out.device), nb, stream())
            return h
    else:
        gemm_nn(n_rows, hid, kin, x, ldx, _f32c(w1), kin, h, hid, a_rows=rows, bias=b1, flags=RELU, b_is_nk=True, math=math)
    gemm_nn(n_rows, nout, hid, h, hid, _f32c(w2), hid, out, ldo, c_rows=out_rows, bias=b2, b_is_nk=True, math=math)
    return h


def _hidden_grad(n_rows, hid, nout, g, ldg, g_rows, w2, h, h_rows, dh):
    """dh[r] = (g[g_rows] @ w2) * (h[r] > 0), r = h_rows or the row itself."""
    if (FUSED_SELF_MLP and hid == 256 and nout == 128 and _precision() == 3 and ldg % 4 == 0 and g.data_ptr() % 16 == 0
            and h.stride(0) % 4 == 0 and h.data_ptr() % 16 == 0 and dh.data_ptr() % 16 == 0):
        # the reference's sizes: fp16 two-term split operands, W2 resident, tcgen05 (tm_selfmlp.cu)
        nb = tm_lib.ws_bytes("tm_selfmlp_rows_dh_ws_bytes")
        call("tm_selfmlp_rows_dh", n_rows, g, ldg, g_rows, _f32c(w2), h, h.stride(0), h_rows, dh, dh.stride(0),
             tm_lib.workspace(nb, dh.device), nb, stream())
        return
    gemm_nn(n_rows, hid, nout, g, ldg, _f32c(w2), hid, dh, hid, a_rows=g_rows, c_rows=h_rows, mask=h, ldmask=hid)


def mlp2_backward(x, ldx, rows, n_rows, w1, w2, h, g, ldg, g_rows=None, need_dx=False, b1=None, wgrad_stream=None,
                  h_rows=None):
    """Gradients of a two-layer MLP.  ``g``: dLoss/d(out) rows (optionally gathered by ``g_rows``).
    ``h``: the hidden activations mlp2_forward returned, or None when they were generated on the fly
    (then ``b1`` is required).  Returns (dw1, db1, dw2, db2, dx or None).
    ``wgrad_stream`` (with ``need_dx``): the data gradient is computed first on the current stream and the four
    parameter gradients on ``wgrad_stream`` (which waits for it); the caller joins that stream before reading them.
    ``h_rows``: row of ``h`` that belongs to MLP row i (default i) -- set when ``rows`` / ``g_rows`` are a subset of
    the rows the forward pass ran on (``tm_graph.BackwardCone``)."""
    hid, kin = w1.shape
    nout = w2.shape[0]
    dev = g.device
    dw2 = torch.empty(nout, hid, dtype=torch.float32, device=dev)
    db2 = torch.empty(nout, dtype=torch.float32, device=dev)
    dw1 = torch.empty(hid, kin, dtype=torch.float32, device=dev)
    db1 = torch.empty(hid, dtype=torch.float32, device=dev)
    if n_rows == 0:
        for t in (dw1, db1, dw2, db2):
            t.zero_()
        return dw1, db1, dw2, db2, None
    if h is None:                                            # generated hidden layer (kin <= 2)
        if b1 is None or need_dx:
            raise RuntimeError("mlp2_backward: generated hidden activations need b1 (and have no dx path)")
        prec = _precision()
        prec = 3 if prec is None else prec
        fused = FUSED_SELF_MLP and hid == 256 and nout == 128 and prec == 3 and ldg % 4 == 0 and g.data_ptr() % 16 == 0
        if fused:
            # the reference's sizes: bias gradient and max |g| in one pass, then the fused weight-gradient kernel
            # (hidden layer generated as a tcgen05 operand, accumulators resident in TMEM: tm_selfmlp.cu)
            gmax = torch.empty(1, dtype=torch.float32, device=dev)
            nbc = tm_lib.ws_bytes("tm_colsum_ws", n_rows, nout)
            call("tm_colsum_absmax", n_rows, nout, g, ldg, g_rows, db2, gmax, tm_lib.workspace(nbc, dev), nbc, stream())
            nb = tm_lib.ws_bytes("tm_selfmlp_wgrad2_ws_bytes")
            call("tm_selfmlp_gen_wgrad2", n_rows, g, ldg, g_rows, x, ldx, rows, kin, _f32c(w1), _f32c(b1), gmax, dw2,
                 tm_lib.workspace(nb, dev), nb, stream())
        else:
            _Aux.run(lambda: _colsum(n_rows, nout, g, ldg, g_rows, db2, 0), g, db2, g_rows)
            nb = tm_lib.ws_bytes("tm_tc_gemm_tn_ws", nout, hid, n_rows)
            call("tm_tc_mlp2_smallk_wgrad2", nout, hid, n_rows, g, ldg, g_rows, x, ldx, rows, kin, _f32c(w1), _f32c(b1), dw2,
                 prec, tm_lib.workspace(nb, dev), nb, tm_lib.err_flag(dev), stream())
        if fused:
            nb = tm_lib.ws_bytes("tm_selfmlp_bwd1_ws_bytes")
            call("tm_selfmlp_gen_bwd1", n_rows, g, ldg, g_rows, x, ldx, rows, kin, _f32c(w1), _f32c(b1), _f32c(w2), dw1, db1,
                 tm_lib.workspace(nb, dev), nb, stream())
        else:
            nb = tm_lib.ws_bytes("tm_tc_mlp1_bwd_ws", hid)
            call("tm_tc_mlp1_bwd_fused", n_rows, hid, nout, g, ldg, g_rows, transpose(_f32c(w2)), None, 0, x, ldx, rows, kin,
                 _f32c(w1), _f32c(b1), dw1, db1, tm_lib.workspace(nb, dev), nb, tm_lib.err_flag(dev), stream())
        aux_join()
        return dw1, db1, dw2, db2, None
    if need_dx and wgrad_stream is not None:
        dh = torch.empty(n_rows, hid, dtype=torch.float32, device=dev)
        gemm_nn(n_rows, hid, nout, g, ldg, _f32c(w2), hid, dh, hid, a_rows=g_rows, mask=h, ldmask=hid)
        dx = torch.empty(n_rows, kin, dtype=torch.float32, device=dev)
        gemm_nn(n_rows, kin, hid, dh, hid, _f32c(w1), kin, dx, kin)
        wgrad_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(wgrad_stream):
            gemm_tn(nout, hid, n_rows, g, ldg, h, hid, dw2, hid, a_rows=g_rows, colsum_a=db2)
            gemm_tn(hid, kin, n_rows, dh, hid, x, ldx, dw1, kin, b_rows=rows, colsum_a=db1)
            aux_join()
        return dw1, db1, dw2, db2, dx
    gemm_tn(nout, hid, n_rows, g, ldg, h, hid, dw2, hid, a_rows=g_rows, b_rows=h_rows, colsum_a=db2)
    if h_rows is not None:
        if need_dx:
            raise RuntimeError("mlp2_backward: h_rows has no dx path")
        # dh lives in the rows of h it belongs to, so the ReLU mask of the epilogue reads the matching row
        dh = torch.empty(h.shape[0], hid, dtype=torch.float32, device=dev)
        _hidden_grad(n_rows, hid, nout, g, ldg, g_rows, w2, h, h_rows, dh)
        gemm_tn(hid, kin, n_rows, dh, hid, x, ldx, dw1, kin, a_rows=h_rows, b_rows=rows, colsum_a=db1)
        aux_join()
        return dw1, db1, dw2, db2, None
    if kin <= 2 and hid <= 256 and hid % 4 == 0 and not need_dx and _precision() == 3:
        # Linear(<=2, hid) first layer: its whole backward (dW1, db1) is a reduction of the hidden
        # gradient, fused into the data-gradient GEMM's epilogue -- dh is never written or re-read
        nb = tm_lib.ws_bytes("tm_tc_mlp1_bwd_ws", hid)
        call("tm_tc_mlp1_bwd_fused", n_rows, hid, nout, g, ldg, g_rows, transpose(_f32c(w2)), h, hid, x, ldx, rows, kin,
             None, None, dw1, db1, tm_lib.workspace(nb, dev), nb, tm_lib.err_flag(dev), stream())
        aux_join()
        return dw1, db1, dw2, db2, None
    dh = torch.empty(n_rows, hid, dtype=torch.float32, device=dev)
    _hidden_grad(n_rows, hid, nout, g, ldg, g_rows, w2, h, None, dh)
    gemm_tn(hid, kin, n_rows, dh, hid, x, ldx, dw1, kin, b_rows=rows, colsum_a=db1)
    dx = None
    if need_dx:
        dx = torch.empty(n_rows, kin, dtype=torch.float32, device=dev)
        gemm_nn(n_rows, kin, hid, dh, hid, _f32c(w1), kin, dx, kin)
    aux_join()
    return dw1, db1, dw2, db2, dx


# --------------------------------------------------------------------------------------------
# GNN propagation
# --------------------------------------------------------------------------------------------
GNN_PARAM_NAMES = ("fc_cell_self.layers.0.weight", "fc_cell_self.layers.0.bias",
                   "fc_cell_self.layers.2.weight", "fc_cell_self.layers.2.bias",
                   "fc_net_self.layers.0.weight", "fc_net_self.layers.0.bias",
                   "fc_net_self.layers.2.weight", "fc_net_self.layers.2.bias",
                   "fc_cell_neigh.layers.0.weight", "fc_cell_neigh.layers.0.bias",
                   "fc_cell_neigh.layers.2.weight", "fc_cell_neigh.layers.2.bias")


def gnn_forward(sched, cell_feat, net_feat, params, save=True, impl=None, x_rows=None):
    """Full propagation over every level.  ``params``: the 12 tensors of GNN_PARAM_NAMES.
    Returns ``(H, saved)``; H is (N, 128) with zeros on pins outside the schedule.
    ``impl``: back end for this call (``tm_gnn_set_impl`` bits; None = the process default, which runs the
    forward as ONE persistent cluster kernel).  The persistent kernel owns every SM while it runs, so a
    caller that overlaps the propagation with other streams (``DesignStep``) asks for 0, the chain of
    small per-level kernels that co-resides with them.
    ``x_rows`` = (rows of ``cell_feat`` for ``sched.cell_class``, rows of ``net_feat`` for ``sched.net_class``) when the
    schedule is a sub-netlist whose pins are renumbered (``tm_graph.ConeGraph``); default: the pin ids themselves."""
    (cs1w, cs1b, cs2w, cs2b, ns1w, ns1b, ns2w, ns2b, cn1w, cn1b, cn2w, cn2b) = [_f32c(p) for p in params]
    if cn2w.shape[0] != D or cs2w.shape[0] != D or ns2w.shape[0] != D or cn1w.shape != (256, D):
        raise RuntimeError("the CUDA propagation kernels are built for out_feat_dim = hidden_feat_dim = 128 "
                           "with 256-wide MLPs (the reference configuration, options.py:10)")
    cell_feat, net_feat = _rowmajor(cell_feat), _rowmajor(net_feat)
    tm_lib.require_cuda(cell_feat, "cell_feat")
    dev = cell_feat.device
    n = sched.n
    S = torch.empty(n, D, dtype=torch.float32, device=dev)
    nc, nn_ = int(sched.cell_class.numel()), int(sched.net_class.numel())
    xc, xn = x_rows if x_rows is not None else (sched.cell_class, sched.net_class)
    hc = mlp2_forward(cell_feat, cell_feat.stride(0), xc, nc, cs1w, cs1b, cs2w, cs2b, S, D,
                      out_rows=sched.cell_class, math=_recurrence_math())
    hn = mlp2_forward(net_feat, net_feat.stride(0), xn, nn_, ns1w, ns1b, ns2w, ns2b, S, D,
                      out_rows=sched.net_class, math=_recurrence_math())
    # every scheduled pin's row is written by its level's kernel: only pins outside the schedule need the zero fill
    # (170 MB = 42 us of the config-2 step when every pin is scheduled, as in the reference's designs)
    if sched.n_sched == n and os.environ.get("TM_H_FILL") != "1":
        H = torch.empty(n, D, dtype=torch.float32, device=dev)
    else:
        H = torch.zeros(n, D, dtype=torch.float32, device=dev)
    ncr = sched.n_cell_rows
    A = LSE = HID = None
    if save:
        A = torch.empty(max(ncr, 1), D, dtype=torch.float32, device=dev)
        LSE = torch.empty(max(ncr, 1), D, dtype=torch.float32, device=dev)
        HID = torch.empty(max(ncr, 1), 256, dtype=torch.float32, device=dev)
    w1t, w2t = transpose(cn1w), transpose(cn2w)
    nb = tm_lib.ws_bytes("tm_gnn_ws_bytes")
    old = tm_lib.lib().tm_gnn_set_impl(impl) if impl is not None else None
    try:
        call("tm_gnn_forward", sched.struct, 0, sched.num_levels, H, S, w1t, cn1b, w2t, cn2b, A, LSE, HID,
             tm_lib.workspace(nb, dev), nb, stream())
    finally:
        if old is not None:
            tm_lib.lib().tm_gnn_set_impl(old)
    saved = dict(H=H, A=A, LSE=LSE, HID=HID, hc=hc, hn=hn, cell_feat=cell_feat, net_feat=net_feat, x_rows=x_rows) if save else None
    return H, saved


def gnn_backward(sched, saved, params, G, cone=None):
    """``G`` (N,128): dLoss/dH, consumed (overwritten with dLoss/d pre-activation).
    Returns the 12 parameter gradients in GNN_PARAM_NAMES order.
    ``cone`` (``tm_graph.BackwardCone`` of the endpoints that seeded G): the weight-gradient contractions run over
    the cone's rows only -- every other row of dLoss/dz is exactly zero."""
    (cs1w, cs1b, cs2w, cs2b, ns1w, ns1b, ns2w, ns2b, cn1w, cn1b, cn2w, cn2b) = [_f32c(p) for p in params]
    dev = G.device
    ncr = sched.n_cell_rows
    GA = torch.empty(max(ncr, 1), D, dtype=torch.float32, device=dev)
    GHID = torch.empty(max(ncr, 1), 256, dtype=torch.float32, device=dev)
    GZC = torch.empty(max(ncr, 1), D, dtype=torch.float32, device=dev)
    nb = tm_lib.ws_bytes("tm_gnn_ws_bytes")
    call("tm_gnn_backward", sched.struct, saved["H"], G, cn1w, cn2w, saved["A"], saved["LSE"], saved["HID"],
         GA, GHID, GZC, tm_lib.workspace(nb, dev), nb, stream())
    dcn2w = torch.empty(D, 256, dtype=torch.float32, device=dev)
    dcn2b = torch.empty(D, dtype=torch.float32, device=dev)
    dcn1w = torch.empty(256, D, dtype=torch.float32, device=dev)
    dcn1b = torch.empty(256, dtype=torch.float32, device=dev)
    cr = cone.crows if cone is not None else None
    if cr is not None:
        ncr = int(cr.numel())
    if ncr > 0:
        gemm_tn(D, 256, ncr, GZC, D, saved["HID"], 256, dcn2w, 256, a_rows=cr, b_rows=cr, colsum_a=dcn2b)
        gemm_tn(256, D, ncr, GHID, 256, saved["A"], D, dcn1w, D, a_rows=cr, b_rows=cr, colsum_a=dcn1b)
        aux_join()
    else:
        for t in (dcn2w, dcn2b, dcn1w, dcn1b):
            t.zero_()
    cf, nf = saved["cell_feat"], saved["net_feat"]
    cpins, npins, cpos = sched.cell_class, sched.net_class, None
    xc, xn = saved.get("x_rows") or (cpins, npins)            # feature rows of the two classes
    if cone is not None:
        if saved.get("x_rows") is not None:
            raise RuntimeError("gnn_backward: a BackwardCone on top of a renumbered sub-netlist is not supported")
        cpins, npins = cone.cell_pins, cone.net_pins
        xc, xn = cpins, npins
        cpos = cone.cell_pos if saved["hc"] is not None else None
    nc, nn_ = int(cpins.numel()), int(npins.numel())
    dcs1w, dcs1b, dcs2w, dcs2b, _ = mlp2_backward(cf, cf.stride(0), xc, nc, cs1w, cs2w,
                                                  saved["hc"], G, D, g_rows=cpins, b1=cs1b, h_rows=cpos)
    dns1w, dns1b, dns2w, dns2b, _ = mlp2_backward(nf, nf.stride(0), xn, nn_, ns1w, ns2w,
                                                  saved["hn"], G, D, g_rows=npins, b1=ns1b,
                                                  h_rows=(cone.net_pos if cone is not None and saved["hn"] is not None else None))
    return (dcs1w, dcs1b, dcs2w, dcs2b, dns1w, dns1b, dns2w, dns2b, dcn1w, dcn1b, dcn2w, dcn2b)


class GnnPropagateFn(torch.autograd.Function):
    """H = propagate(cell_feat, net_feat; 12 PathConv parameters) over a fixed Schedule."""

    @staticmethod
    def forward(ctx, sched, cell_feat, net_feat, need, *params):
        H, saved = gnn_forward(sched, cell_feat, net_feat, [p.detach() for p in params], save=need)
        ctx.sched, ctx.saved = sched, saved
        ctx.params = [p.detach() for p in params]
        return H

    @staticmethod
    def backward(ctx, gH):
        G = gH.contiguous().clone()                      # consumed in place by the sweep
        grads = gnn_backward(ctx.sched, ctx.saved, ctx.params, G)
        return (None, None, None, None) + tuple(grads)


def step_metrics(pred, arrival, required=None, label=None):
    """One launch, one (8,) device tensor [mse, r2, correct, tp, fn, tn, fp, T] instead of the seven
    ``.item()`` round trips of train.py:513-549.  ``label``: int64 (0 = non-critical)."""
    pred, arrival = _f32c(pred.detach().reshape(-1)), _f32c(arrival.detach().reshape(-1))
    tm_lib.require_cuda(pred, "pred")
    out = torch.empty(8, dtype=torch.float32, device=pred.device)
    req = None if required is None else _f32c(required.detach().reshape(-1))
    lab = None if label is None else label.detach().reshape(-1).to(torch.int64).contiguous()
    call("tm_step_metrics", int(pred.numel()), pred, arrival, req, lab, out, stream())
    return out


# --------------------------------------------------------------------------------------------
# mask fusion
# --------------------------------------------------------------------------------------------
def fusion_forward(mask_rows, feat, fcn_w, fcn_b, out, ldo):
    """out[t, 0:128] = fcn(mask_t * feat)  (train.py:501 + model.py:272).  Returns fcn_w^T."""
    J = int(feat.numel())
    if fcn_w.shape != (D, J):
        raise RuntimeError(f"fcn must be Linear({J}, 128); got weight {tuple(fcn_w.shape)}")
    wt = transpose(fcn_w)
    if FUSE_RUNS:
        run_ptr, run_lo, run_hi = mask_rows.runs()
        nb = tm_lib.ws_bytes("tm_fuse_runs_ws", J)
        call("tm_fuse_forward_runs", mask_rows.T, J, D, run_ptr, run_lo, run_hi, feat, wt, _f32c(fcn_b), out, ldo,
             tm_lib.workspace(nb, feat.device), nb, stream())
    else:
        csr = mask_rows.csr
        call("tm_fuse_forward", mask_rows.T, J, D, csr.indptr, csr.cols, mask_rows.rows, feat, wt, _f32c(fcn_b),
             out, ldo, stream())
    return wt


def fusion_backward(mask_rows, feat, wt, g, ldg):
    """Returns (dfeat (J,), dfcn_w (128,J), dfcn_b (128,))."""
    J = int(feat.numel())
    dev = feat.device
    cptr, ct = mask_rows.csc()
    dwt = torch.empty(J, D, dtype=torch.float32, device=dev)
    dF = torch.empty(J, dtype=torch.float32, device=dev)
    call("tm_fuse_backward", mask_rows.T, J, D, cptr, ct, g, ldg, feat, wt, dwt, dF, stream())
    db = colsum(g, mask_rows.T, D, ldg)
    return dF, transpose(dwt), db


class MaskFusionFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mask_rows, feat_flat, fcn_w, fcn_b):
        feat = _f32c(feat_flat.detach()).reshape(-1)
        out = torch.empty(mask_rows.T, D, dtype=torch.float32, device=feat.device)
        wt = fusion_forward(mask_rows, feat, fcn_w.detach(), fcn_b.detach(), out, D)
        ctx.mask_rows, ctx.feat, ctx.wt, ctx.shape = mask_rows, feat, wt, feat_flat.shape
        return out

    @staticmethod
    def backward(ctx, g):
        g = _f32c(g)
        dF, dw, db = fusion_backward(ctx.mask_rows, ctx.feat, ctx.wt, g, D)
        return None, dF.reshape(ctx.shape), dw, db


class MaskedFeatureMap:
    """What the caller may pass as ``path_map`` instead of the dense (T, map^2) product of
    train.py:500-501: the selected sparse mask rows plus the flattened CNN feature map."""

    def __init__(self, mask_rows, feat_map):
        self.mask_rows = mask_rows
        self.feat_map = feat_map

    def to_dense(self):
        csr = self.mask_rows.csr
        T, J = self.mask_rows.T, csr.width
        dense = torch.zeros(T, J, dtype=torch.float32, device=self.feat_map.device)
        rl = self.mask_rows.rows.long()
        deg = (csr.indptr[rl + 1] - csr.indptr[rl]).long()
        t_of = torch.repeat_interleave(torch.arange(T, device=deg.device), deg)
        first = torch.cumsum(deg, 0) - deg
        slot = torch.repeat_interleave(csr.indptr[rl].long(), deg) + \
            (torch.arange(int(deg.sum()), device=deg.device) - first[t_of])
        dense[t_of, csr.cols[slot].long()] = 1.0
        return dense * self.feat_map.reshape(1, -1)


# --------------------------------------------------------------------------------------------
# nn.Linear on the fp32 GEMM core (model.py:15; F.linear semantics, optional fused ReLU)
# --------------------------------------------------------------------------------------------
class LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, relu):
        x2 = _rowmajor(x.detach().reshape(-1, x.shape[-1]))
        tm_lib.require_cuda(x2, "Linear input")
        M, K = x2.shape
        N = w.shape[0]
        y = torch.empty(M, N, dtype=torch.float32, device=x2.device)
        gemm_nn(M, N, K, x2, x2.stride(0), _f32c(w.detach()), K, y, N,
                bias=None if b is None else _f32c(b.detach()), flags=RELU if relu else 0, b_is_nk=True)
        ctx.save = (x2, w.detach(), y if relu else None)
        ctx.has_bias, ctx.xshape = b is not None, x.shape
        ctx.need_dx = x.requires_grad
        return y.reshape(x.shape[:-1] + (N,))

    @staticmethod
    def backward(ctx, g):
        x2, w, y = ctx.save
        N, K = w.shape
        M = x2.shape[0]
        g2 = _f32c(g.reshape(M, N))
        if y is not None:                                  # ReLU backward on the incoming gradient
            gm = torch.empty_like(g2)
            call("tm_leaky_relu_backward", g2.numel(), y, g2, 0.0, gm, stream())
            g2 = gm
        dev = x2.device
        dw = torch.empty(N, K, dtype=torch.float32, device=dev)
        db = torch.empty(N, dtype=torch.float32, device=dev) if ctx.has_bias else None
        if M > 0:
            gemm_tn(N, K, M, g2, N, x2, x2.stride(0), dw, K, colsum_a=db)
            aux_join()
        else:
            dw.zero_()
            if db is not None:
                db.zero_()
        dx = None
        if ctx.need_dx:
            dx = torch.empty(M, K, dtype=torch.float32, device=dev)
            gemm_nn(M, K, N, g2, N, _f32c(w), K, dx, K)
            dx = dx.reshape(ctx.xshape)
        return dx, dw, db, None


def linear(x, w, b=None, relu=False):
    return LinearFn.apply(x, w, b, relu)
