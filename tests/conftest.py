import importlib
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_NAME = "multimodal-fusion-based-pre-routing-timing-prediction-_b200"
PKG = os.path.join(ROOT, PKG_NAME)
GOLD = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def pkg():
    return importlib.import_module(PKG_NAME)


def design_to_oracle(d):
    """SynthDesign -> the dict of CPU tensors oracle.restate.design_step consumes."""
    from oracle import levelize
    ni, ns = levelize.in_csr(d.n, d.net_src, d.net_dst)
    ci, cs = levelize.in_csr(d.n, d.cell_src, d.cell_dst)
    t = torch.from_numpy
    return dict(n=d.n, levels=[t(x.astype(np.int64)) for x in d.level_lists()],
                net_csr=(t(ni).long(), t(ns).long()), cell_csr=(t(ci).long(), t(cs).long()),
                cell_feat=t(d.cell_feat), net_feat=t(d.net_feat), image=t(d.image),
                endpoints=t(d.endpoints), endpoint_level=t(d.level[d.endpoints].astype(np.int64)),
                mask_indptr=t(d.mask_indptr).long(), mask_cols=t(d.mask_cols).long(),
                arrival_time=t(d.arrival_time))


def load_golden_step(name="tiny"):
    z = np.load(os.path.join(GOLD, f"step_{name}.npz"))
    sd_m = {k[6:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("model.")}
    sd_c = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("cnn.")}
    return z, sd_m, sd_c


def assert_close(a, b, rtol=1e-3, atol_scale=1e-4, name="", flip_frac=0.0, flip_factor=50.0, max_bad=0):
    """rtol plus an atol scaled to the tensor's magnitude (SURVEY.md 8d parity gates).

    ``flip_frac`` > 0 is for GRADIENTS that pass through millions of ReLU gates: two correct fp32
    evaluations with different summation order disagree about the sign of the few pre-activations
    that sit within rounding noise of zero (expected count ~ #gates x 2^-22), and each such gate
    switches one gradient row (one element of a bias gradient) on or off.  Up to that fraction of
    elements -- at least two, for the small bias vectors -- may exceed the tolerance, by at most
    ``flip_factor`` x; everything else must meet it.  Forward values never use it."""
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    assert a.shape == b.shape, f"{name}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    scale = float(b.abs().max()) if b.numel() else 0.0
    atol = atol_scale * max(scale, 1e-30)
    err = (a - b).abs()
    tol = atol + rtol * b.abs()
    bad = err > tol
    nbad = int(bad.sum())
    msg = (f"{name}: {nbad}/{a.numel()} mismatches, max abs err {float(err.max()) if a.numel() else 0:.3e}, "
           f"ref scale {scale:.3e}")
    if max_bad > 0:        # an ABSOLUTE number of elements (hidden-layer gate ties at full size), bounded by flip_factor
        assert nbad <= max_bad and not bool((err > flip_factor * tol).any()), msg
    elif flip_frac > 0.0:
        assert nbad <= max(2, flip_frac * a.numel()) and not bool((err > flip_factor * tol).any()), msg
    else:
        assert nbad == 0, msg


def relu_gate_flips(H, Href, near=1e-5, max_flips=16):
    """Evidence instead of a blanket allowance for the GNN gradient checks.  A pin's output gate relu(z) is the only
    discontinuity between ``H`` and the parameter gradients that involves a pre-activation of size ~0: two correct
    evaluations whose z differ in the last bits may disagree about the sign of a z that is ~0, and that switches
    one element of g_z on or off.  Returns the set of channels c for which some pin has ``(H>0) != (Href>0)``,
    after checking that EVERY such disagreement sits at a pre-activation provably within ``near`` x max|H| of zero
    and that there are at most ``max_flips`` of them (measured: 0 on most seeds, 1-2 of 2.1 M gates on config 1)."""
    H = torch.as_tensor(H).detach().double().cpu()
    Href = torch.as_tensor(Href).detach().double().cpu()
    dis = (H > 0) != (Href > 0)
    n = int(dis.sum())
    if n == 0:
        return set()
    scale = float(Href.abs().max())
    worst = float(torch.maximum(H.abs(), Href.abs())[dis].max())
    assert n <= max_flips, f"{n} ReLU gates disagree with the oracle"
    assert worst <= near * scale, f"a disagreeing gate is not near zero: |h| = {worst:.3e} (scale {scale:.3e})"
    return set(int(c) for c in torch.nonzero(dis)[:, 1].tolist())


def assert_grad_close_given_flips(a, b, name, channels, rtol=1e-3, atol_scale=1e-4, flip_factor=50.0):
    """Strict ``assert_close`` everywhere except where a flipped output gate of channel c lands directly: row c of
    a ``layers.2.weight`` (128 x 256) and element c of a ``layers.2.bias`` -- those may be off by the one switched
    g_z element (bounded by ``flip_factor`` x tolerance).  No flips -> no exception anywhere."""
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    assert a.shape == b.shape, f"{name}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    scale = float(b.abs().max()) if b.numel() else 0.0
    tol = atol_scale * max(scale, 1e-30) + rtol * b.abs()
    err = (a - b).abs()
    strict = torch.ones_like(err, dtype=torch.bool)
    if channels and (name.endswith("layers.2.weight") or name.endswith("layers.2.bias")) and a.shape[0] == 128:
        strict[sorted(channels)] = False
    bad = (err > tol) & strict
    assert int(bad.sum()) == 0, (f"{name}: {int(bad.sum())}/{a.numel()} mismatches outside flipped-gate rows, "
                                 f"max abs err {float(err[strict].max()):.3e}, ref scale {scale:.3e}")
    assert not bool((err > flip_factor * tol).any()), f"{name}: an element is off by more than {flip_factor} x tolerance"
