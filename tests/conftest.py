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


def assert_close(a, b, rtol=1e-3, atol_scale=1e-4, name="", flip_frac=0.0, flip_factor=50.0):
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
    if flip_frac > 0.0:
        assert nbad <= max(2, flip_frac * a.numel()) and not bool((err > flip_factor * tol).any()), msg
    else:
        assert nbad == 0, msg
