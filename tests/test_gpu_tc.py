"""tcgen05 / TMEM kernels (tm_tc_*) against fp64 PyTorch references.

precision 3 (3xTF32, the default) and 2 (bf16 x6) must meet the fp32 bar (rtol 1e-3; checked much tighter here);
precision 1 (split bf16 x3) rtol 1e-3; precision 0 (plain bf16) and 4 (single TF32) the bf16 bar (rtol 2e-2)."""
import pytest
import torch
import torch.nn.functional as F

from conftest import assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = {3: (1e-4, 2e-5), 2: (1e-4, 1e-5), 1: (1e-3, 1e-4), 0: (2e-2, 1e-2), 4: (2e-2, 1e-2)}


@pytest.fixture(scope="module")
def lib(pkg):
    import tm_lib
    return tm_lib


def _err():
    return torch.zeros(1, dtype=torch.int32, device=DEV)


@pytest.mark.parametrize("precision", [3, 4, 2, 1, 0])
@pytest.mark.parametrize("M,N,K,b_is_nk", [(128, 32, 64, 1), (300, 256, 36, 1), (1000, 128, 256, 1), (257, 100, 130, 0),
                                           (1350, 576, 288, 1), (77, 1, 576, 1), (500, 256, 2, 1)])
def test_tc_gemm_nn(lib, M, N, K, b_is_nk, precision):
    torch.manual_seed(M + N + K)
    A = torch.randn(M + 3, K, device=DEV)
    B = torch.randn(N, K, device=DEV) if b_is_nk else torch.randn(K, N, device=DEV)
    bias = torch.randn(N, device=DEV)
    rows = torch.randperm(M + 3, device=DEV)[:M].to(torch.int32)
    C = torch.full((M + 3, N), 7.0, device=DEV)
    err = _err()
    lib.call("tm_tc_gemm_nn", M, N, K, A, K, rows, B, B.shape[1], b_is_nk, C, N, rows, bias, None, 0, 1 | 2,
             precision, err, lib.stream())
    Bm = B.double().t() if b_is_nk else B.double()
    ref = torch.relu(A[rows.long()].double() @ Bm + bias.double())
    assert int(err.item()) == 0
    assert_close(C[rows.long()], ref, *TOL[precision], "tc_gemm_nn")
    untouched = torch.ones(M + 3, dtype=torch.bool, device=DEV)
    untouched[rows.long()] = False
    assert bool((C[untouched] == 7.0).all())


@pytest.mark.parametrize("precision", [3, 4, 2, 1, 0])
@pytest.mark.parametrize("M,N,R", [(128, 256, 5000), (256, 36, 3001), (1, 576, 1350), (27, 16, 4096), (256, 128, 20000)])
def test_tc_gemm_tn(lib, M, N, R, precision):
    torch.manual_seed(M + N + R)
    A = torch.randn(R + 3, M, device=DEV)
    B = torch.randn(R + 3, N, device=DEV)
    rows = torch.randperm(R + 3, device=DEV)[:R].to(torch.int32)
    C = torch.empty(M, N, device=DEV)
    nb = lib.ws_bytes("tm_tc_gemm_tn_ws", M, N, R)
    ws = lib.workspace(nb, DEV)
    err = _err()
    lib.call("tm_tc_gemm_tn", M, N, R, A, M, rows, B, N, rows, C, N, 0, precision, ws, nb, err, lib.stream())
    ref = A[rows.long()].double().t() @ B[rows.long()].double()
    assert int(err.item()) == 0
    assert_close(C, ref, *TOL[precision], "tc_gemm_tn")


@pytest.mark.parametrize("precision", [3, 4, 2, 1, 0])
@pytest.mark.parametrize("B,H,W,Cin,Cout,k", [(1, 16, 16, 3, 16, 3), (2, 8, 24, 16, 32, 3), (1, 16, 8, 128, 64, 3),
                                              (1, 12, 12, 2, 32, 9), (1, 16, 16, 16, 1, 1), (1, 32, 32, 64, 128, 3)])
def test_tc_conv(lib, B, H, W, Cin, Cout, k, precision):
    torch.manual_seed(B * H + Cin + Cout + k)
    x = torch.randn(B, Cin, H, W, device=DEV, requires_grad=True)
    w = (torch.randn(Cout, Cin, k, k, device=DEV) * 0.1).requires_grad_(True)
    bias = torch.randn(Cout, device=DEV)
    ref = F.conv2d(x.double(), w.double(), bias.double(), padding=k // 2)
    g = torch.randn(B, Cout, H, W, device=DEV)
    gx, gw = torch.autograd.grad(ref, (x, w), g.double())
    nhwc = lambda t: t.detach().permute(0, 2, 3, 1).contiguous()            # noqa: E731
    xs, gs = nhwc(x), nhwc(g)
    wf = torch.empty(k * k * Cin, Cout, device=DEV)
    wb = torch.empty(k * k * Cout, Cin, device=DEV)
    lib.call("tm_conv_pack_weight", Cout, Cin, k, w.detach().contiguous(), wf, wb, lib.stream())
    err = _err()
    y = torch.empty(B, H, W, Cout, device=DEV)
    lib.call("tm_tc_conv2d_nhwc", B, H, W, Cin, Cout, k, xs, Cin, wf, bias, y, Cout, 0, precision, err, lib.stream())
    assert_close(y.permute(0, 3, 1, 2), ref, *TOL[precision], "tc fprop")
    dx = torch.empty(B, H, W, Cin, device=DEV)
    lib.call("tm_tc_conv2d_nhwc", B, H, W, Cout, Cin, k, gs, Cout, wb, None, dx, Cin, 0, precision, err, lib.stream())
    assert_close(dx.permute(0, 3, 1, 2), gx, *TOL[precision], "tc dgrad")
    nb = lib.ws_bytes("tm_tc_conv2d_wgrad_ws", B, H, W, Cin, Cout, k)
    dwf = torch.empty(k * k * Cin, Cout, device=DEV)
    lib.call("tm_tc_conv2d_wgrad_nhwc", B, H, W, Cin, Cout, k, xs, Cin, gs, Cout, dwf, precision,
             lib.workspace(nb, DEV), nb, err, lib.stream())
    dw = torch.empty(Cout, Cin, k, k, device=DEV)
    lib.call("tm_conv_unpack_wgrad", Cout, Cin, k, dwf, dw, lib.stream())
    assert_close(dw, gw, *TOL[precision], "tc wgrad")
    assert int(err.item()) == 0
