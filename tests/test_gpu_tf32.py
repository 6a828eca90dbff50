"""GPU: the TF32 tensor-core conv / wgrad kernels (fp32 storage, tcgen05 kind::tf32; BASELINE.json configs[1]
"fp32/TF32") against fp32 PyTorch statements of the same ops.

  * precision "tf32x3" -- every product as hi*hi + hi*lo + lo*hi on pre-split operands: must meet the fp32 bar
    (4e-5 per op in the max norm -- measured 5e-7 .. 2e-5, the larger figures on 2k-14k term contractions --, 1e-4 end to end: tests/test_gpu_parity.py);
  * precision "tf32"   -- one TF32 MMA per product, the arithmetic cuDNN gives the reference on a GPU: 10-bit
    mantissas, bounded at 2e-3 per op here."""
import pytest
import torch
import torch.nn.functional as F

from tests import parity_log as PL
from tests.test_gpu_ops import DEV, pack, rel

pytestmark = pytest.mark.gpu
TOL = {"tf32x3": 4e-5, "tf32": 2e-3}


@pytest.fixture(params=["tf32x3", "tf32"])
def precision(request):
    import sd_b200
    prev = sd_b200.get_precision()
    sd_b200.set_precision(request.param)
    yield request.param
    sd_b200.set_precision(prev)


@pytest.mark.parametrize("B,T,K,N,taps,dil", [(3, 50, 12, 20, 3, 1), (2, 360, 270, 320, 3, 16), (4, 37, 208, 270, 1, 1),
                                              (2, 130, 320, 640, 3, 2), (1, 24, 9, 10, 3, 8), (2, 200, 640, 1024, 1, 1)])
def test_tf32_conv_fwd_bias_residual(precision, B, T, K, N, taps, dil):
    from sd_b200 import ops
    torch.manual_seed(0)
    x = torch.randn(B, K, T, device=DEV)
    w = torch.randn(N, K, taps, device=DEV) / (K * taps) ** 0.5
    bias = torch.randn(N, device=DEV)
    res = torch.randn(B, N, T, device=DEV)
    xt, rt = ops.nct_to_btc(x, torch.float32), ops.nct_to_btc(res, torch.float32)
    wf, _ = pack(w, torch.float32)
    out = torch.empty((B, T, ops.rup8(N)), dtype=torch.float32, device=DEV)
    ops.conv_fwd(xt, wf, K=K, N=N, taps=taps, dil=dil, bias=bias, res=rt, out=out)
    ref = F.conv1d(x, w, bias, padding=dil * (taps // 2), dilation=dil) + res
    e = rel(ops.btc_to_nct(out, N), ref)
    PL.record("conv_fwd", e, TOL[precision])
    assert e < TOL[precision]
    assert float(out[:, :, N:].abs().max() if ops.rup8(N) > N else 0.0) == 0.0


def test_tf32_conv_fwd_gelu_nct_glu_stats(precision):
    from sd_b200 import ops, _native as nat
    tol = TOL[precision]
    torch.manual_seed(1)
    B, T, K, N = 3, 75, 40, 48
    x = torch.randn(B, K, T, device=DEV)
    xt = ops.nct_to_btc(x, torch.float32)
    w = torch.randn(N, K, 1, device=DEV) / K ** 0.5
    bias = torch.randn(N, device=DEV)
    wf, _ = pack(w, torch.float32)
    Z = torch.empty((B, N, T), dtype=torch.float32, device=DEV)
    zn2 = torch.zeros((B,), dtype=torch.float32, device=DEV)
    pre = torch.empty((B, T, N), dtype=torch.float32, device=DEV)
    ops.conv_fwd(xt, wf, K=K, N=N, bias=bias, out=Z, preact=pre, rownorm2=zn2, act=nat.ACT_GELU, out_mode=nat.OUT_NCT_F32)
    p_ref = F.conv1d(x, w, bias)
    assert rel(Z, F.gelu(p_ref)) < tol
    assert rel(ops.btc_to_nct(pre, N), p_ref) < tol
    assert rel(zn2, (Z * Z).sum(dim=(1, 2))) < 1e-5
    w3 = torch.randn(N, K, 3, device=DEV) / (3 * K) ** 0.5
    wf3, _ = pack(w3, torch.float32)
    y2 = torch.empty((B, T, N), dtype=torch.float32, device=DEV)
    out = torch.empty((B, T, N // 2), dtype=torch.float32, device=DEV)
    ops.conv_fwd(xt, wf3, K=K, N=N, taps=3, dil=2, bias=bias, out=out, preact=y2, act=nat.ACT_GLU)
    y_ref = F.conv1d(x, w3, bias, padding=2, dilation=2)
    assert rel(ops.btc_to_nct(y2, N), y_ref) < tol
    assert rel(ops.btc_to_nct(out, N // 2), F.glu(y_ref, dim=-2)) < tol
    # odd GLU width (D2 = 21: value / gate halves not 16-byte aligned)
    N2 = 42
    w5 = torch.randn(N2, K, 3, device=DEV) / (3 * K) ** 0.5
    b5 = torch.randn(N2, device=DEV)
    wf5, _ = pack(w5, torch.float32)
    y5 = torch.zeros((B, T, ops.rup8(N2)), dtype=torch.float32, device=DEV)
    o5 = torch.empty((B, T, ops.rup8(N2 // 2)), dtype=torch.float32, device=DEV)
    ops.conv_fwd(xt, wf5, K=K, N=N2, taps=3, dil=1, bias=b5, out=o5, preact=y5, act=nat.ACT_GLU)
    y5_ref = F.conv1d(x, w5, b5, padding=1)
    assert rel(ops.btc_to_nct(y5, N2), y5_ref) < tol
    assert rel(ops.btc_to_nct(o5, N2 // 2), F.glu(y5_ref, dim=-2)) < tol
    stats = torch.zeros((2, N), dtype=torch.float64, device=DEV)
    o2 = torch.empty((B, T, N), dtype=torch.float32, device=DEV)
    ops.conv_fwd(xt, wf3, K=K, N=N, taps=3, dil=2, bias=bias, out=o2, stats=stats)
    assert rel(stats[0], o2.sum(dim=(0, 1)).double(), 1.0) < 1e-5
    assert rel(stats[1], (o2 * o2).sum(dim=(0, 1)).double()) < 1e-5


def test_tf32_subject_grouped_conv(precision):
    from sd_b200 import ops
    torch.manual_seed(2)
    B, T, D, S = 9, 40, 30, 5
    x = torch.randn(B, D, T, device=DEV)
    ws = torch.randn(S, D, D, 1, device=DEV) / D ** 0.5
    ids = torch.tensor([0, 3, 3, 1, 0, 4, 4, 4, 1], dtype=torch.int32, device=DEV)
    Dp = ops.rup8(D)
    wf = torch.empty((S, 1, Dp, Dp), dtype=torch.float32, device=DEV)
    for s in range(S):
        wf[s] = pack(ws[s].contiguous(), torch.float32)[0][0]
    xt = ops.nct_to_btc(x, torch.float32)
    out = torch.empty_like(xt)
    ops.conv_fwd(xt, wf, K=D, N=D, widx=ids, G=S, out=out)
    assert rel(ops.btc_to_nct(out, D), torch.bmm(ws[ids.long(), :, :, 0], x)) < TOL[precision]


@pytest.mark.parametrize("B,T,K,N,taps,dil", [(3, 50, 12, 20, 3, 1), (2, 360, 270, 320, 3, 16), (5, 64, 40, 48, 1, 1),
                                              (2, 100, 320, 640, 3, 2), (40, 360, 320, 320, 3, 4)])
def test_tf32_dgrad_wgrad_match_autograd(precision, B, T, K, N, taps, dil):
    from sd_b200 import ops
    tol = TOL[precision]
    torch.manual_seed(3)
    x = torch.randn(B, K, T, device=DEV, requires_grad=True)
    w = (torch.randn(N, K, taps, device=DEV) / (K * taps) ** 0.5).requires_grad_(True)
    bias = torch.randn(N, device=DEV, requires_grad=True)
    dy = torch.randn(B, N, T, device=DEV)
    F.conv1d(x, w, bias, padding=dil * (taps // 2), dilation=dil).backward(dy)
    xt, dyt = ops.nct_to_btc(x.detach(), torch.float32), ops.nct_to_btc(dy, torch.float32)
    _, wd = pack(w.detach(), torch.float32)
    dx = torch.empty_like(xt)
    ops.conv_fwd(dyt, wd, K=N, N=K, taps=taps, dil=dil, out=dx)
    e = rel(ops.btc_to_nct(dx, K), x.grad)
    PL.record("dgrad", e, tol)
    assert e < tol
    dw = torch.zeros_like(w.detach())
    db = torch.zeros(N, device=DEV)
    ops.conv_wgrad(dyt, xt, dw, K=K, N=N, taps=taps, dil=dil, dbias=db)
    e = rel(dw, w.grad)
    PL.record("wgrad", e, tol)
    assert e < tol
    assert rel(db, bias.grad) < tol


def test_tf32_grouped_wgrad(precision):
    from sd_b200 import ops
    torch.manual_seed(4)
    B, T, D, S = 11, 48, 30, 4
    x = torch.randn(B, D, T, device=DEV)
    dy = torch.randn(B, D, T, device=DEV)
    ids = torch.tensor([2, 0, 0, 3, 2, 2, 0, 3, 3, 3, 0])
    order = torch.argsort(ids, stable=True).int().to(DEV)
    offsets = torch.cat([torch.zeros(1, dtype=torch.long), torch.cumsum(torch.bincount(ids, minlength=S), 0)]).int().to(DEV)
    xt, dyt = ops.nct_to_btc(x, torch.float32), ops.nct_to_btc(dy, torch.float32)
    dws = torch.zeros((S, D, D, 1), device=DEV)
    ops.conv_wgrad(dyt, xt, dws, K=D, N=D, order=order, offsets=offsets, G=S, strides=(D * D, D, 1, 0))
    for s in range(S):
        sel = (ids == s).nonzero().flatten().to(DEV)
        ref = torch.einsum("bnt,bkt->nk", dy[sel], x[sel]) if len(sel) else torch.zeros(D, D, device=DEV)
        assert rel(dws[s, :, :, 0], ref, 1.0) < TOL[precision], s
