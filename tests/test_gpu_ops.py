"""GPU: each C-ABI op against a plain fp32 PyTorch statement of the same math (per-op parity)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _ops():
    from sd_b200 import ops, _native
    return ops, _native


def rel(a, b, floor=1e-30):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max()) / max(float(b.abs().max()), floor)


def pack(w, dtype):
    """(N,K,taps) fp32 -> wf (1,taps,Np,Kp), wd (1,taps,Kp,Np) through the library."""
    ops, nat = _ops()
    N, K, taps = w.shape
    Np, Kp = ops.rup8(N), ops.rup8(K)
    wf = torch.empty((1, taps, Np, Kp), dtype=dtype, device=DEV)
    wd = torch.empty((1, taps, Kp, Np), dtype=dtype, device=DEV)
    nat.call("sd_pack_weight", w.data_ptr(), wf.data_ptr(), wd.data_ptr(), N, K, taps, Np, Kp, ops.code_of(wf),
             torch.cuda.current_stream().cuda_stream)
    return wf, wd


TOL = {torch.float32: 2e-5, torch.bfloat16: 2e-2}


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("impl", ["simt", "auto"])
@pytest.mark.parametrize("B,T,K,N,taps,dil", [(3, 50, 12, 20, 3, 1), (2, 360, 270, 320, 3, 16), (4, 37, 208, 270, 1, 1),
                                              (2, 130, 320, 640, 3, 2), (1, 24, 9, 10, 3, 8)])
def test_conv_fwd_bias_residual(dtype, impl, B, T, K, N, taps, dil):
    ops, nat = _ops()
    ops.set_impl(impl)
    try:
        torch.manual_seed(0)
        x = torch.randn(B, K, T, device=DEV)
        w = torch.randn(N, K, taps, device=DEV) / (K * taps) ** 0.5
        bias = torch.randn(N, device=DEV)
        res = torch.randn(B, N, T, device=DEV)
        xt, rt = ops.nct_to_btc(x, dtype), ops.nct_to_btc(res, dtype)
        wf, _ = pack(w, dtype)
        out = torch.empty((B, T, ops.rup8(N)), dtype=dtype, device=DEV)
        ops.conv_fwd(xt, wf, K=K, N=N, taps=taps, dil=dil, bias=bias, res=rt, out=out)
        ref = F.conv1d(x, w, bias, padding=dil * (taps // 2), dilation=dil) + res
        assert rel(ops.btc_to_nct(out, N), ref) < TOL[dtype]
        assert float(out[:, :, N:].float().abs().max() if ops.rup8(N) > N else 0.0) == 0.0
    finally:
        ops.set_impl("auto")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("impl", ["simt", "auto"])
def test_conv_fwd_gelu_nct_glu_stats(dtype, impl):
    ops, nat = _ops()
    ops.set_impl(impl)
    try:
        torch.manual_seed(1)
        B, T, K, N = 3, 75, 40, 48
        x = torch.randn(B, K, T, device=DEV)
        xt = ops.nct_to_btc(x, dtype)
        w = torch.randn(N, K, 1, device=DEV) / K ** 0.5
        bias = torch.randn(N, device=DEV)
        wf, _ = pack(w, dtype)
        # GELU + NCT fp32 output + saved pre-activation
        Z = torch.empty((B, N, T), dtype=torch.float32, device=DEV)
        pre = torch.empty((B, T, N), dtype=dtype, device=DEV)
        ops.conv_fwd(xt, wf, K=K, N=N, bias=bias, out=Z, preact=pre, act=nat.ACT_GELU, out_mode=nat.OUT_NCT_F32)
        p_ref = F.conv1d(x, w, bias)
        assert rel(Z, F.gelu(p_ref)) < TOL[dtype]
        assert rel(ops.btc_to_nct(pre, N), p_ref) < TOL[dtype]
        # GLU (k=3, dil=2) + saved y2
        w3 = torch.randn(N, K, 3, device=DEV) / (3 * K) ** 0.5
        wf3, _ = pack(w3, dtype)
        y2 = torch.empty((B, T, N), dtype=dtype, device=DEV)
        out = torch.empty((B, T, N // 2), dtype=dtype, device=DEV)
        ops.conv_fwd(xt, wf3, K=K, N=N, taps=3, dil=2, bias=bias, out=out, preact=y2, act=nat.ACT_GLU)
        y_ref = F.conv1d(x, w3, bias, padding=2, dilation=2)
        assert rel(ops.btc_to_nct(y2, N), y_ref) < TOL[dtype]
        assert rel(ops.btc_to_nct(out, N // 2), F.glu(y_ref, dim=-2)) < TOL[dtype]
        # batch statistics of the stored output
        stats = torch.zeros((2, N), dtype=torch.float64, device=DEV)
        o2 = torch.empty((B, T, N), dtype=dtype, device=DEV)
        ops.conv_fwd(xt, wf3, K=K, N=N, taps=3, dil=2, bias=bias, out=o2, stats=stats)
        of = o2.float()
        assert rel(stats[0], of.sum(dim=(0, 1)).double(), 1.0) < 1e-4
        assert rel(stats[1], (of * of).sum(dim=(0, 1)).double()) < 1e-4
    finally:
        ops.set_impl("auto")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("impl", ["simt", "auto"])
def test_subject_grouped_conv_and_dgrad(dtype, impl):
    ops, nat = _ops()
    ops.set_impl(impl)
    try:
        torch.manual_seed(2)
        B, T, D, S = 9, 40, 30, 5
        x = torch.randn(B, D, T, device=DEV)
        ws = torch.randn(S, D, D, 1, device=DEV) / D ** 0.5
        ids = torch.tensor([0, 3, 3, 1, 0, 4, 4, 4, 1], dtype=torch.int32, device=DEV)
        Dp = ops.rup8(D)
        wf = torch.empty((S, 1, Dp, Dp), dtype=dtype, device=DEV)
        wd = torch.empty((S, 1, Dp, Dp), dtype=dtype, device=DEV)
        for s in range(S):
            f, d = pack(ws[s].contiguous(), dtype)
            wf[s], wd[s] = f[0], d[0]
        xt = ops.nct_to_btc(x, dtype)
        out = torch.empty_like(xt)
        ops.conv_fwd(xt, wf, K=D, N=D, widx=ids, G=S, out=out)
        ref = torch.bmm(ws[ids.long(), :, :, 0], x)
        assert rel(ops.btc_to_nct(out, D), ref) < TOL[dtype]
        dx = torch.empty_like(xt)
        ops.conv_fwd(xt, wd, K=D, N=D, widx=ids, G=S, out=dx)     # dgrad == W^T
        ref_t = torch.bmm(ws[ids.long(), :, :, 0].transpose(1, 2), x)
        assert rel(ops.btc_to_nct(dx, D), ref_t) < TOL[dtype]
    finally:
        ops.set_impl("auto")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("impl", ["simt", "auto"])
@pytest.mark.parametrize("B,T,K,N,taps,dil", [(3, 50, 12, 20, 3, 1), (2, 360, 270, 320, 3, 16), (5, 64, 40, 48, 1, 1),
                                              (2, 100, 320, 640, 3, 2)])
def test_conv_dgrad_wgrad_match_autograd(dtype, impl, B, T, K, N, taps, dil):
    ops, nat = _ops()
    ops.set_impl(impl)
    try:
        torch.manual_seed(3)
        x = torch.randn(B, K, T, device=DEV, requires_grad=True)
        w = (torch.randn(N, K, taps, device=DEV) / (K * taps) ** 0.5).requires_grad_(True)
        b = torch.randn(N, device=DEV, requires_grad=True)
        dy = torch.randn(B, N, T, device=DEV)
        F.conv1d(x, w, b, padding=dil * (taps // 2), dilation=dil).backward(dy)
        xt, dyt = ops.nct_to_btc(x.detach(), dtype), ops.nct_to_btc(dy, dtype)
        _, wd = pack(w.detach(), dtype)
        dx = torch.empty_like(xt)
        ops.conv_fwd(dyt, wd, K=N, N=K, taps=taps, dil=dil, out=dx)
        assert rel(ops.btc_to_nct(dx, K), x.grad) < TOL[dtype]
        dw = torch.zeros_like(w)
        db = torch.zeros_like(b)
        ops.conv_wgrad(dyt, xt, dw, K=K, N=N, taps=taps, dil=dil, dbias=db)
        assert rel(dw, w.grad) < TOL[dtype]
        assert rel(db, b.grad) < TOL[dtype]
    finally:
        ops.set_impl("auto")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_grouped_wgrad(dtype):
    ops, nat = _ops()
    torch.manual_seed(4)
    B, T, D, S = 10, 33, 20, 6
    ids = np.array([5, 0, 2, 2, 0, 5, 5, 1, 0, 2])
    h = torch.randn(B, D, T, device=DEV)
    dy = torch.randn(B, D, T, device=DEV)
    order = torch.from_numpy(np.argsort(ids, kind="stable").astype(np.int32)).to(DEV)
    offs = torch.from_numpy(np.concatenate([[0], np.cumsum(np.bincount(ids, minlength=S))]).astype(np.int32)).to(DEV)
    dws = torch.zeros((S, D, D, 1), device=DEV)
    ops.conv_wgrad(ops.nct_to_btc(dy, dtype), ops.nct_to_btc(h, dtype), dws, K=D, N=D, order=order, offsets=offs,
                   G=S, strides=(D * D, D, 1, 0))
    for s in range(S):
        sel = torch.from_numpy(np.where(ids == s)[0]).to(DEV)
        ref = torch.einsum("bnt,bkt->nk", dy[sel], h[sel]) if len(sel) else torch.zeros(D, D, device=DEV)
        assert rel(dws[s, :, :, 0], ref, 1.0) < TOL[dtype], s


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("training", [True, False])
def test_batchnorm_gelu_fwd_bwd(dtype, training):
    ops, nat = _ops()
    torch.manual_seed(5)
    B, T, Cc = 4, 90, 20
    Cp = ops.rup8(Cc)
    y = (torch.randn(B, Cc, T, device=DEV) * 1.5 + 0.3)
    yt = ops.nct_to_btc(y, dtype)
    yq = ops.btc_to_nct(yt, Cc).requires_grad_(True)        # the values the kernels actually see
    bn = torch.nn.BatchNorm1d(Cc).to(DEV)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(); bn.running_mean.normal_(); bn.running_var.uniform_(0.5, 2)
    bn.train(training)
    rm0, rv0 = bn.running_mean.clone(), bn.running_var.clone()
    rm, rv, nbt = rm0.clone(), rv0.clone(), torch.zeros((), dtype=torch.long, device=DEV)
    stats = torch.zeros((2, Cp), dtype=torch.float64, device=DEV)
    nat.call("sd_colstats", yt.data_ptr(), stats.data_ptr(), B * T, Cp, ops.code_of(yt), torch.cuda.current_stream().cuda_stream)
    ss = torch.empty((4, Cp), dtype=torch.float32, device=DEV)
    ops.bn_finalize(stats, Cc, Cp, B * T, bn.weight, bn.bias, rm, rv, nbt, 0.1, 1e-5, training, ss)
    u = torch.empty_like(yt)
    ops.bn_gelu_fwd(yt, ss, u)
    ref = F.gelu(bn(yq))
    assert rel(ops.btc_to_nct(u, Cc), ref) < TOL[dtype]
    if training:
        assert rel(rm, bn.running_mean) < 1e-5 and rel(rv, bn.running_var) < 1e-5 and int(nbt) == 1
    du = torch.randn(B, Cc, T, device=DEV)
    ref.backward(du)
    dut = ops.nct_to_btc(du, dtype)
    red = torch.zeros((2, Cp), dtype=torch.float64, device=DEV)
    dg, db = torch.zeros(Cc, device=DEV), torch.zeros(Cc, device=DEV)
    ops.bn_gelu_bwd(dut, yt, ss, red, dg, db, Cc, training)
    assert rel(ops.btc_to_nct(dut, Cc), yq.grad) < 2 * TOL[dtype]
    assert rel(dg, bn.weight.grad) < 2 * TOL[dtype] and rel(db, bn.bias.grad) < 2 * TOL[dtype]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_glu_and_gelu_backward(dtype):
    ops, nat = _ops()
    torch.manual_seed(6)
    B, T, D2 = 3, 41, 10
    y2 = torch.randn(B, 2 * D2, T, device=DEV)
    y2t = ops.nct_to_btc(y2, dtype)
    y2q = ops.btc_to_nct(y2t, 2 * D2).requires_grad_(True)
    g = torch.randn(B, D2, T, device=DEV)
    F.glu(y2q, dim=-2).backward(g)
    dy2 = torch.empty_like(y2t)
    ops.glu_bwd(ops.nct_to_btc(g, dtype), y2t, dy2, D2)
    assert rel(ops.btc_to_nct(dy2, 2 * D2), y2q.grad) < TOL[dtype]
    p = torch.randn(B, 2 * D2, T, device=DEV)
    pt = ops.nct_to_btc(p, dtype)
    pq = ops.btc_to_nct(pt, 2 * D2).requires_grad_(True)
    du = torch.randn(B, 2 * D2, T, device=DEV)
    F.gelu(pq).backward(du)
    dut = ops.nct_to_btc(du, dtype)
    ops.gelu_bwd(dut, pt)
    assert rel(ops.btc_to_nct(dut, 2 * D2), pq.grad) < TOL[dtype]
    dp = ops.gelu_bwd_nct(du.contiguous(), pt, 2 * D2)
    assert rel(ops.btc_to_nct(dp, 2 * D2), pq.grad) < TOL[dtype]


def test_spatial_attention_weights_fwd_bwd():
    ops, nat = _ops()
    from oracle import restate
    torch.manual_seed(7)
    D1, K, Cc = 30, 6, 23
    loc = restate.synthetic_layout(Cc, 3)
    cos, sin = restate.fourier_tables(K, loc)
    z = torch.complex(torch.rand(D1, K * K), torch.rand(D1, K * K)).requires_grad_(True)
    mask = restate.dropout_mask(loc, 0.3, 4)
    a = z.real @ cos + z.imag @ sin
    w = torch.softmax(a, -1)
    wm = w * mask
    dwm = torch.randn(D1, Cc)
    wm.backward(dwm)
    zr = torch.view_as_real(z.detach()).contiguous().to(DEV)
    w_soft, w_packed = ops.sa_weights_fwd(zr, cos.to(DEV), sin.to(DEV), mask.to(DEV), D1, K * K, Cc, torch.float32)
    assert rel(w_soft, w.detach()) < 1e-5
    assert rel(w_packed[0, 0, :D1, :Cc], wm.detach()) < 1e-5
    assert float(w_packed[0, 0, D1:].abs().max()) == 0 and float(w_packed[0, 0, :, Cc:].abs().max()) == 0
    dz = ops.sa_weights_bwd(dwm.to(DEV), w_soft, mask.to(DEV), cos.to(DEV).t().contiguous(), sin.to(DEV).t().contiguous(), K * K)
    assert rel(dz, torch.view_as_real(z.grad)) < 1e-4


@pytest.mark.parametrize("M,N,D,diag0", [(12, 12, 629, 0), (256, 256, 4096, 0), (64, 16, 1000, 32)])
def test_clip_kernels(M, N, D, diag0):
    """similarity + two-direction CE + gradient against autograd (loss.py:58-79);
    M > N exercises the data-parallel form (global rows x local columns)."""
    ops, nat = _ops()
    torch.manual_seed(8)
    x = torch.randn(M, D, device=DEV)
    z = (torch.randn(N, D, device=DEV) + 0.3 * x[diag0:diag0 + N]).requires_grad_(True)
    temp = torch.tensor([2.3], device=DEV, requires_grad=True)
    xh = x / x.norm(dim=-1, keepdim=True)
    zh = z / z.norm(dim=-1, keepdim=True)
    L = (xh @ zh.T) * torch.exp(temp)
    tgt = torch.arange(N, device=DEV)
    scale = 1.0 / M
    # The kernels see all M (global) rows and the N local columns.  Gradients are those of the global
    # loss restricted to these columns; partial[0] is this rank's share of the loss value (the row-LSE
    # term is counted only for the rank's own N rows so that shares add up across ranks).
    row_lse_t = torch.logsumexp(L, dim=1)
    col_lse_t = torch.logsumexp(L, dim=0)
    diag = L[diag0 + tgt, tgt]
    loss_full = scale * (0.5 * row_lse_t.sum() + 0.5 * col_lse_t.sum() - diag.sum())
    loss_ref = scale * (0.5 * row_lse_t[diag0:diag0 + N].sum() + 0.5 * col_lse_t.sum() - diag.sum())
    if M == N:
        ce = (F.cross_entropy(L, tgt) + F.cross_entropy(L.T, tgt)) / 2
        assert abs(float(ce) - float(loss_full)) < 1e-4 * abs(float(ce))
    loss_full.backward()
    xn2, zn2 = ops.rownorm2(x), ops.rownorm2(z.detach())
    assert rel(xn2, (x * x).sum(1)) < 1e-5
    dots = ops.clip_dots(x, z.detach())
    assert rel(dots, x @ z.detach().T) < 1e-4
    logits, row_stat, col_lse = ops.clip_phase1(dots, xn2, zn2, temp.detach())
    assert rel(logits, L.detach()) < 1e-4
    row_lse = row_stat[:, 0] + torch.log(row_stat[:, 1])
    coef, cz, partial = ops.clip_phase2(logits, row_lse, col_lse, xn2, zn2, temp.detach(), scale, diag0)
    assert rel(partial[0], loss_ref.detach()) < 1e-4
    assert rel(partial[1], temp.grad[0]) < 1e-3
    dz = ops.clip_dz(coef, cz, x, z.detach())
    assert rel(dz, z.grad) < 1e-3


@pytest.mark.parametrize("M,N,D,diag0", [(256, 256, 36864, 0), (64, 48, 4096, 0), (300, 100, 2048, 128), (512, 256, 8192, 256)])
def test_clip_tensor_core_kernels(M, N, D, diag0):
    """TF32 tcgen05 forms of the similarity and gradient GEMMs (bf16 mode) against fp32 matmuls."""
    ops, nat = _ops()
    torch.manual_seed(9)
    x = torch.randn(M, D, device=DEV)
    z = torch.randn(N, D, device=DEV) + 0.3 * x[diag0:diag0 + N]
    ws = torch.empty(nat.lib().sd_clip_dots_workspace_bytes(M, N, D) // 4, device=DEV)
    dots = torch.full((M, N), float("nan"), device=DEV)
    nat.call("sd_clip_dots_tc", x.data_ptr(), z.data_ptr(), dots.data_ptr(), ws.data_ptr(), M, N, D,
             torch.cuda.current_stream().cuda_stream)
    ref = x.double() @ z.double().T
    scale = float((x.norm(dim=1)[:, None] * z.norm(dim=1)[None, :]).max())
    # the tensor core truncates fp32 -> tf32 (10 mantissa bits): a one-sided ~1e-3 relative bias on the
    # large (diagonal) entries, random-sign noise elsewhere
    assert float((dots.double() - ref).abs().max()) / scale < 1.5e-3
    coef = torch.randn(M, N, device=DEV) / M
    Mp = (M + 3) // 4 * 4
    coef_t = torch.zeros(N, Mp, device=DEV)
    coef_t[:, :M] = coef.T
    cz = torch.randn(N, device=DEV) * 0.1
    gs = torch.tensor([0.7], device=DEV)
    dz0 = ops.clip_dz_tc(coef_t, torch.zeros_like(cz), x, z, None)        # GEMM term alone
    assert rel(dz0, coef.double().T @ x.double()) < 3e-3
    dz = ops.clip_dz_tc(coef_t, cz, x, z, gs)
    ref = 0.7 * (coef.double().T @ x.double() - cz.double()[:, None] * z.double())
    assert rel(dz, ref) < 3e-3


@pytest.mark.parametrize("M,N,D,diag0", [(512, 256, 36864, 256), (64, 48, 4096, 0), (300, 100, 2048, 128), (2048, 256, 8192, 512),
                                        (130, 72, 1000, 0)])
def test_clip_bf16_transport_kernels(M, N, D, diag0):
    """Data-parallel CLIP path: rows rounded to bf16 once (+ norms of the rounded rows), kind::f16 GEMMs.
    The GEMMs are exact products of bf16 values accumulated in fp32, so they are checked tightly against
    fp64 matmuls of the ROUNDED operands."""
    ops, nat = _ops()
    torch.manual_seed(11)
    x = torch.randn(M, D, device=DEV)
    z = torch.randn(N, D, device=DEV) + 0.3 * x[diag0:diag0 + N]
    xb, xn2 = ops.cast_rows_bf16(x)
    zb, zn2 = ops.cast_rows_bf16(z)
    assert torch.equal(xb, x.bfloat16())
    assert rel(xn2, (xb.double() ** 2).sum(dim=1)) < 1e-5
    dots = ops.clip_dots(xb, zb)
    ref = xb.double() @ zb.double().T
    scale = float((xn2.sqrt()[:, None] * zn2.sqrt()[None, :]).max())
    assert float((dots.double() - ref).abs().max()) / scale < 2e-6
    coef = torch.randn(M, N, device=DEV) / M
    cz = torch.randn(N, device=DEV) * 0.1
    gs = torch.tensor([0.7], device=DEV)
    dz0 = ops.clip_dz_bf16(coef, torch.zeros_like(cz), xb, z, None)          # GEMM term alone
    assert rel(dz0, coef.bfloat16().double().T @ xb.double()) < 1e-5
    dz = ops.clip_dz_bf16(coef, cz, xb, z, gs)
    ref = 0.7 * (coef.bfloat16().double().T @ xb.double() - cz.double()[:, None] * z.double())
    assert rel(dz, ref) < 1e-5


@pytest.mark.parametrize("B,T,K,N,taps,dil,mode", [
    (100, 360, 320, 320, 3, 4, "res_stats"),      # 300 row tiles: 150 CTA pairs' worth, two column tiles
    (99, 360, 136, 640, 3, 2, "glu"),             # odd number of row tiles: the last pair has a dead half
    (100, 360, 200, 1024, 1, 1, "gelu_nct"),      # final conv: NCT fp32 output + |Z|^2
    (150, 200, 270, 270, 1, 1, "gelu_btc"),
    (100, 300, 64, 48, 3, 16, "res_stats"),
    (100, 360, 1100, 64, 1, 1, "gelu_btc"),       # 18 k-blocks: too many for resident weights, streaming pairs
    (100, 360, 640, 320, 3, 2, "res_stats"),      # GLU-conv data gradient shape: narrow weight-stationary tiles
])
def test_conv_cta_pair_matches_single_cta(B, T, K, N, taps, dil, mode):
    """The cta_group::2 (M=256) conv tiles must reproduce the single-CTA tcgen05 kernel bit for bit
    (same accumulation order) and agree with the fp32 statement of the op."""
    ops, nat = _ops()
    dtype = torch.bfloat16
    torch.manual_seed(3)
    x = torch.randn(B, K, T, device=DEV)
    w = torch.randn(N, K, taps, device=DEV) / (K * taps) ** 0.5
    bias = torch.randn(N, device=DEV)
    xt = ops.nct_to_btc(x, dtype)
    wf, _ = pack(w, dtype)
    Np = ops.rup8(N)
    outs = {}
    try:
        for impl in ("tc_1cta", "tc", "tc_ws"):
            ops.set_impl(impl)
            r = {}
            if mode == "res_stats":
                torch.manual_seed(4)
                rt = ops.nct_to_btc(torch.randn(B, N, T, device=DEV), dtype)
                r["stats"] = torch.zeros((2, Np), dtype=torch.float64, device=DEV)
                r["out"] = torch.empty((B, T, Np), dtype=dtype, device=DEV)
                ops.conv_fwd(xt, wf, K=K, N=N, taps=taps, dil=dil, bias=bias, res=rt, out=r["out"], stats=r["stats"])
                r["res"] = rt
            elif mode == "glu":
                r["pre"] = torch.empty((B, T, Np), dtype=dtype, device=DEV)
                r["out"] = torch.empty((B, T, ops.rup8(N // 2)), dtype=dtype, device=DEV)
                ops.conv_fwd(xt, wf, K=K, N=N, taps=taps, dil=dil, bias=bias, out=r["out"], preact=r["pre"], act=nat.ACT_GLU)
            elif mode == "gelu_nct":
                r["pre"] = torch.empty((B, T, Np), dtype=dtype, device=DEV)
                r["out"] = torch.empty((B, N, T), dtype=torch.float32, device=DEV)
                r["n2"] = torch.zeros((B,), dtype=torch.float32, device=DEV)
                ops.conv_fwd(xt, wf, K=K, N=N, bias=bias, out=r["out"], preact=r["pre"], act=nat.ACT_GELU,
                             out_mode=nat.OUT_NCT_F32, rownorm2=r["n2"])
            else:
                r["pre"] = torch.empty((B, T, Np), dtype=dtype, device=DEV)
                r["out"] = torch.empty((B, T, Np), dtype=dtype, device=DEV)
                ops.conv_fwd(xt, wf, K=K, N=N, bias=bias, out=r["out"], preact=r["pre"], act=nat.ACT_GELU)
            torch.cuda.synchronize()
            outs[impl] = r
    finally:
        ops.set_impl("auto")
    a = outs["tc_1cta"]
    for other in ("tc", "tc_ws"):     # streaming pairs, weight-stationary pairs
        b = outs[other]
        assert torch.equal(a["out"], b["out"]), other
        if "pre" in a:
            assert torch.equal(a["pre"], b["pre"]), other
        if "stats" in a:
            assert rel(b["stats"], a["stats"]) < 1e-6
        if "n2" in a:
            assert rel(b["n2"], a["n2"]) < 1e-5
            assert rel(b["n2"], (b["out"] ** 2).sum(dim=(1, 2))) < 1e-4
    y = F.conv1d(x, w, bias, padding=dil * (taps // 2), dilation=dil)
    if mode == "res_stats":
        assert rel(ops.btc_to_nct(b["out"], N), y + ops.btc_to_nct(a["res"], N)) < TOL[dtype]
    elif mode == "glu":
        assert rel(ops.btc_to_nct(b["out"], N // 2), F.glu(y, dim=-2)) < TOL[dtype]
    elif mode == "gelu_nct":
        assert rel(b["out"], F.gelu(y, approximate="tanh")) < TOL[dtype]
    else:
        assert rel(ops.btc_to_nct(b["out"], N), F.gelu(y, approximate="tanh")) < TOL[dtype]


@pytest.mark.parametrize("B,T,K,N,dil,with_res", [(3, 50, 24, 40, 1, False), (100, 360, 320, 320, 4, True), (2, 360, 640, 320, 2, False),
                                                  (5, 200, 320, 320, 16, True)])
def test_conv_dgrad_with_fused_batchnorm_backward_reduce(B, T, K, N, dil, with_res):
    """sd_conv_args.bnr_y: the data-gradient conv stores g = du * gelu'(scale*y + shift) and accumulates the BatchNorm-backward
    sums in its epilogue.  Against the unfused sequence (conv -> sd_bn_gelu_bwd_reduce -> sd_bn_bwd_apply) and the fp32 math."""
    ops, nat = _ops()
    import sd_b200
    sd_b200.set_precision("bf16")
    dt = torch.bfloat16
    torch.manual_seed(7)
    x = torch.randn(B, K, T, device=DEV)
    w = torch.randn(N, K, 3, device=DEV) / (3 * K) ** 0.5
    y = torch.randn(B, N, T, device=DEV) * 1.5 + 0.3
    res = torch.randn(B, N, T, device=DEV) if with_res else None
    gamma, beta = torch.rand(N, device=DEV) + 0.5, torch.randn(N, device=DEV) * 0.2
    xt, yt = ops.nct_to_btc(x, dt), ops.nct_to_btc(y, dt)
    rt = ops.nct_to_btc(res, dt) if with_res else None
    wf, _ = pack(w, dt)
    Np = ops.rup8(N)
    # BatchNorm scale / shift / mean / invstd of y as the forward would have left them
    yf = yt.float()[:, :, :N]
    mean, var = yf.mean(dim=(0, 1)), yf.var(dim=(0, 1), unbiased=False)
    invstd = (var + 1e-5).rsqrt()
    ss = torch.zeros((4, Np), device=DEV)
    ss[0, :N], ss[1, :N], ss[2, :N], ss[3, :N] = gamma * invstd, beta - mean * gamma * invstd, mean, invstd
    rows = B * T

    def run(fused):
        out = torch.empty((B, T, Np), dtype=dt, device=DEV)
        red = torch.zeros((2, Np), dtype=torch.float64, device=DEV)
        dg, db = torch.zeros(N, device=DEV), torch.zeros(N, device=DEV)
        if fused:
            ops.conv_fwd(xt, wf, K=K, N=N, taps=3, dil=dil, res=rt, out=out, bnr_y=yt, bnr_ss=ss, stats=red)
        else:
            ops.conv_fwd(xt, wf, K=K, N=N, taps=3, dil=dil, res=rt, out=out)
        ops.bn_gelu_bwd(out, yt, ss, red, dg, db, N, True, None, g_ready=fused)
        return out.float()[:, :, :N], red[:, :N].clone(), dg, db

    dy_f, red_f, dg_f, db_f = run(True)
    dy_u, red_u, dg_u, db_u = run(False)
    # fp32 statement of the whole thing (tanh-form GELU derivative is within bf16 noise of the erf form)
    du = F.conv1d(xt.float()[:, :, :K].transpose(1, 2), w.to(dt).float(), None, padding=dil, dilation=dil)
    if with_res:
        du = du + rt.float()[:, :, :N].transpose(1, 2)
    yn = yf.transpose(1, 2)
    pre = (yn * ss[0, :N, None] + ss[1, :N, None]).requires_grad_(True)
    F.gelu(pre).backward(du)
    g = pre.grad
    xhat = (yn - mean[:, None]) * invstd[:, None]
    sg, sgx = g.sum(dim=(0, 2)), (g * xhat).sum(dim=(0, 2))
    dy_ref = (gamma * invstd)[:, None] * (g - sg[:, None] / rows - xhat * sgx[:, None] / rows)

    def nerr(a, b):
        return float((a.double() - b.double()).norm() / b.double().norm())
    assert nerr(red_f[0], red_u[0]) < 5e-3 and nerr(red_f[1], red_u[1]) < 5e-3        # fused sums see the unrounded du
    assert nerr(dy_f, dy_u) < 1e-2
    assert nerr(dy_f.transpose(1, 2), dy_ref) < 1.5e-2 and nerr(dy_u.transpose(1, 2), dy_ref) < 1.5e-2
    assert nerr(dg_f, sgx) < 1e-2 and nerr(db_f, sg) < 1e-2
    assert nerr(dy_f.transpose(1, 2), dy_ref) < 1.2 * nerr(dy_u.transpose(1, 2), dy_ref) + 1e-3
