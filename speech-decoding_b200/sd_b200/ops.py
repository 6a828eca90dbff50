"""Thin tensor-level wrappers over the C ABI (include/sd_b200.h).  PyTorch is
used only for device memory, streams and the autograd graph edge; every
computation below is a launch into libsd_b200.so."""
import os

import torch

from . import _native as nat

_PRECISIONS = {"fp32": (torch.float32, nat.SD_F32), "bf16": (torch.bfloat16, nat.SD_BF16),
               "tf32": (torch.float32, nat.SD_F32), "tf32x3": (torch.float32, nat.SD_F32)}
_precision = os.environ.get("SD_B200_PRECISION", "bf16")


def set_precision(p: str):
    """'bf16' (default; bf16 activations/weight shadows, fp32 accumulate and master parameters),
    'tf32' (fp32 storage, every conv / GEMM on tcgen05 kind::tf32 -- what the reference gets from cuDNN on a GPU),
    'tf32x3' (fp32 storage, every conv product as three TF32 MMAs on pre-split hi/lo operands: fp32-class accuracy
    on the tensor cores, the mode that meets the 1e-4 parity bar) or 'fp32' (fp32 everywhere on the CUDA cores)."""
    global _precision
    if p not in _PRECISIONS:
        raise ValueError("precision must be one of %s" % list(_PRECISIONS))
    _precision = p


def get_precision() -> str:
    return _precision


def dtypes(precision=None):
    return _PRECISIONS[precision or _precision]


_IMPL = "auto"


def get_impl():
    return _IMPL


def set_impl(name: str):
    """'auto' | 'simt' | 'tc' | 'tc_1cta' | 'tc_ws' -- kernel family for conv / wgrad (tests)."""
    global _IMPL
    _IMPL = name
    nat.call("sd_set_impl", {"auto": nat.IMPL_AUTO, "simt": nat.IMPL_SIMT, "tc": nat.IMPL_TC, "tc_1cta": nat.IMPL_TC_1CTA,
                                "tc_ws": nat.IMPL_TC_WS}[name])


def rup8(c: int) -> int:
    return (c + 7) // 8 * 8


def _p(t):
    return None if t is None else t.data_ptr()


_stream_cache = [None]


def _st():
    """raw cudaStream_t of torch's current stream.  torch.cuda.current_stream() costs ~15 us per call, so the
    executor pins the handle for the duration of one forward / backward (see stream_scope)."""
    s = _stream_cache[0]
    return s if s is not None else torch.cuda.current_stream().cuda_stream


class stream_scope:
    def __enter__(self):
        self.prev = _stream_cache[0]
        _stream_cache[0] = torch.cuda.current_stream().cuda_stream
        return self

    def __exit__(self, *a):
        _stream_cache[0] = self.prev
        return False


def require_cuda(t, name):
    if not t.is_cuda:
        raise RuntimeError("sd_b200: %s must be a CUDA tensor -- this implementation has no CPU path "
                           "(got device=%s)" % (name, t.device))


def tensor_core_fp32():
    """fp32 storage with TF32 tensor-core convolutions?  -> 0 (no), 1 (TF32), 3 (3xTF32)"""
    return {"tf32": 1, "tf32x3": 3}.get(_precision, 0)


def tf32_split(x):
    """x fp32 -> (hi, lo) planes of the 3xTF32 kernels; a tensor that already carries its low plane (packed weights:
    engine.WeightPack splits them once per step) is passed through."""
    lo = getattr(x, "_sd_lo", None)
    if lo is not None:
        return x, lo
    x = x.contiguous()
    out = torch.empty((2,) + tuple(x.shape), dtype=torch.float32, device=x.device)
    nat.call("sd_tf32_split", _p(x), _p(out[0]), _p(out[1]), x.numel(), _st())
    return out[0], out[1]


def code_of(t):
    if t.dtype == torch.float32:
        return nat.SD_F32
    if t.dtype == torch.bfloat16:
        return nat.SD_BF16
    raise RuntimeError("sd_b200: unsupported activation dtype %s" % t.dtype)


# ---- layout -------------------------------------------------------------------------------------
def nct_to_btc(x, dtype):
    B, C, T = x.shape
    Cp = rup8(C)
    out = torch.empty((B, T, Cp), dtype=dtype, device=x.device)
    nat.call("sd_nct_to_btc_bf16in" if x.dtype == torch.bfloat16 else "sd_nct_to_btc", _p(x), _p(out), B, C, T, Cp, code_of(out), _st())
    return out


def btc_to_nct(a, C):
    B, T, Cp = a.shape
    out = torch.empty((B, C, T), dtype=torch.float32, device=a.device)
    nat.call("sd_btc_to_nct", _p(a), _p(out), B, C, T, Cp, code_of(a), _st())
    return out


# ---- conv ----------------------------------------------------------------------------------------
def conv_fwd(inp, w, *, K, N, taps=1, dil=1, bias=None, res=None, widx=None, G=1, out=None, preact=None,
             stats=None, rownorm2=None, act=nat.ACT_NONE, out_mode=nat.OUT_BTC, affine=None, bnr_y=None, bnr_ss=None):
    B, T, Kp = inp.shape
    Np = rup8(N)
    code, in_lo, w_lo = code_of(inp), None, None
    tc32 = tensor_core_fp32() if inp.dtype == torch.float32 else 0
    if tc32:
        code = nat.SD_TF32
        if tc32 == 3:
            inp, in_lo = tf32_split(inp)
            w, w_lo = tf32_split(w)
    a = nat.ConvArgs(_p(inp), _p(w), _p(bias), _p(res), _p(widx), _p(out), _p(preact), _p(stats), _p(rownorm2),
                     B, T, K, Kp, N, Np, taps, dil, G, act, out_mode, code, _p(affine), _p(in_lo), _p(w_lo), _p(bnr_y), _p(bnr_ss))
    nat.call("sd_conv_fwd", a, _st())
    return out


_WS = {}


def _wgrad_workspace(device, mb=64):
    """persistent split-K scratch (64 MB; 160 MB for the 3xTF32 mode, which uses twice the split-K slices) for the
    tensor-core wgrad kernels, one per device"""
    key = (device.type, device.index)
    if key not in _WS or _WS[key].numel() * 4 < mb << 20:
        _WS[key] = torch.empty(mb * 256 * 1024, dtype=torch.float32, device=device)
    return _WS[key]


def conv_wgrad(dout, inp, dw, *, K, N, taps=1, dil=1, dbias=None, order=None, offsets=None, G=1,
               strides=None):
    B, T, Np = dout.shape
    Kp = inp.shape[2]
    if strides is None:                      # PyTorch Conv1d weight (N, K, taps)
        strides = (N * K * taps, K * taps, taps, 1)
    code, dout_lo, in_lo = code_of(dout), None, None
    tc32 = tensor_core_fp32() if dout.dtype == torch.float32 else 0
    if tc32:
        code = nat.SD_TF32
        if tc32 == 3:
            if dbias is not None:
                # the bias gradient is a plain column sum with heavy cancellation: fp64 accumulation on the CUDA cores
                # instead of the ones-tile MMA (whose truncating accumulator costs ~1e-4 here)
                scratch = torch.empty((2 * Np,), dtype=torch.float64, device=dout.device)
                nat.call("sd_colsum_add", _p(dout), _p(dbias), _p(scratch), B * T, N, Np, nat.SD_F32, _st())
                dbias = None
            dout, dout_lo = tf32_split(dout)
            inp, in_lo = tf32_split(inp)
    ws = _wgrad_workspace(dout.device, 160 if tc32 == 3 else 64) if (dout.dtype == torch.bfloat16 or tc32) and G == 1 else None
    a = nat.WgradArgs(_p(dout), _p(inp), _p(dw), _p(dbias), _p(order), _p(offsets),
                      B, T, K, Kp, N, Np, taps, dil, G, strides[0], strides[1], strides[2], strides[3],
                      code, _p(ws), ws.numel() * 4 if ws is not None else 0, _p(dout_lo), _p(in_lo))
    nat.call("sd_conv_wgrad", a, _st())


# ---- batchnorm / gelu / glu ---------------------------------------------------------------------
def bn_finalize(stats, C, Cp, n, gamma, beta, rmean, rvar, nbt, momentum, eps, training, ss):
    nat.call("sd_bn_finalize", _p(stats), C, Cp, n, _p(gamma), _p(beta), _p(rmean), _p(rvar), _p(nbt),
             momentum, eps, int(training), _p(ss), _st())


def bn_gelu_fwd(y, ss, u):
    rows = y.shape[0] * y.shape[1]
    nat.call("sd_bn_gelu_fwd", _p(y), _p(ss), _p(u), rows, y.shape[2], code_of(y), _st())


def bn_bwd_fusable(y):
    """can the BatchNorm-backward reduce ride in the epilogue of the conv that produces du?  (bf16 tensor-core path)"""
    return y.dtype == torch.bfloat16 and get_impl() != "simt" and _BNR_FUSE


# Three ways to run the BatchNorm backward, measured at cfg2 on one B200 (same build, same box, 30 steps):
#   stand-alone reduce + apply passes (default)                              6.795 ms/step, conv kernel 914 TFLOP/s
#   reduce fused into the producing conv's epilogue (SD_B200_BNR_FUSE=1)      6.757 ms/step, conv kernel 827 TFLOP/s
#   stand-alone reduce that stores g, apply without gelu' (SD_B200_BN_STORE_G=1)  6.98 ms/step
# The fused epilogue removes a 35 us pass per BatchNorm but costs its conv 23 us: epilogue work is not free next to the
# MMAs (shared-memory traffic, issue slots).  0.5 % is not worth a 10 % slower conv kernel, so it stays opt-in (and tested).
_BNR_FUSE = os.environ.get("SD_B200_BNR_FUSE", "0") == "1"
_BN_STORE_G = os.environ.get("SD_B200_BN_STORE_G", "0") == "1"


def bn_gelu_bwd(du, y, ss, red, dgamma, dbeta, C, training, group=None, g_ready=False):
    """in place: du -> dy (gradient w.r.t. the pre-BN tensor y).  With `group`
    (SyncBN) the two per-channel sums are all-reduced between the passes.
    g_ready: `du` already holds g = du * gelu'(bn(y)) and `red` its sums (written by the producing conv, bnr_y)."""
    rows, Cp = y.shape[0] * y.shape[1], y.shape[2]
    n_stat, dscale = rows, 1.0
    if not g_ready:
        if _BN_STORE_G:      # the reduce pass leaves g in place of du, the apply pass does not recompute the GELU derivative
            nat.call("sd_bn_gelu_bwd_reduce_g", _p(du), _p(y), _p(ss), _p(red), rows, Cp, code_of(y), _st())
            g_ready = True
        else:
            nat.call("sd_bn_gelu_bwd_reduce", _p(du), _p(y), _p(ss), _p(red), rows, Cp, code_of(y), _st())
    if group is not None and training:
        import torch.distributed as dist
        from . import dist as sd_dist
        sd_dist.small_all_reduce_sum_(red, group)
        n_stat = rows * dist.get_world_size(group)
        dscale = 1.0 / dist.get_world_size(group)     # red is global already; the grad all-reduce sums once more
    nat.call("sd_bn_bwd_apply_g" if g_ready else "sd_bn_bwd_apply", _p(du), _p(y), _p(ss), _p(red), _p(dgamma), _p(dbeta), rows,
             n_stat, dscale, C, Cp, int(training), code_of(y), _st())


def glu_bwd(dout, y2, dy2, D2):
    rows = y2.shape[0] * y2.shape[1]
    nat.call("sd_glu_bwd", _p(dout), _p(y2), _p(dy2), rows, D2, y2.shape[2], dout.shape[2], code_of(y2), _st())


def gelu_bwd(du, p):
    rows = p.shape[0] * p.shape[1]
    nat.call("sd_gelu_bwd", _p(du), _p(p), rows, p.shape[2], code_of(p), _st())


def gelu_bwd_nct(dz, p, N):
    B, T, Np = p.shape
    dp = torch.empty_like(p)
    nat.call("sd_gelu_bwd_nct", _p(dz), _p(p), _p(dp), B, N, T, Np, code_of(p), _st())
    return dp


# ---- spatial attention --------------------------------------------------------------------------
def sa_weights_fwd(z_ri, cos, sin, mask, D1, K2, C, dtype):
    D1p, Cp = rup8(D1), rup8(C)
    w_soft = torch.empty((D1, C), dtype=torch.float32, device=z_ri.device)
    w_packed = torch.empty((1, 1, D1p, Cp), dtype=dtype, device=z_ri.device)
    scratch = torch.empty((nat.SA_MPARTS, D1, C), dtype=torch.float32, device=z_ri.device)      # SD_SA_MPARTS partial logits
    nat.call("sd_sa_weights_fwd", _p(z_ri), _p(cos), _p(sin), _p(mask), _p(w_soft), _p(w_packed), _p(scratch),
             D1, K2, C, D1p, Cp, code_of(w_packed), _st())
    return w_soft, w_packed


def sa_weights_bwd(dwm, w_soft, mask, cos_T, sin_T, K2, out=None):
    """cos_T, sin_T: the (C, K^2) transposes of the module's cos/sin buffers"""
    D1, C = w_soft.shape
    dz = out if out is not None else torch.empty((D1, K2, 2), dtype=torch.float32, device=dwm.device)
    nat.call("sd_sa_weights_bwd", _p(dwm), _p(w_soft), _p(mask), _p(cos_T), _p(sin_T), _p(dz), D1, K2, C, _st())
    return dz


# ---- CLIP -----------------------------------------------------------------------------------------
def rownorm2(x2d):
    M, D = x2d.shape
    out = torch.empty((M,), dtype=torch.float32, device=x2d.device)
    nat.call("sd_rownorm2", _p(x2d), _p(out), M, D, _st())
    return out


def clip_tc_ok(x2d, z2d):
    """tensor-core (TF32) CLIP kernels: used by the bf16 mode when the shape allows TMA"""
    M, D = x2d.shape
    return (get_precision() in ("bf16", "tf32") and D % 4 == 0 and D >= 64
            and nat.lib().sd_clip_dots_workspace_bytes(M, z2d.shape[0], D) > 0)


def cast_rows_bf16(x2d):
    """fp32 rows -> (bf16 rows, squared norms of the rounded rows); one pass."""
    M, D = x2d.shape
    y = torch.empty((M, D), dtype=torch.bfloat16, device=x2d.device)
    n2 = torch.empty((M,), dtype=torch.float32, device=x2d.device)
    nat.call("sd_cast_rows_bf16", _p(x2d), _p(y), _p(n2), M, D, _st())
    return y, n2


def rownorm2_bf16(xb):
    """squared norms of rows that are already bf16"""
    M, D = xb.shape
    n2 = torch.empty((M,), dtype=torch.float32, device=xb.device)
    nat.call("sd_rownorm2_bf16", _p(xb), _p(n2), M, D, _st())
    return n2


def clip_bf16_ok(x2d):
    """bf16 transport of the speech rows (data-parallel CLIP): bf16 mode and rows TMA can address"""
    return get_precision() == "bf16" and x2d.shape[1] % 8 == 0 and x2d.shape[1] >= 64


def clip_dots(x2d, z2d, tc=False):
    M, D = x2d.shape
    Nn = z2d.shape[0]
    if x2d.dtype == torch.bfloat16:
        ws_bytes = nat.lib().sd_clip_dots_workspace_bytes(M, Nn, D)
        ws = torch.empty((ws_bytes // 4,), dtype=torch.float32, device=x2d.device)
        dots = torch.empty((M, Nn), dtype=torch.float32, device=x2d.device)
        nat.call("sd_clip_dots_tc_bf16", _p(x2d), _p(z2d), _p(dots), _p(ws), M, Nn, D, _st())
        return dots
    if tc:
        ws_bytes = nat.lib().sd_clip_dots_workspace_bytes(M, Nn, D)
        ws = torch.empty((ws_bytes // 4,), dtype=torch.float32, device=x2d.device)
        dots = torch.empty((M, Nn), dtype=torch.float32, device=x2d.device)
        nat.call("sd_clip_dots_tc", _p(x2d), _p(z2d), _p(dots), _p(ws), M, Nn, D, _st())
        return dots
    dots = torch.zeros((M, Nn), dtype=torch.float32, device=x2d.device)
    nat.call("sd_clip_dots", _p(x2d), _p(z2d), _p(dots), M, Nn, D, _st())
    return dots


def clip_phase1(dots, xn2, zn2, temp):
    M, Nn = dots.shape
    logits = torch.empty_like(dots)
    row_stat = torch.empty((M, 2), dtype=torch.float32, device=dots.device)
    col_lse = torch.empty((Nn,), dtype=torch.float32, device=dots.device)
    nat.call("sd_clip_phase1", _p(dots), _p(xn2), _p(zn2), _p(temp), _p(logits), _p(row_stat), _p(col_lse),
             M, Nn, _st())
    return logits, row_stat, col_lse


def clip_phase2(logits, row_lse, col_lse, xn2, zn2, temp, scale, diag0, want_t=False):
    M, Nn = logits.shape
    coef = torch.empty_like(logits)
    coef_t = torch.zeros((Nn, (M + 3) // 4 * 4), dtype=torch.float32, device=logits.device) if want_t else None
    cz = torch.empty((Nn,), dtype=torch.float32, device=logits.device)
    partial = torch.empty((2,), dtype=torch.float32, device=logits.device)
    nat.call("sd_clip_phase2", _p(logits), _p(row_lse), _p(col_lse), _p(xn2), _p(zn2), _p(temp), scale, diag0,
             _p(coef), _p(coef_t), _p(cz), _p(partial), M, Nn, _st())
    if want_t:
        return coef, coef_t, cz, partial
    return coef, cz, partial


def clip_dz_tc(coef_t, cz, x2d, z2d, gscale=None):
    M, D = x2d.shape
    Nn = z2d.shape[0]
    dz = torch.empty((Nn, D), dtype=torch.float32, device=x2d.device)
    nat.call("sd_clip_dz_tc", _p(coef_t), _p(cz), _p(x2d), _p(z2d), _p(dz), _p(gscale), M, Nn, D, _st())
    return dz


def clip_dz_bf16(coef, cz, xb, z2d, gscale=None):
    """dz with coef (M,Nn) fp32 -> transposed bf16 A operand, xb (M,D) bf16 rows, z2d (Nn,D) fp32."""
    M, D = xb.shape
    Nn = z2d.shape[0]
    Mp = (M + 7) // 8 * 8
    ct = torch.empty((Nn, Mp), dtype=torch.bfloat16, device=xb.device)
    nat.call("sd_clip_coef_t_bf16", _p(coef), _p(ct), M, Nn, Mp, _st())
    dz = torch.empty((Nn, D), dtype=torch.float32, device=xb.device)
    nat.call("sd_clip_dz_tc_bf16", _p(ct), _p(cz), _p(xb), _p(z2d), _p(dz), _p(gscale), M, Nn, D, _st())
    return dz


def clip_dz(coef, cz, x2d, z2d, gscale=None):
    M, D = x2d.shape
    Nn = z2d.shape[0]
    dz = torch.empty((Nn, D), dtype=torch.float32, device=x2d.device)
    nat.call("sd_clip_dz", _p(coef), _p(cz), _p(x2d), _p(z2d), _p(dz), _p(gscale), M, Nn, D, _st())
    return dz
