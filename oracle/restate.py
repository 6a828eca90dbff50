"""TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the reference hot path.

A plain fp32 PyTorch/NumPy re-statement of the algorithm in
  /root/reference/speech_decoding/models.py      (SpatialAttention :14-65,
      SpatialDropout :68-86, SubjectBlock :89-117, ConvBlock :120-166,
      BrainEncoder :169-196, Classifier :199-248)
  /root/reference/speech_decoding/utils/loss.py  (CLIPLoss :28-84)
written functionally over a `state_dict`-shaped dict of tensors (same keys,
shapes and dtypes as the reference's BrainEncoder.state_dict()),
so the same weights drive the reference, this oracle and the CUDA drop-in.

Parity pinning: the reference ships no golden vectors, so
this oracle is pinned against OUTPUTS OF THE UNMODIFIED REFERENCE run in the
build container: tests/golden/*.npz (made by oracle/gen_golden.py, committed)
and, when /root/reference is present, a live side-by-side check
(tests/test_oracle.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference
arm may import this file.  The product path never does.
"""
import math
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------
# args / layout helpers
# --------------------------------------------------------------------------
class Args(SimpleNamespace):
    """dict-with-attribute-access config, as the constructors read it
    (models.py:22-43,93-95,173-178; loss.py:32,36)."""

    def __getitem__(self, k):
        return getattr(self, k)


def make_args(D1=270, D2=320, F_=1024, K=32, d_drop=0.1, num_subjects=27,
              dataset="Gwilliams2022", num_channels=208, last4layers=True,
              reduction="mean", init_temperature=5.1, layout_seed=0):
    return Args(D1=D1, D2=D2, F=F_, K=K, d_drop=d_drop, num_subjects=num_subjects,
                dataset=dataset, root_dir="/nonexistent", num_channels=num_channels,
                preprocs={"last4layers": last4layers}, reduction=reduction,
                init_temperature=init_temperature, layout_seed=layout_seed)


def synthetic_layout(num_channels: int, seed: int = 0) -> torch.Tensor:
    """Stand-in for ch_locations_2d honouring its post-conditions
    (layout.py:37-43): float32 (C,2), per-axis min-max normalised, then
    *0.8+0.1 so values lie in [0.1, 0.9] with exact extremes."""
    rng = np.random.RandomState(1234 + seed)
    loc = rng.rand(num_channels, 2).astype(np.float64)
    loc = (loc - loc.min(axis=0)) / (loc.max(axis=0) - loc.min(axis=0))
    loc = loc * 0.8 + 0.1
    return torch.from_numpy(loc.astype(np.float32))


def fourier_tables(K: int, loc: torch.Tensor):
    """cos/sin buffers, (K*K, C) fp32; row m=(k,l)=(m//K, m%K)
    (models.py:20-26,36-40)."""
    idx = torch.arange(K * K)
    k = idx // K
    l = idx % K
    x, y = loc[:, 0], loc[:, 1]
    # integer grid promoted to float exactly as einsum("k,x->kx", long, float) does
    phi = 2 * torch.pi * (k[:, None] * x[None, :] + l[:, None] * y[None, :])
    return torch.cos(phi), torch.sin(phi)


def dropout_mask(loc: torch.Tensor, d_drop: float, center: int) -> torch.Tensor:
    """(C,) 0/1 mask: sensors closer than d_drop to sensor `center` are dropped
    (models.py:81-83).  `center` is the value np.random.randint(C) returned."""
    d = (loc - loc[center]).norm(dim=-1)
    return torch.where(d < d_drop, 0.0, 1.0)


# --------------------------------------------------------------------------
# forward pieces (autograd supplies the backward)
# --------------------------------------------------------------------------
def spatial_attention(z, cos, sin, X, mask=None):
    """models.py:45-65.  z complex64 (D1,K^2); X (B,C,T); mask (C,) or None."""
    a = z.real @ cos + z.imag @ sin                    # (D1, C)   :49-53
    w = torch.softmax(a, dim=-1)                       # :58
    if mask is not None:                               # :62 -> :77-84
        X = X * mask.to(X.device)[None, :, None]
    return torch.einsum("oi,bit->bot", w, X)           # :65


def subject_block(sd, prefix, X, subject_idxs, mask=None):
    """models.py:111-117."""
    p = prefix
    h = spatial_attention(sd[p + "spatial_attention.z"], sd[p + "spatial_attention.cos"],
                          sd[p + "spatial_attention.sin"], X, mask)
    h = F.conv1d(h, sd[p + "conv.weight"], sd[p + "conv.bias"])            # :113
    outs = []
    for b, s in enumerate(list(subject_idxs)):                             # :114-116
        outs.append(F.conv1d(h[b:b + 1], sd[p + "subject_layer.%d.weight" % int(s)]))
    return torch.cat(outs)


def _bn(sd, p, X, train, momentum=0.1, eps=1e-5):
    """nn.BatchNorm1d defaults (models.py:135,143).  In train mode the running
    statistics inside `sd` are updated in place like the module does."""
    if train:
        sd[p + "num_batches_tracked"] += 1
    return F.batch_norm(X, sd[p + "running_mean"], sd[p + "running_var"],
                        sd[p + "weight"], sd[p + "bias"], training=train,
                        momentum=momentum, eps=eps)


def conv_block(sd, p, k, X, train):
    """models.py:152-166; dilations models.py:133,141,149."""
    d0 = 2 ** ((2 * k) % 5)
    d1 = 2 ** ((2 * k + 1) % 5)
    y = F.conv1d(X, sd[p + "conv0.weight"], sd[p + "conv0.bias"], padding=d0, dilation=d0)
    if k != 0:
        y = y + X
    u = F.gelu(_bn(sd, p + "batchnorm0.", y, train))
    y = F.conv1d(u, sd[p + "conv1.weight"], sd[p + "conv1.bias"], padding=d1, dilation=d1) + u
    u = F.gelu(_bn(sd, p + "batchnorm1.", y, train))
    y = F.conv1d(u, sd[p + "conv2.weight"], sd[p + "conv2.bias"], padding=2, dilation=2)
    return F.glu(y, dim=-2)


def encoder_forward(sd, X, subject_idxs, train=True, mask=None):
    """BrainEncoder.forward, models.py:191-196.  `sd` is a state_dict-shaped
    dict; BN running stats in it are updated when train=True."""
    h = subject_block(sd, "subject_block.", X, subject_idxs, mask if train else None)
    for k in range(5):
        h = conv_block(sd, "conv_blocks.conv%d." % k, k, h, train)
    h = F.gelu(F.conv1d(h, sd["conv_final1.weight"], sd["conv_final1.bias"]))
    h = F.gelu(F.conv1d(h, sd["conv_final2.weight"], sd["conv_final2.bias"]))
    return h


def clip_loss(x, y, temp, reduction="mean", fast=True, return_logits=False):
    """CLIPLoss.forward, loss.py:38-84.  x=Y (speech), y=Z (brain) at the
    reference call site train.py:191."""
    B = x.size(0)
    assert B > 1, "Batch size must be greater than 1."            # loss.py:40
    targets = torch.arange(B, device=x.device)
    if not fast:                                                  # loss.py:46-50
        x_ = x.reshape(1, B, -1)
        y_ = y.reshape(B, 1, -1)
        logits = F.cosine_similarity(x_, y_, dim=-1)
    else:                                                         # loss.py:58-71
        xf = x.reshape(B, -1)
        yf = y.reshape(B, -1)
        xf = xf / xf.norm(dim=-1, keepdim=True)
        yf = yf / yf.norm(dim=-1, keepdim=True)
        logits = (xf @ yf.T) * torch.exp(temp)
    loss = (F.cross_entropy(logits, targets, reduction=reduction)
            + F.cross_entropy(logits.t(), targets, reduction=reduction)) / 2   # loss.py:79
    return (logits, loss) if return_logits else loss


def classifier(Z, Y):
    """Classifier.forward, models.py:208-248, vectorised: similarity[i,j] =
    cos(Z_i, Y_j) with the max(|x||y|, 1e-8) guard (:228), transposed (:233),
    top-1 (:236) and top-10 (:238-243) hit rates against the diagonal."""
    B = Z.size(0)
    x = Z.reshape(B, -1).double()
    y = Y.reshape(B, -1).double()
    den = torch.clamp(x.norm(dim=1)[:, None] * y.norm(dim=1)[None, :], min=1e-8)
    sim = ((x @ y.T) / den).float().T
    diags = torch.arange(B, device=sim.device)
    top1 = (sim.argmax(dim=1) == diags).float().mean().item()
    k = min(10, B)
    top10_idx = torch.topk(sim, k, dim=1, largest=True)[1]
    top10 = float((top10_idx == diags[:, None]).any(dim=1).float().mean())
    return top1, top10, top10_idx, sim


# --------------------------------------------------------------------------
# closed-form backward of the CLIP loss, numpy fp64.
# Used to cross-check autograd on the restatement and to document the formula
# the CUDA kernel implements.
# --------------------------------------------------------------------------
def clip_loss_closed_form(x, y, temp, reduction="mean"):
    x = np.asarray(x, dtype=np.float64).reshape(x.shape[0], -1)
    y = np.asarray(y, dtype=np.float64).reshape(y.shape[0], -1)
    B = x.shape[0]
    s = math.exp(float(temp))
    nx = np.linalg.norm(x, axis=1)
    ny = np.linalg.norm(y, axis=1)
    xh = x / nx[:, None]
    yh = y / ny[:, None]
    L = s * (xh @ yh.T)
    def lse(a, axis):
        m = a.max(axis=axis, keepdims=True)
        return (m + np.log(np.exp(a - m).sum(axis=axis, keepdims=True))).squeeze(axis)
    scale = 1.0 / B if reduction == "mean" else 1.0
    loss = scale * (-np.trace(L) + 0.5 * lse(L, 1).sum() + 0.5 * lse(L, 0).sum())
    P_r = np.exp(L - lse(L, 1)[:, None])
    P_c = np.exp(L - lse(L, 0)[None, :])
    G = 0.5 * scale * (P_r + P_c - 2 * np.eye(B))
    dtemp = (G * L).sum()
    c = (G * L).sum(axis=0)                       # per brain column j
    dy = (s * (G.T @ xh) - c[:, None] * yh) / ny[:, None]
    return loss, L, dy, dtemp


# --------------------------------------------------------------------------
# state-dict construction with the reference's initialisers (appendix A.6)
# --------------------------------------------------------------------------
def init_state_dict(args, C, gen=None):
    """Fresh parameters/buffers with the reference's shapes, dtypes, key names
    and default initialisers (models.py:33; nn.Conv1d / nn.BatchNorm1d
    defaults).  Not bit-identical to constructing the reference under the same
    seed (different draw order) -- parity tests copy weights instead."""
    g = gen or torch.Generator().manual_seed(0)
    sd = {}
    D1, D2, K = args.D1, args.D2, args.K
    Fo = args.F if not args.preprocs["last4layers"] else 1024
    loc = synthetic_layout(C, getattr(args, "layout_seed", 0))
    cos, sin = fourier_tables(K, loc)

    def conv(prefix, co, ci, k, bias=True):
        bound = 1.0 / math.sqrt(ci * k)   # kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(fan_in), +)
        sd[prefix + "weight"] = (torch.rand(co, ci, k, generator=g) * 2 - 1) * bound
        if bias:
            sd[prefix + "bias"] = (torch.rand(co, generator=g) * 2 - 1) * bound

    def bn(prefix, c):
        sd[prefix + "weight"] = torch.ones(c)
        sd[prefix + "bias"] = torch.zeros(c)
        sd[prefix + "running_mean"] = torch.zeros(c)
        sd[prefix + "running_var"] = torch.ones(c)
        sd[prefix + "num_batches_tracked"] = torch.tensor(0, dtype=torch.long)

    sd["subject_block.spatial_attention.z"] = torch.complex(
        torch.rand(D1, K * K, generator=g), torch.rand(D1, K * K, generator=g))
    sd["subject_block.spatial_attention.cos"] = cos
    sd["subject_block.spatial_attention.sin"] = sin
    conv("subject_block.conv.", D1, D1, 1)
    for s in range(args.num_subjects):
        conv("subject_block.subject_layer.%d." % s, D1, D1, 1, bias=False)
    for k in range(5):
        p = "conv_blocks.conv%d." % k
        conv(p + "conv0.", D2, D1 if k == 0 else D2, 3)
        bn(p + "batchnorm0.", D2)
        conv(p + "conv1.", D2, D2, 3)
        bn(p + "batchnorm1.", D2)
        conv(p + "conv2.", 2 * D2, D2, 3)
    conv("conv_final1.", 2 * D2, D2, 1)
    conv("conv_final2.", Fo, 2 * D2, 1)
    return sd, loc


PARAM_SUFFIXES = ("weight", "bias", ".z")


def is_param(key: str) -> bool:
    return key.endswith("weight") or key.endswith("bias") or key.endswith(".z")


def train_step(sd, X, Y, subject_idxs, temp, mask, reduction="mean", train=True):
    """One fwd+bwd of the hot path on the oracle: returns dict with Z, loss,
    logits, grads (keyed like the state dict, absent subjects -> None) and the
    updated BN buffers (in sd)."""
    leaves = {}
    work = {}
    for k, v in sd.items():
        if is_param(k):
            t = v.detach().clone().requires_grad_(True)
            leaves[k] = t
            work[k] = t
        else:
            work[k] = v
    temp_leaf = temp.detach().clone().requires_grad_(True)
    Z = encoder_forward(work, X, subject_idxs, train=train, mask=mask)
    logits, loss = clip_loss(Y, Z, temp_leaf, reduction=reduction, return_logits=True)
    Z.retain_grad()
    loss.backward()
    for k, v in work.items():          # propagate updated BN buffers
        if not is_param(k):
            sd[k] = v
    grads = {k: (t.grad if t.grad is not None else None) for k, t in leaves.items()}
    return dict(Z=Z.detach(), loss=loss.detach(), logits=logits.detach(), dZ=Z.grad,
                grads=grads, dtemp=temp_leaf.grad)


# ------------------------------------------------------------------------------------------------
# the per-batch preprocessing of Gwilliams2022Collator.forward
# (dataclass/gwilliams2022.py:653-661) = baseline_correction_single (utils/preproc_utils.py:128-142)
# followed by scaleAndClamp (utils/preproc_utils.py:69-90: one sklearn RobustScaler per sample, fit
# over time per channel).  The reference hands sklearn a torch tensor, which sklearn's input validation
# converts to FLOAT64 (a torch dtype is not a numpy float dtype), so everything between the baseline
# subtraction and the final `.to(torch.float)` is double arithmetic.  numpy restatement, checked bit for
# bit against the unmodified functions (tests/test_oracle.py, tests/golden/collator.npz):
#   baseline  b = mean_t(x[:L]) in float32 (torch);     y = x - b                      (float32)
#   center    = median_t(y) in float64: middle element, or mean of the two middle ones
#   quantiles q = a + (b - a) * g  (g < 0.5)  |  b - (b - a) * (1 - g)  (g >= 0.5), in float64, on the order
#             statistics at floor((n-1)*p), +1 with g = frac((n-1)*p), p = 0.25 / 0.75   (numpy "linear")
#   scale     = q75 - q25; scale < 10*eps64 -> 1                                       (sklearn _handle_zeros_in_scale)
#   out       = float32( (y - center) / scale ) with the division in float64, then clamp to +-clamp_lim
# ------------------------------------------------------------------------------------------------
def collate_preproc(X, baseline_len_samp, clamp_lim, clamp=True):
    """X: (B, C, T) float32 array/tensor -> float32 numpy array of the same shape."""
    x = np.asarray(X.detach().cpu().numpy() if isinstance(X, torch.Tensor) else X, dtype=np.float32)
    n = x.shape[-1]
    base = torch.from_numpy(x[..., :baseline_len_samp]).mean(dim=-1).numpy()          # torch float32 mean (:138)
    y = (x - base[..., None]).astype(np.float32).astype(np.float64)
    ys = np.sort(y, axis=-1)
    center = ys[..., n // 2] if n % 2 else (ys[..., n // 2 - 1] + ys[..., n // 2]) / 2.0

    def quantile(p):
        v = (n - 1) * p
        lo = int(np.floor(v))
        g = v - lo
        a = ys[..., lo]
        b = ys[..., min(lo + 1, n - 1)]
        d = b - a
        return a + d * g if g < 0.5 else b - d * (1 - g)

    scale = quantile(0.75) - quantile(0.25)
    scale = np.where(scale < 10 * np.finfo(np.float64).eps, 1.0, scale)
    out = ((y - center[..., None]) / scale[..., None]).astype(np.float32)
    if clamp:
        out = np.clip(out, -np.float32(clamp_lim), np.float32(clamp_lim))
    return out
