"""Drop-in for the reference's speech_decoding/models.py.

Same class names, constructor arguments, forward signatures, parameter/buffer
names, shapes and dtypes (so `state_dict()`s are interchangeable and the
reference's train.py runs unchanged: train.py:22,148-150,189,193,225,259), but
every forward/backward is executed by hand-written sm_100a kernels behind the C
ABI in include/sd_b200.h.  `nn.Conv1d` / `nn.BatchNorm1d` objects are kept only
as parameter containers (identical initialisers and state_dict keys); their own
forward is never called.  There is no CPU path: CPU tensors raise.

Precision: sd_b200.set_precision("bf16" | "fp32") or SD_B200_PRECISION.
"""
import numpy as np
import torch
import torch.nn as nn

from sd_b200 import engine, ops
from speech_decoding.utils.layout import ch_locations_2d


class SpatialDropout(nn.Module):
    """One drop centre per forward for the whole batch, drawn from numpy's
    global RNG like the reference (models.py:77-86) so a seeded run makes the
    same draws.  No rescaling by the keep probability."""

    def __init__(self, loc, d_drop):
        super().__init__()
        self.loc = loc                      # (num_channels, 2), plain CPU attribute (models.py:73)
        self.d_drop = d_drop
        self.num_channels = loc.shape[0]
        self._table = None                  # (C, C) device table: row c = mask for centre c
        self._pipe = None

    def _mask_table(self, device):
        if self._table is None or self._table.device != device:
            loc = self.loc.detach().cpu().float()
            dist = (loc[:, None, :] - loc[None, :, :]).norm(dim=-1)          # models.py:82
            self._table = torch.where(dist < self.d_drop, 0.0, 1.0).to(device)  # models.py:83
        return self._table

    def draw_mask(self, device):
        centre = np.random.randint(self.num_channels)                        # models.py:81
        return self._mask_table(device)[centre]

    def forward(self, X):
        assert X.shape[1] == self.num_channels                               # models.py:78
        if not self.training:
            return X
        ops.require_cuda(X, "X")
        mask = self.draw_mask(X.device)
        if self._pipe is None:
            self._pipe = engine.Pipeline([engine.ToBTC(), _ChannelScale(self), engine.ToNCT(self.num_channels)])
        self._pipe.stages[1].mask = mask
        return self._pipe.run(X)


class _ChannelScale(engine.Stage):
    """X * mask[None,:,None] as a 1x1 conv with a diagonal weight (stand-alone SpatialDropout only;
    inside SpatialAttention the mask is folded into the mixing weights instead)."""

    def __init__(self, owner):
        self.owner = owner
        self.mask = None

    def _w(self, run, device):
        C = self.owner.num_channels
        Cp = ops.rup8(C)
        w = torch.zeros((1, 1, Cp, Cp), dtype=run.dtype, device=device)
        w[0, 0, :C, :C] = torch.diag(self.mask).to(run.dtype)
        return w, C

    def forward(self, run, x, sv):
        w, C = self._w(run, x.device)
        out = torch.empty_like(x)
        ops.conv_fwd(x, w, K=C, N=C, out=out)
        if sv is not None:
            sv["w"] = w
        return out

    def backward(self, run, sv, dout, grads, need_dx):
        if not need_dx:
            return None
        dx = torch.empty_like(dout)
        C = self.owner.num_channels
        ops.conv_fwd(dout, sv["w"], K=C, N=C, out=dx)          # diagonal: its own transpose
        return dx


class SpatialAttention(nn.Module):
    """Fourier-parameterised spatial attention (models.py:14-65)."""

    def __init__(self, args):
        super().__init__()
        K = args.K
        grid = torch.arange(K * K)
        k, l = grid // K, grid % K                                # row m = (k, l)   models.py:21-26
        loc = ch_locations_2d(args)                               # models.py:29
        x, y = loc[:, 0], loc[:, 1]
        self.z = nn.Parameter(torch.rand(size=(args.D1, K ** 2), dtype=torch.cfloat))   # models.py:33
        phi = 2 * torch.pi * (k[:, None] * x[None, :] + l[:, None] * y[None, :])        # models.py:36-38
        self.register_buffer("cos", torch.cos(phi))
        self.register_buffer("sin", torch.sin(phi))
        self.spatial_dropout = SpatialDropout(loc, args.d_drop)   # models.py:43
        self._pipe = None

    def forward(self, X):
        """X: (batch_size, num_channels, T) -> (batch_size, D1, T)"""
        if self._pipe is None:
            self._pipe = engine.Pipeline([engine.SpatialAttentionStage(self), engine.ToNCT(self.z.shape[0])])
        return self._pipe.run(X)


class SubjectBlock(nn.Module):
    """SpatialAttention -> Conv1d(1x1, bias) -> per-subject Conv1d(1x1, no bias)  (models.py:89-117)."""

    def __init__(self, args):
        super().__init__()
        self.num_subjects = args.num_subjects
        self.D1 = args.D1
        self.K = args.K
        self.spatial_attention = SpatialAttention(args)
        self.conv = nn.Conv1d(in_channels=self.D1, out_channels=self.D1, kernel_size=1, stride=1)
        self.subject_layer = nn.ModuleList(
            [nn.Conv1d(in_channels=self.D1, out_channels=self.D1, kernel_size=1, bias=False, stride=1)
             for _ in range(self.num_subjects)])
        self._pipe = None

    def _stages(self):
        return [engine.SpatialAttentionStage(self.spatial_attention), engine.SubjectStage(self)]

    def forward(self, X, subject_idxs):
        if self._pipe is None:
            self._pipe = engine.Pipeline(self._stages() + [engine.ToNCT(self.D1)])
        ids = engine.normalize_subject_ids(subject_idxs, self.num_subjects)
        return self._pipe.run(X, ids)


class ConvBlock(nn.Module):
    """Residual dilated conv block with BatchNorm, GELU and a GLU output (models.py:120-166)."""

    def __init__(self, k, D1, D2):
        super().__init__()
        self.k = k
        self.D2 = D2
        self.in_channels = D1 if k == 0 else D2
        self.conv0 = nn.Conv1d(self.in_channels, self.D2, kernel_size=3, padding="same",
                               dilation=2 ** ((2 * k) % 5))
        self.batchnorm0 = nn.BatchNorm1d(num_features=self.D2)
        self.conv1 = nn.Conv1d(self.D2, self.D2, kernel_size=3, padding="same",
                               dilation=2 ** ((2 * k + 1) % 5))
        self.batchnorm1 = nn.BatchNorm1d(num_features=self.D2)
        self.conv2 = nn.Conv1d(self.D2, 2 * self.D2, kernel_size=3, padding="same", dilation=2)
        self._pipe = None

    def forward(self, X):
        if self._pipe is None:
            self._pipe = engine.Pipeline([engine.ToBTC(), engine.ConvBlockStage(self), engine.ToNCT(self.D2)])
        return self._pipe.run(X)


class BrainEncoder(nn.Module):
    """M/EEG (B, C, T) -> latent (B, F, T)   (models.py:169-196)."""

    def __init__(self, args):
        super().__init__()
        self.num_subjects = args.num_subjects
        self.D1 = args.D1
        self.D2 = args.D2
        self.F = args.F if not args.preprocs["last4layers"] else 1024       # models.py:176
        self.K = args.K
        self.dataset_name = args.dataset
        self.subject_block = SubjectBlock(args)
        self.conv_blocks = nn.Sequential()
        for k in range(5):
            self.conv_blocks.add_module(f"conv{k}", ConvBlock(k, self.D1, self.D2))
        self.conv_final1 = nn.Conv1d(in_channels=self.D2, out_channels=2 * self.D2, kernel_size=1)
        self.conv_final2 = nn.Conv1d(in_channels=2 * self.D2, out_channels=self.F, kernel_size=1)
        self._pipe = None

    def pipeline(self):
        if self._pipe is None:
            stages = self.subject_block._stages()
            stages += [engine.ConvBlockStage(b) for b in self.conv_blocks]
            stages += [engine.FinalStage(self)]
            self._pipe = engine.Pipeline(stages)
        return self._pipe

    def forward(self, X, subject_idxs):
        ids = engine.normalize_subject_ids(subject_idxs, self.num_subjects)
        pipe = self.pipeline()
        Z = pipe.run(X, ids)
        if pipe.out_norm2 is not None:
            # |Z_b|^2 came for free from the last conv's epilogue; CLIPLoss picks it up (one pass over Z saved)
            Z._sd_norm2 = pipe.out_norm2
        return Z


class Classifier(nn.Module):
    """Retrieval accuracy of brain latents Z against speech latents Y
    (models.py:199-248): similarity[i,j] = cos(Z_i, Y_j) with the
    max(|x||y|, 1e-8) guard, transposed; top-1 / top-10 hit rate of the
    diagonal.  The B^2 Python loop of the reference is one similarity GEMM."""

    def __init__(self, args):
        super().__init__()
        self.factor = 1

    @torch.no_grad()
    def forward(self, Z: torch.Tensor, Y: torch.Tensor, test=False):
        batch_size = Z.size(0)
        ops.require_cuda(Z, "Z")
        x = Z.reshape(batch_size, -1).float().contiguous()
        y = Y.reshape(batch_size, -1).float().contiguous()
        dots = ops.clip_dots(x, y)
        den = torch.clamp(ops.rownorm2(x).sqrt()[:, None] * ops.rownorm2(y).sqrt()[None, :], min=1e-8)
        similarity = (dots / den).T                                               # models.py:228,233
        diags = torch.arange(batch_size, device=Z.device)
        top1accuracy = (similarity.argmax(dim=1) == diags).to(torch.float).mean().item()
        top10 = torch.topk(similarity, 10, dim=1, largest=True)[1]               # raises for B < 10 like the reference
        top10accuracy = float((top10 == diags[:, None]).any(dim=1).to(torch.float).mean().item())
        return top1accuracy, top10accuracy
