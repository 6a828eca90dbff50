"""Drop-in for the reference's speech_decoding/utils/loss.py (CLIPLoss :28-84,
MSELoss :16-25, torch_exp/torch_log :8-13; train.py:24 star-imports this
module).  CLIPLoss runs on the sm_100a kernels behind include/sd_b200.h:
one streaming similarity pass over the raw (un-normalised) rows, a fused
two-direction softmax cross-entropy on the (B x B) tile, and one streaming
pass for the gradient -- no normalised copies, no logits round trip through
autograd."""
import math

import torch
import torch.nn as nn

from sd_b200 import ops
from sd_b200 import dist as sd_dist


def torch_exp(x: torch.Tensor):
    return torch.exp(x.clamp(max=10))


def torch_log(x: torch.Tensor):
    return torch.log(x.clamp(min=1e-10))


class MSELoss(nn.Module):
    """Sum over (F, T), mean over the batch (loss.py:16-25).  Not on the hot
    path (train.py never calls it); kept importable, plain PyTorch."""

    def __init__(self):
        super().__init__()
        self.mse = nn.MSELoss(reduction="none")

    def forward(self, Y, Z):
        return self.mse(Y, Z).sum(dim=(-1, -2)).mean()


class _ClipFn(torch.autograd.Function):
    """loss, logits = clip(x, z, temp).  x = speech rows (global batch when
    data-parallel), z = local brain rows."""

    @staticmethod
    def forward(ctx, x, z, temp, reduction, use_temp, group, zn2_hint, xn2_hint):
        """x: fp32 rows, or (data-parallel bf16 transport) bf16 rows with their squared norms in xn2_hint."""
        M, Nn = x.shape[0], z.shape[0]
        world, rank = sd_dist.world_rank(group)
        with torch.cuda.device(x.device), ops.stream_scope():
            if x.dtype == torch.bfloat16:
                tc = True
                xn2 = xn2_hint
                zb, zn2 = ops.cast_rows_bf16(z)      # the brain rows as the bf16 GEMM sees them
                dots = ops.clip_dots(x, zb)
                del zb
            else:
                tc = ops.clip_tc_ok(x, z)          # bf16 mode: TF32 tensor-core GEMMs; fp32 mode: exact fp32
                xn2 = ops.rownorm2(x)
                zn2 = zn2_hint if zn2_hint is not None else ops.rownorm2(z)
                dots = ops.clip_dots(x, z, tc=tc)
            t = temp.detach() if use_temp else torch.zeros_like(temp)
            logits, row_stat, col_lse = ops.clip_phase1(dots, xn2, zn2, t)
            row_lse = sd_dist.global_row_lse(row_stat, group)     # one exchange through peer memory + one merge kernel
            scale = 1.0 / M if reduction == "mean" else 1.0
            if x.dtype == torch.bfloat16:      # the bf16 gradient GEMM makes its own transposed bf16 operand
                coef, cz, partial = ops.clip_phase2(logits, row_lse, col_lse, xn2, zn2, t, scale, rank * Nn)
                coef_t = coef                  # placeholder in the saved tuple
            else:
                coef, coef_t, cz, partial = ops.clip_phase2(logits, row_lse, col_lse, xn2, zn2, t, scale, rank * Nn,
                                                            want_t=True)
            if world > 1:
                partial = sd_dist.all_reduce_sum(partial, group)
        ctx.save_for_backward(x, z, coef, cz, partial, logits, xn2, zn2, t, coef_t)
        ctx.use_temp = use_temp
        ctx.tc = tc
        ctx.mark_non_differentiable(logits)
        return partial[0].clone(), logits

    @staticmethod
    def backward(ctx, gloss, _glogits):
        x, z, coef, cz, partial, logits, xn2, zn2, t, coef_t = ctx.saved_tensors
        dx = dz = dtemp = None
        with torch.cuda.device(x.device), ops.stream_scope():
            if ctx.needs_input_grad[1]:
                gs = gloss.detach().float().reshape(1).contiguous()
                if x.dtype == torch.bfloat16:
                    dz = ops.clip_dz_bf16(coef, cz, x, z, gs)
                else:
                    dz = ops.clip_dz_tc(coef_t, cz, x, z, gs) if ctx.tc else ops.clip_dz(coef, cz, x, z, gs)
            if ctx.needs_input_grad[0]:
                assert x.dtype == torch.float32, "speech-side gradient is not available with bf16 transport"
                # symmetric formula for the speech side (appendix A.5); rarely needed (Y carries no grad)
                gl = coef * logits * (xn2.sqrt()[:, None] * zn2.sqrt()[None, :]) / torch.exp(t)
                cx = gl.sum(dim=1) / xn2
                dx = ops.clip_dz(coef.t().contiguous(), cx.contiguous(), z, x, gloss.detach().float().reshape(1).contiguous())
            if ctx.needs_input_grad[2] and ctx.use_temp:
                dtemp = (partial[1] * gloss).reshape(1)
        return dx, dz, dtemp, None, None, None, None, None


class CLIPLoss(nn.Module):
    """Symmetric InfoNCE with a learned temperature (loss.py:28-84).

    forward(x, y, fast=True, return_logits=False): x = speech embeddings
    (B,F,T), y = brain embeddings (B,F,T) at the reference call site
    train.py:191.  `fast=False` is the reference's cosine / no-temperature
    variant (loss.py:46-50) whose logits are transposed.

    Data-parallel: after `sd_b200.dist.attach(loss, group)` x is all-gathered
    so every rank scores its local brain rows against the global speech batch."""

    def __init__(self, args):
        super().__init__()
        self.compute_similarity = nn.CosineSimilarity(dim=-1)
        self.reduction = args.reduction
        self._criterion = nn.CrossEntropyLoss(reduction=args.reduction)
        self.temp = nn.Parameter(torch.tensor([float(args.init_temperature)]))
        self.process_group = None
        self._prefetched = None

    def forward(self, x, y, fast=True, return_logits=False):
        batch_size = x.size(0)
        assert batch_size > 1, "Batch size must be greater than 1."          # loss.py:40
        if self.reduction not in ("mean", "sum"):
            raise NotImplementedError("sd_b200 CLIPLoss supports reduction='mean' or 'sum'")
        ops.require_cuda(x, "x")
        ops.require_cuda(y, "y")
        # speech rows that arrive in bf16 (a frozen wav2vec2 embedding rounded once on the host: half the H2D bytes) feed
        # the bf16 similarity / gradient GEMMs as they are in the bf16 mode; every other combination computes on fp32 rows
        x_bf16 = (x.dtype == torch.bfloat16 and not x.requires_grad and bool(fast) and ops.get_precision() == "bf16"
                  and ops.clip_bf16_ok(x.reshape(batch_size, -1)))
        xf = x.reshape(batch_size, -1).contiguous() if x_bf16 else x.reshape(batch_size, -1).float().contiguous()
        yf = y.reshape(batch_size, -1).float().contiguous()
        group = self.process_group
        xn2 = None
        if group is not None:
            pre = getattr(self, "_prefetched", None)
            self._prefetched = None
            if pre is not None and pre[0] == x.data_ptr() and pre[1] == tuple(x.shape):
                for work in pre[2]:
                    work.wait()                  # gather launched before the encoder forward (DataParallel.prefetch_targets)
                xf, xn2 = pre[3], pre[4]
            else:
                xf, xn2 = sd_dist.gather_speech_rows(xf, group, not x.requires_grad)
        elif x_bf16:
            with torch.cuda.device(x.device), ops.stream_scope():
                xn2 = ops.rownorm2_bf16(xf)
        zn2 = getattr(y, "_sd_norm2", None)      # set by BrainEncoder.forward (fused into its last epilogue)
        if zn2 is not None and (zn2.shape[0] != batch_size or y.dtype != torch.float32):
            zn2 = None
        loss, logits = _ClipFn.apply(xf, yf, self.temp, self.reduction, bool(fast), group, zn2, xn2)
        if return_logits:
            return (logits if fast else logits.t()), loss
        return loss
