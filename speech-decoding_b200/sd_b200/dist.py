"""Data-parallel plumbing (one process per GPU, torch.distributed over NCCL).

The reference is single-device (train.py:31); batch sharding is this build's
addition:

  * every rank runs the full encoder on its shard of the batch;
  * CLIP negatives span the global batch: the speech rows x (input data, no
    gradient) are all-gathered, each rank scores them against its local brain
    rows z, column statistics stay local and only the (M,2) row statistics
    (max, sum-exp) and two scalars cross ranks -- no gradient exchange for dz;
  * parameter gradients are summed across ranks (the loss is the global-batch
    loss, so SUM, not average), launched per stage from inside backward so the
    all-reduce overlaps the remaining backward kernels;
  * optional SyncBN (per-channel statistics all-reduced) so the result equals
    the single-process run on the concatenated batch;
  * subject presence (which per-subject weights receive a gradient at all) is
    agreed on the host over a gloo side-channel, so no device sync is needed.

Everything here is host logic over torch.distributed and runs unchanged on the
gloo backend with CPU tensors (tests/test_dist_cpu.py).
"""
import numpy as np
import torch
import torch.distributed as dist


def world_rank(group=None):
    if group is None or not dist.is_available() or not dist.is_initialized():
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


def all_gather_rows(x, group):
    world, _ = world_rank(group)
    if world == 1:
        return x
    out = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x.contiguous(), group=group)
    return out


def gather_speech_rows(x2d, group, allow_bf16=True, async_op=False):
    """All-gather the speech rows of every rank.  In the bf16 mode (and when no gradient flows to them) the rows are
    rounded to bf16 ONCE on the owning rank together with their squared norms: the gather then moves, and the two
    CLIP GEMMs re-read, half the bytes.  Returns (rows, norms-or-None[, works])."""
    from . import ops
    world, _ = world_rank(group)
    works = []
    if world > 1 and allow_bf16 and x2d.is_cuda and ops.clip_bf16_ok(x2d):
        with torch.cuda.device(x2d.device), ops.stream_scope():
            xb, n2 = ops.cast_rows_bf16(x2d)
        rows = torch.empty((world * xb.shape[0], xb.shape[1]), dtype=xb.dtype, device=xb.device)
        norms = torch.empty((world * n2.shape[0],), dtype=n2.dtype, device=n2.device)
        works.append(dist.all_gather_into_tensor(norms, n2, group=group, async_op=async_op))
        works.append(dist.all_gather_into_tensor(rows, xb, group=group, async_op=async_op))
        keep = (xb, n2)
    else:
        rows = torch.empty((world * x2d.shape[0], x2d.shape[1]), dtype=x2d.dtype, device=x2d.device)
        norms = None
        works.append(dist.all_gather_into_tensor(rows, x2d.contiguous(), group=group, async_op=async_op))
        keep = (x2d,)
    if async_op:
        return rows, norms, works, keep
    return rows, norms


def all_reduce_sum(t, group):
    t = t.contiguous()
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def merge_row_stats(row_stat, group):
    """(M,2) per-rank (max_j, sum_j exp(l - max)) over local columns ->
    the same statistics over the columns of all ranks."""
    m_local = row_stat[:, 0].contiguous()
    m_glob = m_local.clone()
    dist.all_reduce(m_glob, op=dist.ReduceOp.MAX, group=group)
    s = (row_stat[:, 1] * torch.exp(m_local - m_glob)).contiguous()
    dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)
    return torch.stack([m_glob, s], dim=1)


def gather_host_ints(values, host_group):
    """all-gather a small int64 numpy vector over the host (gloo) side channel."""
    world, _ = world_rank(host_group)
    t = torch.from_numpy(np.asarray(values, dtype=np.int64))
    if world == 1:
        return [t.numpy()]
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t, group=host_group)
    return [o.numpy() for o in outs]


class GradReducer:
    """Sums parameter gradients across ranks, stage by stage, while backward
    is still running (hooked into engine.Pipeline)."""

    def __init__(self, group):
        self.group = group
        self.pending = []

    def stage_done(self, flat, lo, hi):
        """All-reduce flat[lo:hi] (one stage's parameter gradients, contiguous in the GradPool --
        absent subjects included as zeros, so the message size is the same on every rank)."""
        if hi <= lo:
            return
        work = dist.all_reduce(flat[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self.pending.append(work)

    def finish(self):
        for work in self.pending:
            work.wait()
        self.pending = []


class DataParallel:
    """Wire a BrainEncoder + CLIPLoss pair for batch-sharded training.

        dp = DataParallel(encoder, loss_fn, group=None, sync_bn=True)
        Z = encoder(X_local, ids_local); loss = loss_fn(Y_local, Z); loss.backward()

    `loss` is the global-batch loss on every rank; parameter .grad's are the
    global-batch gradients (identical on every rank)."""

    def __init__(self, encoder, loss_fn, group=None, sync_bn=True, host_group=None, reserve_sms=None):
        """reserve_sms: SMs left free for the NCCL kernels that run concurrently with the step (speech-row gather
        during the encoder forward, gradient all-reduces during backward).  The conv / wgrad / CLIP kernels are
        persistent, one CTA per SM: if a collective holds even one SM, the CTAs that cannot be placed start only
        when others finish and the launch takes two waves.  Default: SD_B200_DP_RESERVE_SMS or 0 (no cap)."""
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        import os
        if reserve_sms is None:
            reserve_sms = int(os.environ.get("SD_B200_DP_RESERVE_SMS", "0"))
        self.reserve_sms = int(reserve_sms)
        if self.reserve_sms > 0 and torch.cuda.is_available():
            from . import _native as nat
            sms = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
            nat.call("sd_set_sm_limit", max(2, sms - self.reserve_sms))
        self.group = group if group is not None else dist.group.WORLD
        self.host_group = host_group
        if host_group is None and dist.get_backend(self.group) != "gloo":
            self.host_group = dist.new_group(backend="gloo")
        elif host_group is None:
            self.host_group = self.group
        pipe = encoder.pipeline()
        pipe.reducer = GradReducer(self.group)
        pipe.bn_group = self.group if sync_bn else None
        pipe.host_group = self.host_group
        loss_fn.process_group = self.group
        self.loss_fn = loss_fn
        loss_fn._prefetched = None

    def prefetch_targets(self, Y):
        """Start the all-gather of the speech embeddings for the coming step NOW (they are input data, known
        before the encoder runs), so the transfer over NVLink overlaps the encoder forward instead of sitting
        between the encoder and the loss.  Call right before `encoder(X, ids)`; `loss_fn(Y, Z)` with the same
        Y then picks the gathered tensor up.  Optional: without it the loss gathers synchronously."""
        world, _ = world_rank(self.group)
        if world == 1:
            return
        y2 = Y.reshape(Y.shape[0], -1)
        if y2.dtype != torch.float32 or not y2.is_contiguous():
            y2 = y2.float().contiguous()
        rows, norms, works, keep = gather_speech_rows(y2, self.group, not Y.requires_grad, async_op=True)
        self.loss_fn._prefetched = (Y.data_ptr(), tuple(Y.shape), works, rows, norms, keep)
        # the temperature gradient is a partial sum per rank: the loss Function already
        # all-reduces `partial`, so dtemp is global -- nothing more to do for it.
