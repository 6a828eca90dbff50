"""Batch preprocessing on the GPU: what the reference's `Gwilliams2022Collator.forward`
(speech_decoding/dataclass/gwilliams2022.py:653-661) does per batch on the host -- a Python loop of
`baseline_correction_single` (utils/preproc_utils.py:128-142) and one sklearn `RobustScaler` fit per sample
(`scaleAndClamp`, utils/preproc_utils.py:69-90) -- as ONE kernel launch over the stacked batch.

    collate = GpuCollator(args)                      # same `args.preprocs` fields as Gwilliams2022Collator
    X, Y, subject_idxs = collate(batch)              # list of (x, y, subject) items -> X on the GPU, normalised
    X = baseline_scale_clamp(X_raw_on_gpu, 60, 20.0) # or directly on a stacked (B, C, T) device tensor

DataLoader workers cannot touch CUDA, so with `num_workers > 0` keep a stack-only `collate_fn` in the loader and
call `baseline_scale_clamp` on the batch after `X.to(device)` (one line in train.py:187-189)."""
import torch
import torch.nn as nn

from . import _native as nat
from . import ops


def baseline_scale_clamp(X, baseline_len_samp, clamp_lim=20.0, clamp=True, out=None):
    """X: (B, C, T) fp32 CUDA tensor -> same shape, baseline-corrected, robust-scaled and clamped per (sample, channel)."""
    ops.require_cuda(X, "X")
    if X.dtype != torch.float32:
        raise TypeError("baseline_scale_clamp expects float32, got %s" % X.dtype)
    if X.dim() != 3:
        raise ValueError("baseline_scale_clamp expects (batch, channels, time)")
    X = X.contiguous()
    B, C, T = X.shape
    if out is None:
        out = torch.empty_like(X)
    with torch.cuda.device(X.device), ops.stream_scope():
        nat.call("sd_collate_preproc", X.data_ptr(), out.data_ptr(), B * C, T, int(baseline_len_samp), float(clamp_lim),
                 int(bool(clamp)), ops._st())
    return out


class GpuCollator(nn.Module):
    """Same constructor fields and batch contract as the reference's Gwilliams2022Collator; the returned X lives on
    `device` (train.py's later `X.to(device)` is then a no-op), Y and subject_idxs are returned as the reference does."""

    def __init__(self, args, device="cuda"):
        super().__init__()
        self.brain_resample_rate = args.preprocs["brain_resample_rate"]
        self.baseline_len_samp = int(self.brain_resample_rate * args.preprocs["baseline_len_sec"])
        self.clamp = args.preprocs["clamp"]
        self.clamp_lim = args.preprocs["clamp_lim"]
        self.device = torch.device(device)

    def forward(self, batch):
        X = torch.stack([item[0] for item in batch])
        Y = torch.stack([item[1] for item in batch])
        subject_idx = torch.IntTensor([item[2] for item in batch])
        X = baseline_scale_clamp(X.to(self.device, dtype=torch.float32, non_blocking=True), self.baseline_len_samp,
                                 self.clamp_lim, self.clamp)
        return X, Y, subject_idx
