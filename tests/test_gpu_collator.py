"""GPU: the batch-preprocessing kernel (sd_collate_preproc) against the golden outputs of the
reference's baseline_correction_single + scaleAndClamp and against the oracle at the full cfg2 batch size.

Tolerance: everything after the baseline subtraction is float64 in the reference (sklearn) and reproduced exactly;
the baseline itself is torch's float32 mean, whose summation order is not specified -- the kernel sums in float64 and
rounds once, so the baseline can differ by 1 ulp(float32), i.e. <= 5e-7 * |offset| / IQR in the output."""
import os

import numpy as np
import pytest
import torch

from oracle import restate

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLD = os.path.join(os.path.dirname(__file__), "golden", "collator.npz")


@pytest.mark.parametrize("name", ["gw", "odd", "noclamp", "long"])
def test_collator_kernel_matches_reference_fixture(name):
    from sd_b200.preproc import baseline_scale_clamp
    g = np.load(GOLD)
    L, lim, clamp = g[name + "_cfg"]
    x = torch.from_numpy(g[name + "_x"]).to(DEV)
    y = baseline_scale_clamp(x, int(L), float(lim), bool(clamp)).cpu().numpy()
    ref = g[name + "_y"]
    assert np.isfinite(y).all()
    assert np.abs(y - ref).max() <= 5e-6 * max(1.0, float(np.abs(ref).max()))
    assert (y == ref).mean() > 0.5                      # rows whose baseline rounds like torch's are bit-identical


def test_collator_kernel_full_batch_vs_oracle_and_properties():
    from sd_b200.preproc import baseline_scale_clamp, GpuCollator
    rng = np.random.default_rng(0)
    B, C, T = 256, 208, 360                             # BASELINE.json cfg2 batch
    x = (rng.standard_normal((B, C, T)) * np.exp(0.5 * rng.standard_normal((B, C, 1))) + rng.standard_normal((B, C, 1))).astype(np.float32)
    xd = torch.from_numpy(x).to(DEV)
    y = baseline_scale_clamp(xd, 60, 20.0, True)
    yc = y.cpu().numpy()
    ref = restate.collate_preproc(x[:32], 60, 20.0, True)
    assert np.abs(yc[:32] - ref).max() <= 5e-6 * max(1.0, float(np.abs(ref).max()))
    # size-independent properties on the whole batch: median 0 and IQR 1 per row (nothing clamps at 20 here)
    med = np.median(yc.astype(np.float64), axis=-1)
    q = np.percentile(yc.astype(np.float64), [25, 75], axis=-1)
    assert np.abs(med).max() < 1e-6 and np.abs((q[1] - q[0]) - 1.0).max() < 1e-5
    # translation / positive-scale invariance of the whole transform
    y2 = baseline_scale_clamp(xd * 4.0 + 1.0, 60, 20.0, True)
    assert float((y2 - y).abs().max()) < 2e-5
    # in place, and the clamp
    z = xd.clone()
    baseline_scale_clamp(z, 60, 0.5, True, out=z)
    assert float(z.abs().max()) <= 0.5 and torch.equal(z, y.clamp(-0.5, 0.5))
    # the collator module: same constructor fields and batch contract as Gwilliams2022Collator
    class A:
        preprocs = {"brain_resample_rate": 120, "baseline_len_sec": 0.5, "clamp": True, "clamp_lim": 20}
    items = [(torch.from_numpy(x[i]), torch.zeros(4, T), i % 27) for i in range(8)]
    Xb, Yb, ids = GpuCollator(A())(items)
    assert Xb.is_cuda and torch.equal(Xb, y[:8]) and Yb.shape == (8, 4, T) and ids.dtype == torch.int32
    with pytest.raises(RuntimeError):
        baseline_scale_clamp(torch.zeros(1, 2, 4096, device=DEV), 60)      # T > 2048 is rejected loudly


def test_collator_kernel_edge_cases():
    from sd_b200.preproc import baseline_scale_clamp
    x = torch.zeros(2, 3, 1, device=DEV)                # T = 1: median = the sample, IQR = 0 -> scale 1
    assert torch.equal(baseline_scale_clamp(x + 3.0, 1), torch.zeros_like(x))
    x = torch.arange(10, dtype=torch.float32, device=DEV).reshape(1, 1, 10)
    ref = restate.collate_preproc(x.cpu(), 3, 20.0, True)
    assert np.array_equal(baseline_scale_clamp(x, 3).cpu().numpy(), ref)
    assert baseline_scale_clamp(torch.zeros(0, 5, 16, device=DEV), 4).shape == (0, 5, 16)   # empty batch
    with pytest.raises(RuntimeError):
        baseline_scale_clamp(torch.zeros(1, 1, 8, device=DEV), 9)                            # baseline longer than the row
    with pytest.raises(RuntimeError):
        baseline_scale_clamp(torch.zeros(1, 1, 8), 4)                                        # CPU tensors raise
