"""TEST INFRASTRUCTURE ONLY.  Golden vectors for the batch preprocessing of the reference's
Gwilliams2022Collator: runs the UNMODIFIED baseline_correction_single and scaleAndClamp of
/root/reference/speech_decoding/utils/preproc_utils.py (sklearn RobustScaler) on seeded inputs and writes
tests/golden/collator.npz.  Run in the build container:  python oracle/gen_golden_collator.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import   # noqa: E402


def main():
    ref = ref_import.load_preproc_utils()
    rng = np.random.default_rng(5)
    out = {}
    # (name, B, C, T, baseline_len, clamp_lim, clamp)
    cases = [("gw", 3, 7, 360, 60, 20.0, True), ("odd", 2, 5, 97, 11, 3.0, True), ("noclamp", 2, 4, 128, 16, 20.0, False),
             ("long", 1, 3, 1200, 60, 20.0, True)]
    for name, B, C, T, L, lim, clamp in cases:
        x = rng.standard_normal((B, C, T)) * np.exp(rng.standard_normal((B, C, 1))) + 3.0 * rng.standard_normal((B, C, 1))
        x[:, 0, ::7] *= 40.0                      # outliers that hit the clamp
        if name == "odd":
            x[0, 1, :] = 0.25                     # constant channel: IQR = 0 -> scale 1
            x[1, 2, :50] = x[1, 2, 0]             # many ties
        x = torch.from_numpy(x.astype(np.float32))
        y = ref.baseline_correction_single(x.clone(), L)
        y = ref.scaleAndClamp(y, lim, clamp)
        out[name + "_x"] = x.numpy()
        out[name + "_y"] = y.numpy().astype(np.float32)
        out[name + "_cfg"] = np.array([L, lim, float(clamp)], dtype=np.float64)
    path = os.path.join(ROOT, "tests", "golden", "collator.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
