"""Helpers shared by CPU and GPU tests: load tests/golden/*.npz (outputs of the
unmodified reference, made by oracle/gen_golden.py)."""
import glob
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def names():
    # hot-path (encoder + CLIP) fixtures; collator.npz is the batch-preprocessing fixture (tests/test_gpu_collator.py)
    return sorted(n for n in (os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
                  if n != "collator")


def _t(a, key):
    t = torch.from_numpy(np.array(a))
    if key.endswith(".z") and t.dim() == 3 and t.shape[-1] == 2:
        t = torch.view_as_complex(t.contiguous())
    return t


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    g = {"cfg": {}, "sd0": {}, "sd1": {}, "grad": {}}
    for k in z.files:
        if k.startswith("cfg_"):
            v = z[k]
            g["cfg"][k[4:]] = v.item() if v.shape == () else v
        elif k.startswith("sd0/"):
            g["sd0"][k[4:]] = _t(z[k], k)
        elif k.startswith("sd1/"):
            g["sd1"][k[4:]] = _t(z[k], k)
        elif k.startswith("grad/"):
            g["grad"][k[5:]] = _t(z[k], k)
        elif k == "absent_grads":
            g[k] = [str(s) for s in z[k].tolist()]
        else:
            g[k] = torch.from_numpy(np.array(z[k]))
    return g


def rel_err(a, b, floor=0.0):
    """max|a-b| / max(max|b|, floor) -- relative error with an absolute floor
    (some gradients are exactly zero by construction)."""
    a = torch.view_as_real(a) if a.is_complex() else a
    b = torch.view_as_real(b) if b.is_complex() else b
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    den = max(float(b.abs().max()), floor, 1e-30)
    return float((a - b).abs().max()) / den


def rel_l2(a, b, floor=0.0):
    """||a-b||_2 / max(||b||_2, floor): the normwise relative error used for the bf16 mode."""
    a = torch.view_as_real(a) if a.is_complex() else a
    b = torch.view_as_real(b) if b.is_complex() else b
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).norm()) / max(float(b.norm()), floor, 1e-30)
