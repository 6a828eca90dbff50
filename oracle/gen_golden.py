"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz from the UNMODIFIED
reference (run in the build container, where /root/reference exists):

    python oracle/gen_golden.py

Each fixture holds: config, inputs (X, Y, subject ids, dropout centre drawn
from numpy's global RNG exactly as models.py:81 does), the reference's initial
state_dict, and the reference's outputs for one train-mode fwd+bwd
(Z, logits, loss, every parameter gradient, BN buffers after the step) plus an
eval-mode forward with the updated buffers and Classifier's (top1, top10).
The fixtures pin oracle/restate.py (CPU tests) and the CUDA path (GPU tests).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_import            # noqa: E402
from oracle.restate import make_args, synthetic_layout   # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

CONFIGS = {
    # name: (B, C, T, D1, D2, F, K, S, dataset, reduction, seed)
    "tiny_gwilliams": dict(B=12, C=10, T=24, D1=12, D2=16, F=24, K=3, S=4, dataset="Gwilliams2022",
                           reduction="mean", seed=1, d_drop=0.25),
    "small_brennan": dict(B=16, C=60, T=72, D1=30, D2=40, F=64, K=5, S=7, dataset="Brennan2018",
                          reduction="mean", seed=2, d_drop=0.1),
    "odd_sum": dict(B=11, C=13, T=37, D1=9, D2=10, F=17, K=2, S=3, dataset="Gwilliams2022",
                    reduction="sum", seed=3, d_drop=0.3),
}


def to_np(t):
    t = t.detach().cpu()
    if t.is_complex():
        return torch.view_as_real(t).numpy()
    return t.numpy()


def run(name, c):
    torch.manual_seed(c["seed"])
    np.random.seed(c["seed"])
    args = make_args(D1=c["D1"], D2=c["D2"], F_=c["F"], K=c["K"], d_drop=c["d_drop"],
                     num_subjects=c["S"], dataset=c["dataset"], num_channels=c["C"],
                     last4layers=False, reduction=c["reduction"], layout_seed=c["seed"])
    M, L = ref_import.load(lambda a: synthetic_layout(a.num_channels, a.layout_seed))
    enc = M.BrainEncoder(args)
    crit = L.CLIPLoss(args)
    # make BN affine / biases non-trivial so every gradient path is exercised
    with torch.no_grad():
        for n, p in enc.named_parameters():
            if "batchnorm" in n:
                p.add_(0.3 * torch.randn_like(p))
    sd0 = {k: v.clone() for k, v in enc.state_dict().items()}
    B, C, T = c["B"], c["C"], c["T"]
    X = torch.randn(B, C, T).clamp(-20, 20)
    Y = torch.randn(B, c["F"], T)
    ids = torch.randint(0, c["S"], (B,), dtype=torch.int32)
    if c["S"] > 2:
        ids[ids == c["S"] - 1] = 0          # guarantee an absent subject (grad None path)
    out = {"cfg_" + k: np.array(v) for k, v in c.items()}
    out["X"], out["Y"], out["ids"] = to_np(X), to_np(Y), ids.numpy()
    out["loc"] = to_np(synthetic_layout(C, c["seed"]))
    for k, v in sd0.items():
        out["sd0/" + k] = to_np(v)

    # ---- train-mode step; replicate the numpy draw of models.py:81 ----
    enc.train(); crit.train()
    rng_state = np.random.get_state()
    center = np.random.randint(C)
    np.random.set_state(rng_state)          # the forward below makes the same draw
    out["drop_center"] = np.array(center)
    Z = enc(X, ids)
    Z.retain_grad()
    logits, loss = crit(Y, Z, return_logits=True)
    loss.backward()
    out["Z"], out["dZ"], out["logits"], out["loss"] = to_np(Z), to_np(Z.grad), to_np(logits), to_np(loss)
    out["dtemp"] = to_np(crit.temp.grad)
    out["temp"] = to_np(crit.temp)
    absent = []
    for n, p in enc.named_parameters():
        if p.grad is None:
            absent.append(n)
        else:
            out["grad/" + n] = to_np(p.grad)
    out["absent_grads"] = np.array(absent)
    for k, v in enc.state_dict().items():
        if "running" in k or "num_batches" in k:
            out["sd1/" + k] = to_np(v)

    # ---- other CLIPLoss call forms (loss.py:46-50, :81-84) ----
    with torch.no_grad():
        out["loss_slow"] = to_np(crit(Y, Z, fast=False))
        top1, top10 = M.Classifier(args)(Z.detach().contiguous(), Y)
        out["top1"], out["top10"] = np.array(top1), np.array(top10)

    # ---- eval-mode forward with the updated running statistics ----
    enc.eval()
    with torch.no_grad():
        Ze = enc(X, ids)
        out["Z_eval"] = to_np(Ze)
        out["loss_eval"] = to_np(crit(Y, Ze))
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "->", path, "%.1f KB" % (os.path.getsize(path) / 1024), "loss", float(loss), "center", center)


if __name__ == "__main__":
    for n, c in CONFIGS.items():
        run(n, c)
