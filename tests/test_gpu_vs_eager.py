"""GPU: the native path against stock PyTorch eager ON THE SAME B200 (SURVEY §8d "the honest competitor") at the
BASELINE.json cfg2 size.  The eager arm is the oracle restatement (the same stock ops as the reference modules:
cuDNN conv1d, BatchNorm, F.gelu, F.glu, matmul, CrossEntropyLoss) run (a) in fp32 with TF32 allowed and (b) under
torch.autocast(bfloat16).  The native bf16 step must be clearly faster than both; the timings are written to
gpurun_out/eager_vs_native.json when that directory exists."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import restate

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _time(fn, warm, iters):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def test_native_step_beats_stock_eager_on_the_same_gpu():
    import sd_b200
    from speech_decoding.models import BrainEncoder
    from speech_decoding.utils.loss import CLIPLoss
    sd_b200.set_precision("bf16")
    torch.manual_seed(0)
    np.random.seed(0)
    args = restate.make_args()
    enc, crit = BrainEncoder(args).to(DEV).train(), CLIPLoss(args).to(DEV).train()
    B = 256
    X = torch.randn(B, 208, 360, device=DEV).clamp(-20, 20)
    Y = torch.randn(B, 1024, 360, device=DEV)
    ids = torch.randint(0, 27, (B,), dtype=torch.int32)
    params = list(enc.parameters()) + list(crit.parameters())

    def native():
        Z = enc(X, ids)
        loss = crit(Y, Z)
        for p in params:
            p.grad = None
        loss.backward()

    sd = {k: v.detach().clone() for k, v in enc.state_dict().items()}
    temp = crit.temp.detach().clone()
    mask = restate.dropout_mask(enc.subject_block.spatial_attention.spatial_dropout.loc, args.d_drop, 3).to(DEV)
    idl = ids.tolist()

    def eager(autocast):
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            restate.train_step(sd, X, Y, idl, temp, mask)

    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        t_tf32 = _time(lambda: eager(False), 1, 3)
        t_bf16 = _time(lambda: eager(True), 1, 3)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    t_native = _time(native, 3, 10)
    out = {"config": "cfg2 B=256, 208x360, F=1024, fwd + CLIP + bwd, one B200",
           "native_bf16_ms": round(t_native, 3), "eager_tf32_ms": round(t_tf32, 3), "eager_autocast_bf16_ms": round(t_bf16, 3),
           "speedup_vs_eager_tf32": round(t_tf32 / t_native, 2), "speedup_vs_eager_autocast_bf16": round(t_bf16 / t_native, 2)}
    print(json.dumps(out))
    d = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, "eager_vs_native.json"), "w") as f:
            json.dump(out, f, indent=1)
    assert t_native < 0.7 * min(t_tf32, t_bf16), out
