"""GPU (needs >= 2 devices; skipped otherwise): batch-sharded BrainEncoder + CLIPLoss over NCCL equals the
single-process oracle on the concatenated global batch (SyncBN on, fp32 mode: 1e-4)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import restate
from tests import golden_util as G

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmp, precision):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import sd_b200
    from sd_b200.dist import DataParallel
    from speech_decoding.models import BrainEncoder
    from speech_decoding.utils.loss import CLIPLoss
    sd_b200.set_precision(precision)
    dev = torch.device("cuda", rank)
    torch.manual_seed(0)
    args = restate.make_args(D1=40, D2=48, F_=64, K=4, num_subjects=6, num_channels=20, last4layers=False)
    B, C, T = 12, 20, 96                                # per rank
    Xg = torch.randn(world * B, C, T).clamp(-20, 20)
    Yg = torch.randn(world * B, 64, T)
    idg = torch.randint(0, 5, (world * B,), dtype=torch.int32)       # subject 5 absent everywhere
    idg[:B] = idg[:B] % 3                                            # subjects 3,4 absent on rank 0
    enc, crit = BrainEncoder(args).to(dev).train(), CLIPLoss(args).to(dev).train()
    sd0 = {k: v.detach().cpu().clone() for k, v in enc.state_dict().items()}
    DataParallel(enc, crit, sync_bn=True)
    orig = np.random.randint
    np.random.randint = lambda *a, **k: 4                            # same dropout centre on every rank
    try:
        Z = enc(Xg[rank * B:(rank + 1) * B].to(dev), idg[rank * B:(rank + 1) * B])
    finally:
        np.random.randint = orig
    loss = crit(Yg[rank * B:(rank + 1) * B].to(dev), Z)
    loss.backward()
    torch.cuda.synchronize()
    if rank == 0:
        mask = restate.dropout_mask(enc.subject_block.spatial_attention.spatial_dropout.loc, args.d_drop, 4)
        ref = restate.train_step(sd0, Xg, Yg, idg.tolist(), crit.temp.detach().cpu(), mask)
        tol = 1e-4 if precision == "fp32" else 5e-2
        err = G.rel_err if precision == "fp32" else G.rel_l2
        measured = {"loss": G.rel_err(loss, ref["loss"]), "Z": err(Z, ref["Z"][:B]), "grad_worst": 0.0}
        assert measured["loss"] < tol
        assert measured["Z"] < tol
        named = dict(enc.named_parameters())
        for k, g in ref["grads"].items():
            if g is None:
                assert named[k].grad is None, k
                continue
            assert named[k].grad is not None, k
            scale = float(g.abs().max())
            if k.endswith("bias"):
                scale = max(scale, float(ref["grads"][k[:-4] + "weight"].abs().max()))
            e = G.rel_err(named[k].grad, g, floor=scale)
            measured["grad_worst"] = max(measured["grad_worst"], e)
            assert e < (1e-4 if precision == "fp32" else 1e-1), (k, e)
        measured["grad_temp"] = G.rel_err(crit.temp.grad, ref["dtemp"])
        assert measured["grad_temp"] < (1e-4 if precision == "fp32" else 5e-2)
        import json
        json.dump(measured, open(os.path.join(tmp, "measured.json"), "w"))
        sd1 = enc.state_dict()
        for k, v in sd0.items():          # sd0 now holds the oracle's updated running statistics
            if "running" in k:
                assert G.rel_err(sd1[k], v, floor=1e-3) < (1e-3 if precision == "fp32" else 3e-2), k
        open(os.path.join(tmp, "ok"), "w").write("ok")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_two_gpu_data_parallel_matches_global_batch_oracle(tmp_path, precision):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path), precision), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok")
    import json
    from tests import parity_log as PL
    for k, v in json.load(open(tmp_path / "measured.json")).items():
        PL.record(k, v, 1e-4 if precision == "fp32" else (1e-1 if k == "grad_worst" else 5e-2))
