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


def _peer_worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    host = dist.new_group(backend="gloo")
    import sd_b200
    from sd_b200 import dist as sd
    sd_b200.set_precision("bf16")
    # ---- one-kernel all-gather through peer memory: many epochs back to back, fp32 / fp64 payloads of varying size ----
    mb = sd.PeerMailbox(dist.group.WORLD, host, dev)
    for it in range(40):
        n = [2, 640, 4096, 1][it % 4]
        dt = torch.float64 if it % 3 == 0 else torch.float32
        t = (torch.arange(n, device=dev, dtype=dt) + 1000.0 * rank + it)
        g = mb.exchange(t)
        for q in range(world):
            assert torch.equal(g[q], torch.arange(n, device=dev, dtype=dt) + 1000.0 * q + it), (it, q)
        s = mb.all_reduce_sum_(t.clone())
        assert torch.equal(s, sum(torch.arange(n, device=dev, dtype=dt) + 1000.0 * q + it for q in range(world)))
    # ---- copy-engine gather of bf16 rows + norms with the flag fence: several steps, both parity slots ----
    pg = sd.PeerGather(dist.group.WORLD, host, dev)
    rows, D = 24, 4096
    for it in range(6):
        xb = (torch.randn(rows, D, device=dev, generator=torch.Generator(device=dev).manual_seed(100 * it + rank))).to(torch.bfloat16)
        n2 = (xb.float() ** 2).sum(1)
        allrows, allnorms, works = pg.gather(xb, n2)
        for w in works:
            w.wait()
        for q in range(world):
            ref = torch.randn(rows, D, device=dev, generator=torch.Generator(device=dev).manual_seed(100 * it + q)).to(torch.bfloat16)
            assert torch.equal(allrows[q * rows:(q + 1) * rows], ref), (it, q)
            assert torch.equal(allnorms[q * rows:(q + 1) * rows], (ref.float() ** 2).sum(1)), (it, q)
        # the step's closing collective (gradient all-reduce in training) is what makes two slots sufficient
        dist.all_reduce(torch.zeros(1, device=dev))
    torch.cuda.synchronize()
    open(os.path.join(tmp, "ok%d" % rank), "w").write("ok")
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_peer_memory_exchange_and_copy_engine_gather(tmp_path):
    """The two peer-memory primitives of the data-parallel path (sd_peer_exchange mailbox, copy-engine PeerGather with its
    flag fence) against their definitions, over many epochs."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    mp.spawn(_peer_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")
