"""CPU, world_size 2 over gloo: the host-side data-parallel logic of sd_b200/dist.py -- sharded CLIP loss
with global-batch negatives (gather one side, exchange only row statistics), per-stage gradient
all-reduce over slices of the flat gradient pool, and host-side subject-presence agreement.  The device
kernels are stood in for by their closed forms in plain torch (test-only), so what is checked here is
the algebra of the decomposition and the torch.distributed plumbing."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import restate


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _phase1(x, z, temp):               # closed form of sd_clip_dots + sd_clip_phase1
    xn2, zn2 = (x * x).sum(1), (z * z).sum(1)
    logits = torch.exp(temp) * (x @ z.T) / (xn2.sqrt()[:, None] * zn2.sqrt()[None, :])
    m = logits.max(dim=1)[0]
    row_stat = torch.stack([m, torch.exp(logits - m[:, None]).sum(1)], dim=1)
    return logits, row_stat, torch.logsumexp(logits, dim=0), xn2, zn2


def _phase2(logits, row_lse, col_lse, xn2, zn2, temp, scale, diag0):   # closed form of sd_clip_phase2 + sd_clip_dz
    M, N = logits.shape
    eye = torch.zeros(M, N)
    eye[diag0 + torch.arange(N), torch.arange(N)] = 1.0
    G = 0.5 * scale * (torch.exp(logits - row_lse[:, None]) + torch.exp(logits - col_lse[None, :]) - 2 * eye)
    coef = torch.exp(temp) * G / (xn2.sqrt()[:, None] * zn2.sqrt()[None, :])
    cz = (G * logits).sum(0) / zn2
    d = logits[diag0 + torch.arange(N), torch.arange(N)]
    loss_part = scale * (0.5 * row_lse[diag0:diag0 + N].sum() + 0.5 * col_lse.sum() - d.sum())
    return coef, cz, torch.stack([loss_part, (G * logits).sum()])


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sd_b200 import dist as sd
    torch.manual_seed(0)
    B, D = 6, 40                       # per-rank batch
    Yg = torch.randn(world * B, D)
    Zg = torch.randn(world * B, D) + 0.5 * Yg
    temp = torch.tensor([1.3])
    group = dist.group.WORLD
    # ---- sharded CLIP loss ----
    y_loc, z_loc = Yg[rank * B:(rank + 1) * B].contiguous(), Zg[rank * B:(rank + 1) * B].contiguous()
    x_all = sd.all_gather_rows(y_loc, group)
    assert torch.equal(x_all, Yg)
    # the speech-row gather used by CLIPLoss / DataParallel.prefetch_targets: CPU tensors (and the fp32 mode) take the
    # exact fp32 path (no bf16 transport, no norms); the asynchronous form returns the works to wait on
    rows, norms = sd.gather_speech_rows(y_loc, group)
    assert norms is None and torch.equal(rows, Yg)
    rows, norms, works, _keep = sd.gather_speech_rows(y_loc, group, async_op=True)
    for w in works:
        w.wait()
    assert norms is None and torch.equal(rows, Yg)
    logits, row_stat, col_lse, xn2, zn2 = _phase1(x_all, z_loc, temp)
    row_lse = sd.global_row_lse(row_stat, group)          # (CPU / gloo: the NCCL-style merge; CUDA: the peer mailbox)
    merged = sd.merge_row_stats(row_stat, group)
    assert torch.allclose(row_lse, merged[:, 0] + torch.log(merged[:, 1]))
    coef, cz, partial = _phase2(logits, row_lse, col_lse, xn2, zn2, temp, 1.0 / (world * B), rank * B)
    partial = sd.all_reduce_sum(partial, group)
    dz_loc = coef.T @ x_all - cz[:, None] * z_loc
    # single-process oracle on the global batch
    Zr = Zg.clone().requires_grad_(True)
    tr = temp.clone().requires_grad_(True)
    loss_ref = restate.clip_loss(Yg.reshape(world * B, D, 1), Zr.reshape(world * B, D, 1), tr)
    loss_ref.backward()
    assert abs(float(partial[0]) - float(loss_ref)) < 1e-5 * abs(float(loss_ref))
    assert abs(float(partial[1]) - float(tr.grad)) < 1e-4 * max(abs(float(tr.grad)), 1e-3)
    assert torch.allclose(dz_loc, Zr.grad[rank * B:(rank + 1) * B], atol=1e-6, rtol=1e-4)
    # ---- per-stage gradient all-reduce on flat slices ----
    flat = torch.arange(20, dtype=torch.float32) * (rank + 1)
    want = torch.arange(20, dtype=torch.float32) * sum(r + 1 for r in range(world))
    red = sd.GradReducer(group, bucket_bytes=0)            # one all-reduce per stage
    red.stage_done(flat, 12, 20)
    red.stage_done(flat, 0, 12)
    red.finish()
    assert torch.equal(flat, want) and red.launched == 2
    # bucketed: adjacent stage slices (stages finish in reverse parameter order) are coalesced until the bucket is full
    flat = torch.arange(20, dtype=torch.float32) * (rank + 1)
    red = sd.GradReducer(group, bucket_bytes=10 * 4)
    red.stage_done(flat, 16, 20)      # 4 elements: bucket open
    red.stage_done(flat, 9, 16)       # 11 elements >= 10: flushed as [9, 20)
    red.stage_done(flat, 3, 9)        # 6 elements: open
    red.stage_done(flat, 0, 3)        # 9 elements: still open -> finish() flushes [0, 9)
    red.finish()
    assert torch.equal(flat, want) and red.launched == 2
    # a non-adjacent slice closes the open bucket first
    flat = torch.arange(20, dtype=torch.float32) * (rank + 1)
    red = sd.GradReducer(group, bucket_bytes=1 << 20)
    red.stage_done(flat, 15, 20)
    red.stage_done(flat, 0, 5)
    red.finish()
    assert torch.equal(flat[15:], want[15:]) and torch.equal(flat[:5], want[:5]) and red.launched == 2
    assert torch.equal(flat[5:15], torch.arange(5, 15, dtype=torch.float32) * (rank + 1))
    # small in-place all-reduce: without a peer mailbox (CPU) it is the plain collective
    t = torch.full((3,), float(rank + 1), dtype=torch.float64)
    assert torch.equal(sd.small_all_reduce_sum_(t, group), torch.full((3,), 3.0, dtype=torch.float64))
    # ---- subject presence agreed on the host ----
    ids = np.array([0, 3, 3]) if rank == 0 else np.array([5, 5, 1])
    present = np.unique(np.concatenate(sd.gather_host_ints(ids, group)))
    assert present.tolist() == [0, 1, 3, 5]
    collect = sd.gather_host_ints_async(ids, group)        # started in the forward, collected in backward
    assert np.unique(np.concatenate(collect())).tolist() == [0, 1, 3, 5]
    open(os.path.join(tmp, "ok%d" % rank), "w").write("ok")
    dist.destroy_process_group()


def test_data_parallel_host_logic_gloo_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / ("ok%d" % r)) for r in range(world))
