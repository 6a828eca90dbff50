"""Free-running per-phase CUDA-event timing of the data-parallel step (run under torchrun; no host sync inside the
measured loop).  Phases: encoder forward / loss forward / backward (+ gradient all-reduce); in the copy-engine gather
mode also the duration of the peer pushes.  Env: SD_B200_DP_GATHER=peer|nccl, SD_B200_DP_BUCKET_MB, PIPE=1 (start the
gather of the NEXT step's speech rows right after the loss forward, so that it overlaps backward + the next forward),
SYNC_BN=1.  Prints one line per run on rank 0 with the median over ranks and the slowest rank."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "speech-decoding_b200")]
import numpy as np
import torch
import torch.distributed as dist
import bench

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
import sd_b200
from speech_decoding.models import BrainEncoder
from speech_decoding.utils.loss import CLIPLoss
sd_b200.set_precision("bf16")
torch.manual_seed(0); np.random.seed(0)
args = bench.make_args_ns()
enc = BrainEncoder(args).to(dev).train()
crit = CLIPLoss(args).to(dev).train()
dp = None
if world > 1:
    from sd_b200.dist import DataParallel
    dp = DataParallel(enc, crit, sync_bn=os.environ.get("SYNC_BN", "0") == "1")
    if dp.peer is not None:
        dp.peer.timing = []
Xh, Yh, ids = bench.synth(256, 1000 + rank, pin=False)
X, Y = Xh.to(dev), Yh.to(dev)
pipe = os.environ.get("PIPE", "0") == "1"
params = list(enc.parameters()) + list(crit.parameters())
W, K = 5, 20
evs = []
if dp is not None and pipe:
    dp.prefetch_targets(Y)
for it in range(W + K):
    if it == W:
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if dp is not None and dp.peer is not None:
            dp.peer.timing = []
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    if dp is not None and not pipe:
        dp.prefetch_targets(Y)
    Z = enc(X, ids); ev[1].record()
    loss = crit(Y, Z); ev[2].record()
    if dp is not None and pipe:
        dp.prefetch_targets(Y)          # next step's rows: overlaps backward and the next encoder forward
    for p in params:
        p.grad = None
    loss.backward(); ev[3].record()
    if it >= W:
        evs.append(ev)
torch.cuda.synchronize()
ph = np.array([[e[i].elapsed_time(e[i + 1]) for i in range(3)] for e in evs]).mean(axis=0)
total = evs[0][0].elapsed_time(evs[-1][3]) / len(evs)
push = 0.0
if dp is not None and dp.peer is not None and dp.peer.timing:
    push = float(np.mean([a.elapsed_time(b) for a, b in dp.peer.timing[:K]]))
vec = torch.tensor([ph[0], ph[1], ph[2], total, push], device=dev)
if world > 1:
    allv = [torch.empty_like(vec) for _ in range(world)]
    dist.all_gather(allv, vec)
    allv = torch.stack(allv).cpu().numpy()
else:
    allv = vec.cpu().numpy()[None]
if rank == 0:
    med, mx = np.median(allv, axis=0), allv.max(axis=0)
    print("world %d gather=%s bucketMB=%s pipe=%d syncbn=%s | enc fwd %.3f (max %.3f) | loss fwd %.3f (%.3f) | backward %.3f (%.3f) | step %.3f (%.3f) | peer push %.3f (%.3f) ms"
          % (world, os.environ.get("SD_B200_DP_GATHER", "peer"), os.environ.get("SD_B200_DP_BUCKET_MB", "16"), pipe, os.environ.get("SYNC_BN", "0"),
             med[0], mx[0], med[1], mx[1], med[2], mx[2], med[3], mx[3], med[4], mx[4]), flush=True)
if world > 1:
    dist.destroy_process_group()
