"""Per-phase CUDA-event timing of one data-parallel step (run under torchrun): encoder fwd / loss fwd / backward."""
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
    dp = DataParallel(enc, crit, sync_bn=False)
Xh, Yh, ids = bench.synth(256, 1000 + rank, pin=False)
X, Y = Xh.to(dev), Yh.to(dev)
prefetch = os.environ.get("PREFETCH", "1") == "1"
names = ["enc fwd", "loss fwd", "backward"]
acc = np.zeros(3); n = 0
for it in range(13):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev[0].record()
    if dp is not None and prefetch:
        dp.prefetch_targets(Y)
    Z = enc(X, ids); ev[1].record()
    loss = crit(Y, Z); ev[2].record()
    for p in list(enc.parameters()) + list(crit.parameters()):
        p.grad = None
    loss.backward(); ev[3].record()
    torch.cuda.synchronize()
    if it >= 3:
        acc += np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(3)]); n += 1
if rank == 0:
    print("world %d prefetch %d | " % (world, prefetch) + " | ".join("%s %.3f ms" % (nm, v / n) for nm, v in zip(names, acc)) + " | total %.3f ms" % (acc.sum() / n), flush=True)
if world > 1:
    dist.destroy_process_group()
