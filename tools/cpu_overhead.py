"""How long does the host take to enqueue one training step (Python + ctypes + allocator), vs GPU time?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "speech-decoding_b200")]
import torch, numpy as np
import bench
import sd_b200
from speech_decoding.models import BrainEncoder
from speech_decoding.utils.loss import CLIPLoss
dev = torch.device("cuda:0")
args = bench.make_args_ns()
enc = BrainEncoder(args).to(dev).train(); crit = CLIPLoss(args).to(dev).train()
X, Y, ids = bench.synth(256, 1); X, Y = X.to(dev), Y.to(dev)
def step():
    Z = enc(X, ids); loss = crit(Y, Z)
    for p in enc.parameters(): p.grad = None
    crit.temp.grad = None
    loss.backward()
for _ in range(3): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10): step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host enqueue %.2f ms/step ; total %.2f ms/step" % ((t1 - t0) * 100, (t2 - t0) * 100))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(5): step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
