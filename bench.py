#!/usr/bin/env python
"""Benchmark of the BrainEncoder + CLIPLoss training hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision bf16|fp32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE.json configs[1]/[2] -- synthetic Gwilliams2022-shape MEG
(208 sensors, 360 samples, 27 subjects, D1=270, D2=320, F=1024), B=256 per GPU, random-init weights.
A step is  Z = encoder(X, ids); loss = clip(Y, Z); loss.backward()  (+ gradient all-reduce for N>1).

Prints ONE JSON line (rank 0).  See the repo-level prompt/DESIGN.md for the key meanings:
value = samples/s with inputs resident in HBM; e2e = same through the public API from pinned host
buffers incl. H2D copies, Adam step and the loss read-back; roofline = the tcgen05 conv kernel
(forward + data-gradient launches) against the measured bf16 peak; cpu_baseline / --impl reference =
the reference's own unmodified modules (pip-installed under baseline/_ref, see oracle/install_reference.py;
the oracle port oracle/restate.py only if that install is absent) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "speech-decoding_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np   # noqa: E402
import torch         # noqa: E402

CFG = dict(C=208, T=360, S=27, D1=270, D2=320, F=1024, K=32, B=256)
# forward FLOPs per sample (2*MAC), SURVEY.md §8(d) / BASELINE.md §3
ENC_FWD_GF, ENC_BWD_GF = 5.154, 10.267


def step_gflop_per_sample(global_b):
    scale = CFG["T"] / 360.0                      # every encoder FLOP is per time step (T = 360 in SURVEY 8d)
    clip = 2.0 * global_b * CFG["F"] * CFG["T"] / 1e9
    return (ENC_FWD_GF + ENC_BWD_GF) * scale + 2 * clip


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        d = json.load(open(path))
        return dict(tflops_burst=d.get("bf16_tflops", 1590.0), tflops=d.get("bf16_tflops_sustained", 1400.0),
                    hbm_gbs=d.get("hbm_gbs", 6650.0), source="measured")
    return dict(tflops_burst=1590.0, tflops=1400.0, hbm_gbs=6650.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / power / throttle reasons, sampled every ~10 ms by a background process that is started
    before the warm-up; only the samples whose timestamp falls inside the timed region are reported."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.t0 = self.t1 = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "10"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def mark_start(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        import datetime
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if self.p is None:
            return out
        time.sleep(0.05)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, pw = [], []
        reasons = set()
        for r in rows:
            try:
                ts = datetime.datetime.strptime(r[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                if self.t0 is not None and not (self.t0 <= ts <= self.t1):
                    continue
                sm.append(float(r[1])); out["sm_max_mhz"] = float(r[2]); pw.append(float(r[3]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["sm_min_mhz"] = float(min(sm))
            out["power_w"] = float(np.median(pw))
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        return out


def conv_traffic():
    """DRAM bytes per conv_fwd_tc launch (dram__bytes_read.sum + dram__bytes_write.sum averaged over the launches of one
    step) from the committed ncu capture of this same command; None if the capture is not there."""
    try:
        with open(os.path.join(ROOT, "profiles", "r1_conv_fwd_traffic.json")) as f:
            return round(float(json.load(f)["dram_bytes_per_launch"]))
    except Exception:
        return None


class _Args(dict):
    """the attribute-and-item config object the constructors read (models.py:22-43,93-95,173-178; loss.py:32,36)"""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k) from None


def make_args(D1, D2, F_, K, num_subjects, num_channels, last4layers, layout_seed=0):
    return _Args(D1=D1, D2=D2, F=F_, K=K, d_drop=0.1, num_subjects=num_subjects, dataset="Gwilliams2022", root_dir="/nonexistent",
                 num_channels=num_channels, preprocs={"last4layers": last4layers}, reduction="mean", init_temperature=5.1,
                 layout_seed=layout_seed)          # layout_seed: explicit opt-in to the synthetic sensor layout


def make_args_ns():
    return make_args(D1=CFG["D1"], D2=CFG["D2"], F_=CFG["F"], K=CFG["K"], num_subjects=CFG["S"],
                     num_channels=CFG["C"], last4layers=True)


def synth(B, seed, device="cpu", pin=False, T=None):
    T = T or CFG["T"]
    g = torch.Generator().manual_seed(seed)
    X = torch.randn(B, CFG["C"], T, generator=g).clamp_(-20, 20)
    Y = torch.randn(B, CFG["F"], T, generator=g)
    ids = torch.randint(0, CFG["S"], (B,), generator=g, dtype=torch.int32)
    if pin:
        X, Y = X.pin_memory(), Y.pin_memory()
    return X, Y, ids


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline / stock-eager baseline: the reference's OWN modules (unmodified, pip-installed into
# baseline/_ref by oracle/install_reference.py, loaded through oracle/ref_import.py); the oracle port
# (oracle/restate.py) only when neither /root/reference nor baseline/_ref exists.  Never on the product path.
# ------------------------------------------------------------------------------------------------
class ReferenceStep:
    """One training step  Z = enc(X, ids); loss = crit(Y, Z); loss.backward()  (train.py:189-202) of the reference on
    `device`, fp32, random-init weights of the cfg2 architecture."""

    def __init__(self, device, T=None):
        from oracle import ref_import, restate
        self.device = torch.device(device)
        self.args = make_args_ns()
        self.T = T or CFG["T"]
        torch.manual_seed(0)
        np.random.seed(0)
        if ref_import.available():
            M, L = ref_import.load(lambda a: restate.synthetic_layout(a.num_channels, getattr(a, "layout_seed", 0)))
            self.enc = M.BrainEncoder(self.args).to(self.device).train()
            self.crit = L.CLIPLoss(self.args).to(self.device).train()
            self.kind = "reference"
            self.source = ("unmodified reference modules (speech_decoding/models.py, utils/loss.py) from %s"
                           % ("baseline/_ref (pip-installed copy)" if ref_import.where() == "_ref" else "/root/reference"))
            self.params = list(self.enc.parameters()) + list(self.crit.parameters())
        else:
            self.kind = "port"
            self.source = "oracle/restate.py (CPU restatement; the reference install under baseline/_ref is absent)"
            self.sd, loc = restate.init_state_dict(self.args, CFG["C"])
            self.sd = {k: v.to(self.device) for k, v in self.sd.items()}
            self.temp = torch.tensor([5.1], device=self.device)
            self.mask = restate.dropout_mask(loc, self.args.d_drop, 7).to(self.device)
            self.restate = restate

    def data(self, B, seed=123):
        X, Y, ids = synth(B, seed, T=self.T)
        return X.to(self.device), Y.to(self.device), ids

    def __call__(self, X, Y, ids):
        if self.kind == "reference":
            for p in self.params:
                p.grad = None
            Z = self.enc(X, ids)
            loss = self.crit(Y, Z)
            loss.backward()
            return float(loss.detach())
        return float(self.restate.train_step(self.sd, X, Y, ids.tolist(), self.temp, self.mask)["loss"])


def cpu_reference_rate(steps, warmup, sample_b, budget_s=None):
    """Times `warmup` + `steps` reference steps at batch `sample_b` on all host cores.  With `budget_s`, the first step is
    used to project the run time and the batch is cut to 64 when the whole run would not fit (said in `sample`)."""
    torch.set_num_threads(os.cpu_count() or 1)
    ref = ReferenceStep("cpu")
    X, Y, ids = ref.data(sample_b)
    times = []
    note = ""
    i = 0
    while i < warmup + steps:
        t0 = time.perf_counter()
        ref(X, Y, ids)
        dt = time.perf_counter() - t0
        if i == 0 and budget_s is not None and sample_b > 64 and dt * (warmup + steps) > budget_s:
            note = " (B=%d would take %.0f s for the requested steps: sample cut to B=64)" % (sample_b, dt * (warmup + steps))
            sample_b = 64
            X, Y, ids = ref.data(sample_b)
            continue
        if i >= warmup:
            times.append(dt)
        i += 1
    ms = 1e3 * float(np.mean(times))
    return dict(value=sample_b / (ms / 1e3), ms_per_step=ms, cores=torch.get_num_threads(), kind=ref.kind, batch=sample_b,
                steps=len(times), warmup=warmup,
                sample="%s; fwd + CLIP loss + backward, fp32, B=%d of the cfg2 shapes, %d warm-up + %d timed steps on %d "
                       "threads%s" % (ref.source, sample_b, warmup, len(times), torch.get_num_threads(), note))


def gpu_eager_rate(dev, B, steps=3, warmup=1):
    """The reference's own modules, stock PyTorch eager (cuDNN / cuBLAS, TF32 allowed as PyTorch defaults for convs) on
    the same GPU: the honest same-hardware competitor.  Reported next to the CPU baseline."""
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    ref = None
    try:
        ref = ReferenceStep(dev)
        X, Y, ids = ref.data(B)
        for _ in range(warmup):
            ref(X, Y, ids)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            ref(X, Y, ids)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        return {"value": round(B / (ms / 1e3), 1), "unit": "samples/s", "ms_per_step": round(ms, 2), "kind": ref.kind,
                "what": "%s, stock eager on this GPU, fp32 storage with TF32 tensor cores, B=%d" % (ref.source, B)}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
        del ref
        torch.cuda.empty_cache()


def run_reference(a, rank, world):
    if rank != 0:
        return
    r = cpu_reference_rate(max(1, a.steps), max(0, a.warmup), a.batch, budget_s=240.0)
    line = {"impl": "reference", "metric": "BrainEncoder+CLIP train samples/sec", "value": round(r["value"], 2),
            "unit": "samples/s", "n_gpus": a.gpus, "steps": r["steps"], "warmup": r["warmup"], "ms_per_step": round(r["ms_per_step"], 2),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": "cfg2 Gwilliams2022-shape MEG: B=%d, 208 sensors x 360 samples, 27 subjects, D1=270 D2=320 F=1024 "
                                   "K=32; step = encoder fwd + CLIP loss + backward; reference CPU path (fp32, the only "
                                   "precision the reference has) on the host cores" % r["batch"],
                       "global_batch": r["batch"], "same_config_as_repo_arm": bool(r["batch"] == a.batch),
                       "precision_note": "the repo arm computes in bf16 (fp32 master weights / accumulation); the reference has no bf16 path"},
            "cpu_baseline": {"value": round(r["value"], 2), "unit": "samples/s", "cores": r["cores"], "kind": r["kind"],
                             "sample": r["sample"]},
            "e2e": {"value": round(r["value"], 2), "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(index):
    """Pin this process (and therefore its pinned-memory allocations, first touch) to the CPUs NVML reports as local to
    GPU `index`: with 8 ranks each shipping its inputs over PCIe every step, host buffers on the wrong socket make the
    aggregate H2D rate the bottleneck of the e2e number.  Best effort; returns the number of CPUs bound to or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1]
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def dp_self_check(rank, world, dev):
    """N > 1, before the timed run: the batch-sharded step (SyncBN on, fp32 mode, peer-memory exchanges) against the SAME
    global batch run single-process on this rank's GPU through the same kernels -- loss and every parameter gradient.
    (Single-GPU parity with the reference is what tests/ establish; this shows on the benchmark box that sharding does
    not change the result.)  Small shapes, ~2 s.  Returns a dict for config.data_parallel.self_check."""
    import torch.distributed as dist
    import sd_b200
    from sd_b200.dist import DataParallel
    from speech_decoding.models import BrainEncoder
    from speech_decoding.utils.loss import CLIPLoss
    prev = sd_b200.get_precision()
    rng_state = np.random.get_state()
    try:
        sd_b200.set_precision("fp32")
        np.random.seed(4242)             # the same spatial-dropout centre on every rank (models.py:81 draws from numpy's RNG)
        args = make_args(D1=40, D2=48, F_=64, K=4, num_subjects=6, num_channels=20, last4layers=False)
        B, C, T = 12, 20, 96
        g = torch.Generator().manual_seed(4242)
        Xg = torch.randn(world * B, C, T, generator=g).clamp(-20, 20).to(dev)
        Yg = torch.randn(world * B, 64, T, generator=g).to(dev)
        idg = torch.randint(0, 5, (world * B,), generator=g, dtype=torch.int32)
        nets = []
        for _ in range(2):
            torch.manual_seed(7)
            nets.append((BrainEncoder(args).to(dev).train(), CLIPLoss(args).to(dev).train()))
        (enc_dp, crit_dp), (enc_1, crit_1) = nets
        dpc = DataParallel(enc_dp, crit_dp, sync_bn=True)
        st = np.random.get_state()
        sl = slice(rank * B, (rank + 1) * B)
        loss_dp = crit_dp(Yg[sl], enc_dp(Xg[sl], idg[sl]))
        loss_dp.backward()
        np.random.set_state(st)                      # same spatial-dropout centre
        loss_1 = crit_1(Yg, enc_1(Xg, idg))
        loss_1.backward()
        np.random.set_state(st)
        worst, name, problem = 0.0, "", None
        for (k, p), (_, q) in zip(enc_dp.named_parameters(), enc_1.named_parameters()):
            if q.grad is None or p.grad is None:
                if (q.grad is None) != (p.grad is None):      # (no early return: every rank must reach the collectives below)
                    problem = "gradient presence differs for " + k
                continue
            a_, b_ = (torch.view_as_real(t) if t.is_complex() else t for t in (p.grad, q.grad))
            e = float((a_ - b_).abs().max() / b_.abs().max().clamp_min(1e-30))
            if e > worst:
                worst, name = e, k
        el = abs(float(loss_dp.detach()) - float(loss_1.detach())) / abs(float(loss_1.detach()))
        out = {"ok": bool(el < 1e-4 and worst < 1e-4 and problem is None), "loss_rel_err": el, "worst_grad_rel_err": worst,
               "worst_param": name, **({"problem": problem} if problem else {}),
               "what": "fp32 mode, SyncBN, %d ranks x B=%d vs the same global batch on one GPU" % (world, B)}
        t = torch.tensor([0.0 if out["ok"] else 1.0], device=dev)
        dist.all_reduce(t)
        out["ok_all_ranks"] = bool(float(t[0]) == 0.0)
        dpc.close()
        return out
    except Exception as e:                      # never take the benchmark down
        return {"ok": False, "error": "%s: %s" % (type(e).__name__, str(e)[:200])}
    finally:
        sd_b200.set_precision(prev)
        np.random.set_state(rng_state)
        torch.cuda.empty_cache()


def run_ours(a, rank, world, local_rank):
    import torch.distributed as dist
    import sd_b200
    from sd_b200 import _native as nat, ops
    from speech_decoding.models import BrainEncoder
    from speech_decoding.utils.loss import CLIPLoss

    os.environ.setdefault("SD_B200_STRICT_TC", "1")     # a bf16 launch that cannot run on the tcgen05 kernels is an error here
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    numa_cpus = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    self_check = dp_self_check(rank, world, dev) if world > 1 and a.dp_check else None
    sd_b200.set_precision(a.precision)
    torch.manual_seed(0)           # identical replicas on every rank
    np.random.seed(0)              # identical dropout centres on every rank
    args = make_args_ns()
    enc = BrainEncoder(args).to(dev).train()
    crit = CLIPLoss(args).to(dev).train()
    from sd_b200.optim import FusedAdam
    opt = FusedAdam(list(enc.parameters()) + list(crit.parameters()), lr=3e-4)     # train.py:161-163, one launch per step
    dp = None
    if world > 1:
        from sd_b200.dist import DataParallel
        dp = DataParallel(enc, crit, sync_bn=bool(a.sync_bn))
    B = a.batch
    Xh, Yh, ids = synth(B, 1000 + rank, pin=False)
    # speech embeddings: a frozen wav2vec2 output, i.e. dataset content.  In the bf16 mode they are held (host and device)
    # in bf16 -- rounded ONCE when the dataset is built; the bf16 CLIP GEMMs round them anyway -- unless --speech-dtype fp32
    y_bf16 = a.precision == "bf16" and a.speech_dtype == "bf16"
    if y_bf16:
        Yh = Yh.to(torch.bfloat16)
    # sensor windows: the bf16 mode rounds X to bf16 in its first kernel, so shipping them in bf16 gives bit-identical
    # results for half the bytes (tests/test_gpu_parity.py::test_bf16_sensor_input_is_bit_identical)
    x_bf16 = a.precision == "bf16" and a.sensor_dtype == "bf16"
    if x_bf16:
        Xh = Xh.to(torch.bfloat16)
    Xh, Yh = Xh.pin_memory(), Yh.pin_memory()
    X, Y = Xh.to(dev), Yh.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    steps_seen = [0]

    # Data parallel: the speech rows of the NEXT step are exchanged while this step's backward runs (they are input data:
    # a prefetching loader has them one step ahead) -- software pipelining of the one bulk exchange of the path; each
    # step still contains exactly one gather.  --dp-pipeline 0 starts it at the top of the same step instead.
    pipelined = dp is not None and bool(a.dp_pipeline)
    if pipelined:
        dp.prefetch_targets(Y)

    def hot_step():
        steps_seen[0] += 1
        if dp is not None and not pipelined:
            dp.prefetch_targets(Y)          # Y all-gather overlaps the encoder forward
        Z = enc(X, ids)
        loss = crit(Y, Z)
        if pipelined:
            # next step's rows: overlap backward (copy engines, no SM), from the point --dp-push-at names
            dp.prefetch_targets(Y, during_backward=None if a.dp_push_at < 0 else a.dp_push_at)
        for p in opt.param_groups[0]["params"]:
            p.grad = None
        loss.backward()
        return loss

    # counting launches through the C ABI
    counter = {"n": 0}
    orig_call = nat.call

    def counting_call(name, *args_):
        counter["n"] += 1
        return orig_call(name, *args_)

    clocks = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(a.warmup):
        hot_step()
    barrier()
    nat.call = counting_call
    ops.nat.call = counting_call
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if clocks:
        clocks.mark_start()
    e0.record()
    for _ in range(a.steps):
        loss = hot_step()
    e1.record()
    barrier()
    if clocks:
        clocks.mark_end()
    launches = counter["n"]
    nat.call = orig_call
    ops.nat.call = orig_call
    ms = e0.elapsed_time(e1) / a.steps
    clk = clocks.stop() if clocks else None
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    value = world * B / (ms / 1e3)
    loss_val = float(loss.detach())
    del loss          # (no eager autograd graph may outlive this point: the CUDA-graph capture below needs fresh accumulators)

    # ---- roofline of the dominant kernel (tcgen05 conv fwd/dgrad), CUDA events around each launch ----
    roof = None
    if a.precision in ("bf16", "tf32", "tf32x3"):
        recs = []
        orig_conv = ops.conv_fwd

        def timed_conv(inp, w, **kw):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            r = orig_conv(inp, w, **kw)
            e.record()
            Bn, Tn, _ = inp.shape
            recs.append((s, e, 2.0 * Bn * Tn * kw["K"] * kw["N"] * kw.get("taps", 1), kw.get("taps", 1)))
            return r

        ops.conv_fwd = timed_conv
        import sd_b200.engine as eng
        eng.ops.conv_fwd = timed_conv
        nprof = min(a.steps, 5)
        for _ in range(nprof):
            hot_step()
        torch.cuda.synchronize()
        ops.conv_fwd = orig_conv
        eng.ops.conv_fwd = orig_conv
        tot_ms = sum(s.elapsed_time(e) for s, e, _, _ in recs)
        tot_fl = sum(f for _, _, f, _ in recs)
        # the k=3 instantiation alone (the top line of the ncu launch list, 95 % of these FLOPs); the 1x1 launches of the
        # same template are epilogue / HBM bound and pull the all-launch average down
        t3_ms = sum(s.elapsed_time(e) for s, e, _, tp in recs if tp == 3)
        t3_fl = sum(f for _, _, f, tp in recs if tp == 3)
        n3 = sum(1 for r in recs if r[3] == 3)
        peaks = measured_peaks()
        ach = tot_fl / (tot_ms / 1e3) / 1e12
        # TF32 dense peak = half the bf16 figure (datasheet ratio 1.125 / 2.25 PF; MEASURED_PEAKS.json holds bf16 only);
        # 3xTF32 issues three TF32 MMAs per algorithmic product
        div = {"bf16": 1.0, "tf32": 2.0, "tf32x3": 6.0}[a.precision]
        kname = "conv_fwd_tc_kernel" if a.precision == "bf16" else "conv_fwd_tf32_kernel"
        roof = {"kernel": "%s (implicit-GEMM conv forward + data-gradient, %d launches/step)" % (kname, len(recs) // nprof),
                "bound": "tensor", "achieved": round(ach, 1), "peak": round(peaks["tflops"] / div, 1), "unit": "TFLOP/s",
                "frac": round(ach / (peaks["tflops"] / div), 4),
                "frac_of_burst_peak": round(ach / (peaks["tflops_burst"] / div), 4),
                "peak_source": peaks["source"] + " bf16_tflops_sustained" + ("" if div == 1.0 else " / %g (%s)" % (div, a.precision)),
                "avg_launch_ms": round(tot_ms / len(recs), 4), "share_of_step": round(tot_ms / nprof / ms, 3),
                "traffic": conv_traffic() if a.precision == "bf16" else None}
        if n3 and t3_ms > 0:
            a3 = t3_fl / (t3_ms / 1e3) / 1e12
            roof["k3_launches_only"] = {"launches_per_step": n3 // nprof, "achieved": round(a3, 1), "frac": round(a3 / (peaks["tflops"] / div), 4),
                                        "frac_of_burst_peak": round(a3 / (peaks["tflops_burst"] / div), 4),
                                        "avg_launch_ms": round(t3_ms / n3, 4), "share_of_step": round(t3_ms / nprof / ms, 3)}

    # ---- end to end: pinned host inputs, H2D every step (prefetched on a copy stream), Adam, loss read-back ----
    copy_stream = torch.cuda.Stream(device=dev)
    bufs = [(torch.empty_like(X), torch.empty_like(Y)) for _ in range(2)]
    evs = [torch.cuda.Event() for _ in range(2)]

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            bufs[i][0].copy_(Xh, non_blocking=True)
            bufs[i][1].copy_(Yh, non_blocking=True)
            evs[i].record(copy_stream)

    gather_stream = torch.cuda.Stream(device=dev)
    graphed = None
    if world == 1 and a.graph:
        # the whole step (forward, CLIP loss, backward, fused Adam) as ONE CUDA graph (sd_b200.graph)
        from sd_b200.graph import GraphedTrainStep
        graphed = GraphedTrainStep(enc, crit, opt, X, Y, ids)

    def e2e_loop(n):
        prefetch(0)
        last = None
        for i in range(n):
            cur = i & 1
            torch.cuda.current_stream().wait_event(evs[cur])
            Xd, Yd = bufs[cur]
            if graphed is not None:
                steps_seen[0] += 1
                loss = graphed(Xd, Yd, ids)
                if i + 1 < n:
                    prefetch(cur ^ 1)                  # next batch: H2D on the copy stream while the graph runs
                last = loss.item()                     # D2H read of the step's result (train.py:196)
                continue
            if dp is not None and (not pipelined or i == 0):
                dp.prefetch_targets(Yd)
            Z = enc(Xd, ids)
            if i + 1 < n:
                # next batch: issued after the encoder's own small uploads (subject ids) so that those do not queue on the
                # H2D copy engine behind this 265 MB transfer
                prefetch(cur ^ 1)
            loss = crit(Yd, Z)
            if pipelined and i + 1 < n:
                # next step's speech rows: as soon as their H2D copy has landed, on a side stream (norms + peer pushes),
                # so that the main stream never waits for the transfer
                with torch.cuda.stream(gather_stream):
                    gather_stream.wait_event(evs[cur ^ 1])
                    dp.prefetch_targets(bufs[cur ^ 1][1])
            opt.zero_grad(set_to_none=True)
            loss.backward()
            steps_seen[0] += 1
            opt.step()
            last = loss.item()                 # D2H read of the step's result (train.py:196)
        return last

    e2e_loop(2)
    barrier()
    t0 = time.perf_counter()
    e2e_loop(a.steps)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / a.steps
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t[0])
    e2e_value = world * B / (e2e_ms / 1e3)

    if rank != 0:
        return
    gflop = step_gflop_per_sample(world * B)
    line = {"metric": "BrainEncoder+CLIP train samples/sec", "value": round(value, 1), "unit": "samples/s",
            "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": a.precision, "data": "synthetic",
            "config": {"workload": "%s Gwilliams2022-shape MEG: B=%d per GPU, 208 sensors x %d samples, 27 subjects, "
                                   "D1=270 D2=320 F=1024 K=32; step = encoder fwd + CLIP loss + backward%s"
                                   % ("cfg2/cfg3" if CFG["T"] == 360 else "cfg5 long-window" if CFG["T"] == 1200 else "custom-window",
                                      B, CFG["T"], " + grad all-reduce, global-batch CLIP negatives" if world > 1 else ""),
                       "speech_embeddings": "bf16 (rounded once on the host)" if y_bf16 else "fp32",
                       "sensor_windows": "bf16 (bit-identical results: the bf16 mode rounds X in its first kernel)" if x_bf16 else "fp32",
                       "global_batch": world * B,
                       "precision": a.precision + {"bf16": " activations, fp32 master weights/accumulate",
                                                   "tf32": ": fp32 storage, tcgen05 kind::tf32 convolutions / GEMMs (the reference's cuDNN default)",
                                                   "tf32x3": ": fp32 storage, 3xTF32 split convolutions on tcgen05 (fp32-class accuracy), fp32 CLIP",
                                                   "fp32": ": fp32 CUDA-core kernels"}[a.precision],
                       "l2": "working set (>2 GB of activations, >260 MB of inputs) exceeds the 126 MB L2: no flush needed between steps",
                       "sync_bn": bool(a.sync_bn) if world > 1 else None, "loss": round(loss_val, 4)},
            "model_tflops_per_s": round(value * gflop / 1e3, 1),
            "e2e": {"value": round(e2e_value, 1), "unit": "samples/s", "ms_per_step": round(e2e_ms, 3),
                    "h2d_bytes_per_step": int(Xh.numel() * Xh.element_size() + Yh.numel() * Yh.element_size()), "d2h_bytes_per_step": 4,
                    "includes": "H2D of X (%s) and Y (%s) from pinned memory (double-buffered on a copy stream), fwd, loss, backward, "
                                "fused Adam step (sd_adam_step), loss.item()%s" % ("bf16" if x_bf16 else "fp32", "bf16" if y_bf16 else "fp32",
                                "; the step is replayed as one CUDA graph (sd_b200.graph.GraphedTrainStep)" if graphed is not None else "")},
            "gpu_launches": launches, "clocks": clk}
    if dp is not None:
        red = enc.pipeline().reducer
        line["config"]["data_parallel"] = {
            "speech_row_gather": "copy-engine push into CUDA-IPC peer buffers (no SM)" if dp.peer is not None else "NCCL all-gather",
            "speech_row_gather_schedule": "next step's rows, during this step's backward" if pipelined else "this step's rows, during the encoder forward",
            "small_exchanges": "one-kernel all-gather through peer memory (sd_peer_exchange)" if dp.mailbox is not None else "NCCL all-reduce",
            "grad_allreduce_launches_per_step": round(red.launched / max(1, steps_seen[0]), 2),
            "numa_bound_cpus": numa_cpus, "self_check": self_check}
    if roof:
        line["roofline"] = roof
    if world == 1 and not a.no_eager:
        # second stated baseline: the reference's own modules, stock PyTorch eager (cuDNN / cuBLAS, TF32 on) on THIS GPU
        try:
            del enc, crit, opt, graphed
            torch.cuda.empty_cache()
            line["gpu_eager_baseline"] = gpu_eager_rate(dev, B)
        except Exception as e:
            line["gpu_eager_baseline"] = {"error": "%s: %s" % (type(e).__name__, str(e)[:160])}
    if world == 1 and not a.no_cpu:
        r = cpu_reference_rate(2, 1, 64)
        line["cpu_baseline"] = {"value": round(r["value"], 2), "unit": "samples/s", "cores": r["cores"], "kind": r["kind"],
                                "sample": r["sample"]}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "tf32", "tf32x3", "fp32"])
    ap.add_argument("--batch", type=int, default=CFG["B"])
    ap.add_argument("--sync-bn", type=int, default=0)
    ap.add_argument("--dp-check", type=int, default=1, help="N>1: sharded step vs the global batch on one GPU before the timed run")
    ap.add_argument("--dp-push-at", type=int, default=-1,
                    help="pipelined gather: start the pushes after the k-th stage from the end of backward (-1: right after the loss forward)")
    ap.add_argument("--dp-pipeline", type=int, default=1,
                    help="N>1: exchange the next step's speech rows during this step's backward (1) or this step's during its forward (0)")
    ap.add_argument("--graph", type=int, default=1, help="e2e leg at 1 GPU: replay the step as one CUDA graph")
    ap.add_argument("--speech-dtype", default="bf16", choices=["bf16", "fp32"],
                    help="storage of the (frozen) speech embeddings Y in the bf16 mode")
    ap.add_argument("--sensor-dtype", default="bf16", choices=["bf16", "fp32"],
                    help="storage of the sensor windows X in the bf16 mode (bf16 is bit-identical there: X is rounded first thing)")
    ap.add_argument("--window", type=int, default=CFG["T"],
                    help="samples per window: 360 = 3 s (cfg2/cfg3), 1200 = 10 s (BASELINE.json configs[4])")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-eager", action="store_true", help="skip the stock-PyTorch-eager-on-this-GPU baseline")
    a = ap.parse_args()
    CFG["T"] = a.window
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        run_reference(a, rank, world)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(a, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
