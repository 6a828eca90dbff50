"""Micro-benchmarks of single C-ABI ops at the BASELINE.json cfg2 shapes (CUDA events, L2 flushed)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "speech-decoding_b200")]
import torch
from sd_b200 import ops, _native as nat

DEV = "cuda:0"
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=DEV)


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def pack(w, dtype):
    N, K, taps = w.shape
    Np, Kp = ops.rup8(N), ops.rup8(K)
    wf = torch.empty((1, taps, Np, Kp), dtype=dtype, device=DEV)
    wd = torch.empty((1, taps, Kp, Np), dtype=dtype, device=DEV)
    nat.call("sd_pack_weight", w.data_ptr(), wf.data_ptr(), wd.data_ptr(), N, K, taps, Np, Kp, ops.code_of(wf), ops._st())
    return wf, wd


def main():
    B, T = 256, 360
    dt = torch.bfloat16
    res = {}
    for name, K, N, taps, dil in [("conv 320->320 k3", 320, 320, 3, 4), ("conv 320->640 k3", 320, 640, 3, 2),
                                  ("conv 270->320 k3", 270, 320, 3, 1), ("1x1 640->1024", 640, 1024, 1, 1),
                                  ("1x1 320->640", 320, 640, 1, 1), ("1x1 270->270", 270, 270, 1, 1),
                                  ("mix 208->270", 208, 270, 1, 1)]:
        x = torch.randn(B, T, ops.rup8(K), device=DEV).to(dt)
        w = torch.randn(N, K, taps, device=DEV) / (K * taps) ** 0.5
        wf, wd = pack(w, dt)
        out = torch.empty((B, T, ops.rup8(N)), dtype=dt, device=DEV)
        bias = torch.randn(N, device=DEV)
        stats = torch.zeros((2, ops.rup8(N)), dtype=torch.float64, device=DEV)
        flops = 2.0 * B * T * K * N * taps
        for impl in ("tc",):
            ops.set_impl(impl)
            ms = timeit(lambda: ops.conv_fwd(x, wf, K=K, N=N, taps=taps, dil=dil, bias=bias, out=out))
            ms_s = timeit(lambda: ops.conv_fwd(x, wf, K=K, N=N, taps=taps, dil=dil, bias=bias, res=out, out=out, stats=stats))
            res[name] = dict(ms=ms, tflops=flops / ms / 1e9, ms_res_stats=ms_s, tflops_res_stats=flops / ms_s / 1e9)
            print("%-20s %s  plain %.3f ms %.0f TF/s | +res+stats %.3f ms %.0f TF/s" % (name, impl, ms, flops / ms / 1e9, ms_s, flops / ms_s / 1e9), flush=True)
        ops.set_impl("auto")
        dy = torch.randn(B, T, ops.rup8(N), device=DEV).to(dt)
        dw = torch.zeros(N, K, taps, device=DEV)
        db = torch.zeros(N, device=DEV)
        ms = timeit(lambda: ops.conv_wgrad(dy, x, dw, K=K, N=N, taps=taps, dil=dil, dbias=db), iters=3, warm=1)
        res[name]["wgrad_ms"] = ms
        print("%-20s wgrad %.3f ms %.0f TF/s" % (name, ms, flops / ms / 1e9), flush=True)
    # elementwise
    y = torch.randn(B, T, 320, device=DEV).to(dt)
    u = torch.empty_like(y)
    ss = torch.randn(4, 320, device=DEV)
    ms = timeit(lambda: ops.bn_gelu_fwd(y, ss, u))
    print("bn_gelu_fwd %.3f ms  %.0f GB/s" % (ms, 2 * y.numel() * 2 / ms / 1e6))
    ms = timeit(lambda: torch.nn.functional.gelu(y, out=u) if False else u.copy_(y)); print("torch copy_ same shape %.3f ms %.0f GB/s" % (ms, 2 * y.numel() * 2 / ms / 1e6))
    ms = timeit(lambda: torch.nn.functional.gelu(y)); print("torch gelu same shape %.3f ms %.0f GB/s" % (ms, 2 * y.numel() * 2 / ms / 1e6))
    yb = torch.randn(4 * B, T, 320, device=DEV).to(dt); ub = torch.empty_like(yb)
    def many(fn, n=10):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n): fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / n
    ms = many(lambda: ops.bn_gelu_fwd(yb, ss, ub)); print("bn_gelu_fwd 4x rows, back-to-back %.3f ms %.0f GB/s" % (ms, 2 * yb.numel() * 2 / ms / 1e6))
    ms = many(lambda: ub.copy_(yb)); print("torch copy_ 4x rows, back-to-back %.3f ms %.0f GB/s" % (ms, 2 * yb.numel() * 2 / ms / 1e6))
    du = torch.randn_like(yb); red = torch.zeros(2, 320, dtype=torch.float64, device=DEV); dg = torch.zeros(320, device=DEV); db = torch.zeros(320, device=DEV)
    ms = many(lambda: ops.bn_gelu_bwd(du, yb, ss, red, dg, db, 320, True)); print("bn_gelu_bwd (reduce+apply) 4x rows %.3f ms %.0f GB/s" % (ms, 5 * yb.numel() * 2 / ms / 1e6))
    Y = torch.randn(B, 1024 * T, device=DEV)
    Z = torch.randn(B, 1024 * T, device=DEV)
    ms = timeit(lambda: ops.rownorm2(Y)); print("rownorm2 %.3f ms %.0f GB/s" % (ms, Y.numel() * 4 / ms / 1e6))
    ms = timeit(lambda: ops.clip_dots(Y, Z), iters=3, warm=1); print("clip_dots %.3f ms" % ms)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "bench_ops.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
