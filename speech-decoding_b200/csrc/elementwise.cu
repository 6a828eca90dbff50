// HBM-bound kernels of the hot path: layout conversion, weight packing, BatchNorm statistics /
// apply / backward fused with GELU, GLU, GELU backward.  All operate on the channels-last "BTC"
// activation layout (rows = B*T, Cp channels contiguous, Cp % 8 == 0) with 8/16-byte vector access.
#include <stdlib.h>
#include "common.cuh"

namespace sd {

// ---------------------------------------------------------------------------------------------------
// (B,C,T) fp32 <-> (B,T,Cp) T : 32x32 shared-memory tile transpose, both sides coalesced
// ---------------------------------------------------------------------------------------------------
template <typename T, typename TIn = float>
__global__ void nct_to_btc_kernel(const TIn* __restrict__ x, T* __restrict__ out, int C, int Tn, int Cp) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  const TIn* xb = x + (size_t)b * C * Tn;
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    int c = c0 + ty + i, t = t0 + tx;
    tile[ty + i][tx] = (c < C && t < Tn) ? to_f<TIn>(xb[(size_t)c * Tn + t]) : 0.f;
  }
  __syncthreads();
  T* ob = out + (size_t)b * Tn * Cp;
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    int t = t0 + ty + i, c = c0 + tx;
    if (t < Tn && c < Cp) ob[(size_t)t * Cp + c] = from_f<T>(tile[tx][ty + i]);
  }
}

template <typename T>
__global__ void btc_to_nct_kernel(const T* __restrict__ in, float* __restrict__ out, int C, int Tn, int Cp) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const T* ib = in + (size_t)b * Tn * Cp;
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    int t = t0 + ty + i, c = c0 + tx;
    tile[ty + i][tx] = (t < Tn && c < Cp) ? to_f<T>(ib[(size_t)t * Cp + c]) : 0.f;
  }
  __syncthreads();
  float* ob = out + (size_t)b * C * Tn;
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    int c = c0 + ty + i, t = t0 + tx;
    if (c < C && t < Tn) ob[(size_t)c * Tn + t] = tile[tx][ty + i];
  }
}

// ---------------------------------------------------------------------------------------------------
// weight packing: (N,K,taps) fp32 -> wf (taps,Np,Kp), wd (taps,Kp,Np) taps reversed
// ---------------------------------------------------------------------------------------------------
// One 32(n) x 32(k) tile of one weight, all taps: the source rows w[n, k0:k0+32, :] are contiguous (32*taps
// floats), both destinations are written along their contiguous dimension through a shared-memory tile.
template <typename T>
__device__ __forceinline__ void pack_tile(const sd_pack_entry& e, int tile, float (*t)[32][33]) {
  const int tiles_k = (e.Kp + 31) / 32, tiles_n = (e.Np + 31) / 32;
  if (tile >= tiles_k * tiles_n) return;
  const int n0 = (tile / tiles_k) * 32, k0 = (tile % tiles_k) * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  for (int nn = ty; nn < 32; nn += 8) {
    const int n = n0 + nn;
    for (int idx = tx; idx < 32 * e.taps; idx += 32) {
      const int kk = idx / e.taps, j = idx - kk * e.taps;
      float v = 0.f;
      if (n < e.N && k0 + kk < e.K) v = e.w[((int64_t)n * e.K + k0) * e.taps + idx];
      t[j][nn][kk] = v;
    }
  }
  __syncthreads();
  T* wf = reinterpret_cast<T*>(e.wf);
  T* wd = reinterpret_cast<T*>(e.wd);
  for (int j = 0; j < e.taps; ++j) {
    if (wf) {
      for (int nn = ty; nn < 32; nn += 8) {
        const int n = n0 + nn, k = k0 + tx;
        if (n < e.Np && k < e.Kp) wf[((int64_t)j * e.Np + n) * e.Kp + k] = from_f<T>(t[j][nn][tx]);
      }
    }
    if (wd) {
      for (int kk = ty; kk < 32; kk += 8) {
        const int k = k0 + kk, n = n0 + tx;
        if (n < e.Np && k < e.Kp) wd[((int64_t)(e.taps - 1 - j) * e.Kp + k) * e.Np + n] = from_f<T>(t[j][tx][kk]);
      }
    }
  }
}

__global__ void __launch_bounds__(256) pack_weights_kernel(const sd_pack_entry* __restrict__ table) {
  __shared__ float t[3][32][33];
  const sd_pack_entry e = table[blockIdx.y];
  if (e.dtype == SD_BF16) pack_tile<__nv_bfloat16>(e, blockIdx.x, t);
  else pack_tile<float>(e, blockIdx.x, t);
}

__global__ void __launch_bounds__(256) pack_weight_single_kernel(sd_pack_entry e) {
  __shared__ float t[3][32][33];
  if (e.dtype == SD_BF16) pack_tile<__nv_bfloat16>(e, blockIdx.x, t);
  else pack_tile<float>(e, blockIdx.x, t);
}

// ---------------------------------------------------------------------------------------------------
// 8-channel vectors (16 B of bf16 / 32 B of fp32).  Channel-parameterised kernels use a 2-D block:
// threadIdx.x = channel vector (fixed for the thread's lifetime, so per-channel parameters live in
// registers), threadIdx.y = row lane; a block walks a contiguous slab of rows, 2 rows in flight per
// thread.  Memory is touched in fully contiguous runs of Cp elements per row.
// ---------------------------------------------------------------------------------------------------
struct F8 { float v[8]; };

template <typename T> __device__ __forceinline__ F8 ld8(const T* p);
template <> __device__ __forceinline__ F8 ld8<float>(const float* p) {
  F8 r;
  float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
template <> __device__ __forceinline__ F8 ld8<__nv_bfloat16>(const __nv_bfloat16* p) {
  F8 r;
  uint4 q = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
    r.v[2 * i] = f.x; r.v[2 * i + 1] = f.y;
  }
  return r;
}
template <typename T> __device__ __forceinline__ void st8(T* p, const F8& r);
template <> __device__ __forceinline__ void st8<float>(float* p, const F8& r) {
  *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(r.v[4], r.v[5], r.v[6], r.v[7]);
}
template <> __device__ __forceinline__ void st8<__nv_bfloat16>(__nv_bfloat16* p, const F8& r) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 h = __floats2bfloat162_rn(r.v[2 * i], r.v[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ F8 ldp8(const float* p) { return ld8<float>(p); }

// 8 channels as loaded (bf16: one 16-byte register quad): rows waiting in flight are held packed, so that more
// loads fit in the register budget of the HBM-bound channel kernels
template <typename T> struct Raw8;
template <> struct Raw8<float> { float4 a, b; };
template <> struct Raw8<__nv_bfloat16> { uint4 w; };
__device__ __forceinline__ Raw8<float> ldraw8(const float* p) {
  Raw8<float> r;
  r.a = *reinterpret_cast<const float4*>(p);
  r.b = *reinterpret_cast<const float4*>(p + 4);
  return r;
}
__device__ __forceinline__ Raw8<__nv_bfloat16> ldraw8(const __nv_bfloat16* p) {
  Raw8<__nv_bfloat16> r;
  r.w = *reinterpret_cast<const uint4*>(p);
  return r;
}
__device__ __forceinline__ F8 unpack8(const Raw8<float>& r) {
  F8 f;
  f.v[0] = r.a.x; f.v[1] = r.a.y; f.v[2] = r.a.z; f.v[3] = r.a.w;
  f.v[4] = r.b.x; f.v[5] = r.b.y; f.v[6] = r.b.z; f.v[7] = r.b.w;
  return f;
}
__device__ __forceinline__ F8 unpack8(const Raw8<__nv_bfloat16>& r) {
  F8 f;
  const uint32_t w[4] = {r.w.x, r.w.y, r.w.z, r.w.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
    f.v[2 * i] = t.x; f.v[2 * i + 1] = t.y;
  }
  return f;
}


// MODE 0: sum,sumsq of x ; MODE 1: bn+gelu backward reduce over g = du * gelu'(bn(y)) (g is not stored:
// the apply pass recomputes it, which is cheaper than a 59 MB write + read)
template <typename T, int MODE, bool REVERSE = false, bool STORE_G = false>
__global__ void __launch_bounds__(256, MODE == 0 ? 4 : 3)
colreduce_kernel(T* __restrict__ x, const T* __restrict__ y, const float* __restrict__ ss, double* __restrict__ out,
                 int64_t rows, int Cp) {
  extern __shared__ float red_smem[];   // [blockDim.y][Cp] x 2
  const int cv = threadIdx.x, c = cv * 8, ry = threadIdx.y, R = blockDim.y;
  constexpr int UNR = MODE == 0 ? 2 : 4;   // rows in flight per thread (held packed; register budget)
  // persistent blocks (grid = SMs x resident blocks): groups of UNR*R rows strided over the grid -- no tail wave,
  // and one shared-memory reduction + one round of atomics per block
  const int64_t r1 = rows;
  const int64_t r_stride = (int64_t)gridDim.x * UNR * R;
  F8 a, q;
#pragma unroll
  for (int i = 0; i < 8; ++i) a.v[i] = q.v[i] = 0.f;
  F8 sc, sh;
  if (MODE == 1) { sc = ldp8(ss + c); sh = ldp8(ss + Cp + c); }
  for (int64_t r = (int64_t)blockIdx.x * UNR * R + ry; r < r1; r += r_stride) {
    Raw8<T> rv[UNR], ry_[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int64_t rr = r + (int64_t)u * R;
      if (rr < r1) {
        const int64_t rm = REVERSE ? r1 - 1 - rr : rr;     // (sums do not care about the order; the L2 does)
        rv[u] = ldraw8(x + rm * Cp + c);
        if (MODE == 1) ry_[u] = ldraw8(y + rm * Cp + c);
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int64_t rr = r + (int64_t)u * R;
      if (rr >= r1) break;
      const F8 v = unpack8(rv[u]);
      if (MODE == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { a.v[i] += v.v[i]; q.v[i] += v.v[i] * v.v[i]; }
      } else {
        const F8 yy = unpack8(ry_[u]);
        F8 gs;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float g = v.v[i] * gelu_grad_t<T>(fmaf(yy.v[i], sc.v[i], sh.v[i]));
          if (STORE_G) g = to_f<T>(from_f<T>(g));   // the sums are those of g exactly as stored
          gs.v[i] = g;
          a.v[i] += g;
          q.v[i] = fmaf(g, yy.v[i], q.v[i]);        // sum g*y; sum g*xhat = invstd*(sum g*y - mean*sum g) is formed in fp64 later
        }
        if (STORE_G) {   // g replaces du in place: the apply pass then needs no second GELU-derivative evaluation
          const int64_t rm = REVERSE ? r1 - 1 - rr : rr;
          st8<T>(x + rm * Cp + c, gs);
        }
      }
    }
  }
  float* sa = red_smem;
  float* sq = red_smem + R * Cp;
#pragma unroll
  for (int i = 0; i < 8; ++i) { sa[ry * Cp + c + i] = a.v[i]; sq[ry * Cp + c + i] = q.v[i]; }
  __syncthreads();
  const int tid = ry * blockDim.x + cv, nthreads = blockDim.x * R;
  for (int ch = tid; ch < Cp; ch += nthreads) {
    float ta = 0.f, tq = 0.f;
    for (int k = 0; k < R; ++k) { ta += sa[k * Cp + ch]; tq += sq[k * Cp + ch]; }
    atomicAdd(out + ch, (double)ta);
    atomicAdd(out + Cp + ch, (double)tq);
  }
}

// dZ (B,N,T) fp32 + p (B,T,Np) -> dp (B,T,Np) = dZ^T * gelu'(p).  64(t) x 64(c) tiles: dZ is read as float4 along
// t (256 B per 16 lanes), p / dp are accessed as 8-channel vectors (128 B per 8 lanes); Np % 8 == 0.
template <typename T>
__global__ void __launch_bounds__(256)
gelu_bwd_nct_kernel(const float* __restrict__ dz, const T* __restrict__ p, T* __restrict__ dp,
                    int N, int Tn, int Np) {
  __shared__ float tile[64][65];   // [c][t]
  const int b = blockIdx.z, t0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
  const int tid = threadIdx.x;
  const float* zb = dz + (size_t)b * N * Tn;
  const bool vec_t = (Tn & 3) == 0;
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int idx = tid + it * 256;           // 64 rows(c) x 16 float4(t)
    const int cl = idx >> 4, tq = (idx & 15) * 4;
    const int c = c0 + cl, t = t0 + tq;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < N) {
      const float* src = zb + (size_t)c * Tn + t;
      if (vec_t && t + 3 < Tn) v = *reinterpret_cast<const float4*>(src);
      else {
        if (t < Tn) v.x = src[0];
        if (t + 1 < Tn) v.y = src[1];
        if (t + 2 < Tn) v.z = src[2];
        if (t + 3 < Tn) v.w = src[3];
      }
    }
    tile[cl][tq] = v.x; tile[cl][tq + 1] = v.y; tile[cl][tq + 2] = v.z; tile[cl][tq + 3] = v.w;
  }
  __syncthreads();
  const size_t base = (size_t)b * Tn * Np;
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int idx = tid + it * 256;           // 64 rows(t) x 8 channel-vectors
    const int tl = idx >> 3, cv = (idx & 7) * 8;
    const int t = t0 + tl, c = c0 + cv;
    if (t < Tn && c < Np) {
      const size_t o = base + (size_t)t * Np + c;
      F8 pv = ld8<T>(p + o), g;
#pragma unroll
      for (int i = 0; i < 8; ++i) g.v[i] = (c + i < N) ? tile[cv + i][tl] * gelu_grad_t<T>(pv.v[i]) : 0.f;
      st8<T>(dp + o, g);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// BatchNorm finalize (nn.BatchNorm1d semantics)
// ---------------------------------------------------------------------------------------------------
__global__ void bn_finalize_kernel(const double* __restrict__ stats, int C, int Cp, int64_t n,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ rmean, float* __restrict__ rvar, int64_t* __restrict__ nbt,
                                   float momentum, float eps, int training, float* __restrict__ ss) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && training && nbt) *nbt += 1;
  if (c >= Cp) return;
  float scale = 0.f, shift = 0.f, mean = 0.f, invstd = 0.f;
  if (c < C) {
    if (training) {
      double m = stats[c] / (double)n;
      double var = stats[Cp + c] / (double)n - m * m;
      if (var < 0) var = 0;
      mean = (float)m;
      invstd = (float)(1.0 / sqrt(var + (double)eps));
      if (rmean) rmean[c] = (1.f - momentum) * rmean[c] + momentum * mean;
      if (rvar) {
        double unb = n > 1 ? var * (double)n / (double)(n - 1) : var;
        rvar[c] = (1.f - momentum) * rvar[c] + momentum * (float)unb;
      }
    } else {
      mean = rmean[c];
      invstd = rsqrtf(rvar[c] + eps);
    }
    scale = gamma[c] * invstd;
    shift = beta[c] - mean * scale;
  }
  ss[c] = scale; ss[Cp + c] = shift; ss[2 * Cp + c] = mean; ss[3 * Cp + c] = invstd;
}

template <typename T>
__global__ void __launch_bounds__(256, 4)
bn_gelu_fwd_kernel(const T* __restrict__ y, const float* __restrict__ ss, T* __restrict__ u_out, int64_t rows, int Cp) {
  constexpr int UNR = 2;
  const int c = threadIdx.x * 8, ry = threadIdx.y, R = blockDim.y;
  const int64_t r1 = rows, r_stride = (int64_t)gridDim.x * UNR * R;   // persistent blocks, see colreduce_kernel
  const F8 sc = ldp8(ss + c), sh = ldp8(ss + Cp + c);
  for (int64_t r = (int64_t)blockIdx.x * UNR * R + ry; r < r1; r += r_stride) {
    F8 v[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int64_t rr = r + (int64_t)u * R;
      if (rr < r1) v[u] = ld8<T>(y + rr * Cp + c);
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int64_t rr = r + (int64_t)u * R;
      if (rr >= r1) break;
#pragma unroll
      for (int i = 0; i < 8; ++i) v[u].v[i] = gelu_t<T>(fmaf(v[u].v[i], sc.v[i], sh.v[i]));
      st8<T>(u_out + rr * Cp + c, v[u]);
    }
  }
}

// dy = scale * (g - sum_g/n - xhat * sum_gx/n)
template <typename T, bool REVERSE, bool G_READY = false>
__global__ void __launch_bounds__(256, G_READY ? 4 : 3)
bn_bwd_apply_kernel(T* __restrict__ g, const T* __restrict__ y, const float* __restrict__ ss,
                    const double* __restrict__ red, float* __restrict__ dgamma, float* __restrict__ dbeta,
                    int64_t rows, int64_t n_stat, float dscale, int C, int Cp, int training) {
  constexpr int UNR = 2;   // rows in flight per thread, held packed
  const int c = threadIdx.x * 8, ry = threadIdx.y, R = blockDim.y;
  const int64_t r1 = rows, r_stride = (int64_t)gridDim.x * UNR * R;   // persistent blocks, see colreduce_kernel
  if (blockIdx.x == 0) {
    const int tid = ry * blockDim.x + threadIdx.x;
    for (int ch = tid; ch < C; ch += blockDim.x * R) {
      if (dbeta) dbeta[ch] = (float)red[ch] * dscale;
      if (dgamma) dgamma[ch] = (float)(((double)ss[3 * Cp + ch]) * (red[Cp + ch] - (double)ss[2 * Cp + ch] * red[ch])) * dscale;
    }
  }
  const F8 sc = ldp8(ss + c), sh = ldp8(ss + Cp + c), mu = ldp8(ss + 2 * Cp + c), is = ldp8(ss + 3 * Cp + c);
  const float invn = 1.0f / (float)n_stat;
  // g = du * gelu'(sc*y + sh);  dy = sc*g + k2*y + k3 with per-channel constants
  F8 k2, k3;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float sg = (float)red[c + i] * invn;
    const float sx = (float)((double)is.v[i] * (red[Cp + c + i] - (double)mu.v[i] * red[c + i])) * invn;   // sum g*xhat / n
    k2.v[i] = training ? -sc.v[i] * is.v[i] * sx : 0.f;
    k3.v[i] = training ? sc.v[i] * (mu.v[i] * is.v[i] * sx - sg) : 0.f;
  }
  // rows are walked from the END of the tensor: the reduce pass that ran just before walked them front to back, so the
  // tail is what is still resident in the 126 MB L2 (g and y together are 118 MB at cfg2)
  for (int64_t r = (int64_t)blockIdx.x * UNR * R + ry; r < r1; r += r_stride) {
    Raw8<T> rv[UNR], ry_[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int64_t rr = r + (int64_t)u * R;
      if (rr < r1) {
        const int64_t rm = REVERSE ? r1 - 1 - rr : rr;
        rv[u] = ldraw8(g + rm * Cp + c);
        ry_[u] = ldraw8(y + rm * Cp + c);
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int64_t rr = r + (int64_t)u * R;
      if (rr >= r1) break;
      const int64_t rm = REVERSE ? r1 - 1 - rr : rr;
      F8 v = unpack8(rv[u]);
      const F8 yy = unpack8(ry_[u]);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float gg = G_READY ? v.v[i] : v.v[i] * gelu_grad_t<T>(fmaf(yy.v[i], sc.v[i], sh.v[i]));
        v.v[i] = training ? fmaf(sc.v[i], gg, fmaf(k2.v[i], yy.v[i], k3.v[i])) : sc.v[i] * gg;
      }
      st8<T>(g + rm * Cp + c, v);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// GLU over channels (F.glu(X, dim=-2), models.py:164) and GELU backward
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void glu_fwd_kernel(const T* __restrict__ y2, T* __restrict__ out, int64_t rows, int D2, int Np, int Op) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = rows * Op, step = (int64_t)gridDim.x * blockDim.x;
  for (; i < total; i += step) {
    int c = (int)(i % Op);
    int64_t r = i / Op;
    float v = 0.f;
    if (c < D2) v = to_f<T>(y2[r * Np + c]) * sigmoid_f(to_f<T>(y2[r * Np + D2 + c]));
    out[i] = from_f<T>(v);
  }
}

template <typename T>
__global__ void glu_bwd_kernel(const T* __restrict__ dout, const T* __restrict__ y2, T* __restrict__ dy2,
                               int64_t rows, int D2, int Np, int Op) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = rows * Np, step = (int64_t)gridDim.x * blockDim.x;
  for (; i < total; i += step) {
    int c = (int)(i % Np);
    int64_t r = i / Np;
    float v = 0.f;
    if (c < D2) {  // d/da = g * sigmoid(b)
      v = to_f<T>(dout[r * Op + c]) * sigmoid_f(to_f<T>(y2[r * Np + D2 + c]));
    } else if (c < 2 * D2) {  // d/db = g * a * s * (1 - s)
      float s = sigmoid_f(to_f<T>(y2[i]));
      v = to_f<T>(dout[r * Op + (c - D2)]) * to_f<T>(y2[r * Np + (c - D2)]) * s * (1.f - s);
    }
    dy2[i] = from_f<T>(v);
  }
}

// fast paths: D2 % 8 == 0 (so Op == D2, Np == 2*D2): one thread = 8 value channels + their 8 gates
template <typename T>
__global__ void __launch_bounds__(256)
glu_fwd_vec_kernel(const T* __restrict__ y2, T* __restrict__ out, int64_t rows, int D2) {
  const int nv = D2 >> 3;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = rows * nv, step = (int64_t)gridDim.x * blockDim.x;
  for (; i < total; i += step) {
    const int64_t r = i / nv;
    const int c = (int)(i % nv) * 8;
    F8 a = ld8<T>(y2 + r * 2 * D2 + c), b = ld8<T>(y2 + r * 2 * D2 + D2 + c);
#pragma unroll
    for (int k = 0; k < 8; ++k) a.v[k] *= sigmoid_t<T>(b.v[k]);
    st8<T>(out + r * D2 + c, a);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
glu_bwd_vec_kernel(const T* __restrict__ dout, const T* __restrict__ y2, T* __restrict__ dy2, int64_t rows, int D2) {
  const int nv = D2 >> 3;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = rows * nv, step = (int64_t)gridDim.x * blockDim.x;
  for (; i < total; i += step) {
    const int64_t r = i / nv;
    const int c = (int)(i % nv) * 8;
    F8 g = ld8<T>(dout + r * D2 + c), a = ld8<T>(y2 + r * 2 * D2 + c), b = ld8<T>(y2 + r * 2 * D2 + D2 + c);
    F8 da, db;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float sg = sigmoid_t<T>(b.v[k]);
      da.v[k] = g.v[k] * sg;
      db.v[k] = g.v[k] * a.v[k] * sg * (1.f - sg);
    }
    st8<T>(dy2 + r * 2 * D2 + c, da);
    st8<T>(dy2 + r * 2 * D2 + D2 + c, db);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) gelu_bwd_kernel(T* __restrict__ du, const T* __restrict__ p, int64_t n8) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (; i < n8; i += step) {
    F8 v = ld8<T>(du + i * 8), q = ld8<T>(p + i * 8);
#pragma unroll
    for (int k = 0; k < 8; ++k) v.v[k] *= gelu_grad_t<T>(q.v[k]);
    st8<T>(du + i * 8, v);
  }
}

__global__ void add_f64_to_f32_kernel(const double* __restrict__ a, float* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] += (float)a[i];
}

static inline int ew_grid(int64_t n, int threads) {
  int64_t b = (n + threads - 1) / threads;
  const int64_t cap = 148 * 32;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}
// rows handled concurrently by one block of the channel-vector kernels (blockDim = (Cp/8, rows))
// grid of the persistent channel kernels: every SM holds `per_sm` blocks, never more blocks than row groups
static inline int persistent_grid(int64_t rows, int rows_per_iter, int per_sm) {
  const int sms = device_sm_count();
  const int64_t groups = (rows + rows_per_iter - 1) / rows_per_iter;
  const int64_t g = (int64_t)sms * per_sm;
  return (int)(groups < g ? (groups > 0 ? groups : 1) : g);
}

static inline const char* bn_order() {
  static const char* o = nullptr;
  if (!o) {
    const char* e = getenv("SD_B200_BN_ORDER");
    o = (e && strlen(e) == 2) ? e : "rf";
  }
  return o;
}

static inline int chan_block_rows(int Cp) {
  int r = 256 / (Cp / 8);
  return r < 1 ? 1 : (r > 16 ? 16 : r);
}

}  // namespace sd

using namespace sd;

#define DISPATCH_DTYPE(dtype, ...)                                 \
  if ((dtype) == SD_F32) { using T = float; __VA_ARGS__; }         \
  else if ((dtype) == SD_BF16) { using T = __nv_bfloat16; __VA_ARGS__; } \
  else { set_error("unsupported dtype %d", (int)(dtype)); return 1; }

extern "C" {

int sd_nct_to_btc(const float* x, void* out, int B, int C, int T_, int Cp, int dtype, void* stream) {
  SD_REQUIRE(Cp % 8 == 0 && Cp >= C, "sd_nct_to_btc: Cp must be a multiple of 8 and >= C");
  dim3 grid(cdiv(T_, 32), cdiv(Cp, 32), B), block(32, 8);
  DISPATCH_DTYPE(dtype, nct_to_btc_kernel<T><<<grid, block, 0, (cudaStream_t)stream>>>(x, (T*)out, C, T_, Cp));
  return check_launch("nct_to_btc");
}

int sd_nct_to_btc_bf16in(const void* x, void* out, int B, int C, int T_, int Cp, int dtype, void* stream) {
  SD_REQUIRE(Cp % 8 == 0 && Cp >= C, "sd_nct_to_btc_bf16in: Cp must be a multiple of 8 and >= C");
  dim3 grid(cdiv(T_, 32), cdiv(Cp, 32), B), block(32, 8);
  DISPATCH_DTYPE(dtype, nct_to_btc_kernel<T, __nv_bfloat16><<<grid, block, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (T*)out, C, T_, Cp));
  return check_launch("nct_to_btc_bf16in");
}

int sd_btc_to_nct(const void* in, float* out, int B, int C, int T_, int Cp, int dtype, void* stream) {
  dim3 grid(cdiv(T_, 32), cdiv(Cp, 32), B), block(32, 8);
  DISPATCH_DTYPE(dtype, btc_to_nct_kernel<T><<<grid, block, 0, (cudaStream_t)stream>>>((const T*)in, out, C, T_, Cp));
  return check_launch("btc_to_nct");
}

int sd_gelu_bwd_nct(const float* dz, const void* p, void* dp, int B, int N, int T_, int Np, int dtype, void* stream) {
  SD_REQUIRE(Np % 8 == 0, "sd_gelu_bwd_nct: Np %% 8 != 0");
  dim3 grid(cdiv(T_, 64), cdiv(Np, 64), B);
  DISPATCH_DTYPE(dtype, gelu_bwd_nct_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(dz, (const T*)p, (T*)dp, N, T_, Np));
  return check_launch("gelu_bwd_nct");
}

int sd_pack_weight(const float* w, void* wf, void* wd, int N, int K, int taps, int Np, int Kp, int dtype,
                   void* stream) {
  SD_REQUIRE(taps >= 1 && taps <= 3, "sd_pack_weight: taps must be 1..3");
  sd_pack_entry e{w, wf, wd, N, K, taps, Np, Kp, dtype};
  pack_weight_single_kernel<<<cdiv(Np, 32) * cdiv(Kp, 32), 256, 0, (cudaStream_t)stream>>>(e);
  return check_launch("pack_weight");
}

int sd_pack_weights(const sd_pack_entry* table, int n, int max_tiles, void* stream) {
  if (n <= 0) return 0;
  SD_REQUIRE(max_tiles > 0, "sd_pack_weights: max_tiles = max over entries of ceil(Np/32)*ceil(Kp/32)");
  pack_weights_kernel<<<dim3(max_tiles, n), 256, 0, (cudaStream_t)stream>>>(table);
  return check_launch("pack_weights");
}

int sd_colstats(const void* x, double* stats, int64_t rows, int Cp, int dtype, void* stream) {
  SD_REQUIRE(Cp % 8 == 0, "sd_colstats: Cp %% 8 != 0");
  SD_REQUIRE(Cp <= 2048, "sd_colstats: Cp too large");
  dim3 block(Cp / 8, chan_block_rows(Cp));
  const size_t smem = (size_t)2 * block.y * Cp * sizeof(float);
  DISPATCH_DTYPE(dtype, colreduce_kernel<T, 0><<<persistent_grid(rows, 2 * block.y, 4), block, smem, (cudaStream_t)stream>>>((T*)x, nullptr, nullptr, stats, rows, Cp));
  return check_launch("colstats");
}

int sd_colsum_add(const void* x, float* out, double* scratch, int64_t rows, int C, int Cp, int dtype, void* stream) {
  SD_REQUIRE(Cp % 8 == 0 && Cp <= 2048 && C <= Cp, "sd_colsum_add: bad channel counts");
  cudaStream_t st = (cudaStream_t)stream;
  SD_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double) * 2 * Cp, st));
  dim3 block(Cp / 8, chan_block_rows(Cp));
  const size_t smem = (size_t)2 * block.y * Cp * sizeof(float);
  DISPATCH_DTYPE(dtype, colreduce_kernel<T, 0><<<persistent_grid(rows, 2 * block.y, 4), block, smem, st>>>((T*)x, nullptr, nullptr, scratch, rows, Cp));
  if (check_launch("colsum")) return 1;
  add_f64_to_f32_kernel<<<cdiv(C, 128), 128, 0, st>>>(scratch, out, C);
  return check_launch("colsum_add");
}

int sd_bn_finalize(const double* stats, int C, int Cp, int64_t n, const float* gamma, const float* beta,
                   float* running_mean, float* running_var, int64_t* num_batches_tracked, float momentum,
                   float eps, int training, float* ss, void* stream) {
  bn_finalize_kernel<<<cdiv(Cp, 128), 128, 0, (cudaStream_t)stream>>>(stats, C, Cp, n, gamma, beta, running_mean,
                                                                      running_var, num_batches_tracked, momentum,
                                                                      eps, training, ss);
  return check_launch("bn_finalize");
}

int sd_bn_gelu_fwd(const void* y, const float* ss, void* u, int64_t rows, int Cp, int dtype, void* stream) {
  SD_REQUIRE(Cp % 8 == 0 && Cp <= 2048, "sd_bn_gelu_fwd: bad Cp");
  dim3 block(Cp / 8, chan_block_rows(Cp));
  DISPATCH_DTYPE(dtype, bn_gelu_fwd_kernel<T><<<persistent_grid(rows, 2 * block.y, 4), block, 0, (cudaStream_t)stream>>>((const T*)y, ss, (T*)u, rows, Cp));
  return check_launch("bn_gelu_fwd");
}

int sd_bn_gelu_bwd_reduce(void* du_g, const void* y, const float* ss, double* red, int64_t rows, int Cp, int dtype,
                          void* stream) {
  SD_REQUIRE(Cp % 8 == 0 && Cp <= 2048, "sd_bn_gelu_bwd_reduce: bad Cp");
  dim3 block(Cp / 8, chan_block_rows(Cp));
  const size_t smem = (size_t)2 * block.y * Cp * sizeof(float);
  // Walk order of the two BatchNorm-backward passes (A/B switch SD_B200_BN_ORDER = ff | fr | rf): du was just written front
  // to back by the data-gradient conv, so its tail is what the L2 still holds
  if (bn_order()[0] == 'r') {
    DISPATCH_DTYPE(dtype, colreduce_kernel<T, 1, true, false><<<persistent_grid(rows, 4 * block.y, 3), block, smem, (cudaStream_t)stream>>>((T*)du_g, (const T*)y, ss, red, rows, Cp));
  } else {
    DISPATCH_DTYPE(dtype, colreduce_kernel<T, 1, false><<<persistent_grid(rows, 4 * block.y, 3), block, smem, (cudaStream_t)stream>>>((T*)du_g, (const T*)y, ss, red, rows, Cp));
  }
  return check_launch("bn_gelu_bwd_reduce");
}

int sd_bn_gelu_bwd_reduce_g(void* du_g, const void* y, const float* ss, double* red, int64_t rows, int Cp, int dtype,
                            void* stream) {
  SD_REQUIRE(Cp % 8 == 0 && Cp <= 2048, "sd_bn_gelu_bwd_reduce_g: bad Cp");
  dim3 block(Cp / 8, chan_block_rows(Cp));
  const size_t smem = (size_t)2 * block.y * Cp * sizeof(float);
  DISPATCH_DTYPE(dtype, colreduce_kernel<T, 1, false, true><<<persistent_grid(rows, 4 * block.y, 3), block, smem, (cudaStream_t)stream>>>((T*)du_g, (const T*)y, ss, red, rows, Cp));
  return check_launch("bn_gelu_bwd_reduce_g");
}

int sd_bn_bwd_apply(void* g_dy, const void* y, const float* ss, const double* red, float* dgamma, float* dbeta,
                    int64_t rows, int64_t n_stat, float dparam_scale, int C, int Cp, int training, int dtype, void* stream) {
  SD_REQUIRE(Cp % 8 == 0 && Cp <= 2048, "sd_bn_bwd_apply: bad Cp");
  dim3 block(Cp / 8, chan_block_rows(Cp));
  const bool fwd_order = bn_order()[1] != 'r';
  if (fwd_order) {
    DISPATCH_DTYPE(dtype, bn_bwd_apply_kernel<T, false><<<persistent_grid(rows, 2 * block.y, 3), block, 0, (cudaStream_t)stream>>>((T*)g_dy, (const T*)y, ss, red, dgamma, dbeta, rows, n_stat, dparam_scale, C, Cp, training));
  } else {
    DISPATCH_DTYPE(dtype, bn_bwd_apply_kernel<T, true><<<persistent_grid(rows, 2 * block.y, 3), block, 0, (cudaStream_t)stream>>>((T*)g_dy, (const T*)y, ss, red, dgamma, dbeta, rows, n_stat, dparam_scale, C, Cp, training));
  }
  return check_launch("bn_bwd_apply");
}

int sd_bn_bwd_apply_g(void* g_dy, const void* y, const float* ss, const double* red, float* dgamma, float* dbeta,
                      int64_t rows, int64_t n_stat, float dparam_scale, int C, int Cp, int training, int dtype, void* stream) {
  SD_REQUIRE(Cp % 8 == 0 && Cp <= 2048, "sd_bn_bwd_apply_g: bad Cp");
  dim3 block(Cp / 8, chan_block_rows(Cp));
  DISPATCH_DTYPE(dtype, bn_bwd_apply_kernel<T, false, true><<<persistent_grid(rows, 2 * block.y, 4), block, 0, (cudaStream_t)stream>>>((T*)g_dy, (const T*)y, ss, red, dgamma, dbeta, rows, n_stat, dparam_scale, C, Cp, training));
  return check_launch("bn_bwd_apply_g");
}

int sd_glu_fwd(const void* y2, void* out, int64_t rows, int D2, int Np, int Op, int dtype, void* stream) {
  if (D2 % 8 == 0 && Np == 2 * D2 && Op == D2) {
    DISPATCH_DTYPE(dtype, glu_fwd_vec_kernel<T><<<ew_grid(rows * (D2 / 8), 256), 256, 0, (cudaStream_t)stream>>>((const T*)y2, (T*)out, rows, D2));
    return check_launch("glu_fwd_vec");
  }
  DISPATCH_DTYPE(dtype, glu_fwd_kernel<T><<<ew_grid(rows * Op, 256), 256, 0, (cudaStream_t)stream>>>((const T*)y2, (T*)out, rows, D2, Np, Op));
  return check_launch("glu_fwd");
}

int sd_glu_bwd(const void* dout, const void* y2, void* dy2, int64_t rows, int D2, int Np, int Op, int dtype,
               void* stream) {
  if (D2 % 8 == 0 && Np == 2 * D2 && Op == D2) {
    DISPATCH_DTYPE(dtype, glu_bwd_vec_kernel<T><<<ew_grid(rows * (D2 / 8), 256), 256, 0, (cudaStream_t)stream>>>((const T*)dout, (const T*)y2, (T*)dy2, rows, D2));
    return check_launch("glu_bwd_vec");
  }
  DISPATCH_DTYPE(dtype, glu_bwd_kernel<T><<<ew_grid(rows * Np, 256), 256, 0, (cudaStream_t)stream>>>((const T*)dout, (const T*)y2, (T*)dy2, rows, D2, Np, Op));
  return check_launch("glu_bwd");
}

int sd_gelu_bwd(void* du_dp, const void* p, int64_t rows, int Cp, int dtype, void* stream) {
  int64_t n8 = rows * Cp / 8;
  DISPATCH_DTYPE(dtype, gelu_bwd_kernel<T><<<ew_grid(n8, 256), 256, 0, (cudaStream_t)stream>>>((T*)du_dp, (const T*)p, n8));
  return check_launch("gelu_bwd");
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------
// Multi-tensor Adam (the reference's optimizer is torch.optim.Adam over
// brain_encoder.parameters() + loss_func.parameters(), train.py:161-163,201-203): ONE launch updates every
// parameter that received a gradient.  Arithmetic in the order of torch's Adam (lerp, mul + addcmul,
// sqrt / bias_correction2_sqrt + eps, addcdiv), fp32, complex parameters as interleaved real pairs.
// ---------------------------------------------------------------------------------------------------
namespace sd {
__global__ void __launch_bounds__(256) adam_multi_kernel(const sd_adam_entry* __restrict__ table, float beta1, float beta2,
                                                         float eps, float weight_decay) {
  const sd_adam_entry e = table[blockIdx.y];
  const float w1 = 1.0f - beta1, w2 = 1.0f - beta2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < e.n; i += (int64_t)gridDim.x * blockDim.x) {
    float g = e.grad[i];
    const float p = e.param[i];
    if (weight_decay != 0.f) g = fmaf(weight_decay, p, g);              // grad.add(param, alpha=weight_decay)
    float m = e.exp_avg[i], v = e.exp_avg_sq[i];
    m = fmaf(w1, g - m, m);                                             // exp_avg.lerp_(grad, 1 - beta1)
    v = fmaf(w2 * g, g, v * beta2);                                     // exp_avg_sq.mul_(beta2).addcmul_(g, g, 1 - beta2)
    const float denom = sqrtf(v) / e.bias_correction2_sqrt + eps;
    e.exp_avg[i] = m;
    e.exp_avg_sq[i] = v;
    e.param[i] = p - e.step_size * (m / denom);                         // param.addcdiv_(exp_avg, denom, -step_size)
  }
}
}  // namespace sd

extern "C" int sd_adam_step(const sd_adam_entry* table, int n_entries, int blocks_per_entry, float beta1, float beta2,
                            float eps, float weight_decay, void* stream) {
  if (n_entries <= 0) return 0;
  SD_REQUIRE(table != nullptr && blocks_per_entry > 0, "sd_adam_step: bad table / blocks_per_entry");
  sd::adam_multi_kernel<<<dim3(blocks_per_entry, n_entries), 256, 0, (cudaStream_t)stream>>>(table, beta1, beta2, eps,
                                                                                             weight_decay);
  return sd::check_launch("adam_step");
}
