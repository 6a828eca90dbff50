// HBM-bound kernels of the hot path: layout conversion, weight packing, BatchNorm statistics /
// apply / backward fused with GELU, GLU, GELU backward.  All operate on the channels-last "BTC"
// activation layout (rows = B*T, Cp channels contiguous, Cp % 8 == 0) with 8/16-byte vector access.
#include "common.cuh"

namespace sd {

// ---------------------------------------------------------------------------------------------------
// (B,C,T) fp32 <-> (B,T,Cp) T : 32x32 shared-memory tile transpose, both sides coalesced
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void nct_to_btc_kernel(const float* __restrict__ x, T* __restrict__ out, int C, int Tn, int Cp) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  const float* xb = x + (size_t)b * C * Tn;
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    int c = c0 + ty + i, t = t0 + tx;
    tile[ty + i][tx] = (c < C && t < Tn) ? xb[(size_t)c * Tn + t] : 0.f;
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

// dZ (B,N,T) fp32 + p (B,T,Np) -> dp (B,T,Np) = dZ^T * gelu'(p)
template <typename T>
__global__ void gelu_bwd_nct_kernel(const float* __restrict__ dz, const T* __restrict__ p, T* __restrict__ dp,
                                    int N, int Tn, int Np) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const float* zb = dz + (size_t)b * N * Tn;
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    int c = c0 + ty + i, t = t0 + tx;
    tile[ty + i][tx] = (c < N && t < Tn) ? zb[(size_t)c * Tn + t] : 0.f;
  }
  __syncthreads();
  const size_t base = (size_t)b * Tn * Np;
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    int t = t0 + ty + i, c = c0 + tx;
    if (t < Tn && c < Np) {
      size_t o = base + (size_t)t * Np + c;
      float g = (c < N) ? tile[tx][ty + i] * gelu_grad_f(to_f<T>(p[o])) : 0.f;
      dp[o] = from_f<T>(g);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// weight packing: (N,K,taps) fp32 -> wf (taps,Np,Kp), wd (taps,Kp,Np) taps reversed
// ---------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void pack_one(const sd_pack_entry& e, int64_t start, int64_t step) {
  const int64_t total = (int64_t)e.taps * e.Np * e.Kp;
  T* wf = reinterpret_cast<T*>(e.wf);
  T* wd = reinterpret_cast<T*>(e.wd);
  for (int64_t i = start; i < total; i += step) {
    if (wf) {
      int k = (int)(i % e.Kp);
      int n = (int)((i / e.Kp) % e.Np);
      int j = (int)(i / ((int64_t)e.Kp * e.Np));
      float v = (n < e.N && k < e.K) ? e.w[((int64_t)n * e.K + k) * e.taps + j] : 0.f;
      wf[i] = from_f<T>(v);
    }
    if (wd) {
      int n = (int)(i % e.Np);
      int k = (int)((i / e.Np) % e.Kp);
      int j = (int)(i / ((int64_t)e.Kp * e.Np));
      float v = (n < e.N && k < e.K) ? e.w[((int64_t)n * e.K + k) * e.taps + (e.taps - 1 - j)] : 0.f;
      wd[i] = from_f<T>(v);
    }
  }
}

__global__ void pack_weights_kernel(const sd_pack_entry* __restrict__ table) {
  const sd_pack_entry e = table[blockIdx.y];
  int64_t start = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t step = (int64_t)gridDim.x * blockDim.x;
  if (e.dtype == SD_BF16) pack_one<__nv_bfloat16>(e, start, step);
  else pack_one<float>(e, start, step);
}

__global__ void pack_weight_single_kernel(sd_pack_entry e) {
  int64_t start = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t step = (int64_t)gridDim.x * blockDim.x;
  if (e.dtype == SD_BF16) pack_one<__nv_bfloat16>(e, start, step);
  else pack_one<float>(e, start, step);
}

// ---------------------------------------------------------------------------------------------------
// column statistics over (rows, Cp): each thread owns 4 adjacent channels, 8 row-lanes per block
// ---------------------------------------------------------------------------------------------------
constexpr int STAT_ROWS_PER_BLOCK = 256;

template <typename T, int MODE>  // MODE 0: sum,sumsq of x ; MODE 1: bn+gelu backward reduce (in place g)
__global__ void __launch_bounds__(256)
colreduce_kernel(T* __restrict__ x, const T* __restrict__ y, const float* __restrict__ ss, double* __restrict__ out,
                 int64_t rows, int Cp) {
  __shared__ float4 sA[8][32], sB[8][32];
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  const int c = (blockIdx.x * 32 + tx) * 4;
  const int64_t r0 = (int64_t)blockIdx.y * STAT_ROWS_PER_BLOCK;
  const int64_t r1 = min(rows, r0 + STAT_ROWS_PER_BLOCK);
  float4 a = make_float4(0, 0, 0, 0), q = make_float4(0, 0, 0, 0);
  if (c < Cp) {
    float4 sc, sh, mu, is;
    if (MODE == 1) {
      sc = *reinterpret_cast<const float4*>(ss + c);
      sh = *reinterpret_cast<const float4*>(ss + Cp + c);
      mu = *reinterpret_cast<const float4*>(ss + 2 * Cp + c);
      is = *reinterpret_cast<const float4*>(ss + 3 * Cp + c);
    }
    for (int64_t r = r0 + ty; r < r1; r += 8) {
      float4 v = Vec4<T>::ld(x + r * Cp + c);
      if (MODE == 0) {
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        q.x += v.x * v.x; q.y += v.y * v.y; q.z += v.z * v.z; q.w += v.w * v.w;
      } else {
        float4 yy = Vec4<T>::ld(y + r * Cp + c);
        float4 g;
        g.x = v.x * gelu_grad_f(yy.x * sc.x + sh.x);
        g.y = v.y * gelu_grad_f(yy.y * sc.y + sh.y);
        g.z = v.z * gelu_grad_f(yy.z * sc.z + sh.z);
        g.w = v.w * gelu_grad_f(yy.w * sc.w + sh.w);
        Vec4<T>::st(x + r * Cp + c, g);
        a.x += g.x; a.y += g.y; a.z += g.z; a.w += g.w;
        q.x += g.x * (yy.x - mu.x) * is.x; q.y += g.y * (yy.y - mu.y) * is.y;
        q.z += g.z * (yy.z - mu.z) * is.z; q.w += g.w * (yy.w - mu.w) * is.w;
      }
    }
  }
  sA[ty][tx] = a; sB[ty][tx] = q;
  __syncthreads();
  if (ty == 0 && c < Cp) {
#pragma unroll
    for (int i = 1; i < 8; ++i) {
      float4 u = sA[i][tx], w = sB[i][tx];
      a.x += u.x; a.y += u.y; a.z += u.z; a.w += u.w;
      q.x += w.x; q.y += w.y; q.z += w.z; q.w += w.w;
    }
    atomicAdd(out + c + 0, (double)a.x); atomicAdd(out + c + 1, (double)a.y);
    atomicAdd(out + c + 2, (double)a.z); atomicAdd(out + c + 3, (double)a.w);
    atomicAdd(out + Cp + c + 0, (double)q.x); atomicAdd(out + Cp + c + 1, (double)q.y);
    atomicAdd(out + Cp + c + 2, (double)q.z); atomicAdd(out + Cp + c + 3, (double)q.w);
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
__global__ void bn_gelu_fwd_kernel(const T* __restrict__ y, const float* __restrict__ ss, T* __restrict__ u,
                                   int64_t n4, int Cp) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (; i < n4; i += step) {
    int c = (int)((i * 4) % Cp);
    float4 sc = *reinterpret_cast<const float4*>(ss + c);
    float4 sh = *reinterpret_cast<const float4*>(ss + Cp + c);
    float4 v = Vec4<T>::ld(y + i * 4);
    v.x = gelu_f(v.x * sc.x + sh.x); v.y = gelu_f(v.y * sc.y + sh.y);
    v.z = gelu_f(v.z * sc.z + sh.z); v.w = gelu_f(v.w * sc.w + sh.w);
    Vec4<T>::st(u + i * 4, v);
  }
}

// dy = scale * (g - sum_g/n - xhat * sum_gx/n)
template <typename T>
__global__ void bn_bwd_apply_kernel(T* __restrict__ g, const T* __restrict__ y, const float* __restrict__ ss,
                                    const double* __restrict__ red, float* __restrict__ dgamma,
                                    float* __restrict__ dbeta, int64_t rows, int64_t n_stat, int C, int Cp, int training) {
  const int64_t n4 = rows * Cp / 4;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t step = (int64_t)gridDim.x * blockDim.x;
  if (blockIdx.x == 0) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      if (dbeta) dbeta[c] = (float)red[c];
      if (dgamma) dgamma[c] = (float)red[Cp + c];
    }
  }
  const float invn = 1.0f / (float)n_stat;
  for (; i < n4; i += step) {
    int c = (int)((i * 4) % Cp);
    float4 sc = *reinterpret_cast<const float4*>(ss + c);
    float4 v = Vec4<T>::ld(g + i * 4);
    if (training) {
      float4 mu = *reinterpret_cast<const float4*>(ss + 2 * Cp + c);
      float4 is = *reinterpret_cast<const float4*>(ss + 3 * Cp + c);
      float4 yy = Vec4<T>::ld(y + i * 4);
      float sg0 = (float)red[c] * invn, sg1 = (float)red[c + 1] * invn, sg2 = (float)red[c + 2] * invn,
            sg3 = (float)red[c + 3] * invn;
      float sx0 = (float)red[Cp + c] * invn, sx1 = (float)red[Cp + c + 1] * invn,
            sx2 = (float)red[Cp + c + 2] * invn, sx3 = (float)red[Cp + c + 3] * invn;
      v.x = sc.x * (v.x - sg0 - (yy.x - mu.x) * is.x * sx0);
      v.y = sc.y * (v.y - sg1 - (yy.y - mu.y) * is.y * sx1);
      v.z = sc.z * (v.z - sg2 - (yy.z - mu.z) * is.z * sx2);
      v.w = sc.w * (v.w - sg3 - (yy.w - mu.w) * is.w * sx3);
    } else {
      v.x *= sc.x; v.y *= sc.y; v.z *= sc.z; v.w *= sc.w;
    }
    Vec4<T>::st(g + i * 4, v);
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

template <typename T>
__global__ void gelu_bwd_kernel(T* __restrict__ du, const T* __restrict__ p, int64_t n4) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (; i < n4; i += step) {
    float4 v = Vec4<T>::ld(du + i * 4), q = Vec4<T>::ld(p + i * 4);
    v.x *= gelu_grad_f(q.x); v.y *= gelu_grad_f(q.y); v.z *= gelu_grad_f(q.z); v.w *= gelu_grad_f(q.w);
    Vec4<T>::st(du + i * 4, v);
  }
}

static inline int ew_grid(int64_t n, int threads) {
  int64_t b = (n + threads - 1) / threads;
  const int64_t cap = 148 * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
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

int sd_btc_to_nct(const void* in, float* out, int B, int C, int T_, int Cp, int dtype, void* stream) {
  dim3 grid(cdiv(T_, 32), cdiv(Cp, 32), B), block(32, 8);
  DISPATCH_DTYPE(dtype, btc_to_nct_kernel<T><<<grid, block, 0, (cudaStream_t)stream>>>((const T*)in, out, C, T_, Cp));
  return check_launch("btc_to_nct");
}

int sd_gelu_bwd_nct(const float* dz, const void* p, void* dp, int B, int N, int T_, int Np, int dtype, void* stream) {
  dim3 grid(cdiv(T_, 32), cdiv(Np, 32), B), block(32, 8);
  DISPATCH_DTYPE(dtype, gelu_bwd_nct_kernel<T><<<grid, block, 0, (cudaStream_t)stream>>>(dz, (const T*)p, (T*)dp, N, T_, Np));
  return check_launch("gelu_bwd_nct");
}

int sd_pack_weight(const float* w, void* wf, void* wd, int N, int K, int taps, int Np, int Kp, int dtype,
                   void* stream) {
  sd_pack_entry e{w, wf, wd, N, K, taps, Np, Kp, dtype};
  int64_t total = (int64_t)taps * Np * Kp;
  pack_weight_single_kernel<<<ew_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(e);
  return check_launch("pack_weight");
}

int sd_pack_weights(const sd_pack_entry* table, int n, void* stream) {
  if (n <= 0) return 0;
  pack_weights_kernel<<<dim3(64, n), 256, 0, (cudaStream_t)stream>>>(table);
  return check_launch("pack_weights");
}

int sd_colstats(const void* x, double* stats, int64_t rows, int Cp, int dtype, void* stream) {
  SD_REQUIRE(Cp % 8 == 0, "sd_colstats: Cp %% 8 != 0");
  dim3 grid(cdiv(Cp, 128), cdiv(rows, STAT_ROWS_PER_BLOCK)), block(32, 8);
  DISPATCH_DTYPE(dtype, colreduce_kernel<T, 0><<<grid, block, 0, (cudaStream_t)stream>>>((T*)x, nullptr, nullptr, stats, rows, Cp));
  return check_launch("colstats");
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
  int64_t n4 = rows * Cp / 4;
  DISPATCH_DTYPE(dtype, bn_gelu_fwd_kernel<T><<<ew_grid(n4, 256), 256, 0, (cudaStream_t)stream>>>((const T*)y, ss, (T*)u, n4, Cp));
  return check_launch("bn_gelu_fwd");
}

int sd_bn_gelu_bwd_reduce(void* du_g, const void* y, const float* ss, double* red, int64_t rows, int Cp, int dtype,
                          void* stream) {
  dim3 grid(cdiv(Cp, 128), cdiv(rows, STAT_ROWS_PER_BLOCK)), block(32, 8);
  DISPATCH_DTYPE(dtype, colreduce_kernel<T, 1><<<grid, block, 0, (cudaStream_t)stream>>>((T*)du_g, (const T*)y, ss, red, rows, Cp));
  return check_launch("bn_gelu_bwd_reduce");
}

int sd_bn_bwd_apply(void* g_dy, const void* y, const float* ss, const double* red, float* dgamma, float* dbeta,
                    int64_t rows, int64_t n_stat, int C, int Cp, int training, int dtype, void* stream) {
  int64_t n4 = rows * Cp / 4;
  DISPATCH_DTYPE(dtype, bn_bwd_apply_kernel<T><<<ew_grid(n4, 256), 256, 0, (cudaStream_t)stream>>>((T*)g_dy, (const T*)y, ss, red, dgamma, dbeta, rows, n_stat, C, Cp, training));
  return check_launch("bn_bwd_apply");
}

int sd_glu_fwd(const void* y2, void* out, int64_t rows, int D2, int Np, int Op, int dtype, void* stream) {
  DISPATCH_DTYPE(dtype, glu_fwd_kernel<T><<<ew_grid(rows * Op, 256), 256, 0, (cudaStream_t)stream>>>((const T*)y2, (T*)out, rows, D2, Np, Op));
  return check_launch("glu_fwd");
}

int sd_glu_bwd(const void* dout, const void* y2, void* dy2, int64_t rows, int D2, int Np, int Op, int dtype,
               void* stream) {
  DISPATCH_DTYPE(dtype, glu_bwd_kernel<T><<<ew_grid(rows * Np, 256), 256, 0, (cudaStream_t)stream>>>((const T*)dout, (const T*)y2, (T*)dy2, rows, D2, Np, Op));
  return check_launch("glu_bwd");
}

int sd_gelu_bwd(void* du_dp, const void* p, int64_t rows, int Cp, int dtype, void* stream) {
  int64_t n4 = rows * Cp / 4;
  DISPATCH_DTYPE(dtype, gelu_bwd_kernel<T><<<ew_grid(n4, 256), 256, 0, (cudaStream_t)stream>>>((T*)du_dp, (const T*)p, n4));
  return check_launch("gelu_bwd");
}

}  // extern "C"
