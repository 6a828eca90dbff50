// SpatialAttention mixing weights (reference: speech_decoding/models.py:45-65) and SpatialDropout
// (models.py:77-86) folded in as a column mask on the softmax weights.
//
//   a[d,c]  = sum_m Re z[d,m] cos[m,c] + Im z[d,m] sin[m,c]        (models.py:49-53)
//   w       = softmax_c(a)                                          (models.py:58)
//   w~      = w * mask[c]   (dropping sensor c == zeroing column c; no renormalisation, models.py:84)
//
// The logits reach |a| ~ 40 at init, so they are accumulated in fp64 and the softmax is fp32 in every
// precision mode.  The batch-dependent part (w~ applied to X) runs through the implicit-GEMM conv op.
#include "common.cuh"

namespace sd {

constexpr int SA_DT = 16;     // rows of z (output channels d) per block
constexpr int SA_MC = 128;    // frequencies per shared-memory chunk
constexpr int SA_BWD_DT = 8;  // rows of z per block, backward (more, smaller blocks: the kernel is latency-bound)
constexpr int SA_BWD_M = 128; // frequencies (threads) per block, backward

// Partial logits: part[mp][d][c] = sum_{m in partition mp} Re z[d,m] cos[m,c] + Im z[d,m] sin[m,c].
// A block owns SA_DT rows of z and one partition of the K^2 frequencies: every table element it reads
// (coalesced over sensors c) feeds SA_DT FMAs, so the tables are read ceil(D1/16) times in total instead
// of D1 times.  fp32 accumulation over <= K^2/parts terms; the partitions are combined in fp64 by the
// softmax kernel (B200's fp64 FMA rate is ~1/64 of fp32, so fp64 is kept out of the inner loop).
__global__ void __launch_bounds__(256)
sa_logits_kernel(const float* __restrict__ z_ri, const float* __restrict__ cos_t, const float* __restrict__ sin_t,
                 float* __restrict__ part, int D1, int K2, int C, int mparts) {
  __shared__ __align__(16) float zs[SA_MC][2 * SA_DT];   // [m][d] (re,im) pairs: one frequency = 8 LDS.128
  const int d0 = blockIdx.x * SA_DT, mp = blockIdx.y, tid = threadIdx.x;
  const int m_lo = (int)((long long)K2 * mp / mparts), m_hi = (int)((long long)K2 * (mp + 1) / mparts);
  for (int c0 = 0; c0 < C; c0 += 256) {
    const int c = c0 + tid;
    float acc[SA_DT];
#pragma unroll
    for (int i = 0; i < SA_DT; ++i) acc[i] = 0.f;
    for (int mc = m_lo; mc < m_hi; mc += SA_MC) {
      const int mn = min(SA_MC, m_hi - mc);
      __syncthreads();
      for (int i = tid; i < SA_DT * 2 * mn; i += 256) {
        const int dd = i / (2 * mn), r = i % (2 * mn);       // r = 2*m + {re,im}: coalesced global read
        zs[r >> 1][2 * dd + (r & 1)] = (d0 + dd < D1) ? z_ri[((size_t)(d0 + dd) * K2 + mc) * 2 + r] : 0.f;
      }
      __syncthreads();
      if (c < C) {
#pragma unroll 4
        for (int m = 0; m < mn; ++m) {
          const float cv = cos_t[(size_t)(mc + m) * C + c], sv = sin_t[(size_t)(mc + m) * C + c];
          const float4* zr = reinterpret_cast<const float4*>(&zs[m][0]);
#pragma unroll
          for (int i = 0; i < SA_DT / 2; ++i) {
            const float4 q = zr[i];                          // (re,im) of rows 2i and 2i+1
            acc[2 * i] = fmaf(q.x, cv, fmaf(q.y, sv, acc[2 * i]));
            acc[2 * i + 1] = fmaf(q.z, cv, fmaf(q.w, sv, acc[2 * i + 1]));
          }
        }
      }
    }
    if (c < C) {
#pragma unroll
      for (int i = 0; i < SA_DT; ++i)
        if (d0 + i < D1) part[((size_t)mp * D1 + d0 + i) * C + c] = acc[i];
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
sa_softmax_kernel(const float* __restrict__ part, const float* __restrict__ mask, float* __restrict__ w_soft,
                  T* __restrict__ w_packed, int D1, int C, int Cp, int mparts) {
  extern __shared__ float logit[];   // C
  __shared__ float red[32];
  const int d = blockIdx.x, tid = threadIdx.x;
  if (d >= D1) {  // zero rows of the padded weight matrix
    for (int c = tid; c < Cp; c += blockDim.x) w_packed[(size_t)d * Cp + c] = from_f<T>(0.f);
    return;
  }
  float lmax = -INFINITY;
  for (int c = tid; c < C; c += blockDim.x) {
    double acc = 0.0;
    for (int mp = 0; mp < mparts; ++mp) acc += (double)part[((size_t)mp * D1 + d) * C + c];
    logit[c] = (float)acc;
    lmax = fmaxf(lmax, (float)acc);
  }
  lmax = warp_max(lmax);
  if ((tid & 31) == 0) red[tid >> 5] = lmax;
  __syncthreads();
  lmax = red[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); ++i) lmax = fmaxf(lmax, red[i]);
  __syncthreads();
  float lsum = 0.f;
  for (int c = tid; c < C; c += blockDim.x) {
    float e = expf(logit[c] - lmax);
    logit[c] = e;
    lsum += e;
  }
  lsum = warp_sum(lsum);
  if ((tid & 31) == 0) red[tid >> 5] = lsum;
  __syncthreads();
  lsum = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) lsum += red[i];
  const float inv = 1.0f / lsum;
  for (int c = tid; c < Cp; c += blockDim.x) {
    float w = 0.f;
    if (c < C) {
      w = logit[c] * inv;
      w_soft[(size_t)d * C + c] = w;
      if (mask) w *= mask[c];
    }
    w_packed[(size_t)d * Cp + c] = from_f<T>(w);
  }
}

// dw~ -> da (softmax backward through the mask) -> z.grad = da·cos^T + i da·sin^T   (appendix A.1)
// Block = SA_BWD_DT rows d x SA_BWD_M frequencies m; thread = frequency m with 2*SA_BWD_DT accumulators; the tables are
// read through their TRANSPOSES (C, K^2) so that lanes (consecutive m) are coalesced; da is a smem broadcast.
__global__ void __launch_bounds__(SA_BWD_M)
sa_weights_bwd_kernel(const float* __restrict__ dwm, const float* __restrict__ w_soft, const float* __restrict__ mask,
                      const float* __restrict__ cosT, const float* __restrict__ sinT, float* __restrict__ dz_ri,
                      int D1, int K2, int C) {
  extern __shared__ float da[];  // [SA_BWD_DT][C]
  const int d0 = blockIdx.x * SA_BWD_DT, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // softmax backward per row: da = w * (g - sum(g*w)), g = dwm * mask    (one warp per row, 2 rows per warp)
  for (int dd = warp; dd < SA_BWD_DT; dd += SA_BWD_M / 32) {
    const int d = d0 + dd;
    float dot = 0.f;
    if (d < D1)
      for (int c = lane; c < C; c += 32) dot += dwm[(size_t)d * C + c] * (mask ? mask[c] : 1.f) * w_soft[(size_t)d * C + c];
    dot = warp_sum(dot);
    for (int c = lane; c < C; c += 32) {
      float v = 0.f;
      if (d < D1) v = w_soft[(size_t)d * C + c] * (dwm[(size_t)d * C + c] * (mask ? mask[c] : 1.f) - dot);
      da[dd * C + c] = v;
    }
  }
  __syncthreads();
  const int m = blockIdx.y * SA_BWD_M + tid;
  if (m >= K2) return;
  float re[SA_BWD_DT], im[SA_BWD_DT];
#pragma unroll
  for (int i = 0; i < SA_BWD_DT; ++i) re[i] = im[i] = 0.f;
#pragma unroll 4
  for (int c = 0; c < C; ++c) {
    const float cv = cosT[(size_t)c * K2 + m], sv = sinT[(size_t)c * K2 + m];
#pragma unroll
    for (int i = 0; i < SA_BWD_DT; ++i) {
      const float a = da[i * C + c];
      re[i] = fmaf(a, cv, re[i]);
      im[i] = fmaf(a, sv, im[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < SA_BWD_DT; ++i)
    if (d0 + i < D1) *reinterpret_cast<float2*>(dz_ri + ((size_t)(d0 + i) * K2 + m) * 2) = make_float2(re[i], im[i]);
}

}  // namespace sd

using namespace sd;

extern "C" {

int sd_sa_weights_fwd(const float* z_ri, const float* cos_t, const float* sin_t, const float* mask, float* w_soft,
                      void* w_packed, float* scratch, int D1, int K2, int C, int D1p, int Cp, int dtype, void* stream) {
  const int mparts = SD_SA_MPARTS;
  SD_REQUIRE(scratch != nullptr, "sd_sa_weights_fwd: scratch (SD_SA_MPARTS*D1*C floats) is null");
  SD_REQUIRE((size_t)C * sizeof(float) <= 48 * 1024, "sd_sa_weights_fwd: too many sensors");
  cudaStream_t st = (cudaStream_t)stream;
  sa_logits_kernel<<<dim3(cdiv(D1, SA_DT), mparts), 256, 0, st>>>(z_ri, cos_t, sin_t, scratch, D1, K2, C, mparts);
  if (check_launch("sa_logits")) return 1;
  if (dtype == SD_F32)
    sa_softmax_kernel<float><<<D1p, 256, C * sizeof(float), st>>>(scratch, mask, w_soft, (float*)w_packed, D1, C, Cp, mparts);
  else if (dtype == SD_BF16)
    sa_softmax_kernel<__nv_bfloat16><<<D1p, 256, C * sizeof(float), st>>>(scratch, mask, w_soft, (__nv_bfloat16*)w_packed, D1, C, Cp, mparts);
  else {
    set_error("sd_sa_weights_fwd: bad dtype");
    return 1;
  }
  return check_launch("sa_softmax");
}

int sd_sa_weights_bwd(const float* dwm, const float* w_soft, const float* mask, const float* cos_T, const float* sin_T,
                      float* dz_ri, int D1, int K2, int C, void* stream) {
  const size_t smem = (size_t)SA_BWD_DT * C * sizeof(float);
  SD_REQUIRE(smem <= 48 * 1024, "sd_sa_weights_bwd: too many sensors");
  sa_weights_bwd_kernel<<<dim3(cdiv(D1, SA_BWD_DT), cdiv(K2, SA_BWD_M)), SA_BWD_M, smem, (cudaStream_t)stream>>>(dwm, w_soft, mask, cos_T, sin_T,
                                                                                                  dz_ri, D1, K2, C);
  return check_launch("sa_weights_bwd");
}

}  // extern "C"
