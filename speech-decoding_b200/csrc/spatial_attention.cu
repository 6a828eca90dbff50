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

template <typename T>
__global__ void __launch_bounds__(256)
sa_weights_fwd_kernel(const float* __restrict__ z_ri, const float* __restrict__ cos_t, const float* __restrict__ sin_t,
                      const float* __restrict__ mask, float* __restrict__ w_soft, T* __restrict__ w_packed, int D1,
                      int K2, int C, int Cp) {
  extern __shared__ float smem[];
  float* zs = smem;               // 2*K2 interleaved re/im
  float* logit = smem + 2 * K2;   // C
  __shared__ float red[32];
  const int d = blockIdx.x, tid = threadIdx.x;
  if (d >= D1) {  // zero rows of the padded weight matrix
    for (int c = tid; c < Cp; c += blockDim.x) w_packed[(size_t)d * Cp + c] = from_f<T>(0.f);
    return;
  }
  for (int i = tid; i < 2 * K2; i += blockDim.x) zs[i] = z_ri[(size_t)d * 2 * K2 + i];
  __syncthreads();
  // logits: the K^2-long contraction is split over the 8 warps; each lane keeps 4 independent fp32 partial
  // sums (<= K^2/32 terms each) and the 32 partials per logit are combined in fp64 -- B200's fp64 FMA rate is
  // ~1/64 of fp32, so the bulk of the sum must not be fp64; the fp32 partials carry ~1e-6 absolute error on
  // logits of magnitude <= 40, far inside the 1e-4 budget.  Lanes run over sensors: coalesced table reads.
  double* part = reinterpret_cast<double*>(smem + 2 * K2 + ((C + 1) & ~1));   // [8][C]
  {
    const int warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
    const int m0 = (int)((long long)K2 * warp / nw), m1 = (int)((long long)K2 * (warp + 1) / nw);
    for (int c = lane; c < C; c += 32) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      int m = m0;
      for (; m + 1 < m1; m += 2) {
        a0 = fmaf(zs[2 * m], cos_t[(size_t)m * C + c], a0);
        a1 = fmaf(zs[2 * m + 1], sin_t[(size_t)m * C + c], a1);
        a2 = fmaf(zs[2 * m + 2], cos_t[(size_t)(m + 1) * C + c], a2);
        a3 = fmaf(zs[2 * m + 3], sin_t[(size_t)(m + 1) * C + c], a3);
      }
      if (m < m1) {
        a0 = fmaf(zs[2 * m], cos_t[(size_t)m * C + c], a0);
        a1 = fmaf(zs[2 * m + 1], sin_t[(size_t)m * C + c], a1);
      }
      part[warp * C + c] = ((double)a0 + (double)a1) + ((double)a2 + (double)a3);
    }
  }
  __syncthreads();
  float lmax = -INFINITY;
  for (int c = tid; c < C; c += blockDim.x) {
    double acc = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) acc += part[w * C + c];
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
__global__ void __launch_bounds__(256)
sa_weights_bwd_kernel(const float* __restrict__ dwm, const float* __restrict__ w_soft, const float* __restrict__ mask,
                      const float* __restrict__ cos_t, const float* __restrict__ sin_t, float* __restrict__ dz_ri,
                      int K2, int C) {
  extern __shared__ float da[];  // C
  __shared__ float red[32];
  const int d = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  float dot = 0.f;
  for (int c = tid; c < C; c += blockDim.x) {
    float g = dwm[(size_t)d * C + c] * (mask ? mask[c] : 1.f);
    float w = w_soft[(size_t)d * C + c];
    da[c] = g;
    dot += g * w;
  }
  dot = warp_sum(dot);
  if (lane == 0) red[warp] = dot;
  __syncthreads();
  dot = 0.f;
  for (int i = 0; i < nw; ++i) dot += red[i];
  for (int c = tid; c < C; c += blockDim.x) da[c] = w_soft[(size_t)d * C + c] * (da[c] - dot);
  __syncthreads();
  // z.grad[d,m] = sum_c da[d,c] * (cos[m,c] + i sin[m,c]): one thread per frequency m walks its table rows
  // (L1/L2-resident: 2 x K^2 x C floats), da[c] is a shared-memory broadcast; no cross-lane reductions.
  for (int m = blockIdx.y * blockDim.x + tid; m < K2; m += gridDim.y * blockDim.x) {
    const float* cr = cos_t + (size_t)m * C;
    const float* sr = sin_t + (size_t)m * C;
    float re0 = 0.f, re1 = 0.f, im0 = 0.f, im1 = 0.f;
    int c = 0;
    for (; c + 1 < C; c += 2) {
      re0 = fmaf(da[c], cr[c], re0);
      im0 = fmaf(da[c], sr[c], im0);
      re1 = fmaf(da[c + 1], cr[c + 1], re1);
      im1 = fmaf(da[c + 1], sr[c + 1], im1);
    }
    if (c < C) { re0 = fmaf(da[c], cr[c], re0); im0 = fmaf(da[c], sr[c], im0); }
    *reinterpret_cast<float2*>(dz_ri + ((size_t)d * K2 + m) * 2) = make_float2(re0 + re1, im0 + im1);
  }
}

}  // namespace sd

using namespace sd;

extern "C" {

int sd_sa_weights_fwd(const float* z_ri, const float* cos_t, const float* sin_t, const float* mask, float* w_soft,
                      void* w_packed, int D1, int K2, int C, int D1p, int Cp, int dtype, void* stream) {
  size_t smem = (size_t)(2 * K2 + ((C + 1) & ~1)) * sizeof(float) + (size_t)8 * C * sizeof(double);
  SD_REQUIRE(smem <= 200 * 1024, "sd_sa_weights_fwd: K^2/C too large for shared memory");
  if (dtype == SD_F32) {
    if (smem > 48 * 1024) SD_CUDA(cudaFuncSetAttribute(sa_weights_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sa_weights_fwd_kernel<float><<<D1p, 256, smem, (cudaStream_t)stream>>>(z_ri, cos_t, sin_t, mask, w_soft, (float*)w_packed, D1, K2, C, Cp);
  } else if (dtype == SD_BF16) {
    if (smem > 48 * 1024) SD_CUDA(cudaFuncSetAttribute(sa_weights_fwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sa_weights_fwd_kernel<__nv_bfloat16><<<D1p, 256, smem, (cudaStream_t)stream>>>(z_ri, cos_t, sin_t, mask, w_soft, (__nv_bfloat16*)w_packed, D1, K2, C, Cp);
  } else {
    set_error("sd_sa_weights_fwd: bad dtype");
    return 1;
  }
  return check_launch("sa_weights_fwd");
}

int sd_sa_weights_bwd(const float* dwm, const float* w_soft, const float* mask, const float* cos_t,
                      const float* sin_t, float* dz_ri, int D1, int K2, int C, void* stream) {
  sa_weights_bwd_kernel<<<dim3(D1, (K2 + 255) / 256 < 4 ? (K2 + 255) / 256 : 4), 256, C * sizeof(float), (cudaStream_t)stream>>>(dwm, w_soft, mask, cos_t, sin_t, dz_ri, K2, C);
  return check_launch("sa_weights_bwd");
}

}  // extern "C"
