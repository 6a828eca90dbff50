// CUDA-core (fp32-accumulate) implicit-GEMM Conv1d: forward / data-gradient and weight-gradient.
// This is the exact-precision family (SD_F32 activations: fp32 in, fp32 FMA) that the fp32 parity
// tests run, and the on-device cross-check for the tcgen05 family in conv_tc.cu.  It takes any
// shape; no alignment beyond Cp % 8 == 0.
//
// Reference semantics: nn.Conv1d(kernel_size in {1,3}, padding="same", dilation=d) of
// speech_decoding/models.py:97-109,128-150,188-189 and the channel mix einsum models.py:65.
#include "common.cuh"

namespace sd {

int rownorm2_launch(const float* x, float* out, int M, int64_t D, int accumulate, cudaStream_t st);

constexpr int BM = 64, BN = 64, BK = 16;

template <typename T>
__global__ void __launch_bounds__(256)
conv_fwd_simt_kernel(sd_conv_args a, int tiles_per_sample) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int b = blockIdx.x / tiles_per_sample;
  const int t0 = (blockIdx.x % tiles_per_sample) * BM;
  const int n0 = blockIdx.y * BN;
  const int g = a.widx ? a.widx[b] : 0;
  const T* in = reinterpret_cast<const T*>(a.in) + (size_t)b * a.T * a.Kp;
  const T* w = reinterpret_cast<const T*>(a.w) + (size_t)g * a.taps * a.Np * a.Kp;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int lrow = tid >> 2, lk = (tid & 3) * 4;
  for (int j = 0; j < a.taps; ++j) {
    const int shift = (j - (a.taps - 1) / 2) * a.dil;
    const T* wj = w + (size_t)j * a.Np * a.Kp;
    for (int k0 = 0; k0 < a.Kp; k0 += BK) {
      {
        int t = t0 + lrow + shift, k = k0 + lk;
        float4 v = make_float4(0, 0, 0, 0);
        if (t >= 0 && t < a.T && (t0 + lrow) < a.T && k < a.Kp) v = Vec4<T>::ld(in + (size_t)t * a.Kp + k);
        As[lk + 0][lrow] = v.x; As[lk + 1][lrow] = v.y; As[lk + 2][lrow] = v.z; As[lk + 3][lrow] = v.w;
        int n = n0 + lrow;
        float4 u = make_float4(0, 0, 0, 0);
        if (n < a.Np && k < a.Kp) u = Vec4<T>::ld(wj + (size_t)n * a.Kp + k);
        Bs[lk + 0][lrow] = u.x; Bs[lk + 1][lrow] = u.y; Bs[lk + 2][lrow] = u.z; Bs[lk + 3][lrow] = u.w;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
        const float ar[4] = {av.x, av.y, av.z, av.w};
        const float br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) acc[i][jj] = fmaf(ar[i], br[jj], acc[i][jj]);
      }
      __syncthreads();
    }
  }

  // epilogue
  const T* res = reinterpret_cast<const T*>(a.res);
  T* pre = reinterpret_cast<T*>(a.preact);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int t = t0 + ty * 4 + i;
    if (t >= a.T) continue;
    const size_t row = ((size_t)b * a.T + t) * a.Np;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int n = n0 + tx * 4 + jj;
      if (n >= a.Np) continue;
      float v = acc[i][jj];
      if (n < a.N) {
        if (a.bias) v += a.bias[n];
        if (res) v += to_f<T>(res[row + n]);
      } else {
        v = 0.f;
      }
      if (pre) pre[row + n] = from_f<T>(v);
      if (a.act == SD_ACT_GELU) v = (n < a.N) ? gelu_t<T>(v) : 0.f;
      if (a.out_mode == SD_OUT_BTC) {
        reinterpret_cast<T*>(a.out)[row + n] = from_f<T>(v);
      } else if (n < a.N) {
        reinterpret_cast<float*>(a.out)[((size_t)b * a.N + n) * a.T + t] = v;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// weight gradient: dw[g,n,k,j] += sum_{b in g} sum_t dout[b,t,n] * in[b,t+shift_j,k]
// each CTA owns a 64(n) x 64(k) tile for one tap and a contiguous slice of the subject-sorted
// sample list; it flushes with atomics whenever the group changes.
// ---------------------------------------------------------------------------------------------------
constexpr int WT = 16;  // time rows per smem stage

template <typename T>
__global__ void __launch_bounds__(256)
conv_wgrad_simt_kernel(sd_wgrad_args a, int ktiles, int nsplit) {
  __shared__ float As[WT][BN + 4];  // dout tile [t][n]
  __shared__ float Bs[WT][BN + 4];  // in tile   [t][k]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int n0 = (blockIdx.x / ktiles) * BN, k0 = (blockIdx.x % ktiles) * BN;
  const int j = blockIdx.y;
  const int shift = (j - (a.taps - 1) / 2) * a.dil;
  const int per = (a.B + nsplit - 1) / nsplit;
  const int p0 = blockIdx.z * per, p1 = min(a.B, p0 + per);
  const bool do_bias = a.dbias && j == 0 && k0 == 0;

  float acc[4][4], accb[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    accb[i] = 0.f;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) acc[i][jj] = 0.f;
  }
  int cur_g = -1;
  int gi = 0;  // running group cursor

  auto flush = [&](int g) {
    if (g < 0) return;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int n = n0 + ty * 4 + i;
      if (n >= a.N) continue;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int k = k0 + tx * 4 + jj;
        if (k < a.K) atomicAdd(a.dw + (int64_t)g * a.gs + (int64_t)n * a.sn + (int64_t)k * a.sk + (int64_t)j * a.sj, acc[i][jj]);
        acc[i][jj] = 0.f;
      }
    }
  };

  const int lrow = tid >> 4, lc = (tid & 15) * 4;  // 16 rows x 16 float4
  for (int p = p0; p < p1; ++p) {
    const int b = a.sample_order ? a.sample_order[p] : p;
    int g = 0;
    if (a.group_offsets) {
      while (gi + 1 < a.G && p >= a.group_offsets[gi + 1]) ++gi;
      g = gi;
    }
    if (g != cur_g) { flush(cur_g); cur_g = g; }
    const T* dout = reinterpret_cast<const T*>(a.dout) + (size_t)b * a.T * a.Np;
    const T* in = reinterpret_cast<const T*>(a.in) + (size_t)b * a.T * a.Kp;
    for (int t0 = 0; t0 < a.T; t0 += WT) {
      {
        int t = t0 + lrow;
        float4 v = make_float4(0, 0, 0, 0), u = make_float4(0, 0, 0, 0);
        if (t < a.T && n0 + lc < a.Np) v = Vec4<T>::ld(dout + (size_t)t * a.Np + n0 + lc);
        int ts = t + shift;
        if (t < a.T && ts >= 0 && ts < a.T && k0 + lc < a.Kp) u = Vec4<T>::ld(in + (size_t)ts * a.Kp + k0 + lc);
        *reinterpret_cast<float4*>(&As[lrow][lc]) = v;
        *reinterpret_cast<float4*>(&Bs[lrow][lc]) = u;
      }
      __syncthreads();
#pragma unroll
      for (int tt = 0; tt < WT; ++tt) {
        float4 av = *reinterpret_cast<const float4*>(&As[tt][ty * 4]);
        float4 bv = *reinterpret_cast<const float4*>(&Bs[tt][tx * 4]);
        const float ar[4] = {av.x, av.y, av.z, av.w};
        const float br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (tx == 0) accb[i] += ar[i];
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) acc[i][jj] = fmaf(ar[i], br[jj], acc[i][jj]);
        }
      }
      __syncthreads();
    }
  }
  flush(cur_g);
  if (do_bias && tx == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int n = n0 + ty * 4 + i;
      if (n < a.N) atomicAdd(a.dbias + n, accb[i]);
    }
  }
}

int conv_fwd_simt(const sd_conv_args& a_in, cudaStream_t st) {
  sd_conv_args a = a_in;
  const bool glu = a.act == SD_ACT_GLU;
  void* final_out = a.out;
  if (glu) {
    SD_REQUIRE(a.preact != nullptr, "conv_fwd(simt): GLU needs a preact buffer");
    SD_REQUIRE(a.out_mode == SD_OUT_BTC, "conv_fwd(simt): GLU output must be BTC");
    a.out = a.preact;  // write y2 once; glu kernel produces `out`
    a.preact = nullptr;
    a.act = SD_ACT_NONE;
  }
  const int tiles = cdiv(a.T, BM);
  dim3 grid(a.B * tiles, cdiv(a.Np, BN));
  if (a.dtype == SD_F32) conv_fwd_simt_kernel<float><<<grid, 256, 0, st>>>(a, tiles);
  else conv_fwd_simt_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(a, tiles);
  if (check_launch("conv_fwd_simt")) return 1;
  if (glu) {
    const int D2 = a.N / 2, Op = (D2 + 7) / 8 * 8;
    if (sd_glu_fwd(a.out, final_out, (int64_t)a.B * a.T, D2, a.Np, Op, a.dtype, st)) return 1;
  }
  if (a.stats) {
    SD_REQUIRE(a.out_mode == SD_OUT_BTC && !glu && a.act == SD_ACT_NONE, "conv_fwd(simt): stats need a plain BTC output");
    if (sd_colstats(a.out, a.stats, (int64_t)a.B * a.T, a.Np, a.dtype, st)) return 1;
  }
  if (a.rownorm2) {
    SD_REQUIRE(a.out_mode == SD_OUT_NCT_F32, "conv_fwd(simt): rownorm2 needs the NCT fp32 output");
    if (rownorm2_launch(reinterpret_cast<const float*>(final_out), a.rownorm2, a.B, (int64_t)a.N * a.T, 1, st)) return 1;
  }
  return 0;
}

int conv_wgrad_simt(const sd_wgrad_args& a, cudaStream_t st) {
  const int ntiles = cdiv(a.Np, BN), ktiles = cdiv(a.Kp, BN);
  const int tiles = ntiles * ktiles * a.taps;
  int nsplit = cdiv(148 * 4, tiles);
  if (nsplit > a.B) nsplit = a.B;
  if (nsplit < 1) nsplit = 1;
  dim3 grid(ntiles * ktiles, a.taps, nsplit);
  if (a.dtype == SD_F32) conv_wgrad_simt_kernel<float><<<grid, 256, 0, st>>>(a, ktiles, nsplit);
  else conv_wgrad_simt_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(a, ktiles, nsplit);
  return check_launch("conv_wgrad_simt");
}

}  // namespace sd
