// CLIPLoss (reference: speech_decoding/utils/loss.py:38-84), CUDA-core fp32 family.
//
//   x = Y (speech, no grad), z = Z (brain, grad)                      train.py:191
//   logits[i,j] = e^temp * <x_i, z_j> / (|x_i| |z_j|)                 loss.py:64-71
//   loss = (CE(logits, arange) + CE(logits^T, arange)) / 2            loss.py:79
//
// The rows are never normalised in memory: one streaming pass produces raw dot products and squared
// norms, the (M x N) epilogue applies norms and temperature, and the backward is one more streaming
// pass dz = coef^T x - cz * z  (SURVEY appendix A.5).  Multi-GPU: x rows are the all-gathered global
// batch, z rows the local shard; only the (M,2) row statistics cross ranks.
#include "common.cuh"

namespace sd {

// ---- squared row norms ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rownorm2_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t D,
                                                       int64_t chunk) {
  __shared__ float red[8];
  const int i = blockIdx.x;
  const int64_t d0 = (int64_t)blockIdx.y * chunk, d1 = min(D, d0 + chunk);
  const float* row = x + (size_t)i * D;
  float s = 0.f;
  if ((D & 3) == 0) {
    for (int64_t d = d0 + threadIdx.x * 4; d < d1; d += 1024) {
      float4 v = *reinterpret_cast<const float4*>(row + d);
      s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
  } else {
    for (int64_t d = d0 + threadIdx.x; d < d1; d += 256) s += row[d] * row[d];
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(out + i, t);
  }
}

int rownorm2_launch(const float* x, float* out, int M, int64_t D, int accumulate, cudaStream_t st) {
  if (!accumulate) SD_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * M, st));
  int splits = cdiv(148 * 8, M);
  int64_t chunk = (D + splits - 1) / splits;
  chunk = (chunk + 1023) / 1024 * 1024;
  splits = cdiv(D, chunk);
  rownorm2_kernel<<<dim3(M, splits), 256, 0, st>>>(x, out, D, chunk);
  return check_launch("rownorm2");
}

// ---- fp32 rows -> bf16 rows + squared norms of the ROUNDED rows (data-parallel CLIP transport) ------
// One pass: the gathered speech rows travel over NVLink and are re-read by the similarity / gradient GEMMs
// in bf16; the norms are those of the values the GEMMs see, so the cosine stays self-consistent.
__global__ void __launch_bounds__(256) cast_rows_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                             float* __restrict__ out, int64_t D, int64_t chunk) {
  __shared__ float red[8];
  const int i = blockIdx.x;
  const int64_t d0 = (int64_t)blockIdx.y * chunk, d1 = min(D, d0 + chunk);
  const float* row = x + (size_t)i * D;
  __nv_bfloat16* orow = y + (size_t)i * D;
  float s = 0.f;
  for (int64_t d = d0 + threadIdx.x * 8; d < d1; d += 2048) {     // D % 8 == 0 on this path
    const float4 a = *reinterpret_cast<const float4*>(row + d), b = *reinterpret_cast<const float4*>(row + d + 4);
    __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w);
    __nv_bfloat162 p2 = __floats2bfloat162_rn(b.x, b.y), p3 = __floats2bfloat162_rn(b.z, b.w);
    uint4 o;
    o.x = *reinterpret_cast<uint32_t*>(&p0); o.y = *reinterpret_cast<uint32_t*>(&p1);
    o.z = *reinterpret_cast<uint32_t*>(&p2); o.w = *reinterpret_cast<uint32_t*>(&p3);
    *reinterpret_cast<uint4*>(orow + d) = o;
    const float2 f0 = __bfloat1622float2(p0), f1 = __bfloat1622float2(p1), f2 = __bfloat1622float2(p2), f3 = __bfloat1622float2(p3);
    s += f0.x * f0.x + f0.y * f0.y + f1.x * f1.x + f1.y * f1.y + f2.x * f2.x + f2.y * f2.y + f3.x * f3.x + f3.y * f3.y;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(out + i, t);
  }
}

// squared norms of rows that are ALREADY bf16 (speech embeddings shipped from the host in bf16: they are a frozen wav2vec2
// output, rounded once when the dataset is built, so the step never touches an fp32 copy of them)
__global__ void __launch_bounds__(256) rownorm2_bf16_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, int64_t D,
                                                            int64_t chunk) {
  __shared__ float red[8];
  const int i = blockIdx.x;
  const int64_t d0 = (int64_t)blockIdx.y * chunk, d1 = min(D, d0 + chunk);
  const __nv_bfloat16* row = x + (size_t)i * D;
  float s = 0.f;
  for (int64_t d = d0 + threadIdx.x * 8; d < d1; d += 2048) {     // D % 8 == 0 on this path
    const uint4 q = *reinterpret_cast<const uint4*>(row + d);
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[k]));
      s = fmaf(f.x, f.x, fmaf(f.y, f.y, s));
    }
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(out + i, t);
  }
}

// coefT[j, i] (bf16, row stride Mp) = coef[i, j]: A operand of the bf16 gradient GEMM
__global__ void coef_t_bf16_kernel(const float* __restrict__ coef, __nv_bfloat16* __restrict__ ct, int M, int N, int Mp) {
  __shared__ float tile[32][33];
  const int i0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int i = i0 + r, j = j0 + threadIdx.x;
    tile[r][threadIdx.x] = (i < M && j < N) ? coef[(size_t)i * N + j] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int j = j0 + r, i = i0 + threadIdx.x;
    if (j < N && i < Mp) ct[(size_t)j * Mp + i] = __float2bfloat16_rn(tile[threadIdx.x][r]);
  }
}

// ---- dots[i,j] += sum_d x[i,d] z[j,d]  (split-K, 64x64 tiles) -------------------------------------
constexpr int CB = 64, CK = 16;

__device__ __forceinline__ float4 ld4_guard(const float* row, int64_t d, int64_t dend, bool vec) {
  if (vec && d + 3 < dend) return *reinterpret_cast<const float4*>(row + d);
  float4 v = make_float4(0, 0, 0, 0);
  if (d < dend) v.x = row[d];
  if (d + 1 < dend) v.y = row[d + 1];
  if (d + 2 < dend) v.z = row[d + 2];
  if (d + 3 < dend) v.w = row[d + 3];
  return v;
}

__global__ void __launch_bounds__(256)
clip_dots_kernel(const float* __restrict__ x, const float* __restrict__ z, float* __restrict__ dots, int M, int N,
                 int64_t D, int64_t chunk) {
  __shared__ float As[CK][CB + 4];
  __shared__ float Bs[CK][CB + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int i0 = blockIdx.x * CB, j0 = blockIdx.y * CB;
  const int64_t d0 = (int64_t)blockIdx.z * chunk, d1 = min(D, d0 + chunk);
  const bool vec = (D & 3) == 0;
  float acc[4][4] = {};
  const int lrow = tid >> 2, lk = (tid & 3) * 4;
  for (int64_t k0 = d0; k0 < d1; k0 += CK) {
    float4 v = make_float4(0, 0, 0, 0), u = make_float4(0, 0, 0, 0);
    if (i0 + lrow < M) v = ld4_guard(x + (size_t)(i0 + lrow) * D, k0 + lk, d1, vec);
    if (j0 + lrow < N) u = ld4_guard(z + (size_t)(j0 + lrow) * D, k0 + lk, d1, vec);
    As[lk][lrow] = v.x; As[lk + 1][lrow] = v.y; As[lk + 2][lrow] = v.z; As[lk + 3][lrow] = v.w;
    Bs[lk][lrow] = u.x; Bs[lk + 1][lrow] = u.y; Bs[lk + 2][lrow] = u.z; Bs[lk + 3][lrow] = u.w;
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < CK; ++kk) {
      float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float ar[4] = {av.x, av.y, av.z, av.w};
      const float br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int r = i0 + ty * 4 + i, c = j0 + tx * 4 + j;
      if (r < M && c < N) atomicAdd(dots + (size_t)r * N + c, acc[i][j]);
    }
}

// ---- phase 1: scaled logits, row (max,sumexp), column LSE ------------------------------------------
__global__ void __launch_bounds__(256)
clip_rows_kernel(const float* __restrict__ dots, const float* __restrict__ xn2, const float* __restrict__ zn2,
                 const float* __restrict__ temp, float* __restrict__ logits, float* __restrict__ row_stat, int N) {
  __shared__ float red[8];
  const int i = blockIdx.x, tid = threadIdx.x;
  const float s = expf(temp[0]) * rsqrtf(xn2[i]);
  float mx = -INFINITY;
  for (int j = tid; j < N; j += 256) {
    float l = dots[(size_t)i * N + j] * s * rsqrtf(zn2[j]);
    logits[(size_t)i * N + j] = l;
    mx = fmaxf(mx, l);
  }
  mx = warp_max(mx);
  if ((tid & 31) == 0) red[tid >> 5] = mx;
  __syncthreads();
  mx = red[0];
  for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  float sum = 0.f;
  for (int j = tid; j < N; j += 256) sum += expf(logits[(size_t)i * N + j] - mx);
  sum = warp_sum(sum);
  if ((tid & 31) == 0) red[tid >> 5] = sum;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    row_stat[2 * i] = mx;
    row_stat[2 * i + 1] = t;
  }
}

__global__ void __launch_bounds__(256)
clip_cols_kernel(const float* __restrict__ logits, float* __restrict__ col_lse, int M, int N) {
  __shared__ float red[8];
  const int j = blockIdx.x, tid = threadIdx.x;
  float mx = -INFINITY;
  for (int i = tid; i < M; i += 256) mx = fmaxf(mx, logits[(size_t)i * N + j]);
  mx = warp_max(mx);
  if ((tid & 31) == 0) red[tid >> 5] = mx;
  __syncthreads();
  mx = red[0];
  for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  float sum = 0.f;
  for (int i = tid; i < M; i += 256) sum += expf(logits[(size_t)i * N + j] - mx);
  sum = warp_sum(sum);
  if ((tid & 31) == 0) red[tid >> 5] = sum;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    col_lse[j] = mx + logf(t);
  }
}

// ---- phase 2: G, coef, cz, loss / dtemp partials ---------------------------------------------------
__global__ void __launch_bounds__(256)
clip_grad_kernel(const float* __restrict__ logits, const float* __restrict__ row_lse, const float* __restrict__ col_lse,
                 const float* __restrict__ xn2, const float* __restrict__ zn2, const float* __restrict__ temp,
                 float scale, int diag0, float* __restrict__ coef, float* __restrict__ coef_t, float* __restrict__ cz,
                 float* __restrict__ partial, int M, int N) {
  __shared__ float red[8];
  const int i = blockIdx.x, tid = threadIdx.x;
  const float rl = row_lse[i];
  const float sx = expf(temp[0]) * rsqrtf(xn2[i]);
  float gl_sum = 0.f;
  for (int j = tid; j < N; j += 256) {
    float l = logits[(size_t)i * N + j];
    float g = 0.5f * scale * (expf(l - rl) + expf(l - col_lse[j]) - ((i == diag0 + j) ? 2.f : 0.f));
    const float cf = g * sx * rsqrtf(zn2[j]);
    coef[(size_t)i * N + j] = cf;
    if (coef_t) coef_t[(size_t)j * ((M + 3) & ~3) + i] = cf;
    float gl = g * l;
    gl_sum += gl;
    atomicAdd(cz + j, gl / zn2[j]);
  }
  gl_sum = warp_sum(gl_sum);
  if ((tid & 31) == 0) red[tid >> 5] = gl_sum;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(partial + 1, t);
    float lp = 0.f;
    const int j = i - diag0;
    if (j >= 0 && j < N) lp += scale * (0.5f * rl - logits[(size_t)i * N + j] + 0.5f * col_lse[j]);
    atomicAdd(partial, lp);
  }
}

// ---- dz[j,d] = sum_i coef[i,j] x[i,d] - cz[j] z[j,d] ----------------------------------------------
__global__ void __launch_bounds__(256)
clip_dz_kernel(const float* __restrict__ coef, const float* __restrict__ cz, const float* __restrict__ x,
               const float* __restrict__ z, float* __restrict__ dz, const float* __restrict__ gscale, int M, int N,
               int64_t D) {
  __shared__ float As[CK][CB + 4];  // coef[i][j]
  __shared__ float Bs[CK][CB + 4];  // x[i][d]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t d0 = (int64_t)blockIdx.x * CB;
  const int j0 = blockIdx.y * CB;
  const bool vec = (D & 3) == 0;
  float acc[4][4] = {};
  const int lrow = tid >> 4, lc = (tid & 15) * 4;
  for (int i0 = 0; i0 < M; i0 += CK) {
    float4 v = make_float4(0, 0, 0, 0), u = make_float4(0, 0, 0, 0);
    const int i = i0 + lrow;
    if (i < M) {
      const float* cr = coef + (size_t)i * N;
      if (j0 + lc < N) v.x = cr[j0 + lc];
      if (j0 + lc + 1 < N) v.y = cr[j0 + lc + 1];
      if (j0 + lc + 2 < N) v.z = cr[j0 + lc + 2];
      if (j0 + lc + 3 < N) v.w = cr[j0 + lc + 3];
      u = ld4_guard(x + (size_t)i * D, d0 + lc, D, vec);
    }
    *reinterpret_cast<float4*>(&As[lrow][lc]) = v;
    *reinterpret_cast<float4*>(&Bs[lrow][lc]) = u;
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < CK; ++kk) {
      float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float ar[4] = {av.x, av.y, av.z, av.w};
      const float br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(ar[a], br[b], acc[a][b]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int j = j0 + ty * 4 + a;
    if (j >= N) continue;
    const float c = cz[j];
    const float gs = gscale ? gscale[0] : 1.f;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int64_t d = d0 + tx * 4 + b;
      if (d < D) dz[(size_t)j * D + d] = gs * (acc[a][b] - c * z[(size_t)j * D + d]);
    }
  }
}

}  // namespace sd

namespace sd {
bool clip_tc_supported(int M, int N, int64_t D, const void* x, const void* z, bool bf16);
size_t clip_dots_tc_workspace(int M, int N, int64_t D);
int clip_dots_tc(const void* x, const void* z, float* dots, float* workspace, int M, int N, int64_t D, bool bf16, cudaStream_t st);
int clip_dz_tc(const void* coef_t, const float* cz, const void* x, const float* z, float* dz, const float* gscale,
               int M, int N, int64_t D, bool bf16, cudaStream_t st);
}  // namespace sd

using namespace sd;

extern "C" {

int64_t sd_clip_dots_workspace_bytes(int M, int N, int64_t D) {
  if (D % 4 != 0 || D < 64) return 0;
  return (int64_t)clip_dots_tc_workspace(M, N, D);
}

int sd_clip_dots_tc(const float* x, const float* z, float* dots, void* workspace, int M, int N, int64_t D, void* stream) {
  SD_REQUIRE(clip_tc_supported(M, N, D, x, z, false), "sd_clip_dots_tc: unsupported shape/alignment (D %% 4 != 0?)");
  SD_REQUIRE(workspace != nullptr, "sd_clip_dots_tc: workspace is null");
  return clip_dots_tc(x, z, dots, reinterpret_cast<float*>(workspace), M, N, D, false, (cudaStream_t)stream);
}

int sd_clip_dots_tc_bf16(const void* x, const void* z, float* dots, void* workspace, int M, int N, int64_t D, void* stream) {
  SD_REQUIRE(clip_tc_supported(M, N, D, x, z, true), "sd_clip_dots_tc_bf16: unsupported shape/alignment (D %% 8 != 0?)");
  SD_REQUIRE(workspace != nullptr, "sd_clip_dots_tc_bf16: workspace is null");
  return clip_dots_tc(x, z, dots, reinterpret_cast<float*>(workspace), M, N, D, true, (cudaStream_t)stream);
}

int sd_clip_dz_tc_bf16(const void* coef_t, const float* cz, const void* x, const float* z, float* dz, const float* gscale,
                       int M, int N, int64_t D, void* stream) {
  SD_REQUIRE(clip_tc_supported(M, N, D, x, z, true), "sd_clip_dz_tc_bf16: unsupported shape/alignment (D %% 8 != 0?)");
  SD_REQUIRE(coef_t != nullptr && (((uintptr_t)coef_t) & 15) == 0 && (((uintptr_t)dz) & 15) == 0, "sd_clip_dz_tc_bf16: bad pointers");
  return clip_dz_tc(coef_t, cz, x, z, dz, gscale, M, N, D, true, (cudaStream_t)stream);
}

int sd_cast_rows_bf16(const float* x, void* y, float* nrm2, int M, int64_t D, void* stream) {
  SD_REQUIRE(D % 8 == 0 && (((uintptr_t)x) & 15) == 0 && (((uintptr_t)y) & 15) == 0, "sd_cast_rows_bf16: D %% 8 != 0 or unaligned rows");
  cudaStream_t st = (cudaStream_t)stream;
  SD_CUDA(cudaMemsetAsync(nrm2, 0, sizeof(float) * M, st));
  int splits = cdiv(148 * 8, M);
  int64_t chunk = (D + splits - 1) / splits;
  chunk = (chunk + 2047) / 2048 * 2048;
  splits = cdiv(D, chunk);
  cast_rows_bf16_kernel<<<dim3(M, splits), 256, 0, st>>>(x, reinterpret_cast<__nv_bfloat16*>(y), nrm2, D, chunk);
  return check_launch("cast_rows_bf16");
}

// row_lse[i] = log sum_j exp(logit[i][j]) over the columns of ALL ranks from the per-rank (max, sum exp(l - max)) pairs
// (world == 1: just max + log(sum)); replaces the max / sum all-reduces and the eager exp / mul / log glue of loss.py:79's
// cross-entropy over global-batch negatives
__global__ void clip_merge_rows_kernel(const float* __restrict__ g, int world, int M, float* __restrict__ row_lse) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  float m = g[2 * i];
  for (int q = 1; q < world; ++q) m = fmaxf(m, g[((size_t)q * M + i) * 2]);
  float s = 0.f;
  for (int q = 0; q < world; ++q) s += g[((size_t)q * M + i) * 2 + 1] * expf(g[((size_t)q * M + i) * 2] - m);
  row_lse[i] = m + logf(s);
}

int sd_clip_merge_row_stats(const float* gathered, int world, int M, float* row_lse, void* stream) {
  SD_REQUIRE(gathered && row_lse && world > 0 && M > 0, "sd_clip_merge_row_stats: bad arguments");
  clip_merge_rows_kernel<<<cdiv(M, 256), 256, 0, (cudaStream_t)stream>>>(gathered, world, M, row_lse);
  return check_launch("clip_merge_rows");
}

int sd_rownorm2_bf16(const void* x, float* nrm2, int M, int64_t D, void* stream) {
  SD_REQUIRE(D % 8 == 0 && (((uintptr_t)x) & 15) == 0, "sd_rownorm2_bf16: D %% 8 != 0 or unaligned rows");
  cudaStream_t st = (cudaStream_t)stream;
  SD_CUDA(cudaMemsetAsync(nrm2, 0, sizeof(float) * M, st));
  int splits = cdiv(148 * 8, M);
  int64_t chunk = (D + splits - 1) / splits;
  chunk = (chunk + 2047) / 2048 * 2048;
  splits = cdiv(D, chunk);
  rownorm2_bf16_kernel<<<dim3(M, splits), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), nrm2, D, chunk);
  return check_launch("rownorm2_bf16");
}

int sd_clip_coef_t_bf16(const float* coef, void* coef_t, int M, int N, int Mp, void* stream) {
  SD_REQUIRE(Mp >= M && Mp % 8 == 0, "sd_clip_coef_t_bf16: Mp must be M rounded up to a multiple of 8");
  coef_t_bf16_kernel<<<dim3(cdiv(Mp, 32), cdiv(N, 32)), dim3(32, 8), 0, (cudaStream_t)stream>>>(
      coef, reinterpret_cast<__nv_bfloat16*>(coef_t), M, N, Mp);
  return check_launch("clip_coef_t_bf16");
}

int sd_clip_dz_tc(const float* coef_t, const float* cz, const float* x, const float* z, float* dz, const float* gscale,
                  int M, int N, int64_t D, void* stream) {
  SD_REQUIRE(clip_tc_supported(M, N, D, x, z, false), "sd_clip_dz_tc: unsupported shape/alignment (D %% 4 != 0?)");
  SD_REQUIRE(coef_t != nullptr && (((uintptr_t)dz) & 15) == 0, "sd_clip_dz_tc: bad pointers");
  return clip_dz_tc(coef_t, cz, x, z, dz, gscale, M, N, D, false, (cudaStream_t)stream);
}

int sd_rownorm2(const float* x, float* nrm2, int M, int64_t D, void* stream) {
  return rownorm2_launch(x, nrm2, M, D, 0, (cudaStream_t)stream);
}

int sd_clip_dots(const float* x, const float* z, float* dots, int M, int N, int64_t D, void* stream) {
  const int tiles = cdiv(M, CB) * cdiv(N, CB);
  int splits = cdiv(148 * 4, tiles);
  int64_t chunk = (D + splits - 1) / splits;
  chunk = (chunk + CK - 1) / CK * CK;
  splits = cdiv(D, chunk);
  clip_dots_kernel<<<dim3(cdiv(M, CB), cdiv(N, CB), splits), 256, 0, (cudaStream_t)stream>>>(x, z, dots, M, N, D, chunk);
  return check_launch("clip_dots");
}

int sd_clip_phase1(const float* dots, const float* xn2, const float* zn2, const float* temp, float* logits,
                   float* row_stat, float* col_lse, int M, int N, void* stream) {
  clip_rows_kernel<<<M, 256, 0, (cudaStream_t)stream>>>(dots, xn2, zn2, temp, logits, row_stat, N);
  if (check_launch("clip_rows")) return 1;
  clip_cols_kernel<<<N, 256, 0, (cudaStream_t)stream>>>(logits, col_lse, M, N);
  return check_launch("clip_cols");
}

int sd_clip_phase2(const float* logits, const float* row_lse, const float* col_lse, const float* xn2,
                   const float* zn2, const float* temp, float scale, int diag0, float* coef, float* coef_t, float* cz,
                   float* partial, int M, int N, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SD_CUDA(cudaMemsetAsync(cz, 0, sizeof(float) * N, st));
  SD_CUDA(cudaMemsetAsync(partial, 0, sizeof(float) * 2, st));
  clip_grad_kernel<<<M, 256, 0, st>>>(logits, row_lse, col_lse, xn2, zn2, temp, scale, diag0, coef, coef_t, cz, partial, M, N);
  return check_launch("clip_grad");
}

int sd_clip_dz(const float* coef, const float* cz, const float* x, const float* z, float* dz, const float* gscale,
               int M, int N, int64_t D, void* stream) {
  clip_dz_kernel<<<dim3(cdiv(D, CB), cdiv(N, CB)), 256, 0, (cudaStream_t)stream>>>(coef, cz, x, z, dz, gscale, M, N, D);
  return check_launch("clip_dz");
}

}  // extern "C"
