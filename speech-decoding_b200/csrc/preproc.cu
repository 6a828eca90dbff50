// Batch preprocessing of the reference's Gwilliams2022Collator (dataclass/gwilliams2022.py:653-661), SURVEY 8(f)
// rank 2: per (sample, channel) row of T time samples
//   baseline_correction_single (utils/preproc_utils.py:128-142):  y = x - mean(x[:L])
//   scaleAndClamp (utils/preproc_utils.py:69-90): sklearn RobustScaler fit over time = (y - median) / IQR with numpy's
//   linear-interpolated quartiles, IQR < 10*eps -> 1, all in float64 (sklearn upcasts the torch tensor it is
//   handed), rounded to float32, clamped to +-clamp_lim.
// One warp per row: the row is sorted by a bitonic network -- in registers for T <= 512 (16 values per lane,
// shuffles only for the 15 long-distance stages), in shared memory up to T = 2048 -- the three order statistics
// are read by every lane, and the row is normalised from the original samples.  The per-element
// division is a reciprocal multiply plus one FMA correction step in float64 (correctly rounded before the
// final float32 rounding; the plain fp64 divide sequence is ~3x the instructions on a GPU whose fp64 pipe runs
// at 1/64 rate).
#include "common.cuh"

namespace sd {
namespace {

constexpr int ROWS_PER_BLOCK = 8;   // one warp per row

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- T <= 512: the row lives in registers (16 per lane, element e = lane*16 + r) -----------------------------
// Bitonic network with compile-time register indices: compare-exchange distances j < 16 stay inside a lane
// (30 of the 45 stages), j >= 16 exchange whole register files with lane ^ (j/16) by shuffle (15 stages).
template <int K, int J>
__device__ __forceinline__ void bitonic_stage16(float (&v)[16], int lane) {
  if constexpr (J >= 16) {
    const int mask = J >> 4;
    const bool upper = (lane & mask) != 0;                    // this lane holds the higher index of each pair
    const bool up = ((lane << 4) & K) == 0;                   // K >= 32 here: direction is uniform per lane
    const bool keep_max = upper == up;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const float o = __shfl_xor_sync(0xffffffffu, v[r], mask);
      v[r] = keep_max ? fmaxf(v[r], o) : fminf(v[r], o);
    }
  } else {
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      if ((r & J) == 0) {
        const float a = v[r], b = v[r | J];
        const float lo = fminf(a, b), hi = fmaxf(a, b);
        bool up;
        if constexpr (K < 16) up = (r & K) == 0;              // direction bit inside the register index: static
        else up = ((lane << 4) & K) == 0;
        v[r] = up ? lo : hi;
        v[r | J] = up ? hi : lo;
      }
    }
  }
}
template <int K, int J>
__device__ __forceinline__ void bitonic_merge16(float (&v)[16], int lane) {
  bitonic_stage16<K, J>(v, lane);
  if constexpr (J > 1) bitonic_merge16<K, J / 2>(v, lane);
}
template <int K>
__device__ __forceinline__ void bitonic_sort16(float (&v)[16], int lane) {
  if constexpr (K > 2) bitonic_sort16<K / 2>(v, lane);
  bitonic_merge16<K, K / 2>(v, lane);
}

__global__ void __launch_bounds__(ROWS_PER_BLOCK * 32)
collate_preproc_reg_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t rows, int T, int L,
                           float clamp_lim, int clamp) {
  __shared__ float stage[ROWS_PER_BLOCK][512 + 32];           // element e at e + e/16: conflict-free per-lane runs
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * ROWS_PER_BLOCK + warp;
  if (row >= rows) return;
  float* s = stage[warp];
  const float* xr = x + row * T;

  double bs = 0.0;
  for (int i = lane; i < L; i += 32) bs += (double)xr[i];
  const float base = (float)(warp_sum_d(bs) / (double)L);

  for (int i = lane; i < 512; i += 32) s[i + (i >> 4)] = i < T ? xr[i] - base : __int_as_float(0x7f800000);
  __syncwarp();
  float v[16];
#pragma unroll
  for (int r = 0; r < 16; ++r) v[r] = s[lane * 17 + r];
  bitonic_sort16<512>(v, lane);
  __syncwarp();
#pragma unroll
  for (int r = 0; r < 16; ++r) s[lane * 17 + r] = v[r];
  __syncwarp();
  auto sorted = [&](int e) { return (double)s[e + (e >> 4)]; };

  const double center = (T & 1) ? sorted(T >> 1) : (sorted((T >> 1) - 1) + sorted(T >> 1)) / 2.0;
  double q[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const double vi = (double)(T - 1) * (h ? 0.75 : 0.25);
    const int lo = (int)floor(vi);
    const double g = vi - (double)lo;
    const double a = sorted(lo), b = sorted(min(lo + 1, T - 1)), d = b - a;
    q[h] = g < 0.5 ? a + d * g : b - d * (1.0 - g);
  }
  double scale = q[1] - q[0];
  if (scale < 10.0 * 2.220446049250313e-16) scale = 1.0;
  const double rcp = 1.0 / scale;

  float* orow = out + row * T;
  for (int i = lane; i < T; i += 32) {
    const double d = (double)(xr[i] - base) - center;
    double qv = d * rcp;
    qv = fma(fma(-qv, scale, d), rcp, qv);
    float o = (float)qv;
    if (clamp) o = fminf(fmaxf(o, -clamp_lim), clamp_lim);
    orow[i] = o;
  }
}

__global__ void __launch_bounds__(ROWS_PER_BLOCK * 32)
collate_preproc_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t rows, int T, int P, int L,
                       float clamp_lim, int clamp) {
  extern __shared__ float sort_smem[];   // [ROWS_PER_BLOCK][P]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * ROWS_PER_BLOCK + warp;
  if (row >= rows) return;               // whole warp leaves together; only __syncwarp below
  float* s = sort_smem + (size_t)warp * P;
  const float* xr = x + row * T;

  // baseline: mean of the first L samples (float64 sum, rounded to float32 like torch's float mean up to 1 ulp)
  double bs = 0.0;
  for (int i = lane; i < L; i += 32) bs += (double)xr[i];
  const float base = (float)(warp_sum_d(bs) / (double)L);

  for (int i = lane; i < P; i += 32) s[i] = i < T ? xr[i] - base : __int_as_float(0x7f800000);   // +inf padding
  __syncwarp();

  // bitonic sort, ascending
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = lane; i < (P >> 1); i += 32) {
        const int lo = ((i / j) * (j << 1)) + (i % j), hi = lo + j;
        const float a = s[lo], b = s[hi];
        const bool up = (lo & k) == 0;
        if ((a > b) == up) { s[lo] = b; s[hi] = a; }
      }
      __syncwarp();
    }
  }

  // order statistics (every lane computes the same scalars)
  const double center = (T & 1) ? (double)s[T >> 1] : ((double)s[(T >> 1) - 1] + (double)s[T >> 1]) / 2.0;
  double q[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const double v = (double)(T - 1) * (h ? 0.75 : 0.25);
    const int lo = (int)floor(v);
    const double g = v - (double)lo;
    const double a = (double)s[lo], b = (double)s[min(lo + 1, T - 1)], d = b - a;
    q[h] = g < 0.5 ? a + d * g : b - d * (1.0 - g);
  }
  double scale = q[1] - q[0];
  if (scale < 10.0 * 2.220446049250313e-16) scale = 1.0;
  const double rcp = 1.0 / scale;

  float* orow = out + row * T;
  for (int i = lane; i < T; i += 32) {
    const double d = (double)(xr[i] - base) - center;
    double qv = d * rcp;
    qv = fma(fma(-qv, scale, d), rcp, qv);      // one Newton correction: d / scale correctly rounded
    float o = (float)qv;
    if (clamp) o = fminf(fmaxf(o, -clamp_lim), clamp_lim);
    orow[i] = o;
  }
}

}  // namespace
}  // namespace sd

using namespace sd;

extern "C" int sd_collate_preproc(const float* x, float* out, int64_t rows, int T, int baseline_len, float clamp_lim,
                                  int clamp, void* stream) {
  SD_REQUIRE(T >= 1 && T <= 2048, "sd_collate_preproc: T must be in [1, 2048] (got %d)", T);
  SD_REQUIRE(baseline_len >= 1 && baseline_len <= T, "sd_collate_preproc: baseline_len must be in [1, T] (got %d)", baseline_len);
  if (rows <= 0) return 0;                                         // empty batch: nothing to do
  SD_REQUIRE(x != nullptr && out != nullptr, "sd_collate_preproc: null pointer");
  int P = 2;
  while (P < T) P <<= 1;
  const size_t smem = (size_t)ROWS_PER_BLOCK * P * sizeof(float);
  static bool attr_set[SD_MAX_DEVICES];
  if (first_use_on_device(attr_set)) {
    SD_CUDA(cudaFuncSetAttribute(collate_preproc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ROWS_PER_BLOCK * 2048 * 4));
  }
  if (T <= 512)     // register-resident sort (cfg1-cfg4: T = 360)
    collate_preproc_reg_kernel<<<(unsigned)cdiv(rows, ROWS_PER_BLOCK), ROWS_PER_BLOCK * 32, 0, (cudaStream_t)stream>>>(
        x, out, rows, T, baseline_len, clamp_lim, clamp);
  else
    collate_preproc_kernel<<<(unsigned)cdiv(rows, ROWS_PER_BLOCK), ROWS_PER_BLOCK * 32, smem, (cudaStream_t)stream>>>(
        x, out, rows, T, P, baseline_len, clamp_lim, clamp);
  return check_launch("collate_preproc");
}
