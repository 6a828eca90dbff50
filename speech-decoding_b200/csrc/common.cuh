// Shared helpers for the sd_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/sd_b200.h"

namespace sd {

// ---- error plumbing -------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define SD_REQUIRE(cond, ...)            \
  do {                                   \
    if (!(cond)) {                       \
      ::sd::set_error(__VA_ARGS__);      \
      return 1;                          \
    }                                    \
  } while (0)

#define SD_CUDA(expr)                                                                  \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      ::sd::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return 1;                                                                        \
    }                                                                                  \
  } while (0)

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Per-DEVICE once-flags: cudaFuncSetAttribute (and occupancy queries) apply to the current device only, so a process that
// drives several GPUs must repeat them on each.  `flags` is a static array of SD_MAX_DEVICES bools at the call site.
constexpr int SD_MAX_DEVICES = 64;
static inline bool first_use_on_device(bool* flags) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= SD_MAX_DEVICES) return true;     // out of table: just redo the (idempotent) set-up every time
  if (flags[dev]) return false;
  flags[dev] = true;
  return true;
}

// ---- dtype helpers --------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// 4-element vector load/store of T as floats (p must be 4-element aligned)
template <typename T> struct Vec4;
template <> struct Vec4<float> {
  static __device__ __forceinline__ float4 ld(const float* p) { return *reinterpret_cast<const float4*>(p); }
  static __device__ __forceinline__ void st(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
};
template <> struct Vec4<__nv_bfloat16> {
  static __device__ __forceinline__ float4 ld(const __nv_bfloat16* p) {
    uint2 r = *reinterpret_cast<const uint2*>(p);
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&r.x);
    __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&r.y);
    float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
  }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, float4 v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
    __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
    uint2 r;
    r.x = *reinterpret_cast<uint32_t*>(&a);
    r.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = r;
  }
};

// ---- math -----------------------------------------------------------------------------------------
// erf-form GELU, F.gelu default (models.py:158,161,194,195).  erf is evaluated with Abramowitz-Stegun
// 7.1.26 (|abs error| <= 1.5e-7, i.e. fp32-level): one MUFU.RCP + one MUFU.EX2 + 6 FMA instead of the
// ~30-instruction erff(), which made the HBM-bound BatchNorm/GELU passes compute-bound.  The exponential
// exp(-x^2/2) is shared between the CDF and the PDF in the derivative.
struct GeluParts { float cdf, pdf_over_rsqrt2pi; };
__device__ __forceinline__ GeluParts gelu_parts(float x) {
  const float u = fabsf(x) * 0.70710678118654752f;
  const float t = __frcp_rn(fmaf(0.3275911f, u, 1.0f));
  const float e = __expf(-u * u);                       // = exp(-x^2/2)
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float erf_abs = fmaf(-poly * t, e, 1.0f);       // erf(|x|/sqrt2)
  GeluParts g;
  g.cdf = 0.5f * (1.0f + copysignf(erf_abs, x));
  g.pdf_over_rsqrt2pi = e;
  return g;
}
__device__ __forceinline__ float gelu_f(float x) { return x * gelu_parts(x).cdf; }
// d/dx gelu(x) = Phi(x) + x*phi(x)
__device__ __forceinline__ float gelu_grad_f(float x) {
  const GeluParts g = gelu_parts(x);
  return fmaf(x * 0.39894228040143268f, g.pdf_over_rsqrt2pi, g.cdf);
}
// bf16-mode GELU: the tanh form evaluated with the hardware MUFU.TANH (6 instructions, 1 MUFU).  It differs
// from the erf form by < 5e-4 absolute -- below half a bf16 ulp of the stored result -- and keeps the
// HBM-bound BatchNorm/GELU passes under the ~12 instructions/element budget of a 6 TB/s stream.  The
// fp32 mode keeps the erf form above.  Forward and derivative of one mode are the same function.
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float gelu_fast(float x) {
  const float x2 = x * x;
  const float th = tanh_approx(x * fmaf(0.0356774081f, x2, 0.7978845608f));
  const float hx = 0.5f * x;
  return fmaf(hx, th, hx);
}
__device__ __forceinline__ float gelu_grad_fast(float x) {
  const float x2 = x * x;
  const float th = tanh_approx(x * fmaf(0.0356774081f, x2, 0.7978845608f));
  const float du = fmaf(0.1070322243f, x2, 0.7978845608f);       // d/dx of the tanh argument
  const float sech2 = fmaf(-th, th, 1.0f);
  return fmaf(0.5f * x * sech2, du, fmaf(0.5f, th, 0.5f));
}
// precision-mode dispatch: T = activation storage type
template <typename T> __device__ __forceinline__ float gelu_t(float x) { return gelu_f(x); }
template <> __device__ __forceinline__ float gelu_t<__nv_bfloat16>(float x) { return gelu_fast(x); }
template <typename T> __device__ __forceinline__ float gelu_grad_t(float x) { return gelu_grad_f(x); }
template <> __device__ __forceinline__ float gelu_grad_t<__nv_bfloat16>(float x) { return gelu_grad_fast(x); }

__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + __expf(-x)); }
// bf16-mode sigmoid: 0.5*tanh(x/2)+0.5 on MUFU.TANH (3 instructions; |err| < 3e-4, below bf16 resolution)
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(0.5f, tanh_approx(0.5f * x), 0.5f); }
template <typename T> __device__ __forceinline__ float sigmoid_t(float x) { return sigmoid_f(x); }
template <> __device__ __forceinline__ float sigmoid_t<__nv_bfloat16>(float x) { return sigmoid_fast(x); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// launch with programmatic stream serialization allowed (see tc_common.cuh: grid_dep_wait / grid_dep_launch);
// cluster_x > 1 adds a cluster dimension.  The attribute is opt-in: SD_B200_PDL=1.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster_x,
                              Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[2];
  int n = 0;
  if (cluster_x > 1) {
    at[n].id = cudaLaunchAttributeClusterDimension;
    at[n].val.clusterDim.x = cluster_x; at[n].val.clusterDim.y = 1; at[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = at; cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// implementation selector (api.cu)
int current_impl();
// SMs the persistent tensor-core kernels may occupy: the device's SM count, or the cap set with sd_set_sm_limit()
// (data-parallel runs leave a few SMs to the concurrent NCCL kernels: a persistent grid that does not fit is
// serialised into two waves by the first SM a collective holds).
int sm_budget();
int device_sm_count();
void set_sm_limit(int v);

// ---- per-family entry points (defined in the .cu files) ------------------------------------------
int conv_fwd_simt(const sd_conv_args& a, cudaStream_t st);
int conv_wgrad_simt(const sd_wgrad_args& a, cudaStream_t st);
int conv_fwd_tc(const sd_conv_args& a, cudaStream_t st);
int conv_wgrad_tc(const sd_wgrad_args& a, cudaStream_t st);
bool conv_fwd_tc_supported(const sd_conv_args& a);
void set_conv_pair(int on, bool ws);   // conv forward: allow 2-CTA (cta_group::2) tiles / weight-stationary pairs
bool conv_wgrad_tc_supported(const sd_wgrad_args& a);
int conv_fwd_tf32(const sd_conv_args& a, cudaStream_t st);      // conv_tf32.cu
bool conv_fwd_tf32_supported(const sd_conv_args& a);
int conv_wgrad_tf32(const sd_wgrad_args& a, cudaStream_t st);   // wgrad_tf32.cu
bool conv_wgrad_tf32_supported(const sd_wgrad_args& a);

}  // namespace sd
