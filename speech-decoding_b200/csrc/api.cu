// C-ABI glue: error reporting, device query, implementation dispatch for the conv family.
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

namespace sd {

static thread_local char g_err[512] = "";
static int g_impl = SD_IMPL_AUTO;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

int current_impl() { return g_impl; }

static int g_sm_limit = 0;   // 0 = no cap
int device_sm_count() {       // of the CURRENT device (cached per device ordinal)
  static int cache[SD_MAX_DEVICES];
  int dev = 0, n = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < SD_MAX_DEVICES && cache[dev] > 0) return cache[dev];
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  if (n <= 0) n = 148;
  if (dev >= 0 && dev < SD_MAX_DEVICES) cache[dev] = n;
  return n;
}
int sm_budget() {
  const int n = device_sm_count();
  return (g_sm_limit > 0 && g_sm_limit < n) ? g_sm_limit : n;
}
void set_sm_limit(int v) { g_sm_limit = v; }

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("SD_B200_PDL");     // opt-in: measured +0.4 % on the cfg2 step (the launch gaps are ~2 %)
    on = (e && e[0] == '1') ? 1 : 0;
  }
  return on != 0;
}

// A bf16 call the tcgen05 kernels cannot take (odd GLU width, misaligned pointer ...) runs on the CUDA-core kernel, which
// is ~10x slower: never silently -- the first such call of every (entry point, shape) is reported on stderr
// (SD_B200_STRICT_TC=1 turns the report into an error).
static int warn_simt_fallback(const char* what, int N, int K, int taps, int act, int out_mode) {
  static unsigned long long seen[64];
  static int n_seen = 0;
  const unsigned long long key = ((unsigned long long)(what[8] == 'w') << 60) ^ ((unsigned long long)N << 40) ^ ((unsigned long long)K << 20) ^
                                 ((unsigned long long)taps << 8) ^ ((unsigned long long)act << 4) ^ (unsigned long long)out_mode;
  for (int i = 0; i < n_seen; ++i)
    if (seen[i] == key) return 0;
  if (n_seen < 64) seen[n_seen++] = key;
  fprintf(stderr, "sd_b200 WARNING: %s (bf16, N=%d K=%d taps=%d act=%d out=%d) is not supported by the tcgen05 kernels and runs on the "
                  "CUDA-core kernel (~10x slower)\n", what, N, K, taps, act, out_mode);
  return 0;
}
static bool strict_tc() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("SD_B200_STRICT_TC");
    on = (e && e[0] == '1') ? 1 : 0;
  }
  return on != 0;
}

}  // namespace sd

using namespace sd;

extern "C" {

const char* sd_last_error(void) { return g_err; }

int sd_abi_version(void) { return 2; }

int sd_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  SD_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  SD_CUDA(cudaGetDeviceProperties(&p, dev));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  return 0;
}

int sd_set_sm_limit(int n) {
  SD_REQUIRE(n >= 0, "sd_set_sm_limit: negative limit");
  set_sm_limit(n);
  return 0;
}

int sd_set_impl(int impl) {
  SD_REQUIRE(impl >= SD_IMPL_AUTO && impl <= SD_IMPL_TC_WS, "sd_set_impl: bad value %d", impl);
  g_impl = impl >= SD_IMPL_TC_1CTA ? SD_IMPL_TC : impl;
  if (impl != SD_IMPL_SIMT) set_conv_pair(impl == SD_IMPL_TC_1CTA ? 0 : impl == SD_IMPL_AUTO ? 1 : 2, impl == SD_IMPL_TC_WS);
  return 0;
}

int sd_conv_fwd(const sd_conv_args* a, void* stream) {
  SD_REQUIRE(a != nullptr, "sd_conv_fwd: null args");
  SD_REQUIRE(a->taps == 1 || a->taps == 3, "sd_conv_fwd: taps must be 1 or 3 (got %d)", a->taps);
  SD_REQUIRE(a->Kp % 8 == 0 && a->Np % 8 == 0, "sd_conv_fwd: padded channel counts must be multiples of 8");
  SD_REQUIRE(a->dtype == SD_F32 || a->dtype == SD_BF16 || a->dtype == SD_TF32, "sd_conv_fwd: bad dtype");
  SD_REQUIRE(a->B > 0 && a->T > 0, "sd_conv_fwd: empty input");
  const int impl = current_impl();
  if (a->dtype == SD_TF32) {   // fp32 storage, TF32 / 3xTF32 tensor-core math: no CUDA-core stand-in
    SD_REQUIRE(conv_fwd_tf32_supported(*a), "sd_conv_fwd: the TF32 tensor-core path does not support this configuration");
    return conv_fwd_tf32(*a, (cudaStream_t)stream);
  }
  SD_REQUIRE(a->in_lo == nullptr && a->w_lo == nullptr, "sd_conv_fwd: operand low planes need dtype SD_TF32");
  SD_REQUIRE(a->affine == nullptr || (impl != SD_IMPL_SIMT && conv_fwd_tc_supported(*a)),
             "sd_conv_fwd: the fused per-channel affine (eval-mode BatchNorm) needs the tensor-core path with SD_ACT_GELU and a BTC output");
  if (impl == SD_IMPL_TC) {
    SD_REQUIRE(conv_fwd_tc_supported(*a), "sd_conv_fwd: tcgen05 path does not support this configuration");
    return conv_fwd_tc(*a, (cudaStream_t)stream);
  }
  if (impl == SD_IMPL_AUTO && conv_fwd_tc_supported(*a)) return conv_fwd_tc(*a, (cudaStream_t)stream);
  if (a->dtype == SD_BF16 && impl == SD_IMPL_AUTO) {
    SD_REQUIRE(!strict_tc(), "sd_conv_fwd: bf16 configuration (N=%d K=%d taps=%d act=%d out=%d) not supported by the tcgen05 kernels "
               "(SD_B200_STRICT_TC=1)", a->N, a->K, a->taps, a->act, a->out_mode);
    warn_simt_fallback("sd_conv_fwd", a->N, a->K, a->taps, a->act, a->out_mode);
  }
  return conv_fwd_simt(*a, (cudaStream_t)stream);
}

int sd_conv_wgrad(const sd_wgrad_args* a, void* stream) {
  SD_REQUIRE(a != nullptr, "sd_conv_wgrad: null args");
  SD_REQUIRE(a->taps == 1 || a->taps == 3, "sd_conv_wgrad: taps must be 1 or 3");
  SD_REQUIRE(a->Kp % 8 == 0 && a->Np % 8 == 0, "sd_conv_wgrad: padded channel counts must be multiples of 8");
  SD_REQUIRE(a->dtype == SD_F32 || a->dtype == SD_BF16 || a->dtype == SD_TF32, "sd_conv_wgrad: bad dtype");
  const int impl = current_impl();
  if (a->dtype == SD_TF32) {
    SD_REQUIRE(conv_wgrad_tf32_supported(*a), "sd_conv_wgrad: the TF32 tensor-core path does not support this configuration");
    return conv_wgrad_tf32(*a, (cudaStream_t)stream);
  }
  SD_REQUIRE(a->dout_lo == nullptr && a->in_lo == nullptr, "sd_conv_wgrad: operand low planes need dtype SD_TF32");
  if (impl == SD_IMPL_TC) {
    SD_REQUIRE(conv_wgrad_tc_supported(*a), "sd_conv_wgrad: tcgen05 path does not support this configuration");
    return conv_wgrad_tc(*a, (cudaStream_t)stream);
  }
  if (impl == SD_IMPL_AUTO && conv_wgrad_tc_supported(*a)) return conv_wgrad_tc(*a, (cudaStream_t)stream);
  if (a->dtype == SD_BF16 && impl == SD_IMPL_AUTO) {
    SD_REQUIRE(!strict_tc(), "sd_conv_wgrad: bf16 configuration (N=%d K=%d taps=%d) not supported by the tcgen05 kernels "
               "(SD_B200_STRICT_TC=1)", a->N, a->K, a->taps);
    warn_simt_fallback("sd_conv_wgrad", a->N, a->K, a->taps, 0, 0);
  }
  return conv_wgrad_simt(*a, (cudaStream_t)stream);
}

}  // extern "C"
