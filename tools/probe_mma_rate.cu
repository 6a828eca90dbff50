// Probe: steady-state cycles per tcgen05.mma (cta_group::1, kind::f16, M=128, K=16) issued back to back on
// resident shared-memory operands (no TMA in the loop), for several N and operand majors.
#include <cstdio>
#include <vector>
#include "tc_common.cuh"
using namespace sd::tc;

__global__ void __launch_bounds__(128, 1) rate(long long* out, int N, int a_mn, int b_mn, int iters, int same_acc,
                                                const __grid_constant__ CUtensorMap tm, int tma_boxes_per_iter, int row_shift, int pat) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 32 * 1024, bar = base + 200 * 1024, tptr = bar + 16;
  const uint32_t scratch = base + 100 * 1024, tbar = bar + 64;   // 4 x 16 KB ring for the background TMA stream
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar + 32, 1); for (int i = 0; i < 4; ++i) mbar_init(tbar + 8 * i, 1); fence_barrier_init(); }
  // operands: any finite bf16 pattern
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(raw + (base - smem_u32(raw)))[i] = 0x3C003C00u;
  fence_proxy_async();
  if (warp == 0) tmem_alloc(tptr, 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  uint32_t tmem; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tptr));
  if (warp == 1) {
    const uint32_t idesc = make_idesc(1, a_mn, b_mn, 128, N);
    const uint32_t dhi = smem_desc_hi(1024);
    const uint32_t alo = smem_desc_lo(sA + (a_mn ? 0 : row_shift * 128), a_mn ? 8192 : 16), blo = smem_desc_lo(sB + (b_mn ? row_shift * 128 : 0), b_mn ? 9216 : 16);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (elect_one_sync()) {
        if (pat == 0) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t ka = a_mn ? k * 128 : k * 2, kb = b_mn ? k * 128 : k * 2;
            umma_f16(tmem + (same_acc ? 0 : (k & 1) * 256), desc64(alo + ka, dhi), desc64(blo + kb, dhi), idesc, 1);
          }
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t ka = a_mn ? k * 128 : k * 2, kb = b_mn ? k * 128 : k * 2;
#pragma unroll
            for (int j = 0; j < 3; ++j)
              umma_f16(tmem + (pat == 1 ? j * 160 : 0), desc64(alo + ka, dhi), desc64(blo + kb + j * (pat == 3 ? 0 : 32), dhi), idesc, 1);
          }
          umma_commit(bar + 32);
        }
      }
      __syncwarp();
    }
    if (elect_one_sync()) umma_commit(bar);
    __syncwarp();
    mbar_wait(bar, 0);
    long long t1 = clock64();
    if (threadIdx.x == 32) out[0] = t1 - t0;
  }
  if (warp == 2 && tma_boxes_per_iter > 0) {
    // stream 16 KB boxes (128 rows x 128 B) into the ring: tma_boxes_per_iter boxes per 4 MMAs, no consumer
    uint32_t ph = 0; int s = 0;
    const int total = iters * tma_boxes_per_iter;
    long long q0 = clock64();
    for (int i = 0; i < total; ++i) {
      if (i >= 4) { mbar_wait(tbar + 8 * s, ph); }
      if (elect_one_sync()) { mbar_arrive_expect_tx(tbar + 8 * s, 16384); tma_load_3d(scratch + s * 16384, &tm, tbar + 8 * s, 0, (i * 128) % 32768, 0); }
      __syncwarp();
      if (++s == 4) { s = 0; if (i >= 4) ph ^= 1; }
    }
    long long q1 = clock64();
    if (threadIdx.x == 64 && blockIdx.x == 0) out[1] = q1 - q0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  __nv_bfloat16* g; cudaMalloc(&g, 64 * 40000 * 2); cudaMemset(g, 0, 64 * 40000 * 2);
  CUtensorMap tm;
  make_tmap_3d(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, g, 64, 40000, 1, 128, 128 * 40000, 64, 128, 1);
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
  const int iters = 2000;
  for (int maj = 0; maj < 3; ++maj)
    for (int N : {32, 64, 128, 160, 192, 256}) {
      const int a_mn = maj == 2, b_mn = maj >= 1;
      if (N != 160 && N != 256) continue;
      if (N != 160 || maj != 2) continue;
      for (int shift : {0, 1, 2, 3}) {
        const int same = 1, boxes = 0;
        rate<<<148, 128, 210 * 1024>>>(d, N, a_mn, b_mn, iters, same, tm, boxes, 0, shift);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
        long long cc[2] = {0, 0}; cudaMemcpy(cc, d, 16, cudaMemcpyDeviceToHost);
        long long c = cc[0];
        if (boxes) printf("   TMA stream: %.1f B/clk/SM achieved (requested %.1f)\n", (double)iters * boxes * 16384 / cc[1], boxes * 16384 / (4 * 128.0 * N / 256));
        const double per = (double)c / (iters * (shift == 0 ? 4 : 12));
        printf("A %s  B %s  N=%3d  pattern %d (0: 4 MMAs/iter one acc; 1: 12 MMAs/iter over accs 0/160/320 + commit; 2: same acc; 3: same acc+same B): %.1f cycles/MMA (ideal %.0f) -> %.0f%% of peak\n", a_mn ? "MN" : "K ",
               b_mn ? "MN" : "K ", N, shift, per, 128.0 * N / 256, 100.0 * (128.0 * N / 256) / per);
      }
    }
  // all SMs busy at once (power/clock effects): 148 CTAs
  return 0;
}
