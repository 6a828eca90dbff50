// Probe: does a K-major / MN-major SWIZZLE_128B UMMA descriptor work when its start address is offset by a
// number of 128-byte rows that is NOT a multiple of 8 (tap shift inside a halo'd tile)?  Tries base_offset = 0
// and base_offset = (start >> 7) & 7.   nvcc -gencode arch=compute_100a,code=sm_100a -I ../speech-decoding_b200/csrc ...
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "tc_common.cuh"
using namespace sd::tc;

__device__ __forceinline__ uint64_t desc_bo(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t base_off) {
  uint64_t d = make_smem_desc(addr, lbo, sbo);
  d |= (uint64_t)(base_off & 7) << 49;
  return d;
}

// mode 0: K-major A rows shifted (forward conv halo). mode 1: MN-major B K-rows shifted (wgrad halo).
__global__ void __launch_bounds__(128, 1)
probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* out, int shift, int use_bo, int mode) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 64 * 1024, bar = base + 96 * 1024, bar2 = bar + 8, tptr = bar + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar2, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(tptr, 64);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  uint32_t tmem; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tptr));
  if (threadIdx.x == 0) {
    if (mode == 0) {
      mbar_arrive_expect_tx(bar, 256 * 128 + 64 * 128);
      tma_load_3d(sA, &tmA, bar, 0, 0, 0);          // A: 256 rows x 64 bf16, K-major
      tma_load_3d(sB, &tmB, bar, 0, 0, 0);          // B: 64 rows x 64 bf16, K-major
    } else {
      mbar_arrive_expect_tx(bar, 2 * 64 * 128 + 128 * 128);
      tma_load_3d(sA, &tmA, bar, 0, 0, 0);          // A (MN-major): 64 K-rows x 64 m   (two atoms -> M = 128)
      tma_load_3d(sA + 64 * 128, &tmA, bar, 64, 0, 0);
      tma_load_3d(sB, &tmB, bar, 0, 0, 0);          // B (MN-major): 128 K-rows x 64 n  (halo'd: use rows shift..shift+63)
    }
    mbar_wait(bar, 0);
    tc_fence_after();
    for (int k = 0; k < 4; ++k) {
      if (mode == 0) {
        const uint32_t a = sA + shift * 128 + k * 32, b = sB + k * 32;
        const uint32_t idesc = make_idesc(1, 0, 0, 128, 64);
        umma_f16(tmem, desc_bo(a, 16, 1024, use_bo ? (a >> 7) & 7 : 0), desc_bo(b, 16, 1024, 0), idesc, k != 0);
      } else {
        const uint32_t a = sA + k * 2048, b = sB + shift * 128 + k * 2048;
        const uint32_t idesc = make_idesc(1, 1, 1, 128, 64);
        umma_f16(tmem, desc_bo(a, 64 * 128, 1024, 0), desc_bo(b, 128 * 128, 1024, use_bo ? (b >> 7) & 7 : 0), idesc, k != 0);
      }
    }
    umma_commit(bar2);
  }
  mbar_wait(bar2, 0);
  tc_fence_after();
  for (int c = 0; c < 64; c += 16) {
    uint32_t r[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c, r);
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) out[(warp * 32 + lane) * 64 + c + i] = __uint_as_float(r[i]);
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}


int main() {
  // mode 0 data: A[256][64], B[64][64]  -> D[m][n] = sum_k A[m+shift][k] B[n][k]
  // mode 1 data: At[64 k][128 m], Bt[128 k][64 n] -> D[m][n] = sum_k At[k][m] Bt[k+shift][n]
  std::vector<__nv_bfloat16> hA(256 * 64), hB(64 * 64), hAt(64 * 128), hBt(128 * 64);
  srand(1);
  auto rnd = []() { return (float)(rand() % 17 - 8) / 8.f; };
  for (auto& v : hA) v = __float2bfloat16(rnd());
  for (auto& v : hB) v = __float2bfloat16(rnd());
  for (auto& v : hAt) v = __float2bfloat16(rnd());
  for (auto& v : hBt) v = __float2bfloat16(rnd());
  __nv_bfloat16 *dA, *dB, *dAt, *dBt; float* dO;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dAt, hAt.size() * 2); cudaMalloc(&dBt, hBt.size() * 2);
  cudaMalloc(&dO, 128 * 64 * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dAt, hAt.data(), hAt.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dBt, hBt.data(), hBt.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap mA, mB, mAt, mBt;
  make_tmap_3d(&mA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, dA, 64, 256, 1, 128, 128 * 256, 64, 256, 1);
  make_tmap_3d(&mB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, dB, 64, 64, 1, 128, 128 * 64, 64, 64, 1);
  make_tmap_3d(&mAt, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, dAt, 128, 64, 1, 256, 256 * 64, 64, 64, 1);
  make_tmap_3d(&mBt, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, dBt, 64, 128, 1, 128, 128 * 128, 64, 128, 1);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  std::vector<float> hO(128 * 64);
  for (int mode = 0; mode < 2; ++mode)
    for (int use_bo = 0; use_bo < 2; ++use_bo)
      for (int shift : {0, 1, 2, 3, 4, 5, 8, 9, 16, 17, 32}) {
        cudaMemset(dO, 0, 128 * 64 * 4);
        if (mode == 0) probe<<<1, 128, 100 * 1024>>>(mA, mB, dO, shift, use_bo, 0);
        else probe<<<1, 128, 100 * 1024>>>(mAt, mBt, dO, shift, use_bo, 1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d bo %d shift %d: CUDA error %s\n", mode, use_bo, shift, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(hO.data(), dO, 128 * 64 * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < 64; ++n) {
            double ref = 0;
            for (int k = 0; k < 64; ++k) {
              if (mode == 0) ref += (double)__bfloat162float(hA[(m + shift) * 64 + k]) * __bfloat162float(hB[n * 64 + k]);
              else ref += (double)__bfloat162float(hAt[k * 128 + m]) * __bfloat162float(hBt[(k + shift) * 64 + n]);
            }
            maxerr = fmax(maxerr, fabs(ref - hO[m * 64 + n]));
          }
        printf("mode %d (%s) base_offset=%s shift %2d : max err %.4f %s\n", mode, mode ? "MN-major B, K-row shift" : "K-major A, M-row shift",
               use_bo ? "(addr>>7)&7" : "0", shift, maxerr, maxerr < 1e-3 ? "OK" : "WRONG");
      }
  return 0;
}
