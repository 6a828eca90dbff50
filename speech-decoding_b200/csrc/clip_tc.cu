// CLIPLoss streaming GEMMs on tcgen05 (kind::tf32, operands read as fp32 straight from HBM by TMA).
// Reference: speech_decoding/utils/loss.py:60-71 (similarity) and its autograd.
//
//  (1) dots[i,j] = sum_d x[i,d] z[j,d]        M,N <= a few thousand, D = F*T = 368,640: tiny MN, huge K.
//      Split-K: work item = (256-row pair of x, <=256-row tile of z, K slice); both operands K-major;
//      two 128x256 fp32 accumulators fill TMEM; the partial tile goes to a workspace and a small kernel
//      sums the K slices.  HBM-bound by design: every byte of x and z is read once.
//  (2) dz[j,d] = gs * ( sum_i coefT[j,i] x[i,d] - cz[j] z[j,d] )
//      A = coefT (K-major), B = x seen as [d, i] => MN-major (d contiguous), persistent tiles of
//      128 (j) x 128 (d) with double-buffered TMEM; the epilogue prefetches its z row segment with a
//      bulk copy, fuses the projection term and the upstream scale, and bulk-stores 512 contiguous bytes
//      per thread.
#include "tc_common.cuh"

namespace sd {

using namespace tc;

namespace {

constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = (2 + NUM_EPI_WARPS) * 32;
constexpr int TMEM_COLS = 512;
constexpr int SMEM_LIMIT = 227 * 1024;

// ======================================================================================================
// (1) similarity, split-K
// ======================================================================================================
constexpr int DK = 32;                       // 32 fp32 = 128 B per K block row (64 bf16 in the bf16 variant)
constexpr int D_STAGES = 3;
constexpr int D_A_BYTES = 2 * 128 * 128;     // two 128-row halves
constexpr int D_B_BYTES = 256 * 128;
constexpr int D_STAGE_BYTES = D_A_BYTES + D_B_BYTES;

struct DotsParams {
  float* part;   // [nsplit][M][N]
  int M, N, block_n, m_pairs, n_tiles, nsplit;
  long long kblocks_total, kblocks_per_split;
};

// BF16: both operands are bf16 rows (kind::f16, 64 elements per 128-byte K block) -- the data-parallel path,
// where the gathered speech rows travel and are stored in bf16; otherwise tf32 on fp32 rows.
template <bool BF16>
__global__ void __launch_bounds__(NUM_THREADS, 1)
clip_dots_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_z,
                    const DotsParams p) {
  constexpr int KB_ELEMS = BF16 ? 2 * DK : DK;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + D_STAGES * D_STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (D_STAGES + s); };
  const uint32_t tfull_bar = bar_base + 8u * (2 * D_STAGES);
  const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * D_STAGES + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  int item = blockIdx.x;
  const int split = item % p.nsplit; item /= p.nsplit;
  const int n_tile = item % p.n_tiles;
  const int m_pair = item / p.n_tiles;
  const long long kb0 = (long long)split * p.kblocks_per_split;
  long long kb1 = kb0 + p.kblocks_per_split;
  if (kb1 > p.kblocks_total) kb1 = p.kblocks_total;
  const int iters = (int)(kb1 > kb0 ? kb1 - kb0 : 0);

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_x);
    prefetch_tmap(&tmap_z);
    for (int s = 0; s < D_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_smem, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));

  if (warp == 0) {
    int s = 0; uint32_t ph = 0;
    const uint32_t tx = D_A_BYTES + (uint32_t)p.block_n * 128;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(empty_bar(s), ph ^ 1);
      if (elect_one_sync()) {
        const uint32_t sa = smem_base + s * D_STAGE_BYTES, sb = sa + D_A_BYTES;
        const int kc = (int)((kb0 + it) * KB_ELEMS);
        mbar_arrive_expect_tx(full_bar(s), tx);
        tma_load_3d(sa, &tmap_x, full_bar(s), kc, m_pair * 256, 0);
        tma_load_3d(sa + 128 * 128, &tmap_x, full_bar(s), kc, m_pair * 256 + 128, 0);
        tma_load_3d(sb, &tmap_z, full_bar(s), kc, n_tile * p.block_n, 0);
      }
      __syncwarp();
      if (++s == D_STAGES) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc(BF16 ? /*bf16*/ 1 : /*tf32*/ 2, 0, 0, 128, (uint32_t)p.block_n);
    const uint32_t dhi = smem_desc_hi(1024);
    int s = 0; uint32_t ph = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(full_bar(s), ph);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint32_t alo = smem_desc_lo(smem_base + s * D_STAGE_BYTES, 16);
        const uint32_t a2lo = alo + ((128 * 128) >> 4), blo = alo + (D_A_BYTES >> 4);
#pragma unroll
        for (int k = 0; k < DK / 8; ++k) {   // 32 bytes of K per MMA for both element types
          if (BF16) {
            umma_f16(tmem_base, desc64(alo + 2 * k, dhi), desc64(blo + 2 * k, dhi), idesc, (it | k) != 0);
            umma_f16(tmem_base + 256, desc64(a2lo + 2 * k, dhi), desc64(blo + 2 * k, dhi), idesc, (it | k) != 0);
          } else {
            umma_tf32(tmem_base, desc64(alo + 2 * k, dhi), desc64(blo + 2 * k, dhi), idesc, (it | k) != 0);
            umma_tf32(tmem_base + 256, desc64(a2lo + 2 * k, dhi), desc64(blo + 2 * k, dhi), idesc, (it | k) != 0);
          }
        }
        umma_commit(empty_bar(s));
        if (it == iters - 1) umma_commit(tfull_bar);
      }
      __syncwarp();
      if (++s == D_STAGES) { s = 0; ph ^= 1; }
    }
    if (iters == 0 && elect_one_sync()) umma_commit(tfull_bar);
    __syncwarp();
  } else {
    const int ew = warp - 2, quad = warp & 3, half = ew >> 2;   // half selects the 128-row accumulator
    const int i = m_pair * 256 + half * 128 + quad * 32 + lane;
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + half * 256;
    float* prow = p.part + ((size_t)split * p.M + (i < p.M ? i : 0)) * p.N + (size_t)n_tile * p.block_n;
    for (int c = 0; c < p.block_n; c += 16) {
      uint32_t r[16];
      tmem_ld16(taddr + c, r);
      tmem_ld_wait();
      if (i < p.M) {
        const int j0 = n_tile * p.block_n + c;
        if (iters == 0) {
#pragma unroll
          for (int q = 0; q < 16; ++q) r[q] = 0u;
        }
        if (j0 + 16 <= p.N && (p.N & 3) == 0) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<uint4*>(prow + c + 4 * q) = make_uint4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
        } else {
#pragma unroll
          for (int q = 0; q < 16; ++q)
            if (j0 + q < p.N) prow[c + q] = __uint_as_float(r[q]);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

__global__ void sum_partials_kernel(const float* __restrict__ part, float* __restrict__ dots, int64_t mn, int nsplit) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= mn) return;
  float s = 0.f;
  for (int k = 0; k < nsplit; ++k) s += part[(size_t)k * mn + i];
  dots[i] = s;
}

// ======================================================================================================
// (2) gradient GEMM
// ======================================================================================================
constexpr int Z_BM = 128, Z_BN = 128, Z_BK = 32;
constexpr int Z_STAGES = 4;
constexpr int Z_A_BYTES = Z_BM * 128;               // 128 rows(j) x 32 fp32(i)
constexpr int Z_ATOM = Z_BK * 128;                  // 32 K rows(i) x 32 fp32(d) = 4 KB
constexpr int Z_B_BYTES = (Z_BN / 32) * Z_ATOM;     // 4 atoms = 16 KB
constexpr int Z_STAGE_BYTES = Z_A_BYTES + Z_B_BYTES;
constexpr int Z_PITCH = Z_BN * 4 + 16;

struct DzParams {
  const float* z;
  const float* cz;
  const float* gscale;
  float* dz;
  int M, N;
  long long D;
  int j_tiles, num_tiles, k_blocks;
};

// BF16: coefT and x are bf16 (kind::f16): 64 global rows i per k-block; x as the MN-major operand uses the plain
// 128-byte swizzle (64 contiguous d per K row, 8-row atoms, one 8 KB block per 64 columns of d).  z and dz stay fp32.
template <bool BF16>
__global__ void __launch_bounds__(NUM_THREADS, 1)
clip_dz_tc_kernel(const __grid_constant__ CUtensorMap tmap_ct, const __grid_constant__ CUtensorMap tmap_x,
                  const DzParams p) {
  constexpr int KB_ROWS = BF16 ? 2 * Z_BK : Z_BK;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stg_base = smem_base + Z_STAGES * Z_STAGE_BYTES;
  const uint32_t bar_base = stg_base + Z_BM * Z_PITCH;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Z_STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * Z_STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * Z_STAGES + 2 + a); };
  auto res_bar = [&](int w) { return bar_base + 8u * (2 * Z_STAGES + 4 + w); };
  const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * Z_STAGES + 4 + NUM_EPI_WARPS);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_ct);
    prefetch_tmap(&tmap_x);
    for (int s = 0; s < Z_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), NUM_EPI_WARPS); }
    for (int w = 0; w < NUM_EPI_WARPS; ++w) mbar_init(res_bar(w), 32);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_smem, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));

  if (warp == 0) {
    int s = 0; uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int j_tile = tile % p.j_tiles;
      const long long d0 = (long long)(tile / p.j_tiles) * Z_BN;
      for (int kb = 0; kb < p.k_blocks; ++kb) {
        mbar_wait(empty_bar(s), ph ^ 1);
        if (elect_one_sync()) {
          const uint32_t sa = smem_base + s * Z_STAGE_BYTES, sb = sa + Z_A_BYTES;
          mbar_arrive_expect_tx(full_bar(s), Z_STAGE_BYTES);
          tma_load_3d(sa, &tmap_ct, full_bar(s), kb * KB_ROWS, j_tile * Z_BM, 0);
          if (BF16) {
#pragma unroll
            for (int a = 0; a < Z_BN / 64; ++a)
              tma_load_3d(sb + a * (Z_B_BYTES / 2), &tmap_x, full_bar(s), (int)(d0 + 64 * a), kb * KB_ROWS, 0);
          } else {
#pragma unroll
            for (int a = 0; a < Z_BN / 32; ++a)
              tma_load_3d(sb + a * Z_ATOM, &tmap_x, full_bar(s), (int)(d0 + 32 * a), kb * Z_BK, 0);
          }
        }
        __syncwarp();
        if (++s == Z_STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc(BF16 ? /*bf16*/ 1 : /*tf32*/ 2, /*A K-major*/ 0, /*B MN-major*/ 1, Z_BM, Z_BN);
    const uint32_t ahi = smem_desc_hi(1024);
    const uint32_t bhi = BF16 ? smem_desc_hi(1024) : smem_desc_hi(512, /*SWIZZLE_128B_BASE32B*/ 1);
    int s = 0; uint32_t ph = 0; int it_tile = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it_tile) {
      const int acc = it_tile & 1;
      const uint32_t acc_ph = (it_tile >> 1) & 1;
      mbar_wait(tempty_bar(acc), acc_ph ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * 256;
      for (int kb = 0; kb < p.k_blocks; ++kb) {
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t alo = smem_desc_lo(smem_base + s * Z_STAGE_BYTES, 16);
          const uint32_t blo = smem_desc_lo(smem_base + s * Z_STAGE_BYTES + Z_A_BYTES, BF16 ? Z_B_BYTES / 2 : Z_ATOM);
#pragma unroll
          for (int k = 0; k < Z_BK / 8; ++k) {
            // A: 32 B further along the swizzled row.  B tf32: next 8 K-rows (two 4-row swizzle groups, 1024 B);
            // B bf16: next 16 K-rows (two 8-row atoms, 2048 B)
            if (BF16) umma_f16(d_tmem, desc64(alo + 2 * k, ahi), desc64(blo + k * (2048 >> 4), bhi), idesc, (kb | k) != 0);
            else umma_tf32(d_tmem, desc64(alo + 2 * k, ahi), desc64(blo + k * (1024 >> 4), bhi), idesc, (kb | k) != 0);
          }
          umma_commit(empty_bar(s));
          if (kb == p.k_blocks - 1) umma_commit(tfull_bar(acc));
        }
        __syncwarp();
        if (++s == Z_STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else {
    const int ew = warp - 2, quad = warp & 3, hsel = ew >> 2;
    const int row = quad * 32 + lane;
    const uint32_t stg = stg_base + row * Z_PITCH + hsel * (Z_BN / 2) * 4;   // this thread's 64-column segment
    const float gs = p.gscale ? __ldg(p.gscale) : 1.f;
    uint32_t res_ph = 0;
    int it_tile = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it_tile) {
      const int acc = it_tile & 1;
      const uint32_t acc_ph = (it_tile >> 1) & 1;
      const int j = (tile % p.j_tiles) * Z_BM + row;
      const long long d0 = (long long)(tile / p.j_tiles) * Z_BN + hsel * (Z_BN / 2);
      const bool valid = j < p.N;
      long long rem = p.D - d0;
      const int cols = rem <= 0 ? 0 : (rem > Z_BN / 2 ? Z_BN / 2 : (int)rem);   // D % 4 == 0 on this path
      const uint32_t bytes = valid ? (uint32_t)cols * 4 : 0;
      bulk_wait_read0();
      if (bytes) {
        mbar_arrive_expect_tx(res_bar(ew), bytes);
        bulk_load(stg, p.z + (size_t)j * p.D + d0, bytes, res_bar(ew));
      } else {
        mbar_arrive(res_bar(ew));
      }
      mbar_wait(tfull_bar(acc), acc_ph);
      tc_fence_after();
      mbar_wait(res_bar(ew), res_ph);
      res_ph ^= 1;
      const float c = valid ? __ldg(p.cz + j) : 0.f;
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * 256 + hsel * (Z_BN / 2);
#pragma unroll
      for (int cc = 0; cc < Z_BN / 2; cc += 16) {
        uint32_t r[16];
        tmem_ld16(taddr + cc, r);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float4 zz;
          asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(zz.x), "=f"(zz.y), "=f"(zz.z), "=f"(zz.w) : "r"(stg + (cc + 4 * q) * 4));
          float4 o;
          o.x = gs * (__uint_as_float(r[4 * q]) - c * zz.x);
          o.y = gs * (__uint_as_float(r[4 * q + 1]) - c * zz.y);
          o.z = gs * (__uint_as_float(r[4 * q + 2]) - c * zz.z);
          o.w = gs * (__uint_as_float(r[4 * q + 3]) - c * zz.w);
          asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(stg + (cc + 4 * q) * 4), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w) : "memory");
        }
      }
      if (bytes) {
        fence_proxy_async();
        bulk_store(p.dz + (size_t)j * p.D + d0, stg, bytes);
      }
      bulk_commit();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
    }
    bulk_wait0();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

int sms() { return sm_budget(); }

}  // namespace

bool clip_tc_supported(int M, int N, int64_t D, const void* x, const void* z, bool bf16) {
  if (D % (bf16 ? 8 : 4) != 0 || D < 64) return false;    // TMA needs 16-byte row strides
  if (((uintptr_t)x & 15) || ((uintptr_t)z & 15)) return false;
  if (M < 1 || N < 1) return false;
  return true;
}

size_t clip_dots_tc_workspace(int M, int N, int64_t D) {
  const int block_n = N >= 256 ? 256 : (N + 15) / 16 * 16;
  const int m_pairs = (M + 255) / 256, n_tiles = (N + block_n - 1) / block_n;
  int nsplit = sms() / (m_pairs * n_tiles);
  if (nsplit < 1) nsplit = 1;
  return (size_t)nsplit * M * N * sizeof(float);
}

int clip_dots_tc(const void* x, const void* z, float* dots, float* workspace, int M, int N, int64_t D, bool bf16, cudaStream_t st) {
  DotsParams p;
  memset(&p, 0, sizeof(p));
  p.part = workspace; p.M = M; p.N = N;
  p.block_n = N >= 256 ? 256 : (N + 15) / 16 * 16;
  p.m_pairs = (M + 255) / 256;
  p.n_tiles = (N + p.block_n - 1) / p.block_n;
  p.nsplit = sms() / (p.m_pairs * p.n_tiles);
  if (p.nsplit < 1) p.nsplit = 1;
  const int kb_elems = bf16 ? 2 * DK : DK, esz = bf16 ? 2 : 4;
  const CUtensorMapDataType dt = bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  p.kblocks_total = (D + kb_elems - 1) / kb_elems;
  p.kblocks_per_split = (p.kblocks_total + p.nsplit - 1) / p.nsplit;
  CUtensorMap tx, tz;
  if (make_tmap_3d(&tx, dt, x, (uint64_t)D, (uint64_t)M, 1, (uint64_t)D * esz, (uint64_t)D * esz * M, kb_elems, 128, 1)) return 1;
  if (make_tmap_3d(&tz, dt, z, (uint64_t)D, (uint64_t)N, 1, (uint64_t)D * esz, (uint64_t)D * esz * N, kb_elems, (uint32_t)p.block_n, 1)) return 1;
  static bool attr[SD_MAX_DEVICES];
  if (first_use_on_device(attr)) {
    SD_CUDA(cudaFuncSetAttribute(clip_dots_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    SD_CUDA(cudaFuncSetAttribute(clip_dots_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
  }
  const int smem = D_STAGES * D_STAGE_BYTES + 256 + 1024;
  if (bf16) clip_dots_tc_kernel<true><<<p.m_pairs * p.n_tiles * p.nsplit, NUM_THREADS, smem, st>>>(tx, tz, p);
  else clip_dots_tc_kernel<false><<<p.m_pairs * p.n_tiles * p.nsplit, NUM_THREADS, smem, st>>>(tx, tz, p);
  if (check_launch("clip_dots_tc")) return 1;
  const int64_t mn = (int64_t)M * N;
  sum_partials_kernel<<<cdiv(mn, 256), 256, 0, st>>>(workspace, dots, mn, p.nsplit);
  return check_launch("clip_sum_partials");
}

int clip_dz_tc(const void* coef_t, const float* cz, const void* x, const float* z, float* dz, const float* gscale,
               int M, int N, int64_t D, bool bf16, cudaStream_t st) {
  DzParams p;
  memset(&p, 0, sizeof(p));
  p.z = z; p.cz = cz; p.gscale = gscale; p.dz = dz; p.M = M; p.N = N; p.D = D;
  p.j_tiles = (N + Z_BM - 1) / Z_BM;
  p.num_tiles = p.j_tiles * (int)((D + Z_BN - 1) / Z_BN);
  const int kb_rows = bf16 ? 2 * Z_BK : Z_BK;
  p.k_blocks = (M + kb_rows - 1) / kb_rows;
  CUtensorMap tc_, tx;
  if (bf16) {
    const int Mp = (M + 7) / 8 * 8;   // coefT rows are padded to a 16-byte stride by the caller
    const CUtensorMapDataType BF = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    if (make_tmap_3d(&tc_, BF, coef_t, (uint64_t)M, (uint64_t)N, 1, (uint64_t)Mp * 2, (uint64_t)Mp * 2 * N, kb_rows, Z_BM, 1)) return 1;
    if (make_tmap_3d(&tx, BF, x, (uint64_t)D, (uint64_t)M, 1, (uint64_t)D * 2, (uint64_t)D * 2 * M, 64, kb_rows, 1)) return 1;
  } else {
    const int Mp = (M + 3) / 4 * 4;
    if (make_tmap_3d(&tc_, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, coef_t, (uint64_t)M, (uint64_t)N, 1, (uint64_t)Mp * 4, (uint64_t)Mp * 4 * N, Z_BK, Z_BM, 1)) return 1;
    if (make_tmap_3d(&tx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, x, (uint64_t)D, (uint64_t)M, 1, (uint64_t)D * 4, (uint64_t)D * 4 * M, 32, Z_BK, 1, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return 1;
  }
  static bool attr[SD_MAX_DEVICES];
  if (first_use_on_device(attr)) {
    SD_CUDA(cudaFuncSetAttribute(clip_dz_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    SD_CUDA(cudaFuncSetAttribute(clip_dz_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
  }
  const int smem = Z_STAGES * Z_STAGE_BYTES + Z_BM * Z_PITCH + 256 + 1024;
  const int grid = p.num_tiles < sms() ? p.num_tiles : sms();
  if (bf16) clip_dz_tc_kernel<true><<<grid, NUM_THREADS, smem, st>>>(tc_, tx, p);
  else clip_dz_tc_kernel<false><<<grid, NUM_THREADS, smem, st>>>(tc_, tx, p);
  return check_launch("clip_dz_tc");
}

}  // namespace sd
