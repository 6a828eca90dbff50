// tcgen05 / TMEM / TMA Conv1d weight gradient (bf16 operands, fp32 accumulation in tensor memory).
//
//   dw[g, n, c, j] += sum_{b in group g} sum_t dy[b,t,n] * x[b, t + (j-(taps-1)/2)*dil, c]
//   dbias[n]       += sum_{b,t} dy[b,t,n]
// (autograd of nn.Conv1d, speech_decoding/models.py:97-109,128-150,188-189; the grouped form is the
//  per-subject layer models.py:98-116 with samples bucketed by subject id.)
//
// GEMM view per CTA: D[128 out-channels n, BLOCK_C in-channels c] for ONE tap j, contracted over the
// (sample, time) axis of a slice of the batch.  Both operands are read straight from the
// channels-last activations: the contraction index t is the row index, so A = dy^T and B = x^T are
// "MN-major" UMMA operands (64 contiguous channels = one 128-byte swizzle row per time step); the
// tap shift is a TMA row coordinate and rows outside [0,T) are zero-filled (= "same" padding).
// Split-K over the batch: work item = (n-tile, c-tile, tap, group, split); one item per CTA; the fp32
// tile is added to dw (PyTorch (N,K,taps) layout, via element strides) with red.global.add.f32.
// dbias comes from one extra N=16 MMA per k-step against a constant tile of ones.
#include <stdlib.h>
#include "tc_common.cuh"

namespace sd {

using namespace tc;

namespace {

constexpr int BLOCK_MN = 128;     // out-channel tile (UMMA M)
constexpr int BLOCK_T = 64;       // contraction block: 64 time steps
constexpr int ATOM_BYTES = 64 * 128;  // 64 time rows x 128 B (64 channels)
constexpr int MAX_C_ATOMS = 4;    // BLOCK_C <= 256
constexpr int STAGES = 4;
constexpr int TMEM_COLS = 512;
constexpr int BIAS_COL = 256;
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = (2 + NUM_EPI_WARPS) * 32;
constexpr int ONES_BYTES = 2048;

struct WgParams {
  float* dw;
  float* dbias;
  const int* sample_order;
  const int* group_offsets;
  int B, T, N, K, taps, dil, G;
  long long gs, sn, sk, sj;
  int block_c, c_atoms, n_tiles, c_tiles, nsplit, stage_bytes;
  float* ws;        // split-K partials [nsplit][taps][n_tiles*128][c_tiles*block_c] (+ bias [nsplit][n_tiles*128]) or NULL
  float* ws_bias;
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmap_dy, const __grid_constant__ CUtensorMap tmap_x,
                     const WgParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t ones_base = smem_base + STAGES * p.stage_bytes;
  const uint32_t bar_base = ones_base + ONES_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t tfull_bar = bar_base + 8u * (2 * STAGES);
  const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- decode the work item ----
  int item = blockIdx.x;
  const int split = item % p.nsplit; item /= p.nsplit;
  const int g = item % p.G; item /= p.G;
  const int j = item % p.taps; item /= p.taps;
  const int c_tile = item % p.c_tiles;
  const int n_tile = item / p.c_tiles;
  const int n0 = n_tile * BLOCK_MN, c0 = c_tile * p.block_c;
  const int pos0 = p.group_offsets ? p.group_offsets[g] : 0;
  const int pos1 = p.group_offsets ? p.group_offsets[g + 1] : p.B;
  const int cnt = pos1 - pos0;
  const int per = (cnt + p.nsplit - 1) / p.nsplit;
  const int s_begin = pos0 + split * per;
  const int s_end = min(pos1, s_begin + per);
  if (s_begin >= s_end) return;  // uniform for the whole CTA: nothing allocated yet
  const bool do_bias = p.dbias != nullptr && j == 0 && c_tile == 0;
  const int t_blocks = (p.T + BLOCK_T - 1) / BLOCK_T;
  const int shift = (j - (p.taps - 1) / 2) * p.dil;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_dy);
    prefetch_tmap(&tmap_x);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  {  // constant tile of bf16 ones for the bias MMA
    uint32_t* ones = reinterpret_cast<uint32_t*>(smem_gen + STAGES * p.stage_bytes);
    for (int i = threadIdx.x; i < ONES_BYTES / 4; i += NUM_THREADS) ones[i] = 0x3F803F80u;
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_smem, TMEM_COLS);
  grid_dep_launch();   // programmatic dependent launch: the prologue above overlapped the previous kernel's tail
  grid_dep_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));

  if (warp == 0) {
    // ===================== TMA producer (whole warp, one elected lane issues) =====================
    int s = 0;
    uint32_t ph = 0;
    const uint32_t stage_tx = (uint32_t)(2 + p.c_atoms) * ATOM_BYTES;
    for (int pos = s_begin; pos < s_end; ++pos) {
      const int b = p.sample_order ? __ldg(p.sample_order + pos) : pos;
      for (int tb = 0; tb < t_blocks; ++tb) {
        mbar_wait(empty_bar(s), ph ^ 1);
        if (elect_one_sync()) {
          const uint32_t sa = smem_base + s * p.stage_bytes, sb = sa + 2 * ATOM_BYTES;
          mbar_arrive_expect_tx(full_bar(s), stage_tx);
          tma_load_3d(sa, &tmap_dy, full_bar(s), n0, tb * BLOCK_T, b);
          tma_load_3d(sa + ATOM_BYTES, &tmap_dy, full_bar(s), n0 + 64, tb * BLOCK_T, b);
          for (int a = 0; a < p.c_atoms; ++a)
            tma_load_3d(sb + a * ATOM_BYTES, &tmap_x, full_bar(s), c0 + 64 * a, tb * BLOCK_T + shift, b);
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp, one elected lane issues) =====================
    const uint32_t idesc = make_idesc(/*bf16*/ 1, /*A MN-major*/ 1, /*B MN-major*/ 1, BLOCK_MN, (uint32_t)p.block_c);
    const uint32_t idesc_b = make_idesc(1, 1, 1, BLOCK_MN, 16);
    const uint32_t dhi = smem_desc_hi(1024);
    const uint32_t olo = smem_desc_lo(ones_base, ATOM_BYTES);
    int s = 0;
    uint32_t ph = 0;
    const int iters = (s_end - s_begin) * t_blocks;
    // 16-row MMA steps of the last time block that still hold rows t < T (TMA zero-fills the rest: skip them)
    const int k_last = (p.T - (t_blocks - 1) * BLOCK_T + 15) >> 4;
    int tb = 0;
    for (int it = 0; it < iters; ++it) {
      const int k_steps = tb == t_blocks - 1 ? k_last : BLOCK_T / 16;
      if (++tb == t_blocks) tb = 0;
      mbar_wait(full_bar(s), ph);
      tc_fence_after();
      if (elect_one_sync()) {
        // 16 time rows = 2048 B inside each 64-channel atom; atoms are ATOM_BYTES apart (LBO);
        // consecutive 8-row groups are 1024 B apart (SBO)
        const uint32_t alo = smem_desc_lo(smem_base + s * p.stage_bytes, ATOM_BYTES), blo = alo + ((2 * ATOM_BYTES) >> 4);
#pragma unroll
        for (int k = 0; k < BLOCK_T / 16; ++k) {
          if (k >= k_steps) break;
          umma_f16(tmem_base, desc64(alo + k * (2048 >> 4), dhi), desc64(blo + k * (2048 >> 4), dhi), idesc, (it | k) != 0);
          if (do_bias) umma_f16(tmem_base + BIAS_COL, desc64(alo + k * (2048 >> 4), dhi), desc64(olo, dhi), idesc_b, (it | k) != 0);
        }
        umma_commit(empty_bar(s));
        if (it == iters - 1) umma_commit(tfull_bar);
      }
      __syncwarp();
      if (++s == STAGES) { s = 0; ph ^= 1; }
    }
  } else {
    // ===================== epilogue: fp32 tile -> red.add into dw =====================
    const int ew = warp - 2, quad = warp & 3, hsel = ew >> 2;
    const int n = n0 + quad * 32 + lane;
    const int nch = p.block_c >> 4;
    const int ch0 = hsel ? (nch + 1) / 2 : 0, ch1 = hsel ? nch : (nch + 1) / 2;
    mbar_wait_relaxed(tfull_bar, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
    float* dwn = p.dw + (long long)g * p.gs + (long long)n * p.sn + (long long)j * p.sj;
    const int wcols = p.c_tiles * p.block_c, wrows = p.n_tiles * BLOCK_MN;
    float* wsn = p.ws ? p.ws + (((size_t)split * p.taps + j) * wrows + (n0 + quad * 32 + lane)) * wcols + c0 : nullptr;
    for (int c = ch0; c < ch1; ++c) {
      uint32_t r[16];
      tmem_ld16(taddr + c * 16, r);
      tmem_ld_wait();
      if (wsn) {      // split-K partial, summed by wgrad_reduce_kernel (no atomics on the hot tensor)
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<uint4*>(wsn + c * 16 + 4 * q) = make_uint4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
      } else if (n < p.N) {
        const int cb = c0 + c * 16;
        float* dst = dwn + cb;
        if (p.sk == 1 && cb + 16 <= p.K && (reinterpret_cast<uintptr_t>(dst) & 7) == 0) {
          // contiguous input channels (PyTorch layout of a 1x1 weight): vector reductions, 16 / 8 bytes per L2 atomic
          if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * q), "f"(__uint_as_float(r[4 * q])),
                           "f"(__uint_as_float(r[4 * q + 1])), "f"(__uint_as_float(r[4 * q + 2])), "f"(__uint_as_float(r[4 * q + 3])) : "memory");
          } else {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(dst + 2 * q), "f"(__uint_as_float(r[2 * q])),
                           "f"(__uint_as_float(r[2 * q + 1])) : "memory");
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int cc = cb + i;
            if (cc < p.K) atomicAdd(dwn + (long long)cc * p.sk, __uint_as_float(r[i]));
          }
        }
      }
    }
    if (do_bias && hsel == 0) {
      uint32_t r[16];
      tmem_ld16(taddr + BIAS_COL, r);
      tmem_ld_wait();
      if (p.ws_bias) p.ws_bias[(size_t)split * wrows + n0 + quad * 32 + lane] = __uint_as_float(r[0]);
      else if (n < p.N) atomicAdd(p.dbias + n, __uint_as_float(r[0]));
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------------
// k=3 variant: one CTA owns all three taps of its (n-tile, c-tile).  The x tile is loaded ONCE per time
// block with a halo of `dil` rows on each side ((64+2*dil) rows x 64 channels per atom) and the taps are
// UMMA descriptors whose start address is offset by j*dil rows inside that tile -- SWIZZLE_128B operands
// accept arbitrary 128-byte row offsets with base_offset 0 (profiles/r1_probe_umma_desc_row_offset.txt).
// Operand bytes per MMA drop from 40 KB to 17 KB per tap-k-block, which is what the L2->SM ingest rate
// (~45 B/clk/SM) needs for the tensor pipe to stay busy.  Accumulators: tap j at TMEM column j*160.
// ------------------------------------------------------------------------------------------------------
constexpr int W3_STAGES = 4;
constexpr int W3_TAP_COLS = 160;
constexpr int W3_BIAS_COL = 480;

struct WsMaps {
  CUtensorMap m[2];   // split-K workspace as (wcols, wrows, nsplit*taps) fp32; box (w_h, 32, 1) for the two column halves
};

struct Wg3Params {
  float* dw;
  float* dbias;
  const int* sample_order;
  const int* group_offsets;
  int B, T, N, K, dil, G;
  long long gs, sn, sk, sj;
  int block_c, c_atoms, n_tiles, c_tiles, nsplit, stage_bytes, brows, batom_bytes, ring_bytes;
  float* ws;
  float* ws_bias;
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_wgrad3_tc_kernel(const __grid_constant__ CUtensorMap tmap_dy, const __grid_constant__ CUtensorMap tmap_x,
                      const __grid_constant__ WsMaps wm, const Wg3Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t ones_base = smem_base + p.ring_bytes;
  const uint32_t bar_base = ones_base + ONES_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (W3_STAGES + s); };
  const uint32_t tfull_bar = bar_base + 8u * (2 * W3_STAGES);
  const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * W3_STAGES + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  int item = blockIdx.x;
  const int split = item % p.nsplit; item /= p.nsplit;
  const int g = item % p.G; item /= p.G;
  const int c_tile = item % p.c_tiles;
  const int n_tile = item / p.c_tiles;
  const int n0 = n_tile * BLOCK_MN, c0 = c_tile * p.block_c;
  const int pos0 = p.group_offsets ? p.group_offsets[g] : 0;
  const int pos1 = p.group_offsets ? p.group_offsets[g + 1] : p.B;
  const int cnt = pos1 - pos0;
  const int per = (cnt + p.nsplit - 1) / p.nsplit;
  const int s_begin = pos0 + split * per;
  const int s_end = min(pos1, s_begin + per);
  if (s_begin >= s_end) return;
  const bool do_bias = p.dbias != nullptr && c_tile == 0;
  const int t_blocks = (p.T + BLOCK_T - 1) / BLOCK_T;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_dy);
    prefetch_tmap(&tmap_x);
    for (int s = 0; s < W3_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  {
    uint32_t* ones = reinterpret_cast<uint32_t*>(smem_gen + p.ring_bytes);
    for (int i = threadIdx.x; i < ONES_BYTES / 4; i += NUM_THREADS) ones[i] = 0x3F803F80u;
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_smem, TMEM_COLS);
  grid_dep_launch();   // programmatic dependent launch: the prologue above overlapped the previous kernel's tail
  grid_dep_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));

  if (warp == 0) {
    int s = 0;
    uint32_t ph = 0;
    const uint32_t stage_tx = 2u * ATOM_BYTES + (uint32_t)p.c_atoms * (uint32_t)(BLOCK_T + 2 * p.dil) * 128u;
    for (int pos = s_begin; pos < s_end; ++pos) {
      const int b = p.sample_order ? __ldg(p.sample_order + pos) : pos;
      for (int tb = 0; tb < t_blocks; ++tb) {
        mbar_wait(empty_bar(s), ph ^ 1);
        if (elect_one_sync()) {
          const uint32_t sa = smem_base + s * p.stage_bytes, sb = sa + 2 * ATOM_BYTES;
          mbar_arrive_expect_tx(full_bar(s), stage_tx);
          tma_load_3d(sa, &tmap_dy, full_bar(s), n0, tb * BLOCK_T, b);
          tma_load_3d(sa + ATOM_BYTES, &tmap_dy, full_bar(s), n0 + 64, tb * BLOCK_T, b);
          for (int a = 0; a < p.c_atoms; ++a)
            tma_load_3d(sb + a * p.batom_bytes, &tmap_x, full_bar(s), c0 + 64 * a, tb * BLOCK_T - p.dil, b);
        }
        __syncwarp();
        if (++s == W3_STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc(/*bf16*/ 1, 1, 1, BLOCK_MN, (uint32_t)p.block_c);
    const uint32_t idesc_b = make_idesc(1, 1, 1, BLOCK_MN, 16);
    const uint32_t dhi = smem_desc_hi(1024);
    const uint32_t olo = smem_desc_lo(ones_base, ATOM_BYTES);
    const uint32_t tap_step = (uint32_t)(p.dil * 128) >> 4;      // descriptor units (16 B) per tap
    int s = 0;
    uint32_t ph = 0;
    const int iters = (s_end - s_begin) * t_blocks;
    // 16-row MMA steps of the last time block that still hold rows t < T (TMA zero-fills the rest: skip them)
    const int k_last = (p.T - (t_blocks - 1) * BLOCK_T + 15) >> 4;
    int tb = 0;
    for (int it = 0; it < iters; ++it) {
      const int k_steps = tb == t_blocks - 1 ? k_last : BLOCK_T / 16;
      if (++tb == t_blocks) tb = 0;
      mbar_wait(full_bar(s), ph);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint32_t alo = smem_desc_lo(smem_base + s * p.stage_bytes, ATOM_BYTES);
        const uint32_t blo = smem_desc_lo(smem_base + s * p.stage_bytes + 2 * ATOM_BYTES, (uint32_t)p.batom_bytes);
#pragma unroll
        for (int k = 0; k < BLOCK_T / 16; ++k) {
          if (k >= k_steps) break;
          const uint64_t ad = desc64(alo + k * (2048 >> 4), dhi);
#pragma unroll
          for (int j = 0; j < 3; ++j)     // tap j reads x rows t + (j-1)*dil = halo-tile rows (j*dil + ...)
            umma_f16(tmem_base + j * W3_TAP_COLS, ad, desc64(blo + j * tap_step + k * (2048 >> 4), dhi), idesc, (it | k) != 0);
          if (do_bias) umma_f16(tmem_base + W3_BIAS_COL, ad, desc64(olo, dhi), idesc_b, (it | k) != 0);
        }
        umma_commit(empty_bar(s));
        if (it == iters - 1) umma_commit(tfull_bar);
      }
      __syncwarp();
      if (++s == W3_STAGES) { s = 0; ph ^= 1; }
    }
  } else {
    const int ew = warp - 2, quad = warp & 3, hsel = ew >> 2;
    const int n = n0 + quad * 32 + lane;
    const int nch = p.block_c >> 4;
    const int ch0 = hsel ? (nch + 1) / 2 : 0, ch1 = hsel ? nch : (nch + 1) / 2;
    mbar_wait_relaxed(tfull_bar, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
    float* dwn = p.dw + (long long)g * p.gs + (long long)n * p.sn;
    const int wrows = p.n_tiles * BLOCK_MN;
    if (p.ws) {
      // split-K partial tile -> workspace through shared memory (the stage ring is free now) and one TMA
      // tensor store per warp and tap: full-line writes instead of 16-byte scattered stores
      const int seg_cols = (ch1 - ch0) * 16;
      const uint32_t seg_bytes = (uint32_t)seg_cols * 4;
      const uint32_t wbase = smem_base + (uint32_t)ew * 2u * 32u * (uint32_t)(W3_TAP_COLS / 2) * 4u;   // two ping-pong tiles per warp
#pragma unroll 1
      for (int j = 0; j < 3; ++j) {
        const uint32_t tile = wbase + (uint32_t)(j & 1) * 32u * (uint32_t)(W3_TAP_COLS / 2) * 4u;
        if (j == 2) { if (elect_one_sync()) bulk_wait_read1(); __syncwarp(); }
        for (int c = ch0; c < ch1; ++c) {
          uint32_t r[16];
          tmem_ld16(taddr + j * W3_TAP_COLS + c * 16, r);
          tmem_ld_wait();
          const uint32_t dst = tile + lane * seg_bytes + (uint32_t)(c - ch0) * 64;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst + 16 * q), "r"(r[4 * q]), "r"(r[4 * q + 1]), "r"(r[4 * q + 2]), "r"(r[4 * q + 3]) : "memory");
        }
        fence_proxy_async();
        __syncwarp();
        if (elect_one_sync()) {
          if (seg_cols > 0) tma_store_3d(&wm.m[hsel], tile, c0 + ch0 * 16, n0 + quad * 32, split * 3 + j);
          bulk_commit();
        }
      }
      if (elect_one_sync()) bulk_wait0();
      __syncwarp();
    } else if (p.sk == 3 && p.sj == 1) {
      // PyTorch's (N, K, 3) layout: the three taps of 16 consecutive input channels are 48 consecutive floats of this
      // thread's row -> vector reductions (red.global.add.v4.f32: one 16-byte L2 atomic per four values) straight into
      // dw.  The split-K partials never round-trip a workspace and no reduce kernel follows.
      for (int c = ch0; c < ch1; ++c) {
        uint32_t r0[16], r1[16], r2[16];
        tmem_ld16(taddr + c * 16, r0);
        tmem_ld16(taddr + W3_TAP_COLS + c * 16, r1);
        tmem_ld16(taddr + 2 * W3_TAP_COLS + c * 16, r2);
        tmem_ld_wait();
        const int cb = c0 + c * 16;
        if (n >= p.N || cb >= p.K) continue;
        float* dst = dwn + (long long)cb * 3;
        if (cb + 16 <= p.K) {
          float v[48];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            v[3 * i] = __uint_as_float(r0[i]); v[3 * i + 1] = __uint_as_float(r1[i]); v[3 * i + 2] = __uint_as_float(r2[i]);
          }
          if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
            for (int q = 0; q < 12; ++q)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * q), "f"(v[4 * q]), "f"(v[4 * q + 1]),
                           "f"(v[4 * q + 2]), "f"(v[4 * q + 3]) : "memory");
          } else if ((reinterpret_cast<uintptr_t>(dst) & 7) == 0) {
#pragma unroll
            for (int q = 0; q < 24; ++q)
              asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(dst + 2 * q), "f"(v[2 * q]), "f"(v[2 * q + 1]) : "memory");
          } else {
#pragma unroll
            for (int q = 0; q < 48; ++q) atomicAdd(dst + q, v[q]);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            if (cb + i < p.K) {
              atomicAdd(dst + 3 * i, __uint_as_float(r0[i]));
              atomicAdd(dst + 3 * i + 1, __uint_as_float(r1[i]));
              atomicAdd(dst + 3 * i + 2, __uint_as_float(r2[i]));
            }
          }
        }
      }
    } else {
#pragma unroll 1
      for (int j = 0; j < 3; ++j) {
        for (int c = ch0; c < ch1; ++c) {
          uint32_t r[16];
          tmem_ld16(taddr + j * W3_TAP_COLS + c * 16, r);
          tmem_ld_wait();
          if (n < p.N) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int cc = c0 + c * 16 + i;
              if (cc < p.K) atomicAdd(dwn + (long long)cc * p.sk + (long long)j * p.sj, __uint_as_float(r[i]));
            }
          }
        }
      }
    }
    if (do_bias && hsel == 0) {
      uint32_t r[16];
      tmem_ld16(taddr + W3_BIAS_COL, r);
      tmem_ld_wait();
      if (p.ws_bias) p.ws_bias[(size_t)split * wrows + n0 + quad * 32 + lane] = __uint_as_float(r[0]);
      else if (n < p.N) atomicAdd(p.dbias + n, __uint_as_float(r[0]));
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

int pick_block_c3(int kp) {   // multiple of 16, <= 160 (three accumulators must fit 480 TMEM columns)
  int best = 16, best_pad = 1 << 30;
  const int min_tiles = (kp + W3_TAP_COLS - 1) / W3_TAP_COLS;
  for (int nt = min_tiles; nt <= min_tiles + 3; ++nt) {
    int bc = ((kp + nt - 1) / nt + 15) / 16 * 16;
    if (bc > W3_TAP_COLS) continue;
    if (bc * nt < best_pad) { best_pad = bc * nt; best = bc; }
  }
  return best;
}

// dw[n,c,j] += sum_split ws[split][j][n][c];  dbias[n] += sum_split ws_bias[split][n]
__global__ void wgrad_reduce_kernel(const float* __restrict__ ws, const float* __restrict__ ws_bias, float* __restrict__ dw,
                                    float* __restrict__ dbias, int N, int K, int taps, int wrows, int wcols, int nsplit,
                                    long long sn, long long sk, long long sj) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.y, j = blockIdx.z;
  if (c < K) {
    float acc = 0.f;
#pragma unroll 8
    for (int s = 0; s < nsplit; ++s) acc += ws[(((size_t)s * taps + j) * wrows + n) * wcols + c];
    dw[(long long)n * sn + (long long)c * sk + (long long)j * sj] += acc;
  }
  if (ws_bias && j == 0 && blockIdx.x == 0 && threadIdx.x == 0) {
    float acc = 0.f;
    for (int s = 0; s < nsplit; ++s) acc += ws_bias[(size_t)s * wrows + n];
    dbias[n] += acc;
  }
}

// workspace plan for the split-K partials; returns bytes needed (0 = use atomics)
static size_t ws_plan(int nsplit, int taps, int n_tiles, int c_tiles, int block_c, bool bias, size_t* bias_off) {
  if (nsplit <= 1) return 0;
  const size_t wrows = (size_t)n_tiles * BLOCK_MN, wcols = (size_t)c_tiles * block_c;
  size_t bytes = (size_t)nsplit * taps * wrows * wcols * sizeof(float);
  *bias_off = bytes;
  if (bias) bytes += (size_t)nsplit * wrows * sizeof(float);
  return bytes;
}

int pick_block_c(int kp) {
  // multiple of 16, <= 256, minimal padding, fewest tiles on ties
  int best = 16, best_pad = 1 << 30;
  const int min_tiles = (kp + 255) / 256;
  for (int nt = min_tiles; nt <= min_tiles + 3; ++nt) {
    int bc = ((kp + nt - 1) / nt + 15) / 16 * 16;
    if (bc > 256) continue;
    if (bc * nt < best_pad) { best_pad = bc * nt; best = bc; }
  }
  return best;
}

}  // namespace

// split-K reduce, shared with wgrad_tf32.cu
int wgrad_reduce_launch(const float* ws, const float* ws_bias, float* dw, float* dbias, int N, int K, int taps, int wrows,
                        int wcols, int nsplit, long long sn, long long sk, long long sj, cudaStream_t st) {
  wgrad_reduce_kernel<<<dim3(cdiv(K, 128), N, taps), 128, 0, st>>>(ws, ws_bias, dw, dbias, N, K, taps, wrows, wcols, nsplit, sn, sk, sj);
  return check_launch("wgrad_reduce");
}

bool conv_wgrad_tc_supported(const sd_wgrad_args& a) {
  if (a.dtype != SD_BF16) return false;
  if (((uintptr_t)a.dout & 15) || ((uintptr_t)a.in & 15)) return false;
  if ((a.group_offsets == nullptr) != (a.G == 1)) return false;
  return true;
}

static int conv_wgrad3_tc(const sd_wgrad_args& a, void* ws, size_t ws_bytes, cudaStream_t st) {
  Wg3Params p;
  memset(&p, 0, sizeof(p));
  p.dw = a.dw; p.dbias = a.dbias; p.sample_order = a.sample_order; p.group_offsets = a.group_offsets;
  p.B = a.B; p.T = a.T; p.N = a.N; p.K = a.K; p.dil = a.dil; p.G = a.G;
  p.gs = a.gs; p.sn = a.sn; p.sk = a.sk; p.sj = a.sj;
  p.block_c = pick_block_c3(a.Kp);
  p.c_atoms = (p.block_c + 63) / 64;
  p.n_tiles = (a.Np + BLOCK_MN - 1) / BLOCK_MN;
  p.c_tiles = (a.Kp + p.block_c - 1) / p.block_c;
  const int base_items = p.n_tiles * p.c_tiles * a.G;
  const int sms = sm_budget();
  int nsplit = a.G > 1 ? 1 : sms / base_items;
  if (nsplit < 1) nsplit = 1;
  if (nsplit > a.B) nsplit = a.B;
  if (a.G == 1) nsplit = cdiv(a.B, cdiv(a.B, nsplit));      // drop splits that would get no sample
  p.nsplit = nsplit;
  p.brows = (BLOCK_T + 2 * a.dil + 7) / 8 * 8;
  p.batom_bytes = p.brows * 128;
  p.stage_bytes = 2 * ATOM_BYTES + p.c_atoms * p.batom_bytes;
  const int epi_bytes = NUM_EPI_WARPS * 2 * 32 * (W3_TAP_COLS / 2) * 4;     // fp32 staging tiles of the epilogue
  p.ring_bytes = W3_STAGES * p.stage_bytes > epi_bytes ? W3_STAGES * p.stage_bytes : epi_bytes;
  const int smem_bytes = p.ring_bytes + ONES_BYTES + 256 + 1024;
  CUtensorMap tdy, tx;
  if (make_tmap_3d(&tdy, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, a.dout, (uint64_t)a.Np, (uint64_t)a.T, (uint64_t)a.B,
                   (uint64_t)a.Np * 2, (uint64_t)a.T * a.Np * 2, 64, BLOCK_T, 1))
    return 1;
  if (make_tmap_3d(&tx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, a.in, (uint64_t)a.Kp, (uint64_t)a.T, (uint64_t)a.B,
                   (uint64_t)a.Kp * 2, (uint64_t)a.T * a.Kp * 2, 64, (uint32_t)(BLOCK_T + 2 * a.dil), 1))
    return 1;
  static bool attr_set[SD_MAX_DEVICES];
  if (first_use_on_device(attr_set)) {
    SD_CUDA(cudaFuncSetAttribute(conv_wgrad3_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  }
  size_t bias_off = 0;
  const size_t need = ws_plan(nsplit, 3, p.n_tiles, p.c_tiles, p.block_c, a.dbias != nullptr, &bias_off);
  // PyTorch-layout weights take the split-K partials by vector reductions straight into dw (see the kernel); the
  // workspace + reduce-kernel path remains for other layouts (SD_B200_WGRAD_WS=1 forces it: A/B measurements)
  static const bool force_ws = getenv("SD_B200_WGRAD_WS") != nullptr && getenv("SD_B200_WGRAD_WS")[0] == '1';
  const bool direct = a.sk == 3 && a.sj == 1 && !force_ws;
  const bool use_ws = need > 0 && ws != nullptr && ws_bytes >= need && a.G == 1 && !direct;
  p.ws = use_ws ? reinterpret_cast<float*>(ws) : nullptr;
  p.ws_bias = (use_ws && a.dbias) ? reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + bias_off) : nullptr;
  WsMaps wm;
  memset(&wm, 0, sizeof(wm));
  if (use_ws) {
    const int nch = p.block_c / 16;
    const int widths[2] = {(nch + 1) / 2 * 16, nch / 2 * 16};
    const uint64_t wcols = (uint64_t)p.c_tiles * p.block_c, wrows = (uint64_t)p.n_tiles * BLOCK_MN;
    for (int h = 0; h < 2; ++h) {
      if (widths[h] == 0) { wm.m[h] = wm.m[0]; continue; }
      if (make_tmap_3d(&wm.m[h], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, p.ws, wcols, wrows, (uint64_t)nsplit * 3, wcols * 4, wcols * wrows * 4,
                       (uint32_t)widths[h], 32, 1, CU_TENSOR_MAP_SWIZZLE_NONE))
        return 1;
    }
  }
  SD_CUDA(launch_pdl(conv_wgrad3_tc_kernel, dim3(base_items * nsplit), dim3(NUM_THREADS), (size_t)smem_bytes, st, 1, tdy, tx, wm, p));
  if (check_launch("conv_wgrad3_tc")) return 1;
  if (use_ws) {
    wgrad_reduce_kernel<<<dim3(cdiv(a.K, 128), a.N, 3), 128, 0, st>>>(p.ws, p.ws_bias, a.dw, a.dbias, a.N, a.K, 3, p.n_tiles * BLOCK_MN,
                                                                      p.c_tiles * p.block_c, nsplit, a.sn, a.sk, a.sj);
    return check_launch("wgrad_reduce");
  }
  return 0;
}

int conv_wgrad_tc(const sd_wgrad_args& a, cudaStream_t st) {
  void* ws = a.workspace;
  const size_t ws_bytes = (size_t)a.workspace_bytes;
  // k=3: shared halo tile for the three taps, as long as the stage ring fits in shared memory
  if (a.taps == 3) {
    const int bc = pick_block_c3(a.Kp), atoms = (bc + 63) / 64;
    const int brows = (BLOCK_T + 2 * a.dil + 7) / 8 * 8;
    int smem = W3_STAGES * (2 * ATOM_BYTES + atoms * brows * 128);
    if (smem < NUM_EPI_WARPS * 2 * 32 * (W3_TAP_COLS / 2) * 4) smem = NUM_EPI_WARPS * 2 * 32 * (W3_TAP_COLS / 2) * 4;
    smem += ONES_BYTES + 256 + 1024;
    if (smem <= 227 * 1024 && BLOCK_T + 2 * a.dil <= 256) return conv_wgrad3_tc(a, ws, ws_bytes, st);
  }
  WgParams p;
  memset(&p, 0, sizeof(p));
  p.dw = a.dw; p.dbias = a.dbias; p.sample_order = a.sample_order; p.group_offsets = a.group_offsets;
  p.B = a.B; p.T = a.T; p.N = a.N; p.K = a.K; p.taps = a.taps; p.dil = a.dil; p.G = a.G;
  p.gs = a.gs; p.sn = a.sn; p.sk = a.sk; p.sj = a.sj;
  p.block_c = pick_block_c(a.Kp);
  p.c_atoms = (p.block_c + 63) / 64;
  p.n_tiles = (a.Np + BLOCK_MN - 1) / BLOCK_MN;
  p.c_tiles = (a.Kp + p.block_c - 1) / p.block_c;
  const int base_items = p.n_tiles * p.c_tiles * a.taps * a.G;
  const int sms = sm_budget();
  int nsplit = a.G > 1 ? 1 : sms / base_items;
  if (nsplit < 1) nsplit = 1;
  if (nsplit > a.B) nsplit = a.B;
  if (a.G == 1) nsplit = cdiv(a.B, cdiv(a.B, nsplit));      // drop splits that would get no sample
  p.nsplit = nsplit;
  p.stage_bytes = (2 + p.c_atoms) * ATOM_BYTES;
  const int smem_bytes = STAGES * p.stage_bytes + ONES_BYTES + 256 + 1024;

  CUtensorMap tdy, tx;
  if (make_tmap_3d(&tdy, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, a.dout, (uint64_t)a.Np, (uint64_t)a.T, (uint64_t)a.B,
                   (uint64_t)a.Np * 2, (uint64_t)a.T * a.Np * 2, 64, BLOCK_T, 1))
    return 1;
  if (make_tmap_3d(&tx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, a.in, (uint64_t)a.Kp, (uint64_t)a.T, (uint64_t)a.B,
                   (uint64_t)a.Kp * 2, (uint64_t)a.T * a.Kp * 2, 64, BLOCK_T, 1))
    return 1;
  static bool attr_set[SD_MAX_DEVICES];
  if (first_use_on_device(attr_set)) {
    SD_CUDA(cudaFuncSetAttribute(conv_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  }
  size_t bias_off = 0;
  const size_t need = ws_plan(nsplit, a.taps, p.n_tiles, p.c_tiles, p.block_c, a.dbias != nullptr, &bias_off);
  static const bool force_ws = getenv("SD_B200_WGRAD_WS") != nullptr && getenv("SD_B200_WGRAD_WS")[0] == '1';
  const bool direct = a.sk == 1 && !force_ws;      // contiguous channels: vector reductions straight into dw, no reduce kernel
  const bool use_ws = need > 0 && ws != nullptr && ws_bytes >= need && a.G == 1 && !direct;
  p.ws = use_ws ? reinterpret_cast<float*>(ws) : nullptr;
  p.ws_bias = (use_ws && a.dbias) ? reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + bias_off) : nullptr;
  SD_CUDA(launch_pdl(conv_wgrad_tc_kernel, dim3(base_items * nsplit), dim3(NUM_THREADS), (size_t)smem_bytes, st, 1, tdy, tx, p));
  if (check_launch("conv_wgrad_tc")) return 1;
  if (use_ws) {
    wgrad_reduce_kernel<<<dim3(cdiv(a.K, 128), a.N, a.taps), 128, 0, st>>>(p.ws, p.ws_bias, a.dw, a.dbias, a.N, a.K, a.taps,
                                                                          p.n_tiles * BLOCK_MN, p.c_tiles * p.block_c, nsplit, a.sn, a.sk, a.sj);
    return check_launch("wgrad_reduce");
  }
  return 0;
}

}  // namespace sd
