// tcgen05 / TMEM / TMA implicit-GEMM Conv1d (bf16 operands, fp32 accumulation in tensor memory).
//
// Reference semantics: nn.Conv1d(k in {1,3}, padding="same", dilation) + bias (+ residual) of
// speech_decoding/models.py:97-109,128-150,156,160,188-189 with the BatchNorm batch statistics
// (models.py:158,161), GELU (models.py:194-195) and GLU (models.py:164) fused into the epilogue.
//
// GEMM view per CTA tile:  D[128 time rows, BLOCK_N channels] = sum_{tap j} sum_{k-block}
//     A_j[128 x 64] (activations, rows t0+shift_j.., channels-last => K-major, 3-D TMA box whose
//                    out-of-range rows are zero-filled: that *is* the "same" padding and it never
//                    bleeds into the neighbouring sample)
//   x W_j[BLOCK_N x 64]^T (packed weights (G,taps,Np,Kp), K-major; the group g is the subject id)
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane) + TMEM
// allocator, warps 2..9 = epilogue: two warps per TMEM lane quadrant, each thread owns one output row
// and half of the tile's columns.  smem ring of TMA stages; accumulators double-buffered in TMEM so
// the epilogue of tile i overlaps the MMAs of tile i+1.  Epilogue I/O goes through a per-warp smem
// staging tile (32 rows x the warp's column segment) moved by TMA tensor copies: the residual tile is
// prefetched while the MMAs run and the finished tile leaves as one tensor store per warp, so global
// traffic is full-sector both ways, out-of-range rows/columns are clipped by the TMA unit, and the only
// synchronisation is a __syncwarp.
// Persistent grid: one CTA per SM looping over tiles, n-tiles of the same rows adjacent in time so
// the activation slab is re-read from L2, not HBM.
#include <stdlib.h>
#include "tc_common.cuh"

namespace sd {

using namespace tc;

// make TIMELINE=1: per-role clock64 accounting printed by the first CTA (pair) -- tools only, never shipped on
#ifdef SD_TIMELINE
#define TL(...) __VA_ARGS__
#else
#define TL(...)
#endif

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int MAX_BLOCK_N = 256;
constexpr int MAX_A_SLOTS = 4, MAX_W_SLOTS = 8;
constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KB
constexpr int TMEM_COLS = 512;
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = (2 + NUM_EPI_WARPS) * 32;
constexpr int SMEM_LIMIT = 227 * 1024;

struct FwdParams {
  const float* bias;
  const float* affine;   // (2, Np) per-channel scale, shift before the activation (folded eval-mode BatchNorm) or null
  const float* bnr_ss;   // BatchNorm-backward fusion (see the epilogue): (2, Np) scale, shift of the BatchNorm whose input is bnr_y
  int bnr;
  const __nv_bfloat16* res;
  __nv_bfloat16* out_btc;
  float* out_nct;
  __nv_bfloat16* preact;
  double* stats;
  float* rownorm2;
  const int* widx;
  int B, T, N, Np, Kp, taps, dil;
  int block_n, n_tiles, m_tiles_per_sample, num_tiles, k_blocks;
  int act, out_mode, D2, Op;
  // shared-memory plan (byte offsets from the 1024-aligned base)
  int num_mpairs, sa_slots, sw_slots, a_bytes, w_bytes, a_rows, halo, off_w, w0cols, off_stg0, off_stg1, off_stgr, off_bias, off_affine, off_stats, off_bar, cols_alloc;
};

// tensor maps of the epilogue tensors, one per column-half of the tile (the halves may differ in width)
struct EpiMaps {
  CUtensorMap out[2], pre[2], res[2], preb[2];   // preb: gate half of the GLU pre-activation
  CUtensorMap yin[2];                            // BatchNorm-backward fusion: the forward pre-BN tensor y
};

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_round(float a) { return __bfloat162float(__float2bfloat16_rn(a)); }

// column sums of a 32-row x 16-column register tile held one row per lane: 16 shuffles.
// On return every lane holds the sum of column `col_of_lane(lane)` over the warp's 32 rows.
__device__ __forceinline__ float warp_colsum16(float (&v)[16], int lane) {
  {
    const bool hi = lane & 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float send = hi ? v[i] : v[i + 8];
      float keep = hi ? v[i + 8] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  {
    const bool hi = lane & 8;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float send = hi ? v[i] : v[i + 4];
      float keep = hi ? v[i + 4] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  }
  {
    const bool hi = lane & 4;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      float send = hi ? v[i] : v[i + 2];
      float keep = hi ? v[i + 2] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
  }
  {
    const bool hi = lane & 2;
    float send = hi ? v[0] : v[1];
    float keep = hi ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
  return v[0];
}
__device__ __forceinline__ int col_of_lane(int lane) {
  return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
}

// 16 floats -> 16 bf16 (32 B) into shared memory
__device__ __forceinline__ void sts16_bf16(uint32_t saddr, const float (&v)[16]) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(saddr), "r"(pack_bf16x2(v[0], v[1])),
               "r"(pack_bf16x2(v[2], v[3])), "r"(pack_bf16x2(v[4], v[5])), "r"(pack_bf16x2(v[6], v[7])) : "memory");
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(saddr + 16), "r"(pack_bf16x2(v[8], v[9])),
               "r"(pack_bf16x2(v[10], v[11])), "r"(pack_bf16x2(v[12], v[13])), "r"(pack_bf16x2(v[14], v[15])) : "memory");
}
// v += 16 bf16 read from shared memory
__device__ __forceinline__ void lds16_bf16_add(uint32_t saddr, float (&v)[16]) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    uint32_t w[4];
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(saddr + 16 * h));
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 q = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
      float2 f = __bfloat1622float2(q);
      v[8 * h + 2 * i] += f.x;
      v[8 * h + 2 * i + 1] += f.y;
    }
  }
}
__device__ __forceinline__ void lds16_f32_add(uint32_t saddr, float (&v)[16]) {
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    float4 f;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(f.x), "=f"(f.y), "=f"(f.z), "=f"(f.w) : "r"(saddr + 16 * h));
    v[4 * h] += f.x; v[4 * h + 1] += f.y; v[4 * h + 2] += f.z; v[4 * h + 3] += f.w;
  }
}
// v = v * scale + shift with 16 scales / shifts read from shared memory
__device__ __forceinline__ void lds16_f32_affine(uint32_t s_scale, uint32_t s_shift, float (&v)[16]) {
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    float4 a, b;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "r"(s_scale + 16 * h));
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "r"(s_shift + 16 * h));
    v[4 * h] = fmaf(v[4 * h], a.x, b.x); v[4 * h + 1] = fmaf(v[4 * h + 1], a.y, b.y);
    v[4 * h + 2] = fmaf(v[4 * h + 2], a.z, b.z); v[4 * h + 3] = fmaf(v[4 * h + 3], a.w, b.w);
  }
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NUM_EPI_WARPS * 32) : "memory"); }

// PAIR: the two CTAs of a cluster work on two consecutive 128-row tiles of the SAME column tile as one
// M=256 tcgen05.mma.cta_group::2: each CTA loads its own activation tile and only HALF of the weight tile
// (block_n/2 rows), which halves the dominant L2->SM operand stream.  The leader (rank 0) issues the MMAs and
// owns the full / accumulator-empty barriers; commits are multicast to both CTAs; each CTA runs its own epilogue
// on its own 128 TMEM lanes.
template <bool PAIR, bool WS, int TAPS>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                   const __grid_constant__ EpiMaps em, const FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));  // generic pointer to the aligned base
  const uint32_t bar_base = smem_base + p.off_bar;
  // barrier layout (8 B each): a_full[4], a_empty[4], w_full[8], w_empty[8], tmem_full[2], tmem_empty[2], res[8], tmem ptr
  auto afull_bar = [&](int s) { return bar_base + 8u * s; };
  auto aempty_bar = [&](int s) { return bar_base + 8u * (MAX_A_SLOTS + s); };
  auto wfull_bar = [&](int s) { return bar_base + 8u * (2 * MAX_A_SLOTS + s); };
  auto wempty_bar = [&](int s) { return bar_base + 8u * (2 * MAX_A_SLOTS + MAX_W_SLOTS + s); };
  constexpr int NB = 2 * MAX_A_SLOTS + 2 * MAX_W_SLOTS;
  auto tfull_bar = [&](int a) { return bar_base + 8u * (NB + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (NB + 2 + a); };
  auto res_bar = [&](int w) { return bar_base + 8u * (NB + 4 + w); };
  const uint32_t tmem_ptr_smem = bar_base + 8u * (NB + 4 + NUM_EPI_WARPS);
  auto y_bar = [&](int w) { return bar_base + 8u * (NB + 4 + NUM_EPI_WARPS + 1 + w); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool glu = p.act == SD_ACT_GLU;
  const int half_n = p.block_n >> 1;
  const int rank = PAIR ? (int)cluster_ctarank() : 0;
  // Tile walk.  Streaming: tile = (row tile [pair], column tile), column tile fastest, strided over the CTAs
  // [pairs].  Weight-stationary (ws): the pair owns column tile `n_fixed` for the whole kernel and `tile` walks the
  // row-tile pairs, strided over the pairs that own the same column tile.
  constexpr bool ws = WS;
  static_assert(PAIR || !WS, "weight-stationary tiles are CTA-pair tiles");
  int tile_begin = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  int tile_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  int tile_end = p.num_tiles, n_fixed = 0;
  if (ws) {
    n_fixed = tile_begin % p.n_tiles;
    tile_begin = tile_begin / p.n_tiles;
    tile_step = (tile_step - n_fixed + p.n_tiles - 1) / p.n_tiles;
    tile_end = p.num_mpairs;
  }

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_w);
    for (int s = 0; s < p.sa_slots; ++s) { mbar_init(afull_bar(s), 1); mbar_init(aempty_bar(s), 1); }
    if (ws) {   // one single-use "resident weights of k-block kb have landed" barrier per k-block
      for (int s = 0; s < p.k_blocks; ++s) mbar_init(wfull_bar(s), 1);
    } else {
      for (int s = 0; s < p.sw_slots; ++s) { mbar_init(wfull_bar(s), 1); mbar_init(wempty_bar(s), 1); }
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), PAIR ? 2 * NUM_EPI_WARPS : NUM_EPI_WARPS);
    }
    for (int w = 0; w < NUM_EPI_WARPS; ++w) { mbar_init(res_bar(w), 1); mbar_init(y_bar(w), 1); }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (PAIR) tmem_alloc_pair(tmem_ptr_smem, TMEM_COLS);
    else tmem_alloc(tmem_ptr_smem, TMEM_COLS);
  }
  // programmatic dependent launch: everything above overlapped the tail of the previous kernel in the stream;
  // from here on we read what it wrote (activations, BatchNorm scale/shift) -- and let our successor start its prologue
  grid_dep_launch();
  grid_dep_wait();
  // bias (zero beyond N) and statistics accumulators in shared memory
  {
    float* s_bias = reinterpret_cast<float*>(smem_gen + p.off_bias);
    float* s_stats = reinterpret_cast<float*>(smem_gen + p.off_stats);
    if (!glu) {
      for (int i = threadIdx.x; i < p.cols_alloc; i += NUM_THREADS) s_bias[i] = (p.bias && i < p.N) ? p.bias[i] : 0.f;
    } else {  // [0,cols) = value-half bias, [cols, 2*cols) = gate-half bias
      for (int i = threadIdx.x; i < p.cols_alloc; i += NUM_THREADS) {
        s_bias[i] = (p.bias && i < p.D2) ? p.bias[i] : 0.f;
        s_bias[p.cols_alloc + i] = (p.bias && i < p.D2) ? p.bias[p.D2 + i] : 0.f;
      }
    }
    if (p.stats)
      for (int i = threadIdx.x; i < 8 * p.cols_alloc; i += NUM_THREADS) s_stats[i] = 0.f;   // [quadrant][sum, sumsq][col]
    if (p.affine || p.bnr) {   // scale (1 beyond Np) and shift (0 beyond Np)
      const float* src = p.affine ? p.affine : p.bnr_ss;
      float* s_aff = reinterpret_cast<float*>(smem_gen + p.off_affine);
      for (int i = threadIdx.x; i < p.cols_alloc; i += NUM_THREADS) {
        s_aff[i] = i < p.Np ? src[i] : 1.f;
        s_aff[p.cols_alloc + i] = i < p.Np ? src[p.Np + i] : 0.f;
      }
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();   // the peer's barriers must be initialised before any remote arrive / multicast commit
  else __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));


  // Two operand rings.  A ring: one activation tile per k-block, (128 + 2*halo) time rows x 64 channels, loaded
  // ONCE for all taps (halo = dilation for k=3): tap j is the same tile read from row j*dil on -- SWIZZLE_128B
  // descriptors take any 128-byte row offset (profiles/r1_probe_umma_desc_row_offset.txt).  W ring: one
  // BLOCK_N x 64 weight tile per (k-block, tap).  This cuts the L2->SM operand bytes per k-block from
  // 3*(16+W) KB to (16+4*halo/16)+3*W KB; the kernel is bound by that ingest rate (~62 B/clk/SM measured).
  if (warp == 0) {
    // ===================== TMA producer (whole warp, one elected lane issues) =====================
    int sa = 0, sw = 0;
    uint32_t pha = 0, phw = 0;
    const uint32_t a_tx = (uint32_t)p.a_rows * 128u, w_tx = (uint32_t)p.block_n * BLOCK_K * 2;   // per pair if PAIR
    TL(long long tl_t0 = clock64(), tl_ae = 0, tl_we = 0, tl_q;)
    for (int tile = tile_begin; tile < tile_end; tile += tile_step) {
      const int m_tile = ws ? tile : tile / p.n_tiles, n_idx = ws ? n_fixed : tile % p.n_tiles;
      const int m_idx = PAIR ? 2 * m_tile + rank : m_tile;
      const int b = m_idx / p.m_tiles_per_sample;   // PAIR: may be == B for the odd tile out (TMA zero-fills it)
      const int t0 = (m_idx % p.m_tiles_per_sample) * BLOCK_M;
      const int g = p.widx ? __ldg(p.widx + b) : 0;  // (never PAIR: the pair shares one weight tile)
      const int row0 = glu ? n_idx * half_n : n_idx * p.block_n;
      const int row1 = glu ? p.D2 + n_idx * half_n : row0 + half_n;
      for (int kb = 0; kb < p.k_blocks; ++kb) {
        TL(tl_q = clock64();)
        mbar_wait(aempty_bar(sa), pha ^ 1);
        TL(tl_ae += clock64() - tl_q;)
        if (elect_one_sync()) {
          if (PAIR) {
            if (rank == 0) mbar_arrive_expect_tx(afull_bar(sa), 2 * a_tx);
            tma_load_3d_pair(smem_base + sa * p.a_bytes, &tmap_a, mapa_cluster(afull_bar(sa), 0), kb * BLOCK_K, t0 - p.halo, b);
          } else {
            mbar_arrive_expect_tx(afull_bar(sa), a_tx);
            tma_load_3d(smem_base + sa * p.a_bytes, &tmap_a, afull_bar(sa), kb * BLOCK_K, t0 - p.halo, b);
          }
        }
        __syncwarp();
        if (++sa == p.sa_slots) { sa = 0; pha ^= 1; }
        if (ws) {   // resident weights: loaded once, while the first row tile streams in
          if (tile == tile_begin && elect_one_sync()) {
            if (rank == 0) mbar_arrive_expect_tx(wfull_bar(kb), (uint32_t)TAPS * w_tx);
            for (int j = 0; j < TAPS; ++j)
              tma_load_3d_pair(smem_base + p.off_w + (kb * TAPS + j) * p.w_bytes, &tmap_w, mapa_cluster(wfull_bar(kb), 0),
                               kb * BLOCK_K, rank ? row1 : row0, j);
          }
          __syncwarp();
          continue;
        }
        for (int j = 0; j < TAPS; ++j) {
          TL(tl_q = clock64();)
          mbar_wait(wempty_bar(sw), phw ^ 1);
          TL(tl_we += clock64() - tl_q;)
          if (elect_one_sync()) {
            const uint32_t sb = smem_base + p.off_w + sw * p.w_bytes;
            if (PAIR) {   // this CTA's half of the rows only
              if (rank == 0) mbar_arrive_expect_tx(wfull_bar(sw), w_tx);
              tma_load_3d_pair(sb, &tmap_w, mapa_cluster(wfull_bar(sw), 0), kb * BLOCK_K, rank ? row1 : row0, g * TAPS + j);
            } else {
              mbar_arrive_expect_tx(wfull_bar(sw), w_tx);
              tma_load_3d(sb, &tmap_w, wfull_bar(sw), kb * BLOCK_K, row0, g * TAPS + j);
              tma_load_3d(sb + half_n * (BLOCK_K * 2), &tmap_w, wfull_bar(sw), kb * BLOCK_K, row1, g * TAPS + j);
            }
          }
          __syncwarp();
          if (++sw == p.sw_slots) { sw = 0; phw ^= 1; }
        }
      }
    }
    TL(if (lane == 0 && blockIdx.x < 2) printf("blk %d producer: total %lld | wait a-empty %lld w-empty %lld\n", (int)blockIdx.x,
                                               clock64() - tl_t0, tl_ae, tl_we);)
  } else if (warp == 1 && rank == 0) {
    // ===================== MMA issuer (whole warp, one elected lane issues) =====================
    const uint32_t idesc = make_idesc(/*bf16*/ 1, 0, 0, PAIR ? 2 * BLOCK_M : BLOCK_M, (uint32_t)p.block_n);
    const uint32_t dhi = smem_desc_hi(1024);
    const uint32_t tap_step = (uint32_t)(p.halo * 128) >> 4;   // descriptor units per tap (taps==3: halo == dil)
    int sa = 0, sw = 0;
    uint32_t pha = 0, phw = 0;
    int it_tile = 0;
    TL(long long tl_t0 = clock64(), tl_te = 0, tl_af = 0, tl_wf = 0, tl_q;)
    for (int tile = tile_begin; tile < tile_end; tile += tile_step, ++it_tile) {
      const int acc = it_tile & 1;
      const uint32_t acc_ph = (it_tile >> 1) & 1;
      TL(tl_q = clock64();)
      mbar_wait(tempty_bar(acc), acc_ph ^ 1);
      tc_fence_after();
      TL(tl_te += clock64() - tl_q;)
      const uint32_t d_tmem = tmem_base + acc * MAX_BLOCK_N;
      for (int kb = 0; kb < p.k_blocks; ++kb) {
        TL(tl_q = clock64();)
        mbar_wait(afull_bar(sa), pha);
        TL(tl_af += clock64() - tl_q;)
        const uint32_t alo = smem_desc_lo(smem_base + sa * p.a_bytes, 16);
        if constexpr (WS) {
          // resident weights: one elected section issues all TAPS x 4 MMAs of the k-block back to back
          if (it_tile == 0) mbar_wait(wfull_bar(kb), 0);
          tc_fence_after();
          if (elect_one_sync()) {
            const uint32_t blo0 = smem_desc_lo(smem_base + p.off_w + kb * TAPS * p.w_bytes, 16);
            const uint32_t w_step = (uint32_t)p.w_bytes >> 4;
#pragma unroll
            for (int j = 0; j < TAPS; ++j) {
#pragma unroll
              for (int k = 0; k < BLOCK_K / 16; ++k)
                umma_f16_pair(d_tmem, desc64(alo + j * tap_step + 2 * k, dhi), desc64(blo0 + j * w_step + 2 * k, dhi), idesc,
                              (kb | j | k) != 0);
            }
            umma_commit_pair(aempty_bar(sa));
            if (kb == p.k_blocks - 1) umma_commit_pair(tfull_bar(acc));
          }
          __syncwarp();
          if (++sa == p.sa_slots) { sa = 0; pha ^= 1; }
          continue;
        }
        for (int j = 0; j < TAPS; ++j) {
          TL(tl_q = clock64();)
          mbar_wait(wfull_bar(sw), phw);
          tc_fence_after();
          TL(tl_wf += clock64() - tl_q;)
          if (elect_one_sync()) {
            const uint32_t blo = smem_desc_lo(smem_base + p.off_w + sw * p.w_bytes, 16);
            const uint32_t aj = alo + j * tap_step;
#pragma unroll
            for (int k = 0; k < BLOCK_K / 16; ++k) {  // +32 B per 16-element k-step inside the swizzled row
              if (PAIR) umma_f16_pair(d_tmem, desc64(aj + 2 * k, dhi), desc64(blo + 2 * k, dhi), idesc, (kb | j | k) != 0);
              else umma_f16(d_tmem, desc64(aj + 2 * k, dhi), desc64(blo + 2 * k, dhi), idesc, (kb | j | k) != 0);
            }
            if (PAIR) {
              umma_commit_pair(wempty_bar(sw));
              if (j == TAPS - 1) {
                umma_commit_pair(aempty_bar(sa));
                if (kb == p.k_blocks - 1) umma_commit_pair(tfull_bar(acc));
              }
            } else {
              umma_commit(wempty_bar(sw));
              if (j == TAPS - 1) {
                umma_commit(aempty_bar(sa));
                if (kb == p.k_blocks - 1) umma_commit(tfull_bar(acc));
              }
            }
          }
          __syncwarp();
          if (++sw == p.sw_slots) { sw = 0; phw ^= 1; }
        }
        if (++sa == p.sa_slots) { sa = 0; pha ^= 1; }
      }
    }
    TL(if (lane == 0 && blockIdx.x < 2) printf("blk %d mma: tiles %d total %lld | wait acc-empty %lld a-full %lld w-full %lld | issue %lld\n", (int)blockIdx.x,
                                               it_tile, clock64() - tl_t0, tl_te, tl_af, tl_wf, clock64() - tl_t0 - tl_te - tl_af - tl_wf);)
  } else if (warp >= 2) {
    // ===================== epilogue (warps 2..9) =====================
    const int ew = warp - 2;
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const int hsel = ew >> 2;   // which half of the tile's column chunks this warp owns
    const int row = quad * 32 + lane;
    const uint32_t s_bias = smem_base + p.off_bias;
    const uint32_t s_aff = smem_base + p.off_affine;
    float* s_stats = reinterpret_cast<float*>(smem_gen + p.off_stats);
    // chunk range [ch0, ch1) of 16-column chunks owned by this warp
    const int nch = glu ? (half_n >> 4) : (p.block_n >> 4);
    const int ch0 = hsel ? (nch + 1) / 2 : 0;
    const int ch1 = hsel ? nch : (nch + 1) / 2;
    const int seg_col = ch0 * 16, seg_cols = (ch1 - ch0) * 16;
    const uint32_t seg_bytes = (uint32_t)seg_cols * 2;          // one staged row of this warp's segment
    const uint32_t box_bytes = 32u * seg_bytes;
    // staging tiles ([32 rows][seg_cols] bf16, dense) of this warp
    //   non-GLU: stg0 = conv output / pre-activation, stg1 = activated output (GELU with BTC output)
    //   GLU:     stg0 = value half then gate half of the pre-activation, stg1 = gated output
    const uint32_t quad0 = smem_base + p.off_stg0 + quad * (32 * p.block_n * 2);
    const uint32_t stg0 = quad0 + (glu ? 2 : 1) * (hsel ? 32 * p.w0cols * 2 : 0);
    const uint32_t stg0b = stg0 + box_bytes;                                     // GLU gate half
    const uint32_t stg1 = smem_base + p.off_stg1 + quad * (32 * (glu ? half_n : p.block_n) * 2) + (hsel ? 32 * p.w0cols * 2 : 0);
    // residual tile of this warp: its own buffer, so that the NEXT tile's residual can be fetched a whole tile ahead
    const uint32_t stgr = smem_base + p.off_stgr + quad * (32 * p.block_n * 2) + (hsel ? 32 * p.w0cols * 2 : 0);
    const uint32_t my0 = stg0 + lane * seg_bytes, my0b = stg0b + lane * seg_bytes, my1 = stg1 + lane * seg_bytes;
    const uint32_t myr = stgr + lane * seg_bytes;
    const CUtensorMap* m_out = &em.out[hsel];
    const CUtensorMap* m_pre = &em.pre[hsel];
    const CUtensorMap* m_res = &em.res[hsel];
    const CUtensorMap* m_preb = &em.preb[hsel];
    const CUtensorMap* m_yin = &em.yin[hsel];
    uint32_t res_ph = 0, y_ph = 0;
    TL(long long tl_t0 = clock64(), tl_rd = 0, tl_tf = 0, tl_rs = 0, tl_ch = 0, tl_st = 0, tl_q, tl_r;)

    // fetch this warp's residual tile of tile `tl` (non-GLU only)
    auto fetch_residual = [&](const int tl) {
      const int m_tile = ws ? tl : tl / p.n_tiles, n_idx = ws ? n_fixed : tl % p.n_tiles;
      const int m_idx = PAIR ? 2 * m_tile + rank : m_tile;
      const int b = m_idx / p.m_tiles_per_sample;
      const int t_w = (m_idx % p.m_tiles_per_sample) * BLOCK_M + quad * 32;
      const int c0 = n_idx * p.block_n + seg_col;
      if (elect_one_sync()) {
        if (seg_cols > 0 && t_w < p.T && b < p.B && c0 < p.Np) {
          mbar_arrive_expect_tx(res_bar(ew), box_bytes);
          tma_load_3d(stgr, m_res, res_bar(ew), c0, t_w, b);
        } else {
          mbar_arrive(res_bar(ew));
        }
      }
      __syncwarp();
    };
    if (p.res && tile_begin < tile_end) fetch_residual(tile_begin);
    // BatchNorm-backward fusion: this warp's tile of the forward pre-BN tensor y, fetched INTO THE OUTPUT STAGING TILE (the
    // epilogue reads a y chunk, computes g and writes g back to the same place, so no extra shared memory is needed);
    // issued as soon as the previous tile's tensor store has finished reading the staging tile, i.e. early in this
    // tile's mainloop
    auto fetch_y = [&](const int tl) {
      const int m_tile = ws ? tl : tl / p.n_tiles, n_idx = ws ? n_fixed : tl % p.n_tiles;
      const int m_idx = PAIR ? 2 * m_tile + rank : m_tile;
      const int b = m_idx / p.m_tiles_per_sample;
      const int t_w = (m_idx % p.m_tiles_per_sample) * BLOCK_M + quad * 32;
      const int c0 = n_idx * p.block_n + seg_col;
      if (elect_one_sync()) {
        if (seg_cols > 0 && t_w < p.T && b < p.B && c0 < p.Np) {
          mbar_arrive_expect_tx(y_bar(ew), box_bytes);
          tma_load_3d(stg0, m_yin, y_bar(ew), c0, t_w, b);
        } else {
          mbar_arrive(y_bar(ew));
        }
      }
      __syncwarp();
    };

    int it_tile = 0;
    for (int tile = tile_begin; tile < tile_end; tile += tile_step, ++it_tile) {
      const int acc = it_tile & 1;
      const uint32_t acc_ph = (it_tile >> 1) & 1;
      const int m_tile = ws ? tile : tile / p.n_tiles, n_idx = ws ? n_fixed : tile % p.n_tiles;
      const int m_idx = PAIR ? 2 * m_tile + rank : m_tile;
      const int b = m_idx / p.m_tiles_per_sample;
      const int t_w = (m_idx % p.m_tiles_per_sample) * BLOCK_M + quad * 32;   // first row of this warp
      const int t = t_w + lane;
      const bool valid = t < p.T && b < p.B;
      const int n0 = glu ? n_idx * half_n : n_idx * p.block_n;  // first output channel of the tile
      const int lim = glu ? p.Op : p.Np;
      // does this warp's tile intersect the tensor at all?  (TMA clips partial overlap)
      const bool live = seg_cols > 0 && t_w < p.T && b < p.B && n0 + seg_col < lim;

      // the previous tile's tensor stores must have finished reading the staging tiles
      TL(tl_q = clock64();)
      if (elect_one_sync()) bulk_wait_read0();
      __syncwarp();
      if (p.bnr) fetch_y(tile);
      TL(tl_r = clock64(); tl_rd += tl_r - tl_q; tl_q = tl_r;)
      mbar_wait_relaxed(tfull_bar(acc), acc_ph);
      tc_fence_after();
      TL(tl_r = clock64(); tl_tf += tl_r - tl_q; tl_q = tl_r;)
      if (p.res) {
        mbar_wait(res_bar(ew), res_ph);
        res_ph ^= 1;
      }
      if (p.bnr) {
        mbar_wait(y_bar(ew), y_ph);
        y_ph ^= 1;
      }
      TL(tl_r = clock64(); tl_rs += tl_r - tl_q; tl_q = tl_r;)
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * MAX_BLOCK_N;
      float sumsq = 0.f;
      float* nct_row = p.out_nct + (size_t)b * p.N * p.T + (valid ? t : 0);

      if (!glu) {
        for (int c = ch0; c < ch1; ++c) {
          const int cc = c * 16, nb = n0 + cc;
          const uint32_t so = (uint32_t)(c - ch0) * 32;
          uint32_t r[16];
          tmem_ld16(taddr + cc, r);
          tmem_ld_wait();
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
          lds16_f32_add(s_bias + nb * 4, v);
          if (p.res && live) lds16_bf16_add(myr + so, v);
          if (p.affine) lds16_f32_affine(s_aff + nb * 4, s_aff + (p.cols_alloc + nb) * 4, v);
          if (p.bnr) {
            // v = du (gradient w.r.t. u = gelu(bn(y))) for this thread's row and 16 channels.  g = du * gelu'(scale*y +
            // shift), rounded to bf16 as it will be stored; the BatchNorm-backward sums  sum g, sum g*y  over the rows are
            // taken here (the stand-alone reduce pass over du and y disappears), g replaces y in the staging tile
            float yv[16], gy[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              yv[i] = 0.f;
              if (!valid) v[i] = 0.f;       // rows beyond T: their taps reach valid rows, but they are not part of the tensor
            }
            if (live) lds16_bf16_add(my0 + so, yv);
            float xh[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) xh[i] = yv[i];
            lds16_f32_affine(s_aff + nb * 4, s_aff + (p.cols_alloc + nb) * 4, xh);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              v[i] = bf16_round(v[i] * gelu_grad_fast(xh[i]));
              gy[i] = v[i] * yv[i];
            }
            sts16_bf16(my0 + so, v);
            float gs[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) gs[i] = v[i];
            const float cs = warp_colsum16(gs, lane), cq = warp_colsum16(gy, lane);
            if ((lane & 1) == 0 && live) {       // this (quadrant, column) accumulator belongs to this warp: plain read-modify-write
              const int col = nb + col_of_lane(lane);
              s_stats[quad * 2 * p.cols_alloc + col] += cs;
              s_stats[(quad * 2 + 1) * p.cols_alloc + col] += cq;
            }
            continue;
          }
          if (p.act == SD_ACT_GELU) {
            if (p.preact) sts16_bf16(my0 + so, v);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = gelu_fast(v[i]);   // bf16 mode
            if (p.out_mode == SD_OUT_BTC) sts16_bf16(my1 + so, v);
          } else {
            sts16_bf16(my0 + so, v);
          }
          if (p.out_mode == SD_OUT_NCT_F32 && valid) {
            float* dst = nct_row + (size_t)nb * p.T;        // Z[b, nb.., t]: lanes = consecutive t => 128-byte stores
            if (nb + 16 <= p.N) {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                *dst = v[i];
                dst += p.T;
                sumsq = fmaf(v[i], v[i], sumsq);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                if (nb + i < p.N) {
                  dst[(size_t)i * p.T] = v[i];
                  sumsq = fmaf(v[i], v[i], sumsq);
                }
              }
            }
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (p.res && tile + tile_step < tile_end) fetch_residual(tile + tile_step);   // a whole tile ahead
        TL(tl_r = clock64(); tl_ch += tl_r - tl_q; tl_q = tl_r;)
        if (p.stats && !p.bnr && live) {
          // BatchNorm batch statistics of the values exactly as stored (bf16): column sums over this warp's
          // staged tile -- lane = column pair, conflict-free 4-byte reads down the rows
          const int rows_valid = min(32, p.T - t_w);
          for (int cp = lane; cp < (seg_cols >> 1); cp += 32) {
            float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
            const uint32_t col = stg0 + cp * 4;
            if (rows_valid == 32) {   // 8 independent loads in flight, two accumulator chains
              float s2 = 0.f, s3 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
              for (int r0 = 0; r0 < 32; r0 += 8) {
                uint32_t w[8];
#pragma unroll
                for (int i = 0; i < 8; ++i)
                  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w[i]) : "r"(col + (r0 + i) * seg_bytes));
#pragma unroll
                for (int i = 0; i < 8; i += 2) {
                  float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
                  float2 g = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i + 1]));
                  s0 += f.x; s1 += f.y; s2 += g.x; s3 += g.y;
                  q0 = fmaf(f.x, f.x, q0); q1 = fmaf(f.y, f.y, q1); q2 = fmaf(g.x, g.x, q2); q3 = fmaf(g.y, g.y, q3);
                }
              }
              s0 += s2; s1 += s3; q0 += q2; q1 += q3;
            } else {
              for (int r = 0; r < rows_valid; ++r) {
                uint32_t w;
                asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w) : "r"(col + r * seg_bytes));
                float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w));
                s0 += f.x; s1 += f.y;
                q0 = fmaf(f.x, f.x, q0); q1 = fmaf(f.y, f.y, q1);
              }
            }
            // this (quadrant, column pair) accumulator is touched by this lane only: plain read-modify-write
            // (shared-memory float atomics are CAS spin loops and the four quadrants would collide on them)
            float2* acc_s = reinterpret_cast<float2*>(s_stats + quad * 2 * p.cols_alloc + n0 + seg_col + 2 * cp);
            float2* acc_q = reinterpret_cast<float2*>(s_stats + (quad * 2 + 1) * p.cols_alloc + n0 + seg_col + 2 * cp);
            float2 a = *acc_s, q = *acc_q;
            a.x += s0; a.y += s1; q.x += q0; q.y += q1;
            *acc_s = a; *acc_q = q;
          }
        }
        if (elect_one_sync()) {
          if (live) {
            if (p.act == SD_ACT_GELU) {
              if (p.preact) tma_store_3d(m_pre, stg0, n0 + seg_col, t_w, b);
              if (p.out_mode == SD_OUT_BTC) tma_store_3d(m_out, stg1, n0 + seg_col, t_w, b);
            } else {
              tma_store_3d(m_out, stg0, n0 + seg_col, t_w, b);
            }
          }
          bulk_commit();
        }
      } else {
        // GLU: tile columns [0,half) = value channels n0.., [half, 2*half) = gate channels D2+n0..
        for (int c = ch0; c < ch1; ++c) {
          const int cc = c * 16, cb = n0 + cc;
          const uint32_t so = (uint32_t)(c - ch0) * 32;
          uint32_t ra[16], rb[16];
          tmem_ld16(taddr + cc, ra);
          tmem_ld16(taddr + half_n + cc, rb);
          tmem_ld_wait();
          float va[16], vb[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            va[i] = __uint_as_float(ra[i]);
            vb[i] = __uint_as_float(rb[i]);
          }
          lds16_f32_add(s_bias + cb * 4, va);
          lds16_f32_add(s_bias + (p.cols_alloc + cb) * 4, vb);
          if (p.preact) {
            sts16_bf16(my0 + so, va);
            sts16_bf16(my0b + so, vb);
          }
          if (cb + 16 <= p.D2) {
#pragma unroll
            for (int i = 0; i < 16; ++i) va[i] *= sigmoid_fast(vb[i]);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) va[i] = (cb + i < p.D2) ? va[i] * sigmoid_fast(vb[i]) : 0.f;
          }
          sts16_bf16(my1 + so, va);
        }
        fence_proxy_async();
        __syncwarp();
        if (elect_one_sync()) {
          if (live) {
            if (p.preact && n0 + seg_col < p.D2) {   // each half is its own D2-wide tensor: clipped at D2
              tma_store_3d(m_pre, stg0, n0 + seg_col, t_w, b);
              tma_store_3d(m_preb, stg0b, n0 + seg_col, t_w, b);
            }
            tma_store_3d(m_out, stg1, n0 + seg_col, t_w, b);
          }
          bulk_commit();
        }
      }
      if (p.rownorm2) {
        sumsq = warp_sum(sumsq);
        if (lane == 0 && b < p.B) atomicAdd(p.rownorm2 + b, sumsq);
      }
      tc_fence_before();
      __syncwarp();
      TL(tl_r = clock64(); tl_st += tl_r - tl_q; tl_q = tl_r;)
      if (lane == 0) {
        if (PAIR) mbar_arrive_cluster(mapa_cluster(tempty_bar(acc), 0));
        else mbar_arrive(tempty_bar(acc));
      }
    }
    if (elect_one_sync()) bulk_wait0();
    __syncwarp();
    TL(if (lane == 0 && blockIdx.x < 2 && (ew == 0 || ew == 7))
         printf("blk %d epi warp %d: tiles %d total %lld | wait store-read %lld acc-full %lld residual %lld | chunks %lld stats+store %lld (non-GLU)\n",
                (int)blockIdx.x, ew, it_tile, clock64() - tl_t0, tl_rd, tl_tf, tl_rs, tl_ch, tl_st);)
    if (p.stats) {
      epi_bar_sync();
      const int et = threadIdx.x - 64;
      for (int i = et; i < p.n_tiles * p.block_n && i < p.Np; i += NUM_EPI_WARPS * 32) {
        float a = 0.f, q = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          a += s_stats[w * 2 * p.cols_alloc + i];
          q += s_stats[(w * 2 + 1) * p.cols_alloc + i];
        }
        if (a != 0.f || q != 0.f) {
          atomicAdd(p.stats + i, (double)a);
          atomicAdd(p.stats + p.Np + i, (double)q);
        }
      }
    }
  }

  tc_fence_before();
  if (PAIR) cluster_sync_all();   // neither CTA may exit (or free TMEM) while the other can still touch it
  else __syncthreads();
  if (warp == 1) {
    if (PAIR) tmem_dealloc_pair(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// pick the N tile: multiple of `gran`, <= 256, minimal padded total, fewest tiles on ties
int pick_block_n(int n_total, int gran) {
  int best_bn = gran, best_pad = 1 << 30;
  const int min_tiles = (n_total + MAX_BLOCK_N - 1) / MAX_BLOCK_N;
  for (int nt = min_tiles; nt <= min_tiles + 3; ++nt) {
    int bn = ((n_total + nt - 1) / nt + gran - 1) / gran * gran;
    if (bn > MAX_BLOCK_N) continue;
    int pad = bn * nt;
    if (pad < best_pad) { best_pad = pad; best_bn = bn; }
  }
  return best_bn;
}

CUtensorMap ta_dummy() {
  CUtensorMap m;
  memset(&m, 0, sizeof(m));
  return m;
}

int g_conv_ws = 0;   // weight-stationary pairs: measured slower than streaming pairs at the cfg2 shapes (narrower MMAs)
int g_conv_pair = -1;   // -1: not decided yet (SD_B200_CONV_PAIR=0 disables the CTA-pair tiles)

int sm_count() { return sm_budget(); }

}  // namespace

namespace tc {

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tmap_3d(CUtensorMap* m, CUtensorMapDataType dt, const void* base, uint64_t d0, uint64_t d1, uint64_t d2,
                 uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t box0, uint32_t box1, uint32_t box2,
                 CUtensorMapSwizzle swizzle) {
  EncodeTiledFn fn = encode_fn();
  SD_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {box0, box1, box2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, dt, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SD_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d): dims %llu,%llu,%llu strides %llu,%llu box %u,%u,%u",
             (int)r, (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2,
             (unsigned long long)stride1_bytes, (unsigned long long)stride2_bytes, box0, box1, box2);
  return 0;
}

}  // namespace tc

void set_conv_pair(int on, bool ws) {   // on: 0 never, 1 where it pays (default), 2 wherever legal (tests)
  g_conv_pair = on;
  g_conv_ws = ws ? 1 : 0;
}

bool conv_fwd_tc_supported(const sd_conv_args& a) {
  if (a.dtype != SD_BF16) return false;
  if (a.act == SD_ACT_GLU && ((a.N / 2) % 8 != 0 || a.out_mode != SD_OUT_BTC || a.N % 2)) return false;
  if (a.stats && (a.act != SD_ACT_NONE || a.out_mode != SD_OUT_BTC)) return false;   // (forward BN statistics, or with bnr_y the BN-backward sums)
  if (a.rownorm2 && a.out_mode != SD_OUT_NCT_F32) return false;
  if (a.res && ((a.act != SD_ACT_NONE && a.act != SD_ACT_GELU) || a.out_mode != SD_OUT_BTC)) return false;
  if (a.affine && (a.act != SD_ACT_GELU || a.out_mode != SD_OUT_BTC || a.preact || a.stats)) return false;
  if (a.out_mode == SD_OUT_NCT_F32 && a.act != SD_ACT_GELU) return false;
  if (a.bnr_y && (a.act != SD_ACT_NONE || a.out_mode != SD_OUT_BTC || a.preact || a.affine || a.bias || !a.stats || !a.bnr_ss ||
                  ((uintptr_t)a.bnr_y & 15)))
    return false;
  if (((uintptr_t)a.in & 15) || ((uintptr_t)a.w & 15) || ((uintptr_t)a.out & 15) || ((uintptr_t)a.res & 15) ||
      ((uintptr_t)a.preact & 15))
    return false;
  if (a.Np > 4096) return false;
  return true;
}

int conv_fwd_tc(const sd_conv_args& a, cudaStream_t st) {
  const bool glu = a.act == SD_ACT_GLU;
  FwdParams p;
  memset(&p, 0, sizeof(p));
  p.bias = a.bias;
  p.affine = a.affine;
  p.bnr = a.bnr_y != nullptr;
  p.bnr_ss = a.bnr_ss;
  p.res = reinterpret_cast<const __nv_bfloat16*>(a.res);
  p.out_btc = reinterpret_cast<__nv_bfloat16*>(a.out);
  p.out_nct = reinterpret_cast<float*>(a.out);
  p.preact = reinterpret_cast<__nv_bfloat16*>(a.preact);
  p.stats = a.stats;
  p.rownorm2 = a.rownorm2;
  p.widx = a.widx;
  p.B = a.B; p.T = a.T; p.N = a.N; p.Np = a.Np; p.Kp = a.Kp; p.taps = a.taps; p.dil = a.dil;
  p.act = a.act; p.out_mode = a.out_mode;
  p.D2 = glu ? a.N / 2 : 0;
  p.Op = glu ? (p.D2 + 7) / 8 * 8 : 0;
  p.m_tiles_per_sample = (a.T + BLOCK_M - 1) / BLOCK_M;
  p.k_blocks = (a.Kp + BLOCK_K - 1) / BLOCK_K;
  p.halo = a.taps == 3 ? a.dil : 0;
  p.a_rows = BLOCK_M + 2 * p.halo;
  p.a_bytes = (p.a_rows * 128 + 1023) / 1024 * 1024;
  SD_REQUIRE(p.a_rows <= 256, "conv_fwd_tc: dilation %d too large for one activation tile", a.dil);
  if (g_conv_pair < 0) {
    const char* e = getenv("SD_B200_CONV_PAIR");
    g_conv_pair = (e && e[0] == '0') ? 0 : 1;
    e = getenv("SD_B200_CONV_WS");
    g_conv_ws = (e && e[0] == '1') ? 1 : 0;
  }
  // CTA pairs (M=256 MMAs, half the weight bytes per SM) whenever one weight tile serves both row tiles
  // (measured at the cfg2 shapes, tools/bench_conv_pair.py: pairs win for the 3-tap convs, which are bound by the
  //  weight stream, and lose for the short-K 1x1 convs, whose epilogues bound the tile time and get coupled)
  const bool pair = g_conv_pair && a.widx == nullptr && a.B * p.m_tiles_per_sample >= 2 * sm_count() &&
                    (g_conv_pair == 2 || a.taps == 3 || a.out_mode == SD_OUT_NCT_F32);
  const int n_pairs = sm_count() / 2;
  int w0 = 0, w1 = 0, smem_bytes = 0;

  // shared-memory plan for column tile `bn`.  mode 0: one CTA per tile, weights streamed through a ring;
  // 1: CTA pair, weights streamed; 2: CTA pair, weight-stationary (the pair keeps ALL taps x k-blocks of its
  // column tile resident and walks down the row tiles, so only activations stream from L2).
  auto plan = [&](int bn, int mode) -> bool {
    p.block_n = bn;
    if (glu) {
      p.n_tiles = (p.Op + bn / 2 - 1) / (bn / 2);
      p.cols_alloc = p.n_tiles * (bn / 2);
    } else {
      p.n_tiles = (a.Np + bn - 1) / bn;
      p.cols_alloc = p.n_tiles * bn;
    }
    p.w_bytes = (mode >= 1 ? bn / 2 : bn) * BLOCK_K * 2;
    const int nch = (glu ? bn / 2 : bn) / 16;
    w0 = (nch + 1) / 2 * 16; w1 = nch / 2 * 16;            // column widths of the two warp halves
    p.w0cols = w0;
    const bool need_stg1 = glu || (a.act == SD_ACT_GELU && a.out_mode == SD_OUT_BTC);
    const int stg_bytes = (a.act == SD_ACT_GELU && a.preact == nullptr) ? 0 : BLOCK_M * bn * 2;   // GELU without a saved pre-activation stages only its output
    const int stg1_bytes = need_stg1 ? BLOCK_M * (glu ? bn / 2 : bn) * 2 : 0;
    const int stgr_bytes = a.res ? BLOCK_M * bn * 2 : 0;
    const int stats_bytes = a.stats ? 8 * p.cols_alloc * 4 : 0;   // per TMEM quadrant: sums and sums of squares
    const int affine_bytes = (a.affine || a.bnr_y) ? 2 * p.cols_alloc * 4 : 0;
    const int tail = stg_bytes + stg1_bytes + stgr_bytes + (glu ? 2 : 1) * p.cols_alloc * 4 + stats_bytes + affine_bytes + 16 + 512;
    const int ring = SMEM_LIMIT - 1024 - tail;
    int sa, sw, w_region;
    if (mode == 2) {
      if (p.k_blocks > 2 * MAX_W_SLOTS || p.n_tiles > n_pairs) return false;
      w_region = a.taps * p.k_blocks * p.w_bytes;
      sa = (ring - w_region) / p.a_bytes;
      if (sa > MAX_A_SLOTS) sa = MAX_A_SLOTS;
      if (sa < 2) return false;
      sw = 0;
    } else {
      // taps==3: few activation slots (each feeds 3 weight tiles), the rest of the ring holds weight tiles
      sa = a.taps == 3 ? 3 : (ring / (p.a_bytes + p.w_bytes));
      if (a.taps == 3 && ring - sa * p.a_bytes < 5 * p.w_bytes) sa = 2;
      if (sa > MAX_A_SLOTS) sa = MAX_A_SLOTS;
      sw = (ring - sa * p.a_bytes) / p.w_bytes;
      if (sw > MAX_W_SLOTS) sw = MAX_W_SLOTS;
      if (a.taps == 3 && sw > 3 * sa + 3) sw = 3 * sa + 3;
      if (sa < 1 || sw < 2) return false;
      w_region = sw * p.w_bytes;
    }
    p.sa_slots = sa; p.sw_slots = sw;
    p.off_w = sa * p.a_bytes;
    int off = p.off_w + w_region;
    p.off_stg0 = off; off += stg_bytes;
    p.off_stg1 = off; off += stg1_bytes;
    p.off_stgr = off; off += stgr_bytes;
    p.off_bias = off; off += (glu ? 2 : 1) * p.cols_alloc * 4;
    p.off_stats = off; off += stats_bytes;
    p.off_affine = off; off += affine_bytes;
    off = (off + 15) / 16 * 16;
    p.off_bar = off; off += 512;
    smem_bytes = off + 1024;
    return smem_bytes <= SMEM_LIMIT;
  };
  int mode = pair ? 1 : 0;
  const int n_total = glu ? 2 * p.Op : a.Np, gran = glu ? 32 : 16;
  if (pair && g_conv_ws) {   // widest column tile whose weights fit next to >= 2 activation slots
    const int min_tiles = (n_total + MAX_BLOCK_N - 1) / MAX_BLOCK_N;
    for (int nt = min_tiles; nt <= min_tiles + 8 && mode != 2; ++nt) {
      const int bn = ((n_total + nt - 1) / nt + gran - 1) / gran * gran;
      if (bn <= MAX_BLOCK_N && plan(bn, 2)) mode = 2;
    }
  }
  if (mode != 2)
    SD_REQUIRE(plan(pick_block_n(n_total, gran), mode), "conv_fwd_tc: operand rings do not fit (N=%d K=%d dil=%d)", a.N, a.Kp, a.dil);
  p.num_mpairs = (a.B * p.m_tiles_per_sample + 1) / 2;
  static const bool show_plan = getenv("SD_B200_SHOW_PLAN") != nullptr;
  if (show_plan)
    fprintf(stderr, "conv_fwd_tc plan: K=%d N=%d taps=%d dil=%d act=%d out=%d res=%d stats=%d | %s block_n=%d n_tiles=%d a_slots=%d w_slots=%d smem=%d\n",
            a.Kp, a.N, a.taps, a.dil, a.act, a.out_mode, a.res != nullptr, a.stats != nullptr,
            mode == 2 ? "pair weight-stationary" : mode == 1 ? "pair streaming" : "single-CTA streaming", p.block_n, p.n_tiles,
            p.sa_slots, p.sw_slots, smem_bytes);
  p.num_tiles = pair ? p.num_mpairs * p.n_tiles : a.B * p.m_tiles_per_sample * p.n_tiles;

  // tensor maps of the epilogue tensors: (cols, T, B) with a (w, 32, 1) box, no swizzle
  EpiMaps em;
  memset(&em, 0, sizeof(em));
  const int widths[2] = {w0, w1};
  for (int h = 0; h < 2; ++h) {
    const uint32_t w = (uint32_t)widths[h];
    if (w == 0) { em.out[h] = em.out[0]; em.pre[h] = em.pre[0]; em.res[h] = em.res[0]; em.preb[h] = em.preb[0]; em.yin[h] = em.yin[0]; continue; }
    const CUtensorMapDataType BF = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    const CUtensorMapSwizzle NS = CU_TENSOR_MAP_SWIZZLE_NONE;
    if (a.out_mode == SD_OUT_BTC) {
      const uint64_t oc = glu ? (uint64_t)p.Op : (uint64_t)a.Np;
      if (make_tmap_3d(&em.out[h], BF, a.out, oc, (uint64_t)a.T, (uint64_t)a.B, oc * 2, (uint64_t)a.T * oc * 2, w, 32, 1, NS)) return 1;
    } else {
      em.out[h] = ta_dummy();
    }
    if (a.preact) {
      if (glu) {
        const uint64_t d2 = (uint64_t)p.D2;
        if (make_tmap_3d(&em.pre[h], BF, a.preact, d2, (uint64_t)a.T, (uint64_t)a.B, (uint64_t)a.Np * 2, (uint64_t)a.T * a.Np * 2, w, 32, 1, NS)) return 1;
        if (make_tmap_3d(&em.preb[h], BF, reinterpret_cast<const __nv_bfloat16*>(a.preact) + p.D2, d2, (uint64_t)a.T, (uint64_t)a.B,
                         (uint64_t)a.Np * 2, (uint64_t)a.T * a.Np * 2, w, 32, 1, NS)) return 1;
      } else {
        if (make_tmap_3d(&em.pre[h], BF, a.preact, (uint64_t)a.Np, (uint64_t)a.T, (uint64_t)a.B, (uint64_t)a.Np * 2, (uint64_t)a.T * a.Np * 2, w, 32, 1, NS)) return 1;
        em.preb[h] = em.pre[h];
      }
    } else {
      em.pre[h] = ta_dummy(); em.preb[h] = ta_dummy();
    }
    if (a.res) {
      if (make_tmap_3d(&em.res[h], BF, a.res, (uint64_t)a.Np, (uint64_t)a.T, (uint64_t)a.B, (uint64_t)a.Np * 2, (uint64_t)a.T * a.Np * 2, w, 32, 1, NS)) return 1;
    } else {
      em.res[h] = ta_dummy();
    }
    if (a.bnr_y) {
      if (make_tmap_3d(&em.yin[h], BF, a.bnr_y, (uint64_t)a.Np, (uint64_t)a.T, (uint64_t)a.B, (uint64_t)a.Np * 2, (uint64_t)a.T * a.Np * 2, w, 32, 1, NS)) return 1;
    } else {
      em.yin[h] = ta_dummy();
    }
  }

  CUtensorMap ta, tw;
  if (make_tmap_3d(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, a.in, (uint64_t)a.Kp, (uint64_t)a.T, (uint64_t)a.B,
                   (uint64_t)a.Kp * 2, (uint64_t)a.T * a.Kp * 2, BLOCK_K, (uint32_t)p.a_rows, 1))
    return 1;
  if (make_tmap_3d(&tw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, a.w, (uint64_t)a.Kp, (uint64_t)a.Np, (uint64_t)a.G * a.taps,
                   (uint64_t)a.Kp * 2, (uint64_t)a.Np * a.Kp * 2, BLOCK_K, (uint32_t)(p.block_n / 2), 1))
    return 1;

  typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, const EpiMaps, const FwdParams);
  static const KernelFn kernels[3][2] = {
      {conv_fwd_tc_kernel<false, false, 1>, conv_fwd_tc_kernel<false, false, 3>},
      {conv_fwd_tc_kernel<true, false, 1>, conv_fwd_tc_kernel<true, false, 3>},
      {conv_fwd_tc_kernel<true, true, 1>, conv_fwd_tc_kernel<true, true, 3>}};
  static bool attr_set[SD_MAX_DEVICES];
  if (first_use_on_device(attr_set)) {
    for (int m = 0; m < 3; ++m)
      for (int t = 0; t < 2; ++t)
        SD_CUDA(cudaFuncSetAttribute(kernels[m][t], cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
  }
  const KernelFn kernel = kernels[mode][a.taps == 3];
  if (pair) {
    int grid = 2 * (p.num_tiles < n_pairs ? p.num_tiles : n_pairs);
    static int max_pairs_dev[SD_MAX_DEVICES];
    static bool max_pairs_known[SD_MAX_DEVICES];
    int dev = 0;
    cudaGetDevice(&dev);
    const bool cacheable = dev >= 0 && dev < SD_MAX_DEVICES;
    int max_pairs = cacheable && max_pairs_known[dev] ? max_pairs_dev[dev] : -1;
    if (max_pairs < 0) {   // all pairs must be co-resident: the tile walk is statically strided
      cudaLaunchConfig_t q;
      memset(&q, 0, sizeof(q));
      q.gridDim = dim3(grid); q.blockDim = dim3(NUM_THREADS); q.dynamicSmemBytes = SMEM_LIMIT; q.stream = st;
      cudaLaunchAttribute attr;
      attr.id = cudaLaunchAttributeClusterDimension;
      attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
      q.attrs = &attr; q.numAttrs = 1;
      SD_CUDA(cudaOccupancyMaxActiveClusters(&max_pairs, kernels[1][1], &q));
      if (cacheable) { max_pairs_dev[dev] = max_pairs; max_pairs_known[dev] = true; }
    }
    if (2 * max_pairs < grid) grid = 2 * max_pairs;
    SD_REQUIRE(grid >= 2 && (mode != 2 || grid / 2 >= p.n_tiles), "conv_fwd_tc: not enough resident CTA pairs (%d)", grid / 2);
    SD_CUDA(launch_pdl(kernel, dim3(grid), dim3(NUM_THREADS), (size_t)smem_bytes, st, 2, ta, tw, em, p));
  } else {
    int grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();
    SD_CUDA(launch_pdl(kernel, dim3(grid), dim3(NUM_THREADS), (size_t)smem_bytes, st, 1, ta, tw, em, p));
  }
  return check_launch("conv_fwd_tc");
}

}  // namespace sd
