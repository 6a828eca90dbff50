// tcgen05 / TMEM / TMA implicit-GEMM Conv1d on fp32 storage with TF32 tensor-core math (kind::tf32):
// the "fp32/TF32" precision of BASELINE.json configs[1].  Reference semantics as conv_tc.cu:
// nn.Conv1d(k in {1,3}, padding="same", dilation) + bias (+ residual) of speech_decoding/models.py:97-109,
// 128-150,156,160,188-189 with the BatchNorm batch statistics (models.py:158,161), GELU (models.py:194-195)
// and GLU (models.py:164) fused into the epilogue.  On a GPU the reference itself runs these convolutions
// through cuDNN with TF32 (PyTorch's default for convolutions).
//
// Two arithmetic variants, chosen by the caller through the operand planes it passes:
//   * TF32      (in_lo == w_lo == NULL): one kind::tf32 MMA per product -- 10-bit mantissas, ~1e-3.
//   * 3xTF32    (in_lo, w_lo given): every operand arrives pre-split as x = hi + lo with hi = x rounded to TF32
//               (low 13 mantissa bits zero) and lo = the TF32 rounding of x - hi (sd_tf32_split), and the
//               kernel accumulates  hi*hi + hi*lo + lo*hi: three MMAs per product, operand error ~2^-21 -- the mode
//               that meets the 1e-4 parity bar ON the tensor cores.  The tensor core adds into its fp32 accumulator
//               with truncation (a bias of up to one ulp per MMA, all in the same direction), which over a 360-MMA
//               chain is what limits a plain 3xTF32 kernel to ~2e-5 per op; so the hi*hi products of even and odd
//               k-blocks go to two separate TMEM accumulators and the two small cross terms to a third, and the
//               epilogue adds the three in round-to-nearest fp32: chains six times shorter, small terms out of
//               the big sum.  (Three accumulators of <= 160 columns: no double buffering in this variant.)
//
// GEMM view per CTA tile:  D[128 time rows, BLOCK_N channels] = sum_{pass} sum_{tap j} sum_{k-block}
//     A_j[128 x 32] (fp32 activations, channels-last => K-major, rows t0+shift_j.., zero-filled outside [0,T))
//   x W_j[BLOCK_N x 32]^T (packed fp32 weights (G,taps,Np,Kp), K-major)
// 32 fp32 = 128 B = one swizzle row, so descriptors, the shared halo tile for the three taps (row-offset
// descriptors) and the ring protocol are those of the bf16 kernel.  Warp roles (320 threads): warp 0 = TMA
// producer, warp 1 = MMA issuer + TMEM allocator, warps 2..9 = epilogue (two per TMEM lane quadrant, each owning
// half of the tile's columns).  The epilogue reads / writes global memory directly from registers (every thread
// owns one output row: 64 contiguous bytes per 16-column chunk) -- fp32 staging tiles for TMA stores would not
// fit next to the operand rings, and this mode is bound by the 3x (or 2x) slower tensor rate, not by the epilogue.
#include "tc_common.cuh"

namespace sd {

using namespace tc;

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 32;   // 32 fp32 = 128 B = one swizzle row
constexpr int MAX_BLOCK_N = 256;
constexpr int MAX_A_SLOTS = 4, MAX_W_SLOTS = 8;
constexpr int TMEM_COLS = 512;
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = (2 + NUM_EPI_WARPS) * 32;
constexpr int SMEM_LIMIT = 227 * 1024;

struct Tf32Params {
  const float* bias;
  const float* res;
  float* out;
  float* preact;
  double* stats;
  float* rownorm2;
  const int* widx;
  int B, T, N, Np, Kp, taps, dil;
  int block_n, n_tiles, m_tiles_per_sample, num_tiles, k_blocks;
  int act, out_mode, D2, Op;
  int planes;      // 1: TF32, 2: operands split hi/lo (3xTF32)
  int acc_stride;  // 3xTF32: TMEM columns between the three partial accumulators
  int sa_slots, sw_slots, a_bytes, w_bytes, a_rows, halo, off_w, off_bias, off_stats, off_bar, cols_alloc;
};

struct Tf32Maps {
  CUtensorMap a[2], w[2];   // [plane]
};

// column sums of a 32-row x 16-column register tile held one row per lane (see conv_tc.cu)
__device__ __forceinline__ float colsum16(float (&v)[16], int lane) {
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
__device__ __forceinline__ int col_of_lane16(int lane) {
  return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
}

__device__ __forceinline__ void lds16_add(uint32_t saddr, float (&v)[16]) {
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    float4 f;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(f.x), "=f"(f.y), "=f"(f.z), "=f"(f.w) : "r"(saddr + 16 * h));
    v[4 * h] += f.x; v[4 * h + 1] += f.y; v[4 * h + 2] += f.z; v[4 * h + 3] += f.w;
  }
}

template <int TAPS>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_fwd_tf32_kernel(const __grid_constant__ Tf32Maps tm, const Tf32Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + p.off_bar;
  auto afull_bar = [&](int s) { return bar_base + 8u * s; };
  auto aempty_bar = [&](int s) { return bar_base + 8u * (MAX_A_SLOTS + s); };
  auto wfull_bar = [&](int s) { return bar_base + 8u * (2 * MAX_A_SLOTS + s); };
  auto wempty_bar = [&](int s) { return bar_base + 8u * (2 * MAX_A_SLOTS + MAX_W_SLOTS + s); };
  constexpr int NB = 2 * MAX_A_SLOTS + 2 * MAX_W_SLOTS;
  auto tfull_bar = [&](int a) { return bar_base + 8u * (NB + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (NB + 2 + a); };
  const uint32_t tmem_ptr_smem = bar_base + 8u * (NB + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool glu = p.act == SD_ACT_GLU;
  const int half_n = p.block_n >> 1;
  const int tile_begin = (int)blockIdx.x, tile_step = (int)gridDim.x, tile_end = p.num_tiles;

  if (threadIdx.x == 0) {
    for (int pl = 0; pl < p.planes; ++pl) { prefetch_tmap(&tm.a[pl]); prefetch_tmap(&tm.w[pl]); }
    for (int s = 0; s < p.sa_slots; ++s) { mbar_init(afull_bar(s), 1); mbar_init(aempty_bar(s), 1); }
    for (int s = 0; s < p.sw_slots; ++s) { mbar_init(wfull_bar(s), 1); mbar_init(wempty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), NUM_EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_smem, TMEM_COLS);
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
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));

  // A ring: one slot = `planes` activation tiles of (128 + 2*halo) rows x 32 channels for one k-block, loaded once
  // for all taps.  W ring: one slot = `planes` weight tiles BLOCK_N x 32 for one (k-block, tap).
  if (warp == 0) {
    int sa = 0, sw = 0;
    uint32_t pha = 0, phw = 0;
    const uint32_t a_tx = (uint32_t)p.planes * (uint32_t)p.a_rows * 128u, w_tx = (uint32_t)p.planes * (uint32_t)p.block_n * 128u;
    for (int tile = tile_begin; tile < tile_end; tile += tile_step) {
      const int m_idx = tile / p.n_tiles, n_idx = tile % p.n_tiles;
      const int b = m_idx / p.m_tiles_per_sample;
      const int t0 = (m_idx % p.m_tiles_per_sample) * BLOCK_M;
      const int g = p.widx ? __ldg(p.widx + b) : 0;
      const int row0 = glu ? n_idx * half_n : n_idx * p.block_n;
      const int row1 = glu ? p.D2 + n_idx * half_n : row0 + half_n;
      for (int kb = 0; kb < p.k_blocks; ++kb) {
        mbar_wait(aempty_bar(sa), pha ^ 1);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(afull_bar(sa), a_tx);
          for (int pl = 0; pl < p.planes; ++pl)
            tma_load_3d(smem_base + sa * (p.planes * p.a_bytes) + pl * p.a_bytes, &tm.a[pl], afull_bar(sa), kb * BLOCK_K, t0 - p.halo, b);
        }
        __syncwarp();
        if (++sa == p.sa_slots) { sa = 0; pha ^= 1; }
        for (int j = 0; j < TAPS; ++j) {
          mbar_wait(wempty_bar(sw), phw ^ 1);
          if (elect_one_sync()) {
            mbar_arrive_expect_tx(wfull_bar(sw), w_tx);
            for (int pl = 0; pl < p.planes; ++pl) {
              const uint32_t sb = smem_base + p.off_w + sw * (p.planes * p.w_bytes) + pl * p.w_bytes;
              tma_load_3d(sb, &tm.w[pl], wfull_bar(sw), kb * BLOCK_K, row0, g * TAPS + j);
              tma_load_3d(sb + half_n * 128, &tm.w[pl], wfull_bar(sw), kb * BLOCK_K, row1, g * TAPS + j);
            }
          }
          __syncwarp();
          if (++sw == p.sw_slots) { sw = 0; phw ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc(/*tf32*/ 2, 0, 0, BLOCK_M, (uint32_t)p.block_n);
    const uint32_t dhi = smem_desc_hi(1024);
    const uint32_t tap_step = (uint32_t)(p.halo * 128) >> 4;
    const uint32_t a_plane = (uint32_t)p.a_bytes >> 4, w_plane = (uint32_t)p.w_bytes >> 4;
    int sa = 0, sw = 0;
    uint32_t pha = 0, phw = 0;
    int it_tile = 0;
    const bool x3 = p.planes == 2;
    for (int tile = tile_begin; tile < tile_end; tile += tile_step, ++it_tile) {
      // TF32: two accumulator buffers of 256 columns, alternating per tile.  3xTF32: ONE buffer of three accumulators
      // (hi*hi of even k-blocks | hi*hi of odd k-blocks | cross terms), p.acc_stride columns apart.
      const int acc = x3 ? 0 : (it_tile & 1);
      const uint32_t acc_ph = x3 ? (it_tile & 1) : ((it_tile >> 1) & 1);
      mbar_wait(tempty_bar(acc), acc_ph ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * MAX_BLOCK_N;
      for (int kb = 0; kb < p.k_blocks; ++kb) {
        mbar_wait(afull_bar(sa), pha);
        const uint32_t alo = smem_desc_lo(smem_base + sa * (p.planes * p.a_bytes), 16);
        const uint32_t d_main = x3 ? d_tmem + (kb & 1) * p.acc_stride : d_tmem;
        const uint32_t d_small = d_tmem + 2 * p.acc_stride;
        for (int j = 0; j < TAPS; ++j) {
          mbar_wait(wfull_bar(sw), phw);
          tc_fence_after();
          if (elect_one_sync()) {
            const uint32_t blo = smem_desc_lo(smem_base + p.off_w + sw * (p.planes * p.w_bytes), 16);
            const uint32_t aj = alo + j * tap_step;
            const uint32_t first_main = x3 ? (uint32_t)(kb < 2 && j == 0) : (uint32_t)((kb | j) == 0);
#pragma unroll
            for (int k = 0; k < BLOCK_K / 8; ++k)   // +32 B per 8-element k-step inside the swizzled row
              umma_tf32(d_main, desc64(aj + 2 * k, dhi), desc64(blo + 2 * k, dhi), idesc, !(first_main && k == 0));
            if (x3) {
#pragma unroll
              for (int k = 0; k < BLOCK_K / 8; ++k)   // hi * lo
                umma_tf32(d_small, desc64(aj + 2 * k, dhi), desc64(blo + w_plane + 2 * k, dhi), idesc, (kb | j | k) != 0);
#pragma unroll
              for (int k = 0; k < BLOCK_K / 8; ++k)   // lo * hi
                umma_tf32(d_small, desc64(aj + a_plane + 2 * k, dhi), desc64(blo + 2 * k, dhi), idesc, 1);
            }
            umma_commit(wempty_bar(sw));
            if (j == TAPS - 1) {
              umma_commit(aempty_bar(sa));
              if (kb == p.k_blocks - 1) umma_commit(tfull_bar(acc));
            }
          }
          __syncwarp();
          if (++sw == p.sw_slots) { sw = 0; phw ^= 1; }
        }
        if (++sa == p.sa_slots) { sa = 0; pha ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int ew = warp - 2;
    const int quad = warp & 3;
    const int hsel = ew >> 2;
    const uint32_t s_bias = smem_base + p.off_bias;
    float* s_stats = reinterpret_cast<float*>(smem_gen + p.off_stats);
    const int nch = glu ? (half_n >> 4) : (p.block_n >> 4);
    const int ch0 = hsel ? (nch + 1) / 2 : 0;
    const int ch1 = hsel ? nch : (nch + 1) / 2;
    int it_tile = 0;
    const bool x3 = p.planes == 2;
    const bool two_main = x3 && p.k_blocks > 1;     // the odd-k-block accumulator exists
    // 16 accumulator columns; 3xTF32: the sum of the three partial accumulators (round-to-nearest fp32 adds)
    auto load_acc = [&](uint32_t addr, float (&v)[16]) {
      uint32_t r[16];
      tmem_ld16(addr, r);
      if (x3) {
        uint32_t r1[16], r2[16];
        tmem_ld16(addr + 2 * p.acc_stride, r2);
        if (two_main) tmem_ld16(addr + p.acc_stride, r1);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float m = __uint_as_float(r[i]);
          if (two_main) m += __uint_as_float(r1[i]);
          v[i] = m + __uint_as_float(r2[i]);
        }
      } else {
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
      }
    };
    for (int tile = tile_begin; tile < tile_end; tile += tile_step, ++it_tile) {
      const int acc = x3 ? 0 : (it_tile & 1);
      const uint32_t acc_ph = x3 ? (it_tile & 1) : ((it_tile >> 1) & 1);
      const int m_idx = tile / p.n_tiles, n_idx = tile % p.n_tiles;
      const int b = m_idx / p.m_tiles_per_sample;
      const int t = (m_idx % p.m_tiles_per_sample) * BLOCK_M + quad * 32 + lane;
      const bool valid = t < p.T;
      const int n0 = glu ? n_idx * half_n : n_idx * p.block_n;
      mbar_wait_relaxed(tfull_bar(acc), acc_ph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * MAX_BLOCK_N;
      const size_t row = (size_t)b * p.T + (valid ? t : 0);
      float sumsq = 0.f;

      if (!glu) {
        for (int c = ch0; c < ch1; ++c) {
          const int cc = c * 16, nb = n0 + cc;
          float v[16];
          load_acc(taddr + cc, v);
          if (nb >= p.Np) continue;     // (warp-uniform) chunk entirely beyond the tensor
          lds16_add(s_bias + nb * 4, v);
          if (p.res && valid) {
            const float* rs = p.res + row * p.Np + nb;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (nb + 4 * q < p.Np) {
                const float4 f = *reinterpret_cast<const float4*>(rs + 4 * q);
                v[4 * q] += f.x; v[4 * q + 1] += f.y; v[4 * q + 2] += f.z; v[4 * q + 3] += f.w;
              }
            }
          }
          if (p.act == SD_ACT_GELU) {
            if (p.preact && valid) {
              float* dst = p.preact + row * p.Np + nb;
#pragma unroll
              for (int q = 0; q < 4; ++q)
                if (nb + 4 * q < p.Np) *reinterpret_cast<float4*>(dst + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = gelu_f(v[i]);
          }
          if (p.out_mode == SD_OUT_NCT_F32) {
            if (valid) {
              float* dst = p.out + (size_t)b * p.N * p.T + (size_t)nb * p.T + t;   // lanes = consecutive t
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                if (nb + i < p.N) {
                  dst[(size_t)i * p.T] = v[i];
                  sumsq = fmaf(v[i], v[i], sumsq);
                }
              }
            }
          } else if (valid) {
            float* dst = p.out + row * p.Np + nb;
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (nb + 4 * q < p.Np) *reinterpret_cast<float4*>(dst + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          }
          if (p.stats) {
            // BatchNorm batch statistics from the fp32 values exactly as stored; each (quadrant, column) accumulator
            // belongs to one warp, so the owning lane does a plain read-modify-write
            float s[16], q2[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) { s[i] = valid ? v[i] : 0.f; q2[i] = s[i] * s[i]; }
            const float cs = colsum16(s, lane), cq = colsum16(q2, lane);
            if ((lane & 1) == 0) {
              const int col = nb + col_of_lane16(lane);
              s_stats[quad * 2 * p.cols_alloc + col] += cs;
              s_stats[(quad * 2 + 1) * p.cols_alloc + col] += cq;
            }
          }
        }
      } else {
        for (int c = ch0; c < ch1; ++c) {
          const int cc = c * 16, cb = n0 + cc;
          float va[16], vb[16];
          load_acc(taddr + cc, va);
          load_acc(taddr + half_n + cc, vb);
          if (cb >= p.D2) continue;
          lds16_add(s_bias + cb * 4, va);
          lds16_add(s_bias + (p.cols_alloc + cb) * 4, vb);
          if (valid) {
            const bool vec = cb + 16 <= p.D2 && (p.D2 & 3) == 0;     // (warp-uniform) whole chunk, 16-byte aligned halves
            if (p.preact) {
              float* pa = p.preact + row * p.Np + cb;
              if (vec) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  *reinterpret_cast<float4*>(pa + 4 * q) = make_float4(va[4 * q], va[4 * q + 1], va[4 * q + 2], va[4 * q + 3]);
                  *reinterpret_cast<float4*>(pa + p.D2 + 4 * q) = make_float4(vb[4 * q], vb[4 * q + 1], vb[4 * q + 2], vb[4 * q + 3]);
                }
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                  if (cb + i < p.D2) { pa[i] = va[i]; pa[p.D2 + i] = vb[i]; }
              }
            }
            float* dst = p.out + row * p.Op + cb;
            if (vec) {
#pragma unroll
              for (int i = 0; i < 16; ++i) va[i] *= sigmoid_f(vb[i]);
#pragma unroll
              for (int q = 0; q < 4; ++q)
                *reinterpret_cast<float4*>(dst + 4 * q) = make_float4(va[4 * q], va[4 * q + 1], va[4 * q + 2], va[4 * q + 3]);
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (cb + i < p.Op) dst[i] = (cb + i < p.D2) ? va[i] * sigmoid_f(vb[i]) : 0.f;
            }
          }
        }
      }
      if (p.rownorm2) {
        sumsq = warp_sum(sumsq);
        if (lane == 0 && b < p.B) atomicAdd(p.rownorm2 + b, sumsq);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
    }
    if (p.stats) {
      asm volatile("bar.sync 1, %0;" ::"n"(NUM_EPI_WARPS * 32) : "memory");
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
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

int pick_block_n_tf32(int n_total, int gran, int max_bn) {
  int best_bn = gran, best_pad = 1 << 30;
  const int min_tiles = (n_total + max_bn - 1) / max_bn;
  for (int nt = min_tiles; nt <= min_tiles + 3; ++nt) {
    int bn = ((n_total + nt - 1) / nt + gran - 1) / gran * gran;
    if (bn > max_bn) continue;
    int pad = bn * nt;
    if (pad < best_pad) { best_pad = pad; best_bn = bn; }
  }
  return best_bn;
}

}  // namespace

bool conv_fwd_tf32_supported(const sd_conv_args& a) {
  if (a.dtype != SD_TF32) return false;
  if ((a.in_lo == nullptr) != (a.w_lo == nullptr)) return false;
  if (a.act == SD_ACT_GLU && (a.out_mode != SD_OUT_BTC || a.N % 2)) return false;
  if (a.stats && (a.act != SD_ACT_NONE || a.out_mode != SD_OUT_BTC)) return false;
  if (a.rownorm2 && a.out_mode != SD_OUT_NCT_F32) return false;
  if (a.res && ((a.act != SD_ACT_NONE && a.act != SD_ACT_GELU) || a.out_mode != SD_OUT_BTC)) return false;
  if (a.affine) return false;
  if (a.out_mode == SD_OUT_NCT_F32 && a.act != SD_ACT_GELU) return false;
  if (((uintptr_t)a.in & 15) || ((uintptr_t)a.w & 15) || ((uintptr_t)a.out & 15) || ((uintptr_t)a.res & 15) ||
      ((uintptr_t)a.preact & 15) || ((uintptr_t)a.in_lo & 15) || ((uintptr_t)a.w_lo & 15))
    return false;
  if (a.Np > 4096) return false;
  return true;
}

int conv_fwd_tf32(const sd_conv_args& a, cudaStream_t st) {
  const bool glu = a.act == SD_ACT_GLU;
  Tf32Params p;
  memset(&p, 0, sizeof(p));
  p.bias = a.bias;
  p.res = reinterpret_cast<const float*>(a.res);
  p.out = reinterpret_cast<float*>(a.out);
  p.preact = reinterpret_cast<float*>(a.preact);
  p.stats = a.stats;
  p.rownorm2 = a.rownorm2;
  p.widx = a.widx;
  p.B = a.B; p.T = a.T; p.N = a.N; p.Np = a.Np; p.Kp = a.Kp; p.taps = a.taps; p.dil = a.dil;
  p.act = a.act; p.out_mode = a.out_mode;
  p.D2 = glu ? a.N / 2 : 0;
  p.Op = glu ? (p.D2 + 7) / 8 * 8 : 0;
  p.planes = a.in_lo ? 2 : 1;
  p.m_tiles_per_sample = (a.T + BLOCK_M - 1) / BLOCK_M;
  p.k_blocks = (a.Kp + BLOCK_K - 1) / BLOCK_K;
  p.halo = a.taps == 3 ? a.dil : 0;
  p.a_rows = BLOCK_M + 2 * p.halo;
  p.a_bytes = (p.a_rows * 128 + 1023) / 1024 * 1024;
  SD_REQUIRE(p.a_rows <= 256, "conv_fwd_tf32: dilation %d too large for one activation tile", a.dil);

  // shared-memory plan: widest column tile (<= 256, or <= 128 with split operands) whose rings fit
  const int n_total = glu ? 2 * p.Op : a.Np, gran = glu ? 32 : 16;
  int smem_bytes = 0;
  bool ok = false;
  // 3xTF32 keeps three accumulators of block_n columns in the 512 TMEM columns
  for (int max_bn = (p.planes == 2 ? 160 : MAX_BLOCK_N); max_bn >= 64 && !ok; max_bn -= 32) {
    const int bn = pick_block_n_tf32(n_total, gran, max_bn);
    p.block_n = bn;
    if (glu) {
      p.n_tiles = (p.Op + bn / 2 - 1) / (bn / 2);
      p.cols_alloc = p.n_tiles * (bn / 2);
    } else {
      p.n_tiles = (a.Np + bn - 1) / bn;
      p.cols_alloc = p.n_tiles * bn;
    }
    p.w_bytes = bn * 128;
    p.acc_stride = (bn + 31) / 32 * 32;
    const int tail = (glu ? 2 : 1) * p.cols_alloc * 4 + (a.stats ? 8 * p.cols_alloc * 4 : 0) + 16 + 512;
    const int ring = SMEM_LIMIT - 1024 - tail;
    const int a_slot = p.planes * p.a_bytes, w_slot = p.planes * p.w_bytes;
    int sa = 2;
    int sw = (ring - sa * a_slot) / w_slot;
    if (sw > MAX_W_SLOTS) sw = MAX_W_SLOTS;
    if (sw < (a.taps == 3 ? 3 : 2)) continue;
    if (sw > 4 && (ring - 3 * a_slot) / w_slot >= 4) { sa = 3; sw = (ring - sa * a_slot) / w_slot; if (sw > MAX_W_SLOTS) sw = MAX_W_SLOTS; }
    p.sa_slots = sa; p.sw_slots = sw;
    p.off_w = sa * a_slot;
    int off = p.off_w + sw * w_slot;
    p.off_bias = off; off += (glu ? 2 : 1) * p.cols_alloc * 4;
    p.off_stats = off; off += a.stats ? 8 * p.cols_alloc * 4 : 0;
    off = (off + 15) / 16 * 16;
    p.off_bar = off; off += 512;
    smem_bytes = off + 1024;
    ok = smem_bytes <= SMEM_LIMIT;
  }
  SD_REQUIRE(ok, "conv_fwd_tf32: operand rings do not fit (N=%d K=%d dil=%d)", a.N, a.Kp, a.dil);
  p.num_tiles = a.B * p.m_tiles_per_sample * p.n_tiles;

  Tf32Maps tm;
  memset(&tm, 0, sizeof(tm));
  const void* ins[2] = {a.in, a.in_lo};
  const void* ws[2] = {a.w, a.w_lo};
  for (int pl = 0; pl < p.planes; ++pl) {
    if (make_tmap_3d(&tm.a[pl], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, ins[pl], (uint64_t)a.Kp, (uint64_t)a.T, (uint64_t)a.B,
                     (uint64_t)a.Kp * 4, (uint64_t)a.T * a.Kp * 4, BLOCK_K, (uint32_t)p.a_rows, 1))
      return 1;
    if (make_tmap_3d(&tm.w[pl], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, ws[pl], (uint64_t)a.Kp, (uint64_t)a.Np, (uint64_t)a.G * a.taps,
                     (uint64_t)a.Kp * 4, (uint64_t)a.Np * a.Kp * 4, BLOCK_K, (uint32_t)(p.block_n / 2), 1))
      return 1;
  }
  if (p.planes == 1) { tm.a[1] = tm.a[0]; tm.w[1] = tm.w[0]; }

  typedef void (*KernelFn)(const Tf32Maps, const Tf32Params);
  const KernelFn kernel = a.taps == 3 ? conv_fwd_tf32_kernel<3> : conv_fwd_tf32_kernel<1>;
  SD_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
  const int grid = p.num_tiles < sm_budget() ? p.num_tiles : sm_budget();
  kernel<<<grid, NUM_THREADS, smem_bytes, st>>>(tm, p);
  return check_launch("conv_fwd_tf32");
}

// x -> hi = x rounded to the nearest TF32 (cvt.rna: low 13 mantissa bits zero, so it is exactly representable whatever
// the tensor core does with the low bits of its inputs) and lo = rna_tf32(x - hi) (x - hi is exact in fp32; |lo| <=
// 2^-11 |x|, its own rounding error <= 2^-23 |x| and unbiased)
__device__ __forceinline__ float rna_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}
__global__ void __launch_bounds__(256) tf32_split_kernel(const float4* __restrict__ x, float4* __restrict__ hi, float4* __restrict__ lo, int64_t n4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = x[i];
    float4 h, l;
    h.x = rna_tf32(v.x); l.x = rna_tf32(v.x - h.x);
    h.y = rna_tf32(v.y); l.y = rna_tf32(v.y - h.y);
    h.z = rna_tf32(v.z); l.z = rna_tf32(v.z - h.z);
    h.w = rna_tf32(v.w); l.w = rna_tf32(v.w - h.w);
    hi[i] = h;
    lo[i] = l;
  }
}

}  // namespace sd

extern "C" int sd_tf32_split(const float* x, float* hi, float* lo, int64_t n, void* stream) {
  using namespace sd;
  SD_REQUIRE(n % 4 == 0 && !(((uintptr_t)x | (uintptr_t)hi | (uintptr_t)lo) & 15), "sd_tf32_split: n must be a multiple of 4 and the pointers 16-byte aligned");
  if (n == 0) return 0;
  int64_t blocks = (n / 4 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  tf32_split_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(hi),
                                                                  reinterpret_cast<float4*>(lo), n / 4);
  return check_launch("tf32_split");
}
