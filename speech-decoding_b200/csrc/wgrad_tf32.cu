// tcgen05 / TMEM / TMA Conv1d weight gradient on fp32 storage with TF32 tensor-core math (kind::tf32) -- the
// fp32/TF32 precision of BASELINE.json configs[1]; see wgrad_tc.cu for the bf16 version and the reference
// semantics (autograd of nn.Conv1d, speech_decoding/models.py:97-109,128-150,188-189; grouped per-subject form
// models.py:98-116).
//
//   dw[g, n, c, j] += sum_{b in group g} sum_t dy[b,t,n] * x[b, t + (j-(taps-1)/2)*dil, c]
//
// Both operands are read straight from the channels-last fp32 activations: the contraction index t is the row index,
// so A = dy^T and B = x^T are MN-major operands.  32-bit MN-major operands use the 128-byte swizzle with 32-byte
// atoms (UMMA layout type 1, TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B; the layout clip_tc.cu's gradient GEMM uses):
// one "atom column" = 32 contiguous channels (128 B) x BLOCK_T time rows; the pattern repeats every 4 rows
// (SBO = 512), a K = 8 MMA step advances 8 rows (1024 B), consecutive 32-channel blocks are LBO bytes apart.
// Work item = (n-tile, c-tile, tap, group, split-K slice of the batch); one item per CTA; the fp32 tile goes to the
// split-K workspace (summed by wgrad_reduce) or straight to dw with red.global.add.
//
// 3xTF32: with dout_lo / in_lo given, every (sample, time block) is contracted three times -- (dy_hi, x_hi),
// (dy_lo, x_hi), (dy_hi, x_lo) (operands pre-split by sd_tf32_split).  The tensor core adds into its fp32 accumulator
// with truncation (a one-sided bias of up to an ulp per MMA), so the long hi*hi chain is split: even and odd samples
// accumulate into two separate TMEM accumulators, the two small cross passes into a third, the epilogue adds the three
// in round-to-nearest fp32, and twice as many split-K slices are used as SMs.
// dbias comes from an extra N=16 MMA per k-step against a constant tile of ones (passes 0 and 1: dy_hi + dy_lo).
#include "tc_common.cuh"

namespace sd {

using namespace tc;

namespace {

constexpr int BLOCK_MN = 128;           // out-channel tile (UMMA M)
constexpr int BLOCK_T = 32;             // contraction block: 32 time steps
constexpr int ATOM_BYTES = BLOCK_T * 128;   // 32 channels (128 B) x 32 time rows
constexpr int A_ATOMS = BLOCK_MN / 32;
constexpr int STAGES = 4;
constexpr int TMEM_COLS = 512;
constexpr int BIAS_COL = 256;
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = (2 + NUM_EPI_WARPS) * 32;
constexpr int ONES_BYTES = ATOM_BYTES;

struct WgTf32Params {
  float* dw;
  float* dbias;
  const int* sample_order;
  const int* group_offsets;
  int B, T, N, K, taps, dil, G;
  long long gs, sn, sk, sj;
  int block_c, c_atoms, n_tiles, c_tiles, nsplit, stage_bytes, passes, acc_stride, bias_col;
  float* ws;
  float* ws_bias;
};

struct WgTf32Maps {
  CUtensorMap dy[2], x[2];   // [plane]
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_wgrad_tf32_kernel(const __grid_constant__ WgTf32Maps tm, const WgTf32Params p) {
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
    for (int pl = 0; pl < (p.passes > 1 ? 2 : 1); ++pl) { prefetch_tmap(&tm.dy[pl]); prefetch_tmap(&tm.x[pl]); }
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  {  // constant tile of fp32 ones for the bias MMA
    uint32_t* ones = reinterpret_cast<uint32_t*>(smem_gen + STAGES * p.stage_bytes);
    for (int i = threadIdx.x; i < ONES_BYTES / 4; i += NUM_THREADS) ones[i] = 0x3F800000u;
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_smem, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));

  // pass -> (dy plane, x plane): hi*hi, lo*hi, hi*lo
  if (warp == 0) {
    int s = 0;
    uint32_t ph = 0;
    const uint32_t stage_tx = (uint32_t)(A_ATOMS + p.c_atoms) * ATOM_BYTES;
    for (int pass = 0; pass < p.passes; ++pass) {
      const CUtensorMap* mdy = &tm.dy[pass == 1 ? 1 : 0];
      const CUtensorMap* mx = &tm.x[pass == 2 ? 1 : 0];
      for (int pos = s_begin; pos < s_end; ++pos) {
        const int b = p.sample_order ? __ldg(p.sample_order + pos) : pos;
        for (int tb = 0; tb < t_blocks; ++tb) {
          mbar_wait(empty_bar(s), ph ^ 1);
          if (elect_one_sync()) {
            const uint32_t sa = smem_base + s * p.stage_bytes, sb = sa + A_ATOMS * ATOM_BYTES;
            mbar_arrive_expect_tx(full_bar(s), stage_tx);
#pragma unroll
            for (int a = 0; a < A_ATOMS; ++a) tma_load_3d(sa + a * ATOM_BYTES, mdy, full_bar(s), n0 + 32 * a, tb * BLOCK_T, b);
            for (int a = 0; a < p.c_atoms; ++a) tma_load_3d(sb + a * ATOM_BYTES, mx, full_bar(s), c0 + 32 * a, tb * BLOCK_T + shift, b);
          }
          __syncwarp();
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc(/*tf32*/ 2, /*A MN-major*/ 1, /*B MN-major*/ 1, BLOCK_MN, (uint32_t)p.block_c);
    const uint32_t idesc_b = make_idesc(2, 1, 1, BLOCK_MN, 16);
    const uint32_t dhi = smem_desc_hi(512, /*SWIZZLE_128B_BASE32B*/ 1);
    const uint32_t olo = smem_desc_lo(ones_base, ATOM_BYTES);
    int s = 0;
    uint32_t ph = 0;
    const int per_pass = (s_end - s_begin) * t_blocks;
    // 8-row MMA steps of the last time block that still hold rows t < T (TMA zero-fills the rest: skip them)
    const int k_last = (p.T - (t_blocks - 1) * BLOCK_T + 7) >> 3;
    int it = 0;
    uint32_t started = 0;          // bit a: accumulator a has been written
    for (int pass = 0; pass < p.passes; ++pass) {
      const bool bias_pass = do_bias && pass < 2;
      int tb = 0, smp = 0;
      for (int i = 0; i < per_pass; ++i, ++it) {
        const int k_steps = tb == t_blocks - 1 ? k_last : BLOCK_T / 8;
        // 3xTF32: hi*hi of even / odd samples -> accumulators 0 / 1, the cross passes -> accumulator 2
        const int a_idx = p.passes == 1 ? 0 : (pass == 0 ? (smp & 1) : 2);
        // bias column sums: dy_hi of even samples -> bias accumulator 0, dy_hi of odd samples and dy_lo -> accumulator 1
        const int b_idx = p.passes == 1 ? 0 : ((pass == 0 && !(smp & 1)) ? 0 : 1);
        if (++tb == t_blocks) { tb = 0; ++smp; }
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t alo = smem_desc_lo(smem_base + s * p.stage_bytes, ATOM_BYTES), blo = alo + ((A_ATOMS * ATOM_BYTES) >> 4);
          const uint32_t d = tmem_base + a_idx * p.acc_stride;
          const bool fresh = !((started >> a_idx) & 1), fresh_b = !((started >> (4 + b_idx)) & 1);
#pragma unroll
          for (int k = 0; k < BLOCK_T / 8; ++k) {
            if (k >= k_steps) break;
            umma_tf32(d, desc64(alo + k * (1024 >> 4), dhi), desc64(blo + k * (1024 >> 4), dhi), idesc, !(fresh && k == 0));
            if (bias_pass) umma_tf32(tmem_base + p.bias_col + 16 * b_idx, desc64(alo + k * (1024 >> 4), dhi), desc64(olo, dhi), idesc_b, !(fresh_b && k == 0));
          }
          umma_commit(empty_bar(s));
          if (it == p.passes * per_pass - 1) umma_commit(tfull_bar);
        }
        started |= 1u << a_idx;
        if (bias_pass) started |= 1u << (4 + b_idx);
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else {
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
    const bool x3 = p.passes == 3;
    const bool two_main = x3 && (s_end - s_begin) > 1;       // the odd-sample accumulator was written
    for (int c = ch0; c < ch1; ++c) {
      uint32_t r[16];
      tmem_ld16(taddr + c * 16, r);
      if (x3) {
        uint32_t r1[16], r2[16];
        tmem_ld16(taddr + 2 * p.acc_stride + c * 16, r2);
        if (two_main) tmem_ld16(taddr + p.acc_stride + c * 16, r1);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float m = __uint_as_float(r[i]);
          if (two_main) m += __uint_as_float(r1[i]);
          r[i] = __float_as_uint(m + __uint_as_float(r2[i]));
        }
      } else {
        tmem_ld_wait();
      }
      if (wsn) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<uint4*>(wsn + c * 16 + 4 * q) = make_uint4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
      } else if (n < p.N) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int cc = c0 + c * 16 + i;
          if (cc < p.K) atomicAdd(dwn + (long long)cc * p.sk, __uint_as_float(r[i]));
        }
      }
    }
    if (do_bias && hsel == 0) {
      uint32_t r[16];
      tmem_ld16(taddr + p.bias_col, r);
      if (x3) {       // + the second bias accumulator (always written: it takes the dy_lo pass)
        uint32_t r1[16];
        tmem_ld16(taddr + p.bias_col + 16, r1);
        tmem_ld_wait();
        r[0] = __float_as_uint(__uint_as_float(r[0]) + __uint_as_float(r1[0]));
      } else {
        tmem_ld_wait();
      }
      if (p.ws_bias) p.ws_bias[(size_t)split * wrows + n0 + quad * 32 + lane] = __uint_as_float(r[0]);
      else if (n < p.N) atomicAdd(p.dbias + n, __uint_as_float(r[0]));
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

int pick_block_c_tf32(int kp, int max_c) {
  int best = 16, best_pad = 1 << 30;
  const int min_tiles = (kp + max_c - 1) / max_c;
  for (int nt = min_tiles; nt <= min_tiles + 3; ++nt) {
    int bc = ((kp + nt - 1) / nt + 15) / 16 * 16;
    if (bc > max_c) continue;
    if (bc * nt < best_pad) { best_pad = bc * nt; best = bc; }
  }
  return best;
}

}  // namespace

// defined in wgrad_tc.cu
int wgrad_reduce_launch(const float* ws, const float* ws_bias, float* dw, float* dbias, int N, int K, int taps, int wrows,
                        int wcols, int nsplit, long long sn, long long sk, long long sj, cudaStream_t st);

bool conv_wgrad_tf32_supported(const sd_wgrad_args& a) {
  if (a.dtype != SD_TF32) return false;
  if ((a.dout_lo == nullptr) != (a.in_lo == nullptr)) return false;
  if (((uintptr_t)a.dout & 15) || ((uintptr_t)a.in & 15) || ((uintptr_t)a.dout_lo & 15) || ((uintptr_t)a.in_lo & 15)) return false;
  if ((a.group_offsets == nullptr) != (a.G == 1)) return false;
  return true;
}

int conv_wgrad_tf32(const sd_wgrad_args& a, cudaStream_t st) {
  WgTf32Params p;
  memset(&p, 0, sizeof(p));
  p.dw = a.dw; p.dbias = a.dbias; p.sample_order = a.sample_order; p.group_offsets = a.group_offsets;
  p.B = a.B; p.T = a.T; p.N = a.N; p.K = a.K; p.taps = a.taps; p.dil = a.dil; p.G = a.G;
  p.gs = a.gs; p.sn = a.sn; p.sk = a.sk; p.sj = a.sj;
  p.passes = a.dout_lo ? 3 : 1;
  // stage = dy tile (16 KB) + x tile (block_c/32 atoms of 4 KB); 4 stages + ones tile must fit
  // 3xTF32: three accumulators of block_c columns + the bias column share the 512 TMEM columns
  p.block_c = pick_block_c_tf32(a.Kp, p.passes == 3 ? 160 : 256);
  p.acc_stride = p.passes == 3 ? 160 : 0;
  p.bias_col = p.passes == 3 ? 480 : BIAS_COL;
  p.c_atoms = (p.block_c + 31) / 32;
  p.n_tiles = (a.Np + BLOCK_MN - 1) / BLOCK_MN;
  p.c_tiles = (a.Kp + p.block_c - 1) / p.block_c;
  const int base_items = p.n_tiles * p.c_tiles * a.taps * a.G;
  const int sms = sm_budget();
  int nsplit = a.G > 1 ? 1 : sms / base_items;
  if (nsplit < 1) nsplit = 1;
  if (p.passes == 3 && a.G == 1) nsplit *= 2;       // shorter accumulation chains (see the header comment)
  if (nsplit > a.B) nsplit = a.B;
  if (a.G == 1) nsplit = cdiv(a.B, cdiv(a.B, nsplit));
  p.nsplit = nsplit;
  p.stage_bytes = (A_ATOMS + p.c_atoms) * ATOM_BYTES;
  const int smem_bytes = STAGES * p.stage_bytes + ONES_BYTES + 256 + 1024;
  SD_REQUIRE(smem_bytes <= 227 * 1024, "conv_wgrad_tf32: stage ring does not fit");

  WgTf32Maps tm;
  memset(&tm, 0, sizeof(tm));
  const void* dys[2] = {a.dout, a.dout_lo};
  const void* xs[2] = {a.in, a.in_lo};
  const CUtensorMapSwizzle SW = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
  for (int pl = 0; pl < (p.passes > 1 ? 2 : 1); ++pl) {
    if (make_tmap_3d(&tm.dy[pl], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, dys[pl], (uint64_t)a.Np, (uint64_t)a.T, (uint64_t)a.B,
                     (uint64_t)a.Np * 4, (uint64_t)a.T * a.Np * 4, 32, BLOCK_T, 1, SW))
      return 1;
    if (make_tmap_3d(&tm.x[pl], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, xs[pl], (uint64_t)a.Kp, (uint64_t)a.T, (uint64_t)a.B,
                     (uint64_t)a.Kp * 4, (uint64_t)a.T * a.Kp * 4, 32, BLOCK_T, 1, SW))
      return 1;
  }
  if (p.passes == 1) { tm.dy[1] = tm.dy[0]; tm.x[1] = tm.x[0]; }
  SD_CUDA(cudaFuncSetAttribute(conv_wgrad_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));

  void* ws = a.workspace;
  size_t need = 0, bias_off = 0;
  if (nsplit > 1) {
    const size_t wrows = (size_t)p.n_tiles * BLOCK_MN, wcols = (size_t)p.c_tiles * p.block_c;
    need = (size_t)nsplit * a.taps * wrows * wcols * sizeof(float);
    bias_off = need;
    if (a.dbias) need += (size_t)nsplit * wrows * sizeof(float);
  }
  const bool use_ws = need > 0 && ws != nullptr && (size_t)a.workspace_bytes >= need && a.G == 1;
  p.ws = use_ws ? reinterpret_cast<float*>(ws) : nullptr;
  p.ws_bias = (use_ws && a.dbias) ? reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + bias_off) : nullptr;
  conv_wgrad_tf32_kernel<<<base_items * nsplit, NUM_THREADS, smem_bytes, st>>>(tm, p);
  if (check_launch("conv_wgrad_tf32")) return 1;
  if (use_ws)
    return wgrad_reduce_launch(p.ws, p.ws_bias, a.dw, a.dbias, a.N, a.K, a.taps, p.n_tiles * BLOCK_MN, p.c_tiles * p.block_c,
                               nsplit, a.sn, a.sk, a.sj, st);
  return 0;
}

}  // namespace sd
