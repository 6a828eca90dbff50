// Peer-memory plumbing for the data-parallel exchange of the speech rows: each rank owns a receive
// buffer allocated here (cudaMalloc, so that it has a CUDA IPC handle), every other rank maps it through the handle
// and PUSHES its rows into its slot with cudaMemcpyAsync -- the transfer runs on the copy engines over NVLink and
// occupies no SM, so it overlaps the persistent tcgen05 grids of the encoder forward without stealing an SM from
// them (an SM-resident NCCL all-gather kernel next to a 148-CTA persistent grid costs that grid a second wave).
// The reference is single-device (train.py:31); this is the build's batch-sharding addition.
#include "common.cuh"

using namespace sd;

namespace sd {
// src may be PINNED HOST memory (directly addressable from the device under unified addressing): the SMs read it over
// PCIe themselves, so the copy does not queue on the host-to-device copy engine behind a bulk input transfer
__global__ void copy_small_kernel(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src[i];
}
// Arrival fence of the copy-engine all-gather: every peer writes `expected` into its flag word of this rank's receive
// buffer AFTER its rows (same stream, so the copy engine has completed the rows first).  One thread per peer polls; the
// kernel boundary orders the CLIP kernels that follow in the stream after the rows' arrival.  An SM-resident NCCL
// collective as the fence would have to find a free SM next to back-to-back persistent 148-CTA grids (measured: it is
// starved for the whole encoder forward); this kernel is an ordinary in-stream launch.
__global__ void peer_wait_kernel(const int* flags, int world, int expected) {
  const int i = threadIdx.x;
  if (i < world) {
    const volatile int* f = flags + i;
    uint64_t t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (uint32_t spin = 0; *f != expected; ++spin) {
      __nanosleep(200);
      if ((spin & 0xfff) == 0xfff) {
        uint64_t t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 10000000000ull) {   // 10 s: a peer died or the protocol is broken -- abort instead of hanging
          printf("sd_b200: peer gather fence timed out (peer %d: flag %d, expected %d)\n", i, *f, expected);
          __trap();
        }
      }
    }
  }
  __threadfence_system();
}

// ---------------------------------------------------------------------------------------------------------------------
// Small all-gather THROUGH PEER MEMORY in one kernel (the latency-bound exchanges of the data-parallel step: SyncBN
// statistics (2 x 320 doubles), BatchNorm-backward sums, CLIP row statistics (global rows x 2 floats), loss partials).
// Every rank owns a mailbox (IPC-mapped by its peers): [2 parity slots][world][cap bytes] payload + [2][world] flag words.
// The kernel (one CTA) stores this rank's payload into its row of EVERY peer's mailbox over NVLink, publishes the epoch
// number into its flag word there (release, system scope), waits until all peers' flags in its OWN mailbox carry the
// epoch (acquire) and copies the world rows out.  ~10 us at 8 GPUs against ~60 us for a small NCCL all-reduce that has to
// find a free SM next to persistent grids; summation of the gathered rows happens in a fixed order, so every rank gets
// bit-identical results.  Two parity slots suffice: a peer can run at most one exchange ahead (it needs this rank's
// contribution to finish the next one).
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_release_sys(int* p, int v) { asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(512) peer_exchange_kernel(const uint32_t* __restrict__ src, int words, char* const* __restrict__ peers,
                                                            int rank, int world, long long cap, int epoch, uint32_t* __restrict__ out) {
  const int slot = epoch & 1;
  const long long flags_off = 2ll * world * cap;
  for (int q = 0; q < world; ++q) {
    uint32_t* dst = reinterpret_cast<uint32_t*>(peers[q] + ((long long)slot * world + rank) * cap);
    for (int i = threadIdx.x; i < words; i += blockDim.x) dst[i] = src[i];
  }
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < world) {
    st_release_sys(reinterpret_cast<int*>(peers[threadIdx.x] + flags_off) + slot * world + rank, epoch);
    const int* mine = reinterpret_cast<const int*>(peers[rank] + flags_off) + slot * world + threadIdx.x;
    uint64_t t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (uint32_t spin = 0; ld_acquire_sys(mine) != epoch; ++spin) {
      if ((spin & 0xffff) == 0xffff) {
        uint64_t t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 10000000000ull) {
          printf("sd_b200: peer exchange timed out (rank %d waiting for rank %d, epoch %d)\n", rank, (int)threadIdx.x, epoch);
          __trap();
        }
      }
    }
  }
  __syncthreads();
  for (int q = 0; q < world; ++q) {
    const uint32_t* row = reinterpret_cast<const uint32_t*>(peers[rank] + ((long long)slot * world + q) * cap);
    for (int i = threadIdx.x; i < words; i += blockDim.x) out[(long long)q * words + i] = __ldcv(row + i);
  }
}

// out[i] = sum_q in[q][i] in rank order (identical on every rank)
template <typename T>
__global__ void sum_rows_kernel(const T* __restrict__ in, T* __restrict__ out, int world, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  T a = in[i];
  for (int q = 1; q < world; ++q) a += in[(long long)q * n + i];
  out[i] = a;
}

}  // namespace sd

extern "C" {

int sd_peer_exchange(const void* src, int64_t bytes, void* const* peers_dev, int rank, int world, int64_t cap_bytes, int epoch,
                     void* gathered, void* stream) {
  SD_REQUIRE(src && peers_dev && gathered && bytes > 0 && bytes % 4 == 0 && bytes <= cap_bytes && cap_bytes % 16 == 0 &&
             world > 0 && world <= 64 && rank >= 0 && rank < world && epoch > 0, "sd_peer_exchange: bad arguments");
  peer_exchange_kernel<<<1, 512, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint32_t*>(src), (int)(bytes / 4),
                                                          reinterpret_cast<char* const*>(peers_dev), rank, world, (long long)cap_bytes, epoch,
                                                          reinterpret_cast<uint32_t*>(gathered));
  return check_launch("peer_exchange");
}

int sd_sum_rows(const void* in, void* out, int world, int n, int is_f64, void* stream) {
  SD_REQUIRE(in && out && world > 0 && n > 0, "sd_sum_rows: bad arguments");
  if (is_f64) sum_rows_kernel<double><<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const double*>(in), reinterpret_cast<double*>(out), world, n);
  else sum_rows_kernel<float><<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float*>(in), reinterpret_cast<float*>(out), world, n);
  return check_launch("sum_rows");
}

int sd_peer_wait_flags(const int* flags, int world, int expected, void* stream) {
  SD_REQUIRE(flags && world > 0 && world <= 64, "sd_peer_wait_flags: bad arguments");
  peer_wait_kernel<<<1, 64, 0, (cudaStream_t)stream>>>(flags, world, expected);
  return check_launch("peer_wait_flags");
}

int sd_copy_small(void* dst, const void* src, int64_t bytes, void* stream) {
  SD_REQUIRE(dst && src && bytes >= 0 && bytes % 4 == 0 && bytes <= (1 << 22) && !(((uintptr_t)dst | (uintptr_t)src) & 3),
             "sd_copy_small: up to 4 MB, 4-byte granular");
  if (bytes == 0) return 0;
  const int n = (int)(bytes / 4);
  copy_small_kernel<<<n > 4096 ? 16 : 1, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<uint32_t*>(dst), reinterpret_cast<const uint32_t*>(src), n);
  return check_launch("copy_small");
}

int sd_peer_alloc(void** ptr, int64_t bytes) {
  SD_REQUIRE(ptr != nullptr && bytes > 0, "sd_peer_alloc: bad arguments");
  SD_CUDA(cudaMalloc(ptr, (size_t)bytes));
  return 0;
}

int sd_peer_free(void* ptr) {
  if (ptr) SD_CUDA(cudaFree(ptr));
  return 0;
}

int sd_ipc_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

int sd_ipc_get_handle(void* ptr, void* handle_out) {
  SD_REQUIRE(ptr && handle_out, "sd_ipc_get_handle: null argument");
  cudaIpcMemHandle_t h;
  SD_CUDA(cudaIpcGetMemHandle(&h, ptr));
  memcpy(handle_out, &h, sizeof(h));
  return 0;
}

int sd_ipc_open_handle(const void* handle, void** ptr_out) {
  SD_REQUIRE(handle && ptr_out, "sd_ipc_open_handle: null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  SD_CUDA(cudaIpcOpenMemHandle(ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

int sd_ipc_close_handle(void* ptr) {
  if (ptr) SD_CUDA(cudaIpcCloseMemHandle(ptr));
  return 0;
}

// dst / src may live on different devices (unified addressing): copy-engine transfer, asynchronous on `stream`
int sd_memcpy_async(void* dst, const void* src, int64_t bytes, void* stream) {
  SD_REQUIRE(dst && src && bytes >= 0, "sd_memcpy_async: bad arguments");
  if (bytes == 0) return 0;
  SD_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, (cudaStream_t)stream));
  return 0;
}

}  // extern "C"
