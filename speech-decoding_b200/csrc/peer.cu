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

}  // namespace sd

extern "C" {

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
