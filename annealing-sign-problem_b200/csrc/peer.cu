// NVLink peer memory between the one-process-per-GPU ranks of a node (exchange X1).
//
// The reference is single-process; its "exchange" is that every caller holds the whole basis
// (annealing_sign_problem/common.py:146).  Sharded over GPUs, every rank must see all row blocks
// of the sorted basis.  Instead of an NCCL all-gather followed by an index pass, the blocks stay in
// peer-mapped buffers and ONE kernel (gather_index_tma_kernel, exchange_kernels.cuh) pulls them over
// NVLink while it indexes them.  This file holds the plumbing:
//
//   asp_peer_alloc / asp_peer_open / asp_peer_close / asp_peer_free
//       device memory other processes can map (cudaMalloc + CUDA IPC handle, 64 bytes that travel
//       through any host channel, e.g. torch.distributed.all_gather_object);
//   asp_peer_signal / asp_peer_wait
//       epoch flags in peer memory: a rank publishes "my block is complete for epoch e" (or "I have
//       finished reading epoch e") into every peer's flag array with a system-scope release store
//       that is stream-ordered after its earlier work; the consumer spins with system-scope acquire
//       loads.  No host round trip, no NCCL call on the data path.
#include "common.cuh"

namespace asp {

struct PeerFlags {
  unsigned long long *array[16];  // every rank's flag array (own memory or IPC-mapped)
};

__global__ void peer_signal_kernel(const PeerFlags flags, int world, unsigned slot, unsigned long long value) {
  const int p = threadIdx.x;
  if (p >= world || flags.array[p] == nullptr) return;
  __threadfence_system();  // everything this stream wrote before is visible before the flag
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flags.array[p] + slot), "l"(value) : "memory");
}

__global__ void peer_wait_kernel(const unsigned long long *flags, int world, unsigned long long value) {
  const int q = threadIdx.x;
  if (q >= world) return;
  unsigned long long v, t0, t1;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
  unsigned ns = 64;
  for (;;) {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + q) : "memory");
    if (v >= value) break;
    __nanosleep(ns);
    if (ns < 2048) ns <<= 1;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
    if (t1 - t0 > 10000000000ull) __trap();  // a dead peer must not hang the device
  }
}

}  // namespace asp

using namespace asp;

extern "C" {

int asp_peer_alloc(size_t bytes, void **d_ptr, unsigned char *handle) {
  ASP_REQUIRE(d_ptr != nullptr && handle != nullptr && bytes > 0, "asp_peer_alloc: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
  void *p = nullptr;
  ASP_CUDA_CHECK(cudaMalloc(&p, bytes));
  cudaError_t e = cudaMemset(p, 0, bytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    set_error("asp_peer_alloc: %s", cudaGetErrorString(e));
    return ASP_ERR_CUDA;
  }
  std::memcpy(handle, &h, sizeof(h));
  *d_ptr = p;
  return ASP_OK;
}

int asp_peer_open(unsigned char const *handle, void **d_ptr) {
  ASP_REQUIRE(d_ptr != nullptr && handle != nullptr, "asp_peer_open: bad argument");
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, sizeof(h));
  ASP_CUDA_CHECK(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return ASP_OK;
}

int asp_peer_close(void *d_ptr) {
  if (d_ptr) ASP_CUDA_CHECK(cudaIpcCloseMemHandle(d_ptr));
  return ASP_OK;
}

int asp_peer_free(void *d_ptr) {
  if (d_ptr) ASP_CUDA_CHECK(cudaFree(d_ptr));
  return ASP_OK;
}

int asp_peer_signal(uint32_t world, uint64_t *const *d_flags, uint32_t slot, uint64_t value, void *stream) {
  ASP_REQUIRE(world >= 1 && world <= 16 && d_flags != nullptr, "asp_peer_signal: world size must be in 1..16");
  PeerFlags f{};
  for (uint32_t p = 0; p < world; ++p) f.array[p] = reinterpret_cast<unsigned long long *>(d_flags[p]);
  peer_signal_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(f, static_cast<int>(world), slot, value);
  ASP_LAUNCH_CHECK();
  return ASP_OK;
}

int asp_peer_wait(uint32_t world, uint64_t const *d_flags, uint64_t value, void *stream) {
  ASP_REQUIRE(world >= 1 && world <= 16 && d_flags != nullptr, "asp_peer_wait: world size must be in 1..16");
  peer_wait_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const unsigned long long *>(d_flags),
                                                                    static_cast<int>(world), value);
  ASP_LAUNCH_CHECK();
  return ASP_OK;
}

}  // extern "C"
