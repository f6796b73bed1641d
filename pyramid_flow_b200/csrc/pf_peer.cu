// pf_peer.cu — peer memory over NVLink / NVSwitch for the sequence-parallel DiT step (one process per GPU).
//
// The reference moves q/k/v and the attention output between sequence-parallel ranks with list all-to-alls plus
// contiguous()/cat copies on both sides of every attention (trainer_misc/communicate.py:7-24, modeling_flux_block.py:285-321).
// Here the exchange is fused into the producing kernels: the QKV GEMM epilogue stores each head's rows straight into the
// owning rank's buffer through a mapped peer pointer (pf_gemm_desc.peer_*), and the attention epilogue stores each token
// chunk's output straight into the owning rank's `cat` buffer (pf_attn_desc.peer_*).  What is left of the collective is a
// flag barrier (one tiny kernel, below) that orders those remote stores against their consumers, so the whole parallel step is
// ordinary stream-ordered kernels: CUDA-graph capturable, no NCCL call and no copy kernel on the hot path.
//
// Memory comes from cudaMalloc (one allocation per buffer: the CUDA IPC handle then maps the buffer at offset 0 in the peer
// process); handles travel through the host side's torch.distributed object gather (sp.py), once at start-up.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/pf_b200.h"
#include "pf_common.cuh"

namespace pf {

// grp.ptr[i] = member i's flag array (uint32 per member).  One thread per group member.  epoch_counter lives in this rank's memory and is advanced by the kernel itself, so the same
// launch replays correctly inside a CUDA graph.  Every rank runs the same sequence of barriers, so epochs agree.
__global__ void peer_barrier_kernel(PfPeerGroup grp, uint32_t* epoch_counter) {
  const int i = threadIdx.x;
  __shared__ uint32_t s_epoch;
  if (i == 0) s_epoch = *epoch_counter + 1;
  __syncthreads();
  const uint32_t e = s_epoch;
  if (i < grp.n) {
    // all earlier kernels of this stream have completed (kernel boundary); make their remote stores visible system-wide
    // before the flag that announces them
    __threadfence_system();
    uint32_t* remote = reinterpret_cast<uint32_t*>(grp.ptr[i]) + grp.my_index;     // my slot in member i's flag array
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(e) : "memory");
    const uint32_t* mine = reinterpret_cast<const uint32_t*>(grp.ptr[grp.my_index]) + i;   // member i's slot in mine
    uint32_t v;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
    } while (static_cast<int32_t>(v - e) < 0);
  }
  __syncthreads();
  if (i == 0) *epoch_counter = e;
}

// dst_i[dst_offset + k] = src[k] for every group member i (16-byte words): publishes a small result to all peers.
__global__ void peer_bcast_kernel(PfPeerGroup grp, const uint4* src, long long n16, long long dst_offset16) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long k = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; k < n16; k += stride) {
    const uint4 v = src[k];
    for (int i = 0; i < grp.n; ++i) reinterpret_cast<uint4*>(grp.ptr[i])[dst_offset16 + k] = v;
  }
}

}  // namespace pf

extern "C" {

int pf_peer_alloc(int64_t bytes, void** ptr) {
  using namespace pf;
  PF_REQUIRE(ptr != nullptr && bytes > 0, "pf_peer_alloc: bad arguments");
  cudaError_t e = cudaMalloc(ptr, static_cast<size_t>(bytes));
  if (e == cudaSuccess) e = cudaMemset(*ptr, 0, static_cast<size_t>(bytes));
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    set_error("pf_peer_alloc(%lld): %s", static_cast<long long>(bytes), cudaGetErrorString(e));
    return -2;
  }
  return 0;
}

int pf_peer_free(void* ptr) {
  cudaError_t e = cudaFree(ptr);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    pf::set_error("pf_peer_free: %s", cudaGetErrorString(e));
    return -2;
  }
  return 0;
}

int pf_peer_export(void* ptr, void* handle64) {
  using namespace pf;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  PF_REQUIRE(ptr && handle64, "pf_peer_export: null");
  cudaError_t e = cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), ptr);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    set_error("cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
    return -2;
  }
  return 0;
}

int pf_peer_open(const void* handle64, void** peer_ptr) {
  using namespace pf;
  PF_REQUIRE(handle64 && peer_ptr, "pf_peer_open: null");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  cudaError_t e = cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    set_error("cudaIpcOpenMemHandle: %s (peer access over NVLink is required: one process per GPU on one node)",
              cudaGetErrorString(e));
    return -2;
  }
  return 0;
}

int pf_peer_close(void* peer_ptr) {
  cudaError_t e = cudaIpcCloseMemHandle(peer_ptr);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    pf::set_error("cudaIpcCloseMemHandle: %s", cudaGetErrorString(e));
    return -2;
  }
  return 0;
}

int pf_peer_barrier(const PfPeerGroup* grp, uint32_t* epoch_counter, void* stream) {
  using namespace pf;
  PF_REQUIRE(grp && epoch_counter && grp->n >= 1 && grp->n <= PF_MAX_PEERS && grp->my_index >= 0 && grp->my_index < grp->n,
             "pf_peer_barrier: bad group");
  peer_barrier_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(*grp, epoch_counter);
  return check_launch("pf_peer_barrier");
}

int pf_peer_bcast(const PfPeerGroup* dst, const void* src, int64_t bytes, int64_t dst_offset_bytes, void* stream) {
  using namespace pf;
  PF_REQUIRE(dst && src && dst->n >= 1 && dst->n <= PF_MAX_PEERS, "pf_peer_bcast: bad group");
  PF_REQUIRE(bytes > 0 && bytes % 16 == 0 && dst_offset_bytes % 16 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0,
             "pf_peer_bcast: sizes and offsets must be multiples of 16 bytes");
  const long long n16 = bytes / 16;
  int blocks = static_cast<int>((n16 + 255) / 256);
  if (blocks > 296) blocks = 296;
  peer_bcast_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(*dst, static_cast<const uint4*>(src), n16,
                                                                             dst_offset_bytes / 16);
  return check_launch("pf_peer_bcast");
}

}  // extern "C"
