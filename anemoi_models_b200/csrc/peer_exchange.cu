// Halo exchange over NVLink peer memory (one process per GPU, one node): rows are PUSHED straight into the consumer
// rank's buffer with plain stores through a CUDA-IPC mapping, by one kernel that uses every SM.
//
// Replaces the NCCL all-to-all of distributed/halo.py on NVLink-connected ranks.  Measured reason (profiles/r01, run r01l/m):
// a dst-row-sharded graph sends almost all of a rank's halo to ONE neighbour (latitude bands), and a single NCCL
// send/recv pair runs on a couple of channels (~50-100 GB/s), so the all-to-all cost 1.3 ms per 100 MB tensor while
// NVLink 5 moves it in ~0.15 ms.
#include <cstdlib>
#include <cstring>

#include "common.cuh"

using namespace ab2;

extern "C" int ab2_ipc_alloc(size_t bytes, void** dev_ptr, void* handle_out) {
  if (!dev_ptr || !handle_out || bytes == 0) return fail(AB2_ERR_INVALID, "ipc_alloc: bad argument");
  AB2_CUDA_OK(cudaMalloc(dev_ptr, bytes));
  AB2_CUDA_OK(cudaMemset(*dev_ptr, 0, bytes));  // the flag words of the exchange protocol start at epoch 0
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, *dev_ptr);
  if (e != cudaSuccess) {
    cudaFree(*dev_ptr);
    *dev_ptr = nullptr;
    return fail(AB2_ERR_CUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
  memcpy(handle_out, &h, 64);
  return AB2_OK;
}

extern "C" int ab2_ipc_open(const void* handle, void** dev_ptr) {
  if (!handle || !dev_ptr) return fail(AB2_ERR_INVALID, "ipc_open: bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  AB2_CUDA_OK(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return AB2_OK;
}

extern "C" int ab2_ipc_close(void* dev_ptr) {
  if (dev_ptr) AB2_CUDA_OK(cudaIpcCloseMemHandle(dev_ptr));
  return AB2_OK;
}

extern "C" int ab2_ipc_free(void* dev_ptr) {
  if (dev_ptr) AB2_CUDA_OK(cudaFree(dev_ptr));
  return AB2_OK;
}

extern "C" int ab2_memcpy_d2d(void* dst, const void* src, size_t bytes, void* stream) {
  if (bytes == 0) return AB2_OK;
  if (!dst || !src) return fail(AB2_ERR_INVALID, "memcpy_d2d: null pointer");
  AB2_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return AB2_OK;
}

namespace ab2 {

struct PeerTable {
  char* a[AB2_MAX_PEERS];  // base of plane A in every rank's buffer (mapped through IPC; own entry = local pointer)
  char* b[AB2_MAX_PEERS];  // base of plane B
};

// row r of the local planes (src_a, src_b; row index src_row[r] or r) -> row dst_row[r] of rank peer[r]'s planes.
// One warp per row, 16 bytes per lane per step; stores to peer memory are posted writes over NVLink.
__global__ void __launch_bounds__(256)
peer_push_rows_kernel(const char* __restrict__ src_a, const char* __restrict__ src_b, const int* __restrict__ src_row,
                      const int* __restrict__ peer, const int* __restrict__ dst_row, long long n, int row_bytes, PeerTable tab) {
  const int lane = threadIdx.x & 31;
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n; r += nwarps) {
    const size_t so = (size_t)(src_row ? src_row[r] : (int)r) * row_bytes;
    const int p = peer[r];
    const size_t dof = (size_t)dst_row[r] * row_bytes;
    for (int off = lane * 16; off < row_bytes; off += 32 * 16) {
      if (src_a) {
        const uint4 x = ldg16_keep(src_a + so + off);
        *reinterpret_cast<uint4*>(tab.a[p] + dof + off) = x;
      }
      if (src_b) {
        const uint4 y = ldg16_keep(src_b + so + off);
        *reinterpret_cast<uint4*>(tab.b[p] + dof + off) = y;
      }
    }
  }
}

// ---- flag-word barrier (replaces the 1-element NCCL all-reduce that followed every push in round 1) ------------------------
// Every rank owns, inside its IPC buffer, one 32-bit slot per peer.  After the LAST CTA of a push kernel has seen all of the
// kernel's stores fenced at system scope, it stores the exchange's epoch into its own slot on every peer (st.release.sys: the
// rows it pushed are visible before the flag is).  The consumer runs a one-warp kernel that spins (ld.acquire.sys) until the
// slots of all its peers have reached the epoch; stream order then makes the rows visible to the kernels behind it.
// All-to-all signalling (also to peers that receive no rows) keeps the double-buffered planes safe: a rank can only be one
// epoch ahead of the slowest peer.
struct FlagTable {
  unsigned* slot[AB2_MAX_PEERS];  // address of MY slot in every peer's flag array (own entry unused)
};

__device__ __forceinline__ void signal_peers(unsigned* counter, const FlagTable& flags, unsigned epoch, int npeers, int me) {
  __threadfence_system();  // this thread's stores to peer memory
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned ticket = atomicAdd(counter, 1u);
    if (ticket == gridDim.x - 1) {  // every CTA of the launch has fenced its stores
      *counter = 0u;
      __threadfence_system();
      for (int p = 0; p < npeers; ++p)
        if (p != me) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flags.slot[p]), "r"(epoch) : "memory");
    }
  }
}

__global__ void __launch_bounds__(256)
peer_push_rows_signal_kernel(const char* __restrict__ src_a, const char* __restrict__ src_b, const int* __restrict__ src_row,
                             const int* __restrict__ peer, const int* __restrict__ dst_row, long long n, int row_bytes, PeerTable tab,
                             unsigned* counter, FlagTable flags, unsigned epoch, int npeers, int me) {
  const int lane = threadIdx.x & 31;
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n; r += nwarps) {
    const size_t so = (size_t)(src_row ? src_row[r] : (int)r) * row_bytes;
    const int p = peer[r];
    const size_t dof = (size_t)dst_row[r] * row_bytes;
    for (int off = lane * 16; off < row_bytes; off += 32 * 16) {
      if (src_a) *reinterpret_cast<uint4*>(tab.a[p] + dof + off) = ldg16_keep(src_a + so + off);
      if (src_b) *reinterpret_cast<uint4*>(tab.b[p] + dof + off) = ldg16_keep(src_b + so + off);
    }
  }
  signal_peers(counter, flags, epoch, npeers, me);
}

__global__ void peer_wait_flags_kernel(const unsigned* __restrict__ slots, unsigned epoch, int npeers, int me) {
  const int p = threadIdx.x;
  if (p >= npeers || p == me) return;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(slots + p) : "memory");
    if ((int)(v - epoch) >= 0) break;
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > 20000000000ull) __trap();  // 20 s: a lost peer must end in an error, not in a hung GPU
    __nanosleep(64);
  }
}

// dst[idx[s]] += src[s] for s in [0, n): the ids of one call are distinct, so plain read-modify-write (fp32 add)
template <typename T>
__global__ void __launch_bounds__(256)
rows_add_kernel(T* __restrict__ dst, const long long* __restrict__ idx, const T* __restrict__ src, long long n, int chunks) {
  constexpr int VEC = Vec<T>::N;
  const long long total = n * chunks;
  for (long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x; w < total; w += (long long)gridDim.x * blockDim.x) {
    const long long s = w / chunks;
    const size_t off = (size_t)(w - s * chunks) * VEC;
    const size_t D = (size_t)chunks * VEC;
    T* d = dst + (size_t)idx[s] * D + off;
    float a[VEC], b[VEC];
    unpack<T>(*reinterpret_cast<const uint4*>(d), a);
    unpack<T>(ldg16(src + (size_t)s * D + off), b);
#pragma unroll
    for (int i = 0; i < VEC; ++i) a[i] += b[i];
    *reinterpret_cast<uint4*>(d) = pack<T>(a);
  }
}

}  // namespace ab2

extern "C" int ab2_peer_push_rows(const void* src_a, const void* src_b, const int32_t* src_row, const int32_t* peer,
                                  const int32_t* dst_row, int64_t n, int row_bytes, void* const* plane_a, void* const* plane_b,
                                  int npeers, void* stream) {
  if (n == 0) return AB2_OK;
  if (!peer || !dst_row || !plane_a || npeers < 1 || npeers > AB2_MAX_PEERS || row_bytes <= 0 || row_bytes % 16 != 0)
    return fail(AB2_ERR_INVALID, "peer_push_rows: bad argument (row_bytes must be a multiple of 16, npeers <= %d)", AB2_MAX_PEERS);
  PeerTable tab{};
  for (int p = 0; p < npeers; ++p) {
    tab.a[p] = (char*)plane_a[p];
    tab.b[p] = plane_b ? (char*)plane_b[p] : nullptr;
  }
  // NVLink-bound: a few dozen CTAs of posted stores saturate the link; keep the other SMs for the conv kernels that run
  // concurrently on the main stream (AB2_PUSH_CTAS overrides, for experiments)
  static const int max_ctas = [] {
    const char* s = getenv("AB2_PUSH_CTAS");
    return s ? std::max(1, atoi(s)) : 64;
  }();
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + 7) / 8, (int64_t)max_ctas));
  peer_push_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const char*)src_a, plane_b ? (const char*)src_b : nullptr, src_row, peer,
                                                               dst_row, n, row_bytes, tab);
  AB2_LAUNCH_OK("peer_push_rows_kernel");
  return AB2_OK;
}

extern "C" int ab2_peer_push_rows_signal(const void* src_a, const void* src_b, const int32_t* src_row, const int32_t* peer,
                                         const int32_t* dst_row, int64_t n, int row_bytes, void* const* plane_a, void* const* plane_b,
                                         void* counter, void* const* flag_slots, uint32_t epoch, int npeers, int my_rank, void* stream) {
  if (!plane_a || !counter || !flag_slots || npeers < 1 || npeers > AB2_MAX_PEERS || my_rank < 0 || my_rank >= npeers || row_bytes <= 0 ||
      row_bytes % 16 != 0 || (n > 0 && (!peer || !dst_row)))
    return fail(AB2_ERR_INVALID, "peer_push_rows_signal: bad argument (row_bytes must be a multiple of 16, npeers <= %d)", AB2_MAX_PEERS);
  PeerTable tab{};
  FlagTable flags{};
  for (int p = 0; p < npeers; ++p) {
    tab.a[p] = (char*)plane_a[p];
    tab.b[p] = plane_b ? (char*)plane_b[p] : nullptr;
    flags.slot[p] = (unsigned*)flag_slots[p];
  }
  static const int max_ctas = [] {
    const char* s = getenv("AB2_PUSH_CTAS");
    return s ? std::max(1, atoi(s)) : 64;
  }();
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + 7) / 8, (int64_t)max_ctas));
  peer_push_rows_signal_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const char*)src_a, plane_b ? (const char*)src_b : nullptr, src_row,
                                                                      peer, dst_row, n, row_bytes, tab, (unsigned*)counter, flags, epoch,
                                                                      npeers, my_rank);
  AB2_LAUNCH_OK("peer_push_rows_signal_kernel");
  return AB2_OK;
}

extern "C" int ab2_peer_wait_flags(const void* local_slots, uint32_t epoch, int npeers, int my_rank, void* stream) {
  if (!local_slots || npeers < 1 || npeers > AB2_MAX_PEERS) return fail(AB2_ERR_INVALID, "peer_wait_flags: bad argument");
  if (npeers == 1) return AB2_OK;
  peer_wait_flags_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((const unsigned*)local_slots, epoch, npeers, my_rank);
  AB2_LAUNCH_OK("peer_wait_flags_kernel");
  return AB2_OK;
}

extern "C" int ab2_rows_add(void* dst, const int64_t* idx, const void* src, int64_t n, int D, int dtype, void* stream) {
  if (n == 0) return AB2_OK;
  if (!dst || !idx || !src || D <= 0) return fail(AB2_ERR_INVALID, "rows_add: bad argument");
  const int elt = dtype == AB2_F32 ? 4 : 2;
  if (dtype != AB2_F32 && dtype != AB2_BF16) return fail(AB2_ERR_INVALID, "rows_add: bad dtype");
  if ((D * elt) % 16 != 0) return fail(AB2_ERR_UNSUPPORTED, "rows_add: row bytes must be a multiple of 16");
  const int chunks = D * elt / 16;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n * chunks + 255) / 256, (int64_t)num_sms() * 16));
  if (dtype == AB2_F32)
    rows_add_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((float*)dst, (const long long*)idx, (const float*)src, n, chunks);
  else
    rows_add_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)dst, (const long long*)idx,
                                                                          (const __nv_bfloat16*)src, n, chunks);
  AB2_LAUNCH_OK("rows_add_kernel");
  return AB2_OK;
}
