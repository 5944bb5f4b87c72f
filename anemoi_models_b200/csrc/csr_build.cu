// dst-sorted CSR / src-sorted CSC build and the stable dst-chunk edge partition (sm_100a).
//
// Replaces the gather/scatter plan the reference rebuilds on every call inside PyG's
// MessagePassing.propagate (reference layers/conv.py:64,110) and distributed/khop_edges.py:88-130.
// All integer work; the result is bit-exact equal to a stable sort by key.
//
// Algorithm (counting sort + per-segment sort, HBM-bound integer passes):
//   1. histogram of keys (atomicAdd into cnt[nseg])
//   2. exclusive scan cnt -> ptr[nseg+1]               (tiled: 1024 items per CTA + single-CTA pass over tile sums)
//   3. fill: pos = atomicAdd(cursor[key]) ; vals[pos] = original index      (unordered inside a segment)
//   4. sort every segment of `vals` ascending -> exactly the stable order.  One warp per segment with a
//      shuffle bitonic network (<=32), one CTA per longer segment in shared memory (<=8192) or in place in
//      global memory with virtual +inf padding.
#include <climits>

#include "common.cuh"

namespace ab2 {

constexpr int kScanThreads = 256;
constexpr int kScanTile = 1024;  // 4 items per thread
constexpr int kBigSmem = 8192;   // ints sorted in shared memory by one CTA

template <int NT>
__device__ __forceinline__ int block_exclusive_scan(int val, int* total) {
  __shared__ int warp_sums[NT / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = val;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_sums[w] = inc;
  __syncthreads();
  if (w == 0) {
    int ws = lane < NT / 32 ? warp_sums[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, ws, o);
      if (lane >= o) ws += t;
    }
    if (lane < NT / 32) warp_sums[lane] = ws;  // inclusive over warps
  }
  __syncthreads();
  int excl = inc - val + (w > 0 ? warp_sums[w - 1] : 0);
  if (total) *total = warp_sums[NT / 32 - 1];
  __syncthreads();  // warp_sums is reused by the next call
  return excl;
}

__global__ void __launch_bounds__(kScanThreads) scan_tiles_kernel(const int* __restrict__ in, int* __restrict__ out,
                                                                 int* __restrict__ tile_sums, int n) {
  const int base = blockIdx.x * kScanTile + threadIdx.x * 4;
  int v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = (base + i < n) ? in[base + i] : 0;
  const int tsum = v[0] + v[1] + v[2] + v[3];
  int total;
  int excl = block_exclusive_scan<kScanThreads>(tsum, &total);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (base + i < n) out[base + i] = excl;
    excl += v[i];
  }
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) scan_tile_sums_kernel(int* __restrict__ tile_sums, int ntiles) {
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < ntiles; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < ntiles ? tile_sums[i] : 0;
    int total;
    const int excl = block_exclusive_scan<1024>(v, &total);
    const int carry = carry_s;
    if (i < ntiles) tile_sums[i] = excl + carry;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + total;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kScanThreads) scan_add_kernel(int* __restrict__ out, const int* __restrict__ tile_sums, int n) {
  const int base = blockIdx.x * kScanTile + threadIdx.x * 4;
  const int add = tile_sums[blockIdx.x];
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (base + i < n) out[base + i] += add;
}

// exclusive scan of in[0..n) into out[0..n); tile_sums needs ceil(n/1024) ints
static int exclusive_scan(const int* in, int* out, int* tile_sums, int n, cudaStream_t st) {
  if (n <= 0) return 0;
  const int ntiles = (n + kScanTile - 1) / kScanTile;
  scan_tiles_kernel<<<ntiles, kScanThreads, 0, st>>>(in, out, tile_sums, n);
  AB2_LAUNCH_OK("scan_tiles_kernel");
  if (ntiles > 1) {
    scan_tile_sums_kernel<<<1, 1024, 0, st>>>(tile_sums, ntiles);
    AB2_LAUNCH_OK("scan_tile_sums_kernel");
    scan_add_kernel<<<ntiles, kScanThreads, 0, st>>>(out, tile_sums, n);
    AB2_LAUNCH_OK("scan_add_kernel");
  }
  return 0;
}

template <typename K>
__global__ void hist_kernel(const K* __restrict__ keys, int64_t n, int nseg, int* __restrict__ cnt,
                            const K* __restrict__ other, int nother, int* __restrict__ flags) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const K key = keys[t];
    bool bad = key < 0 || key >= (K)nseg;
    if (other) {
      const K o = other[t];
      bad = bad || o < 0 || o >= (K)nother;
    }
    if (bad)
      atomicAdd(&flags[1], 1);
    else
      atomicAdd(&cnt[key], 1);
  }
}

template <typename K>
__global__ void fill_kernel(const K* __restrict__ keys, int64_t n, int nseg, int* __restrict__ cursor,
                            const K* __restrict__ other, int nother, int* __restrict__ vals) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const K key = keys[t];
    bool bad = key < 0 || key >= (K)nseg;
    if (other) {
      const K o = other[t];
      bad = bad || o < 0 || o >= (K)nother;
    }
    if (!bad) vals[atomicAdd(&cursor[key], 1)] = (int)t;
  }
}

__device__ __forceinline__ void cmpx(int& a, int& b) {
  if (a > b) {
    int t = a;
    a = b;
    b = t;
  }
}

// one warp per segment; all-ascending bitonic network ("flip" then "disperse") on 32 lanes
__global__ void __launch_bounds__(256) seg_sort_small_kernel(const int* __restrict__ ptr, int nseg, int* __restrict__ vals,
                                                             int* __restrict__ big_list, int* __restrict__ big_count) {
  const int lane = threadIdx.x & 31;
  const int nwarps = gridDim.x * (blockDim.x >> 5);
  for (int seg = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); seg < nseg; seg += nwarps) {
    const int beg = ptr[seg], len = ptr[seg + 1] - beg;
    if (len <= 1) continue;
    if (len > 32) {
      if (lane == 0) big_list[atomicAdd(big_count, 1)] = seg;
      continue;
    }
    int x = lane < len ? vals[beg + lane] : INT_MAX;
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
      {
        const int partner = lane ^ (k - 1);
        const int y = __shfl_sync(0xffffffffu, x, partner);
        x = lane < partner ? min(x, y) : max(x, y);
      }
#pragma unroll
      for (int j = k >> 2; j > 0; j >>= 1) {
        const int y = __shfl_xor_sync(0xffffffffu, x, j);
        x = (lane & j) == 0 ? min(x, y) : max(x, y);
      }
    }
    if (lane < len) vals[beg + lane] = x;
  }
}

// one CTA per long segment (grid-stride over the list the small kernel produced)
__global__ void __launch_bounds__(1024) seg_sort_big_kernel(const int* __restrict__ ptr, int* __restrict__ vals,
                                                            const int* __restrict__ big_list, const int* __restrict__ big_count) {
  __shared__ int s[kBigSmem];
  const int nbig = *big_count;
  for (int b = blockIdx.x; b < nbig; b += gridDim.x) {
    const int seg = big_list[b];
    const int beg = ptr[seg], len = ptr[seg + 1] - beg;
    int n2 = 64;
    while (n2 < len) n2 <<= 1;
    const int half = n2 >> 1;
    if (n2 <= kBigSmem) {
      for (int i = threadIdx.x; i < n2; i += blockDim.x) s[i] = i < len ? vals[beg + i] : INT_MAX;
      __syncthreads();
      for (int k = 2; k <= n2; k <<= 1) {
        const int hk = k >> 1;
        for (int idx = threadIdx.x; idx < half; idx += blockDim.x) {
          const int blk = idx / hk, o = idx - blk * hk;
          cmpx(s[blk * k + o], s[blk * k + k - 1 - o]);
        }
        __syncthreads();
        for (int j = k >> 2; j > 0; j >>= 1) {
          for (int idx = threadIdx.x; idx < half; idx += blockDim.x) {
            const int i = 2 * j * (idx / j) + (idx % j);
            cmpx(s[i], s[i + j]);
          }
          __syncthreads();
        }
      }
      for (int i = threadIdx.x; i < len; i += blockDim.x) vals[beg + i] = s[i];
      __syncthreads();
    } else {
      int* g = vals + beg;  // in place; indices >= len are virtual +inf and never move (every compare is ascending)
      for (int k = 2; k <= n2; k <<= 1) {
        const int hk = k >> 1;
        for (int idx = threadIdx.x; idx < half; idx += blockDim.x) {
          const int blk = idx / hk, o = idx - blk * hk;
          const int i = blk * k + o, p = blk * k + k - 1 - o;
          if (p < len) {
            int a = g[i], c = g[p];
            if (a > c) {
              g[i] = c;
              g[p] = a;
            }
          }
        }
        __syncthreads();
        for (int j = k >> 2; j > 0; j >>= 1) {
          for (int idx = threadIdx.x; idx < half; idx += blockDim.x) {
            const int i = 2 * j * (idx / j) + (idx % j), p = i + j;
            if (p < len) {
              int a = g[i], c = g[p];
              if (a > c) {
                g[i] = c;
                g[p] = a;
              }
            }
          }
          __syncthreads();
        }
      }
    }
  }
}

// rowidx[p] = segment id of sorted position p (one warp per segment)
__global__ void __launch_bounds__(256) expand_rowidx_kernel(const int* __restrict__ ptr, int nseg, int* __restrict__ rowidx) {
  const int lane = threadIdx.x & 31;
  const int nwarps = gridDim.x * (blockDim.x >> 5);
  for (int seg = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); seg < nseg; seg += nwarps) {
    const int beg = ptr[seg], end = ptr[seg + 1];
    for (int p = beg + lane; p < end; p += 32) rowidx[p] = seg;
  }
}

__global__ void finalize_csr_kernel(const int64_t* __restrict__ src_row, const int* __restrict__ perm, int n,
                                    int* __restrict__ col, int* __restrict__ flags) {
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
    const int t = perm[p];
    col[p] = (int)src_row[t];
    if (t != p) flags[0] = 0;
  }
}

__global__ void finalize_csc_kernel(const int* __restrict__ cpos, const int* __restrict__ rowidx, const int* __restrict__ col, int n,
                                    int* __restrict__ crow, int* __restrict__ csr2csc) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const int p = cpos[t];
    crow[2 * t] = rowidx[p];   // dst of src-sorted position t
    crow[2 * t + 1] = col[p];  // its src (lets a kernel stream CSC edges without walking colptr)
    if (csr2csc) csr2csc[p] = t;  // inverse permutation of cpos
  }
}

__global__ void init_flags_kernel(int* flags) {
  if (threadIdx.x == 0) {
    flags[0] = 1;
    flags[1] = 0;
    flags[2] = 0;
    flags[3] = 0;
  }
}

struct CsrWs {
  int* cnt;        // nmax+1
  int* tile_sums;  // ceil((nmax+1)/1024)+1
  int* big_list;   // nmax
  int* big_count;  // 1 (+3 pad)
  size_t bytes;
};
static CsrWs carve(void* ws, int64_t nmax) {
  auto al = [](size_t x) { return (x + 63) / 64 * 64; };
  CsrWs w;
  size_t off = 0;
  char* base = (char*)ws;
  w.cnt = (int*)(base + off);
  off += al((nmax + 1) * 4);
  w.tile_sums = (int*)(base + off);
  off += al(((nmax + 1 + kScanTile - 1) / kScanTile + 1) * 4);
  w.big_list = (int*)(base + off);
  off += al((nmax > 0 ? nmax : 1) * 4);
  w.big_count = (int*)(base + off);
  off += 64;
  w.bytes = off;
  return w;
}

template <typename K>
static int stable_sort_by_key(const K* keys, int64_t n, int nseg, const K* other, int nother, int* ptr, int* vals,
                              int* flags, const CsrWs& w, cudaStream_t st) {
  const int sms = num_sms();
  AB2_CUDA_OK(cudaMemsetAsync(w.cnt, 0, (size_t)(nseg + 1) * 4, st));
  AB2_CUDA_OK(cudaMemsetAsync(w.big_count, 0, 4, st));
  if (n > 0) {
    const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)sms * 16);
    hist_kernel<K><<<grid, 256, 0, st>>>(keys, n, nseg, w.cnt, other, nother, flags);
    AB2_LAUNCH_OK("hist_kernel");
  }
  if (int rc = exclusive_scan(w.cnt, ptr, w.tile_sums, nseg + 1, st)) return rc;
  if (n > 0 && nseg > 0) {
    AB2_CUDA_OK(cudaMemcpyAsync(w.cnt, ptr, (size_t)nseg * 4, cudaMemcpyDeviceToDevice, st));
    const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)sms * 16);
    fill_kernel<K><<<grid, 256, 0, st>>>(keys, n, nseg, w.cnt, other, nother, vals);
    AB2_LAUNCH_OK("fill_kernel");
    const int wgrid = (int)std::min<int64_t>(((int64_t)nseg + 7) / 8, (int64_t)sms * 8);
    seg_sort_small_kernel<<<wgrid, 256, 0, st>>>(ptr, nseg, vals, w.big_list, w.big_count);
    AB2_LAUNCH_OK("seg_sort_small_kernel");
    seg_sort_big_kernel<<<sms, 1024, 0, st>>>(ptr, vals, w.big_list, w.big_count);
    AB2_LAUNCH_OK("seg_sort_big_kernel");
  }
  return 0;
}

// ---- dst-chunk partition (khop_edges.py:88-130) -------------------------------------------------------
__global__ void chunk_flag_kernel(const int64_t* __restrict__ dst, int64_t n, int64_t lo, int64_t hi, int* __restrict__ flag) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t <= n; t += (int64_t)gridDim.x * blockDim.x) {
    int f = 0;
    if (t < n) {
      const int64_t d = dst[t];
      f = (d >= lo && d < hi) ? 1 : 0;
    }
    flag[t] = f;
  }
}
__global__ void chunk_compact_kernel(const int64_t* __restrict__ dst, int64_t n, int64_t lo, int64_t hi,
                                     const int* __restrict__ pos, int64_t* __restrict__ order, const int64_t* __restrict__ base_ptr,
                                     int64_t* __restrict__ count_out, int64_t* __restrict__ next_base) {
  const int64_t base = *base_ptr;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t d = dst[t];
    if (d >= lo && d < hi) order[base + pos[t]] = t;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    *count_out = pos[n];
    *next_base = base + pos[n];
  }
}

}  // namespace ab2

using namespace ab2;

extern "C" size_t ab2_csr_workspace_bytes(int64_t E, int64_t Ns, int64_t Nd) {
  (void)E;
  return carve(nullptr, std::max<int64_t>(std::max(Ns, Nd), 1)).bytes;
}

extern "C" int ab2_csr_build(const int64_t* edge_index, int64_t E, int64_t Ns, int64_t Nd, int32_t* rowptr, int32_t* col,
                             int32_t* perm, int32_t* rowidx, int32_t* colptr, int32_t* cpos, int32_t* crow, int32_t* csr2csc,
                             int32_t* flags, void* workspace, size_t workspace_bytes, void* stream) {
  if (E < 0 || Ns < 0 || Nd < 0 || E >= INT_MAX || Ns >= INT_MAX || Nd >= INT_MAX)
    return fail(AB2_ERR_UNSUPPORTED, "csr_build: E, Ns, Nd must be in [0, 2^31-1) (got %lld, %lld, %lld)", (long long)E, (long long)Ns, (long long)Nd);
  if (!rowptr || !flags || !workspace || (E > 0 && (!edge_index || !col || !perm || !rowidx)))
    return fail(AB2_ERR_INVALID, "csr_build: null pointer argument");
  const bool want_csc = colptr != nullptr;
  if (want_csc && E > 0 && (!cpos || !crow)) return fail(AB2_ERR_INVALID, "csr_build: colptr given without cpos/crow");
  if (workspace_bytes < ab2_csr_workspace_bytes(E, Ns, Nd)) return fail(AB2_ERR_INVALID, "csr_build: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const CsrWs w = carve(workspace, std::max<int64_t>(std::max(Ns, Nd), 1));
  const int sms = num_sms();
  init_flags_kernel<<<1, 32, 0, st>>>(flags);
  AB2_LAUNCH_OK("init_flags_kernel");
  // pass 1: stable sort of edge ids by dst
  if (int rc = stable_sort_by_key<int64_t>(edge_index + E, E, (int)Nd, edge_index, (int)Ns, rowptr, perm, flags, w, st)) return rc;
  if (E > 0) {
    const int grid = (int)std::min<int64_t>((E + 255) / 256, (int64_t)sms * 16);
    finalize_csr_kernel<<<grid, 256, 0, st>>>(edge_index, perm, (int)E, col, flags);
    AB2_LAUNCH_OK("finalize_csr_kernel");
    if (Nd > 0) {
      const int wgrid = (int)std::min<int64_t>((Nd + 7) / 8, (int64_t)sms * 8);
      expand_rowidx_kernel<<<wgrid, 256, 0, st>>>(rowptr, (int)Nd, rowidx);
      AB2_LAUNCH_OK("expand_rowidx_kernel");
    }
  }
  // pass 2: stable sort of CSR positions by src
  if (want_csc) {
    if (int rc = stable_sort_by_key<int32_t>(col, E, (int)Ns, nullptr, 0, colptr, cpos, flags + 2, w, st)) return rc;
    if (E > 0) {
      const int grid = (int)std::min<int64_t>((E + 255) / 256, (int64_t)sms * 16);
      finalize_csc_kernel<<<grid, 256, 0, st>>>(cpos, rowidx, col, (int)E, crow, csr2csc);
      AB2_LAUNCH_OK("finalize_csc_kernel");
    }
  }
  return AB2_OK;
}

extern "C" size_t ab2_edge_chunks_workspace_bytes(int64_t E) {
  auto al = [](size_t x) { return (x + 63) / 64 * 64; };
  return al((E + 1) * 4) * 2 + al(((E + 1 + kScanTile - 1) / kScanTile + 1) * 4) + 64;
}

extern "C" int ab2_edge_chunks(const int64_t* edge_index, int64_t E, const int64_t* bounds_host, int num_chunks,
                               int64_t* order, int64_t* counts, void* workspace, size_t workspace_bytes, void* stream) {
  if (E < 0 || E >= INT_MAX - 1 || num_chunks < 1) return fail(AB2_ERR_INVALID, "edge_chunks: bad E or num_chunks");
  if (!bounds_host || !counts || !workspace || (E > 0 && (!edge_index || !order))) return fail(AB2_ERR_INVALID, "edge_chunks: null pointer argument");
  if (workspace_bytes < ab2_edge_chunks_workspace_bytes(E)) return fail(AB2_ERR_INVALID, "edge_chunks: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  auto al = [](size_t x) { return (x + 63) / 64 * 64; };
  char* base = (char*)workspace;
  int* flag = (int*)base;
  int* pos = (int*)(base + al((E + 1) * 4));
  int* tile_sums = (int*)(base + 2 * al((E + 1) * 4));
  int64_t* run = (int64_t*)(base + 2 * al((E + 1) * 4) + al(((E + 1 + kScanTile - 1) / kScanTile + 1) * 4));  // [0]=base of this chunk, [1]=next
  AB2_CUDA_OK(cudaMemsetAsync(run, 0, 64, st));
  const int grid = (int)std::min<int64_t>((E + 256) / 256, (int64_t)num_sms() * 16);
  const int64_t* dst = edge_index + E;
  for (int c = 0; c < num_chunks; ++c) {
    const int64_t lo = bounds_host[c], hi = bounds_host[c + 1];
    chunk_flag_kernel<<<grid, 256, 0, st>>>(dst, E, lo, hi, flag);
    AB2_LAUNCH_OK("chunk_flag_kernel");
    if (int rc = exclusive_scan(flag, pos, tile_sums, (int)(E + 1), st)) return rc;
    // run[c&1] holds this chunk's base offset, run[(c+1)&1] receives the next one
    chunk_compact_kernel<<<grid, 256, 0, st>>>(dst, E, lo, hi, pos, order, run + (c & 1), counts + c, run + ((c + 1) & 1));
    AB2_LAUNCH_OK("chunk_compact_kernel");
  }
  return AB2_OK;
}
