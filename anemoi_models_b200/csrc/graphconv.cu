// GraphConv edge path: the non-GEMM pieces, fused (sm_100a).
//
// Reference layers/conv.py:61-76: edges_new = edge_mlp(cat[x_i, x_j, e]) + e ; out = scatter_sum(edges_new, dst).
// The three GEMMs of edge_mlp stay GEMMs (tensor cores); what is fused here is everything around them that the
// reference materialises in HBM: the [E,3D] gather+concat (replaced by a split first layer: pi[dst]+pj[src]+pe),
// the activation, and LayerNorm + residual + scatter-sum (one pass over the dst-sorted CSR, no atomics).
// All kernels are HBM streaming kernels: 16-byte vector loads, fp32 math.
#include <cmath>

#include "common.cuh"

namespace ab2 {

// ---- chunk load/store: N elements of T (N = Vec<T>::N -> one 16-byte access, N = 1 -> scalar) -----------
template <typename T, int N>
__device__ __forceinline__ void load_chunk(const T* p, float (&f)[N]) {
  if constexpr (N == 1) {
    f[0] = to_f<T>(*p);
  } else {
    unpack<T>(ldg16(p), f);
  }
}
template <typename T, int N>
__device__ __forceinline__ void load_chunk_keep(const T* p, float (&f)[N]) {
  if constexpr (N == 1) {
    f[0] = to_f<T>(*p);
  } else {
    unpack<T>(ldg16_keep(p), f);
  }
}
template <typename T, int N>
__device__ __forceinline__ void store_chunk(T* p, const float (&f)[N]) {
  if constexpr (N == 1) {
    *p = from_f<T>(f[0]);
  } else {
    stg16(p, pack<T>(f));
  }
}

template <int ACT>
__device__ __forceinline__ float act_fwd(float x) {
  if constexpr (ACT == 0) return x / (1.f + __expf(-x));                              // SiLU
  if constexpr (ACT == 1) return 0.5f * x * (1.f + erff(x * 0.70710678118654752f));   // GELU (erf form, nn.GELU default)
  if constexpr (ACT == 2) return fmaxf(x, 0.f);                                       // ReLU
  return x;
}
template <int ACT>
__device__ __forceinline__ float act_bwd(float x) {
  if constexpr (ACT == 0) {
    const float s = 1.f / (1.f + __expf(-x));
    return s * (1.f + x * (1.f - s));
  }
  if constexpr (ACT == 1) {
    const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752f));
    return cdf + x * 0.3989422804014327f * __expf(-0.5f * x * x);
  }
  if constexpr (ACT == 2) return x > 0.f ? 1.f : 0.f;
  return 1.f;
}

// ---- h0[t] = act(pi[dst_t] + pj[src_t] + pe[t]) ---------------------------------------------------------
template <typename T, int N, int ACT>
__global__ void __launch_bounds__(256)
edge_gather_add_act_kernel(const T* __restrict__ pi, const T* __restrict__ pj, const T* __restrict__ pe,
                           const int64_t* __restrict__ ei, int64_t E, int chunks, T* __restrict__ h0, T* __restrict__ pre) {
  const size_t D = (size_t)chunks * N;
  const int64_t total = E * chunks;
  for (int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; w < total; w += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = w / chunks;
    const size_t off = (size_t)(w - t * chunks) * N;
    const size_t j = (size_t)ei[t], i = (size_t)ei[E + t];
    float a[N], b[N], c[N];
    load_chunk_keep<T, N>(pi + i * D + off, a);
    load_chunk_keep<T, N>(pj + j * D + off, b);
    load_chunk<T, N>(pe + (size_t)t * D + off, c);
#pragma unroll
    for (int x = 0; x < N; ++x) a[x] = a[x] + b[x] + c[x];
    if (pre) store_chunk<T, N>(pre + (size_t)t * D + off, a);
#pragma unroll
    for (int x = 0; x < N; ++x) a[x] = act_fwd<ACT>(a[x]);
    store_chunk<T, N>(h0 + (size_t)t * D + off, a);
  }
}

// ---- backward: gpe[t] = g[t]*act'(pre[t]) and dpi[i] = segment sum over incoming edges --------------------
// one thread per (dst row, chunk); edges of the row are visited in CSR order.
template <typename T, int N, int ACT>
__global__ void __launch_bounds__(256)
edge_act_bwd_dst_kernel(const T* __restrict__ g, const T* __restrict__ pre, const int* __restrict__ rowptr,
                        const int* __restrict__ perm, int Nd, int chunks, T* __restrict__ gpe, T* __restrict__ dpi) {
  const size_t D = (size_t)chunks * N;
  const int64_t total = (int64_t)Nd * chunks;
  for (int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; w < total; w += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(w / chunks);
    const size_t off = (size_t)(w - (int64_t)i * chunks) * N;
    float acc[N];
#pragma unroll
    for (int x = 0; x < N; ++x) acc[x] = 0.f;
    const int beg = rowptr[i], end = rowptr[i + 1];
    for (int p = beg; p < end; ++p) {
      const size_t t = (size_t)perm[p];
      float gg[N], pp[N];
      load_chunk<T, N>(g + t * D + off, gg);
      load_chunk<T, N>(pre + t * D + off, pp);
#pragma unroll
      for (int x = 0; x < N; ++x) {
        gg[x] *= act_bwd<ACT>(pp[x]);
        acc[x] += gg[x];
      }
      store_chunk<T, N>(gpe + t * D + off, gg);
    }
    if (dpi) store_chunk<T, N>(dpi + (size_t)i * D + off, acc);
  }
}
// dpj[j] = sum of gpe over outgoing edges (CSC order)
template <typename T, int N>
__global__ void __launch_bounds__(256)
edge_act_bwd_src_kernel(const T* __restrict__ gpe, const int* __restrict__ colptr, const int* __restrict__ cpos,
                        const int* __restrict__ perm, int Ns, int chunks, T* __restrict__ dpj) {
  const size_t D = (size_t)chunks * N;
  const int64_t total = (int64_t)Ns * chunks;
  for (int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; w < total; w += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(w / chunks);
    const size_t off = (size_t)(w - (int64_t)j * chunks) * N;
    float acc[N];
#pragma unroll
    for (int x = 0; x < N; ++x) acc[x] = 0.f;
    const int beg = colptr[j], end = colptr[j + 1];
    for (int s = beg; s < end; ++s) {
      const size_t t = (size_t)perm[cpos[s]];
      float gg[N];
      load_chunk_keep<T, N>(gpe + t * D + off, gg);
#pragma unroll
      for (int x = 0; x < N; ++x) acc[x] += gg[x];
    }
    store_chunk<T, N>(dpj + (size_t)j * D + off, acc);
  }
}

// out[i] = sum of g[perm[idx ? idx[s] : s]] over s in [ptr[i], ptr[i+1])   (dst segments: idx = null; src segments: idx = cpos)
template <typename T, int N>
__global__ void __launch_bounds__(256)
edge_segsum_kernel(const T* __restrict__ g, const int* __restrict__ ptr, const int* __restrict__ idx, const int* __restrict__ perm,
                   int n, int chunks, T* __restrict__ out) {
  const size_t D = (size_t)chunks * N;
  const int64_t total = (int64_t)n * chunks;
  for (int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; w < total; w += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(w / chunks);
    const size_t off = (size_t)(w - (int64_t)j * chunks) * N;
    float acc[N];
#pragma unroll
    for (int x = 0; x < N; ++x) acc[x] = 0.f;
    const int beg = ptr[j], end = ptr[j + 1];
    for (int s = beg; s < end; ++s) {
      const size_t t = (size_t)perm[idx ? idx[s] : s];
      float gg[N];
      load_chunk_keep<T, N>(g + t * D + off, gg);
#pragma unroll
      for (int x = 0; x < N; ++x) acc[x] += gg[x];
    }
    store_chunk<T, N>(out + (size_t)j * D + off, acc);
  }
}

// ---- LayerNorm + residual + segment sum -------------------------------------------------------------------
// one warp per dst row; lanes hold CPL chunks of the edge row in registers (D = 32*CPL*N elements max).
__device__ __forceinline__ float wsum(float x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

template <typename T, int N, int CPL>
__global__ void __launch_bounds__(128)
edge_ln_res_segsum_kernel(const T* __restrict__ y, const T* __restrict__ e, const T* __restrict__ gamma, const T* __restrict__ beta,
                          float eps, const int* __restrict__ rowptr, const int* __restrict__ perm, int Nd, int chunks,
                          T* __restrict__ edges_new, T* __restrict__ out, float* __restrict__ mean, float* __restrict__ rstd) {
  const int lane = threadIdx.x & 31;
  const size_t D = (size_t)chunks * N;
  const float invD = 1.f / (float)D;
  float gm[CPL][N], bt[CPL][N];
#pragma unroll
  for (int c = 0; c < CPL; ++c) {
    const int ch = lane + 32 * c;
    if (ch < chunks) {
      load_chunk_keep<T, N>(gamma + (size_t)ch * N, gm[c]);
      load_chunk_keep<T, N>(beta + (size_t)ch * N, bt[c]);
    }
  }
  const int nwarps = gridDim.x * (blockDim.x >> 5);
  for (int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < Nd; i += nwarps) {
    float acc[CPL][N];
#pragma unroll
    for (int c = 0; c < CPL; ++c)
#pragma unroll
      for (int x = 0; x < N; ++x) acc[c][x] = 0.f;
    const int beg = rowptr[i], end = rowptr[i + 1];
    for (int p = beg; p < end; ++p) {
      const size_t t = (size_t)perm[p];
      float yv[CPL][N], s = 0.f;
#pragma unroll
      for (int c = 0; c < CPL; ++c) {
        const int ch = lane + 32 * c;
        if (ch < chunks) {
          load_chunk<T, N>(y + t * D + (size_t)ch * N, yv[c]);
#pragma unroll
          for (int x = 0; x < N; ++x) s += yv[c][x];
        }
      }
      const float mu = wsum(s) * invD;
      float sq = 0.f;
#pragma unroll
      for (int c = 0; c < CPL; ++c) {
        const int ch = lane + 32 * c;
        if (ch < chunks) {
#pragma unroll
          for (int x = 0; x < N; ++x) {
            const float dlt = yv[c][x] - mu;
            sq = fmaf(dlt, dlt, sq);
          }
        }
      }
      const float rs = rsqrtf(wsum(sq) * invD + eps);
      if (lane == 0) {
        mean[t] = mu;
        rstd[t] = rs;
      }
#pragma unroll
      for (int c = 0; c < CPL; ++c) {
        const int ch = lane + 32 * c;
        if (ch < chunks) {
          float ev[N], o[N];
          load_chunk<T, N>(e + t * D + (size_t)ch * N, ev);
#pragma unroll
          for (int x = 0; x < N; ++x) {
            o[x] = fmaf((yv[c][x] - mu) * rs, gm[c][x], bt[c][x]) + ev[x];
            acc[c][x] += o[x];
          }
          store_chunk<T, N>(edges_new + t * D + (size_t)ch * N, o);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      const int ch = lane + 32 * c;
      if (ch < chunks) store_chunk<T, N>(out + (size_t)i * D + (size_t)ch * N, acc[c]);
    }
  }
}

// backward: one warp per edge row (grid-stride); per-CTA partial dgamma/dbeta -> partial[blockIdx][2][D]
template <typename T, int N, int CPL>
__global__ void __launch_bounds__(128)
edge_ln_res_segsum_bwd_kernel(const T* __restrict__ g_edges, const T* __restrict__ g_out, const T* __restrict__ y,
                              const T* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd,
                              const int64_t* __restrict__ ei, int64_t E, int chunks, T* __restrict__ dy, T* __restrict__ de,
                              float* __restrict__ partial) {
  extern __shared__ float sm[];  // [2][D] per-CTA dgamma / dbeta
  const int lane = threadIdx.x & 31;
  const size_t D = (size_t)chunks * N;
  const float invD = 1.f / (float)D;
  for (int x = threadIdx.x; x < 2 * (int)D; x += blockDim.x) sm[x] = 0.f;
  __syncthreads();
  float gm[CPL][N], dg[CPL][N], db[CPL][N];
#pragma unroll
  for (int c = 0; c < CPL; ++c) {
    const int ch = lane + 32 * c;
#pragma unroll
    for (int x = 0; x < N; ++x) dg[c][x] = db[c][x] = 0.f;
    if (ch < chunks) load_chunk_keep<T, N>(gamma + (size_t)ch * N, gm[c]);
  }
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t t = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5); t < E; t += nwarps) {
    const size_t i = (size_t)ei[E + t];
    const float mu = mean[t], rs = rstd[t];
    float gt[CPL][N], xh[CPL][N], s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      const int ch = lane + 32 * c;
      if (ch < chunks) {
        float go[N], yv[N];
        load_chunk_keep<T, N>(g_out + i * D + (size_t)ch * N, go);
        load_chunk<T, N>(y + (size_t)t * D + (size_t)ch * N, yv);
        if (g_edges) {
          float ge[N];
          load_chunk<T, N>(g_edges + (size_t)t * D + (size_t)ch * N, ge);
#pragma unroll
          for (int x = 0; x < N; ++x) go[x] += ge[x];
        }
#pragma unroll
        for (int x = 0; x < N; ++x) {
          gt[c][x] = go[x];
          xh[c][x] = (yv[x] - mu) * rs;
          const float dxh = go[x] * gm[c][x];
          s1 += dxh;
          s2 = fmaf(dxh, xh[c][x], s2);
          dg[c][x] = fmaf(go[x], xh[c][x], dg[c][x]);
          db[c][x] += go[x];
        }
        if (de) store_chunk<T, N>(de + (size_t)t * D + (size_t)ch * N, go);
      }
    }
    s1 = wsum(s1) * invD;
    s2 = wsum(s2) * invD;
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      const int ch = lane + 32 * c;
      if (ch < chunks) {
        float o[N];
#pragma unroll
        for (int x = 0; x < N; ++x) o[x] = rs * (gt[c][x] * gm[c][x] - s1 - xh[c][x] * s2);
        store_chunk<T, N>(dy + (size_t)t * D + (size_t)ch * N, o);
      }
    }
  }
  // warps of the CTA take turns adding into shared memory (fixed order -> deterministic)
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
    if ((int)(threadIdx.x >> 5) == w) {
#pragma unroll
      for (int c = 0; c < CPL; ++c) {
        const int ch = lane + 32 * c;
        if (ch < chunks) {
#pragma unroll
          for (int x = 0; x < N; ++x) {
            sm[(size_t)ch * N + x] += dg[c][x];
            sm[D + (size_t)ch * N + x] += db[c][x];
          }
        }
      }
    }
    __syncthreads();
  }
  for (int x = threadIdx.x; x < 2 * (int)D; x += blockDim.x) partial[(size_t)blockIdx.x * 2 * D + x] = sm[x];
}

__global__ void ln_partial_reduce_kernel(const float* __restrict__ partial, int nparts, int D, float* __restrict__ dgamma,
                                         float* __restrict__ dbeta) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= 2 * D) return;
  float s = 0.f;
  for (int b = 0; b < nparts; ++b) s += partial[(size_t)b * 2 * D + x];
  if (x < D)
    dgamma[x] = s;
  else
    dbeta[x - D] = s;
}

static int stream_grid(int64_t work_items, int threads) {
  const int64_t blocks = (work_items + threads - 1) / threads;
  return (int)std::max<int64_t>(1, std::min<int64_t>(blocks, (int64_t)num_sms() * 16));
}

}  // namespace ab2

using namespace ab2;

#define AB2_ACT_SWITCH(act, CALL) \
  switch (act) {                  \
    case 0: CALL(0); break;       \
    case 1: CALL(1); break;       \
    case 2: CALL(2); break;       \
    default: CALL(3); break;      \
  }

extern "C" int ab2_edge_gather_add_act(const void* pi, const void* pj, const void* pe, const int64_t* edge_index, int64_t E,
                                       int64_t Ns, int64_t Nd, int D, int dtype, int act, void* h0, void* pre, void* stream) {
  (void)Ns;
  (void)Nd;
  if (dtype != AB2_F32 && dtype != AB2_BF16) return fail(AB2_ERR_INVALID, "edge_gather_add_act: bad dtype");
  if (act < 0 || act > 3 || D <= 0 || E < 0) return fail(AB2_ERR_INVALID, "edge_gather_add_act: bad act/D/E");
  if (E == 0) return AB2_OK;
  if (!pi || !pj || !pe || !edge_index || !h0) return fail(AB2_ERR_INVALID, "edge_gather_add_act: null pointer argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int elt = dtype == AB2_F32 ? 4 : 2;
  const bool vec = (D * elt) % 16 == 0;
#define LAUNCH(T, N, A)                                                                                                     \
  edge_gather_add_act_kernel<T, N, A><<<stream_grid(E*(int64_t)(D / N), 256), 256, 0, st>>>((const T*)pi, (const T*)pj, (const T*)pe, \
                                                                                          edge_index, E, D / N, (T*)h0, (T*)pre)
#define CALL(A)                                                   \
  if (dtype == AB2_F32) {                                         \
    if (vec) LAUNCH(float, 4, A); else LAUNCH(float, 1, A);       \
  } else {                                                        \
    if (vec) LAUNCH(__nv_bfloat16, 8, A); else LAUNCH(__nv_bfloat16, 1, A); \
  }
  AB2_ACT_SWITCH(act, CALL)
#undef CALL
#undef LAUNCH
  AB2_LAUNCH_OK("edge_gather_add_act_kernel");
  return AB2_OK;
}

extern "C" int ab2_edge_gather_add_act_bwd(const void* g, const void* pre, const int32_t* rowptr, const int32_t* perm,
                                           const int32_t* colptr, const int32_t* cpos, int64_t E, int64_t Ns, int64_t Nd, int D,
                                           int dtype, int act, void* gpe, void* dpi, void* dpj, void* stream) {
  if (dtype != AB2_F32 && dtype != AB2_BF16) return fail(AB2_ERR_INVALID, "edge_gather_add_act_bwd: bad dtype");
  if (act < 0 || act > 3 || D <= 0 || E < 0) return fail(AB2_ERR_INVALID, "edge_gather_add_act_bwd: bad act/D/E");
  if (!rowptr || !gpe || (E > 0 && (!g || !pre || !perm))) return fail(AB2_ERR_INVALID, "edge_gather_add_act_bwd: null pointer argument");
  if (dpj && (!colptr || (E > 0 && !cpos))) return fail(AB2_ERR_INVALID, "edge_gather_add_act_bwd: dpj requested without the CSC view");
  cudaStream_t st = (cudaStream_t)stream;
  const int elt = dtype == AB2_F32 ? 4 : 2;
  const bool vec = (D * elt) % 16 == 0;
  if (Nd > 0) {
#define LAUNCH(T, N, A)                                                                                                   \
  edge_act_bwd_dst_kernel<T, N, A><<<stream_grid(Nd*(int64_t)(D / N), 256), 256, 0, st>>>((const T*)g, (const T*)pre, rowptr, perm, \
                                                                                        (int)Nd, D / N, (T*)gpe, (T*)dpi)
#define CALL(A)                                                   \
  if (dtype == AB2_F32) {                                         \
    if (vec) LAUNCH(float, 4, A); else LAUNCH(float, 1, A);       \
  } else {                                                        \
    if (vec) LAUNCH(__nv_bfloat16, 8, A); else LAUNCH(__nv_bfloat16, 1, A); \
  }
    AB2_ACT_SWITCH(act, CALL)
#undef CALL
#undef LAUNCH
    AB2_LAUNCH_OK("edge_act_bwd_dst_kernel");
  }
  if (dpj && Ns > 0) {
#define LAUNCH(T, N) \
  edge_act_bwd_src_kernel<T, N><<<stream_grid(Ns*(int64_t)(D / N), 256), 256, 0, st>>>((const T*)gpe, colptr, cpos, perm, (int)Ns, D / N, (T*)dpj)
    if (dtype == AB2_F32) {
      if (vec) LAUNCH(float, 4); else LAUNCH(float, 1);
    } else {
      if (vec) LAUNCH(__nv_bfloat16, 8); else LAUNCH(__nv_bfloat16, 1);
    }
#undef LAUNCH
    AB2_LAUNCH_OK("edge_act_bwd_src_kernel");
  }
  return AB2_OK;
}

extern "C" int ab2_edge_segment_sums(const void* g, const int32_t* rowptr, const int32_t* perm, const int32_t* colptr,
                                     const int32_t* cpos, int64_t E, int64_t Ns, int64_t Nd, int D, int dtype, void* dpi, void* dpj,
                                     void* stream) {
  if (dtype != AB2_F32 && dtype != AB2_BF16) return fail(AB2_ERR_INVALID, "edge_segment_sums: bad dtype");
  if (D <= 0 || E < 0) return fail(AB2_ERR_INVALID, "edge_segment_sums: bad D/E");
  if ((E > 0 && (!g || !perm)) || (dpi && !rowptr) || (dpj && (!colptr || (E > 0 && !cpos))))
    return fail(AB2_ERR_INVALID, "edge_segment_sums: null pointer argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int elt = dtype == AB2_F32 ? 4 : 2;
  const bool vec = (D * elt) % 16 == 0;
#define LAUNCH(T, N, PTR, IDX, CNT, OUT) \
  edge_segsum_kernel<T, N><<<stream_grid((CNT) * (int64_t)(D / N), 256), 256, 0, st>>>((const T*)g, PTR, IDX, perm, (int)(CNT), D / N, (T*)(OUT))
#define CALL(PTR, IDX, CNT, OUT)                                                          \
  if (dtype == AB2_F32) {                                                                 \
    if (vec) LAUNCH(float, 4, PTR, IDX, CNT, OUT); else LAUNCH(float, 1, PTR, IDX, CNT, OUT); \
  } else {                                                                                \
    if (vec) LAUNCH(__nv_bfloat16, 8, PTR, IDX, CNT, OUT); else LAUNCH(__nv_bfloat16, 1, PTR, IDX, CNT, OUT); \
  }
  if (dpi && Nd > 0) {
    CALL(rowptr, (const int*)nullptr, Nd, dpi)
    AB2_LAUNCH_OK("edge_segsum_kernel");
  }
  if (dpj && Ns > 0) {
    CALL(colptr, cpos, Ns, dpj)
    AB2_LAUNCH_OK("edge_segsum_kernel");
  }
#undef CALL
#undef LAUNCH
  return AB2_OK;
}

// chunks-per-lane dispatch: chunks <= 32*CPL
#define AB2_CPL_SWITCH(chunks, CALL)          \
  if ((chunks) <= 32) { CALL(1); }            \
  else if ((chunks) <= 64) { CALL(2); }       \
  else if ((chunks) <= 128) { CALL(4); }      \
  else { CALL(8); }

extern "C" int ab2_edge_ln_res_segsum(const void* y, const void* e, const void* gamma, const void* beta, float eps,
                                      const int32_t* rowptr, const int32_t* perm, int64_t E, int64_t Nd, int D, int dtype,
                                      void* edges_new, void* out, float* mean, float* rstd, void* stream) {
  if (dtype != AB2_F32 && dtype != AB2_BF16) return fail(AB2_ERR_INVALID, "edge_ln_res_segsum: bad dtype");
  if (D <= 0 || E < 0 || Nd < 0) return fail(AB2_ERR_INVALID, "edge_ln_res_segsum: bad D/E/Nd");
  if (Nd == 0) return AB2_OK;
  if (!rowptr || !out || !gamma || !beta || (E > 0 && (!y || !e || !perm || !edges_new || !mean || !rstd)))
    return fail(AB2_ERR_INVALID, "edge_ln_res_segsum: null pointer argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int elt = dtype == AB2_F32 ? 4 : 2;
  const bool vec = (D * elt) % 16 == 0;
  const int n = vec ? 16 / elt : 1;
  const int chunks = D / n;
  if (chunks > 256) return fail(AB2_ERR_UNSUPPORTED, "edge_ln_res_segsum: D=%d too wide (max %d)", D, 256 * n);
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((Nd + 3) / 4, (int64_t)num_sms() * 16));
#define LAUNCH(T, N, CPL)                                                                                                        \
  edge_ln_res_segsum_kernel<T, N, CPL><<<grid, 128, 0, st>>>((const T*)y, (const T*)e, (const T*)gamma, (const T*)beta, eps, rowptr, \
                                                             perm, (int)Nd, chunks, (T*)edges_new, (T*)out, mean, rstd)
#define CALL(CPL)                                                           \
  if (dtype == AB2_F32) {                                                   \
    if (vec) LAUNCH(float, 4, CPL); else LAUNCH(float, 1, CPL);             \
  } else {                                                                  \
    if (vec) LAUNCH(__nv_bfloat16, 8, CPL); else LAUNCH(__nv_bfloat16, 1, CPL); \
  }
  AB2_CPL_SWITCH(chunks, CALL)
#undef CALL
#undef LAUNCH
  AB2_LAUNCH_OK("edge_ln_res_segsum_kernel");
  return AB2_OK;
}

extern "C" int ab2_ln_bwd_parts(void) { return num_sms() * 4; }

extern "C" int ab2_edge_ln_res_segsum_bwd(const void* g_edges, const void* g_out, const void* y, const void* gamma,
                                          const float* mean, const float* rstd, const int64_t* edge_index, int64_t E, int64_t Nd,
                                          int D, int dtype, void* dy, void* de, float* partial, int nparts, float* dgamma,
                                          float* dbeta, void* stream) {
  (void)Nd;
  if (dtype != AB2_F32 && dtype != AB2_BF16) return fail(AB2_ERR_INVALID, "edge_ln_res_segsum_bwd: bad dtype");
  if (D <= 0 || E < 0 || nparts <= 0) return fail(AB2_ERR_INVALID, "edge_ln_res_segsum_bwd: bad D/E/nparts");
  if (!dgamma || !dbeta || !partial || !gamma || (E > 0 && (!g_out || !y || !mean || !rstd || !edge_index || !dy)))
    return fail(AB2_ERR_INVALID, "edge_ln_res_segsum_bwd: null pointer argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int elt = dtype == AB2_F32 ? 4 : 2;
  const bool vec = (D * elt) % 16 == 0;
  const int n = vec ? 16 / elt : 1;
  const int chunks = D / n;
  if (chunks > 256) return fail(AB2_ERR_UNSUPPORTED, "edge_ln_res_segsum_bwd: D=%d too wide (max %d)", D, 256 * n);
  const size_t smem = (size_t)2 * D * sizeof(float);
  if (smem > 48 * 1024) return fail(AB2_ERR_UNSUPPORTED, "edge_ln_res_segsum_bwd: D=%d too wide for the shared-memory reduction", D);
#define LAUNCH(T, N, CPL)                                                                                                          \
  edge_ln_res_segsum_bwd_kernel<T, N, CPL><<<nparts, 128, smem, st>>>((const T*)g_edges, (const T*)g_out, (const T*)y, (const T*)gamma, \
                                                                      mean, rstd, edge_index, E, chunks, (T*)dy, (T*)de, partial)
#define CALL(CPL)                                                           \
  if (dtype == AB2_F32) {                                                   \
    if (vec) LAUNCH(float, 4, CPL); else LAUNCH(float, 1, CPL);             \
  } else {                                                                  \
    if (vec) LAUNCH(__nv_bfloat16, 8, CPL); else LAUNCH(__nv_bfloat16, 1, CPL); \
  }
  AB2_CPL_SWITCH(chunks, CALL)
#undef CALL
#undef LAUNCH
  AB2_LAUNCH_OK("edge_ln_res_segsum_bwd_kernel");
  ln_partial_reduce_kernel<<<(2 * D + 255) / 256, 256, 0, st>>>(partial, nparts, D, dgamma, dbeta);
  AB2_LAUNCH_OK("ln_partial_reduce_kernel");
  return AB2_OK;
}
