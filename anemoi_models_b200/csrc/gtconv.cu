// Fused GraphTransformerConv forward / backward over a dst-sorted CSR (sm_100a).
//
// Replaces reference layers/conv.py:98-142 (+ PyG propagate / utils.softmax / scatter-add), where every
// step materialises an [E,H,C] or [E,H] tensor in HBM.  Here a destination row is owned by a group of
// threads (16 bytes of the row per thread, LPH consecutive lanes per head); the per-edge logit is reduced
// with warp shuffles inside the head group, the segment softmax runs online (flash style, base-2) and the
// weighted sum of (v_j + e_t) accumulates in registers: logits never touch HBM.
//
// HBM-bound: per edge one 16-B-per-lane read of e[t], k[src], v[src]; per dst one read of q and one write of out.
// Backward = two passes, both deterministic (no atomics):
//   bwd_dst (per dst segment):  recompute s, a; dq (registers), de (streamed), and the per-(edge,head) pair
//                               (a, ds/sqrt(C)) to a small [E,H] float2 workspace;
//   bwd_src (per src segment, CSC order): dk_j = sum ds*q_i, dv_j = sum a*g_i  (q, g rows are L2-resident).
#include <cmath>
#include <cstdlib>

#include "gtconv_args.cuh"

namespace ab2 {

constexpr int kThreads = 128;  // CTA size of the vector kernels

// src rows live in two pieces when the graph is dst-row sharded: rows [0, nsplit) in the rank's own k / v shard, rows
// [nsplit, Ns) in the halo buffer received from the peers (k2 / v2).  Single GPU: nsplit = Ns and k2, v2 are never read.
// The halo pointers handed to the kernels are VIRTUAL bases (halo - nsplit*D, computed on the host), so a row address is
// one select of the base plus j*D.  SPLIT=false (one GPU) compiles the select away.
template <bool SPLIT, typename T>
__device__ __forceinline__ T* src_row(T* own, T* halo_virtual, int nsplit, size_t j, size_t D) {
  if constexpr (SPLIT) {
    return (j < (size_t)nsplit ? own : halo_virtual) + j * D;
  } else {
    return own + j * D;
  }
}

template <int LPH>
__device__ __forceinline__ unsigned group_mask() {
  if constexpr (LPH == 32) {
    return 0xffffffffu;
  } else {
    const unsigned lane = threadIdx.x & 31u;
    return ((1u << LPH) - 1u) << (lane & ~(unsigned)(LPH - 1));
  }
}

// ------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------
template <typename T, int LPH, bool SPLIT>
__global__ void __launch_bounds__(kThreads)
gtconv_fwd_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, const T* __restrict__ e,
                  const int* __restrict__ rowptr, const int* __restrict__ col, const int* __restrict__ perm, int Nd,
                  RowMap rm, int H, float qscale, T* __restrict__ out, float* __restrict__ lse2, const T* __restrict__ k2,
                  const T* __restrict__ v2, int nsplit) {
  constexpr int VEC = Vec<T>::N;
  const int lr = threadIdx.x / rm.tpd;
  const int d = blockIdx.x * rm.rpb + lr;
  if (lr >= rm.rpb || d >= Nd) return;
  const int chunk = blockIdx.y * rm.tpd + (threadIdx.x - lr * rm.tpd);  // 16-byte chunk of the row
  const size_t D = (size_t)rm.chunks * VEC;
  const size_t off = (size_t)chunk * VEC;
  const unsigned mask = group_mask<LPH>();

  float qf[VEC];
  unpack<T>(ldg16_keep(q + (size_t)d * D + off), qf);
#pragma unroll
  for (int i = 0; i < VEC; ++i) qf[i] *= qscale;  // log2(e)/sqrt(C): logits come out in base-2 units

  const int beg = rowptr[d], end = rowptr[d + 1];
  float m = -INFINITY, l = 0.f, acc[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[i] = 0.f;

  int jn[kU], tn[kU];  // indices of the next chunk of edges, fetched one chunk ahead of the row loads
#pragma unroll
  for (int u = 0; u < kU; ++u) {
    jn[u] = beg + u < end ? col[beg + u] : 0;
    tn[u] = beg + u < end ? perm[beg + u] : 0;
  }
  for (int p = beg; p < end; p += kU) {
    uint4 kr[kU], er[kU], vr[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      if (p + u < end) {
        const size_t j = (size_t)jn[u], t = (size_t)tn[u];
        kr[u] = ldg16_keep(src_row<SPLIT>(k, k2, nsplit, j, D) + off);
        er[u] = ldg16(e + t * D + off);
        vr[u] = ldg16_keep(src_row<SPLIT>(v, v2, nsplit, j, D) + off);
      } else {
        kr[u] = er[u] = vr[u] = make_uint4(0, 0, 0, 0);
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int pn = p + kU + u;
      jn[u] = pn < end ? col[pn] : 0;
      tn[u] = pn < end ? perm[pn] : 0;
    }
    float s[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      float kf[VEC], ef[VEC];
      unpack<T>(kr[u], kf);
      unpack<T>(er[u], ef);
      float part = 0.f;
#pragma unroll
      for (int i = 0; i < VEC; ++i) part = fmaf(qf[i], kf[i] + ef[i], part);
      s[u] = part;
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) s[u] = group_sum<LPH>(s[u], mask);
    float mn = m;
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      if (p + u >= end) s[u] = -INFINITY;
      mn = fmaxf(mn, s[u]);
    }
    const float corr = fast_exp2(m - mn);  // first chunk: 2^(-inf) = 0
    l *= corr;
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] *= corr;
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const float pw = fast_exp2(s[u] - mn);  // masked edges: 2^(-inf) = 0
      l += pw;
      float vf[VEC], ef[VEC];
      unpack<T>(vr[u], vf);
      unpack<T>(er[u], ef);
#pragma unroll
      for (int i = 0; i < VEC; ++i) acc[i] = fmaf(pw, vf[i] + ef[i], acc[i]);
    }
    m = mn;
  }
  const float inv = 1.f / (l + 1e-16f);  // PyG: exp(s-max) / (sum + 1e-16); an empty segment keeps its zero row
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[i] *= inv;
  stg16(out + (size_t)d * D + off, pack<T>(acc));
  if ((chunk & (LPH - 1)) == 0) lse2[(size_t)d * H + chunk / LPH] = end > beg ? m + log2f(l + 1e-16f) : 0.f;
}

// ------------------------------------------------------------------------------------------------------
// backward, dst pass
// ------------------------------------------------------------------------------------------------------
template <typename T, int LPH, bool SPLIT>
__global__ void __launch_bounds__(kThreads, 4)
gtconv_bwd_dst_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, const T* __restrict__ e,
                      const int* __restrict__ rowptr, const int* __restrict__ col, const int* __restrict__ perm,
                      const int* __restrict__ csr2csc, int Nd, RowMap rm, int H, float qscale, float scale,
                      const T* __restrict__ out, const float* __restrict__ lse2, const T* __restrict__ g, T* __restrict__ dq,
                      T* __restrict__ de, float2* __restrict__ ads, const T* __restrict__ k2, const T* __restrict__ v2,
                      int nsplit) {
  constexpr int VEC = Vec<T>::N;
  const int lr = threadIdx.x / rm.tpd;
  const int d = blockIdx.x * rm.rpb + lr;
  if (lr >= rm.rpb || d >= Nd) return;
  const int chunk = blockIdx.y * rm.tpd + (threadIdx.x - lr * rm.tpd);
  const size_t D = (size_t)rm.chunks * VEC;
  const size_t off = (size_t)chunk * VEC;
  const unsigned mask = group_mask<LPH>();
  const int h = chunk / LPH;
  const bool leader = (chunk & (LPH - 1)) == 0;

  float qf[VEC], gf[VEC], dqa[VEC];
  unpack<T>(ldg16_keep(q + (size_t)d * D + off), qf);
  unpack<T>(ldg16_keep(g + (size_t)d * D + off), gf);
  float Dl;
  {
    float of[VEC];
    unpack<T>(ldg16(out + (size_t)d * D + off), of);
    float part = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) part = fmaf(gf[i], of[i], part);
    Dl = group_sum<LPH>(part, mask);
  }
  const float L = lse2[(size_t)d * H + h];
#pragma unroll
  for (int i = 0; i < VEC; ++i) dqa[i] = 0.f;

  const int beg = rowptr[d], end = rowptr[d + 1];
  int jn[kU], tn[kU], cn[kU];
#pragma unroll
  for (int u = 0; u < kU; ++u) {
    jn[u] = beg + u < end ? col[beg + u] : 0;
    tn[u] = beg + u < end ? perm[beg + u] : 0;
    cn[u] = (ads && beg + u < end) ? csr2csc[beg + u] : 0;
  }
  for (int p = beg; p < end; p += kU) {
    uint4 kr[kU], er[kU], vr[kU];
    size_t ts[kU], cs[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      ts[u] = (size_t)tn[u];
      cs[u] = (size_t)cn[u];
      if (p + u < end) {
        const size_t j = (size_t)jn[u];
        kr[u] = ldg16_keep(src_row<SPLIT>(k, k2, nsplit, j, D) + off);
        er[u] = ldg16(e + ts[u] * D + off);
        vr[u] = ldg16_keep(src_row<SPLIT>(v, v2, nsplit, j, D) + off);
      } else {
        kr[u] = er[u] = vr[u] = make_uint4(0, 0, 0, 0);
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int pn = p + kU + u;
      jn[u] = pn < end ? col[pn] : 0;
      tn[u] = pn < end ? perm[pn] : 0;
      cn[u] = (ads && pn < end) ? csr2csc[pn] : 0;
    }
    float s[kU], gv[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      float kf[VEC], ef[VEC], vf[VEC];
      unpack<T>(kr[u], kf);
      unpack<T>(er[u], ef);
      unpack<T>(vr[u], vf);
      float ps = 0.f, pg = 0.f;
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        ps = fmaf(qf[i], kf[i] + ef[i], ps);
        pg = fmaf(gf[i], vf[i] + ef[i], pg);
      }
      s[u] = ps;
      gv[u] = pg;
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      s[u] = group_sum<LPH>(s[u], mask);
      gv[u] = group_sum<LPH>(gv[u], mask);
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      if (p + u < end) {
        const float a = fast_exp2(fmaf(s[u], qscale, -L));
        const float dss = a * (gv[u] - Dl) * scale;  // d(logit)/sqrt(C)
        float kf[VEC], ef[VEC];
        unpack<T>(kr[u], kf);
        unpack<T>(er[u], ef);
#pragma unroll
        for (int i = 0; i < VEC; ++i) dqa[i] = fmaf(dss, kf[i] + ef[i], dqa[i]);
        if (de) {
          float o[VEC];
#pragma unroll
          for (int i = 0; i < VEC; ++i) o[i] = fmaf(a, gf[i], dss * qf[i]);
          stg16(de + ts[u] * D + off, pack<T>(o));
        }
        if (ads && leader) ads[cs[u] * H + h] = make_float2(a, dss);  // stored at the edge's src-sorted position
      }
    }
  }
  if (dq) stg16(dq + (size_t)d * D + off, pack<T>(dqa));
}

// ------------------------------------------------------------------------------------------------------
// backward, src pass
// ------------------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------------------
// low in-degree forward variant (decoder: 3 edges per dst): a thread group owns kDstRows CONSECUTIVE dst rows,
// whose incoming edges are one contiguous CSR range.  It streams that range in chunks of U edges with a single
// softmax / gradient state that is flushed when the row changes.  With one row per group these graphs are
// latency-bound: a CTA lives for one dependent rowptr -> index -> row -> store chain and moves only a few KB.
// ------------------------------------------------------------------------------------------------------
constexpr int kDstRows = 4;

template <int R>
__device__ __forceinline__ int row_of(int t, const int (&c)[R + 1]) {
  int r = 0;
#pragma unroll
  for (int i = 1; i < R; ++i) r += (t >= c[i]) ? 1 : 0;
  return r;
}
template <int R>
__device__ __forceinline__ uint4 select_row(const uint4 (&x)[R], int r) {
  uint4 y = x[0];
#pragma unroll
  for (int i = 1; i < R; ++i)
    if (r == i) y = x[i];
  return y;
}
template <int R>
__device__ __forceinline__ float select_row(const float (&x)[R], int r) {
  float y = x[0];
#pragma unroll
  for (int i = 1; i < R; ++i)
    if (r == i) y = x[i];
  return y;
}

template <typename T, int LPH, bool SPLIT>
__global__ void __launch_bounds__(kThreads)
gtconv_fwd_rows_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, const T* __restrict__ e,
                       const int* __restrict__ rowptr, const int* __restrict__ col, const int* __restrict__ perm, int Nd,
                       RowMap rm, int H, float qscale, T* __restrict__ out, float* __restrict__ lse2,
                       const T* __restrict__ k2, const T* __restrict__ v2, int nsplit) {
  constexpr int VEC = Vec<T>::N;
  constexpr int R = kDstRows;
  const int lr = threadIdx.x / rm.tpd;
  const long long d0 = ((long long)blockIdx.x * rm.rpb + lr) * R;
  if (lr >= rm.rpb || d0 >= Nd) return;
  const int chunk = blockIdx.y * rm.tpd + (threadIdx.x - lr * rm.tpd);
  const size_t D = (size_t)rm.chunks * VEC;
  const size_t off = (size_t)chunk * VEC;
  const unsigned mask = group_mask<LPH>();
  const bool leader = (chunk & (LPH - 1)) == 0;
  const int h = chunk / LPH;
  const int nrows = (int)min((long long)R, (long long)Nd - d0);

  int c[R + 1];
  uint4 qraw[R];
#pragma unroll
  for (int i = 0; i <= R; ++i) c[i] = rowptr[min(d0 + i, (long long)Nd)];
#pragma unroll
  for (int i = 0; i < R; ++i) qraw[i] = ldg16_keep(q + (size_t)min(d0 + i, (long long)Nd - 1) * D + off);

  float m = -INFINITY, l = 0.f, acc[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
  int cur = 0;

  auto flush_to = [&](int r) {  // finish row `cur` (PyG: divide by sum + 1e-16), zero the edge-less rows up to r
    const float inv = 1.f / (l + 1e-16f);
    float o[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) o[i] = acc[i] * inv;
    stg16(out + (size_t)(d0 + cur) * D + off, pack<T>(o));
    if (leader) lse2[(size_t)(d0 + cur) * H + h] = l > 0.f ? m + log2f(l + 1e-16f) : 0.f;
    for (int z = cur + 1; z < r && z < nrows; ++z) {
      stg16(out + (size_t)(d0 + z) * D + off, make_uint4(0, 0, 0, 0));
      if (leader) lse2[(size_t)(d0 + z) * H + h] = 0.f;
    }
    m = -INFINITY;
    l = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
    cur = r;
  };

  const int beg = c[0], end = c[R];
  int jn[kU], tn[kU];
#pragma unroll
  for (int u = 0; u < kU; ++u) {
    jn[u] = beg + u < end ? col[beg + u] : 0;
    tn[u] = beg + u < end ? perm[beg + u] : 0;
  }
  for (int p = beg; p < end; p += kU) {
    uint4 kr[kU], er[kU], vr[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      if (p + u < end) {
        const size_t j = (size_t)jn[u], t = (size_t)tn[u];
        kr[u] = ldg16_keep(src_row<SPLIT>(k, k2, nsplit, j, D) + off);
        er[u] = ldg16(e + t * D + off);
        vr[u] = ldg16_keep(src_row<SPLIT>(v, v2, nsplit, j, D) + off);
      } else {
        kr[u] = er[u] = vr[u] = make_uint4(0, 0, 0, 0);
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int pn = p + kU + u;
      jn[u] = pn < end ? col[pn] : 0;
      tn[u] = pn < end ? perm[pn] : 0;
    }
    float s[kU];
    int rr[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      rr[u] = row_of<R>(p + u, c);
      float qf[VEC], kf[VEC], ef[VEC];
      unpack<T>(select_row<R>(qraw, rr[u]), qf);
      unpack<T>(kr[u], kf);
      unpack<T>(er[u], ef);
      float part = 0.f;
#pragma unroll
      for (int i = 0; i < VEC; ++i) part = fmaf(qf[i], kf[i] + ef[i], part);
      s[u] = part;
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) s[u] = group_sum<LPH>(s[u], mask) * qscale;
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      if (p + u < end) {
        if (rr[u] != cur) flush_to(rr[u]);
        const float mn = fmaxf(m, s[u]);
        const float corr = fast_exp2(m - mn), pw = fast_exp2(s[u] - mn);
        l = fmaf(l, corr, pw);
        float vf[VEC], ef[VEC];
        unpack<T>(vr[u], vf);
        unpack<T>(er[u], ef);
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] = fmaf(acc[i], corr, pw * (vf[i] + ef[i]));
        m = mn;
      }
    }
  }
  flush_to(nrows);
}

// A thread group owns kSrcRows CONSECUTIVE src rows: their outgoing edges are one contiguous range of the CSC order,
// so the group streams that range in chunks of kU edges (all loads of a chunk in flight together) with a single
// accumulator pair that is flushed whenever the row changes.  (One row per group left the kernel latency-bound:
// the mean out-degree of an encoder graph is 1.4, i.e. one dependent index->row->store chain per CTA.)
constexpr int kSrcRows = 8;

template <typename T, int LPH, bool SPLIT>
__global__ void __launch_bounds__(kThreads, 6)
gtconv_bwd_src_kernel(const T* __restrict__ q, const T* __restrict__ g, const int* __restrict__ colptr,
                      const int* __restrict__ crow, const float2* __restrict__ ads, int src_lo, int Ns, RowMap rm, int H,
                      T* __restrict__ dk, T* __restrict__ dv, T* __restrict__ dk2, T* __restrict__ dv2, int nsplit) {
  // covers src rows [src_lo, Ns)
  constexpr int VEC = Vec<T>::N;
  const int lr = threadIdx.x / rm.tpd;
  const long long j0 = (long long)src_lo + ((long long)blockIdx.x * rm.rpb + lr) * kSrcRows;
  if (lr >= rm.rpb || j0 >= Ns) return;
  const int chunk = blockIdx.y * rm.tpd + (threadIdx.x - lr * rm.tpd);
  const size_t D = (size_t)rm.chunks * VEC;
  const size_t off = (size_t)chunk * VEC;
  const int h = chunk / LPH;
  int c[kSrcRows + 1];  // CSC offsets of the group's rows (rows past Ns are empty)
#pragma unroll
  for (int i = 0; i <= kSrcRows; ++i) c[i] = colptr[min(j0 + i, (long long)Ns)];
  float ka[VEC], va[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) ka[i] = va[i] = 0.f;
  const uint4 zero4 = make_uint4(0, 0, 0, 0);
  int cur = 0;  // row (relative to j0) the accumulators belong to

  auto flush_to = [&](int r) {  // store row `cur`, zero-fill the edge-less rows between, move on to row r
    if (dk) stg16(src_row<SPLIT>(dk, dk2, nsplit, (size_t)(j0 + cur), D) + off, pack<T>(ka));
    if (dv) stg16(src_row<SPLIT>(dv, dv2, nsplit, (size_t)(j0 + cur), D) + off, pack<T>(va));
    for (int z = cur + 1; z < r; ++z) {
      if (j0 + z < Ns) {
        if (dk) stg16(src_row<SPLIT>(dk, dk2, nsplit, (size_t)(j0 + z), D) + off, zero4);
        if (dv) stg16(src_row<SPLIT>(dv, dv2, nsplit, (size_t)(j0 + z), D) + off, zero4);
      }
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) ka[i] = va[i] = 0.f;
    cur = r;
  };

  const int tend = c[kSrcRows];
  for (int tb = c[0]; tb < tend; tb += kU) {
    uint4 qr[kU], gr[kU];
    float2 w[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      if (tb + u < tend) {
        const size_t i = (size_t)crow[2 * (tb + u)];
        w[u] = __ldg(ads + (size_t)(tb + u) * H + h);
        qr[u] = ldg16_keep(q + i * D + off);
        gr[u] = ldg16_keep(g + i * D + off);
      } else {
        w[u] = make_float2(0.f, 0.f);
        qr[u] = gr[u] = zero4;
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int t = tb + u;
      if (t < tend) {
        int r = 0;
#pragma unroll
        for (int i = 1; i < kSrcRows; ++i) r += (t >= c[i]) ? 1 : 0;
        if (r != cur) flush_to(r);
        float qf[VEC], gf[VEC];
        unpack<T>(qr[u], qf);
        unpack<T>(gr[u], gf);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          ka[i] = fmaf(w[u].y, qf[i], ka[i]);
          va[i] = fmaf(w[u].x, gf[i], va[i]);
        }
      }
    }
  }
  // last row with edges, then the trailing edge-less rows of the group
  const int last = (int)min((long long)kSrcRows, (long long)Ns - j0);
  flush_to(last);
}

// The same pass with the bookkeeping done ONCE PER WARP instead of once per thread (row layouts where a warp lies inside one
// row group, i.e. tpd % 32 == 0 -- all the model's shapes).  ncu on the kernel above (run r01y, headline graph): 444 M warp
// instructions of which 48 M are the FMAs -- the per-thread row lookup (7 compares per edge slot), the 9 colptr loads and the
// flush loop made it ~65 % issue-bound at 3.3 TB/s of writes.  Here lane l holds the CSC offset of row l of the group and
// lane l of an edge batch holds (dst, src) of edge l; the per-edge state every thread needs -- the dst index, "is this the
// last edge of its src row" -- comes from one shuffle and one bit of a ballot mask.
constexpr int kSrcRowsW = 24;  // most rows per warp group (<= 31; AB2_SRC_ROWS overrides, for experiments)

template <typename T, int LPH, bool SPLIT>
__global__ void __launch_bounds__(kThreads, 6)
gtconv_bwd_src_warp_kernel(const T* __restrict__ q, const T* __restrict__ g, const int* __restrict__ colptr,
                           const int2* __restrict__ cedge, const float2* __restrict__ ads, int src_lo, int Ns, RowMap rm, int H,
                           T* __restrict__ dk, T* __restrict__ dv, T* __restrict__ dk2, T* __restrict__ dv2, int nsplit,
                           int rows_per_group) {
  constexpr int VEC = Vec<T>::N;
  constexpr unsigned kFull = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int lr = threadIdx.x / rm.tpd;
  const long long j0l = (long long)src_lo + ((long long)blockIdx.x * rm.rpb + lr) * rows_per_group;
  if (lr >= rm.rpb || j0l >= Ns) return;  // warp-uniform: tpd is a multiple of 32
  const int j0 = (int)j0l;
  const int nrows = min(rows_per_group, Ns - j0);
  const int chunk = blockIdx.y * rm.tpd + (threadIdx.x - lr * rm.tpd);
  const size_t D = (size_t)rm.chunks * VEC;
  const size_t off = (size_t)chunk * VEC;
  const T* qo = q + off;
  const T* go = g + off;
  const float2* wo = ads + chunk / LPH;
  const uint4 zero4 = make_uint4(0, 0, 0, 0);

  const int cp = colptr[j0 + min(lane, nrows)];  // lane l: CSC offset of row j0 + l
  const int c0 = __shfl_sync(kFull, cp, 0), cend = __shfl_sync(kFull, cp, nrows);
  int2 ed = c0 + lane < cend ? cedge[c0 + lane] : make_int2(0, -1);  // first edge batch: lane l = (dst, src) of edge c0 + l
  {  // edge-less rows get zeros
    const int cpn = __shfl_down_sync(kFull, cp, 1);
    unsigned empty = __ballot_sync(kFull, lane < nrows && cpn == cp);
    while (empty) {
      const int z = __ffs(empty) - 1;
      empty &= empty - 1;
      if (dk) stg16(src_row<SPLIT>(dk, dk2, nsplit, (size_t)(j0 + z), D) + off, zero4);
      if (dv) stg16(src_row<SPLIT>(dv, dv2, nsplit, (size_t)(j0 + z), D) + off, zero4);
    }
  }
  float ka[VEC], va[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) ka[i] = va[i] = 0.f;

  for (int tb = c0; tb < cend; tb += 32) {
    const int t = tb + lane;
    if (tb != c0) ed = t < cend ? cedge[t] : make_int2(0, -1);
    int nsrc = __shfl_down_sync(kFull, ed.y, 1);  // src row of the next edge (-1 past the group's last edge)
    if (lane == 31) nsrc = t + 1 < cend ? cedge[t + 1].y : -1;
    const unsigned lastm = __ballot_sync(kFull, t < cend && nsrc != ed.y);  // bit l: edge tb + l is the last one of its row
    const int n = min(32, cend - tb);
    for (int u0 = 0; u0 < n; u0 += kU) {
      uint4 qr[kU], gr[kU];
      float2 w[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const size_t i = (size_t)__shfl_sync(kFull, ed.x, (u0 + u) & 31);
        if (u0 + u < n) {
          w[u] = __ldg(wo + (size_t)(tb + u0 + u) * H);
          qr[u] = ldg16_keep(qo + i * D);
          gr[u] = ldg16_keep(go + i * D);
        } else {
          w[u] = make_float2(0.f, 0.f);
          qr[u] = gr[u] = zero4;
        }
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        float qf[VEC], gf[VEC];
        unpack<T>(qr[u], qf);
        unpack<T>(gr[u], gf);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          ka[i] = fmaf(w[u].y, qf[i], ka[i]);
          va[i] = fmaf(w[u].x, gf[i], va[i]);
        }
        if ((lastm >> ((u0 + u) & 31)) & 1u) {  // warp-uniform; only set for u0 + u < n
          const int row = __shfl_sync(kFull, ed.y, (u0 + u) & 31);
          if (dk) stg16(src_row<SPLIT>(dk, dk2, nsplit, (size_t)row, D) + off, pack<T>(ka));
          if (dv) stg16(src_row<SPLIT>(dv, dv2, nsplit, (size_t)row, D) + off, pack<T>(va));
#pragma unroll
          for (int i = 0; i < VEC; ++i) ka[i] = va[i] = 0.f;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// generic kernels: any C (one warp per (row, head), lanes stride over channels).  Used when C*sizeof(T) is not
// a power-of-two multiple of 16 bytes (e.g. C=5, C=24) -- small test shapes; not the tuned path.
// ------------------------------------------------------------------------------------------------------
constexpr int kGenR = 8;  // channels per lane: C <= 256

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

template <typename T, bool SPLIT>
__global__ void __launch_bounds__(kThreads)
gtconv_fwd_generic_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, const T* __restrict__ e,
                          const int* __restrict__ rowptr, const int* __restrict__ col, const int* __restrict__ perm, int Nd, int H,
                          int C, float qscale, T* __restrict__ out, float* __restrict__ lse2, const T* __restrict__ k2,
                          const T* __restrict__ v2, int nsplit) {
  const int lane = threadIdx.x & 31;
  const long long w = (long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  if (w >= (long long)Nd * H) return;
  const int d = (int)(w / H), h = (int)(w % H);
  const size_t D = (size_t)H * C, ho = (size_t)h * C;
  float qf[kGenR], acc[kGenR];
#pragma unroll
  for (int r = 0; r < kGenR; ++r) {
    const int c = lane + 32 * r;
    qf[r] = c < C ? to_f<T>(q[(size_t)d * D + ho + c]) * qscale : 0.f;
    acc[r] = 0.f;
  }
  float m = -INFINITY, l = 0.f;
  const int beg = rowptr[d], end = rowptr[d + 1];
  for (int p = beg; p < end; ++p) {
    const size_t j = (size_t)col[p], t = (size_t)perm[p];
    float vv[kGenR], part = 0.f;
#pragma unroll
    for (int r = 0; r < kGenR; ++r) {
      const int c = lane + 32 * r;
      float kk = 0.f;
      vv[r] = 0.f;
      if (c < C) {
        const float ef = to_f<T>(e[t * D + ho + c]);
        kk = to_f<T>(src_row<SPLIT>(k, k2, nsplit, j, D)[ho + c]) + ef;
        vv[r] = to_f<T>(src_row<SPLIT>(v, v2, nsplit, j, D)[ho + c]) + ef;
      }
      part = fmaf(qf[r], kk, part);
    }
    const float s = warp_sum(part);
    const float mn = fmaxf(m, s);
    const float corr = fast_exp2(m - mn), pw = fast_exp2(s - mn);
    l = fmaf(l, corr, pw);
#pragma unroll
    for (int r = 0; r < kGenR; ++r) acc[r] = fmaf(acc[r], corr, pw * vv[r]);
    m = mn;
  }
  const float inv = 1.f / (l + 1e-16f);
#pragma unroll
  for (int r = 0; r < kGenR; ++r) {
    const int c = lane + 32 * r;
    if (c < C) out[(size_t)d * D + ho + c] = from_f<T>(acc[r] * inv);
  }
  if (lane == 0) lse2[(size_t)d * H + h] = end > beg ? m + log2f(l + 1e-16f) : 0.f;
}

template <typename T, bool SPLIT>
__global__ void __launch_bounds__(kThreads)
gtconv_bwd_dst_generic_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, const T* __restrict__ e,
                              const int* __restrict__ rowptr, const int* __restrict__ col, const int* __restrict__ perm,
                              const int* __restrict__ csr2csc, int Nd, int H, int C, float qscale, float scale, const T* __restrict__ out, const float* __restrict__ lse2,
                              const T* __restrict__ g, T* __restrict__ dq, T* __restrict__ de, float2* __restrict__ ads,
                              const T* __restrict__ k2, const T* __restrict__ v2, int nsplit) {
  const int lane = threadIdx.x & 31;
  const long long w = (long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  if (w >= (long long)Nd * H) return;
  const int d = (int)(w / H), h = (int)(w % H);
  const size_t D = (size_t)H * C, ho = (size_t)h * C;
  float qf[kGenR], gf[kGenR], dqa[kGenR], part = 0.f;
#pragma unroll
  for (int r = 0; r < kGenR; ++r) {
    const int c = lane + 32 * r;
    qf[r] = gf[r] = dqa[r] = 0.f;
    if (c < C) {
      qf[r] = to_f<T>(q[(size_t)d * D + ho + c]);
      gf[r] = to_f<T>(g[(size_t)d * D + ho + c]);
      part = fmaf(gf[r], to_f<T>(out[(size_t)d * D + ho + c]), part);
    }
  }
  const float Dl = warp_sum(part);
  const float L = lse2[(size_t)d * H + h];
  const int beg = rowptr[d], end = rowptr[d + 1];
  for (int p = beg; p < end; ++p) {
    const size_t j = (size_t)col[p], t = (size_t)perm[p];
    float kk[kGenR], ps = 0.f, pg = 0.f;
#pragma unroll
    for (int r = 0; r < kGenR; ++r) {
      const int c = lane + 32 * r;
      kk[r] = 0.f;
      float vv = 0.f;
      if (c < C) {
        const float ef = to_f<T>(e[t * D + ho + c]);
        kk[r] = to_f<T>(src_row<SPLIT>(k, k2, nsplit, j, D)[ho + c]) + ef;
        vv = to_f<T>(src_row<SPLIT>(v, v2, nsplit, j, D)[ho + c]) + ef;
      }
      ps = fmaf(qf[r], kk[r], ps);
      pg = fmaf(gf[r], vv, pg);
    }
    const float s = warp_sum(ps), gv = warp_sum(pg);
    const float a = fast_exp2(fmaf(s, qscale, -L));
    const float dss = a * (gv - Dl) * scale;
#pragma unroll
    for (int r = 0; r < kGenR; ++r) {
      const int c = lane + 32 * r;
      dqa[r] = fmaf(dss, kk[r], dqa[r]);
      if (de && c < C) de[t * D + ho + c] = from_f<T>(fmaf(a, gf[r], dss * qf[r]));
    }
    if (ads && lane == 0) ads[(size_t)csr2csc[p] * H + h] = make_float2(a, dss);
  }
  if (dq) {
#pragma unroll
    for (int r = 0; r < kGenR; ++r) {
      const int c = lane + 32 * r;
      if (c < C) dq[(size_t)d * D + ho + c] = from_f<T>(dqa[r]);
    }
  }
}

template <typename T, bool SPLIT>
__global__ void __launch_bounds__(kThreads)
gtconv_bwd_src_generic_kernel(const T* __restrict__ q, const T* __restrict__ g, const int* __restrict__ colptr,
                              const int* __restrict__ crow, const float2* __restrict__ ads, int src_lo, int Ns, int H, int C,
                              T* __restrict__ dk, T* __restrict__ dv, T* __restrict__ dk2, T* __restrict__ dv2, int nsplit) {
  const int lane = threadIdx.x & 31;
  const long long w = (long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  if (w >= (long long)(Ns - src_lo) * H) return;
  const int j = src_lo + (int)(w / H), h = (int)(w % H);
  const size_t D = (size_t)H * C, ho = (size_t)h * C;
  float ka[kGenR], va[kGenR];
#pragma unroll
  for (int r = 0; r < kGenR; ++r) ka[r] = va[r] = 0.f;
  const int beg = colptr[j], end = colptr[j + 1];
  for (int t = beg; t < end; ++t) {
    const size_t i = (size_t)crow[2 * (size_t)t];
    const float2 w2 = ads[(size_t)t * H + h];
#pragma unroll
    for (int r = 0; r < kGenR; ++r) {
      const int c = lane + 32 * r;
      if (c < C) {
        ka[r] = fmaf(w2.y, to_f<T>(q[i * D + ho + c]), ka[r]);
        va[r] = fmaf(w2.x, to_f<T>(g[i * D + ho + c]), va[r]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < kGenR; ++r) {
    const int c = lane + 32 * r;
    if (c < C) {
      if (dk) src_row<SPLIT>(dk, dk2, nsplit, (size_t)j, D)[ho + c] = from_f<T>(ka[r]);
      if (dv) src_row<SPLIT>(dv, dv2, nsplit, (size_t)j, D)[ho + c] = from_f<T>(va[r]);
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// host dispatch
// ------------------------------------------------------------------------------------------------------
struct Plan {
  bool vector;  // 16-byte vector kernels applicable
  int lph;
  RowMap rm;
  int slices;  // head slices per row (grid.y)
};

static Plan make_plan(int H, int C, int elt) {
  Plan pl{};
  const int head_bytes = C * elt;
  pl.vector = false;
  if (head_bytes % 16 == 0) {
    const int lph = head_bytes / 16;
    if (lph <= 32 && (lph & (lph - 1)) == 0) {
      int hs = 1;  // heads per slice: largest divisor of H with hs*lph <= kThreads
      for (int c = 1; c <= H; ++c)
        if (H % c == 0 && c * lph <= kThreads) hs = c;
      pl.vector = true;
      pl.lph = lph;
      pl.rm.tpd = hs * lph;
      pl.rm.rpb = kThreads / pl.rm.tpd;
      pl.rm.chunks = H * lph;
      pl.slices = H / hs;
    }
  }
  return pl;
}

// a warp never straddles two row groups -> the warp-cooperative src pass applies (AB2_SRC_WARP=0 keeps the per-thread kernel)
static bool src_warp_layout(const Plan& pl) {
  static const bool off = [] {
    const char* s = getenv("AB2_SRC_WARP");
    return s && atoi(s) == 0;
  }();
  return !off && pl.rm.tpd % 32 == 0;
}

// virtual base of a halo buffer: row j >= n_own lives at base + j*D  (never dereferenced for j < n_own)
template <typename T>
static T* vbase(T* halo, const ConvArgs& a) {
  if (!halo) return nullptr;
  return reinterpret_cast<T*>(reinterpret_cast<uintptr_t>(halo) - (uintptr_t)a.n_own * (uintptr_t)a.H * (uintptr_t)a.C * sizeof(T));
}

template <typename T, int LPH, bool SPLIT>
static void launch_fwd(const Plan& pl, const ConvArgs& a) {
  if (a.low_degree) {
    const int groups = (a.Nd + kDstRows - 1) / kDstRows;
    dim3 grid((groups + pl.rm.rpb - 1) / pl.rm.rpb, pl.slices);
    gtconv_fwd_rows_kernel<T, LPH, SPLIT><<<grid, kThreads, 0, a.st>>>((const T*)a.q, (const T*)a.k, (const T*)a.v, (const T*)a.e,
                                                               a.rowptr, a.col, a.perm, a.Nd, pl.rm, a.H, a.qscale, (T*)a.out_w,
                                                               a.lse2_w, vbase((const T*)a.k_halo, a), vbase((const T*)a.v_halo, a), a.n_own);
  } else {
    dim3 grid((a.Nd + pl.rm.rpb - 1) / pl.rm.rpb, pl.slices);
    gtconv_fwd_kernel<T, LPH, SPLIT><<<grid, kThreads, 0, a.st>>>((const T*)a.q, (const T*)a.k, (const T*)a.v, (const T*)a.e, a.rowptr,
                                                          a.col, a.perm, a.Nd, pl.rm, a.H, a.qscale, (T*)a.out_w, a.lse2_w,
                                                          vbase((const T*)a.k_halo, a), vbase((const T*)a.v_halo, a), a.n_own);
  }
}
template <typename T, int LPH, bool SPLIT>
static void launch_bwd_dst(const Plan& pl, const ConvArgs& a) {
  // (a multi-row dst pass was measured on B200: no gain at in-degree 3, slower at in-degree 8 -- one row per group here)
  dim3 grid((a.Nd + pl.rm.rpb - 1) / pl.rm.rpb, pl.slices);
  gtconv_bwd_dst_kernel<T, LPH, SPLIT><<<grid, kThreads, 0, a.st>>>((const T*)a.q, (const T*)a.k, (const T*)a.v, (const T*)a.e, a.rowptr,
                                                            a.col, a.perm, a.csr2csc, a.Nd, pl.rm, a.H, a.qscale, a.scale,
                                                            (const T*)a.out, a.lse2_in, (const T*)a.g, (T*)a.dq, (T*)a.de, a.ads,
                                                            vbase((const T*)a.k_halo, a), vbase((const T*)a.v_halo, a), a.n_own);
}
template <typename T, int LPH, bool SPLIT>
static void launch_bwd_src(const Plan& pl, const ConvArgs& a) {
  if (src_warp_layout(pl)) {
    // rows per warp group: about one 32-edge batch, but at least ~3 waves of CTAs on a small graph.  Measured on B200 (run r01aa),
    // headline graph (out-degree 1.4): 8 rows 0.564 ms, 12: 0.534, 16: 0.526, 24: 0.517, 31: 0.521 (per-thread kernel: 0.658);
    // BASELINE config-1 graphs (10-40 k src rows, fp32): 16 rows per group were too few CTAs (0.063 vs 0.037 ms).
    static const int forced = [] {
      const char* s = getenv("AB2_SRC_ROWS");
      const int v = s ? atoi(s) : 0;
      return v >= 1 && v <= 31 ? v : 0;
    }();
    const int nsrc = a.src_hi - a.src_lo;
    int rpg = forced;
    if (!rpg) {
      const double deg = nsrc > 0 ? (double)a.E / (double)std::max(a.Ns, 1) : 1.0;
      const int by_edges = (int)std::floor(32.0 / std::max(deg, 1.0) + 0.5);
      const int by_waves = nsrc / std::max(1, 3 * 6 * num_sms() * pl.rm.rpb);
      rpg = std::max(2, std::min(kSrcRowsW, std::min(by_edges, by_waves)));
    }
    const int groups = (a.src_hi - a.src_lo + rpg - 1) / rpg;
    dim3 grid((groups + pl.rm.rpb - 1) / pl.rm.rpb, pl.slices);
    gtconv_bwd_src_warp_kernel<T, LPH, SPLIT><<<grid, kThreads, 0, a.st>>>(
        (const T*)a.q, (const T*)a.g, a.colptr, reinterpret_cast<const int2*>(a.crow), a.ads, a.src_lo, a.src_hi, pl.rm, a.H, (T*)a.dk,
        (T*)a.dv, vbase((T*)a.dk_halo, a), vbase((T*)a.dv_halo, a), a.n_own, rpg);
    return;
  }
  const int groups = (a.src_hi - a.src_lo + kSrcRows - 1) / kSrcRows;
  dim3 grid((groups + pl.rm.rpb - 1) / pl.rm.rpb, pl.slices);
  gtconv_bwd_src_kernel<T, LPH, SPLIT><<<grid, kThreads, 0, a.st>>>((const T*)a.q, (const T*)a.g, a.colptr, a.crow, a.ads, a.src_lo, a.src_hi, pl.rm,
                                                            a.H, (T*)a.dk, (T*)a.dv, vbase((T*)a.dk_halo, a), vbase((T*)a.dv_halo, a), a.n_own);
}
template <typename T, bool SPLIT>
static void launch_generic(int which, const ConvArgs& a) {
  const int wpb = kThreads / 32;
  if (which == 0) {
    const unsigned grid = (unsigned)(((long long)a.Nd * a.H + wpb - 1) / wpb);
    gtconv_fwd_generic_kernel<T, SPLIT><<<grid, kThreads, 0, a.st>>>((const T*)a.q, (const T*)a.k, (const T*)a.v, (const T*)a.e, a.rowptr,
                                                             a.col, a.perm, a.Nd, a.H, a.C, a.qscale, (T*)a.out_w, a.lse2_w,
                                                             vbase((const T*)a.k_halo, a), vbase((const T*)a.v_halo, a), a.n_own);
  } else if (which == 1) {
    const unsigned grid = (unsigned)(((long long)a.Nd * a.H + wpb - 1) / wpb);
    gtconv_bwd_dst_generic_kernel<T, SPLIT><<<grid, kThreads, 0, a.st>>>((const T*)a.q, (const T*)a.k, (const T*)a.v, (const T*)a.e,
                                                                 a.rowptr, a.col, a.perm, a.csr2csc, a.Nd, a.H, a.C, a.qscale,
                                                                 a.scale, (const T*)a.out, a.lse2_in, (const T*)a.g, (T*)a.dq,
                                                                 (T*)a.de, a.ads, vbase((const T*)a.k_halo, a), vbase((const T*)a.v_halo, a), a.n_own);
  } else {
    const unsigned grid = (unsigned)(((long long)(a.src_hi - a.src_lo) * a.H + wpb - 1) / wpb);
    gtconv_bwd_src_generic_kernel<T, SPLIT><<<grid, kThreads, 0, a.st>>>((const T*)a.q, (const T*)a.g, a.colptr, a.crow, a.ads, a.src_lo, a.src_hi, a.H,
                                                                 a.C, (T*)a.dk, (T*)a.dv, vbase((T*)a.dk_halo, a), vbase((T*)a.dv_halo, a), a.n_own);
  }
}

#define AB2_DISPATCH_LPH(T, lph, CALL)                                   \
  switch (lph) {                                                         \
    case 1: CALL(T, 1); break;                                           \
    case 2: CALL(T, 2); break;                                           \
    case 4: CALL(T, 4); break;                                           \
    case 8: CALL(T, 8); break;                                           \
    case 16: CALL(T, 16); break;                                         \
    default: CALL(T, 32); break;                                         \
  }

static bool fwd_prefers_tma(int dtype, int H, int C, int64_t E, int64_t Nd);
static bool src_prefers_tma(int dtype, int H, int C, int64_t E, int64_t Ns);

// which: 0 = forward, 1 = backward dst pass, 2 = backward src pass
static int run_conv(int which, int dtype, const ConvArgs& a, const char* name) {
  const Plan pl = make_plan(a.H, a.C, dtype == AB2_F32 ? 4 : 2);
  const bool split = a.n_own < a.Ns;
  if (pl.vector && which == 0 && fwd_prefers_tma(dtype, a.H, a.C, a.E, a.Nd) && try_launch_fwd_tma(dtype, pl.lph, a)) {
    AB2_LAUNCH_OK(name);
    return AB2_OK;
  }
  if (pl.vector && which == 1 && try_launch_bwd_dst_tma(dtype, pl.lph, a)) {
    AB2_LAUNCH_OK(name);
    return AB2_OK;
  }
  if (pl.vector && which == 2 && src_prefers_tma(dtype, a.H, a.C, a.E, a.Ns) && try_launch_bwd_src_tma(dtype, pl.lph, a)) {
    AB2_LAUNCH_OK(name);
    return AB2_OK;
  }
  if (pl.vector) {
#define CALL3(T, L, S)                                   \
  do {                                                   \
    if (which == 0) launch_fwd<T, L, S>(pl, a);          \
    else if (which == 1) launch_bwd_dst<T, L, S>(pl, a); \
    else launch_bwd_src<T, L, S>(pl, a);                 \
  } while (0)
#define CALL(T, L)                 \
  do {                             \
    if (split) CALL3(T, L, true);  \
    else CALL3(T, L, false);       \
  } while (0)
    if (dtype == AB2_F32) {
      AB2_DISPATCH_LPH(float, pl.lph, CALL)
    } else {
      AB2_DISPATCH_LPH(__nv_bfloat16, pl.lph, CALL)
    }
#undef CALL
#undef CALL3
  } else if (dtype == AB2_F32) {
    if (split) launch_generic<float, true>(which, a);
    else launch_generic<float, false>(which, a);
  } else {
    if (split) launch_generic<__nv_bfloat16, true>(which, a);
    else launch_generic<__nv_bfloat16, false>(which, a);
  }
  AB2_LAUNCH_OK(name);
  return AB2_OK;
}

// mean in-degree below which a forward thread group takes kDstRows consecutive dst rows instead of one
// (measured on B200: decoder graph, in-degree 3: 2.42 -> 1.85 ms; processor graph, in-degree 8: no gain).  AB2_ROW_BLOCKS=0/1 forces it.
static bool use_row_blocks(int64_t E, int64_t Nd) {
  static const int forced = [] {
    const char* s = getenv("AB2_ROW_BLOCKS");
    return s ? atoi(s) : -1;
  }();
  if (forced >= 0) return forced != 0;
  return Nd > 0 && E < 6 * Nd;
}

// Forward kernel choice for 2 KB rows, measured on B200 (profiles/r01: A/B runs r01i, r01x).  With row blocks dealt round-robin
// the bulk-copy pipeline wins at every in-degree: 18.6 (encoder) LDG 0.699 -> 0.662 ms, 8 (processor) 0.302 -> 0.208,
// 3 (decoder) row-block LDG 1.85 -> 1.33.  (Its first version, one contiguous row range per CTA, lost on the encoder --
// 0.728 ms -- because neighbouring dst rows were no longer in flight together and k / v rows were fetched from HBM twice.)
// The backward dst pass: 1.00 -> 0.94, 0.455 -> 0.319, 3.62 -> 2.08 ms.  AB2_TMA bit 0 / bit 1 switch them off for A/B runs.
static bool fwd_prefers_tma(int dtype, int H, int C, int64_t E, int64_t Nd) {
  (void)E;
  (void)Nd;
  return tma_applicable(0, dtype, H, C);
}

// Backward src pass for 2 KB rows, measured on B200 (A/B runs r01j, r01n, r01x): at a mean out-degree of 1.4 (encoder) the
// LDG kernel (8-row blocks) and the pipeline (round-robin blocks) are within 2 % (0.657 vs 0.643 ms) -- LDG kept; at
// out-degree 8 (processor) the pipeline with one contiguous row range per CTA wins, 0.190 -> 0.139 ms; at out-degree 40
// (decoder) both sit at the L2/HBM gather limit (1.04 ms).
static bool src_prefers_tma(int dtype, int H, int C, int64_t E, int64_t Ns) {
  static const bool always = getenv("AB2_SRC_TMA_ALWAYS") != nullptr;  // A/B and profiling runs
  return tma_applicable(2, dtype, H, C) && (always || E >= 4 * Ns);
}

static int check_common(const char* fn, int dtype, int64_t Ns, int64_t Nd, int64_t E, int H, int C) {
  if (dtype != AB2_F32 && dtype != AB2_BF16) return fail(AB2_ERR_INVALID, "%s: dtype must be AB2_F32 or AB2_BF16", fn);
  if (Ns < 0 || Nd < 0 || E < 0 || H <= 0 || C <= 0) return fail(AB2_ERR_INVALID, "%s: negative or zero dimension", fn);
  if (Ns >= INT32_MAX || Nd >= INT32_MAX || E >= INT32_MAX) return fail(AB2_ERR_UNSUPPORTED, "%s: sizes must be < 2^31", fn);
  if (C > 32 * kGenR && !make_plan(H, C, dtype == AB2_F32 ? 4 : 2).vector)
    return fail(AB2_ERR_UNSUPPORTED, "%s: head width C=%d not supported (needs C*sizeof power-of-two multiple of 16 B, or C <= %d)", fn, C, 32 * kGenR);
  return 0;
}

static ConvArgs base_args(const void* q, const void* k, const void* v, const void* k_halo, const void* v_halo, int64_t n_own,
                          const void* e, const int32_t* rowptr, const int32_t* col, const int32_t* perm, int64_t Ns, int64_t Nd,
                          int64_t E, int H, int C, void* stream) {
  ConvArgs a{};
  a.q = q; a.k = k; a.v = v; a.e = e;
  a.k_halo = k_halo; a.v_halo = v_halo;
  a.n_own = (int)((k_halo || v_halo) ? n_own : Ns);
  a.rowptr = rowptr; a.col = col; a.perm = perm;
  a.Ns = (int)Ns; a.Nd = (int)Nd; a.E = E; a.H = H; a.C = C;
  a.src_lo = 0; a.src_hi = (int)Ns;
  a.scale = 1.f / sqrtf((float)C);
  a.qscale = kLog2e * a.scale;
  a.low_degree = use_row_blocks(E, Nd);
  a.st = (cudaStream_t)stream;
  return a;
}

static int check_halo(const char* fn, const void* k_halo, const void* v_halo, int64_t n_own, int64_t Ns) {
  if ((k_halo == nullptr) != (v_halo == nullptr)) return fail(AB2_ERR_INVALID, "%s: k_halo and v_halo must be given together", fn);
  if (k_halo && (n_own < 0 || n_own > Ns)) return fail(AB2_ERR_INVALID, "%s: n_own must be in [0, Ns]", fn);
  return 0;
}

}  // namespace ab2

using namespace ab2;

// Name of the kernel a call with these shapes dispatches to (for benchmark / profile bookkeeping).
extern "C" const char* ab2_gtconv_variant(int which, int dtype, int64_t Ns, int64_t Nd, int64_t E, int H, int C) {
  static thread_local char buf[96];
  const Plan pl = make_plan(H, C, dtype == AB2_F32 ? 4 : 2);
  const char* t = dtype == AB2_F32 ? "float" : "__nv_bfloat16";
  const char* base;
  if (!pl.vector) {
    base = which == 0 ? "gtconv_fwd_generic_kernel" : which == 1 ? "gtconv_bwd_dst_generic_kernel" : "gtconv_bwd_src_generic_kernel";
    snprintf(buf, sizeof(buf), "%s<%s>", base, t);
    return buf;
  }
  const bool low = use_row_blocks(E, Nd);
  if (which == 0)
    base = fwd_prefers_tma(dtype, H, C, E, Nd) ? "gtconv_fwd_tma_kernel" : low ? "gtconv_fwd_rows_kernel" : "gtconv_fwd_kernel";
  else if (which == 1)
    base = tma_applicable(1, dtype, H, C) ? "gtconv_bwd_dst_tma_kernel" : "gtconv_bwd_dst_kernel";
  else
    base = src_prefers_tma(dtype, H, C, E, Ns) ? "gtconv_bwd_src_tma_kernel" : src_warp_layout(pl) ? "gtconv_bwd_src_warp_kernel" : "gtconv_bwd_src_kernel";
  snprintf(buf, sizeof(buf), "%s<%s, %d>", base, t, pl.lph);
  return buf;
}

extern "C" int ab2_gtconv_fwd_halo(const void* q, const void* k, const void* v, const void* k_halo, const void* v_halo,
                                   int64_t n_own, const void* e, int dtype, const int32_t* rowptr, const int32_t* col,
                                   const int32_t* perm, int64_t Ns, int64_t Nd, int64_t E, int H, int C, void* out, float* lse2,
                                   void* stream) {
  if (int rc = check_common("gtconv_fwd", dtype, Ns, Nd, E, H, C)) return rc;
  if (int rc = check_halo("gtconv_fwd", k_halo, v_halo, n_own, Ns)) return rc;
  if (Nd == 0) return AB2_OK;
  if (!q || !out || !lse2 || !rowptr || (E > 0 && (!e || !col || !perm)) || (E > 0 && !k_halo && (!k || !v)))
    return fail(AB2_ERR_INVALID, "gtconv_fwd: null pointer argument");
  ConvArgs a = base_args(q, k, v, k_halo, v_halo, n_own, e, rowptr, col, perm, Ns, Nd, E, H, C, stream);
  a.qscale = kLog2e / sqrtf((float)C);
  a.out_w = out;
  a.lse2_w = lse2;
  return run_conv(0, dtype, a, "gtconv_fwd");
}

extern "C" int ab2_gtconv_fwd(const void* q, const void* k, const void* v, const void* e, int dtype, const int32_t* rowptr,
                              const int32_t* col, const int32_t* perm, int64_t Ns, int64_t Nd, int64_t E, int H, int C, void* out,
                              float* lse2, void* stream) {
  return ab2_gtconv_fwd_halo(q, k, v, nullptr, nullptr, Ns, e, dtype, rowptr, col, perm, Ns, Nd, E, H, C, out, lse2, stream);
}

extern "C" size_t ab2_gtconv_bwd_workspace_bytes(int64_t E, int H) { return (size_t)(E > 0 ? E : 1) * (size_t)H * sizeof(float2); }

static int bwd_dst_impl(ConvArgs a, int dtype, const int32_t* csr2csc, const void* out, const float* lse2, const void* g, void* dq,
                        void* de, void* ads_ws, size_t ads_ws_bytes) {
  if (a.Nd == 0) return AB2_OK;
  if (!a.rowptr || !a.q || !out || !lse2 || !g || (a.E > 0 && (!a.e || !a.col || !a.perm)) || (a.E > 0 && !a.k_halo && (!a.k || !a.v)))
    return fail(AB2_ERR_INVALID, "gtconv_bwd_dst: null pointer argument");
  if (ads_ws && ads_ws_bytes < ab2_gtconv_bwd_workspace_bytes(a.E, a.H)) return fail(AB2_ERR_INVALID, "gtconv_bwd_dst: workspace too small");
  if (ads_ws && a.E > 0 && !csr2csc) return fail(AB2_ERR_INVALID, "gtconv_bwd_dst: ads_ws given without csr2csc");
  a.csr2csc = csr2csc;
  a.out = out; a.lse2_in = lse2; a.g = g;
  a.dq = dq; a.de = de;
  a.ads = (float2*)ads_ws;
  return run_conv(1, dtype, a, "gtconv_bwd_dst");
}

static int bwd_src_impl(ConvArgs a, int dtype, const int32_t* colptr, const int32_t* crow, const void* g, const void* ads_ws, void* dk,
                        void* dv, void* dk_halo, void* dv_halo) {
  if (a.Ns == 0 || a.src_hi <= a.src_lo || (!dk && !dv && !dk_halo && !dv_halo)) return AB2_OK;
  if (!colptr || !ads_ws || (a.E > 0 && (!a.q || !g || !crow))) return fail(AB2_ERR_INVALID, "gtconv_bwd_src: null pointer argument");
  if (a.n_own < a.src_hi && ((dk && !dk_halo) || (dv && !dv_halo)))
    return fail(AB2_ERR_INVALID, "gtconv_bwd_src: halo rows in range but dk_halo / dv_halo missing");
  a.colptr = colptr; a.crow = crow; a.g = g;
  a.ads = (float2*)ads_ws;
  a.dk = dk; a.dv = dv; a.dk_halo = dk_halo; a.dv_halo = dv_halo;
  return run_conv(2, dtype, a, "gtconv_bwd_src");
}

extern "C" int ab2_gtconv_bwd_dst(const void* q, const void* k, const void* v, const void* e, int dtype, const int32_t* rowptr,
                                  const int32_t* col, const int32_t* perm, const int32_t* csr2csc, int64_t Ns, int64_t Nd,
                                  int64_t E, int H, int C, const void* out, const float* lse2, const void* g, void* dq, void* de,
                                  void* ads_ws, size_t ads_ws_bytes, void* stream) {
  if (int rc = check_common("gtconv_bwd_dst", dtype, Ns, Nd, E, H, C)) return rc;
  return bwd_dst_impl(base_args(q, k, v, nullptr, nullptr, Ns, e, rowptr, col, perm, Ns, Nd, E, H, C, stream), dtype, csr2csc, out,
                      lse2, g, dq, de, ads_ws, ads_ws_bytes);
}

extern "C" int ab2_gtconv_bwd_src(const void* q, const void* g, int dtype, const int32_t* colptr, const int32_t* crow,
                                  int64_t Ns, int64_t Nd, int64_t E, int H, int C, const void* ads_ws, void* dk, void* dv,
                                  void* stream) {
  if (int rc = check_common("gtconv_bwd_src", dtype, Ns, Nd, E, H, C)) return rc;
  return bwd_src_impl(base_args(q, nullptr, nullptr, nullptr, nullptr, Ns, nullptr, nullptr, nullptr, nullptr, Ns, Nd, E, H, C, stream),
                      dtype, colptr, crow, g, ads_ws, dk, dv, nullptr, nullptr);
}

extern "C" int ab2_gtconv_bwd_src_range(const void* q, const void* g, int dtype, const int32_t* colptr, const int32_t* crow,
                                        int64_t Ns, int64_t Nd, int64_t E, int H, int C, const void* ads_ws, void* dk, void* dv,
                                        int64_t row_begin, int64_t row_end, void* stream) {
  if (int rc = check_common("gtconv_bwd_src_range", dtype, Ns, Nd, E, H, C)) return rc;
  if (row_begin < 0 || row_end > Ns || row_begin > row_end) return fail(AB2_ERR_INVALID, "gtconv_bwd_src_range: bad row range");
  ConvArgs a = base_args(q, nullptr, nullptr, nullptr, nullptr, Ns, nullptr, nullptr, nullptr, nullptr, Ns, Nd, E, H, C, stream);
  a.src_lo = (int)row_begin;
  a.src_hi = (int)row_end;
  return bwd_src_impl(a, dtype, colptr, crow, g, ads_ws, dk, dv, nullptr, nullptr);
}

extern "C" int ab2_gtconv_bwd_dst_halo(const void* q, const void* k, const void* v, const void* k_halo, const void* v_halo,
                                       int64_t n_own, const void* e, int dtype, const int32_t* rowptr, const int32_t* col,
                                       const int32_t* perm, const int32_t* csr2csc, int64_t Ns, int64_t Nd, int64_t E, int H, int C,
                                       const void* out, const float* lse2, const void* g, void* dq, void* de, void* ads_ws,
                                       size_t ads_ws_bytes, void* stream) {
  if (int rc = check_common("gtconv_bwd_dst", dtype, Ns, Nd, E, H, C)) return rc;
  if (int rc = check_halo("gtconv_bwd_dst", k_halo, v_halo, n_own, Ns)) return rc;
  return bwd_dst_impl(base_args(q, k, v, k_halo, v_halo, n_own, e, rowptr, col, perm, Ns, Nd, E, H, C, stream), dtype, csr2csc, out,
                      lse2, g, dq, de, ads_ws, ads_ws_bytes);
}

extern "C" int ab2_gtconv_bwd_src_range_halo(const void* q, const void* g, int dtype, const int32_t* colptr, const int32_t* crow,
                                             int64_t n_own, int64_t Ns, int64_t Nd, int64_t E, int H, int C, const void* ads_ws,
                                             void* dk, void* dv, void* dk_halo, void* dv_halo, int64_t row_begin,
                                             int64_t row_end, void* stream) {
  if (int rc = check_common("gtconv_bwd_src_range", dtype, Ns, Nd, E, H, C)) return rc;
  if (row_begin < 0 || row_end > Ns || row_begin > row_end || n_own < 0 || n_own > Ns)
    return fail(AB2_ERR_INVALID, "gtconv_bwd_src_range: bad row range / n_own");
  ConvArgs a = base_args(q, nullptr, nullptr, nullptr, nullptr, Ns, nullptr, nullptr, nullptr, nullptr, Ns, Nd, E, H, C, stream);
  a.n_own = (int)n_own;
  a.src_lo = (int)row_begin;
  a.src_hi = (int)row_end;
  return bwd_src_impl(a, dtype, colptr, crow, g, ads_ws, dk, dv, dk_halo, dv_halo);
}

extern "C" int ab2_gtconv_bwd_halo(const void* q, const void* k, const void* v, const void* k_halo, const void* v_halo,
                                   int64_t n_own, const void* e, int dtype, const int32_t* rowptr, const int32_t* col,
                                   const int32_t* perm, const int32_t* colptr, const int32_t* csr2csc, const int32_t* crow,
                                   int64_t Ns, int64_t Nd, int64_t E, int H, int C, const void* out, const float* lse2,
                                   const void* g, void* dq, void* dk, void* dv, void* dk_halo, void* dv_halo, void* de,
                                   void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = check_common("gtconv_bwd", dtype, Ns, Nd, E, H, C)) return rc;
  if (int rc = check_halo("gtconv_bwd", k_halo, v_halo, n_own, Ns)) return rc;
  const bool need_src = dk || dv || dk_halo || dv_halo;
  if (need_src && (!workspace || workspace_bytes < ab2_gtconv_bwd_workspace_bytes(E, H)))
    return fail(AB2_ERR_INVALID, "gtconv_bwd: workspace too small");
  if (need_src && (!colptr || (E > 0 && (!csr2csc || !crow)))) return fail(AB2_ERR_INVALID, "gtconv_bwd: dk/dv requested without the CSC view");
  const ConvArgs a = base_args(q, k, v, k_halo, v_halo, n_own, e, rowptr, col, perm, Ns, Nd, E, H, C, stream);
  if (int rc = bwd_dst_impl(a, dtype, csr2csc, out, lse2, g, dq, de, need_src ? workspace : nullptr, workspace_bytes)) return rc;
  if (!need_src) return AB2_OK;
  return bwd_src_impl(a, dtype, colptr, crow, g, workspace, dk, dv, dk_halo, dv_halo);
}

extern "C" int ab2_gtconv_bwd(const void* q, const void* k, const void* v, const void* e, int dtype, const int32_t* rowptr,
                              const int32_t* col, const int32_t* perm, const int32_t* colptr, const int32_t* csr2csc,
                              const int32_t* crow, int64_t Ns, int64_t Nd, int64_t E, int H, int C, const void* out,
                              const float* lse2, const void* g, void* dq, void* dk, void* dv, void* de, void* workspace,
                              size_t workspace_bytes, void* stream) {
  return ab2_gtconv_bwd_halo(q, k, v, nullptr, nullptr, Ns, e, dtype, rowptr, col, perm, colptr, csr2csc, crow, Ns, Nd, E, H, C, out,
                             lse2, g, dq, dk, dv, nullptr, nullptr, de, workspace, workspace_bytes, stream);
}
