// WORK IN PROGRESS -- NOT part of libanemoi_b200.so (not listed in _build.py SOURCES); compiles (`nvcc -c`), never run.
//
// Bulk-copy pipelined variants of the folded-lin_edge kernels (csrc/gtconv_fold.cu) for 2 KB rows: the same
// warp-specialised producer / consumer structure as gtconv_tma.cu (blocks of dst rows dealt round-robin, 2-stage ring, mbarrier
// byte counting), but a stage carries per edge k[src], v[src] (2 KB each) and the 64-byte raw feature row instead of the 2 KB
// e row, and the per-dst "head" copies add the qw (and gw) projection rows.  Round-2 plan: add to SOURCES, call
// ab2_wip_fold_*_tma from ab2_gtconv_fold_fwd / _bwd_dst when they return true, run tests/test_gpu_zz_fold_draft.py.
//
// Invariants to check first when this runs (a mismatch = a hang on the stage's mbarrier):
//   * full(s): one arrival (producer lane 0, with expect_tx) + exactly `tx` bytes of bulk copies:
//       tx = first * head_bytes + n * (2 * 2048 + 64)
//   * empty(s): four arrivals (lane 0 of every consumer warp) after the stage has been read into registers
//   * every bulk copy: size a multiple of 16 B, both addresses 16-B aligned (raw rows: 64 B; qw / gw rows: H*64 B; lse2: H*4 B, H % 4 == 0)
#include <cmath>
#include <cstdlib>

#include "../gtconv_args.cuh"

namespace ab2 {
namespace foldtma {

constexpr int kRowBytes = 2048;
constexpr int kConsumers = 128;
constexpr int kThreadsTotal = kConsumers + 32;
constexpr int kStages = 2;
constexpr int kEdp = 16;
constexpr int kRawBytes = kEdp * 4;

struct __align__(16) StageMeta {
  int row, n, first, last;
  int cs[kU];  // src-sorted positions of the chunk's edges (backward: where (a, ds) goes)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ uint4 lds16(const void* p) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(smem_u32(p)));
  return r;
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

// stage = NHEAD per-dst rows (2 KB slots) | kU k rows | kU v rows | kU raw rows (64 B)
template <int NHEAD>
struct Ring {
  static constexpr int kStageBytes = (NHEAD + 2 * kU) * kRowBytes + kU * kRawBytes;
  static constexpr size_t kBytes = (size_t)kStages * kStageBytes + kStages * sizeof(StageMeta) + 2 * kStages * sizeof(uint64_t);
  char* base;
  __device__ char* head(int s, int i) const { return base + (size_t)s * kStageBytes + (size_t)i * kRowBytes; }
  __device__ char* k(int s, int u) const { return head(s, NHEAD + u); }
  __device__ char* v(int s, int u) const { return head(s, NHEAD + kU + u); }
  __device__ float* raw(int s, int u) const { return reinterpret_cast<float*>(head(s, NHEAD + 2 * kU) + (size_t)u * kRawBytes); }
  __device__ StageMeta* meta(int s) const { return reinterpret_cast<StageMeta*>(base + (size_t)kStages * kStageBytes) + s; }
  __device__ uint64_t* full(int s) const {
    return reinterpret_cast<uint64_t*>(base + (size_t)kStages * kStageBytes + kStages * sizeof(StageMeta)) + s;
  }
  __device__ uint64_t* empty(int s) const { return full(s) + kStages; }
};

struct Args {
  const void *q, *k, *v, *out, *g;
  const float *raw, *qw, *gw, *lse2_in;
  const int *rowptr, *col, *perm, *csr2csc;
  int Nd, H;
  float qscale, scale;
  void *out_w, *dq;
  float *lse2_w, *R, *S;
  float2* ads;
};

// producer: as gtconv_tma.cu producer_loop (round-robin blocks of RB <= 31 dst rows, next block's rowptr / index batches prefetched)
template <typename T, int NHEAD, bool BWD, typename HeadFn>
__device__ __forceinline__ void producer_loop(const Ring<NHEAD>& ring, const Args& a, int RB, uint32_t head_bytes, HeadFn head_copies) {
  const int lane = threadIdx.x & 31;
  const T* kb = (const T*)a.k;
  const T* vb = (const T*)a.v;
  constexpr size_t D = kRowBytes / sizeof(T);
  const int Nd = a.Nd;
  const int nblocks = (Nd + RB - 1) / RB;
  int s = 0;
  uint32_t phase = 0;
  auto load_batch = [&](int base, int pend, int& j, int& t, int& c) {
    const int p = base + lane;
    j = p < pend ? a.col[p] : 0;
    t = p < pend ? a.perm[p] : 0;
    c = (BWD && p < pend) ? a.csr2csc[p] : 0;
  };
  int b = blockIdx.x;
  if (b < nblocks) {
    int r0 = b * RB, r1 = min(r0 + RB, Nd);
    int ptr_cur = a.rowptr[min(r0 + lane, r1)];
    int pb = __shfl_sync(0xffffffffu, ptr_cur, 0), pend = __shfl_sync(0xffffffffu, ptr_cur, r1 - r0);
    int j0, t0, c0, j1, t1, c1;
    load_batch(pb, pend, j0, t0, c0);
    load_batch(pb + 32, pend, j1, t1, c1);
    while (true) {
      const int bn = b + (int)gridDim.x;
      const bool has_next = bn < nblocks;
      const int r0n = bn * RB, r1n = min(r0n + RB, Nd);
      int ptr_nxt = 0;
      if (has_next) ptr_nxt = a.rowptr[min(r0n + lane, r1n)];
      int pbn = 0, pendn = 0, jn0 = 0, tn0 = 0, cn0 = 0, jn1 = 0, tn1 = 0, cn1 = 0;
      bool next_loaded = false;
      int beg = __shfl_sync(0xffffffffu, ptr_cur, 0);
      for (int d = r0; d < r1; ++d) {
        const int end = __shfl_sync(0xffffffffu, ptr_cur, d + 1 - r0);
        int p = beg;
        do {
          const int n = min(kU, end - p);
          if (p >= pb + 32) {
            j0 = j1; t0 = t1; c0 = c1;
            pb += 32;
            load_batch(pb + 32, pend, j1, t1, c1);
          }
          mbar_wait(ring.empty(s), phase ^ 1u);
          const int pp = p + (lane < kU ? lane : 0) - pb;  // 0..63
          const int ja = __shfl_sync(0xffffffffu, j0, pp & 31), jb = __shfl_sync(0xffffffffu, j1, pp & 31);
          const int ta = __shfl_sync(0xffffffffu, t0, pp & 31), tb = __shfl_sync(0xffffffffu, t1, pp & 31);
          const int ca = __shfl_sync(0xffffffffu, c0, pp & 31), cb = __shfl_sync(0xffffffffu, c1, pp & 31);
          const int j = pp < 32 ? ja : jb, t = pp < 32 ? ta : tb, c = pp < 32 ? ca : cb;
          const bool first = p == beg, last = p + n >= end;
          StageMeta* m = ring.meta(s);
          if (lane == 0) {
            m->row = d;
            m->n = n;
            m->first = first;
            m->last = last;
          }
          if (lane < kU) m->cs[lane] = c;
          __syncwarp();
          if (lane == 0) {
            const uint32_t tx = (first ? head_bytes : 0u) + (uint32_t)n * (2u * kRowBytes + (uint32_t)kRawBytes);
            mbar_arrive_expect_tx(ring.full(s), tx);
          }
          __syncwarp();
          if (lane < n) {
            bulk_g2s(ring.k(s, lane), kb + (size_t)j * D, kRowBytes, ring.full(s));
            bulk_g2s(ring.v(s, lane), vb + (size_t)j * D, kRowBytes, ring.full(s));
            bulk_g2s(ring.raw(s, lane), a.raw + (size_t)t * kEdp, kRawBytes, ring.full(s));
          }
          if (first) head_copies(d, s, lane);
          p += n;
          if (++s == kStages) {
            s = 0;
            phase ^= 1u;
          }
        } while (p < end);
        beg = end;
        if (has_next && !next_loaded) {
          pbn = __shfl_sync(0xffffffffu, ptr_nxt, 0);
          pendn = __shfl_sync(0xffffffffu, ptr_nxt, r1n - r0n);
          load_batch(pbn, pendn, jn0, tn0, cn0);
          load_batch(pbn + 32, pendn, jn1, tn1, cn1);
          next_loaded = true;
        }
      }
      if (!has_next) break;
      b = bn; r0 = r0n; r1 = r1n;
      ptr_cur = ptr_nxt;
      pb = pbn; pend = pendn;
      j0 = jn0; t0 = tn0; c0 = cn0;
      j1 = jn1; t1 = tn1; c1 = cn1;
    }
  }
  mbar_wait(ring.empty(s), phase ^ 1u);
  if (lane == 0) {
    ring.meta(s)->row = -1;
    mbar_arrive_expect_tx(ring.full(s), 0);
  }
}

// ---- forward: head slots q, qw -------------------------------------------------------------------------------------------
constexpr int kCtasFwd = 4;

template <typename T, int LPH>
__global__ void __launch_bounds__(kThreadsTotal, kCtasFwd)
fold_fwd_tma_kernel(const __grid_constant__ Args a, int rows_per_block) {
  extern __shared__ __align__(128) char smem_raw[];
  constexpr int VEC = Vec<T>::N;
  constexpr int MPL = kEdp / LPH;
  constexpr size_t D = kRowBytes / sizeof(T);
  Ring<2> ring{smem_raw};
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(ring.full(s), 1);
      mbar_init(ring.empty(s), kConsumers / 32);
    }
    fence_barrier_init();
  }
  __syncthreads();
  const uint32_t qw_bytes = (uint32_t)a.H * kRawBytes;  // <= 2048: H <= 32

  if (threadIdx.x >= kConsumers) {
    const T* qb = (const T*)a.q;
    producer_loop<T, 2, false>(ring, a, rows_per_block, kRowBytes + qw_bytes, [&](int d, int s, int lane) {
      if (lane == kU) bulk_g2s(ring.head(s, 0), qb + (size_t)d * D, kRowBytes, ring.full(s));
      if (lane == kU + 1) bulk_g2s(ring.head(s, 1), a.qw + (size_t)d * a.H * kEdp, qw_bytes, ring.full(s));
    });
    return;
  }

  const int chunk = threadIdx.x;
  const size_t off = (size_t)chunk * 16;
  const unsigned mask = group_mask<LPH>();
  const int lane = threadIdx.x & 31;
  const int h = chunk / LPH, gl = chunk & (LPH - 1);
  const int moff = gl * MPL;
  T* out = (T*)a.out_w;
  float qf[VEC], acc[VEC], qwf[MPL], rr[MPL];
  float m = -INFINITY, l = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) qf[i] = acc[i] = 0.f;
#pragma unroll
  for (int i = 0; i < MPL; ++i) qwf[i] = rr[i] = 0.f;
  int s = 0;
  uint32_t phase = 0;
  while (true) {
    mbar_wait(ring.full(s), phase);
    const StageMeta* mt = ring.meta(s);
    const int row = mt->row, n = mt->n, first = mt->first, last = mt->last;
    if (row < 0) break;
    uint4 kr[kU], vr[kU], qr = make_uint4(0, 0, 0, 0);
    float rw[kU][MPL], qwn[MPL];
#pragma unroll
    for (int i = 0; i < MPL; ++i) qwn[i] = 0.f;
    if (first) {
      qr = lds16(ring.head(s, 0) + off);
      const float* qwp = reinterpret_cast<const float*>(ring.head(s, 1)) + h * kEdp + moff;
#pragma unroll
      for (int i = 0; i < MPL; ++i) qwn[i] = qwp[i];
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      if (u < n) {
        kr[u] = lds16(ring.k(s, u) + off);
        vr[u] = lds16(ring.v(s, u) + off);
        const float* rp = ring.raw(s, u) + moff;
#pragma unroll
        for (int i = 0; i < MPL; ++i) rw[u][i] = rp[i];
      } else {
        kr[u] = vr[u] = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int i = 0; i < MPL; ++i) rw[u][i] = 0.f;
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(ring.empty(s));
    if (++s == kStages) {
      s = 0;
      phase ^= 1u;
    }
    if (first) {
      unpack<T>(qr, qf);
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        qf[i] *= a.qscale;
        acc[i] = 0.f;
      }
#pragma unroll
      for (int i = 0; i < MPL; ++i) {
        qwf[i] = qwn[i] * a.qscale;
        rr[i] = 0.f;
      }
      m = -INFINITY;
      l = 0.f;
    }
    float sc[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      float kf[VEC];
      unpack<T>(kr[u], kf);
      float part = 0.f;
#pragma unroll
      for (int i = 0; i < VEC; ++i) part = fmaf(qf[i], kf[i], part);
#pragma unroll
      for (int i = 0; i < MPL; ++i) part = fmaf(qwf[i], rw[u][i], part);
      sc[u] = part;
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) sc[u] = group_sum<LPH>(sc[u], mask);
    if (n > 0) {
      float mn = m;
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        if (u >= n) sc[u] = -INFINITY;
        mn = fmaxf(mn, sc[u]);
      }
      const float corr = fast_exp2(m - mn);
      l *= corr;
#pragma unroll
      for (int i = 0; i < VEC; ++i) acc[i] *= corr;
#pragma unroll
      for (int i = 0; i < MPL; ++i) rr[i] *= corr;
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const float pw = fast_exp2(sc[u] - mn);
        l += pw;
        float vf[VEC];
        unpack<T>(vr[u], vf);
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] = fmaf(pw, vf[i], acc[i]);
#pragma unroll
        for (int i = 0; i < MPL; ++i) rr[i] = fmaf(pw, rw[u][i], rr[i]);
      }
      m = mn;
    }
    if (last) {
      const float inv = 1.f / (l + 1e-16f);
      float o[VEC];
#pragma unroll
      for (int i = 0; i < VEC; ++i) o[i] = acc[i] * inv;
      stg16(reinterpret_cast<char*>(out + (size_t)row * D) + off, pack<T>(o));
      float* Rp = a.R + ((size_t)row * a.H + h) * kEdp + moff;
#pragma unroll
      for (int i = 0; i < MPL; ++i) Rp[i] = rr[i] * inv;
      if (gl == 0) a.lse2_w[(size_t)row * a.H + h] = l > 0.f ? m + log2f(l + 1e-16f) : 0.f;
    }
  }
}

// ---- backward dst pass: head slots q, g, out, lse2, qw, gw ---------------------------------------------------------------
constexpr int kCtasBwd = 3;

template <typename T, int LPH>
__global__ void __launch_bounds__(kThreadsTotal, kCtasBwd)
fold_bwd_dst_tma_kernel(const __grid_constant__ Args a, int rows_per_block) {
  extern __shared__ __align__(128) char smem_raw[];
  constexpr int VEC = Vec<T>::N;
  constexpr int MPL = kEdp / LPH;
  constexpr size_t D = kRowBytes / sizeof(T);
  Ring<6> ring{smem_raw};
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(ring.full(s), 1);
      mbar_init(ring.empty(s), kConsumers / 32);
    }
    fence_barrier_init();
  }
  __syncthreads();
  const uint32_t lse_bytes = (uint32_t)a.H * 4u;  // H % 4 == 0 (checked on the host)
  const uint32_t w_bytes = (uint32_t)a.H * kRawBytes;

  if (threadIdx.x >= kConsumers) {
    const T* qb = (const T*)a.q;
    const T* gb = (const T*)a.g;
    const T* ob = (const T*)a.out;
    producer_loop<T, 6, true>(ring, a, rows_per_block, 3u * kRowBytes + lse_bytes + 2u * w_bytes, [&](int d, int s, int lane) {
      if (lane == kU) bulk_g2s(ring.head(s, 0), qb + (size_t)d * D, kRowBytes, ring.full(s));
      if (lane == kU + 1) bulk_g2s(ring.head(s, 1), gb + (size_t)d * D, kRowBytes, ring.full(s));
      if (lane == kU + 2) bulk_g2s(ring.head(s, 2), ob + (size_t)d * D, kRowBytes, ring.full(s));
      if (lane == kU + 3) bulk_g2s(ring.head(s, 3), a.lse2_in + (size_t)d * a.H, lse_bytes, ring.full(s));
      if (lane == kU + 4) bulk_g2s(ring.head(s, 4), a.qw + (size_t)d * a.H * kEdp, w_bytes, ring.full(s));
      if (lane == kU + 5) bulk_g2s(ring.head(s, 5), a.gw + (size_t)d * a.H * kEdp, w_bytes, ring.full(s));
    });
    return;
  }

  const int chunk = threadIdx.x;
  const size_t off = (size_t)chunk * 16;
  const unsigned mask = group_mask<LPH>();
  const int lane = threadIdx.x & 31;
  const int h = chunk / LPH, gl = chunk & (LPH - 1);
  const int moff = gl * MPL;
  const bool leader = gl == 0;
  T* dq = (T*)a.dq;
  float2* ads = a.ads;
  float qf[VEC], gf[VEC], dqa[VEC], qwf[MPL], gwf[MPL], ss[MPL];
  float Dl = 0.f, L = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) qf[i] = gf[i] = dqa[i] = 0.f;
#pragma unroll
  for (int i = 0; i < MPL; ++i) qwf[i] = gwf[i] = ss[i] = 0.f;
  int s = 0;
  uint32_t phase = 0;
  while (true) {
    mbar_wait(ring.full(s), phase);
    const StageMeta* mt = ring.meta(s);
    const int row = mt->row, n = mt->n, first = mt->first, last = mt->last;
    if (row < 0) break;
    size_t cs[kU];
    uint4 kr[kU], vr[kU];
    float rw[kU][MPL], qwn[MPL], gwn[MPL];
    uint4 qr = make_uint4(0, 0, 0, 0), gr = qr, orr = qr;
    float Lnew = 0.f;
#pragma unroll
    for (int i = 0; i < MPL; ++i) qwn[i] = gwn[i] = 0.f;
    if (first) {
      qr = lds16(ring.head(s, 0) + off);
      gr = lds16(ring.head(s, 1) + off);
      orr = lds16(ring.head(s, 2) + off);
      Lnew = reinterpret_cast<const float*>(ring.head(s, 3))[h];
      const float* qwp = reinterpret_cast<const float*>(ring.head(s, 4)) + h * kEdp + moff;
      const float* gwp = reinterpret_cast<const float*>(ring.head(s, 5)) + h * kEdp + moff;
#pragma unroll
      for (int i = 0; i < MPL; ++i) {
        qwn[i] = qwp[i];
        gwn[i] = gwp[i];
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      cs[u] = (size_t)mt->cs[u];
      if (u < n) {
        kr[u] = lds16(ring.k(s, u) + off);
        vr[u] = lds16(ring.v(s, u) + off);
        const float* rp = ring.raw(s, u) + moff;
#pragma unroll
        for (int i = 0; i < MPL; ++i) rw[u][i] = rp[i];
      } else {
        kr[u] = vr[u] = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int i = 0; i < MPL; ++i) rw[u][i] = 0.f;
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(ring.empty(s));
    if (++s == kStages) {
      s = 0;
      phase ^= 1u;
    }
    if (first) {
      float of[VEC];
      unpack<T>(qr, qf);
      unpack<T>(gr, gf);
      unpack<T>(orr, of);
      float part = 0.f;
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        part = fmaf(gf[i], of[i], part);
        dqa[i] = 0.f;
      }
#pragma unroll
      for (int i = 0; i < MPL; ++i) {
        qwf[i] = qwn[i];
        gwf[i] = gwn[i];
        ss[i] = 0.f;
      }
      Dl = group_sum<LPH>(part, mask);
      L = Lnew;
    }
    float sc[kU], gv[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      float kf[VEC], vf[VEC];
      unpack<T>(kr[u], kf);
      unpack<T>(vr[u], vf);
      float ps = 0.f, pg = 0.f;
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        ps = fmaf(qf[i], kf[i], ps);
        pg = fmaf(gf[i], vf[i], pg);
      }
#pragma unroll
      for (int i = 0; i < MPL; ++i) {
        ps = fmaf(qwf[i], rw[u][i], ps);
        pg = fmaf(gwf[i], rw[u][i], pg);
      }
      sc[u] = ps;
      gv[u] = pg;
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      sc[u] = group_sum<LPH>(sc[u], mask);
      gv[u] = group_sum<LPH>(gv[u], mask);
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      if (u < n) {
        const float aw = fast_exp2(fmaf(sc[u], a.qscale, -L));
        const float dss = aw * (gv[u] - Dl) * a.scale;
        float kf[VEC];
        unpack<T>(kr[u], kf);
#pragma unroll
        for (int i = 0; i < VEC; ++i) dqa[i] = fmaf(dss, kf[i], dqa[i]);
#pragma unroll
        for (int i = 0; i < MPL; ++i) ss[i] = fmaf(dss, rw[u][i], ss[i]);
        if (leader) ads[cs[u] * a.H + h] = make_float2(aw, dss);
      }
    }
    if (last) {
      if (dq) stg16(reinterpret_cast<char*>(dq + (size_t)row * D) + off, pack<T>(dqa));
      float* Sp = a.S + ((size_t)row * a.H + h) * kEdp + moff;
#pragma unroll
      for (int i = 0; i < MPL; ++i) Sp[i] = ss[i];
    }
  }
}

inline int rows_per_block(int64_t E, int64_t rows) {
  const double deg = rows > 0 ? (double)E / (double)rows : 1.0;
  if (deg >= 12.0) return 1;
  return (int)std::max(1.0, std::min(31.0, std::floor(32.0 / std::max(deg, 1.0) + 0.5)));
}

template <typename T, int LPH>
bool launch(int which, const Args& a, int64_t E, cudaStream_t st) {
  const int rb = rows_per_block(E, a.Nd);
  if (which == 0) {
    auto kern = fold_fwd_tma_kernel<T, LPH>;
    const size_t smem = Ring<2>::kBytes + 128;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return false;
    const int grid = std::max(1, std::min((a.Nd + rb - 1) / rb, num_sms() * kCtasFwd));
    kern<<<grid, kThreadsTotal, smem, st>>>(a, rb);
  } else {
    auto kern = fold_bwd_dst_tma_kernel<T, LPH>;
    const size_t smem = Ring<6>::kBytes + 128;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return false;
    const int grid = std::max(1, std::min((a.Nd + rb - 1) / rb, num_sms() * kCtasBwd));
    kern<<<grid, kThreadsTotal, smem, st>>>(a, rb);
  }
  return true;
}

template <typename T>
bool dispatch(int which, int lph, const Args& a, int64_t E, cudaStream_t st) {
  switch (lph) {
    case 4: return launch<T, 4>(which, a, E, st);
    case 8: return launch<T, 8>(which, a, E, st);
    case 16: return launch<T, 16>(which, a, E, st);
    default: return false;  // 2 KB rows with H <= 32, H % 4 == 0: lanes per head = 128 / H in {4, 8, 16, 32}; 32 has no raw column per lane
  }
}

}  // namespace foldtma
}  // namespace ab2

// Returns true when the pipelined kernel was launched (2 KB rows, 4 <= H <= 32, H % 4 == 0); false -> caller uses the LDG kernel.
extern "C" bool ab2_wip_fold_fwd_tma(const void* q, const void* k, const void* v, const float* raw, const float* qw, int dtype,
                                     const int32_t* rowptr, const int32_t* col, const int32_t* perm, int64_t Nd, int64_t E, int H, int C,
                                     void* out, float* lse2, float* R, void* stream) {
  using namespace ab2;
  const size_t elt = dtype == AB2_F32 ? 4 : 2;
  if ((size_t)H * C * elt != (size_t)foldtma::kRowBytes || H < 4 || H > 32 || H % 4 != 0 || Nd <= 0 || E <= 0) return false;
  foldtma::Args a{};
  a.q = q; a.k = k; a.v = v; a.raw = raw; a.qw = qw;
  a.rowptr = rowptr; a.col = col; a.perm = perm;
  a.Nd = (int)Nd; a.H = H;
  a.scale = 1.f / sqrtf((float)C);
  a.qscale = kLog2e * a.scale;
  a.out_w = out; a.lse2_w = lse2; a.R = R;
  const int lph = (int)((size_t)C * elt / 16);
  return dtype == AB2_F32 ? foldtma::dispatch<float>(0, lph, a, E, (cudaStream_t)stream)
                          : foldtma::dispatch<__nv_bfloat16>(0, lph, a, E, (cudaStream_t)stream);
}

extern "C" bool ab2_wip_fold_bwd_dst_tma(const void* q, const void* k, const void* v, const float* raw, const float* qw, const float* gw,
                                         int dtype, const int32_t* rowptr, const int32_t* col, const int32_t* perm, const int32_t* csr2csc,
                                         int64_t Nd, int64_t E, int H, int C, const void* out, const float* lse2, const void* g, void* dq,
                                         float* S, void* ads, void* stream) {
  using namespace ab2;
  const size_t elt = dtype == AB2_F32 ? 4 : 2;
  if ((size_t)H * C * elt != (size_t)foldtma::kRowBytes || H < 4 || H > 32 || H % 4 != 0 || Nd <= 0 || E <= 0) return false;
  foldtma::Args a{};
  a.q = q; a.k = k; a.v = v; a.raw = raw; a.qw = qw; a.gw = gw;
  a.rowptr = rowptr; a.col = col; a.perm = perm; a.csr2csc = csr2csc;
  a.Nd = (int)Nd; a.H = H;
  a.scale = 1.f / sqrtf((float)C);
  a.qscale = kLog2e * a.scale;
  a.out = out; a.lse2_in = lse2; a.g = g; a.dq = dq; a.S = S; a.ads = (float2*)ads;
  const int lph = (int)((size_t)C * elt / 16);
  return dtype == AB2_F32 ? foldtma::dispatch<float>(1, lph, a, E, (cudaStream_t)stream)
                          : foldtma::dispatch<__nv_bfloat16>(1, lph, a, E, (cudaStream_t)stream);
}
