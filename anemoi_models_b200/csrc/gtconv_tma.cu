// Bulk-copy (TMA engine) pipelined GraphTransformerConv kernels for 2 KB feature rows (D = 1024 bf16 / 512 fp32), sm_100a.
//
// Why: in the LDG kernels every in-flight 16 B costs four registers of the thread that will consume it, and a CTA walks a
// dependent rowptr -> index -> row -> store chain, so at low in-degree (decoder: 3 edges per dst) they are latency-bound.
// Here a CTA is warp-specialised:
//   * one PRODUCER warp streams the CSR edges of the CTA's blocks of dst rows (blocks are dealt round-robin, see producer_loop):
//     it reads the indices (coalesced, 32 edges per load, double-buffered) and issues one `cp.async.bulk` per gathered 2 KB
//     row (k[src], e[perm], v[src], plus q[dst] at the start of a dst row) into a shared-memory ring; completion is counted in
//     bytes on an mbarrier (`complete_tx`);
//   * four CONSUMER warps wait on the stage's mbarrier, read their 16 B of every row with LDS.128 and run the same
//     online-softmax / weighted-sum math as the LDG kernel (logits via lane-group shuffles, fp32 accumulation).
// In-flight data lives in shared memory (up to ~210 KB per SM) instead of registers, index latency is off the consumers'
// critical path, and one instruction moves a whole row.
#include <cmath>
#include <cstdlib>

#include "gtconv_args.cuh"

namespace ab2 {

constexpr int kRowBytes = 2048;
constexpr int kConsumers = 128;               // 4 warps, one 16-byte chunk of the row per thread
constexpr int kTmaThreads = kConsumers + 32;  // + the producer warp
constexpr int kStages = 2;                    // ring depth per CTA; 4 CTAs per SM interleave
constexpr int kCtasPerSm = 4;

struct __align__(16) StageMeta {
  int row;    // dst row of this chunk, -1 = no more work
  int n;      // edges in the chunk (0..kU)
  int first;  // first chunk of its row (q / g / out rows are in the stage)
  int last;   // last chunk of its row
  int t[kU];  // original edge ids (backward: where de goes)
  int cs[kU]; // src-sorted positions (backward: where the softmax weights go)
};

// ---- PTX wrappers ----------------------------------------------------------------------------------------------
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
// one 2 KB (or any multiple of 16 B) global -> shared copy by the TMA engine, completion counted on `bar`
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
__device__ __forceinline__ unsigned tma_group_mask() {
  if constexpr (LPH == 32) {
    return 0xffffffffu;
  } else {
    const unsigned lane = threadIdx.x & 31u;
    return ((1u << LPH) - 1u) << (lane & ~(unsigned)(LPH - 1));
  }
}

// ---- shared-memory ring ------------------------------------------------------------------------------------------
template <int NHEAD>  // rows staged at the start of a dst row: forward q (1); backward q, g, out (3)
struct Ring {
  static constexpr int kStageBytes = (NHEAD + 3 * kU) * kRowBytes;
  static constexpr size_t kBytes = (size_t)kStages * kStageBytes + kStages * sizeof(StageMeta) + 2 * kStages * sizeof(uint64_t);
  char* base;
  __device__ char* head(int s, int i) const { return base + (size_t)s * kStageBytes + (size_t)i * kRowBytes; }
  __device__ char* k(int s, int u) const { return head(s, NHEAD + u); }
  __device__ char* e(int s, int u) const { return head(s, NHEAD + kU + u); }
  __device__ char* v(int s, int u) const { return head(s, NHEAD + 2 * kU + u); }
  __device__ StageMeta* meta(int s) const { return reinterpret_cast<StageMeta*>(base + (size_t)kStages * kStageBytes) + s; }
  __device__ uint64_t* full(int s) const {
    return reinterpret_cast<uint64_t*>(base + (size_t)kStages * kStageBytes + kStages * sizeof(StageMeta)) + s;
  }
  __device__ uint64_t* empty(int s) const { return full(s) + kStages; }
};

// ---- producer: streams the CSR edges of blocks of RB consecutive dst rows; blocks are dealt round-robin to the CTAs ------
// Round-robin matters for L2: at any moment the rows in flight across the grid form one compact window of G*RB dst rows
// (as in the one-row-per-CTA LDG kernels), so the k / v rows their edges share are fetched from HBM once.  A first version
// gave every CTA one contiguous range of rows: ncu showed 12 % (forward) / 10 % (dst pass) extra DRAM reads, and 2.25 GB
// instead of 0.27 GB in the src pass, whose q / g rows stopped being L2-resident.
// The rowptr values and the first two 32-edge index batches of the CTA's NEXT block are requested while the current
// block streams, so a block switch costs no exposed latency.  HEAD(row, stage, lane) issues the per-row copies.
template <typename T, int NHEAD, bool BWD, typename HeadFn>
__device__ __forceinline__ void producer_loop(const Ring<NHEAD>& ring, const ConvArgs& a, int RB, uint32_t head_bytes,
                                              HeadFn head_copies) {
  const int lane = threadIdx.x & 31;
  const T* kb = (const T*)a.k;
  const T* vb = (const T*)a.v;
  const T* k2 = (const T*)a.k_halo;  // virtual bases (halo - n_own*D), see gtconv.cu
  const T* v2 = (const T*)a.v_halo;
  const T* eb = (const T*)a.e;
  constexpr size_t D = kRowBytes / sizeof(T);
  const int Nd = a.Nd;
  const int nblocks = (Nd + RB - 1) / RB;  // RB <= 31: the RB+1 rowptr values of a block live in the lanes of one load
  int s = 0;
  uint32_t phase = 0;
  auto load_batch = [&](int base, int pend, int& j, int& t, int& c) {
    const int p = base + lane;
    j = p < pend ? a.col[p] : 0;
    t = p < pend ? a.perm[p] : 0;
    c = (BWD && a.ads && p < pend) ? a.csr2csc[p] : 0;
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
          if (p >= pb + 32) {  // p advances by <= kU per chunk, so one shift keeps p inside batch 0 and p + n inside batch 1
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
          if (lane < kU) {
            m->t[lane] = t;
            m->cs[lane] = c;
          }
          __syncwarp();
          if (lane == 0) {
            const uint32_t tx = (first ? head_bytes : 0u) + (uint32_t)n * 3u * kRowBytes;
            mbar_arrive_expect_tx(ring.full(s), tx);  // release: the meta stores above are visible to whoever sees the phase flip
          }
          __syncwarp();
          if (lane < n) {
            const T* kp = ((size_t)j < (size_t)a.n_own ? kb : k2) + (size_t)j * D;
            const T* vp = ((size_t)j < (size_t)a.n_own ? vb : v2) + (size_t)j * D;
            bulk_g2s(ring.k(s, lane), kp, kRowBytes, ring.full(s));
            bulk_g2s(ring.e(s, lane), eb + (size_t)t * D, kRowBytes, ring.full(s));
            bulk_g2s(ring.v(s, lane), vp, kRowBytes, ring.full(s));
          }
          if (first) head_copies(d, s, lane);
          p += n;
          if (++s == kStages) {
            s = 0;
            phase ^= 1u;
          }
        } while (p < end);
        beg = end;
        if (has_next && !next_loaded) {  // the next block's rowptr has had a whole row's worth of stages to arrive
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
  // end marker
  mbar_wait(ring.empty(s), phase ^ 1u);
  if (lane == 0) {
    ring.meta(s)->row = -1;
    mbar_arrive_expect_tx(ring.full(s), 0);
  }
}

// rows per block, measured on the headline graph (in-degree 18.6, runs r01x / r01y): forward 0.631 ms at 1 row per block,
// 0.658 at 2, 0.673 at 3, 0.680 at 4, 0.692 at 8; backward dst pass 0.844 / 0.958 / 0.960 (the smaller the window of dst rows
// in flight, the better the L2 reuse of the k / v rows neighbouring dst rows share).  AB2_TMA_RB overrides, for experiments.
static int rows_per_block(int64_t E, int64_t rows) {
  static const int forced = [] {
    const char* s = getenv("AB2_TMA_RB");
    return s ? atoi(s) : 0;
  }();
  if (forced > 0) return std::min(forced, 31);
  const double deg = rows > 0 ? (double)E / (double)rows : 1.0;
  if (deg >= 12.0) return 1;
  return (int)std::max(1.0, std::min(31.0, std::floor(32.0 / std::max(deg, 1.0) + 0.5)));
}

// =====================================================================================================================
// forward
// =====================================================================================================================
template <typename T, int LPH>
__global__ void __launch_bounds__(kTmaThreads, kCtasPerSm)
gtconv_fwd_tma_kernel(const __grid_constant__ ConvArgs a, int rows_per_block) {
  extern __shared__ __align__(128) char smem_raw[];
  constexpr int VEC = Vec<T>::N;
  constexpr size_t D = kRowBytes / sizeof(T);
  Ring<1> ring{smem_raw};
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(ring.full(s), 1);
      mbar_init(ring.empty(s), kConsumers / 32);
    }
    fence_barrier_init();
  }
  __syncthreads();

  if (threadIdx.x >= kConsumers) {
    const T* qb = (const T*)a.q;
    producer_loop<T, 1, false>(ring, a, rows_per_block, kRowBytes, [&](int d, int s, int lane) {
      if (lane == kU) bulk_g2s(ring.head(s, 0), qb + (size_t)d * D, kRowBytes, ring.full(s));
    });
    return;
  }

  // ---- consumers
  const int chunk = threadIdx.x;
  const size_t off = (size_t)chunk * 16;
  const unsigned mask = tma_group_mask<LPH>();
  const int lane = threadIdx.x & 31;
  T* out = (T*)a.out_w;
  float qf[VEC], acc[VEC];
  float m = -INFINITY, l = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) qf[i] = acc[i] = 0.f;
  int s = 0;
  uint32_t phase = 0;
  while (true) {
    mbar_wait(ring.full(s), phase);
    const StageMeta* mt = ring.meta(s);
    const int row = mt->row, n = mt->n, first = mt->first, last = mt->last;
    if (row < 0) break;
    uint4 kr[kU], er[kU], vr[kU], qr = make_uint4(0, 0, 0, 0);
    if (first) qr = lds16(ring.head(s, 0) + off);
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      if (u < n) {
        kr[u] = lds16(ring.k(s, u) + off);
        er[u] = lds16(ring.e(s, u) + off);
        vr[u] = lds16(ring.v(s, u) + off);
      } else {
        kr[u] = er[u] = vr[u] = make_uint4(0, 0, 0, 0);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(ring.empty(s));  // the stage is in registers: hand the slot back to the producer
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
      m = -INFINITY;
      l = 0.f;
    }
    float sc[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      float kf[VEC], ef[VEC];
      unpack<T>(kr[u], kf);
      unpack<T>(er[u], ef);
      float part = 0.f;
#pragma unroll
      for (int i = 0; i < VEC; ++i) part = fmaf(qf[i], kf[i] + ef[i], part);
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
      for (int u = 0; u < kU; ++u) {
        const float pw = fast_exp2(sc[u] - mn);
        l += pw;
        float vf[VEC], ef[VEC];
        unpack<T>(vr[u], vf);
        unpack<T>(er[u], ef);
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] = fmaf(pw, vf[i] + ef[i], acc[i]);
      }
      m = mn;
    }
    if (last) {
      const float inv = 1.f / (l + 1e-16f);
      float o[VEC];
#pragma unroll
      for (int i = 0; i < VEC; ++i) o[i] = acc[i] * inv;
      stg16(reinterpret_cast<char*>(out + (size_t)row * D) + off, pack<T>(o));
      if ((chunk & (LPH - 1)) == 0) a.lse2_w[(size_t)row * a.H + chunk / LPH] = l > 0.f ? m + log2f(l + 1e-16f) : 0.f;
    }
  }
}

template <typename T, int LPH>
static bool launch_fwd_tma_t(const ConvArgs& a) {
  auto kern = gtconv_fwd_tma_kernel<T, LPH>;
  const size_t smem = Ring<1>::kBytes + 128;
  if (!AB2_ENSURE_DYN_SMEM(kern, smem)) return false;
  const int rb = rows_per_block(a.E, a.Nd);
  const int grid = std::max(1, std::min((a.Nd + rb - 1) / rb, num_sms() * kCtasPerSm));
  kern<<<grid, kTmaThreads, smem, a.st>>>(a, rb);
  return true;
}

// =====================================================================================================================
// backward, dst pass (same math as gtconv_bwd_dst_kernel in gtconv.cu)
// =====================================================================================================================
constexpr int kCtasPerSmBwd = 3;

template <typename T, int LPH>
__global__ void __launch_bounds__(kTmaThreads, kCtasPerSmBwd)
gtconv_bwd_dst_tma_kernel(const __grid_constant__ ConvArgs a, int rows_per_block) {
  extern __shared__ __align__(128) char smem_raw[];
  constexpr int VEC = Vec<T>::N;
  constexpr size_t D = kRowBytes / sizeof(T);
  Ring<4> ring{smem_raw};  // head slots: q, g, out, lse2 (H floats)
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(ring.full(s), 1);
      mbar_init(ring.empty(s), kConsumers / 32);
    }
    fence_barrier_init();
  }
  __syncthreads();
  const uint32_t lse_bytes = (uint32_t)a.H * 4u;  // a multiple of 16: H = 128 / LPH >= 4

  if (threadIdx.x >= kConsumers) {
    const T* qb = (const T*)a.q;
    const T* gb = (const T*)a.g;
    const T* ob = (const T*)a.out;
    producer_loop<T, 4, true>(ring, a, rows_per_block, 3u * kRowBytes + lse_bytes, [&](int d, int s, int lane) {
      if (lane == kU) bulk_g2s(ring.head(s, 0), qb + (size_t)d * D, kRowBytes, ring.full(s));
      if (lane == kU + 1) bulk_g2s(ring.head(s, 1), gb + (size_t)d * D, kRowBytes, ring.full(s));
      if (lane == kU + 2) bulk_g2s(ring.head(s, 2), ob + (size_t)d * D, kRowBytes, ring.full(s));
      if (lane == kU + 3) bulk_g2s(ring.head(s, 3), a.lse2_in + (size_t)d * a.H, lse_bytes, ring.full(s));
    });
    return;
  }

  const int chunk = threadIdx.x;
  const size_t off = (size_t)chunk * 16;
  const unsigned mask = tma_group_mask<LPH>();
  const int lane = threadIdx.x & 31;
  const int h = chunk / LPH;
  const bool leader = (chunk & (LPH - 1)) == 0;
  T* dq = (T*)a.dq;
  T* de = (T*)a.de;
  float2* ads = a.ads;
  float qf[VEC], gf[VEC], dqa[VEC];
  float Dl = 0.f, L = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) qf[i] = gf[i] = dqa[i] = 0.f;
  int s = 0;
  uint32_t phase = 0;
  while (true) {
    mbar_wait(ring.full(s), phase);
    const StageMeta* mt = ring.meta(s);
    const int row = mt->row, n = mt->n, first = mt->first, last = mt->last;
    if (row < 0) break;
    size_t ts[kU], cs[kU];
    uint4 kr[kU], er[kU], vr[kU];
    uint4 qr = make_uint4(0, 0, 0, 0), gr = qr, orr = qr;
    float Lnew = 0.f;
    if (first) {
      qr = lds16(ring.head(s, 0) + off);
      gr = lds16(ring.head(s, 1) + off);
      orr = lds16(ring.head(s, 2) + off);
      Lnew = reinterpret_cast<const float*>(ring.head(s, 3))[h];
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      ts[u] = (size_t)mt->t[u];
      cs[u] = (size_t)mt->cs[u];
      if (u < n) {
        kr[u] = lds16(ring.k(s, u) + off);
        er[u] = lds16(ring.e(s, u) + off);
        vr[u] = lds16(ring.v(s, u) + off);
      } else {
        kr[u] = er[u] = vr[u] = make_uint4(0, 0, 0, 0);
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
      Dl = group_sum<LPH>(part, mask);
      L = Lnew;
    }
    float sc[kU], gv[kU];
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
        float kf[VEC], ef[VEC];
        unpack<T>(kr[u], kf);
        unpack<T>(er[u], ef);
#pragma unroll
        for (int i = 0; i < VEC; ++i) dqa[i] = fmaf(dss, kf[i] + ef[i], dqa[i]);
        if (de) {
          float o[VEC];
#pragma unroll
          for (int i = 0; i < VEC; ++i) o[i] = fmaf(aw, gf[i], dss * qf[i]);
          stg16(reinterpret_cast<char*>(de + ts[u] * D) + off, pack<T>(o));
        }
        if (ads && leader) ads[cs[u] * a.H + h] = make_float2(aw, dss);
      }
    }
    if (last && dq) stg16(reinterpret_cast<char*>(dq + (size_t)row * D) + off, pack<T>(dqa));
  }
}

template <typename T, int LPH>
static bool launch_bwd_dst_tma_t(const ConvArgs& a) {
  auto kern = gtconv_bwd_dst_tma_kernel<T, LPH>;
  const size_t smem = Ring<4>::kBytes + 128;
  if (!AB2_ENSURE_DYN_SMEM(kern, smem)) return false;
  const int rb = rows_per_block(a.E, a.Nd);
  const int grid = std::max(1, std::min((a.Nd + rb - 1) / rb, num_sms() * kCtasPerSmBwd));
  kern<<<grid, kTmaThreads, smem, a.st>>>(a, rb);
  return true;
}

// =====================================================================================================================
// backward, src pass: dk_j = sum ds*q_i, dv_j = sum a*g_i over the outgoing edges of src row j (CSC order).
// Producer: per chunk of <= kU consecutive CSC edges of one src row, bulk-copies q[crow], g[crow] (2 KB each) and the
// chunk's slice of the (a, ds) workspace (contiguous in CSC order).  Consumers accumulate and write the row once.
// =====================================================================================================================
constexpr int kCtasPerSmSrc = 5;
constexpr int kAdsSlot = 4096;  // kU * H * 8 bytes with H <= 128

struct SrcRing {
  static constexpr int kStageBytes = 2 * kU * kRowBytes + kAdsSlot;
  static constexpr size_t kBytes = (size_t)kStages * kStageBytes + kStages * sizeof(StageMeta) + 2 * kStages * sizeof(uint64_t);
  char* base;
  __device__ char* q(int s, int u) const { return base + (size_t)s * kStageBytes + (size_t)u * kRowBytes; }
  __device__ char* g(int s, int u) const { return q(s, kU + u); }
  __device__ char* w(int s) const { return q(s, 2 * kU); }
  __device__ StageMeta* meta(int s) const { return reinterpret_cast<StageMeta*>(base + (size_t)kStages * kStageBytes) + s; }
  __device__ uint64_t* full(int s) const {
    return reinterpret_cast<uint64_t*>(base + (size_t)kStages * kStageBytes + kStages * sizeof(StageMeta)) + s;
  }
  __device__ uint64_t* empty(int s) const { return full(s) + kStages; }
};

template <typename T, int LPH>
__global__ void __launch_bounds__(kTmaThreads, kCtasPerSmSrc)
gtconv_bwd_src_tma_kernel(const __grid_constant__ ConvArgs a, int rows_per_cta, int interleave) {
  extern __shared__ __align__(128) char smem_raw[];
  constexpr int VEC = Vec<T>::N;
  constexpr size_t D = kRowBytes / sizeof(T);
  SrcRing ring{smem_raw};
  // One contiguous range of src rows per CTA: at the out-degrees this kernel is chosen for (>= 4) consecutive src rows gather
  // largely the same q / g rows, and that reuse is within the CTA's own stream of stages.  (Two other layouts were measured
  // and dropped -- profiles/r01/ab_r01x, r01y: stages that mix edges of several rows, 0.206 vs 0.139 ms at out-degree 8,
  // and blocks of rows dealt round-robin, 0.175 ms / at out-degree 40 1.41 vs 1.04 ms.)
  // interleave != 0 (high out-degree, see the launcher): CTA c takes the rows src_lo + c, + gridDim.x, + 2 gridDim.x, ...: the rows
  // in flight across the GPU are then ONE window of gridDim.x consecutive src rows, the q / g rows of a dst are wanted by its
  // (neighbouring) src rows at about the same time and are served by L2 instead of being fetched from DRAM once per edge.
  const int r0 = interleave ? (int)min((long long)a.src_lo + (long long)blockIdx.x, (long long)a.src_hi)
                            : (int)min((long long)a.src_lo + (long long)blockIdx.x * rows_per_cta, (long long)a.src_hi);
  const int r1 = interleave ? a.src_hi : (int)min((long long)r0 + rows_per_cta, (long long)a.src_hi);
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(ring.full(s), 1);
      mbar_init(ring.empty(s), kConsumers / 32);
    }
    fence_barrier_init();
  }
  __syncthreads();
  if (r0 >= r1) return;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x >= kConsumers) {
    // ---- producer warp: every stage carries <= kU consecutive CSC edges of ONE src row
    const T* qb = (const T*)a.q;
    const T* gb = (const T*)a.g;
    const int2* cedge = reinterpret_cast<const int2*>(a.crow);  // (dst, src) per src-sorted position
    const uint32_t edge_w = (uint32_t)a.H * 8u;                 // bytes of (a, ds) per edge
    int s = 0;
    uint32_t phase = 0;
    if (interleave) {
      const int stride = (int)gridDim.x;
      const int cnt = (r1 - r0 + stride - 1) / stride;  // rows of this CTA
      auto batch = [&](int b, int e) { return b + lane < e ? cedge[b + lane].x : 0; };
      // lane l keeps the CSC range of the CTA's row number (row_base + l); refreshed every 31 rows so that the NEXT row's range
      // is always at hand (its index batches are requested one row ahead)
      int row_base = 0, plo = 0, phi = 0;
      auto load_ptrs = [&](int ibase) {
        const long long jj = min((long long)r0 + (long long)(ibase + lane) * stride, (long long)a.src_hi - 1);
        plo = a.colptr[jj];
        phi = a.colptr[jj + 1];
      };
      load_ptrs(0);
      int beg = __shfl_sync(0xffffffffu, plo, 0), end = __shfl_sync(0xffffffffu, phi, 0);
      int i0 = batch(beg, end), i1 = batch(beg + 32, end);
      for (int it = 0; it < cnt; ++it) {
        const int j = r0 + it * stride;
        if (it + 1 - row_base >= 32) {
          row_base = it;
          load_ptrs(it);
        }
        beg = __shfl_sync(0xffffffffu, plo, it - row_base);
        end = __shfl_sync(0xffffffffu, phi, it - row_base);
        int n0 = 0, n1 = 0;
        if (it + 1 < cnt) {
          const int nb = __shfl_sync(0xffffffffu, plo, it + 1 - row_base), ne = __shfl_sync(0xffffffffu, phi, it + 1 - row_base);
          n0 = batch(nb, ne);
          n1 = batch(nb + 32, ne);
        }
        int pb = beg, p = beg;
        do {  // at least one stage per row, so that edge-less rows get their zeros written
          const int n = min(kU, end - p);
          if (p >= pb + 32) {
            i0 = i1;
            pb += 32;
            i1 = batch(pb + 32, end);
          }
          mbar_wait(ring.empty(s), phase ^ 1u);
          const int pp = p + (lane < kU ? lane : 0) - pb;
          const int ia = __shfl_sync(0xffffffffu, i0, pp & 31), ib = __shfl_sync(0xffffffffu, i1, pp & 31);
          const int i = pp < 32 ? ia : ib;
          if (lane == 0) {
            StageMeta* m = ring.meta(s);
            m->row = j;
            m->n = n;
            m->first = p == beg;
            m->last = p + n >= end;
            mbar_arrive_expect_tx(ring.full(s), (uint32_t)n * (2u * kRowBytes + edge_w));
          }
          __syncwarp();
          if (lane < n) {
            bulk_g2s(ring.q(s, lane), qb + (size_t)i * D, kRowBytes, ring.full(s));
            bulk_g2s(ring.g(s, lane), gb + (size_t)i * D, kRowBytes, ring.full(s));
          }
          if (lane == kU && n > 0) bulk_g2s(ring.w(s), a.ads + (size_t)p * a.H, (uint32_t)n * edge_w, ring.full(s));
          p += n;
          if (++s == kStages) {
            s = 0;
            phase ^= 1u;
          }
        } while (p < end);
        i0 = n0;
        i1 = n1;
      }
      mbar_wait(ring.empty(s), phase ^ 1u);
      if (lane == 0) {
        ring.meta(s)->row = -1;
        mbar_arrive_expect_tx(ring.full(s), 0);
      }
      return;
    }
    int pb = a.colptr[r0];
    const int pend = a.colptr[r1];
    auto load_batch = [&](int base) { return base + lane < pend ? cedge[base + lane].x : 0; };
    int i0 = load_batch(pb), i1 = load_batch(pb + 32);
    int ptr_base = r0 + 1;
    int next_ptr = a.colptr[min(ptr_base + lane, r1)];
    int beg = pb;
    for (int j = r0; j < r1; ++j) {
      if (j + 1 - ptr_base >= 32) {
        ptr_base = j + 1;
        next_ptr = a.colptr[min(ptr_base + lane, r1)];
      }
      const int end = __shfl_sync(0xffffffffu, next_ptr, j + 1 - ptr_base);
      int p = beg;
      do {  // at least one stage per row, so that edge-less rows get their zeros written
        const int n = min(kU, end - p);
        if (p >= pb + 32) {
          i0 = i1;
          pb += 32;
          i1 = load_batch(pb + 32);
        }
        mbar_wait(ring.empty(s), phase ^ 1u);
        const int pp = p + (lane < kU ? lane : 0) - pb;
        const int ia = __shfl_sync(0xffffffffu, i0, pp & 31), ib = __shfl_sync(0xffffffffu, i1, pp & 31);
        const int i = pp < 32 ? ia : ib;
        if (lane == 0) {
          StageMeta* m = ring.meta(s);
          m->row = j;
          m->n = n;
          m->first = p == beg;
          m->last = p + n >= end;
          mbar_arrive_expect_tx(ring.full(s), (uint32_t)n * (2u * kRowBytes + edge_w));
        }
        __syncwarp();
        if (lane < n) {
          bulk_g2s(ring.q(s, lane), qb + (size_t)i * D, kRowBytes, ring.full(s));
          bulk_g2s(ring.g(s, lane), gb + (size_t)i * D, kRowBytes, ring.full(s));
        }
        if (lane == kU && n > 0) bulk_g2s(ring.w(s), a.ads + (size_t)p * a.H, (uint32_t)n * edge_w, ring.full(s));
        p += n;
        if (++s == kStages) {
          s = 0;
          phase ^= 1u;
        }
      } while (p < end);
      beg = end;
    }
    mbar_wait(ring.empty(s), phase ^ 1u);
    if (lane == 0) {
      ring.meta(s)->row = -1;
      mbar_arrive_expect_tx(ring.full(s), 0);
    }
    return;
  }

  // ---- consumers: accumulate the stages of a row, write it once
  const int chunk = threadIdx.x;
  const size_t off = (size_t)chunk * 16;
  const int h = chunk / LPH;
  T* dk = (T*)a.dk;
  T* dv = (T*)a.dv;
  T* dk2 = (T*)a.dk_halo;  // virtual bases
  T* dv2 = (T*)a.dv_halo;
  float ka[VEC], va[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) ka[i] = va[i] = 0.f;
  int s = 0;
  uint32_t phase = 0;
  while (true) {
    mbar_wait(ring.full(s), phase);
    const StageMeta* mt = ring.meta(s);
    const int row = mt->row, n = mt->n, first = mt->first, last = mt->last;
    if (row < 0) break;
    uint4 qr[kU], gr[kU];
    float2 w[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      if (u < n) {
        qr[u] = lds16(ring.q(s, u) + off);
        gr[u] = lds16(ring.g(s, u) + off);
        w[u] = reinterpret_cast<const float2*>(ring.w(s))[u * a.H + h];
      } else {
        qr[u] = gr[u] = make_uint4(0, 0, 0, 0);
        w[u] = make_float2(0.f, 0.f);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(ring.empty(s));
    if (++s == kStages) {
      s = 0;
      phase ^= 1u;
    }
    if (first) {
#pragma unroll
      for (int i = 0; i < VEC; ++i) ka[i] = va[i] = 0.f;
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
    }
    if (last) {
      const bool own = row < a.n_own;
      if (dk) stg16(reinterpret_cast<char*>((own ? dk : dk2) + (size_t)row * D) + off, pack<T>(ka));
      if (dv) stg16(reinterpret_cast<char*>((own ? dv : dv2) + (size_t)row * D) + off, pack<T>(va));
    }
  }
}

template <typename T, int LPH>
static bool launch_bwd_src_tma_t(const ConvArgs& a) {
  auto kern = gtconv_bwd_src_tma_kernel<T, LPH>;
  const size_t smem = SrcRing::kBytes + 128;
  if (!AB2_ENSURE_DYN_SMEM(kern, smem)) return false;
  const int nrows = a.src_hi - a.src_lo;
  // CTAs = concurrent streams of src rows.  ncu (profiles/r02/ncu_conv_decoder_r02n.md): at out-degree 40 the q / g rows of a
  // dst are fetched from DRAM once per edge (L2 hit rate 6 %): between two consecutive src rows of one CTA the other 739 CTAs
  // move ~95 MB through L2.  AB2_SRC_CTAS_PER_SM trades parallelism for that reuse distance (experiments).
  // Measured (profiles/r02/SUMMARY.md): decoder (out-degree 40) 1.04 / 0.93 / 1.16 / 2.05 ms at 5 / 3 / 2 / 1 CTAs per SM,
  // processor (out-degree 8) 0.137 / 0.169 / 0.221 / 0.379 ms -> 3 per SM at a mean out-degree >= 20, else 5.
  static const int forced = [] {
    const char* s = getenv("AB2_SRC_CTAS_PER_SM");
    return s ? std::max(1, std::min(atoi(s), kCtasPerSmSrc)) : 0;
  }();
  // AB2_SRC_INTERLEAVE = 1 / 0 forces / forbids the interleaved row assignment (default: at a mean out-degree >= 8)
  static const int force_il = [] {
    const char* s = getenv("AB2_SRC_INTERLEAVE");
    return s ? (atoi(s) != 0 ? 1 : 0) : -1;
  }();
  const bool high_degree = (long long)a.E >= 20LL * std::max(a.Ns, 1);
  // interleaved rows from a mean out-degree of 8: decoder (40) 1.04 -> 0.60 ms, processor (8) 0.137 -> 0.131 ms (r02t)
  const bool il_degree = (long long)a.E >= 8LL * std::max(a.Ns, 1);
  const int interleave = force_il >= 0 ? force_il : (il_degree ? 1 : 0);
  const int per_sm = forced ? forced : ((high_degree && !interleave) ? 3 : kCtasPerSmSrc);
  const int ctas = std::max(1, std::min(nrows, num_sms() * per_sm));
  const int rows_per_cta = (nrows + ctas - 1) / ctas;
  const int grid = interleave ? ctas : (nrows + rows_per_cta - 1) / rows_per_cta;
  kern<<<grid, kTmaThreads, smem, a.st>>>(a, rows_per_cta, interleave);
  return true;
}

// AB2_TMA: bit 0 = forward (subject to the degree rule in gtconv.cu), bit 1 = backward dst pass, bit 2 = forward at every
// degree, bit 3 = backward src pass; default 11 (= 1|2|8); 0 sends 2 KB rows to the LDG kernels as well
static int tma_mask() {
  static const int v = [] {
    const char* s = getenv("AB2_TMA");
    return s ? atoi(s) : 11;
  }();
  return v;
}

// halo pointers become virtual bases here as well (row j >= n_own lives at base + j*D)
static ConvArgs with_virtual_halo(const ConvArgs& a, size_t elt) {
  ConvArgs b = a;
  const uintptr_t shift = (uintptr_t)a.n_own * (uintptr_t)a.H * (uintptr_t)a.C * elt;
  if (a.k_halo) b.k_halo = reinterpret_cast<const void*>(reinterpret_cast<uintptr_t>(a.k_halo) - shift);
  if (a.v_halo) b.v_halo = reinterpret_cast<const void*>(reinterpret_cast<uintptr_t>(a.v_halo) - shift);
  return b;
}

bool tma_applicable(int which, int dtype, int H, int C) {
  const size_t elt = dtype == AB2_F32 ? 4 : 2;
  return (tma_mask() & (which == 0 ? 1 : which == 1 ? 2 : 8)) && (size_t)H * C * elt == kRowBytes && H <= 128;
}

bool try_launch_fwd_tma(int dtype, int lph, const ConvArgs& a) {
  const size_t elt = dtype == AB2_F32 ? 4 : 2;
  if (!(tma_mask() & 1) || (size_t)a.H * a.C * elt != kRowBytes || a.Nd <= 0 || a.E <= 0) return false;
  const ConvArgs b = with_virtual_halo(a, elt);
  if (dtype == AB2_BF16) {
    switch (lph) {
      case 1: return launch_fwd_tma_t<__nv_bfloat16, 1>(b);
      case 2: return launch_fwd_tma_t<__nv_bfloat16, 2>(b);
      case 4: return launch_fwd_tma_t<__nv_bfloat16, 4>(b);
      case 8: return launch_fwd_tma_t<__nv_bfloat16, 8>(b);
      case 16: return launch_fwd_tma_t<__nv_bfloat16, 16>(b);
      default: return launch_fwd_tma_t<__nv_bfloat16, 32>(b);
    }
  }
  switch (lph) {
    case 1: return launch_fwd_tma_t<float, 1>(b);
    case 2: return launch_fwd_tma_t<float, 2>(b);
    case 4: return launch_fwd_tma_t<float, 4>(b);
    case 8: return launch_fwd_tma_t<float, 8>(b);
    case 16: return launch_fwd_tma_t<float, 16>(b);
    default: return launch_fwd_tma_t<float, 32>(b);
  }
}

bool try_launch_bwd_dst_tma(int dtype, int lph, const ConvArgs& a) {
  const size_t elt = dtype == AB2_F32 ? 4 : 2;
  if (!(tma_mask() & 2) || (size_t)a.H * a.C * elt != kRowBytes || a.Nd <= 0 || a.E <= 0) return false;
  const ConvArgs b = with_virtual_halo(a, elt);
  if (dtype == AB2_BF16) {
    switch (lph) {
      case 1: return launch_bwd_dst_tma_t<__nv_bfloat16, 1>(b);
      case 2: return launch_bwd_dst_tma_t<__nv_bfloat16, 2>(b);
      case 4: return launch_bwd_dst_tma_t<__nv_bfloat16, 4>(b);
      case 8: return launch_bwd_dst_tma_t<__nv_bfloat16, 8>(b);
      case 16: return launch_bwd_dst_tma_t<__nv_bfloat16, 16>(b);
      default: return launch_bwd_dst_tma_t<__nv_bfloat16, 32>(b);
    }
  }
  switch (lph) {
    case 1: return launch_bwd_dst_tma_t<float, 1>(b);
    case 2: return launch_bwd_dst_tma_t<float, 2>(b);
    case 4: return launch_bwd_dst_tma_t<float, 4>(b);
    case 8: return launch_bwd_dst_tma_t<float, 8>(b);
    case 16: return launch_bwd_dst_tma_t<float, 16>(b);
    default: return launch_bwd_dst_tma_t<float, 32>(b);
  }
}

}  // namespace ab2

namespace ab2 {
bool try_launch_bwd_src_tma(int dtype, int lph, const ConvArgs& a) {
  const size_t elt = dtype == AB2_F32 ? 4 : 2;
  if (!tma_applicable(2, dtype, a.H, a.C) || a.src_hi <= a.src_lo || a.E <= 0 || (!a.dk && !a.dv)) return false;
  ConvArgs b = a;
  const uintptr_t shift = (uintptr_t)a.n_own * (uintptr_t)a.H * (uintptr_t)a.C * elt;
  if (a.dk_halo) b.dk_halo = reinterpret_cast<void*>(reinterpret_cast<uintptr_t>(a.dk_halo) - shift);
  if (a.dv_halo) b.dv_halo = reinterpret_cast<void*>(reinterpret_cast<uintptr_t>(a.dv_halo) - shift);
  if (dtype == AB2_BF16) {
    switch (lph) {
      case 1: return launch_bwd_src_tma_t<__nv_bfloat16, 1>(b);
      case 2: return launch_bwd_src_tma_t<__nv_bfloat16, 2>(b);
      case 4: return launch_bwd_src_tma_t<__nv_bfloat16, 4>(b);
      case 8: return launch_bwd_src_tma_t<__nv_bfloat16, 8>(b);
      case 16: return launch_bwd_src_tma_t<__nv_bfloat16, 16>(b);
      default: return launch_bwd_src_tma_t<__nv_bfloat16, 32>(b);
    }
  }
  switch (lph) {
    case 1: return launch_bwd_src_tma_t<float, 1>(b);
    case 2: return launch_bwd_src_tma_t<float, 2>(b);
    case 4: return launch_bwd_src_tma_t<float, 4>(b);
    case 8: return launch_bwd_src_tma_t<float, 8>(b);
    case 16: return launch_bwd_src_tma_t<float, 16>(b);
    default: return launch_bwd_src_tma_t<float, 32>(b);
  }
}
}  // namespace ab2
