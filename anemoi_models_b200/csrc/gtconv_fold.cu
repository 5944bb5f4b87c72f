// GraphTransformerConv with the block's `lin_edge` folded in (sm_100a) -- ROUND-2 WORK IN PROGRESS, not on the default path.
//
// STATUS: first GPU run (profiles/r01/fold_draft_r01ah.log, the last seconds of round 1's GPU budget): parity against the product path
// conv(q, k, v, lin_edge(raw)) on four head layouts -- fp32 6e-7 .. 8e-7, bf16 5e-3 .. 6e-3 of max|ref|, outputs and all six
// gradients -- and, at the headline sizes (random graph), 6.56 ms per forward+backward against 13.93 ms for lin_edge + conv.
// The algebra is pinned on the CPU (oracle/gtconv.py::gt_conv_edge_folded_f64 against the reference op sequence,
// tests/test_oracle_golden.py) and the host glue against a torch emulation of these kernels (tests/test_host_logic.py).
// Not yet: the bulk-copy pipelined variants, halo rows, the block-level parity tests with the switch on -- so nothing calls
// these kernels unless AB2_EDGE_FOLD=1 is set (ops.gt_conv_folded); tests/test_gpu_zz_fold_draft.py runs them in a subprocess.
//
// Why: at the headline shape e = lin_edge(raw) is read twice and de written once -- 3*E*D*b = 4.6 GB of the step's 11.8 GB --
// and exists only because lin_edge (reference block.py:497, K = 11) is a separate GEMM.  e_t = W raw_t + b is LINEAR in the
// ed <= 15 raw features of the edge, so with W_h [C, ed] the rows of head h (bias folded in as a constant-1 raw column):
//     q_i . e_t       = (W_h^T q_i) . raw_t                 qw_i = W_h^T q_i : 16 floats per (dst, head), a tiny GEMM on the host side
//     sum_t a_t e_t   = W_h (sum_t a_t raw_t)               R_i  = sum_t a_t raw_t : 16 floats per (dst, head), accumulated here
//     g_i . e_t       = (W_h^T g_i) . raw_t                 gw_i
//     sum_t ds_t e_t  = W_h (sum_t ds_t raw_t)              S_i
// so per edge the kernels read k_j, v_j and 64 bytes of raw features; the [E, H, C] tensors e and de never exist.  The
// host side (ops.gt_conv_folded) does the [Nd]-sized einsums with W: qw, gw, out += R W^T, dq += S W^T, dW = g (x) R + q (x) S.
//
// Thread mapping = the LDG kernels of gtconv.cu (16 bytes of a row per thread, LPH lanes per head); lane l of a head group
// additionally owns raw columns [l*MPL, (l+1)*MPL), MPL = 16 / LPH, whose partial dot product joins the q.k partial BEFORE the
// lane-group reduction (no extra shuffles).
#include <cmath>

#include "gtconv_args.cuh"

namespace ab2 {
namespace {

constexpr int kFoldThreads = 128;
constexpr int kEdp = 16;  // padded raw feature count (ed + 1 bias column <= 16)

template <int LPH>
__device__ __forceinline__ unsigned fold_group_mask() {
  if constexpr (LPH == 32) {
    return 0xffffffffu;
  } else {
    const unsigned lane = threadIdx.x & 31u;
    return ((1u << LPH) - 1u) << (lane & ~(unsigned)(LPH - 1));
  }
}

template <int MPL>
__device__ __forceinline__ void load_cols(const float* p, float (&x)[MPL]) {
  if constexpr (MPL == 1) {
    x[0] = __ldg(p);
  } else if constexpr (MPL == 2) {
    const float2 t = __ldg(reinterpret_cast<const float2*>(p));
    x[0] = t.x;
    x[1] = t.y;
  } else {
#pragma unroll
    for (int i = 0; i < MPL; i += 4) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(p + i));
      x[i] = t.x;
      x[i + 1] = t.y;
      x[i + 2] = t.z;
      x[i + 3] = t.w;
    }
  }
}

// forward: out_part_i = sum_t a_t v_j (the W R_i term is added by the caller), lse2, R_i = sum_t a_t raw_t
template <typename T, int LPH>
__global__ void __launch_bounds__(kFoldThreads)
gtconv_fold_fwd_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, const float* __restrict__ raw,
                       const float* __restrict__ qw, const int* __restrict__ rowptr, const int* __restrict__ col,
                       const int* __restrict__ perm, int Nd, RowMap rm, int H, float qscale, T* __restrict__ out,
                       float* __restrict__ lse2, float* __restrict__ R) {
  constexpr int VEC = Vec<T>::N;
  constexpr int MPL = kEdp / LPH;
  const int lr = threadIdx.x / rm.tpd;
  const int d = blockIdx.x * rm.rpb + lr;
  if (lr >= rm.rpb || d >= Nd) return;
  const int chunk = blockIdx.y * rm.tpd + (threadIdx.x - lr * rm.tpd);
  const size_t D = (size_t)rm.chunks * VEC;
  const size_t off = (size_t)chunk * VEC;
  const unsigned mask = fold_group_mask<LPH>();
  const int h = chunk / LPH, gl = chunk & (LPH - 1);
  const int moff = gl * MPL;  // first raw column of this lane

  float qf[VEC], qwf[MPL];
  unpack<T>(ldg16_keep(q + (size_t)d * D + off), qf);
  load_cols<MPL>(qw + ((size_t)d * H + h) * kEdp + moff, qwf);
#pragma unroll
  for (int i = 0; i < VEC; ++i) qf[i] *= qscale;
#pragma unroll
  for (int i = 0; i < MPL; ++i) qwf[i] *= qscale;

  const int beg = rowptr[d], end = rowptr[d + 1];
  float m = -INFINITY, l = 0.f, acc[VEC], rr[MPL];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
#pragma unroll
  for (int i = 0; i < MPL; ++i) rr[i] = 0.f;

  int jn[kU], tn[kU];
#pragma unroll
  for (int u = 0; u < kU; ++u) {
    jn[u] = beg + u < end ? col[beg + u] : 0;
    tn[u] = beg + u < end ? perm[beg + u] : 0;
  }
  for (int p = beg; p < end; p += kU) {
    uint4 kr[kU], vr[kU];
    float rw[kU][MPL];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      if (p + u < end) {
        const size_t j = (size_t)jn[u], t = (size_t)tn[u];
        kr[u] = ldg16_keep(k + j * D + off);
        vr[u] = ldg16_keep(v + j * D + off);
        load_cols<MPL>(raw + t * kEdp + moff, rw[u]);
      } else {
        kr[u] = vr[u] = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int i = 0; i < MPL; ++i) rw[u][i] = 0.f;
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
      float kf[VEC];
      unpack<T>(kr[u], kf);
      float part = 0.f;
#pragma unroll
      for (int i = 0; i < VEC; ++i) part = fmaf(qf[i], kf[i], part);
#pragma unroll
      for (int i = 0; i < MPL; ++i) part = fmaf(qwf[i], rw[u][i], part);
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
    const float corr = fast_exp2(m - mn);
    l *= corr;
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] *= corr;
#pragma unroll
    for (int i = 0; i < MPL; ++i) rr[i] *= corr;
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const float pw = fast_exp2(s[u] - mn);
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
  const float inv = 1.f / (l + 1e-16f);
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[i] *= inv;
  stg16(out + (size_t)d * D + off, pack<T>(acc));
  float* Rp = R + ((size_t)d * H + h) * kEdp + moff;
#pragma unroll
  for (int i = 0; i < MPL; ++i) Rp[i] = rr[i] * inv;
  if (gl == 0) lse2[(size_t)d * H + h] = end > beg ? m + log2f(l + 1e-16f) : 0.f;
}

// backward, dst pass: dq_part_i = sum_t ds_t k_j / sqrt(C) (the W S_i term is added by the caller), S_i = sum_t ds_t raw_t / sqrt(C),
// and (a, ds/sqrt(C)) per (edge, head) at the edge's src-sorted position for the src pass and the raw-gradient kernel
template <typename T, int LPH>
__global__ void __launch_bounds__(kFoldThreads, 4)
gtconv_fold_bwd_dst_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, const float* __restrict__ raw,
                           const float* __restrict__ qw, const float* __restrict__ gw, const int* __restrict__ rowptr,
                           const int* __restrict__ col, const int* __restrict__ perm, const int* __restrict__ csr2csc, int Nd,
                           RowMap rm, int H, float qscale, float scale, const T* __restrict__ out, const float* __restrict__ lse2,
                           const T* __restrict__ g, T* __restrict__ dq, float* __restrict__ S, float2* __restrict__ ads) {
  constexpr int VEC = Vec<T>::N;
  constexpr int MPL = kEdp / LPH;
  const int lr = threadIdx.x / rm.tpd;
  const int d = blockIdx.x * rm.rpb + lr;
  if (lr >= rm.rpb || d >= Nd) return;
  const int chunk = blockIdx.y * rm.tpd + (threadIdx.x - lr * rm.tpd);
  const size_t D = (size_t)rm.chunks * VEC;
  const size_t off = (size_t)chunk * VEC;
  const unsigned mask = fold_group_mask<LPH>();
  const int h = chunk / LPH, gl = chunk & (LPH - 1);
  const int moff = gl * MPL;
  const bool leader = gl == 0;

  float qf[VEC], gf[VEC], dqa[VEC], qwf[MPL], gwf[MPL], ss[MPL];
  unpack<T>(ldg16_keep(q + (size_t)d * D + off), qf);
  unpack<T>(ldg16_keep(g + (size_t)d * D + off), gf);
  load_cols<MPL>(qw + ((size_t)d * H + h) * kEdp + moff, qwf);
  load_cols<MPL>(gw + ((size_t)d * H + h) * kEdp + moff, gwf);
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
#pragma unroll
  for (int i = 0; i < MPL; ++i) ss[i] = 0.f;

  const int beg = rowptr[d], end = rowptr[d + 1];
  int jn[kU], tn[kU], cn[kU];
#pragma unroll
  for (int u = 0; u < kU; ++u) {
    jn[u] = beg + u < end ? col[beg + u] : 0;
    tn[u] = beg + u < end ? perm[beg + u] : 0;
    cn[u] = beg + u < end ? csr2csc[beg + u] : 0;
  }
  for (int p = beg; p < end; p += kU) {
    uint4 kr[kU], vr[kU];
    float rw[kU][MPL];
    size_t cs[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      cs[u] = (size_t)cn[u];
      if (p + u < end) {
        const size_t j = (size_t)jn[u], t = (size_t)tn[u];
        kr[u] = ldg16_keep(k + j * D + off);
        vr[u] = ldg16_keep(v + j * D + off);
        load_cols<MPL>(raw + t * kEdp + moff, rw[u]);
      } else {
        kr[u] = vr[u] = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int i = 0; i < MPL; ++i) rw[u][i] = 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int pn = p + kU + u;
      jn[u] = pn < end ? col[pn] : 0;
      tn[u] = pn < end ? perm[pn] : 0;
      cn[u] = pn < end ? csr2csc[pn] : 0;
    }
    float s[kU], gv[kU];
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
        const float dss = a * (gv[u] - Dl) * scale;
        float kf[VEC];
        unpack<T>(kr[u], kf);
#pragma unroll
        for (int i = 0; i < VEC; ++i) dqa[i] = fmaf(dss, kf[i], dqa[i]);
#pragma unroll
        for (int i = 0; i < MPL; ++i) ss[i] = fmaf(dss, rw[u][i], ss[i]);
        if (leader) ads[cs[u] * H + h] = make_float2(a, dss);
      }
    }
  }
  if (dq) stg16(dq + (size_t)d * D + off, pack<T>(dqa));
  float* Sp = S + ((size_t)d * H + h) * kEdp + moff;
#pragma unroll
  for (int i = 0; i < MPL; ++i) Sp[i] = ss[i];
}

// d raw_t[m] = sum_h a_t,h gw_i[h][m] + (ds_t,h / sqrt(C)) qw_i[h][m]   (i = dst of edge t).  One warp per dst row: lane = (edge slot, m),
// two edges at a time; gw_i / qw_i (2 x H x 64 bytes) stay in L1 across the row's edges.
__global__ void __launch_bounds__(128)
edge_raw_grad_kernel(const float2* __restrict__ ads, const float* __restrict__ qw, const float* __restrict__ gw,
                     const int* __restrict__ rowptr, const int* __restrict__ perm, const int* __restrict__ csr2csc, int Nd, int H,
                     float* __restrict__ draw) {
  const int d = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (d >= Nd) return;
  const int lane = threadIdx.x & 31, m = lane & 15, slot = lane >> 4;
  const float* qwd = qw + (size_t)d * H * kEdp + m;
  const float* gwd = gw + (size_t)d * H * kEdp + m;
  const int beg = rowptr[d], end = rowptr[d + 1];
  for (int p = beg + slot; p < end; p += 2) {
    const size_t c = (size_t)csr2csc[p];
    float acc = 0.f;
    for (int h = 0; h < H; ++h) {
      const float2 w = __ldg(ads + c * H + h);
      acc = fmaf(w.x, __ldg(gwd + h * kEdp), acc);
      acc = fmaf(w.y, __ldg(qwd + h * kEdp), acc);
    }
    draw[(size_t)perm[p] * kEdp + m] = acc;
  }
}

struct FoldPlan {
  bool ok;
  int lph;
  RowMap rm;
  int slices;
};

FoldPlan make_fold_plan(int H, int C, int elt) {
  FoldPlan pl{};
  const int head_bytes = C * elt;
  if (head_bytes % 16 != 0) return pl;
  const int lph = head_bytes / 16;
  if (lph > kEdp || (lph & (lph - 1)) != 0) return pl;  // every lane of a head group owns >= 1 raw column
  int hs = 1;
  for (int c = 1; c <= H; ++c)
    if (H % c == 0 && c * lph <= kFoldThreads) hs = c;
  pl.ok = true;
  pl.lph = lph;
  pl.rm.tpd = hs * lph;
  pl.rm.rpb = kFoldThreads / pl.rm.tpd;
  pl.rm.chunks = H * lph;
  pl.slices = H / hs;
  return pl;
}

struct FoldArgs {
  const void *q, *k, *v, *out, *g;
  const float *raw, *qw, *gw, *lse2_in;
  const int *rowptr, *col, *perm, *csr2csc;
  int Nd, H;
  float qscale, scale;
  void *out_w, *dq;
  float *lse2_w, *R, *S;
  float2* ads;
  cudaStream_t st;
};

template <typename T, int LPH>
void launch_fold(int which, const FoldPlan& pl, const FoldArgs& a) {
  dim3 grid((a.Nd + pl.rm.rpb - 1) / pl.rm.rpb, pl.slices);
  if (which == 0)
    gtconv_fold_fwd_kernel<T, LPH><<<grid, kFoldThreads, 0, a.st>>>((const T*)a.q, (const T*)a.k, (const T*)a.v, a.raw, a.qw, a.rowptr, a.col,
                                                                  a.perm, a.Nd, pl.rm, a.H, a.qscale, (T*)a.out_w, a.lse2_w, a.R);
  else
    gtconv_fold_bwd_dst_kernel<T, LPH><<<grid, kFoldThreads, 0, a.st>>>((const T*)a.q, (const T*)a.k, (const T*)a.v, a.raw, a.qw, a.gw, a.rowptr,
                                                                      a.col, a.perm, a.csr2csc, a.Nd, pl.rm, a.H, a.qscale, a.scale,
                                                                      (const T*)a.out, a.lse2_in, (const T*)a.g, (T*)a.dq, a.S, a.ads);
}

template <typename T>
bool dispatch_fold(int which, const FoldPlan& pl, const FoldArgs& a) {
  switch (pl.lph) {
    case 1: launch_fold<T, 1>(which, pl, a); return true;
    case 2: launch_fold<T, 2>(which, pl, a); return true;
    case 4: launch_fold<T, 4>(which, pl, a); return true;
    case 8: launch_fold<T, 8>(which, pl, a); return true;
    case 16: launch_fold<T, 16>(which, pl, a); return true;
    default: return false;
  }
}

int check_fold(const char* fn, int dtype, int64_t Ns, int64_t Nd, int64_t E, int H, int C, FoldPlan* pl) {
  if (dtype != AB2_F32 && dtype != AB2_BF16) return fail(AB2_ERR_INVALID, "%s: dtype must be AB2_F32 or AB2_BF16", fn);
  if (Ns < 0 || Nd < 0 || E < 0 || H <= 0 || C <= 0) return fail(AB2_ERR_INVALID, "%s: negative size", fn);
  if (Ns > INT32_MAX || Nd > INT32_MAX || E > INT32_MAX) return fail(AB2_ERR_INVALID, "%s: sizes must fit in int32", fn);
  *pl = make_fold_plan(H, C, dtype == AB2_F32 ? 4 : 2);
  if (!pl->ok) return fail(AB2_ERR_UNSUPPORTED, "%s: head width C=%d is not supported by the folded kernels (C*sizeof must be 16..256 bytes, power of two)", fn, C);
  return 0;
}

}  // namespace
}  // namespace ab2

using namespace ab2;

extern "C" int ab2_gtconv_fold_fwd(const void* q, const void* k, const void* v, const float* raw, const float* qw, int dtype,
                                   const int32_t* rowptr, const int32_t* col, const int32_t* perm, int64_t Ns, int64_t Nd, int64_t E,
                                   int H, int C, void* out, float* lse2, float* R, void* stream) {
  FoldPlan pl;
  if (int rc = check_fold("gtconv_fold_fwd", dtype, Ns, Nd, E, H, C, &pl)) return rc;
  if (Nd == 0) return AB2_OK;
  if (!q || !qw || !out || !lse2 || !R || !rowptr || (E > 0 && (!k || !v || !raw || !col || !perm)))
    return fail(AB2_ERR_INVALID, "gtconv_fold_fwd: null pointer argument");
  FoldArgs a{};
  a.q = q; a.k = k; a.v = v; a.raw = raw; a.qw = qw;
  a.rowptr = rowptr; a.col = col; a.perm = perm;
  a.Nd = (int)Nd; a.H = H;
  a.scale = 1.f / sqrtf((float)C);
  a.qscale = kLog2e * a.scale;
  a.out_w = out; a.lse2_w = lse2; a.R = R;
  a.st = (cudaStream_t)stream;
  const bool ok = dtype == AB2_F32 ? dispatch_fold<float>(0, pl, a) : dispatch_fold<__nv_bfloat16>(0, pl, a);
  if (!ok) return fail(AB2_ERR_UNSUPPORTED, "gtconv_fold_fwd: unsupported head layout");
  AB2_LAUNCH_OK("gtconv_fold_fwd_kernel");
  return AB2_OK;
}

extern "C" int ab2_gtconv_fold_bwd_dst(const void* q, const void* k, const void* v, const float* raw, const float* qw, const float* gw,
                                       int dtype, const int32_t* rowptr, const int32_t* col, const int32_t* perm,
                                       const int32_t* csr2csc, int64_t Ns, int64_t Nd, int64_t E, int H, int C, const void* out,
                                       const float* lse2, const void* g, void* dq, float* S, void* ads, void* stream) {
  FoldPlan pl;
  if (int rc = check_fold("gtconv_fold_bwd_dst", dtype, Ns, Nd, E, H, C, &pl)) return rc;
  if (Nd == 0) return AB2_OK;
  if (!q || !qw || !gw || !out || !lse2 || !g || !S || !rowptr || (E > 0 && (!k || !v || !raw || !col || !perm || !csr2csc || !ads)))
    return fail(AB2_ERR_INVALID, "gtconv_fold_bwd_dst: null pointer argument");
  FoldArgs a{};
  a.q = q; a.k = k; a.v = v; a.raw = raw; a.qw = qw; a.gw = gw;
  a.rowptr = rowptr; a.col = col; a.perm = perm; a.csr2csc = csr2csc;
  a.Nd = (int)Nd; a.H = H;
  a.scale = 1.f / sqrtf((float)C);
  a.qscale = kLog2e * a.scale;
  a.out = out; a.lse2_in = lse2; a.g = g; a.dq = dq; a.S = S; a.ads = (float2*)ads;
  a.st = (cudaStream_t)stream;
  const bool ok = dtype == AB2_F32 ? dispatch_fold<float>(1, pl, a) : dispatch_fold<__nv_bfloat16>(1, pl, a);
  if (!ok) return fail(AB2_ERR_UNSUPPORTED, "gtconv_fold_bwd_dst: unsupported head layout");
  AB2_LAUNCH_OK("gtconv_fold_bwd_dst_kernel");
  return AB2_OK;
}

extern "C" int ab2_edge_raw_grad(const void* ads, const float* qw, const float* gw, const int32_t* rowptr, const int32_t* perm,
                                 const int32_t* csr2csc, int64_t Nd, int64_t E, int H, float* draw, void* stream) {
  if (Nd < 0 || E < 0 || H <= 0 || Nd > INT32_MAX || E > INT32_MAX) return fail(AB2_ERR_INVALID, "edge_raw_grad: bad size");
  if (Nd == 0 || E == 0) return AB2_OK;
  if (!ads || !qw || !gw || !rowptr || !perm || !csr2csc || !draw) return fail(AB2_ERR_INVALID, "edge_raw_grad: null pointer argument");
  edge_raw_grad_kernel<<<(unsigned)((Nd + 3) / 4), 128, 0, (cudaStream_t)stream>>>((const float2*)ads, qw, gw, rowptr, perm, csr2csc, (int)Nd, H, draw);
  AB2_LAUNCH_OK("edge_raw_grad_kernel");
  return AB2_OK;
}
