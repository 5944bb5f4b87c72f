// Host-buffer entry point: GraphTransformerConv forward+backward with operands in pinned host memory.
// This is the call a reference-side plugin holding CPU tensors makes (reference layers/conv.py:98 forward and
// its autograd backward); H2D / D2H copies are part of the call and overlap the kernels where the data flow allows:
//   copy stream : k, v, q, e  -> (event) -> fwd ; g -> (event) -> bwd
//   out stream  : out leaves while bwd_dst runs; de, dq leave while bwd_src runs; dk, dv last.
#include <cstdlib>

#include "common.cuh"

using namespace ab2;

namespace {
struct HostWs {
  char *q, *k, *v, *e, *g, *out, *dq, *dk, *dv, *de, *lse2, *ads;
  size_t bytes;
};
HostWs carve_host(void* ws, int64_t Ns, int64_t Nd, int64_t E, int H, int C, int dtype) {
  auto al = [](size_t x) { return (x + 255) / 256 * 256; };
  const size_t elt = dtype == AB2_F32 ? 4 : 2, D = (size_t)H * C;
  const size_t nd = al((size_t)Nd * D * elt), ns = al((size_t)Ns * D * elt), ne = al((size_t)E * D * elt);
  HostWs w;
  char* b = (char*)ws;
  size_t off = 0;
  auto take = [&](size_t n) {
    char* p = b + off;
    off += n;
    return p;
  };
  w.q = take(nd); w.k = take(ns); w.v = take(ns); w.e = take(ne); w.g = take(nd);
  w.out = take(nd); w.dq = take(nd); w.dk = take(ns); w.dv = take(ns); w.de = take(ne);
  w.lse2 = take(al((size_t)Nd * H * 4));
  w.ads = take(al(ab2_gtconv_bwd_workspace_bytes(E, H)));
  w.bytes = off;
  return w;
}
}  // namespace

extern "C" size_t ab2_gtconv_host_workspace_bytes(int64_t Ns, int64_t Nd, int64_t E, int H, int C, int dtype) {
  return carve_host(nullptr, Ns, Nd, E, H, C, dtype).bytes;
}

extern "C" int ab2_gtconv_fwd_bwd_host(const void* q_host, const void* k_host, const void* v_host, const void* e_host,
                                       const void* g_host, int dtype, const int32_t* rowptr, const int32_t* col,
                                       const int32_t* perm, const int32_t* colptr, const int32_t* csr2csc, const int32_t* crow,
                                       int64_t Ns, int64_t Nd, int64_t E, int H, int C, void* out_host, void* dq_host,
                                       void* dk_host, void* dv_host, void* de_host, void* dev_ws, size_t dev_ws_bytes,
                                       void* stream) {
  if (dtype != AB2_F32 && dtype != AB2_BF16) return fail(AB2_ERR_INVALID, "gtconv_fwd_bwd_host: bad dtype");
  if (!q_host || !k_host || !v_host || !e_host || !g_host || !out_host || !dev_ws)
    return fail(AB2_ERR_INVALID, "gtconv_fwd_bwd_host: null pointer argument");
  if (dev_ws_bytes < ab2_gtconv_host_workspace_bytes(Ns, Nd, E, H, C, dtype))
    return fail(AB2_ERR_INVALID, "gtconv_fwd_bwd_host: device workspace too small");
  const HostWs w = carve_host(dev_ws, Ns, Nd, E, H, C, dtype);
  const size_t elt = dtype == AB2_F32 ? 4 : 2, D = (size_t)H * C;
  const size_t nd = (size_t)Nd * D * elt, ns = (size_t)Ns * D * elt, ne = (size_t)E * D * elt;
  cudaStream_t comp = (cudaStream_t)stream, cin = nullptr, cout = nullptr;
  cudaEvent_t ev_in = nullptr, ev_g = nullptr, ev_fwd = nullptr, ev_bwd = nullptr;
  int rc = AB2_OK;
#define TRY(expr)                                                                                                      \
  do {                                                                                                                 \
    cudaError_t _e = (expr);                                                                                           \
    if (_e != cudaSuccess && rc == AB2_OK) rc = fail(AB2_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)
  TRY(cudaStreamCreateWithFlags(&cin, cudaStreamNonBlocking));
  TRY(cudaStreamCreateWithFlags(&cout, cudaStreamNonBlocking));
  TRY(cudaEventCreateWithFlags(&ev_in, cudaEventDisableTiming));
  TRY(cudaEventCreateWithFlags(&ev_g, cudaEventDisableTiming));
  TRY(cudaEventCreateWithFlags(&ev_fwd, cudaEventDisableTiming));
  TRY(cudaEventCreateWithFlags(&ev_bwd, cudaEventDisableTiming));
  if (rc == AB2_OK) {
    // inputs of the forward first, the upstream gradient behind them
    TRY(cudaMemcpyAsync(w.q, q_host, nd, cudaMemcpyHostToDevice, cin));
    TRY(cudaMemcpyAsync(w.k, k_host, ns, cudaMemcpyHostToDevice, cin));
    TRY(cudaMemcpyAsync(w.v, v_host, ns, cudaMemcpyHostToDevice, cin));
    TRY(cudaMemcpyAsync(w.e, e_host, ne, cudaMemcpyHostToDevice, cin));
    TRY(cudaEventRecord(ev_in, cin));
    TRY(cudaMemcpyAsync(w.g, g_host, nd, cudaMemcpyHostToDevice, cin));
    TRY(cudaEventRecord(ev_g, cin));
    TRY(cudaStreamWaitEvent(comp, ev_in, 0));
  }
  if (rc == AB2_OK)
    rc = ab2_gtconv_fwd(w.q, w.k, w.v, w.e, dtype, rowptr, col, perm, Ns, Nd, E, H, C, w.out, (float*)w.lse2, comp);
  if (rc == AB2_OK) {
    TRY(cudaEventRecord(ev_fwd, comp));
    TRY(cudaStreamWaitEvent(cout, ev_fwd, 0));
    TRY(cudaMemcpyAsync(out_host, w.out, nd, cudaMemcpyDeviceToHost, cout));
    TRY(cudaStreamWaitEvent(comp, ev_g, 0));
  }
  if (rc == AB2_OK)
    rc = ab2_gtconv_bwd(w.q, w.k, w.v, w.e, dtype, rowptr, col, perm, colptr, csr2csc, crow, Ns, Nd, E, H, C, w.out,
                        (const float*)w.lse2, w.g, dq_host ? w.dq : nullptr, dk_host ? w.dk : nullptr,
                        dv_host ? w.dv : nullptr, de_host ? w.de : nullptr, w.ads, ab2_gtconv_bwd_workspace_bytes(E, H), comp);
  if (rc == AB2_OK) {
    TRY(cudaEventRecord(ev_bwd, comp));
    TRY(cudaStreamWaitEvent(cout, ev_bwd, 0));
    if (de_host) TRY(cudaMemcpyAsync(de_host, w.de, ne, cudaMemcpyDeviceToHost, cout));
    if (dq_host) TRY(cudaMemcpyAsync(dq_host, w.dq, nd, cudaMemcpyDeviceToHost, cout));
    if (dk_host) TRY(cudaMemcpyAsync(dk_host, w.dk, ns, cudaMemcpyDeviceToHost, cout));
    if (dv_host) TRY(cudaMemcpyAsync(dv_host, w.dv, ns, cudaMemcpyDeviceToHost, cout));
  }
  // the call owns its streams: wait for everything (also on the error path, so nothing is destroyed in flight)
  if (cin) TRY(cudaStreamSynchronize(cin));
  TRY(cudaStreamSynchronize(comp));
  if (cout) TRY(cudaStreamSynchronize(cout));
  if (ev_in) cudaEventDestroy(ev_in);
  if (ev_g) cudaEventDestroy(ev_g);
  if (ev_fwd) cudaEventDestroy(ev_fwd);
  if (ev_bwd) cudaEventDestroy(ev_bwd);
  if (cin) cudaStreamDestroy(cin);
  if (cout) cudaStreamDestroy(cout);
#undef TRY
  return rc;
}

// ------------------------------------------------------------------------------------------------------------------
// Streamed variant: the dst rows are cut into chunks; chunk c+1 is uploaded while chunk c computes and chunk c-1's
// results download, so H2D and D2H (full-duplex PCIe) overlap and the kernels hide behind the copies.
// Needs a dst-sorted edge list (perm == identity: the e / de rows of a chunk are one contiguous slice).
// meta_host: int64 [nchunks][8] = {d0, d1, p0, p1, smax (largest src id referenced by chunks <= c, -1 if none),
//                                  src_final (src rows < src_final have no edge in chunks > c), 0, 0}
// ------------------------------------------------------------------------------------------------------------------
extern "C" int ab2_gtconv_fwd_bwd_host_streamed(const void* q_host, const void* k_host, const void* v_host, const void* e_host,
                                                const void* g_host, int dtype, const int32_t* rowptr, const int32_t* col,
                                                const int32_t* perm, const int32_t* colptr, const int32_t* csr2csc,
                                                const int32_t* crow, int64_t Ns, int64_t Nd, int64_t E, int H, int C,
                                                void* out_host, void* dq_host, void* dk_host, void* dv_host, void* de_host,
                                                const int64_t* meta_host, int nchunks, void* dev_ws, size_t dev_ws_bytes,
                                                void* stream) {
  if (dtype != AB2_F32 && dtype != AB2_BF16) return fail(AB2_ERR_INVALID, "gtconv_fwd_bwd_host_streamed: bad dtype");
  if (!q_host || !k_host || !v_host || !e_host || !g_host || !out_host || !dq_host || !dk_host || !dv_host || !de_host ||
      !dev_ws || !meta_host || nchunks < 1)
    return fail(AB2_ERR_INVALID, "gtconv_fwd_bwd_host_streamed: null pointer argument / bad chunk count");
  if (dev_ws_bytes < ab2_gtconv_host_workspace_bytes(Ns, Nd, E, H, C, dtype))
    return fail(AB2_ERR_INVALID, "gtconv_fwd_bwd_host_streamed: device workspace too small");
  const HostWs w = carve_host(dev_ws, Ns, Nd, E, H, C, dtype);
  const size_t elt = dtype == AB2_F32 ? 4 : 2, D = (size_t)H * C, row = D * elt;
  const size_t ads_bytes = ab2_gtconv_bwd_workspace_bytes(E, H);
  cudaStream_t comp = (cudaStream_t)stream, cin = nullptr, cout = nullptr;
  int rc = AB2_OK;
#define TRY(expr)                                                                                                      \
  do {                                                                                                                 \
    cudaError_t _e = (expr);                                                                                           \
    if (_e != cudaSuccess && rc == AB2_OK) rc = fail(AB2_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)
  TRY(cudaStreamCreateWithFlags(&cin, cudaStreamNonBlocking));
  TRY(cudaStreamCreateWithFlags(&cout, cudaStreamNonBlocking));
  const int nev = 3 * nchunks;
  cudaEvent_t* ev = (cudaEvent_t*)calloc(nev, sizeof(cudaEvent_t));
  if (!ev) rc = fail(AB2_ERR_INVALID, "gtconv_fwd_bwd_host_streamed: out of host memory");
  for (int i = 0; i < nev && rc == AB2_OK; ++i) TRY(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
  auto H2D = [&](char* dst, const void* src_base, size_t lo, size_t hi) {
    if (hi > lo) TRY(cudaMemcpyAsync(dst + lo * row, (const char*)src_base + lo * row, (hi - lo) * row, cudaMemcpyHostToDevice, cin));
  };
  auto D2H = [&](void* dst_base, const char* src, size_t lo, size_t hi) {
    if (hi > lo) TRY(cudaMemcpyAsync((char*)dst_base + lo * row, src + lo * row, (hi - lo) * row, cudaMemcpyDeviceToHost, cout));
  };
  int64_t kv_hi = 0, src_done = 0;
  for (int c = 0; c < nchunks && rc == AB2_OK; ++c) {
    const int64_t* m = meta_host + (size_t)c * 8;
    const int64_t d0 = m[0], d1 = m[1], p0 = m[2], p1 = m[3], smax = m[4];
    const int64_t src_final = c == nchunks - 1 ? Ns : m[5];
    cudaEvent_t ev_in = ev[3 * c], ev_dst = ev[3 * c + 1], ev_src = ev[3 * c + 2];
    // ---- upload: the k/v rows this chunk references beyond what is already resident, then its q, g, e slices
    if (smax + 1 > kv_hi) {
      H2D(w.k, k_host, (size_t)kv_hi, (size_t)(smax + 1));
      H2D(w.v, v_host, (size_t)kv_hi, (size_t)(smax + 1));
      kv_hi = smax + 1;
    }
    H2D(w.q, q_host, (size_t)d0, (size_t)d1);
    H2D(w.g, g_host, (size_t)d0, (size_t)d1);
    H2D(w.e, e_host, (size_t)p0, (size_t)p1);
    TRY(cudaEventRecord(ev_in, cin));
    // ---- compute: forward and the dst pass of the backward on rows [d0, d1)
    TRY(cudaStreamWaitEvent(comp, ev_in, 0));
    if (rc == AB2_OK && d1 > d0) {
      rc = ab2_gtconv_fwd(w.q + d0 * row, w.k, w.v, w.e, dtype, rowptr + d0, col, perm, Ns, d1 - d0, E, H, C, w.out + d0 * row,
                          (float*)w.lse2 + d0 * H, comp);
      if (rc == AB2_OK)
        rc = ab2_gtconv_bwd_dst(w.q + d0 * row, w.k, w.v, w.e, dtype, rowptr + d0, col, perm, csr2csc, Ns, d1 - d0, E, H, C,
                                w.out + d0 * row, (const float*)w.lse2 + d0 * H, w.g + d0 * row, w.dq + d0 * row, w.de, w.ads,
                                ads_bytes, comp);
    }
    TRY(cudaEventRecord(ev_dst, comp));
    // ---- src pass for the src rows whose edges are all behind us
    if (rc == AB2_OK && src_final > src_done) {
      rc = ab2_gtconv_bwd_src_range(w.q, w.g, dtype, colptr, crow, Ns, Nd, E, H, C, w.ads, w.dk, w.dv, src_done, src_final, comp);
    }
    TRY(cudaEventRecord(ev_src, comp));
    // ---- download
    TRY(cudaStreamWaitEvent(cout, ev_dst, 0));
    D2H(out_host, w.out, (size_t)d0, (size_t)d1);
    D2H(dq_host, w.dq, (size_t)d0, (size_t)d1);
    D2H(de_host, w.de, (size_t)p0, (size_t)p1);
    if (src_final > src_done) {
      TRY(cudaStreamWaitEvent(cout, ev_src, 0));
      D2H(dk_host, w.dk, (size_t)src_done, (size_t)src_final);
      D2H(dv_host, w.dv, (size_t)src_done, (size_t)src_final);
      src_done = src_final;
    }
  }
  if (cin) TRY(cudaStreamSynchronize(cin));
  TRY(cudaStreamSynchronize(comp));
  if (cout) TRY(cudaStreamSynchronize(cout));
  if (ev) {
    for (int i = 0; i < nev; ++i)
      if (ev[i]) cudaEventDestroy(ev[i]);
    free(ev);
  }
  if (cin) cudaStreamDestroy(cin);
  if (cout) cudaStreamDestroy(cout);
#undef TRY
  return rc;
}
