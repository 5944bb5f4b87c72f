/*
 * anemoi_b200.h -- C ABI of libanemoi_b200.so (sm_100a).
 *
 * Drop-in boundary for the graph message-passing hot path of ecmwf/anemoi-models.  The reference is pure
 * Python (no FFI of its own); each entry point below names the reference interface it replaces
 * (paths relative to src/anemoi/models/ of the reference tree).  The Python host side
 * (anemoi_models_b200/) binds these with ctypes; INTEGRATION.md shows the stub a reference maintainer adds.
 *
 * Conventions
 *   - Every pointer is a DEVICE pointer on the current CUDA device unless the name ends in `_host`.
 *   - The caller allocates every input, output and workspace; the library never allocates, frees or
 *     retains device memory across calls.  (`*_host` entry points own a small internal staging pool.)
 *   - All work is enqueued on `stream` (a cudaStream_t passed as void*); no hidden synchronisation,
 *     except where a function's comment says it synchronises.
 *   - Return value 0 = success; non-zero = ab2_status.  ab2_last_error() returns a thread-local message.
 *   - Row-major, contiguous tensors.  dtype: AB2_F32 (fp32 in, fp32 out) or AB2_BF16 (bf16 in/out, fp32 math).
 */
#ifndef ANEMOI_B200_H
#define ANEMOI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum { AB2_F32 = 0, AB2_BF16 = 1 } ab2_dtype;

typedef enum {
  AB2_OK = 0,
  AB2_ERR_INVALID = 1,     /* bad argument (shape, dtype, null pointer, workspace too small) */
  AB2_ERR_UNSUPPORTED = 2, /* shape outside what the kernels cover */
  AB2_ERR_CUDA = 3         /* a CUDA runtime call / launch failed */
} ab2_status;

/* Library version (major*10000 + minor*100 + patch). */
int ab2_version(void);
/* Thread-local description of the last non-zero status returned on this thread. */
const char* ab2_last_error(void);
/* Number of CUDA kernels this library has launched in this process so far (benchmark bookkeeping: gpu_launches). */
long long ab2_launch_count(void);

/* ---------------------------------------------------------------------------------------------------
 * dst-sorted CSR (+ src-sorted CSC view) of edge_index.
 * Replaces: PyG MessagePassing._collect/_lift index_select by edge_index[0|1] and the scatter-by-index
 * aggregate (called from layers/conv.py:64,110), i.e. the gather/scatter plan of every conv call.
 * One-off per graph (edge_index is a constant buffer, layers/mapper.py:144-148).
 *
 * edge_index : int64 [2,E], row 0 = src (j), row 1 = dst (i)  (flow source_to_target)
 * rowptr     : int32 [Nd+1]   segment offsets of each dst in the sorted order
 * col        : int32 [E]      src id of sorted position p
 * perm       : int32 [E]      original edge id of sorted position p  (STABLE: ascending within a dst)
 * rowidx     : int32 [E]      dst id of sorted position p
 * colptr     : int32 [Ns+1]   segment offsets of each src in the src-sorted order
 * cpos       : int32 [E]      CSR position p of src-sorted position t (ascending within a src)
 * crow       : int32 [E,2]    (dst id, src id) of src-sorted position t
 * csr2csc    : int32 [E]      src-sorted position t of CSR position p (inverse of cpos); may be NULL
 * flags      : int32 [4]      [0] = 1 if perm is the identity; [1] = number of edges with src/dst out of range
 * Requires E, Ns, Nd < 2^31.  Result is bit-exact equal to torch.sort(edge_index[1], stable=True).
 * ------------------------------------------------------------------------------------------------- */
size_t ab2_csr_workspace_bytes(int64_t E, int64_t Ns, int64_t Nd);
int ab2_csr_build(const int64_t* edge_index, int64_t E, int64_t Ns, int64_t Nd, int32_t* rowptr, int32_t* col,
                  int32_t* perm, int32_t* rowidx, int32_t* colptr, int32_t* cpos, int32_t* crow, int32_t* csr2csc,
                  int32_t* flags, void* workspace, size_t workspace_bytes, void* stream);

/* Stable partition of edges by dst chunk.
 * Replaces: distributed/khop_edges.py:88-130 sort_edges_1hop_chunks / :50-85 sort_edges_1hop_sharding
 * (PyG bipartite_subgraph / k_hop_subgraph masks).  chunk c owns dst rows [bounds_host[c], bounds_host[c+1]).
 * bounds_host : HOST int64 [num_chunks+1]
 * order       : int64 [E]  original edge ids, grouped by chunk, original order inside a chunk
 * counts      : int64 [num_chunks] (device)
 * Edges whose dst lies in no chunk are dropped (their slots at the tail of `order` stay untouched). */
size_t ab2_edge_chunks_workspace_bytes(int64_t E);
int ab2_edge_chunks(const int64_t* edge_index, int64_t E, const int64_t* bounds_host, int num_chunks, int64_t* order,
                    int64_t* counts, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * GraphTransformerConv forward.
 * Replaces: layers/conv.py:98-142 GraphTransformerConv.forward/message + PyG propagate, utils.softmax and
 * the "add" aggregate:  out_i = sum_t softmax_i(q_i.(k_j+e_t)/sqrt(C))_t * (v_j+e_t)  over edges t=(j->i).
 *
 * q [Nd,H,C]; k,v [Ns,H,C]; e [E,H,C] in ORIGINAL edge order (row perm[p] belongs to sorted position p);
 * out [Nd,H,C]; lse2 [Nd,H] fp32 = log2(sum_t 2^(s_t*log2e) + 1e-16) (base-2 log-sum-exp, kept for backward).
 * dst without incoming edges -> zero row (PyG scatter into zeros).
 * ------------------------------------------------------------------------------------------------- */
int ab2_gtconv_fwd(const void* q, const void* k, const void* v, const void* e, int dtype, const int32_t* rowptr,
                   const int32_t* col, const int32_t* perm, int64_t Ns, int64_t Nd, int64_t E, int H, int C,
                   void* out, float* lse2, void* stream);

/* GraphTransformerConv backward (the autograd graph of the ops above in the reference).
 * g = d(out) [Nd,H,C].  Writes dq [Nd,H,C], dk,dv [Ns,H,C], de [E,H,C] (original edge order); any of the four
 * may be NULL to skip it.  workspace: ab2_gtconv_bwd_workspace_bytes(E,H) bytes (per-edge, per-head softmax
 * weight and logit gradient, 8 B each).  Deterministic: no atomics. */
size_t ab2_gtconv_bwd_workspace_bytes(int64_t E, int H);
/* The two passes of the backward, callable on their own (ab2_gtconv_bwd = dst pass, then src pass):
 *   dst pass (per dst segment): dq, de and the per-(edge, head) pair (softmax weight, logit gradient / sqrt(C)),
 *            stored at the edge's SRC-SORTED position csr2csc[p] of ads_ws so that the src pass reads it contiguously;
 *   src pass (per block of consecutive src rows, CSC order): dk_j = sum ds * q_i, dv_j = sum a * g_i. */
int ab2_gtconv_bwd_dst(const void* q, const void* k, const void* v, const void* e, int dtype, const int32_t* rowptr,
                       const int32_t* col, const int32_t* perm, const int32_t* csr2csc, int64_t Ns, int64_t Nd, int64_t E,
                       int H, int C, const void* out, const float* lse2, const void* g, void* dq, void* de, void* ads_ws,
                       size_t ads_ws_bytes, void* stream);
int ab2_gtconv_bwd_src(const void* q, const void* g, int dtype, const int32_t* colptr, const int32_t* crow, int64_t Ns,
                       int64_t Nd, int64_t E, int H, int C, const void* ads_ws, void* dk, void* dv, void* stream);
/* src pass restricted to src rows [row_begin, row_end) (all pointers unshifted) -- lets a caller that streams dst chunks
 * finalise dk / dv for the src rows whose edges are all behind it. */
int ab2_gtconv_bwd_src_range(const void* q, const void* g, int dtype, const int32_t* colptr, const int32_t* crow, int64_t Ns,
                             int64_t Nd, int64_t E, int H, int C, const void* ads_ws, void* dk, void* dv,
                             int64_t row_begin, int64_t row_end, void* stream);
int ab2_gtconv_bwd(const void* q, const void* k, const void* v, const void* e, int dtype, const int32_t* rowptr,
                   const int32_t* col, const int32_t* perm, const int32_t* colptr, const int32_t* csr2csc,
                   const int32_t* crow, int64_t Ns, int64_t Nd, int64_t E, int H, int C, const void* out,
                   const float* lse2, const void* g, void* dq, void* dk, void* dv, void* de, void* workspace,
                   size_t workspace_bytes, void* stream);

/* Name of the kernel a forward (which=0), backward dst pass (1) or backward src pass (2) with these shapes dispatches to
 * (thread-local string; for benchmark and profile bookkeeping). */
const char* ab2_gtconv_variant(int which, int dtype, int64_t Ns, int64_t Nd, int64_t E, int H, int C);

/* dst-row-sharded variants (one process per GPU).  The rank's edges index a COMPACT src space of Ns = n_own + n_halo
 * rows: [0, n_own) are the rank's own k / v rows (k, v), [n_own, Ns) the halo rows received from the peers (k_halo,
 * v_halo -- e.g. the receive buffer of an NCCL all-to-all, or peer memory mapped through NVLink).  No concatenation
 * copy is needed.  Backward writes the gradients of own rows to dk / dv and those of halo rows to dk_halo / dv_halo
 * (to be sent back to their owners).  Replaces the reference's head all-to-all (layers/block.py:366-414). */
int ab2_gtconv_fwd_halo(const void* q, const void* k, const void* v, const void* k_halo, const void* v_halo, int64_t n_own,
                        const void* e, int dtype, const int32_t* rowptr, const int32_t* col, const int32_t* perm,
                        int64_t Ns, int64_t Nd, int64_t E, int H, int C, void* out, float* lse2, void* stream);
/* The two backward passes with split src rows, on their own (used to overlap the halo-gradient exchange: dst pass on
 * the boundary dst rows, src pass on the halo rows, push, then the interior). */
int ab2_gtconv_bwd_dst_halo(const void* q, const void* k, const void* v, const void* k_halo, const void* v_halo,
                            int64_t n_own, const void* e, int dtype, const int32_t* rowptr, const int32_t* col,
                            const int32_t* perm, const int32_t* csr2csc, int64_t Ns, int64_t Nd, int64_t E, int H, int C,
                            const void* out, const float* lse2, const void* g, void* dq, void* de, void* ads_ws,
                            size_t ads_ws_bytes, void* stream);
int ab2_gtconv_bwd_src_range_halo(const void* q, const void* g, int dtype, const int32_t* colptr, const int32_t* crow,
                                  int64_t n_own, int64_t Ns, int64_t Nd, int64_t E, int H, int C, const void* ads_ws,
                                  void* dk, void* dv, void* dk_halo, void* dv_halo, int64_t row_begin, int64_t row_end,
                                  void* stream);
int ab2_gtconv_bwd_halo(const void* q, const void* k, const void* v, const void* k_halo, const void* v_halo, int64_t n_own,
                        const void* e, int dtype, const int32_t* rowptr, const int32_t* col, const int32_t* perm,
                        const int32_t* colptr, const int32_t* csr2csc, const int32_t* crow, int64_t Ns, int64_t Nd,
                        int64_t E, int H, int C, const void* out, const float* lse2, const void* g, void* dq, void* dk,
                        void* dv, void* dk_halo, void* dv_halo, void* de, void* workspace, size_t workspace_bytes,
                        void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Halo exchange over NVLink peer memory (one process per GPU on one node).  Replaces the list-form all_to_all of
 * distributed/transformer.py:21-82 / the all_gather of distributed/primitives.py:57-109 on this path.
 * Buffers that peers write into are whole cudaMalloc allocations owned by the library (CUDA IPC cannot export a
 * sub-allocation of a caching allocator): ab2_ipc_alloc / ab2_ipc_free; peers map them with ab2_ipc_open / _close.
 * ------------------------------------------------------------------------------------------------- */
#define AB2_MAX_PEERS 16
int ab2_ipc_alloc(size_t bytes, void** dev_ptr, void* handle_out /* 64 bytes */);
int ab2_ipc_open(const void* handle /* 64 bytes */, void** dev_ptr);
int ab2_ipc_close(void* dev_ptr);
int ab2_ipc_free(void* dev_ptr);
int ab2_memcpy_d2d(void* dst, const void* src, size_t bytes, void* stream);
/* For r in [0, n): row (src_row ? src_row[r] : r) of the local planes src_a / src_b (row_bytes each, src_b may be NULL)
 * is stored to row dst_row[r] of rank peer[r]'s planes plane_a[peer] / plane_b[peer] (HOST arrays of npeers device
 * pointers: IPC mappings of the peers' buffers, own entry = local buffer).  One kernel, every SM, posted NVLink writes. */
int ab2_peer_push_rows(const void* src_a, const void* src_b, const int32_t* src_row, const int32_t* peer,
                       const int32_t* dst_row, int64_t n, int row_bytes, void* const* plane_a, void* const* plane_b,
                       int npeers, void* stream);
/* The same push followed, in the same kernel, by a flag-word signal: once every CTA's stores are fenced at system scope the
 * last CTA stores `epoch` (st.release.sys) into this rank's slot of every peer's flag array (flag_slots: HOST array of npeers
 * device pointers, entry p = address of MY slot inside peer p's buffer; counter: a zero-initialised device word of this rank).
 * n may be 0 (signal only).  ab2_peer_wait_flags enqueues a one-warp kernel that spins (ld.acquire.sys, 20 s time-out -> trap)
 * until the local slots of all peers have reached `epoch`; kernels behind it in `stream` then see the pushed rows.
 * Replaces the stream-ordered 1-element NCCL all-reduce used as a barrier in round 1. */
int ab2_peer_push_rows_signal(const void* src_a, const void* src_b, const int32_t* src_row, const int32_t* peer,
                              const int32_t* dst_row, int64_t n, int row_bytes, void* const* plane_a, void* const* plane_b,
                              void* counter, void* const* flag_slots, uint32_t epoch, int npeers, int my_rank, void* stream);
int ab2_peer_wait_flags(const void* local_slots, uint32_t epoch, int npeers, int my_rank, void* stream);
/* dst[idx[s]] += src[s], s in [0, n); the ids of one call must be distinct (plain read-modify-write, fp32 add). */
int ab2_rows_add(void* dst, const int64_t* idx, const void* src, int64_t n, int D, int dtype, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * GraphConv edge path (layers/conv.py:61-76): the parts of
 *   edges_new = edge_mlp(cat[x_i, x_j, e]) + e ;  out = scatter_sum(edges_new, dst)
 * that are not plain GEMMs.  The first Linear(3D->D) is split as x_i Wi^T + x_j Wj^T + e We^T so the node
 * terms are computed once per node (pi = x_dst Wi^T + b0 [Nd,D], pj = x_src Wj^T [Ns,D]).
 * ------------------------------------------------------------------------------------------------- */
/* h0[t] = act(pi[dst_t] + pj[src_t] + pe[t]) for every edge t (original order); pre[t] optionally keeps the
 * pre-activation for backward.  act: 0 = SiLU, 1 = GELU(erf), 2 = ReLU, 3 = identity. */
int ab2_edge_gather_add_act(const void* pi, const void* pj, const void* pe, const int64_t* edge_index, int64_t E,
                            int64_t Ns, int64_t Nd, int D, int dtype, int act, void* h0, void* pre, void* stream);
/* Backward of the above: gpre[t] = g[t]*act'(pre[t]) (written to gpe, original order);
 * dpi[i] = sum over edges into i, dpj[j] = sum over edges out of j (CSR / CSC segment sums, deterministic). */
int ab2_edge_gather_add_act_bwd(const void* g, const void* pre, const int32_t* rowptr, const int32_t* perm,
                                const int32_t* colptr, const int32_t* cpos, int64_t E, int64_t Ns, int64_t Nd, int D,
                                int dtype, int act, void* gpe, void* dpi, void* dpj, void* stream);
/* edges_new[t] = LayerNorm(y[t]; gamma, beta, eps) + e[t];  out[i] = sum of edges_new over incoming edges.
 * Also writes mean/rstd [E] fp32 for backward. */
int ab2_edge_ln_res_segsum(const void* y, const void* e, const void* gamma, const void* beta, float eps,
                           const int32_t* rowptr, const int32_t* perm, int64_t E, int64_t Nd, int D, int dtype,
                           void* edges_new, void* out, float* mean, float* rstd, void* stream);
/* Backward: gt[t] = g_edges[t] + g_out[dst_t];  de[t] = gt[t];  dy[t] = LN-backward(gt[t]);
 * dgamma/dbeta [D] fp32 = column sums, reduced deterministically through `partial` ([nparts,2,D] fp32,
 * nparts = ab2_ln_bwd_parts() CTAs each own a fixed slice of the edges). */
int ab2_ln_bwd_parts(void);
int ab2_edge_ln_res_segsum_bwd(const void* g_edges /* may be NULL */, const void* g_out, const void* y,
                               const void* gamma, const float* mean, const float* rstd, const int64_t* edge_index,
                               int64_t E, int64_t Nd, int D, int dtype, void* dy, void* de, float* partial,
                               int nparts, float* dgamma, float* dbeta, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * ROUND-2 WORK IN PROGRESS (off by default, AB2_EDGE_FOLD=1; one green GPU run so far): GraphTransformerConv with the block's
 * `lin_edge` folded in.  Replaces reference layers/block.py:497 (`edges = self.lin_edge(edge_attr)`) + conv.py:98-142:
 * the conv takes the RAW edge features raw [E,16] (fp32; ed <= 15 columns + a constant-1 column that carries the bias,
 * zero padded) and per-dst projections qw = W_h^T q_i, gw = W_h^T g_i ([Nd,H,16] fp32, computed by the caller with one
 * small GEMM) instead of e [E,H,C]; it returns out_part = sum_t a_t v_j and R = sum_t a_t raw_t ([Nd,H,16]); the caller adds
 * W_h R.  Backward dst pass: dq_part = sum_t ds_t k_j / sqrt(C), S = sum_t ds_t raw_t / sqrt(C), and the (a, ds/sqrt(C)) workspace
 * `ads` [E,H] float2 (src-sorted order) that ab2_gtconv_bwd_src and ab2_edge_raw_grad consume.
 * ------------------------------------------------------------------------------------------------- */
int ab2_gtconv_fold_fwd(const void* q, const void* k, const void* v, const float* raw, const float* qw, int dtype,
                        const int32_t* rowptr, const int32_t* col, const int32_t* perm, int64_t Ns, int64_t Nd, int64_t E,
                        int H, int C, void* out, float* lse2, float* R, void* stream);
int ab2_gtconv_fold_bwd_dst(const void* q, const void* k, const void* v, const float* raw, const float* qw, const float* gw,
                            int dtype, const int32_t* rowptr, const int32_t* col, const int32_t* perm, const int32_t* csr2csc,
                            int64_t Ns, int64_t Nd, int64_t E, int H, int C, const void* out, const float* lse2, const void* g,
                            void* dq, float* S, void* ads, void* stream);
/* d raw_t[m] = sum_h a_t,h gw_i[h][m] + ads_t,h.y qw_i[h][m]  ->  draw [E,16] fp32 in ORIGINAL edge order (lin_edge's input gradient) */
int ab2_edge_raw_grad(const void* ads, const float* qw, const float* gw, const int32_t* rowptr, const int32_t* perm,
                      const int32_t* csr2csc, int64_t Nd, int64_t E, int H, float* draw, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Host-buffer entry point (what a reference-side plugin with CPU tensors calls): one GraphTransformerConv
 * forward+backward with q,k,v,e,g in PINNED HOST memory and out,dq,dk,dv,de written back to host memory.
 * Device staging buffers are supplied by the caller (dev_ws, >= ab2_gtconv_host_workspace_bytes()).
 * Copies are chunked and overlapped with the kernels on internal streams; the call returns after everything
 * has landed in host memory (it synchronises).
 * ------------------------------------------------------------------------------------------------- */
size_t ab2_gtconv_host_workspace_bytes(int64_t Ns, int64_t Nd, int64_t E, int H, int C, int dtype);
int ab2_gtconv_fwd_bwd_host(const void* q_host, const void* k_host, const void* v_host, const void* e_host,
                            const void* g_host, int dtype, const int32_t* rowptr, const int32_t* col,
                            const int32_t* perm, const int32_t* colptr, const int32_t* csr2csc, const int32_t* crow,
                            int64_t Ns, int64_t Nd, int64_t E, int H, int C, void* out_host, void* dq_host,
                            void* dk_host, void* dv_host, void* de_host, void* dev_ws, size_t dev_ws_bytes,
                            void* stream);

/* Streamed variant of the call above for dst-sorted edge lists (perm == identity): the dst rows are cut into
 * `nchunks` pieces; chunk c+1 uploads while chunk c computes and chunk c-1 downloads, so the two PCIe directions overlap.
 * meta_host : HOST int64 [nchunks][8] = {d0, d1, p0, p1, smax, src_final, 0, 0} with [d0,d1) the dst rows of the chunk,
 *             [p0,p1) = [rowptr[d0], rowptr[d1]) its edges, smax the largest src id referenced by chunks <= c (-1: none),
 *             src_final the first src id that still has an edge in a later chunk (Ns for the last chunk).
 * All five gradients/outputs are written.  Results are bit-identical to ab2_gtconv_fwd_bwd_host. */
int ab2_gtconv_fwd_bwd_host_streamed(const void* q_host, const void* k_host, const void* v_host, const void* e_host,
                                     const void* g_host, int dtype, const int32_t* rowptr, const int32_t* col,
                                     const int32_t* perm, const int32_t* colptr, const int32_t* csr2csc,
                                     const int32_t* crow, int64_t Ns, int64_t Nd, int64_t E, int H, int C, void* out_host,
                                     void* dq_host, void* dk_host, void* dv_host, void* de_host, const int64_t* meta_host,
                                     int nchunks, void* dev_ws, size_t dev_ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Dense node-side contractions on tcgen05 tensor cores (TMEM accumulators, TMA-staged operands).
 * Replaces the nn.Linear calls inside the graph blocks: layers/block.py:491-499, 615-620 (lin_self / lin_query / lin_key /
 * lin_value), :531-533, 630 (projection + skip), :349-354, 537, 633 (node_dst_mlp), layers/conv.py:53-59 + layers/mlp.py:74-84
 * (GraphConv.edge_mlp), and their autograd backward.
 *
 *   D[M,N] = epilogue( A[M,K] . B[N,K]^T )        bf16 operands, fp32 accumulation
 *
 * Operand layouts: a_mn = 0: A is [M,K] row-major (row stride lda);  a_mn = 1: A is stored as [K,M] (row stride lda).
 *                  b_mn = 0: B is [N,K] row-major (nn.Linear weight); b_mn = 1: B is stored as [K,N].
 *   forward  y = x W^T        : a = x  (a_mn 0), b = W  (b_mn 0)
 *   dgrad    dx = dy W        : a = dy (a_mn 0), b = W  (b_mn 1; W is [K_gemm = out_features, N_gemm = in_features])
 *   wgrad    dW = dy^T x      : a = dy (a_mn 1), b = x  (b_mn 1), splits > 1 (contraction over the rows)
 * Epilogue, in this order (every pointer optional):
 *   acc = row_scale[m] * acc + row_shift[m] * col_vec[n]     (LayerNorm folded into the GEMM: rstd, -rstd*mean, colsum(W'))
 *   acc += bias[n]
 *   acc += gather_a[gather_a_idx[m], n] + gather_b[gather_b_idx[m], n]   (bf16 row tables, row stride ld_gather, int64 indices:
 *                                                              GraphConv's split first layer -- pi[dst_t] + pj[src_t], conv.py:69)
 *   dact_pre != NULL : acc *= act'(dact_pre[m,n])            (dgrad through an activation; bf16 [M,N])
 *   else act != 3    : pre_out[m,n] = acc (bf16, optional side output kept for backward);  acc = act(acc)
 *   acc += residual[m, n]                                    (bf16 or fp32, row stride ld_res)
 *   out[n / seg_cols][m, n % seg_cols] = acc                 (bf16 or fp32; up to 4 column segments, row stride ld_out)
 * act: 0 SiLU, 1 GELU(erf), 2 ReLU, 3 identity.  Requirements: N % 8 == 0, lda / ldb multiples of 8, 16-byte aligned pointers,
 * seg_cols % 32 == 0 (0 = one output).  splits > 1 (split-K): plain single fp32/bf16 output, fp32 partials in `workspace`
 * (ab2_gemm_workspace_bytes), summed in a fixed order by a second kernel -- deterministic.
 * ab2_version() >= 101: the descriptor ends with a_seg / a_seg_len (segmented A operand).
 * ------------------------------------------------------------------------------------------------- */
typedef struct ab2_gemm {
  int64_t M, N, K;
  const void* a;
  int64_t lda;
  const void* b;
  int64_t ldb;
  int32_t a_mn, b_mn;
  void* out[4];
  int64_t ld_out;
  int32_t seg_cols, out_f32;
  const float* bias;
  const float* row_scale;
  const float* row_shift;
  const float* col_vec;
  void* pre_out;
  const void* dact_pre;
  const void* residual;
  int64_t ld_res;
  int32_t res_f32, act;
  int32_t splits, reserved;
  const void* gather_a;
  const int64_t* gather_a_idx;
  const void* gather_b;
  const int64_t* gather_b_idx;
  int64_t ld_gather;
  /* A given as up to 4 tensors of equal shape and row stride lda (a, a_seg[0..2]), cut every a_seg_len elements along K
   * (K-major A: the fused q|k|v|self dgrad reads dq, dk, dv, dself where autograd left them) or along M (MN-major A: the
   * fused wgrad).  a_seg_len = 0: A is the single tensor `a`.  a_seg_len must be a multiple of 64 (K-major) / 256 (MN-major)
   * and divide the segmented extent into at most 4 pieces. */
  const void* a_seg[3];
  int64_t a_seg_len;
} ab2_gemm;
size_t ab2_gemm_workspace_bytes(const ab2_gemm* d);
int ab2_gemm_bf16(const ab2_gemm* d, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * LayerNorm over node rows, feeding the GEMMs above, and the column sum used for bias gradients.
 * Replaces nn.LayerNorm at layers/block.py:487-489, 611 (layer_norm1 / layer_norm2) and :349-354 (node_dst_mlp[0]) plus the
 * fp32 -> bf16 cast autocast inserts in front of every nn.Linear: one pass, output in the dtype the GEMM consumes.
 * x [M,D] (x_dtype), gamma / beta fp32 [D], y [M,D] (y_dtype), mean / rstd fp32 [M] (kept for backward; may be NULL).
 * D % 8 == 0, D <= 2048 (forward) / 1024 (backward).
 * Backward: g [M,D] (g_dtype) -> dx [M,D] (x_dtype) = LN-backward(g) (+ add [M,D] (x_dtype), the gradient of a residual branch,
 * when not NULL); dgamma / dbeta fp32 [D], reduced in a fixed order through `partial` ((ab2_ln_parts() + 1) * 2 * D floats).
 * ab2_colsum: out[n] = sum_m a[m, n] (fp32), a [M,N] with row stride ld; partial: ab2_ln_parts() * N floats.  Deterministic.
 * ------------------------------------------------------------------------------------------------- */
int ab2_ln_parts(void);
int ab2_layernorm_fwd(const void* x, int x_dtype, const float* gamma, const float* beta, float eps, int64_t M, int D, void* y,
                      int y_dtype, float* mean, float* rstd, void* stream);
int ab2_layernorm_bwd(const void* g, int g_dtype, const void* x, int x_dtype, const float* gamma, const float* mean,
                      const float* rstd, int64_t M, int D, const void* add, void* dx, float* partial, float* dgamma, float* dbeta,
                      void* stream);
int ab2_colsum(const void* a, int dtype, int64_t M, int N, int64_t ld, float* partial, float* out, void* stream);
/* Raw edge features in front of `lin_edge` (layers/block.py:497, 618; edge_dim = 11 in the reference configs): x [rows, k] (fp32 or
 * bf16, row stride ld) -> out bf16 [rows, kp], kp >= k a multiple of 8, zero-padded -- the K-major A operand of ab2_gemm_bf16
 * (replaces F.pad + the autocast cast).  ab2_unpad_cast_rows is its backward: g bf16 [rows, kp] -> out [rows, k] (row stride ld). */
int ab2_pad_cast_rows(const void* x, int x_dtype, int64_t rows, int k, int64_t ld, void* out, int kp, void* stream);
int ab2_unpad_cast_rows(const void* g, int64_t rows, int kp, void* out, int out_dtype, int k, int64_t ld, void* stream);
/* dpi[i] = sum of g[t] over the edges t into dst i (CSR order), dpj[j] = sum over the edges out of src j (CSC order); g [E,D] in
 * original edge order.  The node-side gradients of GraphConv's split first layer (conv.py:69) when g = d(pre-activation) already
 * came out of a GEMM epilogue.  Either output may be NULL.  Deterministic. */
int ab2_edge_segment_sums(const void* g, const int32_t* rowptr, const int32_t* perm, const int32_t* colptr, const int32_t* cpos,
                          int64_t E, int64_t Ns, int64_t Nd, int D, int dtype, void* dpi, void* dpj, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ANEMOI_B200_H */
