// Shared between gtconv.cu (LDG kernels + dispatch) and gtconv_tma.cu (bulk-copy pipelined kernels).
#pragma once
#include "common.cuh"

namespace ab2 {

constexpr int kU = 4;  // edges in flight per thread / per pipeline stage

struct RowMap {  // how a CTA's threads map onto (row, 16-byte chunk)
  int tpd;       // threads per row handled by one CTA (= heads-per-slice * LPH)
  int rpb;       // rows per CTA
  int chunks;    // 16-byte chunks in a full row (= D*sizeof(T)/16)
};

// everything a conv launch needs (host side)
struct ConvArgs {
  const void *q, *k, *v, *e;        // k, v: the rank's own src rows [0, n_own)
  const void *k_halo, *v_halo;      // src rows [n_own, Ns) (NULL on one GPU)
  const int *rowptr, *col, *perm, *colptr, *csr2csc, *crow;
  int Ns, Nd, n_own, H, C;
  int src_lo, src_hi;               // src rows the backward src pass covers ([0, Ns) unless a caller streams row ranges)
  int64_t E;
  float qscale, scale;
  const void *out, *g;
  const float* lse2_in;
  void *out_w, *dq, *dk, *dv, *dk_halo, *dv_halo, *de;
  float* lse2_w;
  float2* ads;
  bool low_degree;
  cudaStream_t st;
};

// gtconv_tma.cu: returns true when the pipelined kernel was launched (row = 2048 bytes, vector layout), false -> use the LDG kernel
bool try_launch_fwd_tma(int dtype, int lph, const ConvArgs& a);
bool try_launch_bwd_dst_tma(int dtype, int lph, const ConvArgs& a);
bool try_launch_bwd_src_tma(int dtype, int lph, const ConvArgs& a);
bool tma_applicable(int which, int dtype, int H, int C);  // which: 0 = forward, 1 = backward dst pass, 2 = backward src pass

}  // namespace ab2
