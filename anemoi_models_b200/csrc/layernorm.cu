// LayerNorm forward / backward over node rows, and a deterministic column sum (bias gradients), sm_100a.
//
// Replaces, around the tensor-core GEMMs of the graph blocks (reference paths relative to src/anemoi/models/):
//   layers/block.py:487-489, 611 (layer_norm1 / layer_norm2), :349-354 (node_dst_mlp[0]) -- nn.LayerNorm, which under autocast
//   runs in fp32 and is followed by a separate fp32 -> bf16 cast in front of every nn.Linear (measured in round 1 on the
//   AIFS-like step: 41 ms of LayerNorm kernels + 41 ms of `direct_copy` casts of a 280 ms step).  Here one pass reads the row
//   (fp32 or bf16), keeps it in registers, and writes the normalised row in the dtype the GEMM consumes (bf16) together with
//   mean / rstd for backward.  HBM-bound: b_in + b_out bytes per element.
// Backward: dx = rstd * (g*gamma - mean(g*gamma) - xhat * mean(g*gamma*xhat)); dgamma / dbeta are column sums accumulated in
// registers per CTA over a fixed slice of the rows, written as partials and reduced by a second kernel in a fixed order
// (deterministic, no atomics).
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace ab2 {

constexpr int kLnWarps = 4;
constexpr int kLnParts = 888;  // CTAs of the backward / column-sum kernels = rows of the partial buffers (6 per SM)

template <typename T>
__device__ __forceinline__ void load_row_vec(const T* p, float (&f)[8]);
template <>
__device__ __forceinline__ void load_row_vec<__nv_bfloat16>(const __nv_bfloat16* p, float (&f)[8]) {
  unpack<__nv_bfloat16>(ldg16(p), f);
}
template <>
__device__ __forceinline__ void load_row_vec<float>(const float* p, float (&f)[8]) {
  float a[4], b[4];
  unpack<float>(ldg16(p), a);
  unpack<float>(ldg16(p + 4), b);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[i] = a[i];
    f[4 + i] = b[i];
  }
}
template <typename T>
__device__ __forceinline__ void store_row_vec(T* p, const float (&f)[8]);
template <>
__device__ __forceinline__ void store_row_vec<__nv_bfloat16>(__nv_bfloat16* p, const float (&f)[8]) {
  stg16(p, pack<__nv_bfloat16>(f));
}
template <>
__device__ __forceinline__ void store_row_vec<float>(float* p, const float (&f)[8]) {
  float a[4] = {f[0], f[1], f[2], f[3]}, b[4] = {f[4], f[5], f[6], f[7]};
  stg16(p, pack<float>(a));
  stg16(p + 4, pack<float>(b));
}

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

// one warp per row; lane l owns the 8-element groups l, l + 32, ... (VPL of them): D <= VPL * 256
template <typename TX, typename TY, int VPL>
__global__ void __launch_bounds__(kLnWarps * 32) layernorm_fwd_kernel(const TX* __restrict__ x, const float* __restrict__ gamma,
                                                                     const float* __restrict__ beta, float eps, long long M, int D,
                                                                     TY* __restrict__ y, float* __restrict__ mean_out,
                                                                     float* __restrict__ rstd_out) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * kLnWarps + (threadIdx.x >> 5), nwarps = (long long)gridDim.x * kLnWarps;
  const int ngroups = D >> 3;
  float gm[VPL][8], bt[VPL][8];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int gidx = lane + 32 * i;
    if (gidx < ngroups) {
      load_row_vec<float>(gamma + gidx * 8, gm[i]);
      load_row_vec<float>(beta + gidx * 8, bt[i]);
    }
  }
  for (long long row = warp0; row < M; row += nwarps) {
    float v[VPL][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int gidx = lane + 32 * i;
      if (gidx < ngroups) {
        load_row_vec<TX>(x + row * D + gidx * 8, v[i]);
#pragma unroll
        for (int u = 0; u < 8; ++u) s += v[i][u];
      }
    }
    const float mean = warp_sum(s) / (float)D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      if (lane + 32 * i < ngroups) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float d = v[i][u] - mean;
          q += d * d;
        }
      }
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int gidx = lane + 32 * i;
      if (gidx < ngroups) {
        float o[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) o[u] = (v[i][u] - mean) * rstd * gm[i][u] + bt[i][u];
        store_row_vec<TY>(y + row * D + gidx * 8, o);
      }
    }
    if (lane == 0) {
      if (mean_out != nullptr) mean_out[row] = mean;
      if (rstd_out != nullptr) rstd_out[row] = rstd;
    }
  }
}

// backward: TG = dtype of the incoming gradient, TX = dtype of x and of dx.  partial: [gridDim.x][2][D] fp32.
// (A variant with gamma and the dgamma / dbeta accumulators in shared memory -- 96 instead of 207 registers, 20 instead of 8 rows
// in flight per SM -- was measured SLOWER on the model step: 12.9 vs 9.7 ms over its 38 calls, profiles/r02/bench_model_r02k.json.)
template <typename TG, typename TX, int VPL>
__global__ void __launch_bounds__(kLnWarps * 32) layernorm_bwd_kernel(const TG* __restrict__ g, const TX* __restrict__ x,
                                                                     const float* __restrict__ gamma, const float* __restrict__ mean,
                                                                     const float* __restrict__ rstd, long long M, int D,
                                                                     const TX* __restrict__ add, TX* __restrict__ dx,
                                                                     float* __restrict__ partial) {
  extern __shared__ float red[];  // [kLnWarps][2][D]
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int ngroups = D >> 3;
  // fixed slice of rows per CTA, rows of a slice dealt round-robin to its warps: the summation order is a function of (M, grid)
  const long long per = (M + gridDim.x - 1) / gridDim.x;
  const long long r0 = (long long)blockIdx.x * per, r1 = min(r0 + per, M);
  float gm[VPL][8], dg[VPL][8], db[VPL][8];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int gidx = lane + 32 * i;
#pragma unroll
    for (int u = 0; u < 8; ++u) dg[i][u] = db[i][u] = 0.f;
    if (gidx < ngroups) load_row_vec<float>(gamma + gidx * 8, gm[i]);
  }
  for (long long row = r0 + w; row < r1; row += kLnWarps) {
    const float mu = mean[row], rs = rstd[row];
    float gv[VPL][8], xh[VPL][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int gidx = lane + 32 * i;
      if (gidx < ngroups) {
        load_row_vec<TG>(g + row * D + gidx * 8, gv[i]);
        load_row_vec<TX>(x + row * D + gidx * 8, xh[i]);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          xh[i][u] = (xh[i][u] - mu) * rs;
          dg[i][u] += gv[i][u] * xh[i][u];
          db[i][u] += gv[i][u];
          gv[i][u] *= gm[i][u];
          s1 += gv[i][u];
          s2 += gv[i][u] * xh[i][u];
        }
      }
    }
    s1 = warp_sum(s1) / (float)D;
    s2 = warp_sum(s2) / (float)D;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int gidx = lane + 32 * i;
      if (gidx < ngroups) {
        float o[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) o[u] = rs * (gv[i][u] - s1 - xh[i][u] * s2);
        if (add != nullptr) {  // + gradient arriving through the residual branch (same dtype as x)
          float a[8];
          load_row_vec<TX>(add + row * D + gidx * 8, a);
#pragma unroll
          for (int u = 0; u < 8; ++u) o[u] += a[u];
        }
        store_row_vec<TX>(dx + row * D + gidx * 8, o);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int gidx = lane + 32 * i;
    if (gidx < ngroups) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        red[(w * 2 + 0) * D + gidx * 8 + u] = dg[i][u];
        red[(w * 2 + 1) * D + gidx * 8 + u] = db[i][u];
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * D; c += blockDim.x) {
    float acc = 0.f;
#pragma unroll
    for (int ww = 0; ww < kLnWarps; ++ww) acc += red[ww * 2 * D + c];
    partial[(size_t)blockIdx.x * 2 * D + c] = acc;
  }
}

// backward, second layout (the default): thread t of the CTA owns COLUMNS [8 t, 8 t + 8) for every row of the CTA's slice, so the
// dgamma / dbeta accumulators are 16 registers (not 64 per warp-row lane), ~7 CTAs of D / 8 threads are resident per SM, and a
// thread has the loads of R rows (g, x and the residual-branch gradient) in flight before it needs any of them.  The row sums
// s1, s2 cross the CTA's warps through a double-buffered shared-memory slot: one __syncthreads per R rows.
// Measured on the AIFS-like step (38 calls, 23 GB): warp-per-row kernel above 14.9 ms (1.5 TB/s); see profiles/r02/SUMMARY.md.
template <typename TG, typename TX, int R>
__global__ void __launch_bounds__(256) layernorm_bwd_cols_kernel(const TG* __restrict__ g, const TX* __restrict__ x,
                                                                 const float* __restrict__ gamma, const float* __restrict__ mean,
                                                                 const float* __restrict__ rstd, long long M, int D,
                                                                 const TX* __restrict__ add, TX* __restrict__ dx,
                                                                 float* __restrict__ partial) {
  __shared__ __align__(16) float red[2][R][2][8];  // [buffer][row][s1 | s2][warp]
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int col = tid * 8;
  const bool active = col < D;  // blockDim = D / 8 rounded up to whole warps
  const float inv_d = 1.f / (float)D;
  const long long per = (M + gridDim.x - 1) / gridDim.x;
  const long long r0 = (long long)blockIdx.x * per, r1 = min(r0 + per, M);
  float gm[8], dg[8], db[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) gm[u] = dg[u] = db[u] = 0.f;
  if (active) load_row_vec<float>(gamma + col, gm);
  for (int i = tid; i < 2 * R * 2 * 8; i += blockDim.x) (&red[0][0][0][0])[i] = 0.f;  // slots of warps that do not exist stay zero
  __syncthreads();
  int buf = 0;
  for (long long row = r0; row < r1; row += R, buf ^= 1) {
    float gv[R][8], xh[R][8], av[R][8], rs[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {  // all loads of the R rows first
      const long long rr = row + r;
      const bool ok = active && rr < r1;
#pragma unroll
      for (int u = 0; u < 8; ++u) gv[r][u] = xh[r][u] = av[r][u] = 0.f;
      rs[r] = 0.f;
      if (ok) {
        load_row_vec<TG>(g + rr * D + col, gv[r]);
        load_row_vec<TX>(x + rr * D + col, xh[r]);
        if (add != nullptr) load_row_vec<TX>(add + rr * D + col, av[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const long long rr = row + r;
      float s1 = 0.f, s2 = 0.f;
      if (active && rr < r1) {
        const float mu = mean[rr];
        rs[r] = rstd[rr];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          xh[r][u] = (xh[r][u] - mu) * rs[r];
          dg[u] += gv[r][u] * xh[r][u];
          db[u] += gv[r][u];
          gv[r][u] *= gm[u];
          s1 += gv[r][u];
          s2 += gv[r][u] * xh[r][u];
        }
      }
      s1 = warp_sum(s1);
      s2 = warp_sum(s2);
      if (lane == 0) {
        red[buf][r][0][w] = s1;
        red[buf][r][1][w] = s2;
      }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const long long rr = row + r;
      const float4 a0 = *reinterpret_cast<const float4*>(&red[buf][r][0][0]), a1 = *reinterpret_cast<const float4*>(&red[buf][r][0][4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&red[buf][r][1][0]), b1 = *reinterpret_cast<const float4*>(&red[buf][r][1][4]);
      const float s1 = (((a0.x + a0.y) + (a0.z + a0.w)) + ((a1.x + a1.y) + (a1.z + a1.w))) * inv_d;
      const float s2 = (((b0.x + b0.y) + (b0.z + b0.w)) + ((b1.x + b1.y) + (b1.z + b1.w))) * inv_d;
      if (active && rr < r1) {
        float o[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) o[u] = rs[r] * (gv[r][u] - s1 - xh[r][u] * s2) + av[r][u];
        store_row_vec<TX>(dx + rr * D + col, o);
      }
    }
  }
  if (active) {
    float* p0 = partial + (size_t)blockIdx.x * 2 * D + col;
    float lo[4] = {dg[0], dg[1], dg[2], dg[3]}, hi[4] = {dg[4], dg[5], dg[6], dg[7]};
    *reinterpret_cast<float4*>(p0) = make_float4(lo[0], lo[1], lo[2], lo[3]);
    *reinterpret_cast<float4*>(p0 + 4) = make_float4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<float4*>(p0 + D) = make_float4(db[0], db[1], db[2], db[3]);
    *reinterpret_cast<float4*>(p0 + D + 4) = make_float4(db[4], db[5], db[6], db[7]);
  }
}

// out[c] = sum_p partial[p][c], c < n, in a fixed order: block = 32 columns x 32 part-lanes; lane y sums the parts p = y, y + 32, ...
// (coalesced 128-byte reads), then the 32 lane sums of a column are added in lane order
__global__ void __launch_bounds__(1024) partial_reduce_kernel(const float* __restrict__ partial, int parts, int n, float* __restrict__ out) {
  __shared__ float sm[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f;
  if (c < n)
    for (int p = threadIdx.y; p < parts; p += 32) acc += partial[(size_t)p * n + c];
  sm[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < n) {
    float t = 0.f;
#pragma unroll
    for (int y = 0; y < 32; ++y) t += sm[y][threadIdx.x];
    out[c] = t;
  }
}

// column sums of a [M, N] matrix (row stride ld): partial[blockIdx.x][N]; thread t owns columns 8t .. 8t+7 of a 8*blockDim-wide
// column block, rows of the CTA's fixed slice in order
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ a, long long M, int N, long long ld, float* __restrict__ partial) {
  const long long per = (M + gridDim.x - 1) / gridDim.x;
  const long long r0 = (long long)blockIdx.x * per, r1 = min(r0 + per, M);
  for (int c0 = threadIdx.x * 8; c0 < N; c0 += blockDim.x * 8) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    long long row = r0;
    for (; row + 4 <= r1; row += 4) {  // 4 independent 16 B loads in flight per thread
      float f[4][8];
#pragma unroll
      for (int i = 0; i < 4; ++i) load_row_vec<T>(a + (row + i) * ld + c0, f[i]);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] += f[i][u];
    }
    for (; row < r1; ++row) {
      float f[8];
      load_row_vec<T>(a + row * ld + c0, f);
#pragma unroll
      for (int u = 0; u < 8; ++u) acc[u] += f[u];
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) partial[(size_t)blockIdx.x * N + c0 + u] = acc[u];
  }
}

template <typename TX, typename TY>
static int ln_fwd_dispatch(const void* x, const float* gamma, const float* beta, float eps, long long M, int D, void* y, float* mean,
                           float* rstd, cudaStream_t st) {
  const int vpl = (D / 8 + 31) / 32;
  long long blocks = (M + kLnWarps - 1) / kLnWarps;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
#define AB2_LN_FWD(V)                                                                                                       \
  layernorm_fwd_kernel<TX, TY, V><<<(unsigned)blocks, kLnWarps * 32, 0, st>>>((const TX*)x, gamma, beta, eps, M, D, (TY*)y, mean, rstd)
  if (vpl <= 1) AB2_LN_FWD(1);
  else if (vpl <= 2) AB2_LN_FWD(2);
  else if (vpl <= 4) AB2_LN_FWD(4);
  else if (vpl <= 8) AB2_LN_FWD(8);
  else return fail(AB2_ERR_UNSUPPORTED, "layernorm: D = %d is wider than 2048", D);
#undef AB2_LN_FWD
  AB2_LAUNCH_OK("layernorm_fwd_kernel");
  return AB2_OK;
}

template <typename TG, typename TX>
static int ln_bwd_dispatch(const void* g, const void* x, const float* gamma, const float* mean, const float* rstd, long long M, int D,
                           const void* add, void* dx, float* partial, int* parts_used, cudaStream_t st) {
  *parts_used = kLnParts;
  // AB2_LN_BWD=rows keeps the warp-per-row kernel (A/B runs); the column-owner kernel takes D <= 2048
  static const bool by_rows = [] {
    const char* e = getenv("AB2_LN_BWD");
    return e != nullptr && strcmp(e, "rows") == 0;
  }();
  if (!by_rows && D <= 2048) {
    const int threads = ((D / 8 + 31) / 32) * 32;
    // one wave: as many CTAs as are resident at once (888 CTAs at 5 per SM ran as 740 + a tail of 148), at most kLnParts
    int per_sm = 0;
    AB2_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, layernorm_bwd_cols_kernel<TG, TX, 2>, threads, 0));
    *parts_used = std::max(1, std::min(kLnParts, num_sms() * std::max(per_sm, 1)));
    layernorm_bwd_cols_kernel<TG, TX, 2><<<*parts_used, threads, 0, st>>>((const TG*)g, (const TX*)x, gamma, mean, rstd, M, D,
                                                                           (const TX*)add, (TX*)dx, partial);
    AB2_LAUNCH_OK("layernorm_bwd_cols_kernel");
    return AB2_OK;
  }
  const int vpl = (D / 8 + 31) / 32;
  const size_t smem = (size_t)kLnWarps * 2 * D * sizeof(float);
#define AB2_LN_BWD(V)                                                                                                       \
  layernorm_bwd_kernel<TG, TX, V><<<kLnParts, kLnWarps * 32, smem, st>>>((const TG*)g, (const TX*)x, gamma, mean, rstd, M, D, \
                                                                         (const TX*)add, (TX*)dx, partial)
  if (vpl <= 1) AB2_LN_BWD(1);
  else if (vpl <= 2) AB2_LN_BWD(2);
  else if (vpl <= 4) AB2_LN_BWD(4);
  else return fail(AB2_ERR_UNSUPPORTED, "layernorm backward: D = %d is wider than 1024", D);
#undef AB2_LN_BWD
  AB2_LAUNCH_OK("layernorm_bwd_kernel");
  return AB2_OK;
}

// Rows of k raw features (the reference's edge_dim = 11: 3 geometric + 8 trainable columns, fp32) -> bf16 rows of kp >= k columns,
// zero-padded: what the `lin_edge` GEMM reads (16-byte aligned rows).  F.pad + .to(bfloat16) ran as two strided elementwise
// kernels at ~90 GB/s (5.4 ms of the AIFS-like step for 36 calls).  One thread per (row, 8-column group).
template <typename TI>
__global__ void pad_cast_rows_kernel(const TI* __restrict__ x, long long rows, int k, long long ld, __nv_bfloat16* __restrict__ out, int kp) {
  const int groups = kp >> 3;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long row = i / groups;
  const int c0 = (int)(i - row * groups) * 8;
  if (row >= rows) return;
  float f[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) f[u] = (c0 + u < k) ? to_f(x[row * ld + c0 + u]) : 0.f;
  stg16(out + row * kp + c0, pack<__nv_bfloat16>(f));
}
// backward of the above: the first k columns of bf16 [rows, kp] -> [rows, k] in the dtype of the raw features
template <typename TO>
__global__ void unpad_cast_rows_kernel(const __nv_bfloat16* __restrict__ g, long long rows, int kp, TO* __restrict__ out, int k, long long ld) {
  const int groups = kp >> 3;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long row = i / groups;
  const int c0 = (int)(i - row * groups) * 8;
  if (row >= rows || c0 >= k) return;
  float f[8];
  unpack<__nv_bfloat16>(ldg16(g + row * kp + c0), f);
#pragma unroll
  for (int u = 0; u < 8; ++u)
    if (c0 + u < k) out[row * ld + c0 + u] = from_f<TO>(f[u]);
}

}  // namespace ab2

using namespace ab2;

extern "C" int ab2_pad_cast_rows(const void* x, int x_dtype, int64_t rows, int k, int64_t ld, void* out, int kp, void* stream) {
  if (rows < 0 || k <= 0 || kp < k || kp % 8 != 0 || ld < k) return fail(AB2_ERR_INVALID, "pad_cast_rows: need 0 < k <= kp, kp %% 8 == 0, ld >= k");
  if (rows == 0) return AB2_OK;
  if (x == nullptr || out == nullptr) return fail(AB2_ERR_INVALID, "pad_cast_rows: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const long long n = (long long)rows * (kp / 8);
  const unsigned blocks = (unsigned)((n + 255) / 256);
  if (x_dtype == AB2_F32) pad_cast_rows_kernel<float><<<blocks, 256, 0, st>>>((const float*)x, rows, k, ld, (__nv_bfloat16*)out, kp);
  else if (x_dtype == AB2_BF16) pad_cast_rows_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)x, rows, k, ld, (__nv_bfloat16*)out, kp);
  else return fail(AB2_ERR_INVALID, "pad_cast_rows: bad dtype");
  AB2_LAUNCH_OK("pad_cast_rows_kernel");
  return AB2_OK;
}

extern "C" int ab2_unpad_cast_rows(const void* g, int64_t rows, int kp, void* out, int out_dtype, int k, int64_t ld, void* stream) {
  if (rows < 0 || k <= 0 || kp < k || kp % 8 != 0 || ld < k) return fail(AB2_ERR_INVALID, "unpad_cast_rows: need 0 < k <= kp, kp %% 8 == 0, ld >= k");
  if (rows == 0) return AB2_OK;
  if (g == nullptr || out == nullptr) return fail(AB2_ERR_INVALID, "unpad_cast_rows: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const long long n = (long long)rows * (kp / 8);
  const unsigned blocks = (unsigned)((n + 255) / 256);
  if (out_dtype == AB2_F32) unpad_cast_rows_kernel<float><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)g, rows, kp, (float*)out, k, ld);
  else if (out_dtype == AB2_BF16) unpad_cast_rows_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)g, rows, kp, (__nv_bfloat16*)out, k, ld);
  else return fail(AB2_ERR_INVALID, "unpad_cast_rows: bad dtype");
  AB2_LAUNCH_OK("unpad_cast_rows_kernel");
  return AB2_OK;
}

extern "C" int ab2_ln_parts(void) { return kLnParts; }

extern "C" int ab2_layernorm_fwd(const void* x, int x_dtype, const float* gamma, const float* beta, float eps, int64_t M, int D, void* y,
                                 int y_dtype, float* mean, float* rstd, void* stream) {
  if (M < 0 || D <= 0 || D % 8 != 0) return fail(AB2_ERR_UNSUPPORTED, "layernorm: D must be a positive multiple of 8 (got %d)", D);
  if (M == 0) return AB2_OK;
  if (x == nullptr || y == nullptr || gamma == nullptr || beta == nullptr) return fail(AB2_ERR_INVALID, "layernorm: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (x_dtype == AB2_BF16 && y_dtype == AB2_BF16) return ln_fwd_dispatch<__nv_bfloat16, __nv_bfloat16>(x, gamma, beta, eps, M, D, y, mean, rstd, st);
  if (x_dtype == AB2_F32 && y_dtype == AB2_BF16) return ln_fwd_dispatch<float, __nv_bfloat16>(x, gamma, beta, eps, M, D, y, mean, rstd, st);
  if (x_dtype == AB2_F32 && y_dtype == AB2_F32) return ln_fwd_dispatch<float, float>(x, gamma, beta, eps, M, D, y, mean, rstd, st);
  if (x_dtype == AB2_BF16 && y_dtype == AB2_F32) return ln_fwd_dispatch<__nv_bfloat16, float>(x, gamma, beta, eps, M, D, y, mean, rstd, st);
  return fail(AB2_ERR_INVALID, "layernorm: bad dtype");
}

extern "C" int ab2_layernorm_bwd(const void* g, int g_dtype, const void* x, int x_dtype, const float* gamma, const float* mean,
                                 const float* rstd, int64_t M, int D, const void* add, void* dx, float* partial, float* dgamma,
                                 float* dbeta, void* stream) {
  if (M < 0 || D <= 0 || D % 8 != 0) return fail(AB2_ERR_UNSUPPORTED, "layernorm: D must be a positive multiple of 8 (got %d)", D);
  if (g == nullptr || x == nullptr || gamma == nullptr || mean == nullptr || rstd == nullptr || dx == nullptr || partial == nullptr ||
      dgamma == nullptr || dbeta == nullptr)
    return fail(AB2_ERR_INVALID, "layernorm backward: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  int rc, parts = kLnParts;
  if (g_dtype == AB2_BF16 && x_dtype == AB2_BF16) rc = ln_bwd_dispatch<__nv_bfloat16, __nv_bfloat16>(g, x, gamma, mean, rstd, M, D, add, dx, partial, &parts, st);
  else if (g_dtype == AB2_BF16 && x_dtype == AB2_F32) rc = ln_bwd_dispatch<__nv_bfloat16, float>(g, x, gamma, mean, rstd, M, D, add, dx, partial, &parts, st);
  else if (g_dtype == AB2_F32 && x_dtype == AB2_F32) rc = ln_bwd_dispatch<float, float>(g, x, gamma, mean, rstd, M, D, add, dx, partial, &parts, st);
  else if (g_dtype == AB2_F32 && x_dtype == AB2_BF16) rc = ln_bwd_dispatch<float, __nv_bfloat16>(g, x, gamma, mean, rstd, M, D, add, dx, partial, &parts, st);
  else return fail(AB2_ERR_INVALID, "layernorm backward: bad dtype");
  if (rc) return rc;
  // partial is [kLnParts][2][D]: dgamma = columns [0, D), dbeta = [D, 2D) of the reduced row
  partial_reduce_kernel<<<(2 * D + 31) / 32, dim3(32, 32), 0, st>>>(partial, parts, 2 * D, partial + (size_t)kLnParts * 2 * D);
  AB2_LAUNCH_OK("partial_reduce_kernel");
  AB2_CUDA_OK(cudaMemcpyAsync(dgamma, partial + (size_t)kLnParts * 2 * D, D * sizeof(float), cudaMemcpyDeviceToDevice, st));
  AB2_CUDA_OK(cudaMemcpyAsync(dbeta, partial + (size_t)kLnParts * 2 * D + D, D * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return AB2_OK;
}

extern "C" int ab2_colsum(const void* a, int dtype, int64_t M, int N, int64_t ld, float* partial, float* out, void* stream) {
  if (M < 0 || N <= 0 || N % 8 != 0) return fail(AB2_ERR_UNSUPPORTED, "colsum: N must be a positive multiple of 8 (got %d)", N);
  if (a == nullptr || partial == nullptr || out == nullptr) return fail(AB2_ERR_INVALID, "colsum: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == AB2_BF16) colsum_kernel<__nv_bfloat16><<<kLnParts, 256, 0, st>>>((const __nv_bfloat16*)a, M, N, ld, partial);
  else if (dtype == AB2_F32) colsum_kernel<float><<<kLnParts, 256, 0, st>>>((const float*)a, M, N, ld, partial);
  else return fail(AB2_ERR_INVALID, "colsum: bad dtype");
  AB2_LAUNCH_OK("colsum_kernel");
  partial_reduce_kernel<<<(N + 31) / 32, dim3(32, 32), 0, st>>>(partial, kLnParts, N, out);
  AB2_LAUNCH_OK("partial_reduce_kernel");
  return AB2_OK;
}
