// Dense node-side contractions on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM), sm_100a.
//
// Replaces, inside the graph blocks of the reference (paths relative to src/anemoi/models/):
//   layers/block.py:491-499, 615-620   lin_self / lin_query / lin_key / lin_value  (concatenated: [q|self], [k|v], one GEMM each)
//   layers/block.py:531-533, 630       projection(out + x_r) + x_skip               (bias + residual epilogue)
//   layers/block.py:349-354, 537, 633  node_dst_mlp: Linear -> act -> Linear + res   (bias + activation / bias + residual epilogues)
//   layers/conv.py:53-59 + mlp.py:74-84 GraphConv.edge_mlp Linear layers             (bias + activation epilogues)
// and their backward (dgrad with the activation derivative in the epilogue, split-K wgrad).
//
//   D[M,N] = epilogue( A[M,K] . B[N,K]^T ),  bf16 operands, fp32 accumulation in tensor memory.
//
// Structure (one persistent CTA per SM, 10 warps; CG = 2: the two CTAs of a cluster -- the two SMs of a TPC -- work as ONE
// 256 x BN tile with tcgen05.mma.cta_group::2: each CTA stages its 128 rows of A and its HALF of B, so the B bytes pulled
// from L2 and the shared-memory reads per flop are halved; measured: 1-CTA 128 x 256 tiles are L2-feed-bound at 75-85 % of
// cuBLAS, profiles/r02/gemm_probe_r02a.jsonl):
//   warp 0      TMA producer: cp.async.bulk.tensor (SWIZZLE_128B boxes) of the A / B k-blocks into a shared-memory ring,
//               completion counted in bytes on the stage's `full` mbarrier (CG = 2: both CTAs' copies signal the LEADER's);
//   warp 1      MMA issuer (leader CTA): the warp walks its loops converged, ONE elected lane issues tcgen05.mma.kind::f16
//               (128*CG x BN x 16 per instruction) on the staged tiles; tcgen05.commit releases the stage (`empty`, multicast to both CTAs) and, after the last k-block,
//               publishes the accumulator (`tmem_full`).  The warp also owns the TMEM allocation (512 columns = two
//               accumulator stages of BN <= 256);
//   warps 2..9  epilogue: tcgen05.ld (32 lanes x 32 columns per instruction) of the finished accumulator while the MMA warp
//               already works on the next tile in the other TMEM stage; bias / LayerNorm-fold / activation (+ pre-activation
//               side output) / activation-derivative / residual; a thread owns a ROW of the accumulator, so bf16 results go
//               through per-warp shared-memory patch buffers (TMA SWIZZLE_64B layout) and leave as TMA stores of 32 x 32
//               boxes (storing straight from registers writes 32 rows x 16 B per instruction, half a sector each -- measured
//               2-4x slower on epilogue-bound shapes); outputs optionally split by column segments (q | self, k | v land in
//               separate tensors).
// Operands may be K-major (contraction contiguous in memory: x [M,K], W [N,K]) or MN-major (contraction strided: W as the
// B operand of dgrad, dY^T and x as the operands of wgrad), selected per operand in the instruction descriptor; no transposes
// are materialised.  wgrad splits the (long) contraction over CTAs into fp32 partials that a second kernel sums in a fixed
// order (deterministic, no atomics).
#include <cuda.h>

#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <climits>
#include <mutex>

#include "common.cuh"

namespace ab2 {
namespace tc {

constexpr int BM = 128;     // accumulator rows per CTA tile = TMEM lanes
constexpr int BK = 64;      // 64 bf16 = 128 B = one SWIZZLE_128B atom along the contraction
constexpr int UMMA_K = 16;  // contraction per tcgen05.mma for 16-bit operands
constexpr int kBoxBytes = 64 * 64 * 2;  // one 64 x 64 bf16 box (MN-major operands are loaded box by box)

// Epilogue patches: per epilogue warp TWO buffers of 32 rows x 64 bytes (32 bf16 columns), dense, 16-byte chunks XOR-swizzled
// with the row (the TMA SWIZZLE_64B pattern: chunk ^= (row >> 1) & 3) -- conflict-free both for "lane = row" accesses and for
// "8 rows x 4 chunks" accesses, and the layout a TMA store reads.  Two buffers so that the store of one chunk (or of the
// pre-activation) is still being read by the TMA engine while the warp fills the other.
constexpr int kPatchBuf = 32 * 64;
constexpr int kPatchBytes = 2 * kPatchBuf;

// EW = number of epilogue warps (8 or 16: two or four per scheduler -- the epilogue is a latency chain of TMEM load, gathers and
// the shared-memory transpose, so the epilogue-heavy shapes want four; the price is one pipeline stage of shared memory)
template <int BN, int CG, int EW>
struct Cfg {
  static constexpr int kThreads = 64 + 32 * EW;
  static constexpr int kBRows = BN / CG;  // rows of B this CTA stages (CG = 2: half of the tile's columns)
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = kBRows * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagesRaw = (227 * 1024 - 2048 - EW * kPatchBytes) / kStageBytes;  // EW=16: CG=2: 5 (BN=256), 7 (BN=128); EW=8: 6 / 8
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kBarBytes = 1024;
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarBytes + EW * kPatchBytes + 1024;  // + 1024 B alignment slack
};

struct Epi {
  void* out[4];          // column segment s of the result goes to out[s] (row stride ld_out elements)
  long long ld_out;
  int seg_cols;          // width of a column segment (multiple of 32); >= N for a single output
  int out_f32;           // 1: fp32 stores, 0: bf16
  const float* bias;     // [N] or null
  const float* row_scale;  // [M] or null:  acc = row_scale[m] * acc + row_shift[m] * col_vec[n]   (LayerNorm folded into the GEMM)
  const float* row_shift;  // [M]
  const float* col_vec;    // [N]
  int act;               // 0 SiLU, 1 GELU(erf), 2 ReLU, 3 identity  (ops.ACT_CODES)
  void* pre_out;         // bf16 [M,N] (ld = N): pre-activation side output, kept for backward; null to skip
  const void* dact_pre;  // bf16 [M,N] (ld = N): result *= act'(dact_pre)  (dgrad through an activation); null to skip
  const void* residual;  // [M,N] row stride ld_res, bf16 or fp32; added last
  long long ld_res;
  int res_f32;
  const __nv_bfloat16* gather[2];  // row tables added before the activation: acc += gather[i][gather_idx[i][m], n]
  const long long* gather_idx[2];
  long long ld_gather;
};

struct Args {
  int M, N, K;
  int tiles_m, tiles_n, splits, kb_per_split, kb_total;
  long long split_stride;  // elements between the fp32 partials of two splits (split-K only)
  int k_lbo, k_sbo, mn_lbo, mn_sbo;  // descriptor byte offsets (fixed by the tile layout; overridable for bring-up probes only)
  int a_seg_len;  // A given as up to 4 tensors, cut every a_seg_len elements along K (K-major A) or along M (MN-major A)
  Epi epi;
};

// A may be SEGMENTED: [dq | dk | dv | dself] of the fused q/k/v/self projection arrive from autograd as separate tensors; the
// producer picks the tensor map by k-block (dgrad, K-major A) or by tile row (wgrad, MN-major A) instead of a concatenation pass
struct AMaps {
  CUtensorMap m[4];
};
// bf16 results leave through TMA stores (box = 32 columns x 32 rows, SWIZZLE_64B): one map per output segment + the pre-activation
struct OMaps {
  CUtensorMap out[4];
  CUtensorMap pre;
};

// ---- PTX wrappers ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.b32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol error must end in a trapped launch (cudaErrorLaunchFailure), never in a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();  // ~2 s
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(map), "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}
// CG = 2: the copy lands in THIS CTA's shared memory, its bytes are counted on the LEADER CTA's mbarrier (cluster address)
__device__ __forceinline__ void tma_load_2d_cg2(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1)
               : "memory");
}
// one lane of the (converged) warp; the warp keeps walking the loops together, so loop counters, addresses and descriptors stay
// warp-uniform (uniform registers) -- with the loops inside `if (lane == 0)` ptxas treated every tcgen05 / TMA operand as
// divergent and wrapped each instruction in an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall: ~120 issue slots per k-block on
// the MMA warp, as long as the four MMAs of the k-block take (ncu source view, profiles/r02/SUMMARY.md)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.b32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {  // shared::cluster address of `addr` in CTA `rank`
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_commit_cg2(uint32_t bar) {  // arrives on the barrier at this offset in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void tc_mma_cg2(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void sts16(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds16(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  return r;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor (SWIZZLE_128B, descriptor version 1 = sm_100): start address, leading / stride byte
// offsets in 16 B units.  K-major tile [rows][64]: rows are 128 B apart inside an 8-row swizzle atom, atoms 1024 B apart
// (SBO); LBO is not used.  MN-major tile = 64 x 64 boxes [k][64 mn]: 8 k-rows form an atom, atoms 1024 B apart along k
// (SBO), the next 64 mn-elements live in the next box, 8192 B further (LBO).
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  const uint32_t lo = ((addr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
  const uint32_t hi = ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
  return ((uint64_t)hi << 32) | lo;
}

__device__ __forceinline__ float rcp_approx(float x) {  // one MUFU.RCP, no refinement / denormal slow path (|rel err| ~ 1e-7)
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float tanh_approx(float x) {  // one MUFU.TANH (|rel err| <= 2^-11, far below bf16 resolution)
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// ---- epilogue math on PAIRS of fp32 values: sm_100 has packed FFMA2 / FMUL2 / FADD2 (two fp32 lanes per instruction), and the
// epilogue of the activation shapes is issue-bound (ncu, profiles/r02/ncu_gemm_perf_mlp1_r02n.md: 900 warp instructions per
// 32 x 32 chunk at 55 % issue-slot use while the tensor pipe idles 46 % of the time) -- so every polynomial step below is one
// instruction for two elements; only the special-function ops (MUFU) and min / max / sign stay per element.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float a, float b) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 bc2(float c) { return pk2(c, c); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ void add_pair(float& a, float& b, float c0, float c1) { upk2(add2(pk2(a, b), pk2(c0, c1)), a, b); }

// u(x) = -0.5 erfc(|x| / sqrt 2) for two values, and e = exp(-x^2 / 2): Abramowitz & Stegun 7.1.26 (|error| of erfc < 1.5e-7) with
// the -0.5 folded into the coefficients; one MUFU.RCP and one MUFU.EX2 per element, everything else packed.  libdevice's erff (and
// __frcp_rn's refinement + slow path) cost 27 instructions and three branches per element in the first version of this kernel.
__device__ __forceinline__ void half_erfc_pair(f32x2 x, f32x2 ax, f32x2& u, f32x2& e) {
  float d0, d1, s0, s1;
  upk2(fma2(ax, bc2(0.3275911f * 0.70710678118654752f), bc2(1.f)), d0, d1);
  const f32x2 t = pk2(rcp_approx(d0), rcp_approx(d1));
  f32x2 p = fma2(bc2(-0.5f * 1.061405429f), t, bc2(0.5f * 1.453152027f));
  p = fma2(p, t, bc2(-0.5f * 1.421413741f));
  p = fma2(p, t, bc2(0.5f * 0.284496736f));
  p = fma2(p, t, bc2(-0.5f * 0.254829592f));
  const f32x2 y = mul2(x, bc2(0.84932180028801907f));  // sqrt(0.5 log2 e): exp(-x^2 / 2) = 2^(-y^2)
  upk2(mul2(y, y), s0, s1);
  e = pk2(fast_exp2(-s0), fast_exp2(-s1));
  u = mul2(mul2(p, t), e);
}
// act(x) for two values.  GELU (erf form, nn.GELU()): x Phi(x) = max(x, 0) - |x| * 0.5 erfc(|x| / sqrt 2)  -- no sign select, no
// 1 + erf cancellation.  SiLU: x * sigmoid(x) with the sigmoid through ONE MUFU.TANH (exp2 + reciprocal would be two, and at
// K = 512 the SFU pipe -- 16 lanes per SM and clock -- would then need as long as the MMAs of the tile).
template <int ACT>
__device__ __forceinline__ void act_pair(float& x0, float& x1) {
  if constexpr (ACT == 0) {
    const f32x2 x = pk2(x0, x1);
    float h0, h1;
    upk2(mul2(x, bc2(0.5f)), h0, h1);
    const f32x2 sg = fma2(pk2(tanh_approx(h0), tanh_approx(h1)), bc2(0.5f), bc2(0.5f));
    upk2(mul2(x, sg), x0, x1);
  } else if constexpr (ACT == 1) {
    const f32x2 x = pk2(x0, x1), ax = pk2(fabsf(x0), fabsf(x1));
    f32x2 u, e;
    half_erfc_pair(x, ax, u, e);
    upk2(fma2(ax, u, pk2(fmaxf(x0, 0.f), fmaxf(x1, 0.f))), x0, x1);
  } else if constexpr (ACT == 2) {
    x0 = fmaxf(x0, 0.f);
    x1 = fmaxf(x1, 0.f);
  }
}
// (f0, f1) *= act'(x0, x1)
template <int ACT>
__device__ __forceinline__ void act_grad_mul_pair(float& f0, float& f1, float x0, float x1) {
  if constexpr (ACT == 0) {  // s (1 + x (1 - s))
    const f32x2 x = pk2(x0, x1);
    float h0, h1;
    upk2(mul2(x, bc2(0.5f)), h0, h1);
    const f32x2 sg = fma2(pk2(tanh_approx(h0), tanh_approx(h1)), bc2(0.5f), bc2(0.5f));
    const f32x2 gr = mul2(sg, fma2(x, fma2(sg, bc2(-1.f), bc2(1.f)), bc2(1.f)));
    upk2(mul2(pk2(f0, f1), gr), f0, f1);
  } else if constexpr (ACT == 1) {  // Phi(x) + x phi(x);  Phi = 0.5 + copysign(0.5 + u, x),  phi = e / sqrt(2 pi)
    const f32x2 x = pk2(x0, x1), ax = pk2(fabsf(x0), fabsf(x1));
    f32x2 u, e;
    half_erfc_pair(x, ax, u, e);
    float c0, c1;
    upk2(add2(u, bc2(0.5f)), c0, c1);
    const f32x2 cdf = add2(pk2(copysignf(c0, x0), copysignf(c1, x1)), bc2(0.5f));
    const f32x2 gr = fma2(mul2(x, bc2(0.3989422804014327f)), e, cdf);
    upk2(mul2(pk2(f0, f1), gr), f0, f1);
  } else if constexpr (ACT == 2) {
    f0 = x0 > 0.f ? f0 : 0.f;
    f1 = x1 > 0.f ? f1 : 0.f;
  }
}
// the activation code is a runtime argument: dispatch ONCE per 32-value chunk (warp-uniform), never per element -- a per-element
// `switch` made ptxas emit every activation's code for every element (measured: ~165 instructions per element, epilogue-bound GEMMs
// at 15-35 % of cuBLAS, profiles/r02/gemm_probe_r02c.jsonl)
template <int ACT>
__device__ __forceinline__ void act_chunk(float (&f)[32]) {
#pragma unroll
  for (int j = 0; j < 32; j += 2) act_pair<ACT>(f[j], f[j + 1]);
}
__device__ __forceinline__ void act_chunk_rt(float (&f)[32], int act) {
  if (act == 0) act_chunk<0>(f);
  else if (act == 1) act_chunk<1>(f);
  else if (act == 2) act_chunk<2>(f);
}
template <int ACT>
__device__ __forceinline__ void act_grad_mul8(float (&f)[32], int j, const float (&p)[8]) {
#pragma unroll
  for (int u = 0; u < 8; u += 2) act_grad_mul_pair<ACT>(f[j + u], f[j + u + 1], p[u], p[u + 1]);
}

// ---- the kernel --------------------------------------------------------------------------------------------------
// swizzled address of 16-byte chunk `c` of row `r` in a patch buffer
__device__ __forceinline__ uint32_t patch_addr(uint32_t buf, int r, int c) { return buf + r * 64 + ((c ^ ((r >> 1) & 3)) << 4); }

// One warp's patch buffers.  A buffer may be refilled only after the TMA store that read it has finished reading: lane 0 issues
// every store of the warp (bulk groups are per thread) and keeps at most one group pending while the other buffer is filled.
struct Patch {
  uint32_t base;
  uint32_t cur;  // 0 or kPatchBuf
  __device__ __forceinline__ uint32_t acquire(int lane) {  // the buffer to fill next (warp-uniform call)
    cur ^= (uint32_t)kPatchBuf;
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    __syncwarp();
    return base + cur;
  }
};

// bf16 chunk of 32 rows x 32 columns held one ROW per lane (pk[j] = columns 2j, 2j+1 of row `lane`, packed bf16x2) -> global
// memory: the lanes write their rows into a patch buffer, ONE lane hands the buffer to the TMA engine (which clips rows / columns
// beyond the matrix).  No shared-memory read-back, no store instructions, no address arithmetic in the warp: the epilogue of the
// activation shapes had the LSU data pipe 70 % busy (ncu, profiles/r02/ncu_gemm_perf_mlp1_r02w.md).
__device__ __forceinline__ void store_chunk_tma(Patch& p, int lane, const uint32_t (&pk)[16], const CUtensorMap* map, int col, int row0) {
  const uint32_t b = p.acquire(lane);
#pragma unroll
  for (int j = 0; j < 4; ++j) sts16(patch_addr(b, lane, j), make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]));
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the async proxy
  __syncwarp();
  if (lane == 0) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(col), "r"(row0), "r"(b)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
}
// global memory -> one ROW per lane, for the per-element inputs of the epilogue (residual, saved pre-activation, gathered node
// rows): 32 rows x 64 bytes, row r at `row_ptr` of lane r (0 = no such row).  With every lane loading ITS row, one load
// instruction touched 32 different 128-byte lines for 16 bytes each; here each instruction reads 8 rows x 64 contiguous bytes
// and the rows are handed out through a patch buffer.
__device__ __forceinline__ void load_chunk_rows(Patch& p, int lane, const void* row_ptr, int bytes_left, uint4 (&out)[4]) {
  const uint32_t b = p.acquire(lane);
  const int c16 = lane & 3, rsub = lane >> 2;
  const unsigned long long mine = reinterpret_cast<unsigned long long>(row_ptr);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = i * 8 + rsub;
    const unsigned long long q = __shfl_sync(0xffffffffu, mine, r);
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (q != 0ull && c16 * 16 < bytes_left) v = ldg16_keep(reinterpret_cast<const char*>(q) + c16 * 16);
    sts16(patch_addr(b, r, c16), v);
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 4; ++j) out[j] = lds16(patch_addr(b, lane, j));
  __syncwarp();
}
__device__ __forceinline__ void pack_chunk(const float (&f)[32], uint32_t (&pk)[16]) {
#pragma unroll
  for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(f[2 * j], f[2 * j + 1]);
}

template <int BN, bool A_MN, bool B_MN, int CG, int EW>
__global__ void __launch_bounds__(64 + 32 * EW, 1) gemm_tc_kernel(const __grid_constant__ AMaps tmAs,
                                                             const __grid_constant__ CUtensorMap tmB,
                                                             const __grid_constant__ OMaps tmO, const Args g) {
  using C = Cfg<BN, CG, EW>;
  constexpr int S = C::kStages;
  constexpr int TM = BM * CG;  // rows of the tile the CTA pair (or the single CTA) accumulates
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B atoms are 1024 B aligned
  const uint32_t bars = base + S * C::kStageBytes;
  auto full = [&](int s) { return bars + 8u * s; };
  auto empty = [&](int s) { return bars + 8u * (S + s); };
  auto tfull = [&](int a) { return bars + 8u * (2 * S + a); };
  auto tempty = [&](int a) { return bars + 8u * (2 * S + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * S + 4);
  const uint32_t patches = bars + C::kBarBytes;
  auto sA = [&](int s) { return base + s * C::kStageBytes; };
  auto sB = [&](int s) { return base + s * C::kStageBytes + C::kABytes; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta = CG == 2 ? cluster_ctarank() : 0u;  // 0 = leader (issues the MMAs, owns the full / tmem_empty barriers)
  const int cluster_id = CG == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int num_clusters = CG == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full(s), CG);  // one arrival per CTA's producer (the leader's carries the byte count of both)
      mbar_init(empty(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull(a), 1);
      mbar_init(tempty(a), CG * EW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if constexpr (CG == 2) cluster_sync_all();  // the peer's barriers exist before anything can signal them
  if (warp == 1) {  // TMEM: 512 columns (the whole SM's tensor memory; one CTA per SM)
    if constexpr (CG == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  // work item t -> (split, output tile): the tiles of ONE K range are consecutive, so the CTA pairs in flight at any time read
  // the same rows of both operands (split-K wgrad streams both from DRAM; with the splits of one tile consecutive instead, the
  // pairs in flight covered ~5 tiles x 16 K ranges and every operand slab came from DRAM once per tile that uses it)
  const int tiles_mn = g.tiles_m * g.tiles_n;
  const int num_tiles = tiles_mn * g.splits;

  if (warp == 0) {
    // ===== TMA producer (whole warp walks the loop, one elected lane issues) =====
    {
      uint32_t c = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        const int split = t / tiles_mn, r = t % tiles_mn;
        const int n0 = (r % g.tiles_n) * BN + (int)cta * C::kBRows, m0 = (r / g.tiles_n) * TM + (int)cta * BM;
        const int kb0 = split * g.kb_per_split, kb1 = min(kb0 + g.kb_per_split, g.kb_total);
        for (int kb = kb0; kb < kb1; ++kb, ++c) {
          const int s = c % S;
          mbar_wait(empty(s), ((c / S) & 1u) ^ 1u);
          uint32_t fb = full(s);
          if constexpr (CG == 2) fb = mapa(fb, 0);
          if (!elect_one()) continue;
          auto load = [&](uint32_t dst, const CUtensorMap* m, int c0, int c1) {
            if constexpr (CG == 2) tma_load_2d_cg2(dst, m, fb, c0, c1); else tma_load_2d(dst, m, fb, c0, c1);
          };
          if constexpr (!A_MN) {
            const int ka = kb * BK, seg = ka / g.a_seg_len;  // a_seg_len is a multiple of BK (or INT_MAX: one tensor)
            load(sA(s), &tmAs.m[seg], ka - seg * g.a_seg_len, m0);
          } else {
            const int seg = m0 / g.a_seg_len, ma = m0 - seg * g.a_seg_len;  // a_seg_len is a multiple of the tile rows
#pragma unroll
            for (int i = 0; i < BM / 64; ++i) load(sA(s) + i * kBoxBytes, &tmAs.m[seg], ma + 64 * i, kb * BK);
          }
          if constexpr (!B_MN) {
            load(sB(s), &tmB, kb * BK, n0);
          } else {
#pragma unroll
            for (int i = 0; i < C::kBRows / 64; ++i) load(sB(s) + i * kBoxBytes, &tmB, n0 + 64 * i, kb * BK);
          }
          // copies first, arrival second: the peer's arrival on the leader's barrier is a remote operation and must not sit in
          // front of its own loads (a `.release.cluster` arrival there cost an ERRBAR per k-block: 34 % tensor-pipe activity)
          if constexpr (CG == 2) {
            if (cta == 0) mbar_arrive_expect_tx(full(s), 2 * C::kStageBytes); else mbar_arrive_cluster(fb);
          } else {
            mbar_arrive_expect_tx(fb, C::kStageBytes);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer: warp 1 of the leader CTA walks the loops converged, one elected lane issues =====
    if (cta == 0) {
      // instruction descriptor: D fp32, A / B bf16, operand majors, N >> 3 at bit 17, M >> 4 at bit 24
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                             ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
      constexpr uint32_t kAStep = (A_MN ? (UMMA_K * 128) : (UMMA_K * 2)) >> 4;  // descriptor address units per UMMA_K
      constexpr uint32_t kBStep = (B_MN ? (UMMA_K * 128) : (UMMA_K * 2)) >> 4;
      // shared-memory descriptors: only the 14-bit start-address field (low word) changes with stage and k step, and it cannot
      // carry out of its field (shared memory ends below 256 KB), so the per-MMA work is one 32-bit add per operand
      const uint64_t da0 = A_MN ? smem_desc(sA(0), g.mn_lbo, g.mn_sbo) : smem_desc(sA(0), g.k_lbo, g.k_sbo);
      const uint64_t db0 = B_MN ? smem_desc(sB(0), g.mn_lbo, g.mn_sbo) : smem_desc(sB(0), g.k_lbo, g.k_sbo);
      const uint32_t da_hi = (uint32_t)(da0 >> 32), db_hi = (uint32_t)(db0 >> 32);
      const uint32_t da_lo0 = (uint32_t)da0, db_lo0 = (uint32_t)db0;
      auto desc64 = [](uint32_t hi, uint32_t lo) { return ((uint64_t)hi << 32) | lo; };
      uint32_t c = 0, it = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters, ++it) {
        const int split = t / tiles_mn;
        const int kb0 = split * g.kb_per_split, kb1 = min(kb0 + g.kb_per_split, g.kb_total);
        const uint32_t a = it & 1u;
        mbar_wait(tempty(a), ((it >> 1) & 1u) ^ 1u);  // the epilogue (of both CTAs) has drained this accumulator stage
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + a * BN;
        for (int kb = kb0; kb < kb1; ++kb, ++c) {
          const uint32_t s = c % S;
          mbar_wait(full(s), (c / S) & 1u);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t soff = (s * (uint32_t)C::kStageBytes) >> 4;
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t dak = desc64(da_hi, da_lo0 + soff + k * kAStep), dbk = desc64(db_hi, db_lo0 + soff + k * kBStep);
              const uint32_t acc = (kb > kb0 || k > 0) ? 1u : 0u;
              if constexpr (CG == 2) tc_mma_cg2(d_tmem, dak, dbk, idesc, acc); else tc_mma(d_tmem, dak, dbk, idesc, acc);
            }
            // stage reusable (in both CTAs) once these MMAs have read it
            if constexpr (CG == 2) tc_commit_cg2(empty(s)); else tc_commit(empty(s));
            if (kb == kb1 - 1) {  // accumulator complete
              if constexpr (CG == 2) tc_commit_cg2(tfull(a)); else tc_commit(tfull(a));
            }
          }
          __syncwarp();
        }
      }
      if constexpr (CG == 2) {
        // the peer's epilogue arrives on THIS CTA's tmem_empty barriers: see its last arrivals in before leaving
        if (it >= 1) mbar_wait(tempty((it - 1) & 1u), ((it - 1) >> 1) & 1u);
        if (it >= 2) mbar_wait(tempty((it - 2) & 1u), ((it - 2) >> 1) & 1u);
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue =====
    const int e = warp - 2;
    const int quad = warp & 3;  // a warp may only touch TMEM lanes [32 * (warp % 4), +32)
    const int half = e >> 2;    // EW / 4 warps per lane quadrant: each takes an equal share of the BN columns
    constexpr int kChunks = BN / 32 / (EW / 4);
    const Epi& ep = g.epi;
    Patch patch;
    patch.base = patches + e * kPatchBytes;
    patch.cur = 0;
    uint32_t it = 0;
    for (int t = cluster_id; t < num_tiles; t += num_clusters, ++it) {
      const int split = t / tiles_mn, r = t % tiles_mn;
      const int n0 = (r % g.tiles_n) * BN, m0 = (r / g.tiles_n) * TM + (int)cta * BM;
      const uint32_t a = it & 1u;
      mbar_wait(tfull(a), (it >> 1) & 1u);
      tc_fence_after();
      const int row0 = m0 + quad * 32;  // first row of this warp's 32 x 32 chunks
      const int row = row0 + lane;
      const bool row_ok = row < g.M;
      float rs = 1.f, rt = 0.f;
      if (ep.row_scale != nullptr && row_ok) {
        rs = ep.row_scale[row];
        rt = ep.row_shift[row];
      }
      long long grow[2] = {0, 0};  // this thread's row of each gather table (one index load per tile)
#pragma unroll
      for (int i = 0; i < 2; ++i)
        if (ep.gather[i] != nullptr && row_ok) grow[i] = ep.gather_idx[i][row];
#pragma unroll 1
      for (int ch = 0; ch < kChunks; ++ch) {
        const int cc = (half * kChunks + ch) * 32;
        uint32_t v[32];
        tc_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + a * BN + cc, v);
        tc_wait_ld();
        if (ch == kChunks - 1) {  // this warp has read its share of the stage: hand it back before the stores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (CG == 2) mbar_arrive_cluster(mapa(tempty(a), 0)); else mbar_arrive(tempty(a));
          }
        }
        const int col0 = n0 + cc;
        if (row0 >= g.M || col0 >= g.N) continue;  // warp-uniform
        const int cols_left = g.N - col0;
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        if (ep.row_scale != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (j < cols_left) {
              const float4 cv = __ldg(reinterpret_cast<const float4*>(ep.col_vec + col0 + j));
              f[j] = rs * f[j] + rt * cv.x;
              f[j + 1] = rs * f[j + 1] + rt * cv.y;
              f[j + 2] = rs * f[j + 2] + rt * cv.z;
              f[j + 3] = rs * f[j + 3] + rt * cv.w;
            }
          }
        }
        if (ep.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (j < cols_left) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + j));
              add_pair(f[j], f[j + 1], b.x, b.y);
              add_pair(f[j + 2], f[j + 3], b.z, b.w);
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          if (ep.gather[i] != nullptr) {  // warp-uniform
            uint4 rows[4];
            load_chunk_rows(patch, lane, row_ok ? ep.gather[i] + grow[i] * ep.ld_gather + col0 : nullptr, cols_left * 2, rows);
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              float p[8];
              unpack<__nv_bfloat16>(rows[j / 8], p);  // zeros beyond the matrix
#pragma unroll
              for (int u = 0; u < 8; u += 2) add_pair(f[j + u], f[j + u + 1], p[u], p[u + 1]);
            }
          }
        }
        if (ep.dact_pre != nullptr) {
          uint4 rows[4];
          load_chunk_rows(patch, lane, row_ok ? reinterpret_cast<const __nv_bfloat16*>(ep.dact_pre) + (size_t)row * g.N + col0 : nullptr,
                          cols_left * 2, rows);
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            float p[8];
            unpack<__nv_bfloat16>(rows[j / 8], p);
            if (ep.act == 0) act_grad_mul8<0>(f, j, p);
            else if (ep.act == 1) act_grad_mul8<1>(f, j, p);
            else if (ep.act == 2) act_grad_mul8<2>(f, j, p);
          }
        } else if (ep.act != 3) {
          if (ep.pre_out != nullptr) {
            // the activation sees the bf16-rounded pre-activation, so that backward (which only has the rounded value)
            // differentiates exactly the function forward applied
            uint32_t pk[16];
            pack_chunk(f, pk);
            store_chunk_tma(patch, lane, pk, &tmO.pre, col0, row0);
#pragma unroll
            for (int j = 0; j < 16; ++j) {  // bf16 -> fp32 is a shift / a mask
              f[2 * j] = __uint_as_float(pk[j] << 16);
              f[2 * j + 1] = __uint_as_float(pk[j] & 0xffff0000u);
            }
          }
          act_chunk_rt(f, ep.act);
        }
        if (ep.residual != nullptr) {
          if (ep.res_f32) {  // 32 fp32 columns = two 64-byte halves
            const float* rp = reinterpret_cast<const float*>(ep.residual) + (size_t)row * ep.ld_res + col0;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              uint4 rows[4];
              load_chunk_rows(patch, lane, row_ok ? rp + 16 * h : nullptr, (cols_left - 16 * h) * 4, rows);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float p[4];
                unpack<float>(rows[j], p);
                add_pair(f[16 * h + 4 * j], f[16 * h + 4 * j + 1], p[0], p[1]);
                add_pair(f[16 * h + 4 * j + 2], f[16 * h + 4 * j + 3], p[2], p[3]);
              }
            }
          } else {
            uint4 rows[4];
            load_chunk_rows(patch, lane,
                            row_ok ? reinterpret_cast<const __nv_bfloat16*>(ep.residual) + (size_t)row * ep.ld_res + col0 : nullptr,
                            cols_left * 2, rows);
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              float p[8];
              unpack<__nv_bfloat16>(rows[j / 8], p);
#pragma unroll
              for (int u = 0; u < 8; u += 2) add_pair(f[j + u], f[j + u + 1], p[u], p[u + 1]);
            }
          }
        }
        const int seg = col0 / ep.seg_cols;
        const int cs = col0 - seg * ep.seg_cols;
        if (ep.out_f32) {
          if (row_ok) {
            float* op = reinterpret_cast<float*>(ep.out[seg]) + (size_t)split * g.split_stride + (size_t)row * ep.ld_out + cs;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (j < cols_left) {
                float p[4] = {f[j], f[j + 1], f[j + 2], f[j + 3]};
                stg16(op + j, pack<float>(p));
              }
            }
          }
        } else {
          uint32_t pk[16];
          pack_chunk(f, pk);
          store_chunk_tma(patch, lane, pk, &tmO.out[seg], cs, row0);
        }
      }
    }
  }

  if (warp >= 2 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // this warp's TMA stores are complete
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CG == 2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// out[i] = sum_s part[s][i]  (fixed order), cast to bf16 or kept fp32
__global__ void splitk_reduce_kernel(const float* __restrict__ part, long long n, int splits, long long stride, void* out,
                                     int out_f32) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  float4 acc = *reinterpret_cast<const float4*>(part + i);
  for (int s = 1; s < splits; ++s) {
    const float4 p = *reinterpret_cast<const float4*>(part + (long long)s * stride + i);
    acc.x += p.x;
    acc.y += p.y;
    acc.z += p.z;
    acc.w += p.w;
  }
  if (out_f32) {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + i) = acc;
  } else {
    uint2 o = make_uint2(pack_bf16x2(acc.x, acc.y), pack_bf16x2(acc.z, acc.w));
    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + i) = o;
  }
}

// ---- host side ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &st) == cudaSuccess && st == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// 2-D bf16 tensor map: `inner` contiguous elements, `outer` rows `ld` elements apart; box = 64 x box_rows, SWIZZLE_128B
static int make_map(CUtensorMap* m, const void* ptr, long long inner, long long outer, long long ld, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) return fail(AB2_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15u) != 0 || (ld * 2) % 16 != 0)
    return fail(AB2_ERR_INVALID, "gemm operand must be 16-byte aligned with a row stride that is a multiple of 8 elements");
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(AB2_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return AB2_OK;
}

// output map: bf16 [outer rows][inner columns], row stride ld elements; box = 32 columns x 32 rows, SWIZZLE_64B (TMA stores)
static int make_out_map(CUtensorMap* m, const void* ptr, long long inner, long long outer, long long ld) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) return fail(AB2_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15u) != 0 || (ld * 2) % 16 != 0)
    return fail(AB2_ERR_INVALID, "gemm bf16 output must be 16-byte aligned with a row stride that is a multiple of 8 elements");
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {32u, 32u};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(AB2_ERR_CUDA, "cuTensorMapEncodeTiled (output) failed with CUresult %d", (int)r);
  return AB2_OK;
}

template <int BN, bool A_MN, bool B_MN, int CG, int EW>
static int launch(const AMaps& ta, const CUtensorMap& tb, const OMaps& to, const Args& a, cudaStream_t st) {
  using C = Cfg<BN, CG, EW>;
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN, CG, EW>;
  if (!AB2_ENSURE_DYN_SMEM(kern, C::kSmemBytes)) return fail(AB2_ERR_CUDA, "gemm_tc_kernel: cannot reserve %d B of shared memory", C::kSmemBytes);
  const int tiles = a.tiles_m * a.tiles_n * a.splits;
  const int slots = num_sms() / CG;  // CTAs (CG = 1) or CTA pairs (CG = 2) that can be resident
  int grid = tiles < slots ? tiles : slots;
  const char* env = getenv("AB2_GEMM_GRID");
  if (env != nullptr && atoi(env) > 0 && atoi(env) < grid) grid = atoi(env);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(grid * CG));
  cfg.blockDim = dim3(C::kThreads);
  cfg.dynamicSmemBytes = C::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  AB2_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, ta, tb, to, a));
  AB2_LAUNCH_OK("gemm_tc_kernel");
  return AB2_OK;
}

}  // namespace tc
}  // namespace ab2

using namespace ab2;

extern "C" size_t ab2_gemm_workspace_bytes(const ab2_gemm* d) {
  if (d == nullptr || d->splits <= 1) return 0;
  return (size_t)d->splits * (size_t)d->M * (size_t)d->N * sizeof(float);
}

// cuTensorMapEncodeTiled is a DRIVER API call and needs a context current on the calling thread.  A thread that has only ever been
// handed device 0 by PyTorch (an autograd worker whose first operation is one of these GEMMs) has made no runtime call yet, so
// nothing has bound the primary context: the encode then fails with CUDA_ERROR_INVALID_CONTEXT (201).  One runtime call binds it.
static void ensure_context_on_this_thread() {
  static thread_local bool bound = false;
  if (!bound) {
    cudaFree(nullptr);
    bound = true;
  }
}

extern "C" int ab2_gemm_bf16(const ab2_gemm* d, void* workspace, size_t workspace_bytes, void* stream) {
  if (d == nullptr) return fail(AB2_ERR_INVALID, "ab2_gemm_bf16: null descriptor");
  const long long M = d->M, N = d->N, K = d->K;
  if (M < 0 || N <= 0 || K <= 0 || M >= (1LL << 31) || N >= (1LL << 31) || K >= (1LL << 31))
    return fail(AB2_ERR_INVALID, "ab2_gemm_bf16: bad sizes M=%lld N=%lld K=%lld", M, N, K);
  if (M == 0) return AB2_OK;
  if (N % 8 != 0) return fail(AB2_ERR_UNSUPPORTED, "ab2_gemm_bf16: N must be a multiple of 8 (got %lld)", N);
  if (d->a == nullptr || d->b == nullptr || d->out[0] == nullptr) return fail(AB2_ERR_INVALID, "ab2_gemm_bf16: null operand");
  const int splits = d->splits > 1 ? d->splits : 1;
  tc::Args a;
  memset(&a, 0, sizeof(a));
  a.M = (int)M;
  a.N = (int)N;
  a.K = (int)K;
  const int BN = (N > 128) ? 256 : 128;
  int CG = 2;  // CTA pairs (cta_group::2) unless asked otherwise (A/B experiments)
  if (const char* cg = getenv("AB2_GEMM_CG")) CG = atoi(cg) == 1 ? 1 : 2;
  a.tiles_m = (int)((M + tc::BM * CG - 1) / (tc::BM * CG));
  a.tiles_n = (int)((N + BN - 1) / BN);
  a.kb_total = (int)((K + tc::BK - 1) / tc::BK);
  a.splits = splits;
  a.kb_per_split = (a.kb_total + splits - 1) / splits;
  a.k_lbo = 16;
  a.k_sbo = 1024;
  a.mn_lbo = tc::kBoxBytes;
  a.mn_sbo = 1024;
  if (const char* dbg = getenv("AB2_GEMM_DESC")) sscanf(dbg, "%d,%d,%d,%d", &a.k_lbo, &a.k_sbo, &a.mn_lbo, &a.mn_sbo);
  if ((long long)a.kb_per_split * (splits - 1) >= a.kb_total && splits > 1)
    return fail(AB2_ERR_INVALID, "ab2_gemm_bf16: %d splits leave an empty split for %d k-blocks", splits, a.kb_total);
  tc::Epi& e = a.epi;
  for (int i = 0; i < 4; ++i) e.out[i] = d->out[i];
  e.ld_out = d->ld_out;
  e.seg_cols = d->seg_cols > 0 ? d->seg_cols : (int)(((N + 31) / 32) * 32);
  if (e.seg_cols % 32 != 0) return fail(AB2_ERR_INVALID, "ab2_gemm_bf16: seg_cols must be a multiple of 32");
  const long long nseg = (N + e.seg_cols - 1) / e.seg_cols;
  if (nseg > 4) return fail(AB2_ERR_INVALID, "ab2_gemm_bf16: at most 4 output segments");
  for (int i = 0; i < nseg; ++i)
    if (e.out[i] == nullptr) return fail(AB2_ERR_INVALID, "ab2_gemm_bf16: output segment %d is null", i);
  e.out_f32 = d->out_f32;
  e.bias = d->bias;
  e.row_scale = d->row_scale;
  e.row_shift = d->row_shift;
  e.col_vec = d->col_vec;
  if (e.row_scale != nullptr && (e.row_shift == nullptr || e.col_vec == nullptr))
    return fail(AB2_ERR_INVALID, "ab2_gemm_bf16: row_scale needs row_shift and col_vec");
  e.act = d->act;
  e.pre_out = d->pre_out;
  e.dact_pre = d->dact_pre;
  e.residual = d->residual;
  e.ld_res = d->ld_res;
  e.res_f32 = d->res_f32;
  e.gather[0] = reinterpret_cast<const __nv_bfloat16*>(d->gather_a);
  e.gather[1] = reinterpret_cast<const __nv_bfloat16*>(d->gather_b);
  e.gather_idx[0] = reinterpret_cast<const long long*>(d->gather_a_idx);
  e.gather_idx[1] = reinterpret_cast<const long long*>(d->gather_b_idx);
  e.ld_gather = d->ld_gather;
  for (int i = 0; i < 2; ++i)
    if ((e.gather[i] != nullptr) != (e.gather_idx[i] != nullptr)) return fail(AB2_ERR_INVALID, "ab2_gemm_bf16: a gather table needs its index array");
  void* final_out = e.out[0];
  if (splits > 1) {
    if (nseg != 1 || e.bias || e.row_scale || e.act != 3 || e.dact_pre || e.residual || e.pre_out || e.gather[0] || e.gather[1])
      return fail(AB2_ERR_INVALID, "ab2_gemm_bf16: split-K supports a plain single output only");
    const size_t need = (size_t)splits * M * N * sizeof(float);
    if (workspace == nullptr || workspace_bytes < need)
      return fail(AB2_ERR_INVALID, "ab2_gemm_bf16: split-K workspace too small (%zu < %zu)", workspace_bytes, need);
    if (d->ld_out != N) return fail(AB2_ERR_INVALID, "ab2_gemm_bf16: split-K needs a contiguous output (ld_out == N)");
    e.out[0] = workspace;
    e.out_f32 = 1;
    a.split_stride = M * N;
  }
  ensure_context_on_this_thread();
  tc::AMaps ta;
  CUtensorMap tb;
  int rc;
  // K-major operand: memory [rows][K] -> inner = K, outer = rows, box 64 x (rows this CTA stages).  MN-major: memory [K][rows]
  // -> inner = rows, outer = K, 64 x 64 boxes.
  a.a_seg_len = INT_MAX;
  int nsegA = 1;
  const void* aptr[4] = {d->a, d->a_seg[0], d->a_seg[1], d->a_seg[2]};
  if (d->a_seg_len > 0) {
    const long long extent = d->a_mn ? M : K;  // the segmented dimension: K of a K-major A, M of an MN-major A
    const long long unit = d->a_mn ? (long long)tc::BM * CG : tc::BK;
    nsegA = (int)((extent + d->a_seg_len - 1) / d->a_seg_len);
    if (d->a_seg_len % unit != 0 || extent % d->a_seg_len != 0 || nsegA > 4 || d->a_seg_len >= INT_MAX)
      return fail(AB2_ERR_INVALID, "ab2_gemm_bf16: a_seg_len must be a multiple of %lld that divides %lld into at most 4 segments", unit, extent);
    for (int i = 1; i < nsegA; ++i)
      if (aptr[i] == nullptr) return fail(AB2_ERR_INVALID, "ab2_gemm_bf16: A segment %d is null", i);
    a.a_seg_len = (int)d->a_seg_len;
  }
  for (int i = 0; i < nsegA; ++i) {
    const long long segM = (d->a_mn && nsegA > 1) ? d->a_seg_len : M, segK = (!d->a_mn && nsegA > 1) ? d->a_seg_len : K;
    rc = d->a_mn ? tc::make_map(&ta.m[i], aptr[i], segM, K, d->lda, 64) : tc::make_map(&ta.m[i], aptr[i], segK, M, d->lda, tc::BM);
    if (rc) return rc;
  }
  for (int i = nsegA; i < 4; ++i) ta.m[i] = ta.m[0];  // unused slots hold a valid map (no second encode: ~1 us of host time each)
  rc = d->b_mn ? tc::make_map(&tb, d->b, N, K, d->ldb, 64) : tc::make_map(&tb, d->b, K, N, d->ldb, BN / CG);
  if (rc) return rc;
  tc::OMaps to;
  memset(&to, 0, sizeof(to));
  if (!e.out_f32) {  // bf16 results leave through TMA stores; fp32 ones (split-K partials, fp32 outputs) are stored directly
    for (int i = 0; i < nseg; ++i) {
      const long long width = std::min<long long>(e.seg_cols, N - (long long)i * e.seg_cols);
      rc = tc::make_out_map(&to.out[i], e.out[i], width, M, e.ld_out);
      if (rc) return rc;
    }
  }
  if (e.pre_out != nullptr) {
    rc = tc::make_out_map(&to.pre, e.pre_out, N, M, N);
    if (rc) return rc;
  }
  cudaStream_t st = (cudaStream_t)stream;
#define AB2_GEMM_DISPATCH(BNV, CGV, EWV)                                                       \
  do {                                                                                         \
    if (!d->a_mn && !d->b_mn) rc = tc::launch<BNV, false, false, CGV, EWV>(ta, tb, to, a, st);     \
    else if (!d->a_mn && d->b_mn) rc = tc::launch<BNV, false, true, CGV, EWV>(ta, tb, to, a, st);  \
    else if (d->a_mn && d->b_mn) rc = tc::launch<BNV, true, true, CGV, EWV>(ta, tb, to, a, st);    \
    else rc = tc::launch<BNV, true, false, CGV, EWV>(ta, tb, to, a, st);                           \
  } while (0)
  // epilogue warps: 16 when the epilogue does special-function work or gathers per element (activation, activation derivative,
  // row tables), 8 (and one more pipeline stage) for bias / residual epilogues -- measured both ways on every block shape
  // (profiles/r02/gemm_probe_r02h_ew8.jsonl vs _ew16.jsonl); AB2_GEMM_EW overrides for A/B runs
  int EW = (e.act != 3 || e.dact_pre || e.gather[0] || e.gather[1]) ? 16 : 8;
  if (const char* ew = getenv("AB2_GEMM_EW")) EW = atoi(ew) == 16 ? 16 : 8;
  if (CG == 1) {  // single-CTA tiles: A/B experiments only
    if (BN == 256) AB2_GEMM_DISPATCH(256, 1, 8); else AB2_GEMM_DISPATCH(128, 1, 8);
  } else if (EW == 16) {
    if (BN == 256) AB2_GEMM_DISPATCH(256, 2, 16); else AB2_GEMM_DISPATCH(128, 2, 16);
  } else {
    if (BN == 256) AB2_GEMM_DISPATCH(256, 2, 8); else AB2_GEMM_DISPATCH(128, 2, 8);
  }
#undef AB2_GEMM_DISPATCH
  if (rc) return rc;
  if (splits > 1) {
    const long long n = M * N;  // N % 8 == 0 -> n % 4 == 0
    const int threads = 256;
    const long long blocks = (n / 4 + threads - 1) / threads;
    tc::splitk_reduce_kernel<<<(unsigned)blocks, threads, 0, st>>>(reinterpret_cast<const float*>(workspace), n, splits, M * N, final_out,
                                                                    d->out_f32);
    AB2_LAUNCH_OK("splitk_reduce_kernel");
  }
  return AB2_OK;
}
