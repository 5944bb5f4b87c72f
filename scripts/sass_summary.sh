#!/bin/bash
# SASS evidence without a GPU: per hot kernel (bf16, LPH=8 = the headline shape), registers and the counts of the
# instructions that carry the design (bulk copies, mbarrier waits, 128-bit loads/stores, shuffles).
LIB=anemoi_models_b200/lib/libanemoi_b200.so
OUT=${1:-profiles/r01/sass_summary.txt}
{
echo "# cuobjdump -sass / -res-usage of $LIB (sm_100a), kernels of the headline shape (bf16, 8 lanes per head)"
for k in gtconv_fwd_tma_kernelI13__nv_bfloat16Li8 gtconv_bwd_dst_tma_kernelI13__nv_bfloat16Li8 gtconv_bwd_src_tma_kernelI13__nv_bfloat16Li8 \
         gtconv_bwd_src_warp_kernelI13__nv_bfloat16Li8ELb0 gtconv_fwd_kernelI13__nv_bfloat16Li8ELb0 gtconv_bwd_dst_kernelI13__nv_bfloat16Li8ELb0; do
  fn=$(cuobjdump -res-usage $LIB 2>/dev/null | grep -o "Function [^:]*$k[^:]*" | head -1 | sed 's/Function //')
  [ -z "$fn" ] && continue
  res=$(cuobjdump -res-usage $LIB 2>/dev/null | grep -A1 "Function $fn:" | tail -1)
  sass=$(cuobjdump -sass -fun "$fn" $LIB 2>/dev/null)
  c() { echo "$sass" | grep -cE "$1"; }
  echo
  echo "## $(echo $fn | c++filt | cut -c1-110)"
  echo "   $res"
  echo "   UBLKCP (cp.async.bulk, TMA engine): $(c 'UBLKCP')   SYNCS (mbarrier arrive / try_wait): $(c 'SYNCS')   LDS.128: $(c 'LDS\.128')"
  echo "   LDG.E.128: $(c 'LDG\.E\.128')   STG.E(.NA).128: $(c 'STG\.E.*128')   SHFL: $(c 'SHFL')   VOTE: $(c 'VOTE')   MUFU.EX2: $(c 'MUFU\.EX2')   FFMA: $(c 'FFMA')"
done
} > $OUT
cat $OUT
