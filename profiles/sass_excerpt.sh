#!/bin/bash
# Blackwell-specific instructions in the shipped library, per kernel:
#   bash profiles/sass_excerpt.sh > profiles/r2_sass.txt
# (cuobjdump -sass of smvs_b200/libsmvs_b200.so, built for sm_100a)
LIB=${1:-smvs_b200/libsmvs_b200.so}
echo "# cuobjdump -sass $LIB | per-kernel counts of the instructions the design relies on"
echo "# UBLKCP = cp.async.bulk (TMA engine, 1-D bulk copy), SYNCS = mbarrier, LDG.E.NA.EFL2.256 = 256-bit"
echo "# no-L1-allocate evict-first load, VIMNMX3 / VIADDMNMX = DPX 16x2 min, HSET2 / HFMA2 = fp16x2 census,"
echo "# DFMA = fp64 FMA, REDUX / CREDUX = warp reduction to a uniform register; grid barrier of cg_kernel: MEMBAR.ALL.GPU + REDG.E.ADD.STRONG.GPU (red.release.gpu), LDG.E.STRONG.GPU + CCTL.IVALL (ld.acquire.gpu)"
cuobjdump -sass "$LIB" 2>/dev/null | awk '
/Function : /{name=$3; sub(/^_ZN5smvsb/,"",name); next}
{
  n=split("UBLKCP UTMALDG SYNCS LDG.E.NA.EFL2.256 LDG.E.128 VIMNMX3 VIADDMNMX HSET2 HFMA2 DFMA DADD DMUL REDUX MEMBAR.ALL.GPU REDG.E.ADD.STRONG.GPU LDG.E.STRONG.GPU CCTL.IVALL ATOMG", pats, " ")
  for (i=1;i<=n;i++) if (index($0, pats[i])>0) cnt[name,pats[i]]++
  names[name]=1
}
END{
  for (k in names){
    line=""
    for (i=1;i<=n;i++) if (cnt[k,pats[i]]>0) line=line sprintf(" %s=%d", pats[i], cnt[k,pats[i]])
    if (line!="") print k ":" line
  }
}' | sed 's/_GLOBAL__N__[0-9a-f_]*_cu_[0-9a-f]*//' | sort
echo
echo "# excerpts"
for pat in "UBLKCP" "SYNCS.ARRIVE" "REDG.E.ADD.STRONG.GPU" "LDG.E.STRONG.GPU R" "CCTL.IVALL" "LDG.E.NA.EFL2.256" "VIMNMX3.U16x2" "VIADDMNMX.U16x2" "HSET2.BF.LT" "HFMA2" "CREDUX\|REDUX"; do
  echo "## $pat"
  cuobjdump -sass "$LIB" 2>/dev/null | grep -m 3 "$pat" | sed 's/^ *//' | cut -c1-120
done
