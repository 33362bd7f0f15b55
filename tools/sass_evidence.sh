#!/bin/bash
# Static evidence from the built library (no GPU needed): per-kernel register / stack / static-smem usage and the
# tensor-core / TMEM / async-copy SASS mnemonics per kernel.  bash tools/sass_evidence.sh > profiles/r2_sass_resource_usage.txt
LIB=${1:-sac_expert_b200/libsaceo.so}
echo "# $(date -u +%FT%TZ)  $LIB  ($(stat -c %s $LIB) bytes), nvcc $(nvcc --version | grep -o 'V[0-9.]*')"
echo "## resource usage (cuobjdump --dump-resource-usage): REG / STACK (spill frame) / static SHARED; dynamic smem is set at launch"
cuobjdump --dump-resource-usage $LIB 2>/dev/null | grep -A1 "Function" | paste - - | sed 's/^ Function //; s/CONSTANT.*//' | while read name rest; do
  echo "$(echo $name | c++filt | sed 's/saceo:://; s/(.*//') | $rest"; done | sort
echo
echo "## SASS mnemonics per kernel (cuobjdump -sass): UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTCBAR = tcgen05.commit,"
echo "## LDGSTS = cp.async, SYNCS = mbarrier, UTCCP = tcgen05.cp, UBLKCP/UTMALDG = TMA bulk copies (cp.async.bulk), UBLKPF = cp.async.bulk.prefetch.L2, HMMA = mma.sync (warp-level tensor path)"
cuobjdump -sass $LIB 2>/dev/null | awk '
  /Function :/ { name=$3 }
  /UTCHMMA/ {c[name,"UTCHMMA"]++} /LDTM/ {c[name,"LDTM"]++} /STTM/ {c[name,"STTM"]++} /UTCBAR/ {c[name,"UTCBAR"]++}
  /LDGSTS/ {c[name,"LDGSTS"]++} /SYNCS/ {c[name,"SYNCS"]++} /UTCCP/ {c[name,"UTCCP"]++} /UBLKCP|UTMALDG/ {c[name,"TMA"]++} /UBLKPF/ {c[name,"UBLKPF"]++} /HMMA\./ {c[name,"HMMA"]++}
  /Function :/ { names[name]=1 }
  END { for (n in names) { s=""; split("UTCHMMA LDTM STTM UTCBAR LDGSTS SYNCS UTCCP TMA UBLKPF HMMA", k, " ");
          for (i=1;i<=10;i++) if (c[n,k[i]]>0) s=s" "k[i]"="c[n,k[i]]; if (s!="") print n" :"s } }' | while read name rest; do
  echo "$(echo $name | c++filt | sed 's/saceo:://; s/(.*//') $rest"; done | sort
