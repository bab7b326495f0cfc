#!/bin/bash
# SASS opcode histogram of the shipped library (profiles/r2_sass_histogram.txt):  bash tools/sass_histogram.sh > profiles/r2_sass_histogram.txt
cd "$(dirname "$0")/.."
SO=self-supervised-vision_b200/ssv_b200/libssv_b200.so
TMP=$(mktemp)
cuobjdump -sass "$SO" | grep -E '^\s+/\*[0-9a-f]{4}\*/' > "$TMP"
echo "# SASS opcode histogram of $SO (cuobjdump -sass, sm_100a, all kernels; end of round 2)"
echo "# tcgen05.mma -> UTCHMMA, tcgen05.ld/st -> LDTM/STTM, cp.async.bulk.tensor load / store / reduce -> UTMALDG / UTMASTG / UTMAREDG,"
echo "# cp.async.bulk -> UBLKCP, mbarrier -> SYNCS; HMMA / HGMMA (legacy mma.sync / wgmma) must be 0.  multimem.st compiles to"
echo "# STG.E.*.STRONG.SYS on a multicast address."
echo "# total instructions: $(wc -l < "$TMP")"
for m in UTCHMMA UTCQMMA UTCBAR LDTM STTM UTMALDG UTMASTG UTMAREDG UBLKCP UTCCP HMMA HGMMA MUFU.EX2 SYNCS MULTIMEM FFMA2 FADD2 FMUL2 RED.E SHFL; do
  printf "%-10s %s\n" "$m" "$(grep -cE "[^A-Z]$m" "$TMP")"
done
rm -f "$TMP"
