#!/bin/bash
# SASS evidence of the Blackwell paths: per kernel of libcagroup3d_b200.so, the counts of the tcgen05 / TMEM /
# bulk-copy mnemonics.   bash tools/sass_summary.sh > profiles/r2_sass_summary.txt
LIB=cagroup3d_b200/libcagroup3d_b200.so
echo "# cuobjdump -sass $LIB | per-function instruction counts (sm_100a); UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st,"
echo "# UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk, LDGSTS = cp.async, SYNCS = mbarrier ops; kernels without any of them omitted"
printf "%8s %5s %5s %7s %7s %7s %6s  %s\n" UTCHMMA LDTM STTM UTCBAR UBLKCP LDGSTS SYNCS kernel
cuobjdump -sass $LIB | awk '
/Function :/ { if (name != "") emit(); name=$3; for (k in c) delete c[k]; next }
/^ +\/\*[0-9a-f]+\*\/ / {
  if ($0 ~ /UTCHMMA/) c["UTCHMMA"]++; if ($0 ~ / LDTM/) c["LDTM"]++; if ($0 ~ / STTM/) c["STTM"]++; if ($0 ~ /UTCBAR/) c["UTCBAR"]++;
  if ($0 ~ /UBLKCP/) c["UBLKCP"]++; if ($0 ~ /LDGSTS/) c["LDGSTS"]++; if ($0 ~ /SYNCS/) c["SYNCS"]++ }
function emit() { if (c["UTCHMMA"]+c["LDTM"]+c["STTM"]+c["UBLKCP"] > 0) printf "%8d %5d %5d %7d %7d %7d %6d  %s\n", c["UTCHMMA"], c["LDTM"], c["STTM"], c["UTCBAR"], c["UBLKCP"], c["LDGSTS"], c["SYNCS"], name }
END { emit() }' | while read -r a b c d e f g name; do
  dn=$(echo "$name" | cu++filt 2>/dev/null | python3 -c "import sys,re; t=sys.stdin.read().strip(); t=re.sub(r'\\(anonymous namespace\\)::|<unnamed>::|^void |\\((int|bool)\\)','',t); print(re.sub(r'\\((Tc|Ts|Wg)Args.*','',t))")
  printf "%8s %5s %5s %7s %7s %7s %6s  %s\n" $a $b $c $d $e $f $g "$dn"
done
