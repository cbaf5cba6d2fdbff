#!/bin/bash
# SASS evidence of the Blackwell paths: per kernel of libcagroup3d_b200.so, the counts of the tcgen05 / TMEM /
# bulk-copy mnemonics.   bash tools/sass_summary.sh > profiles/r2_sass_summary.txt
LIB=cagroup3d_b200/libcagroup3d_b200.so
echo "# cuobjdump -sass $LIB | per-function counts (sm_100a); UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit,"
echo "# UBLKCP = cp.async.bulk, UTMALDG = cp.async.bulk.tensor (TMA), LDGSTS = cp.async, SYNCS = mbarrier"
printf "%-78s %8s %6s %6s %7s %7s %8s %7s %6s\n" kernel UTCHMMA LDTM STTM UTCBAR UBLKCP UTMALDG LDGSTS SYNCS
cuobjdump -sass $LIB | awk '
/Function :/ { if (name != "") emit(); name=$3; for (k in c) delete c[k]; next }
/UTCHMMA/ {c["UTCHMMA"]++} /LDTM/ {c["LDTM"]++} /STTM/ {c["STTM"]++} /UTCBAR/ {c["UTCBAR"]++} /UBLKCP/ {c["UBLKCP"]++}
/UTMALDG/ {c["UTMALDG"]++} /LDGSTS/ {c["LDGSTS"]++} /SYNCS/ {c["SYNCS"]++}
function emit() { if (c["UTCHMMA"]+c["LDTM"]+c["STTM"]+c["UBLKCP"]+c["UTMALDG"] > 0) printf "%-78s %8d %6d %6d %7d %7d %8d %7d %6d\n", substr(name,1,78), c["UTCHMMA"], c["LDTM"], c["STTM"], c["UTCBAR"], c["UBLKCP"], c["UTMALDG"], c["LDGSTS"], c["SYNCS"] }
END { emit() }' | while read -r line; do n=$(echo "$line" | awk '{print $1}'); d=$(echo "$n" | c++filt 2>/dev/null | sed 's/(anonymous namespace):://; s/<unnamed>:://; s/void //' | cut -c1-78); echo "$line" | awk -v d="$d" '{printf "%-78s %8s %6s %6s %7s %7s %8s %7s %6s\n", d, $2,$3,$4,$5,$6,$7,$8,$9}'; done
