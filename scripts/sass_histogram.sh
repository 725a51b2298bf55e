#!/bin/bash
# SASS opcode histogram of the in-tree product library: evidence that the conv engine is tcgen05 (UTCHMMA, LDTM/STTM, UTCBAR) fed by
# TMA (UTMALDG / UTMASTG / UTMAREDG), per kernel.  Runs anywhere cuobjdump is (no GPU needed):  bash scripts/sass_histogram.sh > profiles/rNN_sass_histogram.txt
LIB=${1:-fcd-gan-pytorch_b200/libfcd_b200.so}
echo "# cuobjdump -sass $LIB  ($(date -u +%Y-%m-%d), $(cuobjdump --version | tail -1))"
cuobjdump -sass "$LIB" > /tmp/fcd_sass.txt
echo "# whole library"
for op in UTCHMMA UTCQMMA UTCBAR UTCCP LDTM STTM UTMALDG UTMASTG UTMAREDG UBLKCP SYNCS.EXCH SYNCS.ARRIVE ELECT HMMA FFMA DFMA ATOMG RED; do
  printf "%-14s %6d\n" $op $(grep -c "[ .]$op" /tmp/fcd_sass.txt)
done
echo "# per kernel (tcgen05 / TMA users only): UTCHMMA LDTM UTMALDG UTMASTG+UTMAREDG"
awk '/Function :/ {name=$3} /UTCHMMA/ {m[name]++} /LDTM/ {l[name]++} /UTMALDG/ {t[name]++} /UTMASTG|UTMAREDG/ {s[name]++}
     END {for (n in m) printf "%6d %4d %4d %4d  %s\n", m[n], l[n], t[n], s[n], n}' /tmp/fcd_sass.txt | sort -k5 | c++filt | cut -c1-200
