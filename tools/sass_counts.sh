#!/bin/bash
# SASS evidence for the shipped library (B200_PROFILING.md "What proves a Blackwell-native kernel"):
#   bash tools/sass_counts.sh > profiles/r02_sass_counts.txt
LIB=diff-mining_b200/libdm_b200.so
echo "# cuobjdump -sass $LIB | grep -c <mnemonic>      ($(date -u +%Y-%m-%dT%H:%MZ), $(git rev-parse --short HEAD 2>/dev/null))"
echo "# library: $(stat -c %s $LIB) bytes; ldd shows no cuBLAS / cuDNN:"
ldd $LIB | grep -E "cublas|cudnn|cutlass" || echo "#   (none)"
SASS=$(mktemp)
cuobjdump -sass $LIB > $SASS
for m in UTCHMMA "UTCHMMA.2CTA" LDTM STTM UTMALDG UTMASTG UBLKCP "MUFU.EX2"; do
  printf "%-28s %s\n" "$m" "$(grep -c -F "$m" $SASS)"
done
printf "%-28s %s\n" "HMMA (legacy mma.sync)" "$(grep -c -E '(^|[^C])HMMA' $SASS)"
printf "%-28s %s\n" "HGMMA / QGMMA (wgmma)" "$(grep -c -E 'HGMMA|QGMMA|IGMMA' $SASS)"
echo "# per kernel (UTCHMMA / LDTM / UTMALDG / MUFU.EX2):"
awk '/Function :/ {name=$3} /UTCHMMA/ {a[name]++} /LDTM/ {b[name]++} /UTMALDG/ {c[name]++} /MUFU.EX2/ {d[name]++} END {for (n in a) printf "%6d %6d %6d %6d  %s\n", a[n], b[n], c[n], d[n], n}' $SASS | sort -k5 | c++filt 2>/dev/null | cut -c1-200
rm -f $SASS
