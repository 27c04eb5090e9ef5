#!/bin/bash
# SASS mnemonic counts of the shipped library (tcgen05 / TMA / TMEM / cp.async evidence), per kernel family.
so=${1:-safer2-recommender_b200/libfrecsys_b200.so}
echo "# cuobjdump -sass $so  (sm_100a), $(date -u +%F)"
cuobjdump -sass "$so" > /tmp/frx_all.sass
echo "## whole library"
grep -oE "\b(UTCHMMA|UTCQMMA|UTMALDG[A-Z0-9.]*|UTMASTG[A-Z0-9.]*|LDTM[A-Z0-9.]*|STTM[A-Z0-9.]*|UTCBAR|UTCCP|LDGSTS[A-Z0-9.]*|FFMA2|HMMA[A-Z0-9.]*|SYNCS[A-Z0-9.]*|UBLKCP[A-Z0-9.]*|F2FP[A-Z0-9.]*)\b" /tmp/frx_all.sass | sed 's/\..*//' | sort | uniq -c | sort -rn
echo "## per kernel: UTCHMMA / LDTM / STTM / LDGSTS / UTMALDG / FFMA2"
awk '/Function :/ {name=$3} /UTCHMMA/ {a[name]++} /LDTM/ {b[name]++} /STTM/ {c[name]++} /LDGSTS/ {d[name]++} /UTMALDG/ {e[name]++} /FFMA2/ {f[name]++} END {for (n in a) printf "%6d %5d %5d %5d %5d %6d  %s\n", a[n], b[n], c[n], d[n], e[n], f[n], n}' /tmp/frx_all.sass | sort -k7 | c++filt | cut -c1-200
