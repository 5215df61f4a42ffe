#!/bin/bash
# SASS evidence that the hot kernels are Blackwell-native: per-kernel counts of the tcgen05 / TMEM / TMA mnemonics in the in-tree
# libsat_b200.so (B200_PROFILING.md: UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = TMA tensor load, UBLKCP = 1-D bulk copy,
# UTCBAR = tcgen05.commit, SYNCS = mbarrier ops).      tools/sass_summary.sh > profiles/r02_sass_summary.txt
SO="$(dirname "$0")/../show-attend-and-tell-pytorch-lightning_b200/libsat_b200.so"
echo "cuobjdump -sass $(basename "$SO") (sm_100a): mnemonic counts per kernel (kernels without any of them are omitted)"
/usr/local/cuda/bin/cuobjdump -sass "$SO" | awk '
  /Function :/ { name=$3; next }
  /UTCHMMA/ { a[name]++ } /LDTM/ { b[name]++ } /UTMALDG/ { c[name]++ } /UBLKCP/ { d[name]++ } /UTCBAR/ { e[name]++ } /SYNCS/ { f[name]++ }
  END { for (n in f) if (a[n]+b[n]+c[n]+d[n]+e[n] > 0) printf "%6d %6d %7d %6d %6d %6d  %s\n", a[n], b[n], c[n], d[n], e[n], f[n], n }' \
  | sort -k7 | (echo "UTCHMMA   LDTM UTMALDG UBLKCP UTCBAR  SYNCS  kernel (mangled)"; cat) | cut -c1-220
echo
echo "totals:"
/usr/local/cuda/bin/cuobjdump -sass "$SO" | grep -o "UTCHMMA\|LDTM\|UTMALDG\|UBLKCP\|UTCBAR\|SYNCS\|UTMAPF\|HMMA\|MUFU.TANH\|MUFU.EX2" | sort | uniq -c
