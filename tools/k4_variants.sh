#!/bin/bash
# K4 tuning run: tools/bench_extra.py --section loss against libraries built with -DMSPL_UWCE_MINB / -DMSPL_UWCE_BWD_P
# (make -C mspl_b200/csrc OUT=<dir> EXTRA="-D..."), one JSON-lines block per variant.
out=${1:-gpurun_out/k4_variants.txt}
: > "$out"
for lib in mspl_b200/lib/libmspl_b200.so tools/_k4var/*/libmspl_b200.so; do
  echo "# $lib" >> "$out"
  MSPL_B200_LIB=$PWD/$lib python tools/bench_extra.py --section loss 2>/dev/null | grep -E "b64|b256" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('%-42s %.4f ms  %.3f' % (d['name'], d['ms_median'], d['frac_of_measured_hbm_peak']))" >> "$out"
done
cat "$out"
