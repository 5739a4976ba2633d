#!/bin/bash
# sweep the ECCKD_TUNE walk-kernel variants (library must be built with EXTRA=-DECCKD_TUNE)
for v in ${VARIANTS:-0 1 2 3 4 5 6 7}; do
  for cfg in "lw f64" "sw f64" "lw f32"; do
    echo -n "variant $v: "; ECCKD_WALK_VARIANT=$v timeout 120 python tools/prof_one.py $cfg walk ${NCOL:-65536} 4 | sort -t: -k2 -n | head -1
  done
done
