#!/bin/bash
# Sweep the ECCKD_TUNE walk-kernel variants (build first: make -C rte-ecckd_b200/csrc -B EXTRA=-DECCKD_TUNE).
# Variant numbers are listed next to ECCKD_VARIANT in csrc/ecckd_kernels.cu.  Other tuning-build switches:
#   ECCKD_DBG=<bits>   1 no writer stores, 2 index warps idle after the first chunk, 8 no cache refills,
#                      64 no Planck warps, 128 Planck warps only   (timing experiments: results are wrong)
#   ECCKD_CH=1|2, ECCKD_BLOCKS=<n>   force tiles per chunk / blocks per g-group (results stay correct)
#   ECCKD_NO_TMA=1     writer-warp path (also honoured by the production library)
for v in ${VARIANTS:-0 106 107 114 118 122 123 1 2 3 4 5 6 7 8}; do
  for cfg in ${CONFIGS:-"lw f64" "sw f64"}; do
    echo -n "variant $v: "; ECCKD_WALK_VARIANT=$v timeout 120 python tools/prof_one.py $cfg walk ${NCOL:-65536} ${REPS:-10}
  done
done
