# tuning-build experiments for the quad kernel (make -C rte-ecckd_b200/csrc -B EXTRA=-DECCKD_TUNE)
for v in ${VARIANTS:-0 1 2 4 6 7 8 9 10}; do for b in lw sw; do echo -n "VAR=$v "; ECCKD_QUAD_VARIANT=$v timeout -s KILL 60 python tools/prof_one.py $b f64 walk 65536 6; done; done
for v in ${VARIANTS32:-0 4}; do for b in lw sw; do echo -n "VAR=$v "; ECCKD_QUAD_VARIANT=$v timeout -s KILL 60 python tools/prof_one.py $b f32 walk 65536 6; done; done
for d in ${DBGS:-64 66 128}; do for b in lw; do echo -n "VAR=${DV:-6} DBG=$d "; ECCKD_QUAD_VARIANT=${DV:-6} ECCKD_DBG=$d timeout -s KILL 60 python tools/prof_one.py $b f64 walk 65536 6; done; done
