for r in 1 2; do for d in 0 1024; do for b in lw; do echo -n "DBG=$d "; ECCKD_DBG=$d timeout -s KILL 60 python tools/prof_one.py $b f64 walk 65536 8; done; done; done
for d in 128 1152; do echo -n "planck-only DBG=$d "; ECCKD_DBG=$d timeout -s KILL 60 python tools/prof_one.py lw f64 walk 65536 8; done
ECCKD_QUAD_VARIANT=2 ECCKD_DBG=0 timeout -s KILL 60 python tools/prof_one.py lw f64 walk 65536 8
ECCKD_QUAD_VARIANT=2 ECCKD_DBG=1024 timeout -s KILL 60 python tools/prof_one.py lw f64 walk 65536 8
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -q -x -k "gcm_block or golden or quad_kernel_equals" 2>&1 | tail -2
