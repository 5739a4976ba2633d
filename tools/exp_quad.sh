# Tuning-build experiments for the quad kernel.  Build first:  make -C rte-ecckd_b200/csrc -B EXTRA=-DECCKD_TUNE
#   ECCKD_QUAD_VARIANT=<n>  alternative instantiation (register split, LUT width, ring sizes, staged refills ...: the table
#                           next to ECCKD_QUAD_VARIANT in csrc/ecckd_quad.cu).  NOTE: the run-time experiment switches cost the
#                           tuning build registers (it spills where the production build does not): use it to find out WHERE
#                           time goes; decide between variants with two PRODUCTION builds (a -D macro around the
#                           instantiation) copied to the box and timed by bench.py inside one gpurun call
#   ECCKD_CARVEOUT=<pct>    preferred shared-memory carve-out (100: smallest L1), ECCKD_VERBOSE=1 prints the shared-memory size
#   ECCKD_DBG=<bits>        timing experiments, results are WRONG: 1 no TMA stores, 2 index warps idle after the first ring
#                           fill, 4 no arithmetic, 8 no cache refills, 16 no staging stores / proxy fence, 64 no Planck warps,
#                           128 Planck warps only, 8192 refill loads issued with false predicates after the first fill
#   ECCKD_BLOCKS=<n>        blocks per g-group (default: one per SM)
# Compare only numbers from ONE run of this script: the B200 boxes of the pool differ by several per cent.
for v in ${VARIANTS:-0 1 2}; do for b in lw sw; do echo -n "VAR=$v "; ECCKD_QUAD_VARIANT=$v timeout -s KILL 60 python tools/prof_one.py $b f64 walk 65536 6; done; done
for v in ${VARIANTS32:-0}; do for b in lw sw; do echo -n "VAR=$v "; ECCKD_QUAD_VARIANT=$v timeout -s KILL 60 python tools/prof_one.py $b f32 walk 65536 6; done; done
for d in ${DBGS:-64 66 128}; do for b in lw; do echo -n "VAR=${DV:-0} DBG=$d "; ECCKD_QUAD_VARIANT=${DV:-0} ECCKD_DBG=$d timeout -s KILL 60 python tools/prof_one.py $b f64 walk 65536 6; done; done
