#!/bin/bash
# Round-2 third measurement pass: GPU tests (with the primitive-collider tests), config-5 labels at n = 256, steady-state ncu capture
# of a dexterous hand on its product path (fp64 build).  Results are copied from gpurun_out/ to profiles/ by hand.
O=gpurun_out
set -x
python -m pytest tests -m gpu -x -q > $O/r2c_tests.log 2>&1; tail -3 $O/r2c_tests.log
python tools/cfg5_labels.py 256 > $O/cfg5_labels.log 2>&1; tail -1 $O/cfg5_labels.log
MGS_STEADY_F64=1 python tools/profile_steady.py allegro 1024 1200 100 > $O/r2c_steady_allegro_plain.log 2>&1 && tail -1 $O/r2c_steady_allegro_plain.log && \
MGS_STEADY_F64=1 MGS_STEADY_REPS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:mgs_rollout --launch-skip 1 --launch-count 1 -f \
    -o $O/prof_r2c_allegro_f64_steady python tools/profile_steady.py allegro 1024 1200 100 > $O/r2c_steady_allegro_ncu.log 2>&1
ls -la $O/prof_r2c_allegro_f64_steady.ncu-rep
