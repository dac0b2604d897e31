#!/bin/bash
# Round-2 final measurements on one B200 (results are copied from gpurun_out/ to profiles/ by hand)
O=gpurun_out
set -x
python bench.py --steps 20 --warmup 5 > $O/bench_r2_ours.json 2> $O/bench_r2_ours.err
python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_r2_reference.json 2> $O/bench_r2_reference.err
MGS_LABELS_OUT=label_agreement_r2.json python tools/label_agreement.py 512 > $O/label_agreement_r2.log 2>&1
python bench.py --workload mixed --steps 1 --warmup 3 > $O/bench_r2_mixed.json 2> $O/bench_r2_mixed.err
python bench.py --workload clutter_shadow --steps 2 --warmup 3 > $O/bench_r2_clutter_shadow.json 2> $O/bench_r2_clutter_shadow.err
python tools/first50.py 16 > $O/first50_r2.log 2>&1; cp $O/first50.json $O/first50_r2.json
python tools/profile_clutter.py 148 300 20 80 > $O/r2_final_wide_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mgs_rollout_kernel_wide --launch-skip 181 --launch-count 1 -f -o $O/prof_r2_wide_clutter_final python tools/profile_clutter.py 148 300 20 80 > $O/r2_final_wide_ncu.log 2>&1
tail -2 $O/r2_final_wide_plain.log
