#!/bin/bash
# Round-2 ncu captures (GPU box, one GPU).  Every command runs once WITHOUT ncu first; numbers printed under ncu are not bench values.
#   steady hold-phase launches (per-function / stall breakdown, warp-instructions per env-step) of the two bench grippers,
#   DRAM traffic of one full bench launch per workload, and the launch list of a short bench run.
set -x
O=gpurun_out
for g in robotiq2f85:1480 panda:2368; do
  name=${g%%:*}; n=${g##*:}
  python tools/profile_steady.py $name $n 1200 100 > $O/r2_steady_${name}_plain.log 2>&1 || exit 1
  MGS_STEADY_REPS=1 ncu --set full --clock-control none --import-source on -k regex:mgs_rollout --launch-skip 1 --launch-count 1 -f \
      -o $O/prof_r2_${name}_steady python tools/profile_steady.py $name $n 1200 100 > $O/r2_steady_${name}_ncu.log 2>&1
done
for w in robotiq panda; do
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:mgs_rollout --launch-skip 3 --launch-count 1 \
      --csv --log-file $O/traffic_r2_$w.csv python bench.py --workload $w --no-also --no-cpu --steps 1 --warmup 3 > $O/r2_traffic_$w.log 2>&1
done
python bench.py --steps 2 --warmup 1 --no-cpu --no-hands > $O/r2_launchlist_plain.json 2> $O/r2_launchlist_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_r2_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-hands > $O/r2_launchlist_ncu.log 2>&1
ls -la $O/prof_r2_*steady* $O/traffic_r2_* $O/launches_r2_bench.csv
