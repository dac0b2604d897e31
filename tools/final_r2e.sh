#!/bin/bash
# Round-2 fifth measurement pass (LEAP fp64 on four warp-environments per SM: 12 contacts, 32-pair shared MPR cache): GPU tests, LEAP bench line.
O=gpurun_out
set -x
python -m pytest tests -m gpu -q > $O/r2e_tests.log 2>&1; tail -5 $O/r2e_tests.log
python bench.py --workload leap --no-also --steps 2 --warmup 3 > $O/bench_r2e_leap.json 2> $O/bench_r2e_leap.err; head -c 300 $O/bench_r2e_leap.json; echo; grep -o '"capacity": {[^}]*}' $O/bench_r2e_leap.json; grep -o '"envs_per_sm": [0-9]*' $O/bench_r2e_leap.json
