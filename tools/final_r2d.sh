#!/bin/bash
# Round-2 fourth measurement pass (first-pass capacity selection: the Allegro hand's fp64 path on four warp-environments per SM):
# GPU tests, the Allegro bench line, its label agreement.
O=gpurun_out
set -x
python -m pytest tests -m gpu -x -q > $O/r2d_tests.log 2>&1; tail -3 $O/r2d_tests.log
python bench.py --workload allegro --no-also --steps 2 --warmup 3 > $O/bench_r2d_allegro.json 2> $O/bench_r2d_allegro.err; head -c 400 $O/bench_r2d_allegro.json; echo
MGS_LABELS_OUT=label_agreement_r2d_allegro.json python tools/label_agreement.py 512 allegro > $O/label_agreement_r2d_allegro.log 2>&1; tail -1 $O/label_agreement_r2d_allegro.log
