#!/bin/bash
# Round-2 second measurement pass (after the closed-form primitive colliders): GPU tests, label agreement of the two hands whose
# oracle labels changed, short bench lines.  Results are copied from gpurun_out/ to profiles/ by hand.
O=gpurun_out
set -x
python -m pytest tests -m gpu -x -q > $O/r2b_tests.log 2>&1; tail -3 $O/r2b_tests.log
MGS_LABELS_OUT=label_agreement_r2b_hands.json python tools/label_agreement.py 512 allegro,shadow > $O/label_agreement_r2b_hands.log 2>&1; tail -1 $O/label_agreement_r2b_hands.log
python bench.py --workload clutter_shadow --steps 2 --warmup 3 > $O/bench_r2b_clutter_shadow.json 2> $O/bench_r2b_clutter_shadow.err; tail -c 600 $O/bench_r2b_clutter_shadow.json
python bench.py --steps 4 --warmup 3 > $O/bench_r2b_ours.json 2> $O/bench_r2b_ours.err; head -c 300 $O/bench_r2b_ours.json
