#!/bin/bash
# fp32 label agreement of ablation builds (build_ab/lib_<variant>.so) on the marginal candidate sets
# usage (GPU box): tools/ablation_labels.sh "variant ..." "gripper,gripper"
for v in $1; do
  echo "== $v"
  MGS_B200_SO=$PWD/build_ab/lib_$v.so MGS_LABELS_SKIP_F64=1 MGS_LABELS_OUT=abl_$v.json python tools/label_agreement.py 512 $2 2>&1 | tail -1
done
