#!/bin/bash
# One `ncu --set full` capture per hot kernel of the cfg3 step (B200_PROFILING.md recipe); reports land in gpurun_out/.
# Usage (on the GPU box): tools/ncu_capture.sh <tag>
tag=${1:-r1b}
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
for st in 1 2 3; do
  $NCU -k regex:warp_variance_c8 --launch-skip 3 -c 1 -o gpurun_out/${tag}_warp_c8h_s${st} \
      python tools/prof_warp.py --mode c8h --stages $st --reps 1 > gpurun_out/${tag}_warp_c8h_s${st}.log 2>&1
done
for spec in "conv0 1" "conv0 3" "prob 3" "conv11 3" "conv1 3"; do
  set -- $spec
  $NCU -k regex:conv3d_umma_kernel --launch-skip 3 -c 1 -o gpurun_out/${tag}_$1_s$2 \
      python tools/prof_conv.py --layers $1 --stages $2 --reps 1 > gpurun_out/${tag}_$1_s$2.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
