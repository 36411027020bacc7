#!/bin/bash
# End-of-round-2 `ncu --set full` captures of the reworked conv kernels (one launch each); reports land in gpurun_out/.
tag=${1:-r2z}
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
for spec in "conv0 3" "prob 3" "conv1 3" "conv11 3"; do
  set -- $spec
  $NCU -k regex:conv3d_umma_kernel --launch-skip 3 -c 1 -o gpurun_out/${tag}_$1_s$2 \
      python tools/prof_conv.py --layers $1 --stages $2 --reps 1 > gpurun_out/${tag}_$1_s$2.log 2>&1
done
# a flat 2D extractor layer (8 -> 8, five 1600x1184 images): the second conv launch of a whole-model step
$NCU -k regex:conv3d_umma_kernel --launch-skip 1 -c 1 -o gpurun_out/${tag}_flat2d_conv01 \
    python tools/one_step.py > gpurun_out/${tag}_flat2d_conv01.log 2>&1
ls -la gpurun_out/*.ncu-rep
