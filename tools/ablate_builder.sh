#!/bin/bash
# Bound analysis of the fused warp + variance builder: the shipped L1-gather kernel against three ablation
# builds of the same kernel (python -m mvs_b200.csrc.build --variant ablK MVS_C8_ABLATE=K), cfg3 stage shapes, CUDA events,
# L2 flushed.  1 = no tap arithmetic, 2 = no gathers, 3 = gathers + stores only.  Output: gpurun_out/<tag>_ablate.txt
tag=${1:-r2}
out=gpurun_out/${tag}_ablate.txt
mkdir -p gpurun_out
: > $out
for v in "" abl1 abl2 abl3; do
  lib=mvs_b200/libmvs_b200${v:+_$v}.so
  echo "== ${v:-shipped} ($lib)" >> $out
  MVS_B200_LIB=$PWD/$lib python tools/prof_warp.py --mode c8g --reps 10 2>&1 | grep '"cfg"' | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print('   stage', r['stage'], 'ms', r['ms_median'], 'GB/s', r['GBps'], 'frac', r['frac_of_measured_peak'])" >> $out
done
cat $out
