#!/bin/bash
# Sweeps the step-chunking heuristic of the UMMA conv (MVS_UMMA_WAVES / MVS_UMMA_MIN_STEPS) over all cfg3 layers.
mkdir -p gpurun_out
for w in 0 2 3 4 6 8 12 24; do
  for ms in 0; do
    MVS_UMMA_WAVES=$w MVS_UMMA_MIN_STEPS=$ms python tools/prof_conv.py --reps 3 > gpurun_out/sweep_w${w}_m${ms}.jsonl 2>&1
  done
done
python - <<'PY'
import json,glob,collections
tab=collections.defaultdict(dict)
for f in sorted(glob.glob('gpurun_out/sweep_w*_m*.jsonl')):
    key=f.split('sweep_')[1].split('.jsonl')[0]
    for l in open(f):
        try: r=json.loads(l)
        except Exception: continue
        tab[(r['stage'],r['layer'])][key]=r['ms']
keys=sorted({k for v in tab.values() for k in v}, key=lambda s:int(s.split('_')[0][1:]))
print('layer', *keys)
for k,v in tab.items():
    print(k, *[v.get(x) for x in keys])
PY
