#!/bin/bash
# Launch list of one bench step under ncu (cold-cache, serialised: compare SHARES, not absolutes) -> gpurun_out/<tag>_launches.csv
tag=${1:-r2}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv \
    python tools/one_step.py > gpurun_out/${tag}_launches.log 2>&1
