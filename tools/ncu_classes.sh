#!/bin/bash
# per-family ncu profile: one report for all cases, summarised on the box (the .ncu-rep is too big to bring back)
set -x
mkdir -p gpurun_out/cls /tmp/rep
TAG=${1:-r02}
ncu --set full --import-source on --clock-control none --nvtx --nvtx-include "prof/" -k regex:faa_augment -f -o /tmp/rep/cls \
    python tools/ncu_classes.py > gpurun_out/cls/${TAG}_run.log 2>&1
NAMES=$(grep '^CASE' gpurun_out/cls/${TAG}_run.log | awk '{print $2}' | tr '\n' ' ')
ncu -i /tmp/rep/cls.ncu-rep --page raw --csv > gpurun_out/cls/${TAG}_raw.csv
ncu -i /tmp/rep/cls.ncu-rep --page source --csv --print-source cuda,sass | python tools/ncu_src_agg.py 45 $NAMES > gpurun_out/cls/${TAG}_src_agg.txt
ls -la /tmp/rep gpurun_out/cls
