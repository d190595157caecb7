#!/bin/bash
# per-family ncu profile of a few cases: one report, summarised on the box (the .ncu-rep is too big to bring back)
# usage: ncu_classes.sh TAG case [case...]
set -x
TAG=$1; shift
mkdir -p gpurun_out/cls /tmp/rep
export FAA_CHAIN=${FAA_CHAIN:-0}
ncu --set full --import-source on --clock-control none --nvtx --nvtx-include "prof/" -k regex:faa_augment -f -o /tmp/rep/cls \
    python tools/ncu_classes.py "$@" > gpurun_out/cls/${TAG}_run.log 2>&1
ncu -i /tmp/rep/cls.ncu-rep --page raw --csv > gpurun_out/cls/${TAG}_raw.csv
timeout 600 ncu -i /tmp/rep/cls.ncu-rep --page source --csv --print-source cuda,sass > /tmp/rep/src.csv
python tools/ncu_src_agg.py 60 < /tmp/rep/src.csv > gpurun_out/cls/${TAG}_src_agg.txt
ls -la /tmp/rep gpurun_out/cls
