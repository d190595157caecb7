#!/bin/bash
# round-2 evidence: launch list of the bench command, full ncu capture of the pixel kernels of one mix step, Mixup kernels
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 10 --warmup 3 --no-cpu --no-also > gpurun_out/r02_launches_bench.log 2>&1
MIX_N=4 ncu --set full --import-source on --clock-control none -k regex:faa_ -s 15 -c 6 -f -o gpurun_out/prof_r02 \
    python tools/mix_probe.py > gpurun_out/r02_prof.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:mix -s 2 -c 3 -f -o gpurun_out/prof_r02_mixup \
    python tools/mixup_ncu_target.py > gpurun_out/r02_prof_mixup.log 2>&1
ls -la gpurun_out/prof_r02.ncu-rep gpurun_out/prof_r02_mixup.ncu-rep
