#!/bin/bash
for e in "FAA_X=0" "FAA_CHAIN=1" "FAA_X=0" "FAA_CHAIN=1" "FAA_CHAIN=2" "FAA_CHAIN=1 FAA_MID_BANDS=4 FAA_MID_THREADS=256"; do
  env $e python tools/mix_probe.py
done
