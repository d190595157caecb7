#!/bin/bash
for e in "FAA_X=0" "FAA_MID_BANDS=4 FAA_MID_THREADS=256" "FAA_X=0" "FAA_MID_BANDS=4 FAA_MID_THREADS=256" "FAA_MID=0" "FAA_CHAIN=1" "FAA_OCTETS=0"; do
  env $e python tools/mix_probe.py
done
