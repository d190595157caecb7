#!/bin/bash
# the bench mix under the scheduling switches (each line: env -> us/step)
for e in "FAA_CHAIN=1" "FAA_CHAIN=2" "FAA_CHAIN=0" "FAA_CHAIN=1 FAA_MID=0" "FAA_CHAIN=0 FAA_MID=0" "FAA_CHAIN=1 FAA_MID_BANDS=8" "FAA_CHAIN=1 FAA_MID_BANDS=2"; do
  env $e python tools/mix_probe.py
done
