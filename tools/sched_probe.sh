#!/bin/bash
for e in "FAA_X=0" "FAA_X=0" "FAA_MID=0" "FAA_CHAIN=1"; do
  env $e python tools/mix_probe.py
done
