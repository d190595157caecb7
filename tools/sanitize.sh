#!/bin/bash
# compute-sanitizer logs for profiles/ (memcheck, racecheck, synccheck); usage: sanitize.sh TAG
TAG=${1:-r02}
mkdir -p gpurun_out/san
for tool in ${TOOLS:-memcheck racecheck}; do      # synccheck: see profiles/r02_sanitizer_notes.txt
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_target.py > gpurun_out/san/${TAG}_$tool.txt 2>&1
  tail -3 gpurun_out/san/${TAG}_$tool.txt
done
