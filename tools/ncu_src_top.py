#!/usr/bin/env python
"""Top source lines of an ncu report with SourceCounters (--import-source on): samples and instructions per line.
Usage: ncu_src_top.py report.ncu-rep [N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname = None; hdr = None; acc = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    if r[2] != "-": continue                      # SASS rows carry an address; keep the per-line aggregates
    try:
        smp = int(r[hdr.index("# Samples")]); ins = int(r[hdr.index("Instructions Executed")])
    except ValueError:
        continue
    if smp or ins: acc.append((smp, ins, fname, int(r[0]), r[1].strip()[:90]))
ts = sum(a[0] for a in acc); ti = sum(a[1] for a in acc)
print("total samples %d, warp instructions %d" % (ts, ti))
for smp, ins, f, ln, src in sorted(acc, reverse=True)[:N]:
    print("%5.1f%% smp %5.1f%% ins  %s:%d  %s" % (100.0 * smp / ts, 100.0 * ins / ti, f, ln, src))
