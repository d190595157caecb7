#!/usr/bin/env python
"""stdin: `ncu -i rep --page source --csv --print-source cuda,sass`; stdout: per kernel result the executed warp
instructions by source line (top N) and by SASS opcode.  Usage: ... | ncu_src_agg.py [N] [case names...]"""
import collections, csv, sys
N = int(sys.argv[1]) if len(sys.argv) > 1 else 40
names = sys.argv[2:]
res = []          # one entry per kernel result
cur = None; fname = None
for row in csv.reader(sys.stdin):
    if not row:
        continue
    if row[0] == "File Path":
        fname = row[1].split("/")[-1]; continue
    if row[0] == "Function Name":
        if cur is None or cur["fn"] != row[1] or cur.get("closed"):
            cur = {"fn": row[1], "lines": collections.Counter(), "samp": collections.Counter(), "ops": collections.Counter(),
                   "src": {}, "seen": set()}
            res.append(cur)
        continue
    if row[0] == "Line No":
        hdr = row; continue
    if cur is None or len(row) < 9:
        continue
    try:
        ins = int(row[7]); smp = int(row[6])
    except ValueError:
        continue
    if row[0].isdigit():
        k = (fname, int(row[0]))
        cur["lines"][k] += ins; cur["samp"][k] += smp; cur["src"][k] = row[1].strip()[:100]
    elif row[2].startswith("0x"):
        if (fname, row[2]) in cur["seen"]:
            continue
        op = row[3].split()
        if op and op[0].startswith("@"):
            op = op[1:]
        if op:
            cur["ops"][op[0].split(".")[0]] += ins
# ncu prints one block per (result, file): merge consecutive blocks of the same function until the SASS repeats
for i, r in enumerate(res):
    if sum(r["lines"].values()) < 200000:
        continue
    tot = sum(r["ops"].values()) or 1
    tl = sum(r["lines"].values()) or 1
    ts = sum(r["samp"].values()) or 1
    print("=" * 100)
    print("result %d %s  %s   warp-inst(by sass) %d" % (i, names[i // 2] if i // 2 < len(names) else "", r["fn"][:70], tot))
    print("  opcodes: " + "  ".join("%s %.1f%%" % (o, 100.0 * c / tot) for o, c in r["ops"].most_common(28)))
    for k, c in r["lines"].most_common(N):
        print("  %9d %5.1f%%  smp %5.1f%%  %s:%d  %s" % (c, 100.0 * c / tl, 100.0 * r["samp"][k] / ts, k[0], k[1], r["src"][k]))
