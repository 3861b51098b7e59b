"""Opcode-level stall summary of one kernel from an ncu report: python tools/ncu_ops.py <report> <kernel-regex>"""
import collections
import csv
import subprocess
import sys

rep, k = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + k],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(r for r in rows if "Source" in r and "# Samples" in r)
isrc, isamp, iinst = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
data = []
for r in rows:
    if len(r) != len(hdr):
        continue
    try:
        data.append((r[isrc].strip(), float(r[isamp] or 0), float(r[iinst] or 0)))
    except ValueError:
        continue
tot = sum(d[1] for d in data) or 1.0
toti = sum(d[2] for d in data) or 1.0
agg, aggi = collections.Counter(), collections.Counter()
for src, smp, ins in data:
    t = src.split()
    if not t:
        continue
    op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
    op = op.split(".")[0]
    agg[op] += smp
    aggi[op] += ins
print("==", k, "samples", tot, "warp instr", toti, "sass rows", len(data))
for op, v in agg.most_common(14):
    print("  %-12s %5.1f%% smp  %5.1f%% ins" % (op, v / tot * 100, aggi[op] / toti * 100))
for src, smp, ins in sorted(data, key=lambda d: -d[1])[:8]:
    print("    %5.1f%%  %s" % (smp / tot * 100, src[:110]))
