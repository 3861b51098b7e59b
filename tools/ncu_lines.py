"""Per-source-line stall samples / instruction counts of one kernel in an ncu report:
  python tools/ncu_lines.py <report.ncu-rep> <kernel-regex> [top]"""
import collections, csv, subprocess, sys
rep, k = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + k, "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = hdr = None
agg, samp, src = collections.Counter(), collections.Counter(), {}
for r in rows:
    if r and r[0] == "File Path": cur = r[1].split("/")[-1]; hdr = None; continue
    if r and r[0] == "Function Name": continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    try: ln = int(r[0])
    except ValueError: continue
    try: n = float(r[hdr.index("Instructions Executed")] or 0); s = float(r[hdr.index("# Samples")] or 0)
    except ValueError: continue
    agg[(cur, ln)] += n; samp[(cur, ln)] += s; src[(cur, ln)] = r[1]
tot, ts = sum(agg.values()) or 1, sum(samp.values()) or 1
print("total warp inst %d  samples %d" % (tot, ts))
for key, v in samp.most_common(top):
    print("%5.1f%% smp %5.1f%% inst  %s:%d  %s" % (v / ts * 100, agg[key] / tot * 100, key[0], key[1], src[key].strip()[:105]))
