"""SASS evidence per kernel of libsqlx.so: counts of the Blackwell tensor / TMA / TMEM mnemonics, packed-fp32 math, waterfall
loops around uniform-datapath instructions, registers and spills.   python tools/sass_summary.py [lib] > profiles/...txt"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "sfmnext-impl_b200/lib/libsqlx.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
regs = {}
cur = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        cur = m.group(1)
        continue
    m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", line)
    if m and cur:
        regs[cur] = tuple(int(g) for g in m.groups())
keys = ["UTCHMMA", "UTMALDG", "LDTM", "STTM", "UTCBAR", "SYNCS", "MUFU.EX2", "FFMA2", "BRA.U.ANY"]
counts = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    counts[cur]["inst"] += 1
    for k in keys:
        if op.startswith(k):
            counts[cur][k] += 1
dem = subprocess.run(["cu++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines()
print("%-78s %6s %5s %5s | %s" % ("kernel (sm_100a SASS of %s)" % lib.split("/")[-1], "inst", "regs", "stack", "  ".join("%s" % k for k in keys)))
for (mangled, c), name in zip(counts.items(), dem):
    cut = name.find(">(")
    name = (name[:cut + 1] if cut >= 0 else name.split("(")[0]).replace("void ", "").replace("sqlx::", "")
    r = regs.get(mangled, (0, 0, 0))
    print("%-78s %6d %5d %5d | %s" % (name[:78], c["inst"], r[0], r[1], "  ".join("%*d" % (len(k), c[k]) for k in keys)))
