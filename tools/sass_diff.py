"""Which kernels of an object file changed?  Compares the SASS of every function of two .o files (addresses stripped):
  python tools/sass_diff.py old.o new.o
Used to show that adding an opt-in template instantiation leaves the default kernels byte-identical to the build the GPU
tests ran on (e.g. old.o = the same source at the last GPU-verified commit, compiled with the Makefile's flags)."""
import hashlib
import re
import subprocess
import sys


def funcs(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    res, cur = {}, None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            res[cur] = []
        elif cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            res[cur].append(re.sub(r"/\*[0-9a-f]*\*/", "", line))
    return {k: hashlib.md5("\n".join(v).encode()).hexdigest() for k, v in res.items()}


if __name__ == "__main__":
    a, b = funcs(sys.argv[1]), funcs(sys.argv[2])
    print("functions: %d -> %d" % (len(a), len(b)))
    for k in sorted(a):
        if k not in b:
            print("missing  ", k)
        elif a[k] != b[k]:
            print("CHANGED  ", k)
    for k in sorted(b):
        if k not in a:
            print("new      ", k)
