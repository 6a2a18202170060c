"""Which kernels differ between two builds of libpmw?  usage: python tools/sass_diff.py A.so B.so
Compares the SASS instruction streams function by function (addresses and encodings ignored).  Used to
show that a change which is meant to leave the production kernels alone really does: e.g. that the default
build is instruction-for-instruction the one the GPU suite last ran against."""
import re
import subprocess
import sys


def funcs(so):
    out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    d, cur = {}, None
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            d[cur] = []
            continue
        if cur:
            mm = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(.*?);", ln)
            if mm:
                d[cur].append(mm.group(1).strip())
    return d


a, b = funcs(sys.argv[1]), funcs(sys.argv[2])
only = sorted(set(a) ^ set(b))
diff = sorted(k for k in a if k in b and a[k] != b[k])
print(f"{len(a)} / {len(b)} kernels; in one build only: {len(only)}; different SASS: {len(diff)}")
for k in only + diff:
    print("  ", k[:110])
sys.exit(1 if (only or diff) else 0)
