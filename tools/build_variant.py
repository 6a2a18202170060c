"""Builds an alternative libpmw (development A/B runs): python tools/build_variant.py TAG -DMACRO=VALUE ...
-> pyminiweather_b200/variants/libpmw_TAG.so (select it with PMW_LIB=...)."""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyminiweather_b200 import _lib
tag, flags = sys.argv[1], sys.argv[2:]
out = os.path.join(ROOT, "pyminiweather_b200", "variants", f"libpmw_{tag}.so")
os.makedirs(os.path.dirname(out), exist_ok=True)
env = dict(os.environ); env.pop("CC", None); env.pop("CXX", None)
cmd = [_lib.nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-fmad=false", "-lineinfo", "-std=c++17",
       "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v", "-DPMW_DEV", "-o", out] + flags + [os.path.join(ROOT, "pyminiweather_b200", "csrc", "pmw_api.cu")]
res = subprocess.run(cmd, env=env, capture_output=True, text=True)
if res.returncode:
    print(res.stderr[-3000:]); sys.exit(1)
for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n.*?(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", res.stderr):
    n = m.group(1)
    if re.search(r"sweep_x\w*ILi2ELi1ELb0ELb0|sweep_z\w*ILi1ELb0ELb0", n):
        print(tag, n[:44], "stack", m.group(2), "spill", m.group(3), "regs", m.group(5))
print("built", out)
