"""scratch: registers / spills per trace_kernel instantiation (ptxas -v of trace.cu)"""
import re, subprocess, sys, os
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "raytracercuda_b200", "csrc", "trace.cu")
extra = sys.argv[2:]
out = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false", "-Xptxas", "-v", "-c", "-o", "/tmp/regs.o", src] + extra, capture_output=True, text=True).stderr
cur = None
for l in out.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", l)
    if m:
        name = m.group(1); k = re.search(r"(\w+_kernel)I((?:L[ib][01]E)+)", name)
        cur = (k.group(1) + "<" + ",".join(re.findall(r"L[ib]([01])E", k.group(2))) + ">") if k else re.sub(r"^_ZN\w*?(\d+)(\w+_kernel).*", r"\2", name)[:50]
    m = re.search(r"Used (\d+) registers", l)
    if m and cur: regs = m.group(1)
    m2 = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", l)
    if m2 and cur: stack = m2.groups()
    if "Used" in l and cur:
        print("%-40s regs %s stack/spillst/spillld %s" % (cur, regs, stack)); cur = None
