"""dev/sass_census.py -- per-kernel SASS census of libxyz_b200.so (the reference's `task ptx` idea, Taskfile.yaml:27-146,
applied to sm_100a): registers / spills / shared memory from `cuobjdump -res-usage`, and counts of the instructions that
prove what each kernel is built from (`cuobjdump -sass`):
  UBLKCP   TMA 1-D bulk copies (cp.async.bulk)          SYNCS    mbarrier operations
  LDGSTS   cp.async                                     FFMA2/FADD2/FMUL2  packed fp32 (Blackwell)
  MUFU     special-function unit (ex2, rcp, ...)        REDG/RED/ATOMG      global reductions / atomics
  MATCH    match.any (warp-aggregated add_grad)         SHFL     warp shuffles        DFMA/DADD/DMUL  fp64
Usage: python dev/sass_census.py [path/to/lib.so] > profiles/sass_census_rNN.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "xyz-autodiff-cuda_b200", "lib", "libxyz_b200.so")
MNEMONICS = ["UBLKCP", "SYNCS", "LDGSTS", "FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "FMUL", "MUFU", "REDG", "RED", "ATOMG",
             "ATOMS", "MATCH", "SHFL", "DFMA", "DADD", "DMUL", "LDS", "STS", "LDG", "STG", "BAR", "MEMBAR", "LOP3"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), stdout=subprocess.PIPE, text=True).stdout.splitlines()
    res = {}
    for m, d in zip(names, out):
        d = d.replace("xyzb::(anonymous namespace)::", "").replace("void ", "")
        tu = re.search(r"_GLOBAL__N__[0-9a-f]+_\d+_(\w+?)_cu", m)
        res[m] = re.sub(r"\(.*", "", d) + (f" [{tu.group(1)}.cu]" if tu else "")
    return res


res = subprocess.run(["cuobjdump", "-res-usage", lib], stdout=subprocess.PIPE, text=True).stdout
usage = {}
cur = None
for line in res.splitlines():
    m = re.match(r"\s*Function (\S+):", line)
    if m:
        cur = m.group(1)
        continue
    m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
    if m and cur:
        usage[cur] = tuple(int(v) for v in m.groups())
sass = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True).stdout
counts = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        counts[cur]["_total"] += 1
        for mn in MNEMONICS:
            if op == mn:
                counts[cur][mn] += 1
names = demangle(list(counts))
print(f"# SASS census of {os.path.relpath(lib, ROOT)} (sm_100a), one line per kernel; zero counts omitted")
print("# kernel | regs stack(spill bytes) static-smem | instructions | mnemonic counts")
for k in sorted(counts, key=lambda k: names[k]):
    r = usage.get(k, (0, 0, 0, 0))
    c = counts[k]
    body = " ".join(f"{mn}={c[mn]}" for mn in MNEMONICS if c[mn])
    print(f"{names[k]} | regs={r[0]} stack={r[1]} smem={r[2]} | {c['_total']} | {body}")
