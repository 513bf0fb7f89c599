#!/usr/bin/env python3
"""SASS digest of the built library: per kernel the instruction count, registers / spill bytes, and the counts of the
instructions that show what the code is made of -- 256-bit read-only loads (LDG.E.ENL2.256.CONSTANT, Blackwell's
ld.global.nc.v8.f32), 64-bit shared-memory stack accesses, warp votes, PRMT plane decodes, FFMA.

    python tools/sass_digest.py > profiles/r02_sass_digest.txt
"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "luxcore_b200", "lib", "libluxrays_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
usage = {}
cur = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        cur = m.group(1)
    elif cur and "REG:" in line:
        usage[cur] = line.strip()
        cur = None
kernels, name = {}, None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        kernels[name] = []
    elif name and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
        kernels[name].append(re.sub(r"/\*.*?\*/", "", line).strip())
demangle = subprocess.run(["c++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines()
print("SASS digest of luxcore_b200/lib/libluxrays_b200.so (sm_100a; cuobjdump -sass / -res-usage)\n")
print("%-86s %6s %4s %6s | %8s %7s %7s %6s %6s %6s" % ("kernel", "instr", "regs", "spillB", "LDG.256", "STS.64", "LDS.64", "VOTE", "PRMT", "FFMA"))
for (mangled, ins), pretty in sorted(zip(kernels.items(), demangle), key=lambda kv: kv[1]):
    if "cub" in pretty:
        continue
    u = usage.get(mangled, "")
    regs = re.search(r"REG:(\d+)", u); stack = re.search(r"STACK:(\d+)", u)
    cnt = lambda pat: sum(1 for i in ins if re.search(pat, i))
    print("%-86s %6d %4s %6s | %8d %7d %7d %6d %6d %6d" % (pretty.replace("lrb::", "")[:86], len(ins), regs.group(1) if regs else "?", stack.group(1) if stack else "?",
          cnt(r"LDG\.E\.ENL2\.256"), cnt(r"STS\.64"), cnt(r"LDS\.64"), cnt(r"\bVOTEU?\."), cnt(r"\bPRMT\b"), cnt(r"\bFFMA\b")))
print("\nSTACK = bytes of local memory per thread (register spills in cold paths and the 16-float matrix of the motion sampler).")
