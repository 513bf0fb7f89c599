"""Static instruction counts of the sections of the persistent trace kernel, from the built library.

tools/warp_model.py weights the phases a warp issues with the number of SASS instructions of the
corresponding code.  This tool derives those numbers from libluxrays_b200.so instead of from a hand
count: it extracts the sm_100a cubin (cuobjdump -xelf), disassembles it with line information
(nvdisasm -g; the library is built with -lineinfo) and attributes every instruction of
TracePersistent<false, true, false, false> to the source function its line belongs to.

    python tools/sass_costs.py            # prints the table and the COST dict for warp_model.py

Static counts: a section with a rarely taken side path (the spilling push path of NodeStep) counts it
separately; helpers shared by two phases (Ld256, ChildEntry) go to the phase of the nearest
unambiguous instruction.
"""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "luxcore_b200", "csrc")
LIB = os.path.join(ROOT, "luxcore_b200", "lib", "libluxrays_b200.so")
KERNEL = "_ZN3lrb15TracePersistentILb0ELb1ELb0ELb0EEEvNS_9TraceArgsE"

# function -> section
SECTION = {
    "NodeStep": "node_phase", "SlotEntry": "node_phase", "QByte": "node_phase",
    "pushFast": "node_phase", "pushFastIf": "node_phase", "room": "node_phase", "push": "node_slow_push",
    "TriStep": "tri_phase", "TriangleTest": "tri_phase", "Dot3": "tri_phase",
    "Resolve": "pop_trip", "PopOnce": "pop_trip", "pop": "pop_trip", "empty": "pop_trip", "NeedsResolve": "inner_fixed",
    "WorkOf": "inner_fixed", "VoteTrianglePhase": "inner_fixed", "VoteEnterInstances": "inner_fixed",
    "InitRay": "refill", "SetRay": "refill", "LoadRay": "refill", "reset": "refill", "init": "outer_fixed",
    "StoreHit": "store", "StoreHitTo": "store", "WriteHit": "store", "ForwardMaskedHit": "store",
    "Ld256": None, "ChildEntry": None,      # shared: resolved by neighbourhood
}


def function_ranges(path):
    """[(first_line, name)] of the functions / methods defined in a source file (a line that opens a body)."""
    out = []
    pat = re.compile(r"^\s*(?:template\s*<[^>]*>\s*)?(?:LRB_HD|__device__|__global__|static|inline)[^;{]*?\b([A-Za-z_]\w*)\s*\([^;]*$")
    with open(path) as f:
        for i, line in enumerate(f, 1):
            m = pat.match(line)
            if m and not line.strip().startswith("//"):
                name = m.group(1)
                k = re.search(r"\b(TracePersistent|TraceStatic|RayKeyKernel)\s*\(", line)     # __global__ ... __launch_bounds__(...) Name(
                out.append((i, k.group(1) if k else name))
    return out


def owner(ranges, line):
    name = None
    for first, n in ranges:
        if first <= line:
            name = n
        else:
            break
    return name


def main():
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.startswith("device.") and f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    lines = dis.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith(".text." + KERNEL + ":"))
    end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith("\t.section") or lines[i].startswith(".text.")), len(lines))
    ranges = {"traverse.h": function_ranges(os.path.join(CSRC, "traverse.h")),
              "trace_kernels.cuh": function_ranges(os.path.join(CSRC, "trace_kernels.cuh"))}
    # the kernel body itself: inner loop = from the `do {` of the traversal loop to its `while`
    src = open(os.path.join(CSRC, "trace_kernels.cuh")).read().splitlines()
    k0 = next(i for i, l in enumerate(src, 1) if "TracePersistent(const TraceArgs a)" in l)
    do_line = next(i for i, l in enumerate(src, 1) if i > k0 and l.strip() == "do {")
    while_line = next(i for i, l in enumerate(src, 1) if i > do_line and "while (nLive >= floorLanes)" in l)

    insts = []      # (section or None, opcode)
    cur = ("trace_kernels.cuh", k0)
    for l in lines[start:end]:
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
        if not m:
            continue
        f, ln = cur
        sec = "other"
        if f in ranges:
            fn = owner(ranges[f], ln)
            if fn == "TracePersistent":
                sec = "inner_fixed" if do_line <= ln <= while_line else "outer_fixed"
            elif fn in SECTION:
                sec = SECTION[fn]
            else:
                sec = "other:" + str(fn)
        insts.append([sec, m.group(1)])
    # shared helpers take the section of the nearest classified neighbour (look ahead first: loads lead a phase)
    for i, (sec, _) in enumerate(insts):
        if sec is None:
            for j in list(range(i + 1, min(i + 40, len(insts)))) + list(range(i - 1, max(i - 40, -1), -1)):
                if insts[j][0] in ("node_phase", "tri_phase", "refill"):
                    insts[i][0] = "gate" if (insts[j][0] == "tri_phase" and j < i) else insts[j][0]
                    break
            else:
                insts[i][0] = "other"
    count = {}
    for sec, _ in insts:
        count[sec] = count.get(sec, 0) + 1
    print("%s: %d SASS instructions" % (KERNEL, len(insts)))
    for sec in sorted(count, key=lambda s: -count[s]):
        print("  %-18s %5d" % (sec, count[sec]))
    return count


if __name__ == "__main__":
    main()
