#!/usr/bin/env python3
"""Summarise an .ncu-rep (read on the CPU box): key counters + instruction mix + stall samples."""
import csv, collections, subprocess, sys, io

def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max"]

def main():
    path = sys.argv[1]
    hdr, units, rows = raw(path)
    for r in rows:
        print("== kernel:", r[hdr.index("Kernel Name")][:70])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print("  %-88s %s %s" % (w, r[i], units[i]))
        stalls = []
        for i, h in enumerate(hdr):
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(r[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        print("  stall reasons (warps per issue):", ", ".join("%s %.2f" % (n, v) for v, n in stalls[:8]))
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]
    ia, ie, it, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Avg. Threads Executed"), hdr.index("# Samples")
    ops, thr, samp = collections.Counter(), collections.Counter(), collections.Counter()
    tot = 0
    for r in rows[2:]:
        if len(r) <= ie:
            continue
        try:
            n = int(r[ie])
        except ValueError:
            continue
        toks = r[ia].split()
        if not toks:
            continue
        op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
        op = op.split(".")[0]
        ops[op] += n
        tot += n
        try:
            thr[op] += n * float(r[it]); samp[op] += int(r[isamp])
        except ValueError:
            pass
    print("-- instruction mix (warp instructions, avg active threads, stall samples)")
    for op, n in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 16):
        print("  %-10s %14d %5.1f%%  avgthreads %5.1f  samples %d" % (op, n, 100.0 * n / tot, thr[op] / max(1, n), samp[op]))

if __name__ == "__main__":
    main()
