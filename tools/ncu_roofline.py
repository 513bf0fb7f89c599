#!/usr/bin/env python3
"""Per-launch roofline figures of ONE kernel launch from an `ncu --set full` capture, written into
profiles/traffic.json under the workload's name (bench.py reads them into `roofline`).

    python tools/ncu_roofline.py <file.ncu-rep> <workload> <rays in that launch> [note]

What is recorded (all per launch, from the capture itself -- nothing is modelled):
  dram bytes (dram__bytes_read.sum + dram__bytes_write.sum), L2 bytes (lts__t_sectors.sum x 32), the bytes the
  kernel's global loads asked L1 for (l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum x 32), warp instructions,
  and the utilisation counters that say which resource bounds the kernel."""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return None


SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3,
         "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}


def main():
    path, workload, rays = sys.argv[1], sys.argv[2], int(sys.argv[3])
    note = sys.argv[4] if len(sys.argv) > 4 else ""
    hdr, units, rows = raw(path)
    r = rows[0]

    def get(name, scaled=True):
        if name not in hdr:
            return None
        i = hdr.index(name)
        v = num(r[i])
        if v is None:
            return None
        return v * SCALE.get(units[i], 1.0) if scaled else v

    stalls = {}
    for i, h in enumerate(hdr):
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            v = num(r[i])
            if v is not None:
                stalls[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = round(v, 3)
    top = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:5])
    dram_r, dram_w = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
    lts = get("lts__t_sectors.sum", False)
    l1 = get("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", False)
    inst = get("smsp__inst_executed.sum", False)
    try:
        head = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    except Exception:
        head = ""
    e = {
        "kernel": r[hdr.index("Kernel Name")],
        "rays_in_launch": rays,
        "duration_ms_under_ncu": get("gpu__time_duration.sum"),
        "dram_bytes_per_launch": (dram_r or 0) + (dram_w or 0),
        "dram_bytes_read": dram_r, "dram_bytes_write": dram_w,
        "lts_bytes_per_launch": lts * 32.0 if lts else None,
        "l1_global_load_bytes_per_launch": l1 * 32.0 if l1 else None,
        "warp_instructions_per_launch": inst,
        "warp_instructions_per_ray": round(inst / rays, 2) if inst else None,
        "issue_active_pct": get("smsp__issue_active.avg.pct_of_peak_sustained_active", False),
        "alu_pipe_pct": get("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", False),
        "fma_pipe_pct": get("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", False),
        "l1_data_pipe_pct": get("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", False),
        "lts_throughput_pct": get("lts__throughput.avg.pct_of_peak_sustained_elapsed", False),
        "dram_throughput_pct": get("dram__throughput.avg.pct_of_peak_sustained_elapsed", False),
        "lts_hit_rate_pct": get("lts__t_sector_hit_rate.pct", False),
        "l1_hit_rate_pct": get("l1tex__t_sector_hit_rate.pct", False),
        "threads_per_instruction": get("smsp__thread_inst_executed_per_inst_executed.ratio", False),
        "warps_active_pct": get("sm__warps_active.avg.pct_of_peak_sustained_active", False),
        "registers_per_thread": get("launch__registers_per_thread", False),
        "grid_size": get("launch__grid_size", False),
        "top_stalls_warps_per_issue": top,
        "source": "profiles/%s (ncu --set full --clock-control none, one launch)%s" % (os.path.basename(path).replace(".ncu-rep", "_ncu_summary.txt"), (" -- " + note) if note else ""),
        "captured_at_commit": head,
    }
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        t = json.load(open(tpath))
    except Exception:
        t = {}
    t[workload] = e
    json.dump(t, open(tpath, "w"), indent=1)
    print(json.dumps(e, indent=1))


if __name__ == "__main__":
    main()
