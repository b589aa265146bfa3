"""Summarise ncu output brought back in gpurun_out/ into small tracked files under profiles/.

  python scripts/ncu_summary.py launches gpurun_out/launches.csv profiles/rN_launches.md
  python scripts/ncu_summary.py full gpurun_out/prof.ncu-rep profiles/rN_full.csv
"""
import collections
import csv
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__shared_mem_per_block_static", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        agg.setdefault(r[ki].split("(")[0], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list summary of {src} (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare shares)\n\n")
        f.write("| kernel | launches | mean us | share |\n|---|---|---|---|\n")
        for k, v in agg.items():
            f.write(f"| {k} | {len(v)} | {sum(v) / len(v) / 1e3:.1f} | {100 * sum(v) / tot:.1f}% |\n")
    print(open(dst).read())


def short(name):
    """'void k_advect<0>(const DevParams *, ...)' -> 'k_advect'"""
    import re
    return re.sub(r"<.*$", "", name.split("(")[0].replace("void ", "")).strip()


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [short(r[idx["Kernel Name"]]) for r in rows[2:]])
        for m in KEEP:
            if m in idx:
                w.writerow([m, units[idx[m]]] + [r[idx[m]] for r in rows[2:]])
    print(open(dst).read())
    # DRAM traffic per launch for bench.py's roofline.traffic
    import json, os
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    n = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    tj = os.path.join(os.path.dirname(dst), "traffic.json")
    t = json.load(open(tj)) if os.path.exists(tj) else {}
    for r in rows[2:]:
        rd = float(r[idx["dram__bytes_read.sum"]]) * scale[units[idx["dram__bytes_read.sum"]]]
        wr = float(r[idx["dram__bytes_write.sum"]]) * scale[units[idx["dram__bytes_write.sum"]]]
        t[short(r[idx["Kernel Name"]])] = {"dram_bytes_per_launch": rd + wr, "n_particles": n, "source": os.path.basename(dst)}
    json.dump(t, open(tj, "w"), indent=1)


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
