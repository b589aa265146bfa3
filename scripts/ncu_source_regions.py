#!/usr/bin/env python3
"""Where a kernel's issued instructions go, from the per-instruction counters of an `ncu --set full --import-source on`
capture: runs of consecutive SASS instructions with the same execution count (= straight-line regions and loop bodies),
with their share of all executed warp instructions, their share of the stall samples and the average number of active
threads.  A loop body shows up as a region whose count is a multiple of the number of warp-particles.

    python scripts/ncu_source_regions.py gpurun_out/prof_r2_final.ncu-rep k_advect k_density k_relax > profiles/..._regions.md
"""
import csv
import io
import subprocess
import sys


def regions(rep, kern, min_share=0.01):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern],
                         capture_output=True, text=True).stdout
    lines = txt.splitlines()
    start = [i for i, l in enumerate(lines) if l.startswith('"Kernel Name"')]
    if not start:
        return None
    seg = lines[start[0] + 1:(start[1] if len(start) > 1 else None)]          # first captured launch of that kernel
    rows = list(csv.DictReader(io.StringIO("\n".join(seg))))
    ins = [(int(r["Instructions Executed"]), int(r["# Samples"]), float(r["Avg. Threads Executed"] or 0), r["Source"].strip())
           for r in rows]
    tot = sum(e for e, _, _, _ in ins)
    smp = max(1, sum(s for _, s, _, _ in ins))
    thr = sum(int(r["Thread Instructions Executed"]) for r in rows) / max(tot, 1)
    out, cur = [], [0]
    for i in range(1, len(ins)):
        if abs(ins[i][0] - ins[cur[-1]][0]) <= 0.03 * max(ins[cur[-1]][0], 1):
            cur.append(i)
        else:
            out.append(cur)
            cur = [i]
    out.append(cur)
    table = []
    for rg in out:
        e = sum(ins[i][0] for i in rg)
        s = sum(ins[i][1] for i in rg)
        if e / tot >= min_share or s / smp >= min_share:
            table.append((rg[0], rg[-1], len(rg), ins[rg[0]][0], e / tot, s / smp, sum(ins[i][2] for i in rg) / len(rg), ins[rg[0]][3]))
    return {"instructions": len(ins), "executed": tot, "threads_per_instruction": thr, "samples": smp, "regions": table}


if __name__ == "__main__":
    rep = sys.argv[1]
    print(f"# per-instruction execution counts of {rep} (ncu --page source), grouped into regions of equal count\n")
    for k in sys.argv[2:]:
        r = regions(rep, k)
        if not r:
            continue
        print(f"## {k}: {r['instructions']} SASS instructions, {r['executed']} warp instructions executed, "
              f"{r['threads_per_instruction']:.1f} active threads per instruction, {r['samples']} stall samples\n")
        print("| SASS lines | length | executions per instruction | share of executed instructions | share of stall samples | active threads | first instruction |")
        print("|---|---|---|---|---|---|---|")
        for a, b, n, e, si, ss, t, src in r["regions"]:
            print(f"| {a}-{b} | {n} | {e} | {si:.1%} | {ss:.1%} | {t:.1f} | `{src[:44]}` |")
        print()
