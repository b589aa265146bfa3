#!/usr/bin/env python3
"""Loop bodies of a kernel in the SASS of a built library: for every backward branch, the number of instructions
between its target and the branch, and the instruction mix of the largest ones.

    python scripts/sass_loops.py sph_b200/libsph_b200.so k_advect k_density
    python scripts/sass_loops.py sph_b200/variants/packed.so k_advect k_density

Static evidence for instruction-issue-bound kernels: the unrolled main candidate loop of a gather
handles 4 candidates per trip (SPH_UNROLL), so its length / 4 is the number of issued instructions per candidate.
"""
import collections
import re
import subprocess
import sys


def kernels(lib):
    txt = subprocess.run(["cuobjdump", "-sass", lib], check=True, capture_output=True, text=True).stdout
    out, name = collections.OrderedDict(), None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"^void ", "", name).split("(")[0]
            out[name] = []
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and name:
            out[name].append((int(m.group(1), 16), m.group(2).strip()))
    return out


def loops(instrs):
    addr = {a: i for i, (a, _) in enumerate(instrs)}
    found = []
    for i, (a, text) in enumerate(instrs):
        m = re.search(r"\bBRA\s+(?:\S+,\s*)?(0x[0-9a-f]+)", text)
        if m:
            t = int(m.group(1), 16)
            if t <= a and t in addr:
                found.append((i - addr[t] + 1, addr[t], i))
    return sorted(found, reverse=True)


if __name__ == "__main__":
    ks = kernels(sys.argv[1])
    for want in sys.argv[2:]:
        for name, ins in ks.items():
            if name.replace("<false>", "") != want and name != want:
                continue
            print(f"## {name}: {len(ins)} instructions")
            for n, b, e in loops(ins)[:6]:
                mix = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for _, t in ins[b:e + 1])
                top = ", ".join(f"{k} {v}" for k, v in mix.most_common(9))
                print(f"  loop of {n:4d} instructions at {ins[b][0]:#06x}: {top}")
