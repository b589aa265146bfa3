#!/usr/bin/env python3
"""Per-kernel hash of the SASS in a built library (cuobjdump -sass), addresses and encodings stripped.

    python scripts/sass_hash.py sph_b200/libsph_b200.so              # print {kernel: [n_instr, hash]}
    python scripts/sass_hash.py sph_b200/libsph_b200.so --save F     # also write them to F
    python scripts/sass_hash.py sph_b200/libsph_b200.so --against F  # diff with a saved set

Used to show that a source refactor leaves the machine code of a kernel that was measured on the B200
untouched (same hash == same instructions in the same order), without a GPU.
"""
import hashlib
import json
import re
import subprocess
import sys


def kernel_hashes(lib):
    txt = subprocess.run(["cuobjdump", "-sass", lib], check=True, capture_output=True, text=True).stdout
    out, name, h, n = {}, None, None, 0
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name:
                out[name] = [n, h.hexdigest()[:16]]
            name, h, n = short(m.group(1)), hashlib.sha256(), 0
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if m and name:
            h.update(m.group(1).strip().encode())
            n += 1
    if name:
        out[name] = [n, h.hexdigest()[:16]]
    return out


def short(name):
    """k_advect<false, false> is the plain kernel (the default path): it keeps the name it had before it became a
    template; <true, false> is the stabilised gather's, "k_advect<true>" in the records; the HOLD instantiations
    (a slab's steps between two exchanges, added after the round-2b capture) are "...+hold"."""
    d = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    d = re.sub(r"^void ", "", d).split("(")[0]
    m = re.fullmatch(r"k_advect<(true|false), (true|false)>", d)
    if m:
        d = "k_advect" + ("<true>" if m.group(1) == "true" else "") + ("+hold" if m.group(2) == "true" else "")
    return d.replace("<false>", "")


if __name__ == "__main__":
    hs = kernel_hashes(sys.argv[1])
    if "--save" in sys.argv:
        with open(sys.argv[sys.argv.index("--save") + 1], "w") as f:
            json.dump(hs, f, indent=1, sort_keys=True)
    if "--against" in sys.argv:
        with open(sys.argv[sys.argv.index("--against") + 1]) as f:
            ref = json.load(f)
        bad = 0
        for k in sorted(set(hs) | set(ref)):
            a, b = hs.get(k), ref.get(k)
            state = "same" if a == b else ("NEW" if b is None else ("GONE" if a is None else "CHANGED"))
            bad += state in ("CHANGED", "GONE")
            print(f"{state:8s} {k:18s} {a} {'' if a == b else b}")
        sys.exit(1 if bad else 0)
    else:
        for k, v in sorted(hs.items()):
            print(f"{k:18s} {v[0]:6d} {v[1]}")
