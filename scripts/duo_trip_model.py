"""Work model of the SPH_DENSITY_DUO formulation of k_density (two adjacent sorted entries per thread over the union of
their candidate ranges) on settled states from tests/golden: loop trips and candidate loads per
particle against the shipped formulation (two candidates per trip + one left-over trip per odd segment).  CPU only;
a count of work, not a time."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from common import load_golden  # noqa: E402

DIV = 2


def model(name, warm, near=5):
    z, t, tank_w, tank_h, h, _ = load_golden(name)
    st = z[f"w{warm}_state"]
    ch = h / DIV
    gx = np.floor(st["x"] / ch).astype(int); gy = np.floor(st["y"] / ch).astype(int)
    wx = gx.max() + 1; rows = gy.max() + 1
    key = gy * wx + gx
    order = np.lexsort((np.arange(len(st)), key))
    gx, gy, key = gx[order], gy[order], key[order]
    n = len(st)
    start = np.searchsorted(key, np.arange(wx * rows + 1))

    def ranges(i):
        out = []
        c0, c1 = max(gx[i] - DIV, 0), min(gx[i] + DIV, wx - 1)
        for d in range(-DIV, DIV + 1):
            r = gy[i] + d
            out.append((0, 0) if r < 0 or r >= rows else (start[r * wx + c0], start[r * wx + c1 + 1]))
        return out
    R = [ranges(i) for i in range(n)]
    trips0 = loads0 = rows0 = 0
    for i in range(n):
        for d, (b, e) in enumerate(R[i]):
            segs = [(b, i), (i + 1, e)] if d == DIV else [(b, e)]
            for jb, je in segs:
                L = max(je - jb, 0)
                trips0 += L // 2 + L % 2; loads0 += L
            rows0 += 1
    trips1 = loads1 = rows1 = passes2 = 0
    for a in range(0, n, 2):
        b_ = a + 1
        if b_ < n and gy[a] == gy[b_] and abs(gx[b_] - gx[a]) <= near:
            for (b0, e0), (b1, e1) in zip(R[a], R[b_]):
                L = max(e0, e1) - min(b0, b1) if e0 > b0 and e1 > b1 else (e0 - b0) + (e1 - b1)
                trips1 += L; loads1 += L; rows1 += 1
        else:
            passes2 += 1
            for i in ([a, b_] if b_ < n else [a]):
                for (b, e) in R[i]:
                    trips1 += e - b; loads1 += e - b; rows1 += 1
    cand = loads0 / n
    print(f"{name:14s} n={n} candidates/particle {cand:5.1f} | shipped: trips {trips0 / n:5.1f} loads {loads0 / n:5.1f} row set-ups {rows0 / n:.1f}"
          f" | duo: trips {trips1 / n:5.1f} loads {loads1 / n:5.1f} row set-ups {rows1 / n:.2f} serial pairs {100.0 * passes2 / (n / 2):.1f} %")


if __name__ == "__main__":
    for name, warm in (("default1508", 400), ("block3000", 150), ("goo_rect1508", 300)):
        model(name, warm)
