"""How much of a permutation the two sorts of a step repair (round 1's verdict, "Next round" item 3, asked for this
figure): the fraction of particles whose SORT CELL (side h/2) differs (a) between the relaxed position of step t and the
predicted position of step t + 1 -- what sort 1 repairs -- and (b) between the predicted and the relaxed position of
step t -- what sort 2 repairs; and the fraction of entries whose INDEX in the cell-sorted arrays differs, which is what
an in-place repair would have to move (a counting sort packs the cells densely: one particle that changes cell shifts
every entry between its old and its new place by one).

CPU only: the product's kernel source compiled for the host (tests/emu -- TEST INFRASTRUCTURE); a dam-break block like
bench.py's, scaled down.    python scripts/sort_cell_changes.py [particles] [pre-roll steps] [measured steps]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import sph_b200  # noqa: E402
from emu.build_emu import build  # noqa: E402

sph_b200._lib = sph_b200._bind(C.CDLL(build()))


def cells(a, h, wx):
    ch = np.float32(h) / np.float32(2.0)
    return np.floor(a["y"] / ch).astype(np.int64) * wx + np.floor(a["x"] / ch).astype(np.int64)


def sorted_index(key, uid):
    order = np.lexsort((uid, key))          # the library's canonical order: cell, then uid
    idx = np.empty(len(order), np.int64)
    idx[order] = np.arange(len(order))
    return idx


def main():
    n_req = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    pre = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    water_frac = 0.5
    tank_w = 15.0 * (n_req / 1500.0 / water_frac * 0.58) ** 0.5        # (about bench.py's problem_dims: constant density)
    prob = sph_b200.make_problem(n_req, tank_w=tank_w, water_frac=water_frac)
    t = sph_b200.default_params(prob["h"], prob["tank_w"], prob["tank_h"], "x")
    t.mover_center_x = 0.75 * prob["tank_w"]
    c = sph_b200.Context(prob["tank_w"], prob["tank_h"], prob["h"], prob["n_global"] + 64)
    c.set_params(t)
    n = c.init_lattice(prob)
    c.step(pre)
    h = prob["h"]
    wx = int(np.floor(prob["tank_w"] / (h / 2))) + 2
    a0, u = c.download()                                        # uid order: relaxed positions of the last step
    k_rel = cells(a0, h, wx)
    f1, f2, m1, m2 = [], [], [], []
    for _ in range(steps):
        c.advect(); c.sort()
        ap, _ = c.download()                                    # predicted positions
        k_pred = cells(ap, h, wx)
        c.density(); c.relax(); c.sort()
        ar, _ = c.download()
        k_new = cells(ar, h, wx)
        f1.append(np.mean(k_pred != k_rel)); f2.append(np.mean(k_new != k_pred))
        m1.append(np.mean(sorted_index(k_pred, u) != sorted_index(k_rel, u)))
        m2.append(np.mean(sorted_index(k_new, u) != sorted_index(k_pred, u)))
        k_rel = k_new
    st = c.status()
    print(f"{n} particles, {pre} steps from the lattice, then {steps} measured; capacity_overflow {st.capacity_overflow}")
    print(f"sort 1 (relaxed -> next predicted): {100 * np.mean(f1):.1f} % of the particles change sort cell, "
          f"{100 * np.mean(m1):.1f} % of the entries change index")
    print(f"sort 2 (predicted -> relaxed):      {100 * np.mean(f2):.1f} % of the particles change sort cell, "
          f"{100 * np.mean(m2):.1f} % of the entries change index")


if __name__ == "__main__":
    main()
