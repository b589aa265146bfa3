"""Shared helpers for the test-suite (golden loading, tolerances)."""
import ctypes as C
import os

import numpy as np

from oracle.oracle import PARTICLE, Tunable

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# ---- stated parity tolerances (DESIGN.md "Parity") -------------------------------------------
# CUDA vs the gather oracle: same algorithm, same summation order; differences are FMA contraction
# and the fp32 evaluation of the fp64 tail of fluid.c:591.  A few ulps of the coordinate.
ULPS_POS = 8            # |dx| <= ULPS_POS * ulp(tank_w)
# gather (Jacobi) vs the reference's in-place (Gauss-Seidel) sweep, ONE step from a shared snapshot
# (the reference's own forward-vs-reverse sweep differs by 7e-3 h max, 4e-4 h rms)
ONE_STEP_MAX_H = 1e-2
ONE_STEP_RMS_H = 1e-3
# density after the hash: summation order only
DENSITY_REL = 1e-5
# long run (last 200 of 1200 steps, time averaged): chaotic trajectories, stable statistics
STAT_REL = 0.02         # mean density, max density, mean height
STAT_REL_MAXDENS = 0.05
KE_REL = 0.25           # kinetic energy per particle
# ... but never tighter than the REFERENCE's own sensitivity to the (arbitrary) order of its particle array:
# the same lattice run through the sequential oracle as is / reversed / shuffled moves these statistics by the
# amounts recorded in golden/order_sensitivity.json (make_order_sensitivity.py), e.g. the mean height of the
# "goo" heap by 14 % and the kinetic energy of the default fluid by 32 %.  An implementation that sums the
# pairs in yet another order gets 1.5 x that spread where it exceeds the figures above.
ORDER_SPREAD_FACTOR = 1.5


# The goo heap under the stabilised gather settles into one of two packings -- mean density 10.03 or 10.18
# (+0.7 % / +2.2 % of the reference's 9.97), mean height 0.640 or 0.557 (reference 0.667, own order spread 14 %) --
# and a 1-ulp change of ONE initial coordinate decides which (gather oracle against itself; the CUDA source lands
# on the other one than the oracle from the unperturbed lattice).  So for this case the mean-density bar is 3 %.
# Its kinetic energy per particle is a creeping heap's: 5e-4, 3e-3 and 1.2e-2 in three such runs (reference 2e-4, the
# reference's own order spread reaches 1.6e-2), so the absolute KE bar is 5e-2 here; a heap that does NOT settle
# has KE 3.5 (the plain gather).
GOO_STABILISED_WIDEN = {"mean_density": 0.03, "ke_abs": 0.05}


def longrun_tolerances(name):
    """-> dict(mean_density, max_density, mean_height, ke_rel, ke_abs) for check_long_run_statistics"""
    import json
    z = json.load(open(os.path.join(GOLDEN, "order_sensitivity.json")))[name]
    d = [ORDER_SPREAD_FACTOR * v for v in z["max_rel_deviation"]]
    return {"mean_density": max(STAT_REL, d[0]), "max_density": max(STAT_REL_MAXDENS, d[1]),
            "mean_height": max(STAT_REL, d[2]), "ke_rel": max(KE_REL, min(d[3], 1.0)),
            "ke_abs": max(1e-3, ORDER_SPREAD_FACTOR * z["max_abs_ke"] if d[3] > 1.0 else 1e-3)}


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    t = Tunable()
    C.memmove(C.byref(t), z["tunable"].tobytes(), 64)
    tank_w, tank_h, h, spacing = [float(v) for v in z["geom"]]
    return z, t, tank_w, tank_h, h, spacing


def pos_err_h(a, b, h):
    d = np.hypot(a["x"].astype("f8") - b["x"], a["y"].astype("f8") - b["y"]) / h
    return float(d.max()), float(np.sqrt((d ** 2).mean()))


def vel_err(a, b):
    d = np.hypot(a["v_x"].astype("f8") - b["v_x"], a["v_y"].astype("f8") - b["v_y"])
    return float(d.max()), float(np.sqrt((d ** 2).mean()))


def bits_equal(a, b, fields):
    return all(np.array_equal(a[f].view("u4"), b[f].view("u4")) for f in fields)


def ulp32(x):
    return float(np.spacing(np.float32(x)))


def random_state(n, tank_w, tank_h, seed, vmax=3.0, clustered=False):
    """Seeded particle soup inside the tank (for property tests; not a physical state)."""
    rng = np.random.default_rng(seed)
    a = np.zeros(n, PARTICLE)
    if clustered:
        cx = rng.uniform(0.2, 0.8, (8, 2)) * [tank_w, tank_h]
        k = rng.integers(0, 8, n)
        p = cx[k] + rng.normal(0, 0.03 * tank_w, (n, 2))
        a["x"] = np.clip(p[:, 0], 0, tank_w - 0.002); a["y"] = np.clip(p[:, 1], 0, tank_h - 0.002)
    else:
        a["x"] = rng.uniform(0, tank_w - 0.002, n); a["y"] = rng.uniform(0, tank_h - 0.002, n)
    a["v_x"] = rng.uniform(-vmax, vmax, n); a["v_y"] = rng.uniform(-vmax, vmax, n)
    a["x_prev"] = a["x"]; a["y_prev"] = a["y"]
    a["id"] = np.arange(n)
    return a
