"""How much do the REFERENCE's own long-run statistics depend on the order of its particle array?

The reference sweeps its pairs in place, in pointer order (fluid.c:442-472, :577-607), and that order is
arbitrary: it is the lattice order at start-up and changes with every migration (communication.c:370-431).
This script runs the sequential oracle (oracle/sph_oracle.c orc_seq_*, pinned bit-exact to the reference,
tests/test_oracle_pin.py) on the same initial lattice with the array as is, reversed, and in two random
orders, and records the statistics tests/parity_checks.py::check_long_run_statistics compares (last 200 of
1200 steps: mean density, max density, mean height, kinetic energy per particle) relative to the golden run.
The spread is the floor under any tolerance for an implementation that sums pairs in another order
(tests/common.py LONGRUN_TOL).  Output: tests/golden/order_sensitivity.json.  ~2 minutes."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import parity_checks as pc                                             # noqa: E402
from common import load_golden                                         # noqa: E402
from oracle.oracle import GatherOracle, SeqOracle, lattice, make_problem   # noqa: E402

CASES = {"default1508": dict(n_request=1500), "goo_rect1508": dict(n_request=1500),
         "block3000": dict(n_request=3000, tank_w=21.2, water_frac=0.5), "zerog1508": dict(n_request=1500),
         "gas1508": dict(n_request=1500)}


def stats(name, perm_of):
    z, t, tw, th, h, _ = load_golden(name)
    a, _ = lattice(make_problem(**CASES[name]))
    s = SeqOracle(len(a) + 64, tw, th, t)
    s.load(a[perm_of(len(a))].copy())
    acc = []
    for k in range(1200):
        s.step()
        if k >= 1000 and k % 10 == 9:
            st = s.store()
            d = pc.density_of(lambda *args: GatherOracle(*args), tw, th, h, t, st)
            acc.append([d.mean(), d.max(), st["y"].mean(), 0.5 * (st["v_x"] ** 2 + st["v_y"] ** 2).mean()])
    return np.array(acc).mean(axis=0), z["longrun_stats"]


def main():
    rng = np.random.default_rng(0)
    orders = {"as_is": lambda n: np.arange(n), "reversed": lambda n: np.arange(n)[::-1],
              "random1": lambda n: rng.permutation(n), "random2": lambda n: rng.permutation(n)}
    out = {}
    for name in CASES:
        rows = {}
        for label, perm in orders.items():
            got, ref = stats(name, perm)
            rows[label] = {"value": [float(v) for v in got], "ratio_to_golden": [float(v) for v in got / ref]}
            print(name, label, np.round(got / ref, 3))
        r = np.array([rows[k]["ratio_to_golden"] for k in rows])
        v = np.array([rows[k]["value"] for k in rows])
        out[name] = {"columns": ["mean_density", "max_density", "mean_height", "ke_per_particle"], "runs": rows,
                     "max_rel_deviation": [float(x) for x in np.abs(r - 1).max(axis=0)],
                     "max_abs_ke": float(v[:, 3].max())}
    json.dump(out, open(os.path.join(HERE, "order_sensitivity.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
