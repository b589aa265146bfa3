"""The sequential oracle (oracle/sph_oracle.c orc_seq_*) is BIT-EXACT with the unmodified
reference: against the committed golden vectors everywhere, and against the live reference
(oracle/_ref/libsph_ref.so) where it exists."""
import numpy as np
import pytest

from common import bits_equal, load_golden
from oracle.oracle import PARTICLE, Ref, SeqOracle, lattice, make_problem, orc

F_ALL = ("x", "y", "v_x", "v_y", "x_prev", "y_prev", "density", "density_near", "pressure", "pressure_near")
CASES = [("default1508", 100), ("default1508", 400), ("goo_rect1508", 300), ("block3000", 150),
         ("zerog1508", 200), ("gas1508", 200)]


def seq_pairs(seq):
    nc, ni, _, _, _ = seq.lists()
    owner = np.repeat(np.arange(len(nc)), nc)
    flat = np.concatenate([ni[i, :nc[i]] for i in range(len(nc))]) if nc.sum() else np.zeros(0, "i4")
    a = np.minimum(owner, flat).astype("u8"); b = np.maximum(owner, flat).astype("u8")
    return np.sort((a << np.uint64(32)) | b), nc


@pytest.mark.parametrize("name,warm", CASES)
def test_seq_oracle_matches_golden_stage_by_stage(name, warm):
    z, t, tank_w, tank_h, h, _ = load_golden(name)
    st = z[f"w{warm}_state"]
    seq = SeqOracle(len(st) + 8, tank_w, tank_h, t)
    seq.load(st)
    seq.hash(False)                                   # lists as the reference left them (fluid.c:341)
    p, fwd = seq_pairs(seq)
    assert np.array_equal(p, z[f"w{warm}_pairs"])
    assert np.array_equal(fwd, z[f"w{warm}_fwd"])
    L = orc()
    cells = np.array([L.orc_hash_val(float(x), float(y), h, seq.lists()[4][0]) for x, y in zip(st["x"], st["y"])], "u4")
    assert np.array_equal(cells, z[f"w{warm}_cells"])
    seq.apply_gravity(); seq.viscosity(); seq.predict()
    assert bits_equal(seq.store(), z[f"w{warm}_advect"], ("x", "y", "v_x", "v_y", "x_prev", "y_prev"))
    seq.hash(True)
    assert bits_equal(seq.store(), z[f"w{warm}_density"], ("density", "density_near"))
    assert np.array_equal(seq_pairs(seq)[0], z[f"w{warm}_pairs_pred"])
    seq.relax(); seq.update_velocities()
    assert bits_equal(seq.store(), z[f"w{warm}_relaxed"], F_ALL)
    seq.hash(False)
    for _ in range(9):
        seq.step()
    assert bits_equal(seq.store(), z[f"w{warm}_after10"], F_ALL)


def test_partition_matches_golden():
    z = np.load(__import__("os").path.join(__import__("common").GOLDEN, "partition.npz"))["rows"]
    done = set()
    for n_req, tank_w, frac, nranks, rank, spacing, xs, lx, sx, ex, ng in z:
        key = (n_req, tank_w, frac, nranks)
        if key in done:
            continue
        done.add(key)
        prob = make_problem(int(n_req), tank_w=tank_w, water_frac=frac, nranks=int(nranks))
        rows = z[(z[:, 0] == n_req) & (z[:, 1] == tank_w) & (z[:, 2] == frac) & (z[:, 3] == nranks)]
        assert np.float32(prob["spacing"]) == np.float32(rows[0, 5])
        assert prob["n_global"] == int(rows[0, 10])
        for r in rows:
            sc, nc, s, e = prob["slabs"][int(r[4])]
            assert (sc, nc) == (int(r[6]), int(r[7]))
            assert np.float32(s) == np.float32(r[8]) and np.float32(e) == np.float32(r[9])


@pytest.mark.ref
@pytest.mark.skipif(not Ref.available(), reason="oracle/_ref/libsph_ref.so not built (no /root/reference here)")
def test_seq_oracle_matches_live_reference():
    ref = Ref(1500)
    prob = make_problem(1500)
    a, uid = lattice(prob)
    st = ref.state()
    assert np.array_equal(a["x"], st["x"]) and np.array_equal(a["y"], st["y"])     # geometry.c:29-59
    assert prob["n_global"] == ref.n and np.float32(prob["h"]) == np.float32(ref.h)
    seq = SeqOracle(ref.n + 8, ref.tank_w, ref.tank_h, ref.tunable)
    seq.load(st)
    for step in range(120):
        ref.step(); seq.step()
        assert bits_equal(ref.state(), seq.store(), F_ALL), f"diverged at step {step}"


@pytest.mark.ref
@pytest.mark.skipif(not Ref.available(), reason="oracle/_ref/libsph_ref.so not built")
def test_seq_oracle_matches_live_reference_on_random_soup():
    """Hostile input: random overlapping particles, coincident pairs, particles on the walls."""
    from common import random_state
    ref = Ref(1500)
    a = random_state(1200, ref.tank_w, ref.tank_h, seed=7, clustered=True)
    a[10] = a[11]; a[12]["x"] = 0.0; a[12]["y"] = 0.0; a[13]["x"] = 0.0; a[13]["y"] = 0.0   # coincident + corner
    a[14]["x"] = ref.tank_w; a[15]["y"] = ref.tank_h                                        # exactly on max
    ref.set_state(a)
    seq = SeqOracle(2000, ref.tank_w, ref.tank_h, ref.tunable)
    seq.load(a)
    ref.hash_fluid(False); seq.hash(False)
    for step in range(5):
        ref.step(); seq.step()
        r, s = ref.state(), seq.store()
        ok = all(np.array_equal(r[f].view("u4"), s[f].view("u4")) or
                 np.array_equal(np.isnan(r[f]), np.isnan(s[f])) and np.array_equal(r[f][~np.isnan(r[f])].view("u4"), s[f][~np.isnan(s[f])].view("u4"))
                 for f in F_ALL)
        assert ok, f"diverged at step {step}"
