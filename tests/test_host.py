"""C host layer (sph_b200/host, include/sph_host.h) against the reference's golden vectors and the
oracle's restatement: partition, lattice, parameter presets, load balancer, partition add/remove."""
import ctypes as C
import os

import numpy as np
import pytest

import sph_b200
from common import GOLDEN
from oracle import oracle as orc_mod


def test_host_symbols_exported(built_lib):
    L = C.CDLL(built_lib)
    for s in sph_b200.HOST_SYMBOLS:
        assert hasattr(L, s)
    hdr = open(os.path.join(os.path.dirname(built_lib), "..", "include", "sph_host.h")).read()
    import re
    assert set(re.findall(r"\b(sph_host_[a-z0-9_]+)\s*\(", hdr)) == set(sph_b200.HOST_SYMBOLS)


def test_partition_matches_reference_golden(built_lib):
    z = np.load(os.path.join(GOLDEN, "partition.npz"))["rows"]
    for n_req, tank_w, frac, nranks in {tuple(r[:4]) for r in z}:
        prob = sph_b200.make_problem(int(n_req), tank_w=tank_w, water_frac=frac, nranks=int(nranks))
        rows = z[(z[:, 0] == n_req) & (z[:, 1] == tank_w) & (z[:, 2] == frac) & (z[:, 3] == nranks)]
        assert np.float32(prob["spacing"]) == np.float32(rows[0, 5]) and prob["n_global"] == int(rows[0, 10])
        for r in rows:
            sc, nc, s, e = prob["slabs"][int(r[4])]
            assert (sc, nc) == (int(r[6]), int(r[7]))
            assert np.float32(s) == np.float32(r[8]) and np.float32(e) == np.float32(r[9])


def test_lattice_matches_reference_golden(built_lib):
    z = np.load(os.path.join(GOLDEN, "default1508.npz"))
    prob = sph_b200.make_problem(1500)
    a, uid = sph_b200.lattice(prob)
    # the golden w100 state is 100 steps in; the multirank r1 fixture has uids; check lattice through the oracle
    b, ub = orc_mod.lattice(orc_mod.make_problem(1500))
    assert np.array_equal(a["x"], b["x"]) and np.array_equal(a["y"], b["y"]) and np.array_equal(uid, ub)
    assert len(a) == 1508 and np.float32(prob["h"]) == z["geom"][2]
    # 3-slab split: the slabs tile the global lattice
    p3 = sph_b200.make_problem(1500, nranks=3)
    parts = [sph_b200.lattice(p3, r) for r in range(3)]
    alluid = np.concatenate([u for _, u in parts])
    assert np.array_equal(np.sort(alluid), np.arange(1508))
    allx = np.concatenate([p["x"] for p, _ in parts])[np.argsort(alluid)]
    assert np.array_equal(allx, a["x"][np.argsort(uid)])


def test_default_params_and_presets(built_lib):
    for preset in "xyab":
        t = sph_b200.default_params(0.58, 15.0, 8.4375, preset)
        o = orc_mod.default_tunable(0.58, 15.0, 8.4375, preset)
        assert bytes(C.string_at(C.addressof(t), 64)) == bytes(C.string_at(C.addressof(o), 64))


def test_balance_matches_oracle_on_random_cases(built_lib):
    rng = np.random.default_rng(5)
    L = orc_mod.orc()
    for _ in range(300):
        n = int(rng.integers(1, 9))
        h = float(np.float32(rng.uniform(0.3, 1.0)))
        cuts = np.sort(rng.uniform(0, 50, n - 1)).astype("f4")
        xs = np.concatenate([[0], cuts, [50]]).astype("f4")
        edges = [(float(xs[i]), float(xs[i + 1])) for i in range(n)]
        counts = rng.integers(0, 4000, n).astype("i4")
        got = sph_b200.balance(edges, counts, h)
        m = (orc_mod.Tunable * n)()
        for r, (s, e) in enumerate(edges):
            m[r].smoothing_radius = h; m[r].node_start_x = s; m[r].node_end_x = e
        L.orc_balance(m, n, counts.ctypes.data_as(C.c_void_p), int(counts.sum()))
        want = [(float(m[r].node_start_x), float(m[r].node_end_x)) for r in range(n)]
        assert got == want


@pytest.mark.ref
@pytest.mark.skipif(not orc_mod.Ref.available(), reason="oracle/_ref not built")
def test_balance_matches_reference_harness_restatement(built_lib):
    """refh_balance carries the arithmetic of renderer.c:427-477 next to the reference build."""
    R = orc_mod.Ref.lib()
    rng = np.random.default_rng(9)
    for _ in range(100):
        n = int(rng.integers(2, 9)); h = float(np.float32(rng.uniform(0.3, 1.0)))
        xs = np.concatenate([[0], np.sort(rng.uniform(0, 30, n - 1)), [30]]).astype("f4")
        edges = [(float(xs[i]), float(xs[i + 1])) for i in range(n)]
        counts = rng.integers(0, 3000, n).astype("i4")
        m = (orc_mod.Tunable * n)()
        for r, (s, e) in enumerate(edges):
            m[r].smoothing_radius = h; m[r].node_start_x = s; m[r].node_end_x = e
        R.refh_balance(m, n, counts.ctypes.data_as(C.c_void_p), int(counts.sum()))
        assert sph_b200.balance(edges, counts, h) == [(float(m[r].node_start_x), float(m[r].node_end_x)) for r in range(n)]


def test_remove_and_add_partition(built_lib):
    """controls.c:405-455: removing parks the last slab at end+1 and widens its left neighbour;
    adding splits the last active slab in half."""
    L = sph_b200._host()
    m = (sph_b200.Tunable * 4)()
    for r in range(4):
        m[r].smoothing_radius = 0.5; m[r].node_start_x = 5.0 * r; m[r].node_end_x = 5.0 * (r + 1); m[r].active = bytes([1])
    n = L.sph_host_remove_partition(m, 4)
    assert n == 3 and m[2].node_end_x == 20.0 and m[3].node_start_x == 21.0 and m[3].node_end_x == 21.0 and m[3].active == bytes([0])
    n = L.sph_host_add_partition(m, n, 4)
    assert n == 4 and m[2].node_end_x == 15.0 and m[3].node_start_x == 15.0 and m[3].node_end_x == 20.0 and m[3].active == bytes([1])
    assert L.sph_host_remove_partition(m, 1) == 1 and L.sph_host_add_partition(m, 4, 4) == 4


def test_one_exchange_edge_filter_keeps_interior_slabs_wider_than_the_layer():
    from sph_b200.slab import keep_slabs_wider_than
    old = [(0.0, 4.0), (4.0, 6.1), (6.1, 10.0), (10.0, 15.0)]
    # edge 0|1 moves right (slab 1 shrinks to 2.0 < 2.05: undone), edge 2|3 moves left (slab 2 shrinks to 3.8: fine)
    new = [(0.0, 4.1), (4.1, 6.1), (6.1, 9.9), (9.9, 15.0)]
    out = keep_slabs_wider_than(old, new, 2.05, 4)
    assert out == [(0.0, 4.0), (4.0, 6.1), (6.1, 9.9), (9.9, 15.0)]
    # boundary slabs may be as narrow as the reference allows: nothing lies beyond them
    new = [(0.0, 3.9), (3.9, 6.1), (6.1, 10.0), (10.0, 15.0)]
    assert keep_slabs_wider_than(old, new, 5.0, 4) == new
    # unchanged edges stay unchanged, parked slabs (beyond nactive) are not looked at
    assert keep_slabs_wider_than(old, old, 100.0, 3) == old


def test_time_proportional_edge_policy_converges_and_respects_the_minimum_width():
    """sph_host_balance_time (not the reference's policy): edges move in proportion to the measured imbalance of the two
    slabs they separate.  A synthetic cost density that varies by +-30 % across 8 slabs is levelled to within 3 % in 150
    frames at two smoothing radii per frame (the reference's h/8 per frame would need over a thousand), no slab ever gets
    narrower than the minimum, the outer edges stay on the tank, and identical inputs give identical edges."""
    import sph_b200
    h = 0.58
    dens = lambda x: 1.0 + 0.3 * np.sin(x / 120.0) - 0.25 * (x > 700)
    edges = [(100.0 * r, 100.0 * (r + 1)) for r in range(8)]

    def times(edges):
        out = []
        for a, b in edges:
            xs = np.linspace(a, b, 2001)
            out.append(int(1000 * np.trapezoid(dens(xs), xs) / 100.0))
        return out
    first = max(times(edges)) / np.mean(times(edges))
    for _ in range(150):
        t = times(edges)
        again = sph_b200.balance_time(edges, t, h, 8, gain=0.5, max_shift_h=2.0, min_width_h=7.0)
        edges = sph_b200.balance_time(edges, t, h, 8, gain=0.5, max_shift_h=2.0, min_width_h=7.0)
        assert edges == again
        assert all(b - a >= 7.0 * h - 1e-4 for a, b in edges)
        assert all(abs(edges[r][1] - edges[r + 1][0]) < 1e-5 for r in range(7))
    t = times(edges)
    assert first > 1.3 and max(t) / np.mean(t) < 1.03, (first, t)
    assert edges[0][0] == 0.0 and edges[-1][1] == 800.0
    # a slab at the minimum width is not shrunk further, however slow it is
    tight = [(0.0, 10.0), (10.0, 10.0 + 7.0 * h), (10.0 + 7.0 * h, 30.0)]
    out = sph_b200.balance_time(tight, [100, 900, 100], h, 3, gain=0.5, max_shift_h=2.0, min_width_h=7.0)
    assert out[1][1] - out[1][0] >= 7.0 * h - 1e-5


def test_time_proportional_edge_policy_never_leaves_a_slab_narrower_than_the_minimum():
    """Random slab layouts (2-8 slabs, some barely wider than the minimum) and random measured times, 30 frames each: after
    every call the slabs still tile the tank, no edge moved further than the step limit and NO slab is narrower than the
    minimum.  The first version tested the edges in one left-to-right pass and accepted edge e against a move of edge
    e + 1 that it then withdrew: 49 of 3000 such runs ended with a slab up to 2 h narrower than the layer its
    neighbours' ghosts need.  (One exchange per step was shielded by slab.keep_slabs_wider_than; two exchanges per step
    with --balance time was not.)"""
    import random
    import sph_b200
    h = 0.58
    moved = 0
    for seed in range(400):
        rng = random.Random(seed)
        K = rng.randint(2, 8)
        W = rng.uniform(40, 600) * h
        layer = rng.choice([2.0, 3.5, 7.0])
        cuts = sorted(rng.uniform(0, W) for _ in range(K - 1))
        edges = [(a, b) for a, b in zip([0.0] + cuts, cuts + [W])]
        if min(b - a for a, b in edges) < (layer + 0.1) * h:
            continue
        busy = [rng.randint(1, 400) for _ in range(K)]
        for _ in range(30):
            new = sph_b200.balance_time(edges, busy, h, K, gain=0.5, max_shift_h=2.0, min_width_h=layer)
            assert abs(new[0][0]) < 1e-6 and abs(new[-1][1] - W) < 1e-3 and all(new[r][1] == new[r + 1][0] for r in range(K - 1))
            assert all(b - a >= layer * h - 1e-4 for a, b in new), (seed, [(b - a) / h for a, b in edges], [(b - a) / h for a, b in new], busy)
            assert all(abs(new[r][1] - edges[r][1]) <= 2.0 * h + 1e-4 for r in range(K - 1))
            # ... and every slab keeps a layer of its OLD extent (a slab whose two edges move the same way keeps its
            # width while its old and new extents drift apart: in the step the edges land, its neighbour's ghosts would
            # still be owned by the slab beyond -- soak run 93072 of tests/fuzz/fuzz_slabs.py)
            for r in range(K - 1):
                if new[r][1] < edges[r][1]:
                    assert new[r][1] - edges[r][0] >= layer * h - 1e-4, (seed, r, "left slab", (new[r][1] - edges[r][0]) / h)
                elif new[r][1] > edges[r][1]:
                    assert edges[r + 1][1] - new[r][1] >= layer * h - 1e-4, (seed, r, "right slab", (edges[r + 1][1] - new[r][1]) / h)
            moved += new != edges
            edges = new
            busy = [max(1, int(b * rng.uniform(0.8, 1.25))) for b in busy]
    assert moved > 1000


def test_edge_filter_of_the_one_exchange_modes_holds_for_random_layouts():
    """slab.keep_slabs_wider_than behind the reference's balancer (h/8 per frame, renderer.c:427-477) on random layouts and
    random counts: no interior slab ever gets narrower than the ghost layer.  One pass over the edges was not enough -- a
    slab whose two edges had both been moving right came out 0.05 h too narrow once the second move was undone (1 of
    93 000 calls of a randomised check); the filter now repeats until nothing is undone."""
    import random
    import sph_b200
    from sph_b200.slab import keep_slabs_wider_than
    h = 0.58
    calls = 0
    for seed in list(range(300)) + [2267]:
        rng = random.Random(seed)
        K = rng.randint(3, 8)
        W = rng.uniform(40, 300) * h
        layer = rng.choice([3.5, 4.5, 7.0])
        cuts = sorted(rng.uniform(0, W) for _ in range(K - 1))
        edges = [(a, b) for a, b in zip([0.0] + cuts, cuts + [W])]
        if min(b - a for a, b in edges[1:-1]) < layer * h:
            continue
        for _ in range(40):
            counts = [rng.randint(0, 4000) for _ in range(K)]
            new = keep_slabs_wider_than(edges, sph_b200.balance(edges, counts, h, K), layer * h, K)
            assert all(new[r][1] == new[r + 1][0] for r in range(K - 1))
            assert all(b - a >= layer * h - 1e-5 for a, b in new[1:-1]), (seed, [(b - a) / h for a, b in edges], [(b - a) / h for a, b in new])
            edges = new
            calls += 1
    assert calls > 3000


def test_balancer_and_partition_controls_match_the_compiled_reference_functions(built_lib):
    """check_partition_left (renderer.c:427-477), remove_partition and add_partition (controls.c:405-455) as COMPILED from
    the reference (oracle/_ref/libref_full.so, the unmodified renderer.c / controls.c), called on a render_t built here,
    against sph_host_balance / sph_host_remove_partition / sph_host_add_partition: random layouts of 1-8 slabs, random
    counts, sequences of 30 calls -- every edge float for float, the number of active slabs too.  (The other balance tests
    compare with restatements; a run of 90 000 such calls found no difference.)"""
    import random
    import sph_b200
    ref = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle", "_ref", "libref_full.so")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref not built")
    R = C.CDLL(ref)
    T = sph_b200.Tunable

    class Render(C.Structure):          # the leading members of render_t (renderer.h:44-61)
        _fields_ = [("sim_width", C.c_float), ("sim_height", C.c_float), ("screen_width", C.c_float), ("screen_height", C.c_float),
                    ("selected_parameter", C.c_int), ("node_params", C.c_void_p), ("master_params", C.c_void_p),
                    ("num_compute_procs", C.c_int), ("num_compute_procs_active", C.c_int), ("rest", C.c_char * 64)]
    assert Render.master_params.offset == 32 and Render.num_compute_procs_active.offset == 44
    h = 0.580948

    def reference(edges, nactive, fn, counts=None):
        K = len(edges)
        m = (T * K)()
        for r, (s, e) in enumerate(edges):
            m[r].smoothing_radius = h; m[r].node_start_x = s; m[r].node_end_x = e
        rs = Render(); rs.master_params = rs.node_params = C.addressof(m); rs.num_compute_procs = K; rs.num_compute_procs_active = nactive
        if fn == "balance":
            R.check_partition_left(C.byref(rs), (C.c_int * K)(*counts), C.c_int(sum(counts)))
        else:
            getattr(R, fn + "_partition")(C.byref(rs))
        return [(float(m[r].node_start_x), float(m[r].node_end_x)) for r in range(K)], rs.num_compute_procs_active

    calls = 0
    for seed in range(150):
        rng = random.Random(seed)
        K = rng.randint(1, 8)
        W = rng.uniform(10, 300)
        cuts = sorted(rng.uniform(0, W) for _ in range(K - 1))
        edges = [(float(np.float32(a)), float(np.float32(b))) for a, b in zip([0.0] + cuts, cuts + [W])]
        nactive = K
        for _ in range(30):
            op = rng.choice(["balance"] * 6 + ["remove", "add"])
            if op == "balance":
                counts = [rng.randint(0, 3000) * 2 if r < nactive else 0 for r in range(K)]     # coordinate counts, renderer.c:280,290
                want, wn = reference(edges, nactive, op, counts)
                got, gn = sph_b200.balance(edges, counts, h, nactive), nactive
            elif op == "remove":
                want, wn = reference(edges, nactive, op)
                got, gn = sph_b200.remove_partition(edges, h, nactive)
            else:
                want, wn = reference(edges, nactive, op)
                got, gn = sph_b200.add_partition(edges, h, nactive)
            assert gn == wn, (seed, op)
            assert [tuple(np.float32(x) for x in e) for e in got] == [tuple(np.float32(x) for x in e) for e in want], (seed, op, edges)
            edges, nactive = got, gn
            calls += 1
    assert calls == 4500


def test_presets_match_the_compiled_reference_functions(built_lib):
    """set_fluid_x / y / a / b of the compiled controls.c (controls.c:344-401) applied to a parameter block against
    sph_host_preset on the same block: all 64 bytes."""
    import sph_b200
    ref = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle", "_ref", "libref_full.so")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref not built")
    R = C.CDLL(ref)
    L = C.CDLL(built_lib)

    class Render(C.Structure):          # the leading members of render_t (renderer.h:44-61)
        _fields_ = [("sim_width", C.c_float), ("sim_height", C.c_float), ("screen_width", C.c_float), ("screen_height", C.c_float),
                    ("selected_parameter", C.c_int), ("node_params", C.c_void_p), ("master_params", C.c_void_p),
                    ("num_compute_procs", C.c_int), ("num_compute_procs_active", C.c_int), ("rest", C.c_char * 64)]
    for before in "xyab":
        for which in "xyab":
            m = (sph_b200.Tunable * 2)()
            for r in range(2):
                m[r] = sph_b200.default_params(0.58, 15.0, 8.4375, before)
            mine = sph_b200.default_params(0.58, 15.0, 8.4375, before)
            rs = Render(); rs.master_params = rs.node_params = C.addressof(m); rs.num_compute_procs = rs.num_compute_procs_active = 2
            getattr(R, "set_fluid_" + which)(C.byref(rs))
            assert L.sph_host_preset(C.byref(mine), C.c_char(which.encode())) == 0
            for r in range(2):
                assert bytes(C.string_at(C.addressof(m[r]), 64)) == bytes(C.string_at(C.addressof(mine), 64)), (before, which, r)


def test_autopilot_matches_the_compiled_reference_when_gl_x_is_rederived(built_lib):
    """update_inactive_state of the compiled renderer.c (renderer.c:491-531: the mover's idle path, which also resets the
    preset and the mover size every frame) against sph_host_mover_autopilot, 1200 frames (six reversals) in three tank
    sizes: bit-identical mover centres when the caller forms gl_x from the centre before each call, as the reference
    does (renderer.c:494); with gl_x carried over the path is the same up to the frame at which it reverses."""
    import sph_b200
    ref = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle", "_ref", "libref_full.so")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref not built")
    R = C.CDLL(ref)
    L = C.CDLL(built_lib)

    class Render(C.Structure):          # render_t (renderer.h:44-61)
        _fields_ = [("sim_width", C.c_float), ("sim_height", C.c_float), ("screen_width", C.c_float), ("screen_height", C.c_float),
                    ("selected_parameter", C.c_int), ("node_params", C.c_void_p), ("master_params", C.c_void_p),
                    ("num_compute_procs", C.c_int), ("num_compute_procs_active", C.c_int), ("show_dividers", C.c_bool),
                    ("pause", C.c_bool), ("quit_mode", C.c_bool), ("last_activity_time", C.c_double),
                    ("exit_menu_state", C.c_void_p), ("return_value", C.c_int), ("liquid", C.c_bool)]
    f32 = np.float32
    for tank_w in (15.0, 41.7, 387.3):
        tank_h = float(f32(tank_w) / f32(16.0 / 9.0))
        m = (sph_b200.Tunable * 1)()
        m[0] = sph_b200.default_params(0.58, tank_w, tank_h, "b")
        rs = Render(); rs.sim_width = tank_w; rs.sim_height = tank_h; rs.master_params = rs.node_params = C.addressof(m)
        rs.num_compute_procs = rs.num_compute_procs_active = 1
        mine = sph_b200.default_params(0.58, tank_w, tank_h, "b")
        kept = sph_b200.default_params(0.58, tank_w, tank_h, "b")
        gl_kept = C.c_float(f32(kept.mover_center_x) / (f32(tank_w) * f32(0.5)) - f32(1.0)); d_kept = C.c_int(1)
        d = C.c_int(1)
        far = 0.0
        for frame in range(1200):
            R.update_inactive_state(C.byref(rs))
            gl = C.c_float(f32(mine.mover_center_x) / (f32(tank_w) * f32(0.5)) - f32(1.0))
            L.sph_host_mover_autopilot(C.byref(mine), C.c_float(tank_w), C.c_float(tank_h), C.byref(gl), C.byref(d))
            assert (m[0].mover_center_x, m[0].mover_center_y) == (mine.mover_center_x, mine.mover_center_y), (tank_w, frame)
            L.sph_host_mover_autopilot(C.byref(kept), C.c_float(tank_w), C.c_float(tank_h), C.byref(gl_kept), C.byref(d_kept))
            far = max(far, abs(kept.mover_center_x - m[0].mover_center_x) / tank_w)
        # carried over: never further from the reference's path than a few frames' travel (0.005 of the width per frame)
        assert far <= 6 * 0.005 + 1e-6, (tank_w, far)
        # (the idle path also puts the fluid back to preset x and the mover to 2 x 2: controls.c:344-357, :333-338)
        want = sph_b200.default_params(0.58, tank_w, tank_h, "x")
        assert L.sph_host_preset(C.byref(mine), C.c_char(b"x")) == 0
        for fld in ("k", "k_near", "k_spring", "sigma", "beta", "rest_density", "g"):
            assert getattr(m[0], fld) == getattr(want, fld) == getattr(mine, fld), fld
