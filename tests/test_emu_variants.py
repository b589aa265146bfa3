"""Build-flag variants of the kernels, checked in the kernel-source emulator (tests/emu) before they are ever
timed on the B200.

SPH_RELAX_PD4=1: k_relax's neighbour walk reads one 16-byte (x, y, density, density_near) record instead of two.
SPH_PACKED=1 (+ SPH_PACKED_RELAX=1 for k_relax's pair physics): the candidate loops of k_advect / k_coupling /
k_density take two candidates per trip with packed FP32 instructions (FADD2 / FMUL2 / FFMA2).  Every packed operation is the same round-to-nearest operation as its
scalar counterpart and the order of every sum is kept, so in the emulator (where neither build contracts
a*b+c) the variant must reproduce the default build BIT FOR BIT -- positions, velocities, densities, neighbour
masks' effect on the relaxation, with and without the stabilised viscosity gather, odd and even range lengths,
ghost entries and all."""
import ctypes as C

import numpy as np
import pytest

import sph_b200
from common import load_golden
from emu.build_emu import build as build_emu
from test_gpu_parity import as_sph


def run(libpath, name, warm, steps, gamma, monkeypatch):
    monkeypatch.setattr(sph_b200, "_lib", sph_b200._bind(C.CDLL(libpath)))
    z, t, tank_w, tank_h, h, _ = load_golden(name)
    st = z[f"w{warm}_state"]
    c = sph_b200.Context(tank_w, tank_h, h, len(st) + 64)
    c.set_params(as_sph(t)); c.set_viscosity_stabilisation(gamma); c.upload(st)
    c.step(steps)
    c.advect(); c.sort(); c.density()
    dens, _ = c.download()
    c.relax(); c.sort()
    out, _ = c.download()
    return dens, out


@pytest.mark.parametrize("name,warm,gamma", [("default1508", 400, 0.0), ("block3000", 150, 0.0), ("goo_rect1508", 300, 0.0),
                                             ("goo_rect1508", 300, 0.5), ("gas1508", 200, 0.5)])
def test_packed_fp32_variant_is_bit_identical_in_the_emulator(built_lib, monkeypatch, name, warm, gamma):
    base = build_emu()
    packed = build_emu(defines=("SPH_PACKED=1", "SPH_PACKED_RELAX=1", "SPH_RELAX_PD4=1"), name="libsph_emu_packed.so")
    d0, a0 = run(base, name, warm, 12, gamma, monkeypatch)
    d1, a1 = run(packed, name, warm, 12, gamma, monkeypatch)
    for f in ("density", "density_near", "x", "y"):
        assert np.array_equal(d0[f].view("u4"), d1[f].view("u4")), ("density stage", f)
    for f in ("x", "y", "v_x", "v_y"):
        assert np.array_equal(a0[f].view("u4"), a1[f].view("u4")), ("after the step", f)


@pytest.mark.parametrize("packed", [False, True])
def test_trimmed_advect_loop_agrees_with_the_gather_oracle_and_redoes_rows_where_the_clamp_binds(built_lib, monkeypatch, packed):
    """SPH_TRIM=1: k_advect applies a pair's half impulse without the per-component clamp and redoes, with the exact
    body, only the rows in which the clamp COULD bind (max |t| h > 2.5).  (1) Rounding-level agreement with the gather
    oracle on settled fluids (no row is redone).  (2) A violent goo state -- every second particle thrown at 5 units/s
    against its neighbours, so that the clamp binds in many rows -- must still agree with the oracle's clamped
    impulses: without the redo the velocities would be off by whole units."""
    import parity_checks as pc
    from test_gpu_parity import Cuda, make_oracle
    defs = ("SPH_TRIM=1",) + (("SPH_PACKED=1", "SPH_PACKED_RELAX=1") if packed else ())
    lib = build_emu(defines=defs, name="libsph_emu_trim%d.so" % packed)
    monkeypatch.setattr(sph_b200, "_lib", sph_b200._bind(C.CDLL(lib)))
    mk = lambda *a: Cuda(*a)
    pc.check_tight_vs_gather_oracle(mk, make_oracle, "default1508", 400, steps=3)
    pc.check_tight_vs_gather_oracle(mk, make_oracle, "goo_rect1508", 300, steps=2)
    # (2)
    z, t, tank_w, tank_h, h, _ = load_golden("goo_rect1508")
    st = z["w300_state"].copy()
    rng = np.random.default_rng(5)
    st["v_x"] = np.where(np.arange(len(st)) % 2 == 0, 5.0, -5.0).astype("f4") * rng.uniform(0.5, 1.0, len(st)).astype("f4")
    st["v_y"] = rng.uniform(-5, 5, len(st)).astype("f4")
    b = mk(tank_w, tank_h, h, len(st) + 64); o = make_oracle(tank_w, tank_h, h, len(st) + 64)
    for s in (b, o):
        s.set_params(t); s.upload(st)
    for s in (b, o):
        s.advect(); s.sort()
    a, ua = b.download(); r, ur = o.download()
    assert np.array_equal(ua, ur)
    # predicted positions = x + v dt: a unit of velocity error is dt = 8.3e-3 of position
    assert np.abs(a["x"] - r["x"]).max() <= 1e-5 and np.abs(a["y"] - r["y"]).max() <= 1e-5
    # the clamp did bind for some pair of this state (brute force over the pairs; fluid.c:451-459 halved)
    x, y, vx, vy = (st[f].astype("f8") for f in ("x", "y", "v_x", "v_y"))
    dx = x[None, :] - x[:, None]; dy = y[None, :] - y[:, None]
    r = np.hypot(dx, dy); np.fill_diagonal(r, np.inf)
    near = r <= h
    with np.errstate(divide="ignore", invalid="ignore"):
        u = ((vx[:, None] - vx[None, :]) * dx + (vy[:, None] - vy[None, :]) * dy) / r
        imp = 0.5 * t.time_step * (1 - r / h) * (t.sigma * u + t.beta * u * u)
    binds = near & (u > 0) & ((np.abs(imp * dx / r) > 2.5) | (np.abs(imp * dy / r) > 2.5))
    assert binds.sum() > 20, binds.sum()


@pytest.mark.parametrize("packed", [False, True])
@pytest.mark.parametrize("name,warm,gamma", [("default1508", 400, 0.0), ("block3000", 150, 0.0), ("goo_rect1508", 300, 0.5)])
def test_branch_free_relax_loop_is_bit_identical_in_the_emulator(built_lib, monkeypatch, name, warm, gamma, packed):
    """SPH_RELAX_BF=1: k_relax walks a row's candidates in order with mask-predicated loads instead of chasing the
    set bits; a masked-off candidate contributes an exact zero, so the default build's bits must come out."""
    defs = ("SPH_RELAX_BF=1",) + (("SPH_PACKED=1", "SPH_PACKED_RELAX=1") if packed else ())
    base = build_emu()
    bf = build_emu(defines=defs, name="libsph_emu_bf%d.so" % packed)
    d0, a0 = run(base, name, warm, 12, gamma, monkeypatch)
    d1, a1 = run(bf, name, warm, 12, gamma, monkeypatch)
    for f in ("x", "y", "v_x", "v_y"):
        assert np.array_equal(a0[f].view("u4"), a1[f].view("u4")), ("after the step", f)


def test_branch_free_relax_loop_redoes_particles_with_coincident_neighbours(built_lib, monkeypatch):
    """The hostile soup of the edge-case test (coincident particles, pile-ups in the corners, particles on the walls)
    through the SPH_RELAX_BF build: the coincident-pair rules (fluid.c:583-588) live in the exact walk the fast loop
    falls back to, so the result must equal the default build's bit for bit -- including rows longer than a mask."""
    from test_gpu_parity import Cuda
    outs = []
    for lib in (build_emu(), build_emu(defines=("SPH_RELAX_BF=1", "SPH_PACKED=1", "SPH_PACKED_RELAX=1"), name="libsph_emu_bf1.so")):
        monkeypatch.setattr(sph_b200, "_lib", sph_b200._bind(C.CDLL(lib)))
        z, t, tank_w, tank_h, h, _ = load_golden("default1508")
        rng = np.random.default_rng(11)
        n = 1200
        st = np.zeros(n, z["w400_state"].dtype)
        st["x"] = rng.uniform(0, tank_w, n); st["y"] = rng.uniform(0, tank_h * 0.3, n)
        st["x"][:200] = st["x"][200:400]; st["y"][:200] = st["y"][200:400]          # coincident pairs
        st["x"][400:560] = rng.uniform(0, 0.4 * h, 160); st["y"][400:560] = rng.uniform(0, 0.4 * h, 160)   # a corner pile: rows > 32
        st["x"][560:600] = 0.0; st["y"][600:640] = 0.0                              # on the walls
        st["v_x"] = rng.uniform(-3, 3, n); st["v_y"] = rng.uniform(-3, 3, n)
        st["x_prev"] = st["x"]; st["y_prev"] = st["y"]
        b = Cuda(tank_w, tank_h, h, n + 64)
        b.set_params(t); b.upload(st); b.step(6)
        outs.append(b.download()[0])
    for f in ("x", "y", "v_x", "v_y"):
        assert np.array_equal(outs[0][f].view("u4"), outs[1][f].view("u4")), f


# the default build's sort since round 2's second half: uid-only scatter, reorder walking the source order, one-barrier
# scan over 1024-cell tiles that skips empty tiles, k_relax inputs staged with cp.async; R2A = what round 2 shipped first
SORT_DEFS = ("SPH_SORT_SRC=1", "SPH_SCAN_FAST=1")
R2A = ("SPH_SORT_SRC=0", "SPH_SCAN_FAST=0", "SPH_ASYNC=0", "SCAN_ITEMS=8")


def run_order(libpath, name, warm, steps, gamma, monkeypatch):
    """like run(), returning the resident ORDER (uids as stored) after each sort as well"""
    monkeypatch.setattr(sph_b200, "_lib", sph_b200._bind(C.CDLL(libpath)))
    z, t, tank_w, tank_h, h, _ = load_golden(name)
    st = z[f"w{warm}_state"]
    c = sph_b200.Context(tank_w, tank_h, h, len(st) + 64)
    c.set_params(as_sph(t)); c.set_viscosity_stabilisation(gamma); c.upload(st)
    c.step(steps)
    c.advect(); c.sort()
    a, ua = c.download(order=sph_b200.ORDER_CELL, include_halo=True)
    c.density(); c.relax(); c.sort()
    b, ub = c.download(order=sph_b200.ORDER_CELL, include_halo=True)
    s = c.status()
    return a, ua, b, ub, tuple(getattr(s, f) for f, _ in s._fields_)


@pytest.mark.parametrize("name,warm,gamma,items", [("default1508", 400, 0.0, 4), ("block3000", 150, 0.0, 4), ("goo_rect1508", 300, 0.5, 4),
                                                   ("block3000", 150, 0.0, 8)])
def test_source_order_sort_is_bit_identical_in_the_emulator(built_lib, monkeypatch, name, warm, gamma, items):
    """SPH_SORT_SRC=1 + SPH_SCAN_FAST=1: the sort's kernels restructured (uid-only scatter, reorder walking the source
    order, one-barrier scan that skips empty tiles).  A sort has exactly one correct result -- cells in key order,
    ascending uid inside a cell -- so order, payload and every counter must equal the default build's.
    (items = 4: scan tiles of 1024 cells, so that these small tanks have several tiles and empty ones among them.)"""
    base = build_emu(defines=R2A, name="libsph_emu_r2a.so")
    var = build_emu(defines=SORT_DEFS + ("SCAN_ITEMS=%d" % items,), name="libsph_emu_sortsrc%d.so" % items)
    r0 = run_order(base, name, warm, 12, gamma, monkeypatch)
    r1 = run_order(var, name, warm, 12, gamma, monkeypatch)
    for k in (0, 2):
        for f in ("x", "y", "v_x", "v_y", "x_prev", "y_prev"):
            if f in r0[k].dtype.names:
                assert np.array_equal(r0[k][f].view("u4"), r1[k][f].view("u4")), (k, f)
    assert np.array_equal(r0[1], r1[1]) and np.array_equal(r0[3], r1[3])
    assert r0[4] == r1[4]


def test_source_order_sort_on_the_hostile_soup(built_lib, monkeypatch):
    """coincident particles, a corner pile (cells of > 32 entries: the rank loop matters), particles on the walls"""
    from test_gpu_parity import Cuda
    outs = []
    for lib in (build_emu(defines=R2A, name="libsph_emu_r2a.so"), build_emu()):
        monkeypatch.setattr(sph_b200, "_lib", sph_b200._bind(C.CDLL(lib)))
        z, t, tank_w, tank_h, h, _ = load_golden("default1508")
        rng = np.random.default_rng(11)
        n = 1200
        st = np.zeros(n, z["w400_state"].dtype)
        st["x"] = rng.uniform(0, tank_w, n); st["y"] = rng.uniform(0, tank_h * 0.3, n)
        st["x"][:200] = st["x"][200:400]; st["y"][:200] = st["y"][200:400]
        st["x"][400:560] = rng.uniform(0, 0.4 * h, 160); st["y"][400:560] = rng.uniform(0, 0.4 * h, 160)
        st["x"][560:600] = 0.0; st["y"][600:640] = 0.0
        st["v_x"] = rng.uniform(-3, 3, n); st["v_y"] = rng.uniform(-3, 3, n)
        st["x_prev"] = st["x"]; st["y_prev"] = st["y"]
        b = Cuda(tank_w, tank_h, h, n + 64)
        b.set_params(t); b.upload(st); b.step(6)
        outs.append(b.download(order=sph_b200.ORDER_CELL))
    assert np.array_equal(outs[0][1], outs[1][1])
    for f in ("x", "y", "v_x", "v_y"):
        assert np.array_equal(outs[0][0][f].view("u4"), outs[1][0][f].view("u4")), f


@pytest.mark.parametrize("trip", [2])
@pytest.mark.parametrize("name,warm,gamma", [("block3000", 150, 0.0), ("goo_rect1508", 300, 0.5)])
def test_relax_walk_with_the_rare_pairs_out_of_line_is_bit_identical_in_the_emulator(built_lib, monkeypatch, name, warm, gamma, trip):
    """SPH_RELAX_RARE=1: the mask walk's pair physics without the coincident-pair rules; flagged particles are redone."""
    base = build_emu()
    var = build_emu(defines=("SPH_RELAX_RARE=1", "SPH_RELAX_TRIP=%d" % trip), name="libsph_emu_rare%d.so" % trip)
    d0, a0 = run(base, name, warm, 12, gamma, monkeypatch)
    d1, a1 = run(var, name, warm, 12, gamma, monkeypatch)
    for f in ("x", "y", "v_x", "v_y"):
        assert np.array_equal(a0[f].view("u4"), a1[f].view("u4")), ("after the step", f)


def test_relax_walk_with_the_rare_pairs_out_of_line_redoes_coincident_particles(built_lib, monkeypatch):
    """the hostile soup again: 200 coincident pairs, a corner pile with rows longer than a mask, particles on the walls"""
    from test_gpu_parity import Cuda
    outs = []
    for lib in (build_emu(), build_emu(defines=("SPH_RELAX_RARE=1", "SPH_RELAX_TRIP=2"), name="libsph_emu_rare2.so")):
        monkeypatch.setattr(sph_b200, "_lib", sph_b200._bind(C.CDLL(lib)))
        z, t, tank_w, tank_h, h, _ = load_golden("default1508")
        rng = np.random.default_rng(11)
        n = 1200
        st = np.zeros(n, z["w400_state"].dtype)
        st["x"] = rng.uniform(0, tank_w, n); st["y"] = rng.uniform(0, tank_h * 0.3, n)
        st["x"][:200] = st["x"][200:400]; st["y"][:200] = st["y"][200:400]
        st["x"][400:560] = rng.uniform(0, 0.4 * h, 160); st["y"][400:560] = rng.uniform(0, 0.4 * h, 160)
        st["x"][560:600] = 0.0; st["y"][600:640] = 0.0
        st["v_x"] = rng.uniform(-3, 3, n); st["v_y"] = rng.uniform(-3, 3, n)
        st["x_prev"] = st["x"]; st["y_prev"] = st["y"]
        b = Cuda(tank_w, tank_h, h, n + 64)
        b.set_params(t); b.upload(st); b.step(6)
        outs.append(b.download()[0])
    for f in ("x", "y", "v_x", "v_y"):
        assert np.array_equal(outs[0][f].view("u4"), outs[1][f].view("u4")), f


ALL_R2B = ("SPH_PAIRMASK=1", "SPH_RELAX_RARE=1")


@pytest.mark.parametrize("defs,tag", [(ALL_R2B, "r2b")])
@pytest.mark.parametrize("name,warm,gamma", [("block3000", 150, 0.0), ("goo_rect1508", 300, 0.5)])
def test_masked_pair_trips_are_bit_identical_in_the_emulator(built_lib, monkeypatch, name, warm, gamma, defs, tag):
    """SPH_PAIRMASK=1: a row's odd last candidate rides in a masked pair trip (k_advect, k_coupling, k_density) instead of
    a scalar left-over loop.  The masked slot adds an exact zero, so densities, masks (through the relaxation they
    drive), positions and velocities must equal the default build's -- alone and together with the round's other
    restructurings (r2b)."""
    base = build_emu()
    var = build_emu(defines=defs, name="libsph_emu_%s.so" % tag)
    d0, a0 = run(base, name, warm, 12, gamma, monkeypatch)
    d1, a1 = run(var, name, warm, 12, gamma, monkeypatch)
    for f in ("density", "density_near", "x", "y"):
        assert np.array_equal(d0[f].view("u4"), d1[f].view("u4")), ("density stage", f)
    for f in ("x", "y", "v_x", "v_y"):
        assert np.array_equal(a0[f].view("u4"), a1[f].view("u4")), ("after the step", f)


def test_all_round2b_variants_on_the_hostile_soup(built_lib, monkeypatch):
    """coincident pairs, a corner pile, particles on the walls, through every restructured loop at once; capacity
    exactly the particle count, so the masked slot of the very last range reads the padding"""
    from test_gpu_parity import Cuda
    outs = []
    for lib in (build_emu(defines=R2A, name="libsph_emu_r2a.so"), build_emu(defines=ALL_R2B, name="libsph_emu_r2b.so")):
        monkeypatch.setattr(sph_b200, "_lib", sph_b200._bind(C.CDLL(lib)))
        z, t, tank_w, tank_h, h, _ = load_golden("default1508")
        rng = np.random.default_rng(11)
        n = 1200
        st = np.zeros(n, z["w400_state"].dtype)
        st["x"] = rng.uniform(0, tank_w, n); st["y"] = rng.uniform(0, tank_h * 0.3, n)
        st["x"][:200] = st["x"][200:400]; st["y"][:200] = st["y"][200:400]
        st["x"][400:560] = rng.uniform(0, 0.4 * h, 160); st["y"][400:560] = rng.uniform(0, 0.4 * h, 160)
        st["x"][560:600] = 0.0; st["y"][600:640] = 0.0
        st["x"][640:660] = tank_w - 0.001; st["y"][640:660] = tank_h - 0.001      # the very last cell of the table
        st["v_x"] = rng.uniform(-3, 3, n); st["v_y"] = rng.uniform(-3, 3, n)
        st["x_prev"] = st["x"]; st["y_prev"] = st["y"]
        b = Cuda(tank_w, tank_h, h, n)
        b.set_params(t); b.upload(st); b.step(6)
        outs.append(b.download(order=sph_b200.ORDER_CELL))
    assert np.array_equal(outs[0][1], outs[1][1])
    for f in ("x", "y", "v_x", "v_y"):
        assert np.array_equal(outs[0][0][f].view("u4"), outs[1][0][f].view("u4")), f


@pytest.mark.parametrize("name,warm,gamma", [("block3000", 150, 0.0), ("goo_rect1508", 300, 0.5)])
def test_staged_inputs_and_deferred_slot_are_bit_identical_in_the_emulator(built_lib, monkeypatch, name, warm, gamma):
    """SPH_ASYNC=7 + SPH_DEFER=1: per-particle inputs staged through shared memory one particle ahead (cp.async on the
    GPU, a plain copy here) and the arrival-slot store deferred by one particle: the same values reach the same
    arithmetic, so every bit and the resident order must equal the default build's."""
    base = build_emu(defines=R2A, name="libsph_emu_r2a.so")
    var = build_emu(defines=("SPH_ASYNC=7", "SPH_DEFER=1"), name="libsph_emu_async.so")
    r0 = run_order(base, name, warm, 12, gamma, monkeypatch)
    r1 = run_order(var, name, warm, 12, gamma, monkeypatch)
    for k in (0, 2):
        for f in ("x", "y", "v_x", "v_y"):
            assert np.array_equal(r0[k][f].view("u4"), r1[k][f].view("u4")), (k, f)
    assert np.array_equal(r0[1], r1[1]) and np.array_equal(r0[3], r1[3])
    assert r0[4] == r1[4]


R2B_FINAL = ("SPH_ADVECT_PV4=1",)


@pytest.mark.parametrize("name,warm,gamma", [("block3000", 150, 0.0)])
def test_interleaved_candidate_records_are_bit_identical_in_the_emulator(built_lib, monkeypatch, name, warm, gamma):
    """SPH_ADVECT_PV4=1 on top of the source-order sort and the staged k_relax inputs: k_advect reads (x, y, vx, vy)
    records written by sort 2's reorder.  Also through a save / restore (the records are rebuilt)."""
    base = build_emu()
    var = build_emu(defines=R2B_FINAL, name="libsph_emu_pv4.so")
    r0 = run_order(base, name, warm, 12, gamma, monkeypatch)
    r1 = run_order(var, name, warm, 12, gamma, monkeypatch)
    for k in (0, 2):
        for f in ("x", "y", "v_x", "v_y"):
            assert np.array_equal(r0[k][f].view("u4"), r1[k][f].view("u4")), (k, f)
    assert np.array_equal(r0[1], r1[1]) and np.array_equal(r0[3], r1[3])
    # save, run on, restore, run again: same result as the first time
    monkeypatch.setattr(sph_b200, "_lib", sph_b200._bind(C.CDLL(var)))
    z, t, tank_w, tank_h, h, _ = load_golden(name)
    st = z[f"w{warm}_state"]
    c = sph_b200.Context(tank_w, tank_h, h, len(st) + 64)
    c.set_params(as_sph(t)); c.set_viscosity_stabilisation(gamma); c.upload(st)
    c.step(3); c.state_save(); c.step(5)
    a, _ = c.download()
    c.state_restore(); c.step(5)
    b, _ = c.download()
    for f in ("x", "y", "v_x", "v_y"):
        assert np.array_equal(a[f].view("u4"), b[f].view("u4")), f


@pytest.mark.parametrize("name,warm,gamma", [("goo_rect1508", 300, 0.5)])
def test_rows_from_the_sort_key_are_bit_identical_in_the_emulator(built_lib, monkeypatch, name, warm, gamma):
    """SPH_KEYROWS=1 (with the source-order sort and the staged k_relax inputs): candidate rows from the cell key"""
    base = build_emu()
    var = build_emu(defines=("SPH_KEYROWS=1",), name="libsph_emu_keyrows.so")
    r0 = run_order(base, name, warm, 12, gamma, monkeypatch)
    r1 = run_order(var, name, warm, 12, gamma, monkeypatch)
    for k in (0, 2):
        for f in ("x", "y", "v_x", "v_y"):
            assert np.array_equal(r0[k][f].view("u4"), r1[k][f].view("u4")), (k, f)
    assert np.array_equal(r0[1], r1[1]) and np.array_equal(r0[3], r1[3])


def test_row_prefetch_build_is_bit_identical_in_the_emulator(built_lib, monkeypatch):
    """SPH_PREFETCH=7 (+ SPH_ASYNC=7): prefetches change no value; the build must run and agree"""
    base = build_emu()
    var = build_emu(defines=("SPH_ASYNC=7", "SPH_PREFETCH=7"), name="libsph_emu_pf.so")
    r0 = run_order(base, "block3000", 150, 12, 0.0, monkeypatch)
    r1 = run_order(var, "block3000", 150, 12, 0.0, monkeypatch)
    for k in (0, 2):
        for f in ("x", "y", "v_x", "v_y"):
            assert np.array_equal(r0[k][f].view("u4"), r1[k][f].view("u4")), (k, f)
