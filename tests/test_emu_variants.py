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


# ---------------------------------------------------------------------------------------------------------------
# Round 2, second half.  The default build changed (profiles/r2_variants.md): source-order sort (uid-only scatter,
# reorder with coalesced payload loads, four entries per trip), one-barrier scan over 1024-cell tiles that skips empty
# tiles, k_relax inputs staged one particle ahead (cp.async on the GPU, a plain copy here), k_relax's mask walk with the
# coincident-pair rules out of line.  R2A = the build round 2 shipped first: every one of those switched back.
# A sort has exactly one correct result (cells in key order, ascending uid inside a cell) and none of the other changes
# touches an operation or its order, so resident order, payload bits and counters must be equal.
# (Three emulator builds for the whole section: each costs ~10 s of the CPU suite.)
# ---------------------------------------------------------------------------------------------------------------
from emu.build_emu import VARIANTS  # noqa: E402  (one list of defines per library name: conftest prebuilds them side by side)

R2A = VARIANTS["libsph_emu_r2a.so"]
# other sizes of the same machinery: 2048-cell tiles, two entries per sort trip, two neighbours per relax trip
SIZES = VARIANTS["libsph_emu_sizes.so"]
# everything that was measured and rejected, switched on together: masked pair trips, staged inputs in all three
# gathers, deferred slot store, (x, y, vx, vy) candidate records, rows from the sort key, L1 prefetch of the next rows
REJECTED = VARIANTS["libsph_emu_rejected.so"]


def run_order(libpath, name, warm, steps, gamma, monkeypatch):
    """like run(), returning the resident ORDER (uids as stored, ghosts included) after each sort as well"""
    monkeypatch.setattr(sph_b200, "_lib", sph_b200._bind(C.CDLL(libpath)))
    z, t, tank_w, tank_h, h, _ = load_golden(name)
    st = z[f"w{warm}_state"]
    c = sph_b200.Context(tank_w, tank_h, h, len(st) + 64)
    c.set_params(as_sph(t)); c.set_viscosity_stabilisation(gamma); c.upload(st)
    c.step(steps)
    c.advect(); c.sort()
    a, ua = c.download(order=sph_b200.ORDER_CELL, include_halo=True)
    c.density()
    d, _ = c.download(order=sph_b200.ORDER_CELL, include_halo=True)
    c.relax(); c.sort()
    b, ub = c.download(order=sph_b200.ORDER_CELL, include_halo=True)
    s = c.status()
    return a, ua, b, ub, tuple(getattr(s, f) for f, _ in s._fields_), d


def same_run(r0, r1):
    for k in (0, 2):
        for f in ("x", "y", "v_x", "v_y"):
            assert np.array_equal(r0[k][f].view("u4"), r1[k][f].view("u4")), (k, f)
    for f in ("density", "density_near"):
        assert np.array_equal(r0[5][f].view("u4"), r1[5][f].view("u4")), f
    assert np.array_equal(r0[1], r1[1]) and np.array_equal(r0[3], r1[3])
    assert r0[4] == r1[4]


def soup(lib, monkeypatch, cap_extra=0):
    """coincident pairs, a corner pile (cells of > 32 entries: rank loops, rows longer than a mask), particles on the
    walls and in the very last cell of the table; capacity = particle count, so the last range ends at the padding"""
    from test_gpu_parity import Cuda
    monkeypatch.setattr(sph_b200, "_lib", sph_b200._bind(C.CDLL(lib)))
    z, t, tank_w, tank_h, h, _ = load_golden("default1508")
    rng = np.random.default_rng(11)
    n = 1200
    st = np.zeros(n, z["w400_state"].dtype)
    st["x"] = rng.uniform(0, tank_w, n); st["y"] = rng.uniform(0, tank_h * 0.3, n)
    st["x"][:200] = st["x"][200:400]; st["y"][:200] = st["y"][200:400]
    st["x"][400:560] = rng.uniform(0, 0.4 * h, 160); st["y"][400:560] = rng.uniform(0, 0.4 * h, 160)
    st["x"][560:600] = 0.0; st["y"][600:640] = 0.0
    st["x"][640:660] = tank_w - 0.001; st["y"][640:660] = tank_h - 0.001
    st["v_x"] = rng.uniform(-3, 3, n); st["v_y"] = rng.uniform(-3, 3, n)
    st["x_prev"] = st["x"]; st["y_prev"] = st["y"]
    b = Cuda(tank_w, tank_h, h, n + cap_extra)
    b.set_params(t); b.upload(st); b.step(6)
    return b.download(order=sph_b200.ORDER_CELL)


def same_soup(a, b):
    assert np.array_equal(a[1], b[1])
    for f in ("x", "y", "v_x", "v_y"):
        assert np.array_equal(a[0][f].view("u4"), b[0][f].view("u4")), f


@pytest.mark.parametrize("name,warm,gamma", [("default1508", 400, 0.0), ("block3000", 150, 0.0), ("goo_rect1508", 300, 0.5)])
def test_default_build_equals_the_build_round_2_shipped_first(built_lib, monkeypatch, name, warm, gamma):
    """(block3000 has 6240 sort cells = 7 scan tiles of the default build, the upper ones empty: the tile-skipping path)"""
    same_run(run_order(build_emu(defines=R2A, name="libsph_emu_r2a.so"), name, warm, 12, gamma, monkeypatch),
             run_order(build_emu(), name, warm, 12, gamma, monkeypatch))


def test_default_build_equals_the_build_round_2_shipped_first_on_the_hostile_soup(built_lib, monkeypatch):
    """the coincident-pair rules live in the exact walk the straight-line loop falls back to"""
    same_soup(soup(build_emu(defines=R2A, name="libsph_emu_r2a.so"), monkeypatch), soup(build_emu(), monkeypatch))


def test_tile_and_trip_sizes_do_not_matter(built_lib, monkeypatch):
    var = build_emu(defines=SIZES, name="libsph_emu_sizes.so")
    same_run(run_order(build_emu(), "block3000", 150, 12, 0.0, monkeypatch), run_order(var, "block3000", 150, 12, 0.0, monkeypatch))
    same_soup(soup(build_emu(), monkeypatch), soup(var, monkeypatch))


@pytest.mark.parametrize("name,warm,gamma", [("block3000", 150, 0.0), ("goo_rect1508", 300, 0.5)])
def test_measured_and_rejected_variants_still_agree(built_lib, monkeypatch, name, warm, gamma):
    """the A/B flags stay in the source with their numbers (profiles/r2_variants.md); they must keep building and
    keep producing the default build's bits"""
    var = build_emu(defines=REJECTED, name="libsph_emu_rejected.so")
    same_run(run_order(build_emu(), name, warm, 12, gamma, monkeypatch), run_order(var, name, warm, 12, gamma, monkeypatch))


def test_rejected_variants_on_the_soup_and_through_a_restore(built_lib, monkeypatch):
    var = build_emu(defines=REJECTED, name="libsph_emu_rejected.so")
    same_soup(soup(build_emu(), monkeypatch), soup(var, monkeypatch))
    # SPH_ADVECT_PV4 keeps (x, y, vx, vy) records beside the arrays: rebuilt by sph_state_restore
    monkeypatch.setattr(sph_b200, "_lib", sph_b200._bind(C.CDLL(var)))
    z, t, tank_w, tank_h, h, _ = load_golden("block3000")
    st = z["w150_state"]
    c = sph_b200.Context(tank_w, tank_h, h, len(st) + 64)
    c.set_params(as_sph(t)); c.upload(st)
    c.step(3); c.state_save(); c.step(5)
    a, _ = c.download()
    c.state_restore(); c.step(5)
    b, _ = c.download()
    for f in ("x", "y", "v_x", "v_y"):
        assert np.array_equal(a[f].view("u4"), b[f].view("u4")), f
