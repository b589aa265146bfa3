"""The gather (Jacobi) oracle against the reference's golden vectors: binning and neighbour sets
bit-exact, one step within the stated tolerance, long-run statistics within tolerance.  This is the
CPU leg of the same assertions the CUDA path has to pass in test_gpu_parity.py."""
import numpy as np
import pytest

import parity_checks as pc
from oracle.oracle import GatherOracle, lattice, make_problem

CASES = [("default1508", 100), ("default1508", 400), ("goo_rect1508", 300), ("block3000", 150),
         ("zerog1508", 200), ("gas1508", 200)]


def make(tank_w, tank_h, h, capacity):
    return GatherOracle(tank_w, tank_h, h, capacity)


@pytest.mark.parametrize("name,warm", CASES)
def test_binning_and_neighbour_sets_bit_exact(name, warm):
    pc.check_binning_and_neighbours_exact(make, name, warm)


@pytest.mark.parametrize("name,warm", CASES)
def test_one_step_within_tolerance(name, warm):
    pc.check_stages_vs_reference(make, name, warm)


@pytest.mark.parametrize("name,warm", CASES)
def test_density_on_reference_positions(name, warm):
    pc.check_density_exact_positions(make, name, warm)


@pytest.mark.parametrize("name,warm", CASES[:2])
def test_ten_steps_bounded(name, warm):
    pc.check_ten_steps_bounded(make, name, warm)


LONGRUN = {"default1508": dict(n_request=1500), "block3000": dict(n_request=3000, tank_w=21.2, water_frac=0.5),
           "zerog1508": dict(n_request=1500), "gas1508": dict(n_request=1500), "goo_rect1508": dict(n_request=1500)}


@pytest.mark.parametrize("name", ["default1508", "block3000", "zerog1508", "gas1508"])
def test_long_run_statistics(name):
    a, _ = lattice(make_problem(**LONGRUN[name]))
    pc.check_long_run_statistics(make, name, a)


def test_goo_preset_needs_the_stabilised_viscosity_gather():
    """Why the stabilised viscosity gather is ON BY DEFAULT for blocks with dt*sigma >= 0.5.  With the "goo" preset (sigma 100, beta 10:
    dt sigma = 0.83 per pair) the viscosity impulses summed from frozen velocities overshoot and the fluid
    never settles (kinetic energy 3.5 against the reference's 2e-4, heap three times too high), while the
    reference's in-place sweep damps every pair without overshoot.  The symmetric damping proposed in
    oracle/sph_oracle.c (orc_g_set_viscosity_stabilisation, gamma 0.5) settles it within the reference's own
    order sensitivity and leaves the stable presets bit-identical.  Both libraries engage it by default for such
    blocks (sph_create / orc_g_create: gamma 0.5, threshold 0.5); this test pins the gap of the plain gather
    (gamma = 0), the fix, and that the DEFAULT is the fix."""
    a, _ = lattice(make_problem(**LONGRUN["goo_rect1508"]))
    with pytest.raises(AssertionError):
        pc.check_long_run_statistics(make, "goo_rect1508", a, prepare=lambda g: g.set_viscosity_stabilisation(0.0))
    from common import GOO_STABILISED_WIDEN
    pc.check_long_run_statistics(make, "goo_rect1508", a, widen=GOO_STABILISED_WIDEN)       # the default
    pc.check_long_run_statistics(make, "goo_rect1508", a, prepare=lambda g: g.set_viscosity_stabilisation(0.5),
                                 widen=GOO_STABILISED_WIDEN)
    # ... and from a lattice with ONE coordinate moved by one ulp, which lands in the other packing of the heap
    b = a.copy(); b["x"][700] = np.nextafter(b["x"][700], np.float32(100))
    pc.check_long_run_statistics(make, "goo_rect1508", b, prepare=lambda g: g.set_viscosity_stabilisation(0.5),
                                 widen=GOO_STABILISED_WIDEN)


def test_stabilised_viscosity_leaves_stable_presets_bit_identical():
    from common import load_golden
    for name in ("default1508", "block3000", "gas1508"):
        z, t, tank_w, tank_h, h, _ = load_golden(name)
        a, _ = lattice(make_problem(**LONGRUN[name]))
        out = []
        for gamma in (0.0, 0.5):
            g = make(tank_w, tank_h, h, len(a) + 64)
            g.set_params(t); g.set_viscosity_stabilisation(gamma); g.upload(a); g.step(300)
            out.append(g.download()[0])
        for f in ("x", "y", "v_x", "v_y"):
            assert np.array_equal(out[0][f].view("u4"), out[1][f].view("u4")), (name, f)


def test_preset_cycle_with_moving_mover_tracks_the_reference():
    """BASELINE.json config 4 in small, on the CPU oracles: dam-break block, mover on the render rank's autopilot
    path (renderer.c:513-531), presets a -> b -> x -> y for 256 steps each (controls.c:344-401), every block
    landing in the last sub-step of its frame.  At the end of each phase the gather (with the stabilised
    viscosity, which only ever engages in the y phase) is compared with the sequential restatement of the
    reference: these are transients of a chaotic system, so the bars are loose; the plain gather is shown to
    lose the y phase."""
    import ctypes as C
    import sph_b200
    from oracle.oracle import SeqOracle, default_tunable
    n_req, phase_steps = 3000, 256
    prob = make_problem(n_req, tank_w=15.0 * np.sqrt(n_req / 750.0), water_frac=0.5)
    a, uid = lattice(prob)
    L = sph_b200._host()

    def run(kind):
        t = default_tunable(prob["h"], prob["tank_w"], prob["tank_h"])
        ts = sph_b200.Tunable(); C.memmove(C.byref(ts), C.byref(t), 64)
        if kind == "seq":
            o = SeqOracle(len(a) + 64, prob["tank_w"], prob["tank_h"], t); o.load(a)
        else:
            o = make(prob["tank_w"], prob["tank_h"], prob["h"], len(a) + 64); o.set_params(t)
            o.set_viscosity_stabilisation(0.5 if kind == "stabilised" else 0.0)
            o.upload(a, uid)
        gl_x, direction = C.c_float(-0.2), C.c_int(1)
        out = {}
        for preset in "abxy":
            for _ in range(phase_steps // 4):
                L.sph_host_mover_autopilot(C.byref(ts), prob["tank_w"], prob["tank_h"], C.byref(gl_x), C.byref(direction))
                L.sph_host_preset(C.byref(ts), preset.encode())
                C.memmove(C.byref(t), C.byref(ts), 64)
                if kind == "seq":
                    o.step(); o.step(); o.step(); o.step(queued=t)
                else:
                    o.step(3); o.queue_params(t); o.step(1)
            st = o.store() if kind == "seq" else o.download()[0]
            d = pc.density_of(make, prob["tank_w"], prob["tank_h"], prob["h"], t, st)
            out[preset] = np.array([d.mean(), st["y"].mean()])
        return out

    ref, plain, stab = run("seq"), run("plain"), run("stabilised")
    for preset in "abx":                         # the stable presets: both gathers track the reference
        for got in (plain, stab):
            assert np.all(np.abs(got[preset] / ref[preset] - 1) <= 0.05), (preset, got[preset], ref[preset])
    assert np.all(np.abs(stab["y"] / ref["y"] - 1) <= [0.06, 0.2]), (stab["y"], ref["y"])      # measured: +2.5 %, +10 %
    assert plain["y"][1] > 2 * ref["y"][1], "the plain gather is expected to lose the goo phase (known gap)"
