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
    """KNOWN GAP of the shipped gather.  With the "goo" preset (sigma 100, beta 10:
    dt sigma = 0.83 per pair) the viscosity impulses summed from frozen velocities overshoot and the fluid
    never settles (kinetic energy 3.5 against the reference's 2e-4, heap three times too high), while the
    reference's in-place sweep damps every pair without overshoot.  The symmetric damping proposed in
    oracle/sph_oracle.c (orc_g_set_viscosity_stabilisation, gamma 0.5) settles it within the reference's own
    order sensitivity and leaves the stable presets bit-identical.  The CUDA path implements the plain
    gather only; this test pins both facts so that neither can change unnoticed."""
    a, _ = lattice(make_problem(**LONGRUN["goo_rect1508"]))
    with pytest.raises(AssertionError):
        pc.check_long_run_statistics(make, "goo_rect1508", a)
    pc.check_long_run_statistics(make, "goo_rect1508", a, prepare=lambda g: g.set_viscosity_stabilisation(0.5))


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
