"""The kernel SOURCE and the C-ABI layer exercised without a GPU.

tests/emu compiles sph_b200/csrc/sph_capi.cu (and the kernels it includes) unchanged with g++ against a fake
CUDA runtime (fibers for the threads of a block, closures for graph capture) into tests/emu/_build/libsph_emu.so.
This module points the ctypes binding at that library and runs the bodies of the GPU parity tests on it, so a
change to a kernel or to the stage/buffer/graph logic is checked against the reference's golden vectors and the
gather oracle in this container.  It is a development aid and a regression net for the host logic: races,
memory-model and scheduling behaviour are invisible to it, MUFU approximations and FFMA contraction are not
reproduced, and nothing here counts as parity of the product -- that is tests/test_gpu_*.py on the B200.
The product library never loads the emulator (test_product_library_has_no_emulator_in_it).
"""
import ctypes as C
import subprocess

import numpy as np
import pytest

import parity_checks as pc
import sph_b200
import test_gpu_parity as gpu
import test_zy_gpu_stabilised_and_feed as late
from emu.build_emu import build as build_emu


@pytest.fixture(autouse=True)
def emulated_library(monkeypatch):
    monkeypatch.setattr(sph_b200, "_lib", sph_b200._bind(C.CDLL(build_emu())))
    yield


# the GPU tests' own bodies (parametrisation travels with the functions, the module-level gpu mark does not)
from test_gpu_parity import (  # noqa: E402,F401
    test_binning_and_neighbour_sets_bit_exact,
    test_one_step_within_tolerance_of_reference,
    test_density_on_reference_positions,
    test_rounding_level_agreement_with_gather_oracle,
    test_ten_steps_bounded,
    test_graph_step_equals_staged_and_is_deterministic,
    test_queued_params_land_between_predict_and_relax,
    test_edge_cases,
    test_pack_coords_matches_reference_formula,
    test_device_side_lattice_equals_host_lattice,
    test_mover_autopilot_and_preset_cycle,
    test_long_run_statistics_default,
)
from test_zy_gpu_stabilised_and_feed import (  # noqa: E402,F401
    test_stabilised_viscosity_rounding_level_agreement_with_gather_oracle,
    test_stabilised_viscosity_engages_on_goo_and_leaves_stable_presets_bit_identical,
    test_stabilisation_threshold_selects_the_pass_per_parameter_block,
    test_long_run_statistics_goo_with_stabilised_viscosity,
    test_asynchronous_coordinate_feed_equals_the_synchronous_one,
)


def test_full_size_properties_in_small(built_lib):
    gpu.test_full_size_properties(built_lib, 20_000)


def test_config4_properties_in_small(built_lib):
    prob, b, n0, coords, per_frame = late.run_config4(6000, 32, 4)
    late.check_config4(prob, b, n0, coords, per_frame, 4)


def test_product_library_has_no_emulator_in_it(built_lib):
    """libsph_b200.so is nvcc output with device code and real CUDA runtime calls; the emulator's symbols
    exist only in tests/emu/_build/libsph_emu.so."""
    syms = subprocess.run(["nm", "-D", "--defined-only", built_lib], capture_output=True, text=True, check=True).stdout
    assert "emu" not in syms.lower()
    elf = subprocess.run(["cuobjdump", "-lelf", built_lib], capture_output=True, text=True).stdout
    assert "sm_100a" in elf


def test_no_device_is_an_error_not_a_fallback(built_lib, monkeypatch):
    """The C-ABI layer's own no-GPU path (sph_create: "no CUDA device: sph_b200 has no CPU path"), driven by the
    fake runtime reporting zero devices."""
    monkeypatch.setenv("SPH_EMU_DEVICES", "0")
    with pytest.raises(sph_b200.SphError, match="no CPU path"):
        sph_b200.Context(10.0, 5.0, 0.5, 64)


@pytest.mark.parametrize("order", ["reverse", "shuffle"])
def test_results_do_not_depend_on_block_or_thread_order(built_lib, monkeypatch, order):
    """The counting sort takes its arrival slots from atomics and the messages their entries from atomic cursors;
    k_reorder then orders every cell by uid.  Run the blocks and the threads inside them in reverse and in
    shuffled order: positions, velocities, densities and the coordinate feed must not change by a bit."""
    from common import load_golden
    z, t, tank_w, tank_h, h, _ = load_golden("block3000")
    st = z["w150_state"]

    def run():
        c = sph_b200.Context(tank_w, tank_h, h, len(st) + 64)
        c.set_viscosity_stabilisation(0.5)
        c.set_params(gpu.as_sph(t)); c.upload(st); c.step(12)
        c.advect(); c.sort(); c.density()
        d, _ = c.download()
        c.relax(); c.sort()
        a, u = c.download(order=sph_b200.ORDER_CELL)
        return d, a, u, c.pack_coords()

    base = run()
    monkeypatch.setenv("SPH_EMU_ORDER", order)
    other = run()
    for f in ("density", "density_near"):
        assert np.array_equal(base[0][f].view("u4"), other[0][f].view("u4")), f
    for f in ("x", "y", "v_x", "v_y"):
        assert np.array_equal(base[1][f].view("u4"), other[1][f].view("u4")), f
    assert np.array_equal(base[2], other[2])                      # the device order itself is canonical
    assert np.array_equal(base[3], other[3])


def test_state_snapshot_restores_the_same_future(built_lib):
    pc.check_state_snapshot(sph_b200.Context, gpu.as_sph)


def test_clamped_impulses_take_the_exact_rows(built_lib):
    pc.check_clamped_impulses(gpu.mk, gpu.make_oracle)


def test_status_of_a_fresh_upload_counts_the_forward_lists_exactly(built_lib):
    """sph_status.neighbor_overflow is documented as exact for the current sorted state.  The exact count used to run only
    once a density pass had raised its conservative running count, so a state that had just been uploaded reported 0
    whatever it held (found by tests/fuzz/fuzz_soup.py: 2315 particles with forward lists above the reference's 400
    entries, hash.c:188,223, reported as none).  Emulator only: the fix is host logic around a kernel the GPU suite runs."""
    from oracle.oracle import PARTICLE, GatherOracle, default_tunable
    rng = np.random.default_rng(7)
    n, tank_w = 3000, 15.0
    tank_h = tank_w * 9.0 / 16.0
    t = default_tunable(0.580948, tank_w, tank_h, "x")
    h = t.smoothing_radius
    a = np.zeros(n, PARTICLE)
    a["x"] = np.clip(0.5 * tank_w + rng.normal(0, 0.5 * h, n), 0, tank_w - 0.002)       # one cluster: ~1000 particles within h
    a["y"] = np.clip(0.5 * tank_h + rng.normal(0, 0.5 * h, n), 0, tank_h - 0.002)
    a["x_prev"], a["y_prev"], a["id"] = a["x"], a["y"], np.arange(n)
    b = sph_b200.Context(tank_w, tank_h, h, n + 64)
    b.set_params(gpu.as_sph(t)); b.upload(a)
    o = GatherOracle(tank_w, tank_h, h, n + 64)
    o.set_params(t); o.upload(a)
    (fu, fc), (gu, gc) = b.forward_counts(), o.forward_counts()
    assert np.array_equal(fc[np.argsort(fu)], gc[np.argsort(gu)])
    want = int((gc > 400).sum())
    assert want > 100
    assert b.status().neighbor_overflow == want
    b.step(1)                                                    # and it stays exact once the running count is primed
    fu, fc = b.forward_counts()
    assert b.status().neighbor_overflow == int((fc > 400).sum())
