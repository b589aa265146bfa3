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
import os
import subprocess

import numpy as np
import pytest

import sph_b200
import test_gpu_parity as gpu
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
    test_stabilised_viscosity_rounding_level_agreement_with_gather_oracle,
    test_stabilised_viscosity_engages_on_goo_and_leaves_stable_presets_bit_identical,
    test_stabilisation_threshold_selects_the_pass_per_parameter_block,
    test_asynchronous_coordinate_feed_equals_the_synchronous_one,
)


def test_long_run_statistics_goo_with_stabilised_viscosity(built_lib):
    gpu.test_long_run_statistics_goo_with_stabilised_viscosity(built_lib)


def test_full_size_properties_in_small(built_lib):
    gpu.test_full_size_properties(built_lib, 20_000)


def test_config4_properties_in_small(built_lib):
    prob, b, n0, coords, per_frame = gpu.run_config4(6000, 32, 4)
    gpu.check_config4(prob, b, n0, coords, per_frame, 4)


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
