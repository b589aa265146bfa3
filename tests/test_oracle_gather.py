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


def test_long_run_statistics_default():
    a, _ = lattice(make_problem(1500))
    pc.check_long_run_statistics(make, "default1508", a)
