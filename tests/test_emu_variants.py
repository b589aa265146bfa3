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
