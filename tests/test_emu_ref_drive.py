"""The reference's UNMODIFIED start_simulation() driving the reference-named entry points (sph_b200/host/ref_api.c:
arming, lazy attach, host mirror) without a GPU: the GPU driver binary of tests/test_gpu_ref_drive.py with the
kernel-source emulator library (tests/emu) preloaded in front of libsph_b200.so.  Frames must equal, bit for bit,
those of the handle API on the same emulator.  Host-logic regression net only; the claim itself is the GPU test."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import sph_b200
from emu.build_emu import build as build_emu
from oracle.oracle import lattice, make_problem
from test_ref_drive import GPU_DRIVE, HOT, bindings, pack, read_drive


@pytest.mark.skipif(not os.path.exists(GPU_DRIVE), reason="oracle/_ref not built")
@pytest.mark.parametrize("mirror_every,stab", [(None, None), ("4", None), (None, "0.5,0.0")])
def test_unmodified_reference_driver_on_the_emulated_library(built_lib, tmp_path, monkeypatch, mirror_every, stab):
    emu = build_emu()
    out = str(tmp_path / "emu.bin")
    env = dict(os.environ, LD_PRELOAD=emu)
    if mirror_every:
        env["SPH_REF_MIRROR_EVERY"] = mirror_every
    if stab:
        env["SPH_VISC_STAB"] = stab                     # the reference's entry points have no call for it
    r = subprocess.run([GPU_DRIVE, "--frames", "6", "--out", out], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, (r.stdout[-300:], r.stderr[-800:])
    b = bindings(r.stdout)
    assert b["start_simulation"].endswith("libref_driver.so")
    assert all(b[k].endswith("libsph_emu.so") for k in HOT), b
    n, w, h, first, frames = read_drive(out)
    prob = make_problem(1500)
    a, uid = lattice(prob)
    monkeypatch.setattr(sph_b200, "_lib", sph_b200._bind(C.CDLL(emu)))

    def sph(t):
        o = sph_b200.Tunable(); C.memmove(C.byref(o), C.byref(t), 64); return o

    ctx = sph_b200.Context(w, h, float(first.smoothing_radius), n)
    if stab:
        ctx.set_viscosity_stabilisation(*[float(v) for v in stab.split(",")])
    ctx.set_params(sph(first))
    ctx.upload(a, uid)
    coords = np.zeros(2 * n, "i2")
    for k, (blk, got) in enumerate(frames):
        assert ctx.run_frame(sph(blk), 4, coords) == n
        s, _ = ctx.download()
        assert np.array_equal(pack(s["x"], s["y"], w, h), got), k


def test_reference_call_order_on_the_emulated_library(built_lib, monkeypatch):
    """tests/test_ref_api.py's GPU test (the reference's call sequence through the reference-named entry points
    equals sph_step bit for bit) with the emulator library in place of libsph_b200.so."""
    import test_ref_api
    emu = build_emu()
    monkeypatch.setattr(sph_b200, "_lib", sph_b200._bind(C.CDLL(emu)))
    test_ref_api.test_reference_call_order_reproduces_sph_step(emu)


def test_reference_named_attach_refuses_more_than_one_rank(built_lib):
    """start/finishHaloExchange and transferOOBParticles are single-rank no-ops; a host that announces more than one
    compute rank must be told so, not handed isolated slabs."""
    L = C.CDLL(build_emu())
    L.sph_ref_last_error.restype = C.c_char_p
    L.sph_ref_set_rank(1, 3)
    try:
        assert L.sph_ref_attach(None, None, None, None, 0) != 0
        assert b"ONE compute rank" in L.sph_ref_last_error()
    finally:
        L.sph_ref_set_rank(0, 1)
