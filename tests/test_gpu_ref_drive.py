"""The reference's UNMODIFIED start_simulation() (fluid.c:71-395) running on the B200 path: the driver is
the reference's own object code (oracle/_ref/libref_driver.so), every hot-path function it calls is
resolved to libsph_b200.so by link order, the library attaches itself and mirrors the
host AoS the driver packs its frames from.  The frames a renderer would receive must be, bit for bit,
the ones the handle API produces for the same parameter blocks.  Run on the B200 box: pytest -m gpu."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle.oracle import lattice, make_problem
from test_ref_drive import GPU_DRIVE, HOT, bindings, pack, read_drive

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not os.path.exists(GPU_DRIVE), reason="oracle/_ref not built")
@pytest.mark.parametrize("mirror_every", [None, "4"])
def test_unmodified_reference_driver_runs_on_the_gpu_path(built_lib, tmp_path, mirror_every):
    import sph_b200
    out = str(tmp_path / "gpu.bin")
    env = dict(os.environ)
    if mirror_every:
        env["SPH_REF_MIRROR_EVERY"] = mirror_every       # the driver packs every 4th step (fluid.c:105)
    r = subprocess.run([GPU_DRIVE, "--frames", "6", "--out", out], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, (r.stdout[-300:], r.stderr[-800:])
    b = bindings(r.stdout)
    assert b["start_simulation"].endswith("libref_driver.so")
    assert all(b[k].endswith("libsph_b200.so") for k in HOT), b
    n, w, h, first, frames = read_drive(out)
    prob = make_problem(1500)
    assert (n, w, h) == (prob["n_global"], prob["tank_w"], prob["tank_h"])
    a, uid = lattice(prob)

    def sph(t):
        o = sph_b200.Tunable(); C.memmove(C.byref(o), C.byref(t), 64); return o

    ctx = sph_b200.Context(w, h, float(first.smoothing_radius), n)
    ctx.set_params(sph(first))
    ctx.upload(a, uid)
    coords = np.zeros(2 * n, "i2")
    for k, (blk, got) in enumerate(frames):
        assert ctx.run_frame(sph(blk), 4, coords) == n
        s, _ = ctx.download()                              # uid order == the driver's pointer order
        assert np.array_equal(pack(s["x"], s["y"], w, h), got), k
        # and the device-side feed is the same set of pixels (device order is the sort's)
        assert np.array_equal(np.sort(coords.reshape(n, 2).view("i4").ravel()), np.sort(got.view("i4").ravel())), k
