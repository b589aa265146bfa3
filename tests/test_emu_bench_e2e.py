"""bench.py's end-to-end leg (SingleGpu.e2e: synchronous frames, then frames pipelined through
sph_run_frame_async / sph_coords_wait, with the cross-check that decides which figure is reported) run on the
kernel-source emulator, so that the code the driver runs at round end has been executed before it reaches a GPU.
Timing means nothing here; the protocol, the bookkeeping and the fallback do."""
import ctypes as C
import types

import numpy as np
import pytest

import sph_b200
from emu.build_emu import build as build_emu


@pytest.fixture()
def emu_bench(monkeypatch):
    import torch
    monkeypatch.setattr(sph_b200, "_lib", sph_b200._bind(C.CDLL(build_emu())))
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    import bench
    prob = sph_b200.make_problem(3000, tank_w=15.0 * float(np.sqrt(3000 / 750.0)), water_frac=0.5)
    t = sph_b200.default_params(prob["h"], prob["tank_w"], prob["tank_h"])
    sim = bench.SingleGpu(sph_b200, prob, t, types.SimpleNamespace(cuda_stream=None), None)
    sim.init_lattice()
    sim.run(20)
    return bench, sim, torch


def test_pipelined_e2e_is_reported_when_it_matches_the_synchronous_feed(emu_bench):
    bench, sim, torch = emu_bench
    out = sim.e2e(5, torch.zeros(16, dtype=torch.uint8))
    assert out["pipelined"] is True and "pipelined_error" not in out
    assert out["steps"] == 20 and out["seconds"] > 0 and out["sync_seconds"] > 0
    assert out["d2h_per_step"] == sim.ctx.status().n_local and out["h2d_per_step"] == 16


def test_e2e_falls_back_to_the_synchronous_protocol(emu_bench, monkeypatch):
    bench, sim, torch = emu_bench

    def broken(*a, **k):
        raise sph_b200.SphError("sph_run_frame_async -> 3: simulated failure")
    monkeypatch.setattr(sim.ctx, "run_frame_async", broken)
    out = sim.e2e(3, torch.zeros(16, dtype=torch.uint8))
    assert out["pipelined"] is False and "simulated failure" in out["pipelined_error"]
    assert out["steps"] == 12 and out["seconds"] > 0
