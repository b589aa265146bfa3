#!/usr/bin/env python
"""BASELINE.json config 4 as a timed workload: 4 M particles, dam-break block, mover sphere on the render rank's
autopilot path (renderer.c:513-531, step in simulation units: sph_host_mover_autopilot_ex) meeting the collapsing
water, presets cycled a -> b -> x -> y every 64 frames (controls.c:344-401; 256 steps per preset), every frame through
the public frame call with host buffers (sph_run_frame_async: parameter block in, int16 coordinates out).
Prints one JSON line: particle-steps/s per preset phase and overall, launches per frame, the reference's capacities.
    python scripts/bench_cfg4.py [--particles 4000000] [--cycles 1]"""
import argparse
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import numpy as np
    import torch
    import sph_b200
    ap = argparse.ArgumentParser()
    ap.add_argument("--particles", type=int, default=4_000_000)
    ap.add_argument("--frames-per-preset", type=int, default=64)
    ap.add_argument("--cycles", type=int, default=1)
    a = ap.parse_args()
    prob = sph_b200.make_problem(a.particles, tank_w=15.0 * float(np.sqrt(a.particles / 750.0)), water_frac=0.5)
    ts = sph_b200.default_params(prob["h"], prob["tank_w"], prob["tank_h"])
    stream = torch.cuda.Stream()
    b = sph_b200.Context(prob["tank_w"], prob["tank_h"], prob["h"], prob["n_global"] + 4096, stream=stream.cuda_stream)
    L = sph_b200._host()
    gl_x, direction = C.c_float(ts.mover_width / prob["tank_w"] + 0.002), C.c_int(-1)
    dx_gl = 0.01 * 15.0 / prob["tank_w"]
    L.sph_host_mover_autopilot_ex(C.byref(ts), prob["tank_w"], prob["tank_h"], C.byref(gl_x), C.byref(direction), 0.0)
    b.set_params(ts)
    n0 = b.init_lattice(prob)
    bufs = [torch.empty(2 * (n0 + 4096), dtype=torch.int16).pin_memory().numpy() for _ in range(2)]
    phases, prev, f = [], None, 0
    for cyc in range(a.cycles):
        for preset in "abxy":
            torch.cuda.synchronize()
            l0, t0 = b.launches, time.perf_counter()
            for _ in range(a.frames_per_preset):
                L.sph_host_mover_autopilot_ex(C.byref(ts), prob["tank_w"], prob["tank_h"], C.byref(gl_x), C.byref(direction), dx_gl)
                L.sph_host_preset(C.byref(ts), preset.encode())
                ticket = b.run_frame_async(ts, 4, bufs[f % 2])        # frame f goes in ...
                if prev is not None:
                    b.coords_wait(prev)                               # ... before frame f-1 is collected
                prev = ticket
                f += 1
            b.coords_wait(prev); prev = None
            torch.cuda.synchronize()
            secs = time.perf_counter() - t0
            st = b.status()
            phases.append({"preset": preset, "particle_steps_per_s": n0 * 4 * a.frames_per_preset / secs, "ms_per_step": 1e3 * secs / (4 * a.frames_per_preset),
                           "launches_per_frame": (b.launches - l0) / a.frames_per_preset, "max_bucket": st.max_bucket,
                           "bucket_overflow": st.bucket_overflow, "neighbor_overflow": st.neighbor_overflow,
                           "mean_neighbours": 2.0 * b.L.sph_get_pairs(b.h, None, 0) / n0})
    out, u = b.download()
    ok = bool(np.array_equal(u, np.arange(n0, dtype=u.dtype)) and np.all(np.isfinite(out["x"])) and np.all(np.abs(out["v_x"]) <= 5.0))
    total_steps = 4 * a.frames_per_preset * len(phases)
    total_s = sum(4 * a.frames_per_preset * p["ms_per_step"] * 1e-3 for p in phases)
    print(json.dumps({"workload": f"BASELINE config 4: {n0} particles, mover on the autopilot path, presets a/b/x/y x {a.frames_per_preset} frames, "
                                  "through sph_run_frame_async with host buffers (e2e)",
                      "value": n0 * total_steps / total_s, "unit": "particle-steps/s", "phases": phases, "state_ok": ok,
                      "h2d_bytes_per_frame": 64, "d2h_bytes_per_frame": 4 * n0}))


if __name__ == "__main__":
    main()
