#!/usr/bin/env python3
"""Where a frame's time goes beyond its four steps (1 M particles, one GPU): host clock around `frames` frames of
  A  sph_step(4)                                           four captured steps, nothing else
  B  sph_step(3); sph_queue_params; sph_step(1)            + the parameter block landing in the last step
  C  B + sph_pack_coords_async / sph_coords_wait           = sph_run_frame_async, pipelined (what bench.py's e2e times)
  D  A + sph_pack_coords_async / sph_coords_wait           the coordinate feed without the parameter block
no L2 flush anywhere (as in the e2e section of bench.py)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import sph_b200 as sph

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 300
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
preroll = int(sys.argv[3]) if len(sys.argv) > 3 else 1005
prob = sph.make_problem(n, tank_w=bench.problem_dims(n, 0.5), water_frac=0.5, nranks=1)
t = sph.default_params(prob["h"], prob["tank_w"], prob["tank_h"], "x")
t.mover_center_x = 0.75 * prob["tank_w"]
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    cap = prob["n_global"] + 4096
    c = sph.Context(prob["tank_w"], prob["tank_h"], prob["h"], cap, stream=stream.cuda_stream)
    c.set_params(t)
    c.init_lattice(prob)
    c.step(preroll)
    c.state_save()
    bufs = [torch.empty(2 * cap, dtype=torch.int16).pin_memory().numpy() for _ in range(2)]

    def timed(fn, warm=4):
        c.state_restore()
        for f in range(warm):
            fn(f, True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for f in range(frames):
            fn(f, False)
        fn(-1, False)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / frames * 1e6

    def A(f, w):
        if f >= 0: c.step(4)

    def B(f, w):
        if f >= 0:
            c.step(3); c.queue_params(t); c.step(1)

    tick = {}

    def feed(f, stepper):
        if f >= 0:
            stepper()
            tick[f % 2] = c.pack_coords_async(bufs[f % 2])
            if (f - 1) % 2 in tick and f > 0:
                c.coords_wait(tick.pop((f - 1) % 2))
        else:
            for k in list(tick):
                c.coords_wait(tick.pop(k))

    def C_(f, w):
        feed(f, lambda: (c.step(3), c.queue_params(t), c.step(1)))
        if w and f == 3: feed(-1, None)

    def D(f, w):
        feed(f, lambda: c.step(4))
        if w and f == 3: feed(-1, None)

    for name, fn in (("A four steps", A), ("B + queued parameter block", B), ("C + coordinate feed (= run_frame_async)", C_),
                     ("D four steps + coordinate feed", D), ("A again", A)):
        print(f"{name:45s} {timed(fn):8.1f} us per frame", flush=True)
