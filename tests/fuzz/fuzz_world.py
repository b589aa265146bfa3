"""TEST INFRASTRUCTURE -- the reference's WHOLE PROGRAM under a randomised "user", K compute ranks against one.

oracle/_ref/sph_ref_world_gpu is the reference's unmodified main(), renderer.c (load balancer, parameter scatter,
coordinate gather) and controls.c, headless (oracle/ref_build/ref_world.c, render_stubs.c), with every hot-path call of
its compute ranks bound to the library -- here the kernel-source emulator, preloaded in front of it.  Per seed the
headless user presses a random sequence of keys (fluid presets x / a / b, remove_partition / add_partition,
controls.c:344-455) while the mover is dragged across the tank; the frames the renderer draws with K = 2-4 compute ranks
(slab messages through the host's MPI_Sendrecv, sph_exchange_via_host) must hold the pixels it draws with ONE rank.
    python tests/fuzz/fuzz_world.py FIRST_SEED COUNT"""
import os
import random
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(os.path.dirname(HERE)), os.path.dirname(HERE)]
from emu.build_emu import build  # noqa: E402
from test_ref_drive import WORLD_GPU, read_world  # noqa: E402


def run(seed, tmp):
    rng = random.Random(seed)
    K = rng.choice([2, 3, 4])
    frames = rng.randint(12, 26)
    keys, active = [], K
    for f in range(2, frames - 1):
        if rng.random() < 0.35:
            k = rng.choice(["x", "a", "b", "remove", "add"])
            if k == "remove" and active <= 1:
                continue
            if k == "add" and active >= K:
                continue
            active += {"remove": -1, "add": 1}.get(k, 0)
            keys.append(f"{f}:{k}")
    env = dict(os.environ, LD_PRELOAD=build(), SPH_RENDER_SCRIPT=" ".join(keys))
    if rng.random() < 0.5:
        env["SPH_REF_MIRROR_EVERY"] = "4"
    outs = {}
    for k in (1, K):
        outs[k] = os.path.join(tmp, f"w{seed}_{k}.bin")
        r = subprocess.run([WORLD_GPU, "--ranks", str(k), "--frames", str(frames), "--out", outs[k]], capture_output=True, text=True,
                           timeout=900, env=env)
        assert r.returncode == 0, (k, r.stdout[-300:], r.stderr[-600:])
        assert "sph_ref_api:" not in r.stderr, (k, r.stderr[-600:])
    _, w, h, one = read_world(outs[1])
    _, wk, hk, many = read_world(outs[K])
    assert (w, h) == (wk, hk) and len(one) == len(many) == frames
    for f in range(frames):
        assert len(many[f][2]) == len(one[f][2]), ("particle count", f, len(many[f][2]), len(one[f][2]), keys)
        a = np.sort(many[f][2].copy().view("i8").ravel()); b = np.sort(one[f][2].copy().view("i8").ravel())
        assert np.array_equal(a, b), ("pixels differ", f, int((a != b).sum()), keys)
    for p in outs.values():
        os.remove(p)
    return f"ok K={K} frames={frames} keys={' '.join(keys)}"


if __name__ == "__main__":
    s0, n = int(sys.argv[1]), int(sys.argv[2])
    failed = 0
    with tempfile.TemporaryDirectory() as tmp:
        for seed in range(s0, s0 + n):
            try:
                print("seed", seed, run(seed, tmp), flush=True)
            except AssertionError as e:
                failed += 1
                print("seed", seed, "FAIL", str(e)[:600], flush=True)
    sys.exit(1 if failed else 0)
