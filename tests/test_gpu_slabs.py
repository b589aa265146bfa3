"""Multi-GPU slab path on real GPUs (NCCL).  Needs >= 2 devices; launched as a torchrun job by the
test itself so that `pytest -m gpu` on a 1-GPU box simply skips it."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(ngpus() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("world,transport,n_req,steps", [
    (2, "p2p", 40000, 240), (2, "collective", 40000, 240), (4, "p2p", 40000, 240), (8, "p2p", 40000, 240),

    # big enough that every kernel runs its full grid: the exchange kernel waits for the neighbour inside the
    # kernel, which deadlocks unless its grid is fully co-resident (regression test)
    (2, "p2p", 600000, 40)])
    # (steps < 0 selects the goo preset with the stabilised viscosity gather, k_coupling on ghosts too: that case lives in
    #  tests/test_zy_gpu_stabilised_and_feed.py with the other first-hardware-run tests)
def test_two_gpu_slabs_match_single_gpu_bit_for_bit(tmp_path, built_lib, world, transport, n_req, steps):
    if ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    goo = steps < 0
    steps = abs(steps)
    base = str(tmp_path / "slab")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(HERE, "slab_gpu_worker.py"), base, str(n_req), str(steps), transport] + (["goo_stabilised"] if goo else [])
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    parts = [np.load(f"{base}.rank{r}.npz") for r in range(world)]
    state = np.concatenate([p["state"] for p in parts]); uid = np.concatenate([p["uid"] for p in parts])
    assert all(int(p["overflow"].sum()) == 0 for p in parts)
    single = np.load(f"{base}.single.npz")
    assert np.array_equal(np.sort(uid), single["uid"])
    order = np.argsort(uid)
    for f in ("x", "y", "v_x", "v_y"):
        assert np.array_equal(state[f][order].view("u4"), single["state"][f].view("u4")), f
    counts = [len(p["uid"]) for p in parts]
    assert max(counts) - min(counts) <= 0.3 * len(uid) / world, counts
