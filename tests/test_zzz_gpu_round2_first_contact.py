"""GPU tests of what round 2 added and that therefore met the hardware last: they sort after EVERY other GPU file, so a
surprise here cannot mask tests that were already green on the B200 (`pytest -x` stops at the first failure).
The same bodies run on the kernel-source emulator in tests/test_emu_*.py."""
import os
import subprocess
import sys

import numpy as np
import pytest

import parity_checks as pc
from test_gpu_parity import as_sph, make_oracle, mk

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_state_snapshot_restores_the_same_future(built_lib):
    import sph_b200
    pc.check_state_snapshot(sph_b200.Context, as_sph)


def test_clamped_impulses_take_the_exact_rows(built_lib):
    """SPH_TRIM (default since round 2): rows in which the impulse clamp can bind are redone with the exact body."""
    pc.check_clamped_impulses(mk, make_oracle)


def test_reference_caps_bite_and_the_cuda_path_reports_instead_of_dropping(built_lib):
    import test_oracle_caps as caps
    from test_gpu_parity import Cuda
    caps.check(*caps.run_teleporting_mover(lambda *a: Cuda(*a)))


def test_config4_bench_script_runs(built_lib):
    """scripts/bench_cfg4.py (BASELINE config 4 as a timed workload) at a reduced size."""
    import json
    r = subprocess.run([sys.executable, os.path.join(HERE, "..", "scripts", "bench_cfg4.py"), "--particles", "200000", "--frames-per-preset", "8"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert d["state_ok"] and d["value"] > 0 and [p["preset"] for p in d["phases"]] == list("abxy")
    assert d["phases"][3]["launches_per_frame"] > d["phases"][2]["launches_per_frame"]      # the goo phase runs the extra pass


def ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("world,mode,n_req,steps", [(2, "onex:1:count", 40000, 240), (2, "onex:2:time", 40000, 240),
                                                   (2, "onex:4:time", 120000, 120), (4, "onex:2:time", 120000, 120),
                                                   (8, "onex:2:time", 400000, 80)])
def test_slabs_with_one_exchange_every_few_steps_match_single_gpu_bit_for_bit(tmp_path, built_lib, world, mode, n_req, steps):
    """One exchange per step as a mode of the context, neighbours meeting every 1 / 2 / 4 steps, edges balanced on
    measured slab time: peer-memory exchange over NVLink, bit-identical to one GPU."""
    if ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    base = str(tmp_path / "slab")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29519", os.path.join(HERE, "slab_gpu_worker.py"), base, str(n_req), str(steps), "p2p", mode]
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
