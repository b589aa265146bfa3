"""bench.py's main() -- the command the driver runs on the B200 at round end -- executed from argv to JSON line in
this container: tests/emu/run_bench_emulated.py points the binding at the kernel-source emulator and replaces
torch.cuda by stand-ins.  The figures are meaningless; the contract keys and the code path are what is checked."""
import json
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
CONTRACT = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "roofline", "clocks", "e2e", "gpu_launches")


@pytest.mark.parametrize("extra,gather", [([], "library default"), (["--preset", "y"], "library default: stabilised"),
                                          (["--preset", "y", "--visc-stab", "0"], "forced")])
def test_bench_main_runs_end_to_end_on_the_emulator(built_lib, extra, gather):
    cmd = [sys.executable, os.path.join(HERE, "emu", "run_bench_emulated.py"), "--particles", "3000", "--steps", "4",
           "--warmup", "3", "--preroll", "8", "--no-cpu-baseline"] + extra
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "stdout must carry exactly one JSON line"
    d = json.loads(lines[0])
    assert all(k in d for k in CONTRACT), [k for k in CONTRACT if k not in d]
    assert d["metric"] == "particle-steps/sec" and d["n_gpus"] == 1 and d["steps"] == 4 and d["dtype"] == "f32"
    assert d["config"]["viscosity_gather"].startswith(gather)
    assert set(d["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
    assert d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert "pipelined" in d["e2e"]["protocol"] and d["e2e"]["synchronous_value"] > 0 and "pipelined_error" not in d["e2e"]
    # 9 launches per step (10 with the stabilised gather's extra pass) in the timed region, which repeats the block
    # of --steps steps until enough device time has been measured and reports the median block with its spread
    tr = d["config"]["timed_region"]
    assert tr["blocks"] >= 3 and tr["steps_per_block"] == 4 and len(tr["block_ms_min_p10_median_p90_max"]) == 5
    assert d["gpu_launches"] == tr["blocks"] * 4 * (10 if gather.endswith("stabilised") else 9)
    assert d["config"]["integrity"]["nobody_lost_or_duplicated"] and d["config"]["integrity"]["capacity_overflow"] == 0
