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


def _free_port():
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,extra", [(2, ["--no-parity-check"]), (3, ["--force-cfg3", "--cfg3-particles", "2000"])])
def test_bench_main_runs_as_several_ranks_on_emulated_slabs(built_lib, world, extra):
    """The command the driver's scaling run launches -- torch.distributed.run, one rank per GPU, bench.py --gpus N -- from
    argv to JSON line on emulated slabs over gloo: the parity pre-check (N slabs against one, 40 k particles), the
    exchange period, time-balanced edges, the reductions over ranks, the per-slab table, integrity counters and (forced
    at N = 3) the config-3 section that --gpus 8 adds.  Found on the way: the roofline's dominant "kernel" was taken over
    all stage times, "exchange" included -- a KeyError whenever the collective transport's send/recv was the longest stage."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(HERE, "emu", "run_bench_emulated.py"), "--gpus", str(world),
           "--particles", "3000", "--steps", "4", "--warmup", "3", "--preroll", "8", "--no-cpu-baseline", "--min-timed-ms", "50"] + extra
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, "exactly one JSON line, printed by rank 0"
    d = json.loads(lines[0])
    assert all(k in d for k in CONTRACT), [k for k in CONTRACT if k not in d]
    assert d["n_gpus"] == world and d["scaling"] == "weak" and d["value"] > 0 and d["e2e"]["value"] > 0
    c = d["config"]
    if "--no-parity-check" not in extra:
        assert c["slab_parity"]["result"] == "bit-identical" and c["slab_parity"]["slabs"] == world
    assert c["integrity"]["nobody_lost_or_duplicated"] and c["integrity"]["capacity_overflow"] == 0 and c["integrity"]["exchange_timeouts"] == 0
    assert c["exchange_period_steps"] == 2 and c["parallelism"] == f"slab{world}"
    per_slab = [v for k, v in c.items() if k.startswith("per_slab_")][0]
    assert len(per_slab) == world and sum(row[0] for row in per_slab) == c["integrity"]["particles_resident"]
    assert d["roofline"]["kernel"].startswith("k_")
    if "--force-cfg3" in extra:
        c3 = c["cfg3_16m"]
        assert "failed" not in c3 and c3["value"] > 0 and c3["particles_resident"] == c3["particles_created"] and c3["capacity_overflow"] == 0
    else:
        assert "cfg3_16m" not in c


def test_smoke_entry_point_runs_on_the_emulator(built_lib, monkeypatch, capsys):
    """__graft_entry__.smoke() -- the driver's first call on the B200 -- with the binding pointed at the emulator: its
    three comparisons (gather oracle at rounding level, reference restatement and live reference within the one-step
    tolerance) and its status checks execute here before they do there."""
    import ctypes as C

    import sph_b200
    from emu.build_emu import build as build_emu
    sys.path.insert(0, os.path.dirname(HERE))
    import __graft_entry__ as entry
    monkeypatch.setattr(sph_b200, "_lib", sph_b200._bind(C.CDLL(build_emu())))
    entry.smoke()
    out = capsys.readouterr().out
    assert "smoke ok: n=1508" in out


@pytest.mark.parametrize("world", [1, 2])
def test_reference_arm_prints_one_line_from_rank_zero(world):
    """bench.py --impl reference, as the driver launches it (under torch.distributed.run for N > 1: rank 0 alone runs the
    reference's CPU implementation and prints, the others exit 0): contract keys, impl, a cpu_baseline of kind
    "reference" that describes this run, an e2e object with zero copies."""
    ref = os.path.join(HERE, "..", "oracle", "_ref", "sph_ref_run")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref not built")
    bench = os.path.join(HERE, "..", "bench.py")
    tail = ["--impl", "reference", "--gpus", str(world), "--steps", "2", "--warmup", "3", "--particles", "3000", "--preroll", "5"]
    cmd = [sys.executable, bench] + tail if world == 1 else \
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
         "--master-port", str(_free_port()), bench] + tail
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "particle-steps/sec" and d["n_gpus"] == world and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["value"] == d["value"] > 0 and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert f"{3000 * world} particles requested" in d["config"]["workload"]
