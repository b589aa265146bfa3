#!/usr/bin/env python
"""bench.py -- particle-steps/s of the TinySPH compute-rank timestep on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--particles PER_GPU]

Workload (BASELINE.json configs[1]): 2-D dam-break block, 1 M particles per GPU, default fluid
preset 'x', lattice spacing and h of the reference (s0 = 0.2905, h = 0.5809), tank scaled so the
water block (left half, full height) holds N particles; mover sphere (diameter 2/15 W) parked in
the dry half at (0.75 W, 0.35 H).  A "step" is one iteration of the loop at fluid.c:270-372
(dt = 1/120).  Weak scaling: N GPUs -> N x 1 M particles in N x-slabs.

One JSON line on stdout (rank 0).  Keys beyond the base contract:
  roofline      dominant kernel vs the measured HBM peak, algorithmic bytes
  cpu_baseline  the unmodified reference (oracle/_ref/sph_ref_run) on this box's host cores
  e2e           frames through sph_run_frame: parameter block H2D + 4 steps + int16 coords D2H
"""
import argparse
import json
import os
import re
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALG_BYTES = {            # algorithmic bytes per particle per launch
    "advect": 24, "sort": 48 + 8 + 4, "density": 16, "relax": 40, "step": 200,
}
HBM_FALLBACK_GBS = 6650.0


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    return HBM_FALLBACK_GBS, "fallback"


def measured_traffic(kernel, n_particles):
    """DRAM bytes per launch of `kernel` from the last `ncu --set full` capture summarised in
    profiles/traffic.json (scripts/ncu_summary.py); only reported for the particle count it was taken at."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None
    t = json.load(open(p))
    e = t.get(kernel.split("+")[0])
    if not e or abs(e["n_particles"] - n_particles) > 0.02 * n_particles:
        return None
    return e["dram_bytes_per_launch"]


def profiled_limiter(kernel, source="r2b_full.csv", launch_seconds=None, sm_mhz=None, n_particles=None, sms=148):
    """What actually bounds `kernel` according to the last `ncu --set full` capture summarised under profiles/
    (NOT measured in this run): issue-slot utilisation, active lanes per instruction and DRAM throughput.  The
    HBM roofline fraction is small because these gathers are instruction-issue bound.
    `issue_roofline`: the second roof SURVEY.md 8(d) asks for -- the capture's warp-instruction count per launch
    (valid for the particle count and state it was taken at) over THIS run's launch time, against
    SMs x 4 schedulers x the SM clock sampled during this run."""
    import csv
    p = os.path.join(ROOT, "profiles", source)
    try:
        rows = list(csv.reader(open(p)))
        k = kernel.split("+")[0]
        col = rows[0].index(k)
        get = lambda m: float(next(r[col] for r in rows if r[0] == m))
        out = {"issue_active_pct": get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
               "active_lanes_per_instruction": get("smsp__thread_inst_executed_per_inst_executed.ratio"),
               "dram_pct_of_peak": get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
               "source": f"profiles/{source} (ncu capture of this build's kernels, not this run)"}
        try:
            t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(k)
            if launch_seconds and sm_mhz and t and n_particles and abs(t["n_particles"] - n_particles) <= 0.02 * n_particles:
                inst = get("smsp__inst_executed.sum")
                peak = sms * 4 * sm_mhz * 1e6
                out["issue_roofline"] = {"warp_instructions_per_launch": inst, "achieved": inst / launch_seconds, "peak": peak,
                                         "unit": "warp-instructions/s", "frac": inst / launch_seconds / peak,
                                         "note": "instruction count from the committed capture, launch time and SM clock from this run"}
        except Exception:
            pass
        return out
    except Exception:       # a missing or reshaped summary must never cost the bench line
        return None


def problem_dims(n, water_frac):
    import numpy as np
    tank_w = 15.0 * float(np.sqrt(n / (1500.0 * water_frac)))
    return tank_w


class ClockSampler:
    """nvidia-smi clocks during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
        except OSError:
            self.p = None

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        out = self.p.communicate()[0]
        sm, mx, reasons = [], None, set()
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the unmodified reference on the host cores
# ------------------------------------------------------------------------------------------------
def host_mem_available_bytes():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) * 1024
    except OSError:
        pass
    return 64 << 30


def run_reference_cpu(n, water_frac, steps, warmup, ranks=None, budget_s=150.0, shipped=True):
    """Runs the unmodified reference (reference TUs + mini-MPI, oracle/_ref/) on the host cores: `shipped` = built with
    the flags the reference ships (-O3 -ffast-math, makefile:5: sph_ref_run_shipped), else the IEEE -O2 build that
    defines parity (sph_ref_run).  Falls back to the C restatement (oracle port, 1 core) when no binary is there.
    The sample is bounded by the time budget, by the reference's own 32-bit limit (2N*400 in an unsigned at
    fluid.c:203,206 overflows above 5.3 M particles) and by the host's memory (~13 KB of touched capacity per
    particle, fluid.c:156,202-203)."""
    from oracle.oracle import ref_binary
    cores = os.cpu_count() or 1
    exe = ref_binary()
    if exe and shipped and os.path.exists(exe + "_shipped"):
        exe = exe + "_shipped"
    elif exe and shipped:
        shipped = False
    if exe:
        ranks = ranks or max(1, min(cores, 64))
        est_rate = 0.5e6 * ranks                       # particle-steps/s, conservative
        by_time = est_rate * budget_s / max(steps + warmup, 1)
        by_mem = 0.5 * host_mem_available_bytes() / 13e3
        n_s = int(min(n, 5_200_000, by_mem, max(20000, by_time)))
        bound = ("the full size" if n_s == n else "the reference's 32-bit capacity limit" if n_s == 5_200_000 else
                 "host memory" if n_s == int(by_mem) else f"the {budget_s:.0f} s time budget")
        tank_w = problem_dims(n_s, water_frac)
        cmd = [exe, "--ranks", str(ranks), "--n", str(n_s), "--tank-w", f"{tank_w:.6f}",
               "--tank-h", f"{tank_w * 9.0 / 16.0:.6f}", "--water-frac", str(water_frac),
               "--mover-x-frac", "0.75", "--steps", str(steps), "--warmup", str(warmup), "--balance", "1"]
        t0 = time.time()
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=budget_s * 4)
        line = [l for l in out.stdout.splitlines() if l.startswith("{")]
        if out.returncode != 0 or not line:
            raise RuntimeError(f"{os.path.basename(exe)} failed rc={out.returncode}: {out.stderr[-400:]}")
        r = json.loads(line[-1])
        return {"value": r["particle_steps_per_s"], "unit": "particle-steps/s", "cores": ranks, "kind": "reference",
                "build": "-O3 -ffast-math, the flags the reference ships (makefile:5)" if shipped else
                         "-O2 IEEE (no fast-math, no FMA contraction): the build that defines parity",
                "sample": f"{r['n_global']} particles x {steps} steps (+{warmup} warm-up) from the lattice, "
                          f"{ranks} compute ranks over mini-MPI, load balancer on, host has {cores} cores, "
                          f"wall {time.time() - t0:.1f}s; size bounded by {bound}"
                          + ("" if n_s == n else f" (requested {n}: the RATE is what is compared, extrapolated)"),
                "ms_per_step": 1e3 * r["seconds"] / max(steps, 1), "n_particles": r["n_global"]}
    # oracle port: single core
    import ctypes as C
    from oracle.oracle import SeqOracle, default_tunable, lattice, make_problem
    n_s = int(min(n, 200000))
    prob = make_problem(n_s, tank_w=problem_dims(n_s, water_frac), water_frac=water_frac)
    a, _ = lattice(prob)
    t = default_tunable(prob["h"], prob["tank_w"], prob["tank_h"]); t.mover_center_x = 0.75 * prob["tank_w"]
    seq = SeqOracle(len(a) + 8, prob["tank_w"], prob["tank_h"], t)
    seq.load(a)
    for _ in range(warmup):
        seq.step()
    t0 = time.time()
    for _ in range(steps):
        seq.step()
    dt = time.time() - t0
    return {"value": len(a) * steps / dt, "unit": "particle-steps/s", "cores": 1, "kind": "port",
            "sample": f"{len(a)} particles x {steps} steps from the lattice, oracle/sph_oracle.c orc_seq_step",
            "ms_per_step": 1e3 * dt / max(steps, 1), "n_particles": len(a)}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n = args.n * args.gpus
    # same initial condition and the same number of steps before the timed region as our arm
    r = run_reference_cpu(n, args.water_frac, args.steps, args.preroll + args.warmup, shipped=True)
    try:    # the parity build beside it, on a shorter pre-roll (same state class, bounded time)
        ri = run_reference_cpu(n, args.water_frac, args.steps, min(args.preroll + args.warmup, 300), budget_s=40.0, shipped=False)
        ieee = {k: ri[k] for k in ("value", "unit", "cores", "kind", "build", "sample")}
    except Exception as e:
        ieee = {"value": None, "sample": f"failed: {e}"}
    line = {
        "impl": "reference", "metric": "particle-steps/sec", "value": r["value"], "unit": "particle-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"2D dam-break block, {n} particles requested ({r['n_particles']} in the bounded sample), "
                               "preset x, CPU reference", "water_frac": args.water_frac},
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "build", "sample")},
        "cpu_baseline_ieee_build": ieee,
        "e2e": {"value": r["value"], "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def main_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import sph_b200

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import faulthandler
    import signal
    faulthandler.register(signal.SIGUSR1, all_threads=True)     # kill -USR1 <pid>: where is it stuck?

    def mark(msg):
        if os.environ.get("SPH_BENCH_TRACE"):
            print(f"[bench rank {rank}] {msg} t={time.time():.1f}", file=sys.stderr, flush=True)
    # stdout carries exactly one JSON line: libraries that chat on fd 1 (NCCL prints its version there)
    # are sent to stderr until the result is printed
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; sph_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    peak, peak_kind = measured_peak()

    n_global_req = args.n * world
    tank_w = problem_dims(n_global_req, args.water_frac)
    prob = sph_b200.make_problem(n_global_req, tank_w=tank_w, water_frac=args.water_frac, nranks=world)
    t = sph_b200.default_params(prob["h"], prob["tank_w"], prob["tank_h"], args.preset)
    t.mover_center_x = 0.75 * prob["tank_w"]
    stream = torch.cuda.Stream()
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    parity = None
    if world > 1 and not args.no_parity_check:
        parity = slab_parity_check(sph_b200, rank, world, stream, args)
        mark("parity check")
    cfg3 = None
    with torch.cuda.stream(stream):
        if world == 1:
            sim = SingleGpu(sph_b200, prob, t, stream, args)
        else:
            from sph_b200.slab import SlabRunner
            sim = SlabRunner(prob, t, rank, world, stream, capacity_factor=2.0, balance_policy=args.balance,
                             halo_width=args.halo_width, exchange_period=args.exchange_period,
                             exchanges_per_step=args.exchanges_per_step, time_proportional=not args.time_fixed_step)
        if args.visc_stab is not None:                  # default: the library's (gamma 0.5 for blocks with dt*sigma >= 0.5)
            sim.ctx.set_viscosity_stabilisation(args.visc_stab)
        mark("created")
        sim.init_lattice()
        mark("lattice")
        sampler = ClockSampler(local_rank)   # nvidia-smi needs ~0.2 s to start: begin before the pre-roll,
        sampler.start()                      # stop after the last measured phase; everything in between is load
        sim.run(args.preroll)
        mark("preroll enqueued")
        sim.run(args.warmup)
        barrier()
        mark("warm")
        launches0 = sim.launches
        if world > 1:
            sim.ctx.exchange_times(reset=True)
        if os.environ.get("SPH_PROFILE"):      # ncu --profile-from-start off: instrument the timed region only
            torch.cuda.profiler.start()
        # ---- timed region: blocks of K steps, L2 flushed between steps, device time per step from CUDA events on the
        # stream the steps are launched on; every block bracketed by barrier + synchronize on both sides.  The block is
        # repeated until at least 0.5 s of device time has been measured (a single block of the driver's 20 steps is
        # 5 ms), and the MEDIAN block is reported with the spread.
        # Every block times the SAME steps: the state after the pre-roll is snapshotted in device memory and restored
        # before each block (the dam-break keeps compressing -- 22 neighbours per particle after 1000 steps, 60 after
        # 4000 -- so blocks timed one after the other would not be the same work).
        sim.state_save()

        def timed_block():
            sim.state_restore()
            barrier()
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
            for k in range(args.steps):
                flush_buf.zero_()
                ev[k][0].record(stream)
                sim.run(1)
                ev[k][1].record(stream)
            barrier()
            return float(sum(a.elapsed_time(b) for a, b in ev))

        wall0 = time.perf_counter()
        block_ms = [timed_block()]
        first_ms = block_ms[0]
        if world > 1:
            first = torch.tensor([first_ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(first, op=dist.ReduceOp.MAX)
            first_ms = float(first.item())
        repeats = int(min(args.max_repeats, max(3, -(-args.min_timed_ms // max(first_ms, 1e-3)))))
        for _ in range(repeats - 1):
            block_ms.append(timed_block())
        mark("timed")
        wall = time.perf_counter() - wall0
        launches = sim.launches - launches0
        xt, n_meet = sim.ctx.exchange_times(reset=True) if world > 1 else (None, 0)
        if os.environ.get("SPH_PROFILE"):
            torch.cuda.profiler.stop()
        if world > 1:
            blocks = torch.tensor(block_ms, device="cuda", dtype=torch.float64)
            dist.all_reduce(blocks, op=dist.ReduceOp.MAX)          # every block: the slowest rank
            block_ms = [float(x) for x in blocks.tolist()]
        block_ms = sorted(block_ms)
        total_ms = statistics.median(block_ms)
        # ---- the same K steps back to back, state L2-resident (informational)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sim.state_restore()
        barrier()
        e0.record(stream); sim.run(args.steps); e1.record(stream)
        barrier()
        b2b_ms = e0.elapsed_time(e1)
        # ---- per-stage device times for the roofline (stage API, flushed between steps)
        mark("b2b")
        sim.state_restore()
        stage_ms = sim.stage_times(min(args.steps, 20), flush_buf)
        mark("stages")
        # ---- end to end through the frame call with host buffers: enough frames for ~0.5 s
        # (like the timed region, in blocks that start from the restored state: frames timed one after the other march
        #  into a denser fluid -- 300 frames after the pre-roll a step costs 242 us instead of 208, scripts/diag_e2e.py
        #  -- and would not be the work `value` is quoted on)
        e2e_frames = int(min(64, max(8, 0.5 * args.min_timed_ms / (4.0 * max(total_ms / args.steps, 1e-3)))))
        e2e_block = max(8, args.steps // 4)
        sim.state_restore()
        e2e = sim.e2e(e2e_frames, flush_buf, block=e2e_block)
        mark("e2e")
        clocks = sampler.stop()
        sim.state_restore()
        stats = sim.stats()
        mark("stats")

    if (world == 8 or args.force_cfg3) and world > 1 and not args.no_cfg3:
        # (the main section's simulation stays alive: its names and settings go into the line below)
        try:
            cfg3 = run_cfg3(sph_b200, rank, world, stream, args, flush_buf, barrier)
        except Exception as e:
            cfg3 = {"failed": repr(e)[:300]}
        mark("cfg3")

    per_rank = None
    if world > 1:
        # per-slab picture: population, neighbours, pure compute time (the three gather kernels) per step
        mine = torch.tensor([stats["n_local"], stats["n_halo"], stats["mean_neighbours"],
                             1e3 * (stage_ms["advect"] + stage_ms["density"] + stage_ms["relax"]),
                             1e3 * (stage_ms["sort1"] + stage_ms["sort2"]),
                             xt["send_us"], xt["wait_us"], xt["unpack_us"], n_meet], device="cuda", dtype=torch.float64)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = [[round(float(v), 1) for v in r.tolist()] for r in allr]
        tmax = torch.tensor([b2b_ms, e2e["seconds"], e2e.get("sync_seconds", 0.0)], device="cuda", dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        b2b_ms, e2e["seconds"], sync_s = [float(x) for x in tmax.tolist()]
        if "sync_seconds" in e2e:
            e2e["sync_seconds"] = sync_s
        okall = torch.tensor([1.0 if e2e.get("pipelined") else 0.0], device="cuda", dtype=torch.float64)
        dist.all_reduce(okall, op=dist.ReduceOp.MIN)
        e2e["pipelined"] = bool(okall.item() > 0)
        cnt = torch.tensor([stats["n_local"], launches, stats["capacity_overflow"], stats["msg_overflow"],
                            stats["exchange_timeouts"]], device="cuda", dtype=torch.float64)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        n_global = int(cnt[0].item()); launches = int(cnt[1].item())
        integrity = {"particles_resident": n_global, "particles_created": prob["n_global"],
                     "nobody_lost_or_duplicated": n_global == prob["n_global"],
                     "capacity_overflow": int(cnt[2].item()), "msg_overflow": int(cnt[3].item()),
                     "exchange_timeouts": int(cnt[4].item())}
    else:
        n_global = stats["n_local"]
        integrity = {"particles_resident": n_global, "particles_created": prob["n_global"],
                     "nobody_lost_or_duplicated": n_global == prob["n_global"],
                     "capacity_overflow": stats["capacity_overflow"], "msg_overflow": stats["msg_overflow"], "exchange_timeouts": 0}

    if rank == 0:
        ms_per_step = total_ms / args.steps
        value = n_global * args.steps / (total_ms * 1e-3)
        q = lambda f: block_ms[min(len(block_ms) - 1, int(f * len(block_ms)))]
        # (the dominant KERNEL: "exchange" is the collective transport's send/recv between kernels, empty with peer stores)
        dom = max((k for k in stage_ms if k != "exchange"), key=lambda k: stage_ms[k])
        n_per_launch = stats["n_local"] + stats["n_halo"]
        dom_key = "sort" if dom.startswith("sort") else dom
        achieved = ALG_BYTES[dom_key] * n_per_launch / (stage_ms[dom] * 1e-3) / 1e9
        step_gbs = ALG_BYTES["step"] * n_global / world / (ms_per_step * 1e-3) / 1e9
        line = {
            "metric": "particle-steps/sec", "value": value, "unit": "particle-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": f"2D dam-break block, {n_global} particles ({args.n} requested per GPU), preset {args.preset}, "
                            f"h={prob['h']:.6f}, tank {prob['tank_w']:.1f}x{prob['tank_h']:.1f}, {world} x-slab(s)",
                "viscosity_gather": ("library default: stabilised gather (gamma 0.5) for blocks with dt*sigma >= 0.5 (goo), plain otherwise"
                                     if args.visc_stab is None else f"forced: gamma {args.visc_stab} for every block (0 = plain)"),
                "state": f"{args.preroll} pre-roll steps + {args.warmup} warm-up steps from the lattice, snapshotted in device memory "
                         "and restored before every timed block (and before the stage, e2e and statistics sections)",
                "mean_neighbours_per_particle": stats["mean_neighbours"], "max_bucket": stats["max_bucket"],
                "l2": "flushed between timed steps (256 MiB write, outside the per-step CUDA-event brackets)",
                "timed_region": {"blocks": len(block_ms), "steps_per_block": args.steps, "statistic": "median block",
                                 "block_ms_min_p10_median_p90_max": [round(x, 4) for x in (block_ms[0], q(0.1), total_ms, q(0.9), block_ms[-1])],
                                 "spread_rel": round((q(0.9) - q(0.1)) / total_ms, 4),
                                 "bracket": "barrier + torch.cuda.synchronize() on both sides of every block; max over ranks per block"},
                "integrity": integrity,
                "slab_parity": parity,
                "l2_resident_value": n_global * args.steps / (b2b_ms * 1e-3),
                "l2_resident_ms_per_step": b2b_ms / args.steps,
                "wall_s_timed_region": wall,
                "stage_ms": stage_ms,
                "step_hbm_frac": step_gbs / peak,
                "parallelism": f"slab{world}",
                "exchanges_per_step": getattr(sim, "exchanges", None),
                "exchange_period_steps": getattr(sim, "exchange_period", None),
                "edge_policy": ("particle count, dead band 1/15 (renderer.c:427-477)" if args.balance == "count" else
                                "work estimate per slab (sph_copy_load), dead band 1/40 -- NOT the reference's policy" if args.balance == "cost" else
                                "measured device time of each slab between meetings (sph_copy_work), dead band 1/100 -- NOT the reference's "
                                "policy; the result does not depend on where the edges are") if world > 1 else None,
                "per_slab_[n_local,n_ghost,neighbours,gather_us,sort_us,meet_send_us,meet_wait_us,meet_unpack_us,meetings_in_timed_region]": per_rank,
            },
            "roofline": {"bound": "hbm", "kernel": sim.kernel_name(dom), "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak,
                         "traffic": measured_traffic(sim.kernel_name(dom), n_per_launch), "peak_source": peak_kind,
                         "algorithmic_bytes_per_particle": ALG_BYTES[dom_key],
                         "profiled_limiter": profiled_limiter(sim.kernel_name(dom), launch_seconds=stage_ms[dom] * 1e-3,
                                                              sm_mhz=(clocks or {}).get("sm_mhz"), n_particles=n_per_launch)},
            "clocks": clocks,
            "e2e": {"value": n_global * e2e["steps"] / e2e["seconds"], "unit": "particle-steps/s",
                    "h2d_bytes_per_step": e2e["h2d_per_step"], "d2h_bytes_per_step": e2e["d2h_per_step"],
                    "protocol": ("frames pipelined like the reference's MPI_Isend of its frame (fluid.c:283-287, :354-365): "
                                 "per frame a 64-byte parameter block H2D, 4 steps, int16 (x,y) per particle D2H into pinned "
                                 "host memory, frame f collected after frame f+1 was submitted; one host clock around each block of "
                                 f"{max(8, args.steps // 4)} frames, the state restored before every block (outside the clock) so that "
                                 "the frames are the steps `value` is quoted on; no L2 flush inside (flushed and L2-resident "
                                 "kernel rates: value / config.l2_resident_value)"
                                 if e2e.get("pipelined") else
                                 "per frame: 64-byte parameter block H2D, 4 steps, int16 (x,y) per particle D2H "
                                 "into pinned host memory (fluid.c:293-294, :354-365); L2 flushed before every frame"),
                    **({"synchronous_value": n_global * e2e["steps"] / e2e["sync_seconds"]} if e2e.get("pipelined") else {}),
                    **({"pipelined_error": e2e["pipelined_error"]} if e2e.get("pipelined_error") else {})},
            "gpu_launches": launches,
        }
        if cfg3 is not None:
            line["config"]["cfg3_16m"] = cfg3
        if not args.no_cpu_baseline and world == 1:
            for key, shipped in (("cpu_baseline", True), ("cpu_baseline_ieee_build", False)):
                try:
                    r = run_reference_cpu(args.n, args.water_frac, steps=args.cpu_steps, warmup=args.cpu_warmup, budget_s=20.0,
                                          shipped=shipped)
                    line[key] = {k: r[k] for k in ("value", "unit", "cores", "kind", "build", "sample")}
                except Exception as e:  # the baseline is reported, never fatal
                    line[key] = {"value": None, "unit": "particle-steps/s", "cores": 0, "kind": "reference",
                                 "sample": f"failed: {e}"}
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def slab_parity_check(sph_b200, rank, world, stream, args, n_req=None, steps=40):
    """Correctness evidence inside the scaling run itself: a small dam-break block on the SAME N slabs (peer-memory
    exchange, rebalancer on, the bench's exchange period) against ONE slab on rank 0, compared per uid BIT FOR BIT."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from sph_b200.slab import SlabRunner
    if n_req is None:
        # 40 k particles, or enough for every slab of the block to be twice as wide as the ghost layer of the
        # bench's exchange period (3.5 h per step between exchanges; the block is half of a tank 15 sqrt(n / 750) wide)
        layer = 3.5 * max(args.exchange_period, 1) * 0.58
        n_req = int(max(40000, 750 * (world * 2.0 * layer / 7.5) ** 2))
    prob = sph_b200.make_problem(n_req, tank_w=problem_dims(n_req, args.water_frac), water_frac=args.water_frac, nranks=world)
    t = sph_b200.default_params(prob["h"], prob["tank_w"], prob["tank_h"], args.preset)
    t.mover_center_x = 0.4 * prob["tank_w"]                 # the mover inside the water, across slab edges
    out = {"particles": prob["n_global"], "steps": steps, "slabs": world}
    ok_local, a, uid, bad = 1, None, None, 0
    try:
        with torch.cuda.stream(stream):
            sim = SlabRunner(prob, t, rank, world, stream, capacity_factor=3.0, balance_policy=args.balance,
                             halo_width=args.halo_width, exchange_period=args.exchange_period,
                             exchanges_per_step=args.exchanges_per_step, time_proportional=not args.time_fixed_step)
            sim.init_lattice()
            sim.run(steps)
            a, uid = sim.ctx.download()
            st = sim.ctx.status()
            bad = st.capacity_overflow + st.msg_overflow + st.exchange_timeouts
        torch.cuda.synchronize()
        del sim
    except Exception as e:     # evidence, not a gate -- but every rank must still take part in the collectives below
        ok_local, out["result"] = 0, f"check failed to run on rank {rank}: {e!r}"[:300]
    try:
        parts = [None] * world
        dist.all_gather_object(parts, (a, uid, bad, ok_local, out.get("result")))
        if not all(p[3] for p in parts):
            out["result"] = next(p[4] for p in parts if not p[3])
        elif rank == 0:
            state = np.concatenate([p[0] for p in parts]); uids = np.concatenate([p[1] for p in parts])
            p1 = sph_b200.make_problem(n_req, tank_w=problem_dims(n_req, args.water_frac), water_frac=args.water_frac)
            with torch.cuda.stream(stream):
                one = sph_b200.Context(p1["tank_w"], p1["tank_h"], p1["h"], p1["n_global"] + 64, stream=stream.cuda_stream)
                one.set_params(t); one.init_lattice(p1)
                # the slab runner queues a parameter block (the rebalancer's edges) in the last sub-step of every
                # frame; its physics does not change, so the single slab just steps
                one.step(steps)
                ref, ru = one.download()
            order = np.argsort(uids)
            same_set = len(uids) == len(ru) and np.array_equal(uids[order], ru)
            same_bits = same_set and all(np.array_equal(state[f][order].view("u4"), ref[f].view("u4")) for f in ("x", "y", "v_x", "v_y"))
            out["result"] = ("bit-identical" if same_bits else "MISMATCH" if same_set else "PARTICLES LOST OR DUPLICATED")
            out["overflow_or_timeout_counters"] = int(sum(p[2] for p in parts))
    except Exception as e:
        out["result"] = f"check failed to run: {e!r}"[:300]
    return out


def run_cfg3(sph_b200, rank, world, stream, args, flush_buf, barrier):
    """BASELINE.json config 3 inside the 8-GPU run: 2 M particles per GPU = 16 M, 8 x-slabs, halo exchange and load
    balancing on; same timing rules as the main section, shorter pre-roll."""
    import torch
    import torch.distributed as dist
    from sph_b200.slab import SlabRunner
    n_per = args.cfg3_particles
    prob = sph_b200.make_problem(n_per * world, tank_w=problem_dims(n_per * world, args.water_frac), water_frac=args.water_frac, nranks=world)
    t = sph_b200.default_params(prob["h"], prob["tank_w"], prob["tank_h"], args.preset)
    t.mover_center_x = 0.75 * prob["tank_w"]
    preroll = min(args.preroll, 600)
    with torch.cuda.stream(stream):
        sim = SlabRunner(prob, t, rank, world, stream, capacity_factor=2.0, balance_policy=args.balance,
                         halo_width=args.halo_width, exchange_period=args.exchange_period,
                             exchanges_per_step=args.exchanges_per_step, time_proportional=not args.time_fixed_step)
        sim.init_lattice()
        sim.run(preroll + args.warmup)
        barrier()
        blocks = []
        for _ in range(5):
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
            for k in range(args.steps):
                flush_buf.zero_()
                ev[k][0].record(stream); sim.run(1); ev[k][1].record(stream)
            barrier()
            blocks.append(float(sum(a.elapsed_time(b) for a, b in ev)))
        st = sim.stats()
    bt = torch.tensor(blocks, device="cuda", dtype=torch.float64)
    dist.all_reduce(bt, op=dist.ReduceOp.MAX)
    cnt = torch.tensor([st["n_local"], st["capacity_overflow"], st["msg_overflow"], st["exchange_timeouts"]], device="cuda", dtype=torch.float64)
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    med = statistics.median(bt.tolist())
    n_global = int(cnt[0].item())
    del sim
    return {"workload": f"2D dam-break block, {n_global} particles on {world} x-slabs, preset {args.preset}", "value": n_global * args.steps / (med * 1e-3),
            "unit": "particle-steps/s", "ms_per_step": med / args.steps, "blocks": len(blocks), "steps_per_block": args.steps,
            "state": f"{preroll} pre-roll + {args.warmup} warm-up steps from the lattice", "mean_neighbours_per_particle": st["mean_neighbours"],
            "step_hbm_frac_per_gpu": ALG_BYTES["step"] * n_global / world / (med / args.steps * 1e-3) / 1e9 / measured_peak()[0],
            "particles_resident": n_global, "particles_created": prob["n_global"], "capacity_overflow": int(cnt[1].item()),
            "msg_overflow": int(cnt[2].item()), "exchange_timeouts": int(cnt[3].item())}


class SingleGpu:
    """One GPU, one slab, whole tank: sph_step from a CUDA graph."""

    def __init__(self, sph, prob, t, stream, args):
        import numpy as np
        self.sph, self.prob, self.t = sph, prob, t
        self.np = np
        self.stream = stream
        self.cap = prob["n_global"] + 4096
        self.ctx = sph.Context(prob["tank_w"], prob["tank_h"], prob["h"], self.cap, stream=stream.cuda_stream)
        self.ctx.set_params(t)

    @property
    def launches(self):
        return self.ctx.launches

    def init_lattice(self):
        self.ctx.init_lattice(self.prob)            # filled on the device: no host AoS

    def run(self, n):
        if n > 0:
            self.ctx.step(n)

    def state_save(self): self.ctx.state_save()
    def state_restore(self): self.ctx.state_restore()

    def stage_times(self, nsteps, flush_buf):
        import torch
        names = ("advect", "sort1", "density", "relax", "sort2")
        acc = {k: 0.0 for k in names}
        c = self.ctx
        for _ in range(nsteps):
            flush_buf.zero_()
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
            evs[0].record(self.stream); c.advect()
            evs[1].record(self.stream); c.sort()
            evs[2].record(self.stream); c.density()
            evs[3].record(self.stream); c.relax()
            evs[4].record(self.stream); c.sort()
            evs[5].record(self.stream)
            torch.cuda.synchronize()
            for i, k in enumerate(names):
                acc[k] += evs[i].elapsed_time(evs[i + 1])
        return {k: v / nsteps for k, v in acc.items()}

    def kernel_name(self, stage):
        return {"advect": "k_advect", "density": "k_density", "relax": "k_relax",
                "sort1": "k_scan_apply+k_scatter+k_reorder",
                "sort2": "k_scan_apply+k_scatter+k_reorder"}[stage]

    def e2e(self, frames, flush_buf, block=None):
        """Frames through the public frame call with HOST buffers, in blocks of `block` frames that each start from the
        restored state (restore outside the clock).  Two protocols are timed:
        synchronous  sph_run_frame: the call returns when the frame's coordinates are in host memory (L2 flushed
                     before every frame, outside the clock) -- the protocol of the round-1 records;
        pipelined    sph_run_frame_async / sph_coords_wait: frame f is collected after frame f+1 has been
                     submitted, so its copy overlaps the next frame's steps, which is what the reference's
                     compute rank does with MPI_Isend (fluid.c:283-287, :354-365).  One clock around all frames,
                     every frame's parameter block goes in and every frame's coordinates come out inside it.
        The pipelined figure is reported when it ran and its last frame equals what the synchronous call packs
        from the same state; otherwise the synchronous one is."""
        block = max(1, min(block or frames, frames))
        self.ctx.state_save()                          # the state every block starts from: the one e2e() was called in
        out = self._e2e_sync(frames, flush_buf, block)
        out["pipelined"] = False
        try:
            pipe = self._e2e_pipelined(frames, block)
            if pipe.pop("ok"):
                pipe["sync_seconds"] = out["seconds"]
                pipe["pipelined"] = True
                out = pipe
            else:
                out["pipelined_error"] = "last frame differs from the synchronous feed"
        except Exception as e:  # the verified protocol's number stands
            out["pipelined_error"] = repr(e)[:200]
        return out

    def _e2e_pipelined(self, frames, block):
        import torch
        np = self.np
        bufs = [torch.empty(2 * self.cap, dtype=torch.int16).pin_memory().numpy() for _ in range(2)]
        c = self.ctx
        for f in range(2):
            c.coords_wait(c.run_frame_async(self.t, 4, bufs[f]))
        secs, done, n = 0.0, 0, 0
        while done < frames:
            c.state_restore()                          # every block times the frames `value`'s blocks time
            torch.cuda.synchronize()
            tickets = []
            t0 = time.perf_counter()
            for f in range(block):
                tickets.append(c.run_frame_async(self.t, 4, bufs[f % 2]))
                if f > 0:
                    n = c.coords_wait(tickets[f - 1])
            n = c.coords_wait(tickets[-1])             # (the last frame's copy is exposed once per block)
            secs += time.perf_counter() - t0
            done += block
        last = bufs[(block - 1) % 2][:2 * n].copy()
        ok = bool(np.array_equal(last, c.pack_coords().ravel()[:2 * n]))
        return {"ok": ok, "seconds": secs, "steps": 4 * done, "h2d_per_step": 64 / 4, "d2h_per_step": 4 * n / 4,
                "frames_per_block": block}

    def _e2e_sync(self, frames, flush_buf, block):
        import torch
        coords = torch.empty(2 * self.cap, dtype=torch.int16).pin_memory()
        xy = coords.numpy()
        for _ in range(2):
            self.ctx.run_frame(self.t, 4, xy)
        secs, done, n = 0.0, 0, 0
        while done < frames:
            self.ctx.state_restore()
            for _ in range(block):
                flush_buf.zero_()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                n = self.ctx.run_frame(self.t, 4, xy)      # returns after the coordinates are in host memory
                secs += time.perf_counter() - t0
            done += block
        return {"seconds": secs, "steps": 4 * done, "h2d_per_step": 64 / 4, "d2h_per_step": 4 * n / 4, "frames_per_block": block}

    def stats(self):
        s = self.ctx.status()
        npairs = self.ctx.L.sph_get_pairs(self.ctx.h, None, 0)
        return {"n_local": s.n_local, "n_halo": s.n_halo, "max_bucket": s.max_bucket,
                "mean_neighbours": 2.0 * npairs / max(s.n_local, 1),
                "capacity_overflow": s.capacity_overflow, "msg_overflow": s.msg_overflow, "exchange_timeouts": 0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--particles", dest="n", type=int, default=1_000_000, help="particles per GPU")
    ap.add_argument("--preroll", type=int, default=1000, help="untimed steps before warm-up (state preparation)")
    ap.add_argument("--water-frac", type=float, default=0.5)
    ap.add_argument("--preset", default="x")
    ap.add_argument("--halo-width", type=float, default=None,
                    help="ghost-layer width in h at N > 1 (default: the build's, 2; the one-exchange build: 3.5, 4.5 with --visc-stab)")
    ap.add_argument("--visc-stab", type=float, default=None, metavar="GAMMA",
                    help="force the stabilised viscosity gather with this gamma for every block (0 = plain gather everywhere); "
                         "default: the library's own rule (gamma 0.5 where dt*sigma >= 0.5, i.e. the goo preset, DESIGN.md 5b)")
    ap.add_argument("--time-fixed-step", action="store_true",
                    help="--balance time with the reference's fixed edge step of h/8 per frame instead of a step proportional to the imbalance")
    ap.add_argument("--exchanges-per-step", type=int, default=1, choices=[1, 2],
                    help="N > 1: 2 = neighbours meet after the prediction and after the relaxation, like the reference (fluid.c:310-348); "
                         "1 = once, ghosts relaxed redundantly (default)")
    ap.add_argument("--exchange-period", type=int, default=2,
                    help="N > 1, one exchange per step: neighbours meet every this many steps (ghost layer 3.5 h per step; default 2)")
    ap.add_argument("--cpu-steps", type=int, default=20, help="timed steps of the cpu_baseline sample")
    ap.add_argument("--cpu-warmup", type=int, default=300, help="untimed steps of the cpu_baseline sample (bounded: the "
                    "reference arm, --impl reference, runs the full pre-roll)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--min-timed-ms", type=float, default=500.0, help="repeat the block of --steps timed steps until this much device time is measured")
    ap.add_argument("--max-repeats", type=int, default=400)
    ap.add_argument("--no-parity-check", action="store_true", help="N > 1: skip the small N-slab-vs-1-slab bit comparison before the timed region")
    ap.add_argument("--no-cfg3", action="store_true", help="--gpus 8: skip the second timed section on BASELINE config 3 (16 M particles)")
    ap.add_argument("--force-cfg3", action="store_true", help="run that section at any N > 1 (to exercise the code path on fewer GPUs)")
    ap.add_argument("--cfg3-particles", type=int, default=2_000_000, help="particles per GPU of that section (8 x 2 M = config 3's 16 M)")
    ap.add_argument("--balance", default="time", choices=["count", "cost", "time"],
                    help="slab edge policy at N > 1: each slab's measured device time between meetings (default), the reference's "
                         "particle counts (renderer.c:427-477), or the modelled work estimate; results are identical bit for bit")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.exchanges_per_step == 2:
        args.exchange_period = 1
    if args.impl == "reference":
        return main_reference(args)
    return main_ours(args)


if __name__ == "__main__":
    sys.exit(main())
