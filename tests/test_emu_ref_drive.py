"""The reference's UNMODIFIED start_simulation() driving the reference-named entry points (sph_b200/host/ref_api.c:
arming, lazy attach, host mirror) without a GPU: the GPU driver binary of tests/test_gpu_ref_drive.py with the
kernel-source emulator library (tests/emu) preloaded in front of libsph_b200.so.  Frames must equal, bit for bit,
those of the handle API on the same emulator.  Host-logic regression net only; the claim itself is the GPU test."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import sph_b200
from emu.build_emu import build as build_emu
from oracle.oracle import lattice, make_problem
from test_ref_drive import GPU_DRIVE, HOT, RESTART, WORLD_CPU, WORLD_GPU, bindings, pack, read_drive, read_world


@pytest.mark.skipif(not os.path.exists(GPU_DRIVE), reason="oracle/_ref not built")
@pytest.mark.parametrize("mirror_every,stab", [(None, None), ("4", None), (None, "0.5,0.0")])
def test_unmodified_reference_driver_on_the_emulated_library(built_lib, tmp_path, monkeypatch, mirror_every, stab):
    emu = build_emu()
    out = str(tmp_path / "emu.bin")
    env = dict(os.environ, LD_PRELOAD=emu)
    if mirror_every:
        env["SPH_REF_MIRROR_EVERY"] = mirror_every
    if stab:
        env["SPH_VISC_STAB"] = stab                     # the reference's entry points have no call for it
    r = subprocess.run([GPU_DRIVE, "--frames", "6", "--out", out], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, (r.stdout[-300:], r.stderr[-800:])
    b = bindings(r.stdout)
    assert b["start_simulation"].endswith("libref_driver.so")
    assert all(b[k].endswith("libsph_emu.so") for k in HOT), b
    n, w, h, first, frames = read_drive(out)
    prob = make_problem(1500)
    a, uid = lattice(prob)
    monkeypatch.setattr(sph_b200, "_lib", sph_b200._bind(C.CDLL(emu)))

    def sph(t):
        o = sph_b200.Tunable(); C.memmove(C.byref(o), C.byref(t), 64); return o

    ctx = sph_b200.Context(w, h, float(first.smoothing_radius), n)
    if stab:
        ctx.set_viscosity_stabilisation(*[float(v) for v in stab.split(",")])
    ctx.set_params(sph(first))
    ctx.upload(a, uid)
    coords = np.zeros(2 * n, "i2")
    for k, (blk, got) in enumerate(frames):
        assert ctx.run_frame(sph(blk), 4, coords) == n
        s, _ = ctx.download()
        assert np.array_equal(pack(s["x"], s["y"], w, h), got), k


def frames_of_all_ranks(out, ranks):
    """-> world, per frame: [coords of rank 0, rank 1, ...] from <out>.r<rank> (oracle/ref_build/ref_drive.c --ranks)"""
    parts = [read_drive(f"{out}.r{r}") for r in range(ranks)]
    n, w, h = parts[0][:3]                       # only compute rank 0 tells its render rank (fluid.c:167-171)
    return n, w, h, [[p[4][k][1] for p in parts] for k in range(len(parts[0][4]))], [p[3] for p in parts]


def check_ranks_against_one_rank(drive, env, tmp_path, ranks, frames, libname, wobble=0, explicit=0):
    """The unmodified driver as `ranks` compute ranks (one slab each, over the mini-MPI, glue object linked in)
    against the same driver as ONE rank: every frame must hold the same pixels, bit for bit, whoever owns them."""
    one, many = str(tmp_path / "one.bin"), str(tmp_path / "many.bin")
    r1 = subprocess.run([drive, "--frames", str(frames), "--out", one], capture_output=True, text=True, timeout=300, env=env)
    assert r1.returncode == 0, (r1.stdout[-300:], r1.stderr[-800:])
    rk = subprocess.run([drive, "--ranks", str(ranks), "--wobble", str(wobble), "--explicit", str(explicit), "--frames", str(frames),
                         "--out", many], capture_output=True, text=True, timeout=600, env=env)
    assert rk.returncode == 0, (rk.stdout[-300:], rk.stderr[-800:])
    b = bindings(rk.stdout)
    assert b["start_simulation"].endswith("libref_driver.so")
    assert all(b[k].endswith(libname) for k in HOT), b
    assert "sph_ref_api:" not in rk.stderr, rk.stderr[-800:]          # no stage reported an error
    if explicit:
        # sph_ref_set_rank + sph_ref_set_transport instead of the glue: two send/receive pairs per exchange, two
        # exchanges per step, one at the attach; the kill_sim scatter arrives in the 4th sub-step of frame `frames`,
        # after three more whole steps (fluid.c:293-304)
        calls = [int(line.split(": ")[1]) for line in rk.stdout.splitlines() if line.startswith("explicit transport calls")]
        assert calls == [2 * (2 * (4 * frames + 3) + 1)] * ranks, calls
    n, w, h, first, f1 = read_drive(one)
    nk, wk, hk, fk, firsts = frames_of_all_ranks(many, ranks)
    assert (nk, wk, hk) == (n, w, h) and len(fk) == len(f1) == frames
    # the slabs are partitionProblem's (geometry.c:101-160): contiguous, covering the tank
    assert firsts[0].node_start_x == 0.0 and firsts[-1].node_end_x == np.float32(w)
    if wobble:      # the stubs moved the interior edges the way the balancer does (h/8 per frame) and back again
        blocks = [read_drive(f"{many}.r1")[4][k][0] for k in range(frames)]
        assert blocks[4].node_start_x > blocks[1].node_start_x and blocks[frames - 1].node_start_x == blocks[0].node_start_x
    moved = False
    for k in range(frames):
        counts = [len(c) for c in fk[k]]
        assert sum(counts) == n, (k, counts)                          # nobody lost, nobody duplicated
        moved |= counts != [len(c) for c in fk[0]]
        got = np.sort(np.concatenate(fk[k]).copy().view("i4").ravel())
        want = np.sort(f1[k][1].copy().view("i4").ravel())
        assert np.array_equal(got, want), k
    assert moved                                                      # particles did change owner on the way


@pytest.mark.skipif(not os.path.exists(GPU_DRIVE), reason="oracle/_ref not built")
@pytest.mark.parametrize("ranks,mirror_every,wobble", [(3, None, 0), (2, "4", 0), (4, None, 1), (3, "4", 1)])
def test_unmodified_reference_driver_with_several_compute_ranks_on_the_emulated_library(built_lib, tmp_path, ranks, mirror_every,
                                                                                        wobble):
    """BASELINE config 1's shape (mpirun -n 4: 1 render + 3 compute ranks) through the reference-named entry points:
    identify_oob_particles / startHaloExchange move the slab messages through the host's MPI_Sendrecv
    (sph_b200/host/glue/sph_ref_mpi_glue.c -> sph_exchange_via_host); slab edges that move in mid-run follow at the
    next predict_positions."""
    env = dict(os.environ, LD_PRELOAD=build_emu())
    if mirror_every:
        env["SPH_REF_MIRROR_EVERY"] = mirror_every
    check_ranks_against_one_rank(GPU_DRIVE, env, tmp_path, ranks, 10, "libsph_emu.so", wobble)


@pytest.mark.skipif(not os.path.exists(GPU_DRIVE), reason="oracle/_ref not built")
@pytest.mark.skipif(not os.path.exists(GPU_DRIVE), reason="oracle/_ref not built")
def test_host_that_announces_rank_and_transport_itself_on_the_emulated_library(built_lib, tmp_path):
    """sph_ref_set_rank + sph_ref_set_transport (a host that is edited anyway) instead of the glue object's weak hook."""
    check_ranks_against_one_rank(GPU_DRIVE, dict(os.environ, LD_PRELOAD=build_emu()), tmp_path, 3, 10, "libsph_emu.so", 1, explicit=1)


@pytest.mark.skipif(not os.path.exists(GPU_DRIVE), reason="oracle/_ref not built")
def test_several_compute_ranks_through_the_reference_names_on_the_one_exchange_build(built_lib, tmp_path):
    """-DSPH_ONE_EXCHANGE=1 relaxes its ghosts itself: the startHaloExchange after updateVelocities has nothing to move
    (sph_exchanges_per_step() == 1) and the frames are still the one-rank frames."""
    lib = build_emu(defines=("SPH_ONE_EXCHANGE=1",), name="libsph_emu_sph_one_exchange1.so")
    check_ranks_against_one_rank(GPU_DRIVE, dict(os.environ, LD_PRELOAD=lib), tmp_path, 3, 10, os.path.basename(lib), 1)


KEYS = "3:remove 6:b 9:add 12:a 15:remove 17:add 18:x"       # what the headless "user" presses, frame:key (render_stubs.c)


def check_whole_program(world, env, tmp_path, ranks, frames, libname, script=None):
    """The reference's whole program (its main(), renderer.c with its load balancer, controls.c; oracle/ref_build/
    ref_world.c) with the compute ranks' hot path in the library: K ranks must draw the pixels ONE rank draws."""
    outs = {}
    if script:
        env = dict(env, SPH_RENDER_SCRIPT=script)
    for k in (1, ranks):
        outs[k] = str(tmp_path / f"world{k}.bin")
        r = subprocess.run([world, "--ranks", str(k), "--frames", str(frames), "--out", outs[k]], capture_output=True,
                           text=True, timeout=600, env=env)
        assert r.returncode == 0, (r.stdout[-300:], r.stderr[-800:])
        assert "sph_ref_api:" not in r.stderr, r.stderr[-800:]
        b = bindings(r.stdout)
        assert all(b[k_].endswith("libref_full.so") for k_ in ("start_renderer", "check_partition_left", "start_simulation")), b
        assert all(b[k_].endswith(libname) for k_ in HOT), b
    K1, w, h, one = read_world(outs[1])
    K, wk, hk, many = read_world(outs[ranks])
    assert (K1, K, wk, hk) == (1, ranks, w, h) and len(one) == len(many) == frames
    for f in range(frames):
        assert np.array_equal(many[f][1], one[f][1])                                  # same mover path
        assert np.array_equal(np.sort(many[f][2].copy().view("i8").ravel()), np.sort(one[f][2].copy().view("i8").ravel())), f
    # the reference's balancer did move a slab edge on the way (renderer.c:427-477), and the pixels did not notice
    assert any(not np.array_equal(many[f][0], many[0][0]) for f in range(frames))
    if script:
        # remove_partition (controls.c:405-426) parked the last slab outside the tank: it drained into its neighbour
        # through the migration path; add_partition (:429-455) split the last active slab and it filled up again
        parked = [f for f in range(frames) if many[f][0][-1, 0] > w]
        assert parked and parked[0] == 3 and (frames - 1) not in parked
        assert many[4][0][-2, 1] == np.float32(w)


@pytest.mark.skipif(not os.path.exists(WORLD_GPU), reason="oracle/_ref not built")
@pytest.mark.parametrize("ranks", [3, 4, 8])
def test_whole_reference_program_with_its_renderer_on_the_emulated_library(built_lib, tmp_path, ranks):
    # (4 ranks: both runs under a shuffled block / thread order of the emulator -- the arrival order of the atomics must not matter)
    # (3 ranks: the host mirror only at the driver's frame rate, every 4th step, fluid.c:105)
    extra = {4: {"SPH_EMU_ORDER": "shuffle"}, 3: {"SPH_REF_MIRROR_EVERY": "4"}}.get(ranks, {})
    check_whole_program(WORLD_GPU, dict(os.environ, LD_PRELOAD=build_emu(), **extra), tmp_path, ranks, 14, "libsph_emu.so")


GOO_KEYS = "2:y 3:remove 9:add 12:x"      # the goo preset needs SPH_VISC_STAB


@pytest.mark.skipif(not os.path.exists(WORLD_GPU), reason="oracle/_ref not built")
@pytest.mark.parametrize("keys,stab", [(KEYS, None), (GOO_KEYS, "0.5,0.5")])
def test_reference_controls_park_and_re_add_a_slab_and_switch_presets_on_the_emulated_library(built_lib, tmp_path, keys, stab):
    """The reference's own remove_partition / add_partition / set_fluid_b / set_fluid_x / set_fluid_y (controls.c,
    unmodified) pressed in mid-run while its balancer keeps moving the edges: three ranks still draw the one-rank
    pixels in every frame -- also through the goo phase with the stabilised viscosity gather switched on by environment."""
    env = dict(os.environ, LD_PRELOAD=build_emu())
    if stab:
        env["SPH_VISC_STAB"] = stab
    check_whole_program(WORLD_GPU, env, tmp_path, 3, 20, "libsph_emu.so", keys)


def check_config1_against_the_pure_reference(env, tmp_path, frames=120):
    """BASELINE config 1 -- the reference's default dam-break, `mpirun -n 4` (its render rank + 3 compute ranks) -- run
    twice as the reference's whole unmodified program: once pure (sph_ref_world_cpu), once with the compute ranks'
    hot path in the library (sph_ref_world_gpu).  480 steps, mover dragged through the water, the reference's balancer
    active in both.  Per-particle agreement is not defined over such a run (Gauss-Seidel scatter vs gather, chaotic
    system: SURVEY.md 8(c)); the drawn frames must agree in their statistics, in GL units (the screen is 2 x 2):
    centre of mass and spread of every frame within 0.01 vertically (measured: 0.002) and 0.02 horizontally
    (measured: 0.004), and the first 10 frames, before the sweeps' order matters, within 1e-4."""
    outs = []
    for exe, e in ((WORLD_CPU, dict(os.environ)), (WORLD_GPU, env)):
        out = str(tmp_path / (os.path.basename(exe) + ".bin"))
        r = subprocess.run([exe, "--ranks", "3", "--frames", str(frames), "--out", out], capture_output=True, text=True,
                           timeout=900, env=e)
        assert r.returncode == 0, (r.stdout[-300:], r.stderr[-800:])
        assert "sph_ref_api:" not in r.stderr, r.stderr[-800:]
        outs.append(read_world(out)[3])
    ref, got = outs
    assert len(ref) == len(got) == frames
    worst = np.zeros(4)
    for k, (a, b) in enumerate(zip(ref, got)):
        assert len(a[2]) == len(b[2]) == 1508 and np.array_equal(a[1], b[1])          # everybody drawn; same mover path
        d = np.abs([a[2][:, 1].mean() - b[2][:, 1].mean(), a[2][:, 1].std() - b[2][:, 1].std(),
                    a[2][:, 0].mean() - b[2][:, 0].mean(), a[2][:, 0].std() - b[2][:, 0].std()])
        assert np.all(d <= (1e-4 if k < 10 else np.array([0.01, 0.01, 0.02, 0.02]))), (k, d)
        worst = np.maximum(worst, d)
    # the water did collapse and slosh (the statistics above are not those of a fluid at rest)
    assert ref[0][2][:, 1].mean() - ref[60][2][:, 1].mean() > 0.5
    return worst


@pytest.mark.skipif(not os.path.exists(WORLD_GPU), reason="oracle/_ref not built")
def test_config1_whole_program_statistics_against_the_pure_reference_on_the_emulated_library(built_lib, tmp_path):
    check_config1_against_the_pure_reference(dict(os.environ, LD_PRELOAD=build_emu()), tmp_path)


def check_restart_from_a_moving_fluid(env, tmp_path, ranks=3, steps=8):
    """oracle/ref_build/ref_restart.c: a host on the reference-named entry points with EXPLICIT sph_ref_set_rank /
    sph_ref_set_transport / sph_ref_attach, no host mirror, sph_ref_sync_to_host at the end -- starting from a fluid
    that already moves.  The attach must hand every slab its ghost layer (one exchange), or the first viscosity pass
    misses the neighbours across the edges; with it, `ranks` slabs equal one slab bit for bit (without it they do
    not: checked by hand when this was written)."""
    def rd(path, w=15.0, h=8.4375):
        raw = open(path, "rb").read()
        n = int(np.frombuffer(raw, "i4", 1)[0])
        state = np.frombuffer(raw, "f4", 4 * n, 4).reshape(n, 4)
        # the slab's device-side frame (sph_ref_pack_coords) is the reference's formula on these positions (fluid.c:358-361)
        feed = np.frombuffer(raw, "i2", 2 * n, 4 + 16 * n).reshape(n, 2)
        want = pack(state[:, 0].copy(), state[:, 1].copy(), w, h)
        assert np.array_equal(np.sort(feed.copy().view("i4").ravel()), np.sort(want.copy().view("i4").ravel()))
        return state

    def canon(a):
        return a[np.lexsort(a.view("u4").T[::-1])].view("u4")

    for k in (1, ranks):
        r = subprocess.run([RESTART, "--ranks", str(k), "--steps", str(steps), "--out", str(tmp_path / f"restart{k}.bin")],
                           capture_output=True, text=True, timeout=300, env=env)
        assert r.returncode == 0 and "sph_ref_api:" not in r.stderr, (r.stdout[-300:], r.stderr[-800:])
    one = rd(tmp_path / "restart1.bin.r0")
    parts = [rd(tmp_path / f"restart{ranks}.bin.r{r}") for r in range(ranks)]
    assert len(one) == 1508 and sum(len(p) for p in parts) == 1508
    assert [len(p) for p in parts] != [522, 493, 493]                    # particles crossed the edges on the way
    assert np.abs(one[:, 2:]).max() > 3.0                                # and it is not a fluid at rest
    assert np.array_equal(canon(one), canon(np.concatenate(parts)))


@pytest.mark.skipif(not os.path.exists(RESTART), reason="oracle/_ref not built")
def test_restart_from_a_moving_fluid_through_the_reference_names_on_the_emulated_library(built_lib, tmp_path):
    check_restart_from_a_moving_fluid(dict(os.environ, LD_PRELOAD=build_emu()), tmp_path)


def test_reference_call_order_on_the_emulated_library(built_lib, monkeypatch):
    """tests/test_ref_api.py's GPU test (the reference's call sequence through the reference-named entry points
    equals sph_step bit for bit) with the emulator library in place of libsph_b200.so."""
    import test_ref_api
    emu = build_emu()
    monkeypatch.setattr(sph_b200, "_lib", sph_b200._bind(C.CDLL(emu)))
    test_ref_api.test_reference_call_order_reproduces_sph_step(emu)


def test_reference_named_attach_refuses_more_than_one_rank(built_lib):
    """start/finishHaloExchange and transferOOBParticles are single-rank no-ops; a host that announces more than one
    compute rank must be told so, not handed isolated slabs."""
    L = C.CDLL(build_emu())
    L.sph_ref_last_error.restype = C.c_char_p
    L.sph_ref_set_rank(1, 3)
    try:
        assert L.sph_ref_attach(None, None, None, None, 0) != 0
        assert b"ONE compute rank" in L.sph_ref_last_error()
    finally:
        L.sph_ref_set_rank(0, 1)
