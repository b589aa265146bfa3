"""The slab kernels' logic without a GPU: message packing in k_advect / k_relax, k_unpack, windows of grid
columns, migration, the per-frame rebalancing -- the CUDA source compiled for the host (tests/emu) as the engine
of the real multi-rank driver over gloo, with the buffers moved by torch.distributed (the "collective"
transport; the peer-memory transport needs NVLink and is covered by tests/test_gpu_slabs.py).
N slabs must equal one slab BIT FOR BIT, and both must sit at rounding level from the gather oracle."""
import numpy as np
import pytest

from emu.backend import EmuSlab
from oracle.oracle import default_tunable, lattice, make_problem
from test_slab_gloo import run_world


def emu_single(prob, t, steps, gamma=0.0):
    a, uid = lattice(prob)
    e = EmuSlab(prob["tank_w"], prob["tank_h"], prob["h"], len(a) + 64, 1, 0, 1)
    e.set_viscosity_stabilisation(gamma)
    e.set_params(t); e.upload(a, uid); e.step(steps)
    return e.download()


@pytest.mark.parametrize("world", [2, 3])
def test_emulated_slabs_equal_single_slab_bit_for_bit(tmp_path, built_lib, world):
    steps = 80
    parts = run_world(tmp_path, world, 1500, steps, True, "emu")
    state = np.concatenate([p["state"] for p in parts]); uid = np.concatenate([p["uid"] for p in parts])
    assert all(int(p["overflow"].sum()) == 0 for p in parts)
    assert np.array_equal(np.sort(uid), np.arange(1508)), "particles lost or duplicated in migration"
    order = np.argsort(uid)
    prob = make_problem(1500)
    ref, _ = emu_single(prob, default_tunable(prob["h"], prob["tank_w"], prob["tank_h"]), steps)
    for f in ("x", "y", "v_x", "v_y"):
        assert np.array_equal(state[f][order].view("u4"), ref[f].view("u4")), f


def test_emulated_four_slabs_block_with_mover_across_an_edge(tmp_path, built_lib):
    n_req, steps = 12000, 40
    parts = run_world(tmp_path, 4, n_req, steps, True, "emu_block")
    state = np.concatenate([p["state"] for p in parts]); uid = np.concatenate([p["uid"] for p in parts])
    assert all(int(p["overflow"].sum()) == 0 for p in parts), [p["overflow"] for p in parts]
    prob = make_problem(n_req, tank_w=15.0 * float(np.sqrt(n_req / 750.0)), water_frac=0.5)
    t = default_tunable(prob["h"], prob["tank_w"], prob["tank_h"]); t.mover_center_x = 0.4 * prob["tank_w"]
    ref, ru = emu_single(prob, t, steps)
    assert np.array_equal(np.sort(uid), ru)
    order = np.argsort(uid)
    for f in ("x", "y", "v_x", "v_y"):
        assert np.array_equal(state[f][order].view("u4"), ref[f].view("u4")), f


def test_emulated_stabilised_viscosity_is_decomposition_independent(tmp_path, built_lib):
    """k_coupling forms C for ghosts too (their neighbours sit in the 2h ghost layer), so the scaled impulses of
    the goo preset are the same bits on 3 slabs as on one."""
    steps = 80
    parts = run_world(tmp_path, 3, 1500, steps, True, "emu_goo_stabilised")
    state = np.concatenate([p["state"] for p in parts]); uid = np.concatenate([p["uid"] for p in parts])
    assert all(int(p["overflow"].sum()) == 0 for p in parts)
    assert np.array_equal(np.sort(uid), np.arange(1508))
    order = np.argsort(uid)
    prob = make_problem(1500)
    t = default_tunable(prob["h"], prob["tank_w"], prob["tank_h"], preset="y")
    ref, _ = emu_single(prob, t, steps, gamma=0.5)
    plain, _ = emu_single(prob, t, steps)
    for f in ("x", "y", "v_x", "v_y"):
        assert np.array_equal(state[f][order].view("u4"), ref[f].view("u4")), f
    assert not np.array_equal(ref["x"].view("u4"), plain["x"].view("u4"))     # the pass did engage


@pytest.mark.parametrize("config,world,n_req,steps,halo", [("emu", 3, 1500, 80, "0"), ("emu_block", 4, 12000, 40, "0"),
                                                         ("emu_goo_stabilised", 3, 1500, 80, "4.5")])
def test_one_exchange_build_equals_single_slab_bit_for_bit(tmp_path, built_lib, monkeypatch, config, world, n_req, steps, halo):
    """-DSPH_ONE_EXCHANGE=1: neighbours meet once per step; the ghosts travel with x_prev in a >= 3 h layer and are
    relaxed redundantly, so the next viscosity pass has their velocities without a second message.  Same bits as
    one slab (and therefore as the two-exchange build), with the rebalancer moving the edges underneath."""
    monkeypatch.setenv("SPH_EMU_DEFINES", "SPH_ONE_EXCHANGE=1")
    monkeypatch.setenv("SPH_EMU_HALO_WIDTH", halo)
    parts = run_world(tmp_path, world, n_req, steps, True, config)
    assert all(int(p["exchanges"][0]) == 1 for p in parts), "the workers did not run the one-exchange build"
    state = np.concatenate([p["state"] for p in parts]); uid = np.concatenate([p["uid"] for p in parts])
    assert all(int(p["overflow"].sum()) == 0 for p in parts), [p["overflow"] for p in parts]
    block, goo = "block" in config, "goo" in config
    prob = make_problem(n_req, tank_w=15.0 * float(np.sqrt(n_req / 750.0)), water_frac=0.5) if block else make_problem(n_req)
    t = default_tunable(prob["h"], prob["tank_w"], prob["tank_h"], preset="y" if goo else "x")
    if block:
        t.mover_center_x = 0.4 * prob["tank_w"]
    ref, ru = emu_single(prob, t, steps, gamma=0.5 if goo else 0.0)
    assert np.array_equal(np.sort(uid), ru), "particles lost or duplicated in migration"
    order = np.argsort(uid)
    for f in ("x", "y", "v_x", "v_y"):
        assert np.array_equal(state[f][order].view("u4"), ref[f].view("u4")), f


@pytest.mark.parametrize("defines", ["", "SPH_ONE_EXCHANGE=1"])
def test_emulated_partition_removed_and_added_back(tmp_path, built_lib, monkeypatch, defines):
    """Runtime partition control on the CUDA source (controls.c:405-455): the last of three slabs is parked outside the
    tank at step 43, drains through the migration path, and is added back at step 123 -- in the shipped build and in
    the one-exchange build.  Nothing is lost and the result is the single-slab result bit for bit."""
    monkeypatch.setenv("SPH_EMU_DEFINES", defines)
    steps = 200
    parts = run_world(tmp_path, 3, 1500, steps, True, "emu_elastic")
    state = np.concatenate([p["state"] for p in parts]); uid = np.concatenate([p["uid"] for p in parts])
    assert np.array_equal(np.sort(uid), np.arange(1508))
    assert all(int(p["overflow"][0]) == 0 for p in parts), [p["overflow"] for p in parts]
    drained = [h for h in parts[2]["history"] if h[0] == -1]
    assert drained and drained[0][1] == 0, "the parked slab still held particles when it was added back"
    assert len(parts[2]["uid"]) > 100, "the re-added slab did not refill"
    prob = make_problem(1500)
    ref, _ = emu_single(prob, default_tunable(prob["h"], prob["tank_w"], prob["tank_h"]), steps)
    order = np.argsort(uid)
    for f in ("x", "y", "v_x", "v_y"):
        assert np.array_equal(state[f][order].view("u4"), ref[f].view("u4")), f


@pytest.mark.parametrize("config,world,n_req,steps,period", [("emu", 3, 1500, 80, 2), ("emu_block", 4, 12000, 40, 2),
                                                           ("emu_block", 2, 12000, 44, 4), ("emu_goo_stabilised", 2, 1500, 80, 2),
                                                           ("emu_elastic", 3, 1500, 200, 2)])
def test_exchange_period_equals_single_slab_bit_for_bit(tmp_path, built_lib, monkeypatch, config, world, n_req, steps, period):
    """sph_set_exchange_period (one-exchange build): neighbours meet every `period` steps; in between every slab
    advances its ghosts itself (k_advect no longer drops them, k_relax relaxes them), in a layer of 3.5 h per step
    (4.5 h with the stabilised viscosity gather).  Every frame's parameter block (the rebalancer's new edges) forces
    an exchange in its own step.  Still the single-slab bits: plain fluid on 3 slabs, the block with the mover on an
    edge on 4 slabs and (4 steps between meetings) on 2, goo with the stabilised gather, and a slab parked and re-added."""
    monkeypatch.setenv("SPH_EMU_DEFINES", "SPH_ONE_EXCHANGE=1")
    monkeypatch.setenv("SPH_EMU_XPERIOD", str(period))
    monkeypatch.setenv("SPH_EMU_HALO_WIDTH", str((4.5 if "goo" in config else 3.5) * period))
    parts = run_world(tmp_path, world, n_req, steps, True, config)
    assert all(int(p["exchanges"][0]) == 1 for p in parts), "the workers did not run the one-exchange build"
    # the ranks really met less often: every `period`-th step, plus the steps in which the rebalancer's block landed
    # (one per frame of 4 steps; they coincide when the period divides 4) -- never in every step
    met = [int(p["n_exchanges"][0]) for p in parts]
    assert all(m == met[0] for m in met) and steps // period <= met[0] <= steps // period + steps // 4 + 2 < steps, (met, steps)
    state = np.concatenate([p["state"] for p in parts]); uid = np.concatenate([p["uid"] for p in parts])
    assert all(int(p["overflow"].sum()) == 0 for p in parts), [p["overflow"] for p in parts]
    block, goo = "block" in config, "goo" in config
    prob = make_problem(n_req, tank_w=15.0 * float(np.sqrt(n_req / 750.0)), water_frac=0.5) if block else make_problem(n_req)
    t = default_tunable(prob["h"], prob["tank_w"], prob["tank_h"], preset="y" if goo else "x")
    if block:
        t.mover_center_x = 0.4 * prob["tank_w"]
    ref, ru = emu_single(prob, t, steps, gamma=0.5 if goo else 0.0)
    assert np.array_equal(np.sort(uid), ru), "particles lost or duplicated in migration"
    order = np.argsort(uid)
    for f in ("x", "y", "v_x", "v_y"):
        assert np.array_equal(state[f][order].view("u4"), ref[f].view("u4")), f


def test_parked_slab_with_a_small_message_capacity_drains_without_losing_anyone(tmp_path, built_lib, monkeypatch):
    """A slab parked outside the tank (controls.c:405-426) hands over ALL its particles, and its window of grid columns
    leaves the tank with it.  With a message capacity below its population the emigrants that do not fit must stay
    resident (binned into the nearest window cell) and follow in the next steps -- round 1 dropped them for good
    (314 of 1508 at capacity 400).  Nobody is lost; msg_overflow says that particles had to wait."""
    monkeypatch.setenv("SPH_EMU_MSG_CAPACITY", "400")
    steps = 200
    parts = run_world(tmp_path, 3, 1500, steps, True, "emu_elastic")
    uid = np.concatenate([p["uid"] for p in parts])
    assert np.array_equal(np.sort(uid), np.arange(1508)), (len(uid), len(np.unique(uid)))
    assert all(int(p["overflow"][0]) == 0 for p in parts), [p["overflow"] for p in parts]      # capacity_overflow: none lost
    assert sum(int(p["overflow"][1]) for p in parts) > 0                                       # emigrants did wait
    drained = [h for h in parts[2]["history"] if h[0] == -1]
    assert drained and drained[0][1] == 0, "the parked slab still held particles when it was added back"
    assert len(parts[2]["uid"]) > 100, "the re-added slab did not refill"


def test_parameters_set_in_the_middle_of_a_slab_step_do_not_move_the_window_under_the_sort(built_lib, monkeypatch):
    """sph_set_params between sph_advect and sph_sort on a slab whose edges move by 3 h: round 1 re-derived the window
    at once, the sort then ran with keys binned for the OLD window, and 91 of 1508 particles vanished with
    capacity_overflow still 0.  Now the physics applies at once and the edges land with the next sph_advect;
    sph_set_edges in mid-step is refused."""
    import sph_b200
    from emu.backend import use_emulator
    from oracle.oracle import lattice
    monkeypatch.setattr(sph_b200, "_lib", sph_b200._lib)      # use_emulator() rebinds the module's library: undone after the test
    sph = use_emulator()
    prob = make_problem(1500, nranks=2)
    h = prob["h"]
    ctxs = []
    for r in range(2):
        a, uid = lattice(prob, r)
        c = sph.Context(prob["tank_w"], prob["tank_h"], h, 4096, msg_capacity=2048, rank=r, nranks=2)
        t = sph.default_params(h, prob["tank_w"], prob["tank_h"])
        t.node_start_x, t.node_end_x = prob["slabs"][r][2], prob["slabs"][r][3]
        c.set_params(t); c.upload(a, uid)
        ctxs.append((c, t))

    def exchange(which):
        bufs = [[np.ctypeslib.as_array((sph.C.c_ubyte * nb).from_address(p)) for p in ptrs]
                for ptrs, nb in (c.exchange_pointers(which) for c, _ in ctxs)]
        bufs[1][1][:] = bufs[0][2]      # rank 1 receives from its left what rank 0 sent to its right
        bufs[0][3][:] = bufs[1][0]      # and the other way round

    shift = 3 * h
    for step in range(12):
        for c, _ in ctxs:
            c.advect()
        if step == 5:
            for r, (c, t) in enumerate(ctxs):
                t2 = t.copy()
                if r == 0: t2.node_end_x = t.node_end_x + shift
                else: t2.node_start_x = t.node_start_x + shift
                with pytest.raises(sph.SphError):
                    c.set_edges(t2.node_start_x, t2.node_end_x)
                c.set_params(t2)
                ctxs[r] = (c, t2)
        exchange(0)
        for c, _ in ctxs:
            c.sort(); c.density(); c.relax()
        exchange(1)
        for c, _ in ctxs:
            c.sort()
    uid = np.concatenate([c.download()[1] for c, _ in ctxs])
    assert np.array_equal(np.sort(uid), np.arange(1508)), len(uid)
    for c, _ in ctxs:
        st = c.status()
        assert st.capacity_overflow == 0 and st.msg_overflow == 0
    # the edges did move: rank 0 now owns more than its half
    assert ctxs[0][0].status().n_local > 1508 // 2 + 100


@pytest.mark.parametrize("margin,world", [(None, 3), ("-3000", 2)])
def test_bench_e2e_leg_on_emulated_slabs(tmp_path, built_lib, monkeypatch, margin, world):
    """SlabRunner.e2e as bench.py calls it at N > 1 (blocks of frames from the restored state, both protocols): the
    pipelined figure must be the one reported (its last frame equals the synchronous feed's), every rank must time the
    same number of steps, and running the leg twice must end in the same state bit for bit (the restore works under
    the pipelined coordinate feed).  Three slabs: the middle one has two neighbours, like every interior slab of the 4-
    and 8-GPU runs."""
    # margin -3000: the asynchronous copy brings fewer entries than the slab holds (~6000), so that every frame takes the
    # path of a slab that outgrew its estimate -- sph_coords_wait fetches the remainder from the ticket's own device frame
    if margin:
        monkeypatch.setenv("SPH_FEED_MARGIN_ENTRIES", margin)
    monkeypatch.setenv("SPH_EMU_E2E", "1")
    monkeypatch.setenv("SPH_EMU_XPERIOD", "2")
    monkeypatch.setenv("SPH_EMU_DEFINES", "SPH_ONE_EXCHANGE=1")
    monkeypatch.setenv("SPH_EMU_HALO_WIDTH", "7.0")
    parts = run_world(tmp_path, world, 12000, 12, True, "emu_block")
    assert len(parts) == world
    for p in parts:
        out = eval(str(p["e2e"][0]))
        assert out["pipelined"] is True and "pipelined_error" not in out, out
        assert out["steps"] == 16 and out["seconds"] > 0 and out["sync_seconds"] > 0, out
        assert out["repeatable"] is True, out
        # D2H accounting from what the library copied: about the slab's population (+ 1/8 + margin), not its capacity
        # (bytes per step = entries per frame here: 4 bytes per entry, 4 steps per frame; a slab holds ~6000 particles;
        #  the default margin of 4096 entries dominates at this size)
        lo, hi = (4500, 7500) if margin else (7000, 10500)        # (three slabs: ~4000 particles each)
        assert lo < out["d2h_per_step"] < hi, out


@pytest.mark.parametrize("exchanges", [2, 1])
def test_slabs_equal_one_slab_while_no_particle_is_thrown_past_the_ghost_layer(built_lib, monkeypatch, exchanges):
    """The condition behind "N slabs == 1 slab bit for bit" (DESIGN.md 6), pinned from both sides.  Ordinary motion is
    bounded by the velocity clamp (0.07 h per step, fluid.c:613-625); the one unbounded displacement on the path is the
    mover's push-out (fluid.c:663-685).  A sphere of radius 3.4 h whose centre sits 1 h beyond a slab edge throws the
    particles between the edge and its centre 2.4 h deep into the neighbour's slab.  If the mover ARRIVES at a walk (1 h per
    frame: 8 x the render rank's autopilot step, renderer.c:513-531) the fluid gives way ahead of it and the slabs stay
    bit-identical to one slab; if it is teleported into the fluid in a single step (a mouse click, controls.c), particles are
    thrown further than the ghost layer reaches, their owner relaxes them for one step without their new neighbours, and a
    few come out differently from the one-slab run -- as they do in the reference, whose halo is h wide
    (communication.c:137-140).  Nobody is lost, nothing overflows."""
    import sph_b200
    from emu.backend import use_emulator
    monkeypatch.setattr(sph_b200, "_lib", sph_b200._lib)      # use_emulator() rebinds the module's library: undone after the test
    sph = use_emulator()
    n_req = 6000
    tank_w = 15.0 * float(np.sqrt(n_req / 1500.0))
    prob, p1 = sph.make_problem(n_req, tank_w=tank_w, nranks=2), sph.make_problem(n_req, tank_w=tank_w)
    h = prob["h"]
    edges = [(s, e) for (_, _, s, e) in prob["slabs"]]

    def run(jump_h):
        t0 = sph.default_params(h, prob["tank_w"], prob["tank_h"], "x")
        radius = 0.5 * t0.mover_width
        assert radius > 3.2 * h
        t0.mover_center_x = edges[0][1] + h
        t0.mover_center_y = -2.0 * radius                    # below the floor: touches nothing yet
        ctxs = []
        for r in range(2):
            c = sph.Context(prob["tank_w"], prob["tank_h"], h, 2 * prob["n_global"] + 4096, msg_capacity=4096, rank=r, nranks=2,
                            halo_width=0.0 if exchanges == 2 else 3.5, exchanges_per_step=0 if exchanges == 2 else 1)
            t = t0.copy(); t.node_start_x, t.node_end_x = edges[r]
            c.set_params(t); c.init_lattice(prob, r)
            ctxs.append(c)
        one = sph.Context(p1["tank_w"], p1["tank_h"], h, p1["n_global"] + 64)
        one.set_params(t0); one.init_lattice(p1)

        def exchange(which):
            bufs = [[np.ctypeslib.as_array((sph.C.c_ubyte * nb).from_address(p)) for p in ptrs]
                    for ptrs, nb in (c.exchange_pointers(which) for c in ctxs)]
            bufs[1][1][:] = bufs[0][2]
            bufs[0][3][:] = bufs[1][0]

        tcur = t0.copy()
        for frame in range(10):
            for sub in range(4):
                if sub == 3 and frame >= 4:                  # the frame's parameter block (fluid.c:293-294)
                    tcur = tcur.copy()
                    tcur.mover_center_y = min(tcur.mover_center_y + jump_h * h, 0.3 * prob["tank_h"])
                    for r, c in enumerate(ctxs):
                        t = tcur.copy(); t.node_start_x, t.node_end_x = edges[r]
                        c.queue_params(t)
                    one.queue_params(tcur)
                for c in ctxs:
                    c.advect()
                exchange(0)
                for c in ctxs:
                    c.sort(); c.density(); c.relax()
                if exchanges == 2:
                    exchange(1)
                for c in ctxs:
                    c.sort()
                one.step(1)
        parts = [c.download() for c in ctxs]
        uid = np.concatenate([p[1] for p in parts]); state = np.concatenate([p[0] for p in parts])
        ref, ru = one.download()
        assert np.array_equal(np.sort(uid), ru), "particles lost or duplicated"
        for c in ctxs:
            st = c.status()
            assert st.capacity_overflow == 0 and st.msg_overflow == 0
        o = np.argsort(uid)
        d = np.maximum(np.abs(state["x"][o] - ref["x"]), np.abs(state["y"][o] - ref["y"])) / h
        moved = float(np.abs(ref["y"] - sph.lattice(p1)[0]["y"]).max() / h)
        return int((d > 0).sum()), float(d.max()), moved

    differing, worst, moved = run(1.0)
    assert differing == 0 and moved > 1.0, (differing, worst, moved)         # the mover did plough through the fluid
    differing, worst, _ = run(100.0)
    assert 0 < differing < 0.05 * prob["n_global"] and worst < 1.0, (differing, worst)


def test_parked_slab_keeps_its_waiting_emigrants_between_two_exchanges(built_lib, monkeypatch):
    """Exchange period 2 (the multi-GPU bench's default) and a slab that is parked (controls.c:405-426) with more
    particles than one message takes: the emigrants that wait for the next exchange must survive the step in between,
    in which nobody is handed over.  Found by tests/fuzz/fuzz_slabs.py: the prediction kernel of such a step dropped
    every local outside the slab's window -- all that the parked slab had not sent yet (792 of 6068 particles in the run
    that showed it), reported as capacity_overflow.  Those steps now run the HOLD instantiation of k_advect."""
    import sph_b200
    from emu.backend import use_emulator
    monkeypatch.setattr(sph_b200, "_lib", sph_b200._lib)      # use_emulator() rebinds the module's library: undone after the test
    sph = use_emulator()
    K, n_req, period = 3, 6000, 2
    tank_w = 15.0 * float(np.sqrt(n_req / 1500.0))
    prob = sph.make_problem(n_req, tank_w=tank_w, nranks=K)
    h = prob["h"]
    edges = [(s, e) for (_, _, s, e) in prob["slabs"]]
    t0 = sph.default_params(h, prob["tank_w"], prob["tank_h"], "x")
    t0.mover_center_y = -10.0 * prob["tank_h"]               # out of the way
    ctxs = []
    for r in range(K):
        c = sph.Context(prob["tank_w"], prob["tank_h"], h, 2 * prob["n_global"] + 4096, msg_capacity=1400, rank=r, nranks=K,
                        halo_width=3.5 * period, exchanges_per_step=1)
        c.set_exchange_period(period)
        t = t0.copy(); t.node_start_x, t.node_end_x = edges[r]
        c.set_params(t); c.init_lattice(prob, r)
        ctxs.append(c)
    assert ctxs[K - 1].status().n_local > 1400               # more than one message takes

    def exchange():
        bufs = [[np.ctypeslib.as_array((sph.C.c_ubyte * nb).from_address(p)) for p in ptrs]
                for ptrs, nb in (c.exchange_pointers(0) for c in ctxs)]
        for r in range(K - 1):
            bufs[r + 1][1][:] = bufs[r][2]
            bufs[r][3][:] = bufs[r + 1][0]

    n_active, populations = K, []
    for step in range(24):
        if step == 7:                                         # the frame's parameter block parks the last slab
            edges, n_active = sph.remove_partition(edges, h, n_active)
            for r, c in enumerate(ctxs):
                t = t0.copy(); t.node_start_x, t.node_end_x = edges[r]; t.active = bytes([1 if r < n_active else 0])
                c.queue_params(t)
        for c in ctxs:
            c.advect()
        if all(c.exchange_due for c in ctxs):
            exchange()
        else:
            assert not any(c.exchange_due for c in ctxs)
        for c in ctxs:
            c.sort(); c.density(); c.relax(); c.sort()
        populations.append(ctxs[K - 1].status().n_local)
    uid = np.concatenate([c.download()[1] for c in ctxs])
    assert np.array_equal(np.sort(uid), np.arange(prob["n_global"])), (len(uid), populations)      # nobody lost, nobody duplicated
    assert all(c.status().capacity_overflow == 0 for c in ctxs)
    assert ctxs[K - 1].status().msg_overflow > 0             # emigrants did have to wait ...
    waiting = [p for p in populations[7:] if 0 < p < populations[6]]
    assert len(waiting) >= 2, populations                    # ... through at least one step without an exchange
    assert populations[-1] == 0, populations                 # and the parked slab did drain


def test_a_slab_must_keep_a_layer_of_its_old_extent_when_both_of_its_edges_move(built_lib, monkeypatch):
    """The second half of the condition behind "N slabs == 1 slab" (DESIGN.md 6), pinned from both sides: in the step in
    which new edges land, a slab's neighbour needs as ghosts a strip that must already be that slab's.  Three slabs, the
    middle one 2.75 h wide, two exchanges per step (2 h layer), both of its edges moved 2 h to the left in one parameter
    block: its width stays, but its old and new extents overlap by 0.75 h only -- most of what the right slab needs as
    ghosts is, for that one step, still owned by the LEFT slab, which is not its neighbour; a few particles beside the
    new edge come out ulps away from the one-slab run (soak run 93072 of tests/fuzz/fuzz_slabs.py).  With the middle
    slab 4.25 h wide (layer + move, and a quarter) the same move is bit-identical.  sph_host_balance_time and
    slab.keep_slabs_wider_than refuse the first move since (tests/test_host.py); nobody is lost either way."""
    import sph_b200
    from emu.backend import use_emulator
    monkeypatch.setattr(sph_b200, "_lib", sph_b200._lib)      # use_emulator() rebinds the module's library: undone after the test
    sph = use_emulator()
    n_req = 3000
    tank_w = 15.0 * float(np.sqrt(n_req / 750.0))
    p1 = sph.make_problem(n_req, tank_w=tank_w, water_frac=0.5)
    h = p1["h"]
    a0, uid0 = sph.lattice(p1)

    def run(width_h):
        t0 = sph.default_params(h, p1["tank_w"], p1["tank_h"], "x")
        # a sphere that sits in the fluid from the start, across the edges that will move: it stirs the lattice (on the
        # undisturbed lattice the missing ghosts are exactly h away from the nearest local and weigh nothing)
        t0.mover_center_x = 12.5 * h; t0.mover_center_y = 0.5 * p1["tank_h"]
        left = 14.5 * h - width_h * h
        edges = [(0.0, left), (left, 14.5 * h), (14.5 * h, p1["tank_w"])]
        ctxs = []
        for r in range(3):
            c = sph.Context(p1["tank_w"], p1["tank_h"], h, 2 * p1["n_global"] + 4096, msg_capacity=4096, rank=r, nranks=3)
            t = t0.copy(); t.node_start_x, t.node_end_x = edges[r]
            c.set_params(t)
            mine = (a0["x"] >= edges[r][0]) & (a0["x"] < edges[r][1]) if r < 2 else a0["x"] >= edges[r][0]
            c.upload(a0[mine], uid0[mine])
            ctxs.append(c)
        one = sph.Context(p1["tank_w"], p1["tank_h"], h, p1["n_global"] + 64)
        one.set_params(t0); one.upload(a0, uid0)

        def exchange(which):
            bufs = [[np.ctypeslib.as_array((sph.C.c_ubyte * nb).from_address(p)) for p in ptrs]
                    for ptrs, nb in (c.exchange_pointers(which) for c in ctxs)]
            for r in range(2):
                bufs[r + 1][1][:] = bufs[r][2]
                bufs[r][3][:] = bufs[r + 1][0]

        for frame in range(9):
            for sub in range(4):
                if sub == 3 and frame == 6:                  # one parameter block: both edges of the middle slab 2 h to the left
                    edges = [(0.0, edges[0][1] - 2.0 * h), (edges[1][0] - 2.0 * h, edges[1][1] - 2.0 * h), (edges[2][0] - 2.0 * h, edges[2][1])]
                    for r, c in enumerate(ctxs):
                        t = t0.copy(); t.node_start_x, t.node_end_x = edges[r]
                        c.queue_params(t)
                    one.queue_params(t0)
                for c in ctxs:
                    c.advect()
                exchange(0)
                for c in ctxs:
                    c.sort(); c.density(); c.relax()
                exchange(1)
                for c in ctxs:
                    c.sort()
                one.step(1)
        parts = [c.download() for c in ctxs]
        uid = np.concatenate([p[1] for p in parts]); state = np.concatenate([p[0] for p in parts])
        ref, ru = one.download()
        assert np.array_equal(np.sort(uid), ru), "particles lost or duplicated"
        for c in ctxs:
            st = c.status()
            assert st.capacity_overflow == 0 and st.msg_overflow == 0
        o = np.argsort(uid)
        d = np.maximum(np.abs(state["x"][o] - ref["x"]), np.abs(state["y"][o] - ref["y"])) / h
        return int((d > 0).sum()), float(d.max())

    differing, worst = run(4.25)
    assert differing == 0, (differing, worst)
    differing, worst = run(2.75)
    assert 0 < differing < 0.2 * p1["n_global"] and worst < 1e-2, (differing, worst)


def test_a_mover_resting_on_a_wall_beside_an_edge_walks_particles_across_it_after_the_migration(built_lib, monkeypatch):
    """A third face of the condition behind "N slabs == 1 slab" (DESIGN.md 6), pinned from both sides.  A sphere that
    stands in the fluid from the start is harmless (the first prediction's boundary pass throws what it covers BEFORE
    that step's migration) -- unless it also rests on a WALL next to a slab edge: a particle under it is pushed out
    radially, i.e. below the floor, the tank clamp (fluid.c:656-744) puts it back on the floor INSIDE the sphere, and the
    next boundary pass -- the relaxation's, after the migration -- pushes it out again, now almost horizontally: it walks
    along the floor to the sphere's rim, across the edge and 2.5 h beyond it, while it still belongs to the slab it came
    from, whose ghost layer (2 h) does not reach its new neighbours.  Soak runs 401053 / 420879 / 430956 of
    tests/fuzz/fuzz_slabs.py (3 of 4000): 16 particles ulps away from the one-slab run in the second step.  The same
    sphere 8 h above the floor: bit-identical.  Nobody is lost either way."""
    import sph_b200
    from emu.backend import use_emulator
    monkeypatch.setattr(sph_b200, "_lib", sph_b200._lib)      # use_emulator() rebinds the module's library: undone after the test
    sph = use_emulator()
    n_req = 6000
    tank_w = 15.0 * float(np.sqrt(n_req / 750.0))
    prob = sph.make_problem(n_req, tank_w=tank_w, water_frac=0.5, nranks=2)
    p1 = sph.make_problem(n_req, tank_w=tank_w, water_frac=0.5)
    h = prob["h"]
    edges = [(s, e) for (_, _, s, e) in prob["slabs"]]
    assert abs(edges[0][1] / h - 18.0) < 1e-3

    def run(my_h):
        t0 = sph.default_params(h, prob["tank_w"], prob["tank_h"], "b")
        t0.mover_center_x = 15.694 * h; t0.mover_center_y = my_h * h      # 2.3 h left of the edge
        ctxs = []
        for r in range(2):
            c = sph.Context(prob["tank_w"], prob["tank_h"], h, 2 * prob["n_global"] + 4096, msg_capacity=4096, rank=r, nranks=2)
            t = t0.copy(); t.node_start_x, t.node_end_x = edges[r]
            c.set_params(t); c.init_lattice(prob, r)
            ctxs.append(c)
        one = sph.Context(p1["tank_w"], p1["tank_h"], h, p1["n_global"] + 64)
        one.set_params(t0); one.init_lattice(p1)

        def exchange(which):
            bufs = [[np.ctypeslib.as_array((sph.C.c_ubyte * nb).from_address(p)) for p in ptrs]
                    for ptrs, nb in (c.exchange_pointers(which) for c in ctxs)]
            bufs[1][1][:] = bufs[0][2]
            bufs[0][3][:] = bufs[1][0]

        for step in range(2):
            for c in ctxs:
                c.advect()
            exchange(0)
            for c in ctxs:
                c.sort(); c.density(); c.relax()
            exchange(1)
            for c in ctxs:
                c.sort()
            one.step(1)
        parts = [c.download() for c in ctxs]
        uid = np.concatenate([p[1] for p in parts]); state = np.concatenate([p[0] for p in parts])
        ref, ru = one.download()
        assert np.array_equal(np.sort(uid), ru), "particles lost or duplicated"
        for c in ctxs:
            st = c.status()
            assert st.capacity_overflow == 0 and st.msg_overflow == 0
        o = np.argsort(uid)
        d = np.maximum(np.abs(state["x"][o] - ref["x"]), np.abs(state["y"][o] - ref["y"])) / h
        return int((d > 0).sum()), float(d.max())

    differing, worst = run(8.0)
    assert differing == 0, (differing, worst)
    differing, worst = run(0.989)
    assert 0 < differing < 100 and worst < 1e-2, (differing, worst)
