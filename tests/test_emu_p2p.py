"""The peer-memory exchange protocol without NVLink: two emulated devices (two host threads of this process, each
running the CUDA source compiled for the host, tests/emu) map each other's exchange block and run whole slab
steps from captured graphs -- send_messages, the last-block publication, the arrival flags and the device-side
waits of k_unpack, double buffering by step parity.  The result must equal the single-slab run bit for bit.
Timing, co-residency and the memory model of real GPUs are out of reach here (tests/test_gpu_slabs.py)."""
import ctypes as C
import threading

import numpy as np
import pytest

import sph_b200
from emu.build_emu import build as build_emu
from oracle.oracle import default_tunable, lattice, make_problem
from test_gpu_parity import as_sph


@pytest.fixture(autouse=True)
def spin_timeout(monkeypatch):
    monkeypatch.setenv("SPH_SPIN_TIMEOUT_MS", "60000")
    yield


@pytest.mark.parametrize("world,goo,one_exchange,period", [(2, False, False, 1), (3, True, False, 1), (3, False, True, 1), (2, True, True, 1),
                                                           (3, False, True, 2), (2, False, True, 4), (2, True, True, 2)])
def test_peer_memory_slab_steps_equal_single_slab_bit_for_bit(built_lib, monkeypatch, world, goo, one_exchange, period):
    """period > 1 (sph_set_exchange_period): the graph of a step exists with and without the exchange kernel, and the
    message sequence numbers / buffer parity advance with the EXCHANGES, not with the steps."""
    # one exchange per step is a mode of the context (sph_config.exchanges_per_step = 1) or the default of a
    # -DSPH_ONE_EXCHANGE=1 build: the cases with a period of 2 take the first route on the DEFAULT build
    runtime_mode = one_exchange and period == 2
    lib = build_emu(defines=("SPH_ONE_EXCHANGE=1",), name="libsph_emu_sph_one_exchange1.so") if one_exchange and not runtime_mode else build_emu()
    monkeypatch.setattr(sph_b200, "_lib", sph_b200._bind(C.CDLL(lib)))
    halo = (4.5 if goo else 3.5) * period if one_exchange else 2.0
    n_req, steps = 3000, 80
    tank_w = 15.0 * float(np.sqrt(n_req / 750.0))
    prob = make_problem(n_req, tank_w=tank_w, water_frac=0.5, nranks=world)
    p1 = make_problem(n_req, tank_w=tank_w, water_frac=0.5)

    def params(p, rank=None):
        t = default_tunable(p["h"], p["tank_w"], p["tank_h"], preset="y" if goo else "x")
        t.mover_center_x = 0.3 * p["tank_w"]
        if rank is not None:
            t.node_start_x, t.node_end_x = p["slabs"][rank][2], p["slabs"][rank][3]
        return as_sph(t)

    ctxs = []
    for r in range(world):
        a, uid = lattice(prob, r)
        c = sph_b200.Context(prob["tank_w"], prob["tank_h"], prob["h"], 2 * len(a) + 4096, msg_capacity=2048,
                             device=r, rank=r, nranks=world, halo_width=halo, exchanges_per_step=1 if runtime_mode else 0)
        assert c.exchanges_per_step == (1 if one_exchange else 2)
        if period > 1:
            c.set_exchange_period(period)
        if goo:
            c.set_viscosity_stabilisation(0.5)
        c.set_params(params(prob, r))
        c.upload(a, uid)
        ctxs.append(c)
    handles = [c.p2p_handle() for c in ctxs]
    for r, c in enumerate(ctxs):
        c.p2p_connect(handles[r - 1] if r > 0 else None, handles[r + 1] if r < world - 1 else None)

    errors = []

    def changed(ts):                 # the render rank's scatter in mid-run: mover moved, stiffer fluid (fluid.c:293-294)
        t2 = ts.copy(); t2.mover_center_x = 0.45 * prob["tank_w"]; t2.mover_center_y = 0.2 * prob["tank_h"]; t2.k = 0.35
        return t2

    def run(c, r):
        try:
            for k in range(steps // 8):
                if k == 4:
                    c.queue_params(changed(params(prob, r)))     # lands inside the next step, on every slab alike
                c.step(8)            # graph replays; neighbours drift apart by at most one exchange
            c.synchronize()
        except Exception as e:       # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=run, args=(c, r)) for r, c in enumerate(ctxs)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not errors, errors
    assert not any(t.is_alive() for t in threads), "a slab is still waiting for its neighbour"

    parts = [c.download() for c in ctxs]
    state = np.concatenate([p[0] for p in parts]); uid = np.concatenate([p[1] for p in parts])
    if world > 1:
        # launches of a slab: 4 for the upload (k_bin_upload + one sort of three kernels), 9 per step (10 with the goo
        # pass), and one k_unpack per meeting: the slabs met every `period` steps (+ once for the queued block's own step)
        if one_exchange:
            met = ctxs[0].launches - 4 - (9 + (1 if goo else 0)) * steps
            assert steps // period <= met <= steps // period + 2, (met, steps, period)
    for c in ctxs:
        s = c.status()
        assert s.capacity_overflow == 0 and s.msg_overflow == 0, (s.capacity_overflow, s.msg_overflow)
        assert b"timed out" not in sph_b200.lib().sph_last_error(c.h)

    a1, u1 = lattice(p1)
    one = sph_b200.Context(p1["tank_w"], p1["tank_h"], p1["h"], len(a1) + 64)
    if goo:
        one.set_viscosity_stabilisation(0.5)
    one.set_params(params(p1)); one.upload(a1, u1)
    one.step(32); one.queue_params(changed(params(p1))); one.step(steps - 32)
    ref, ru = one.download()
    assert np.array_equal(np.sort(uid), ru), "particles lost or duplicated in migration"
    order = np.argsort(uid)
    for f in ("x", "y", "v_x", "v_y"):
        assert np.array_equal(state[f][order].view("u4"), ref[f].view("u4")), f
    assert sum(c.status().migrated_left + c.status().migrated_right for c in ctxs) >= 0


@pytest.mark.parametrize("one_exchange,moving", [(False, False), (True, False), (False, True), (True, True)])
def test_hostile_soup_across_three_devices(built_lib, monkeypatch, one_exchange, moving):
    """Clustered random particles at rest, coincident pairs -- one inside a slab, two straddling the slab edges exactly
    (on the edge, and one ulp to its left) -- through the peer-memory protocol on three emulated devices: the owner
    rule of the coincident-particle nudge (fluid.c:583-586, hash.c:178-224) and the strict </> of the migration
    test (fluid.c:494-497) must give the single-slab bits.  `moving`: the soup keeps its random velocities -- a
    restart from a moving snapshot -- and every slab calls sph_refresh_ghosts + sph_sort after its upload, so that
    the viscosity pass of the first step sees the neighbours across the edges."""
    from common import random_state
    lib = build_emu(defines=("SPH_ONE_EXCHANGE=1",), name="libsph_emu_sph_one_exchange1.so") if one_exchange else build_emu()
    monkeypatch.setattr(sph_b200, "_lib", sph_b200._bind(C.CDLL(lib)))
    world, n, steps = 3, 3000, 20
    prob = make_problem(n, nranks=world)
    tw, th, h = prob["tank_w"], prob["tank_h"], prob["h"]

    def params(rank=None):
        t = default_tunable(h, tw, th)
        if rank is not None:
            t.node_start_x, t.node_end_x = prob["slabs"][rank][2], prob["slabs"][rank][3]
        return as_sph(t)

    for seed in (1, 2):
        a = random_state(n, tw, th, seed=seed, clustered=True)
        e1, e2 = prob["slabs"][1][2], prob["slabs"][2][2]
        a[10] = a[11]
        a[20]["x"] = a[21]["x"] = e1; a[21]["y"] = a[20]["y"]
        a[22]["x"] = np.nextafter(np.float32(e2), np.float32(0)); a[23]["x"] = e2; a[23]["y"] = a[22]["y"]
        a["x_prev"] = a["x"]; a["y_prev"] = a["y"]; a["id"] = np.arange(n)
        if not moving:
            a["v_x"] = 0; a["v_y"] = 0
        uid = np.arange(n, dtype="u4")
        ctxs = []
        # owner = the first slab whose right edge is not exceeded (adjacent edges come out of partitionProblem by two
        # different expressions, geometry.c:141-147, and may differ in the last bit)
        ends = np.array([prob["slabs"][r][3] for r in range(world)], "f4"); ends[-1] = np.inf
        owner = np.searchsorted(ends, a["x"], side="left")
        for r in range(world):
            own = owner == r
            c = sph_b200.Context(tw, th, h, 2 * n, msg_capacity=4096, device=r, rank=r, nranks=world,
                                 halo_width=3.5 if one_exchange else 2.0)
            c.set_params(params(r)); c.upload(a[own], uid[own]); ctxs.append(c)
        hs = [c.p2p_handle() for c in ctxs]
        for r, c in enumerate(ctxs):
            c.p2p_connect(hs[r - 1] if r > 0 else None, hs[r + 1] if r < world - 1 else None)
        def run(c):
            if moving:
                c.refresh_ghosts(); c.sort()
            c.step(steps)
        threads = [threading.Thread(target=run, args=(c,)) for c in ctxs]
        for t in threads:
            t.start()
        for t in threads:
            t.join(timeout=300)
        parts = [c.download() for c in ctxs]
        st = np.concatenate([p[0] for p in parts]); u = np.concatenate([p[1] for p in parts])
        assert all(c.status().capacity_overflow == 0 and c.status().msg_overflow == 0 for c in ctxs)
        one = sph_b200.Context(tw, th, h, n + 64)
        one.set_params(params()); one.upload(a, uid); one.step(steps)
        ref, ru = one.download()
        assert np.array_equal(np.sort(u), ru), seed
        order = np.argsort(u)
        for f in ("x", "y", "v_x", "v_y"):
            assert np.array_equal(st[f][order].view("u4"), ref[f].view("u4")), (seed, f)


def test_exchange_via_host_callback_equals_single_slab(built_lib, monkeypatch):
    """sph_exchange_via_host: the exchange for a host with a plain (host-memory) sendrecv, here Python queues between
    three threads standing in for MPI_Sendrecv.  Stage by stage through the handle API, as INTEGRATION.md 4 describes."""
    import queue
    monkeypatch.setattr(sph_b200, "_lib", sph_b200._bind(C.CDLL(build_emu())))
    world, n_req, steps = 3, 3000, 40
    tank_w = 15.0 * float(np.sqrt(n_req / 750.0))
    prob = make_problem(n_req, tank_w=tank_w, water_frac=0.5, nranks=world)
    p1 = make_problem(n_req, tank_w=tank_w, water_frac=0.5)
    wires = {(a, b): queue.Queue() for a in range(world) for b in range(world) if abs(a - b) == 1}

    def params(p, rank=None):
        t = default_tunable(p["h"], p["tank_w"], p["tank_h"]); t.mover_center_x = 0.3 * p["tank_w"]
        if rank is not None:
            t.node_start_x, t.node_end_x = p["slabs"][rank][2], p["slabs"][rank][3]
        return as_sph(t)

    ctxs = []
    for r in range(world):
        a, uid = lattice(prob, r)
        c = sph_b200.Context(prob["tank_w"], prob["tank_h"], prob["h"], 2 * len(a) + 4096, msg_capacity=2048, device=r, rank=r, nranks=world)
        c.set_params(params(prob, r)); c.upload(a, uid); ctxs.append(c)
    errors = []

    def run(c, r):
        def sendrecv(send, to_side, nrecv, from_side):
            if send is not None:
                wires[(r, r - 1 if to_side == 0 else r + 1)].put(send)
            return wires[(r - 1 if from_side == 0 else r + 1, r)].get(timeout=120) if nrecv else None
        try:
            for _ in range(steps):
                c.advect(); c.exchange_via_host(0, sendrecv); c.sort(); c.density(); c.relax()
                c.exchange_via_host(1, sendrecv); c.sort()
        except Exception as e:       # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=run, args=(c, r)) for r, c in enumerate(ctxs)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not errors, errors
    parts = [c.download() for c in ctxs]
    state = np.concatenate([p[0] for p in parts]); uid = np.concatenate([p[1] for p in parts])
    assert all(c.status().capacity_overflow == 0 and c.status().msg_overflow == 0 for c in ctxs)
    a1, u1 = lattice(p1)
    one = sph_b200.Context(p1["tank_w"], p1["tank_h"], p1["h"], len(a1) + 64)
    one.set_params(params(p1)); one.upload(a1, u1); one.step(steps)
    ref, ru = one.download()
    assert np.array_equal(np.sort(uid), ru)
    order = np.argsort(uid)
    for f in ("x", "y", "v_x", "v_y"):
        assert np.array_equal(state[f][order].view("u4"), ref[f].view("u4")), f


@pytest.mark.parametrize("one_exchange,period", [(False, 1), (True, 2)])
def test_snapshot_and_restore_on_peer_memory_slabs(built_lib, monkeypatch, one_exchange, period):
    """sph_state_save / sph_state_restore on two slabs: the message sequence numbers keep counting through a restore
    (the neighbours' arrival flags only grow), the first step after it exchanges, and the replayed steps are the
    single-slab steps bit for bit."""
    lib = build_emu(defines=("SPH_ONE_EXCHANGE=1",), name="libsph_emu_sph_one_exchange1.so") if one_exchange else build_emu()
    monkeypatch.setattr(sph_b200, "_lib", sph_b200._bind(C.CDLL(lib)))
    world, n_req = 2, 3000
    tank_w = 15.0 * float(np.sqrt(n_req / 750.0))
    prob = make_problem(n_req, tank_w=tank_w, water_frac=0.5, nranks=world)
    p1 = make_problem(n_req, tank_w=tank_w, water_frac=0.5)

    def params(p, rank=None):
        t = default_tunable(p["h"], p["tank_w"], p["tank_h"])
        t.mover_center_x = 0.3 * p["tank_w"]
        if rank is not None:
            t.node_start_x, t.node_end_x = p["slabs"][rank][2], p["slabs"][rank][3]
        return as_sph(t)

    ctxs = []
    for r in range(world):
        a, uid = lattice(prob, r)
        c = sph_b200.Context(prob["tank_w"], prob["tank_h"], prob["h"], 2 * len(a) + 4096, msg_capacity=2048, device=r, rank=r,
                             nranks=world, halo_width=3.5 * period if one_exchange else 2.0)
        if period > 1:
            c.set_exchange_period(period)
        c.set_params(params(prob, r)); c.upload(a, uid)
        ctxs.append(c)
    handles = [c.p2p_handle() for c in ctxs]
    for r, c in enumerate(ctxs):
        c.p2p_connect(handles[r - 1] if r > 0 else None, handles[r + 1] if r < world - 1 else None)
    errors, results = [], {}

    def run(c, r):
        try:
            c.step(13); c.state_save(); c.step(9)
            first = c.download()
            c.step(4)                       # a future that is undone
            c.state_restore(); c.step(9)
            results[r] = (first, c.download())
            c.synchronize()
        except Exception as e:       # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=run, args=(c, r)) for r, c in enumerate(ctxs)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not errors, errors
    assert not any(t.is_alive() for t in threads)
    # (which slab OWNS a particle near the edge may differ between the two passes when the period is > 1: migration
    #  happens at exchange steps, and the restore forces one; the particles themselves must not differ)
    def union(k):
        st = np.concatenate([results[r][k][0] for r in range(world)]); u = np.concatenate([results[r][k][1] for r in range(world)])
        o = np.argsort(u)
        return st[o], u[o]
    (a, ua), (b, ub) = union(0), union(1)
    assert np.array_equal(ua, ub)
    for f in ("x", "y", "v_x", "v_y"):
        assert np.array_equal(a[f].view("u4"), b[f].view("u4")), f
    a1, u1 = lattice(p1)
    one = sph_b200.Context(p1["tank_w"], p1["tank_h"], p1["h"], len(a1) + 64)
    one.set_params(params(p1)); one.upload(a1, u1); one.step(22)
    ref, ru = one.download()
    state = np.concatenate([results[r][1][0] for r in range(world)]); uid = np.concatenate([results[r][1][1] for r in range(world)])
    order = np.argsort(uid)
    assert np.array_equal(uid[order], ru)
    for f in ("x", "y", "v_x", "v_y"):
        assert np.array_equal(state[f][order].view("u4"), ref[f].view("u4")), f
