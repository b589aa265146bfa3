"""TEST INFRASTRUCTURE -- randomised runs of the PEER-MEMORY slab protocol on the kernel-source emulator.

Per seed: 2-3 emulated devices (host threads of this process; a cudaIpc handle is just the pointer) map each other's
exchange blocks and run one script of calls -- sph_step(n) in pieces of random length (captured graphs: message
sequence numbers, arrival flags, device-side waits, parity double buffering), queued parameter blocks with moved slab
edges and a walking mover, snapshots and restores at barriers (as bench.py does between its timed blocks), a last slab
that is parked and re-added -- two exchanges per step or one exchange every 1 / 2 / 4 steps.  The slabs drift apart as far
as the protocol lets them (a random sleep before every call).  Result against ONE slab given the same script: bit for bit.
    python tests/fuzz/fuzz_p2p.py FIRST_SEED COUNT"""
import ctypes as C
import os
import random
import sys
import threading
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(os.path.dirname(HERE)), os.path.dirname(HERE)]
os.environ.setdefault("SPH_SPIN_TIMEOUT_MS", "60000")
import sph_b200 as sph  # noqa: E402
from emu.build_emu import build  # noqa: E402

sph._lib = sph._bind(C.CDLL(build()))


def run(seed):
    rng = random.Random(seed)
    K = rng.choice([2, 3])
    E = rng.choice([0, 1, 2, 4])                       # 0: two exchanges per step
    n_req = rng.choice([1500, 3000])
    water = rng.choice([0.5, 1.0])
    goo = rng.random() < 0.25
    tank_w = 15.0 * float(np.sqrt(n_req / (1500.0 * water)))
    prob = sph.make_problem(n_req, tank_w=tank_w, water_frac=water, nranks=K)
    p1 = sph.make_problem(n_req, tank_w=tank_w, water_frac=water)
    h = prob["h"]
    layer = (3.0 if goo else 2.0) if E == 0 else (4.5 if goo else 3.5) * E
    edges = [(s, e) for (_, _, s, e) in prob["slabs"]]
    if min(e - s for s, e in edges) < (layer + 0.6) * h:
        return "skipped (slabs narrower than the layer)"
    elastic = K == 3 and rng.random() < 0.4
    # a parked slab with more particles than a message takes: its emigrants wait, also through the steps between two
    # exchanges (the HOLD instantiation of k_advect); such a run is no longer the one-slab run -- conservation only
    small = elastic and rng.random() < 0.5
    msg_cap = 700 if small else 4096
    t0 = sph.default_params(h, prob["tank_w"], prob["tank_h"], "y" if goo else rng.choice(["x", "a", "b"]))
    t0.mover_center_x = rng.random() * prob["tank_w"]; t0.mover_center_y = rng.random() * prob["tank_h"]

    # ---- one script for everybody
    script, tcur, n_active, saved, total = [], t0.copy(), K, False, 0
    for _ in range(rng.randint(10, 24)):
        op = rng.choice(["step", "step", "queue", "queue", "save", "restore", "park"])
        if op == "step":
            n = rng.randint(1, 9); script.append(("step", n)); total += n
        elif op == "queue":
            tcur = tcur.copy()
            tcur.mover_center_x = min(max(tcur.mover_center_x + (2 * rng.random() - 1) * h, 0.0), prob["tank_w"])
            tcur.mover_center_y = min(max(tcur.mover_center_y + (2 * rng.random() - 1) * h, 0.0), prob["tank_h"])
            new = list(edges)
            for r in range(n_active - 1):
                d = rng.choice([0, 1, -1, 4, -4]) * 0.125 * h
                e = new[r][1] + d
                if e - new[r][0] >= (layer + 0.6) * h and new[r + 1][1] - e >= (layer + 0.6) * h:
                    new[r] = (new[r][0], e); new[r + 1] = (e, new[r + 1][1])
            edges = new
            script.append(("queue", tcur.copy(), list(edges), n_active))
            n = rng.randint(1, 5); script.append(("step", n)); total += n       # a queued block lands in the next step
        elif op == "save":
            script.append(("save", tcur.copy(), list(edges), n_active)); saved = True
        elif op == "restore" and saved:
            script.append(("restore",))
            _, tcur, edges, n_active = next(s for s in reversed(script) if s[0] == "save")
            tcur, edges = tcur.copy(), list(edges)
        elif op == "park" and elastic:
            if n_active == K:
                edges, n_active = sph.remove_partition(list(edges), h, n_active)
            else:
                edges, n_active = sph.add_partition(list(edges), h, n_active)
            script.append(("queue", tcur.copy(), list(edges), n_active))
            n = rng.randint(2, 6); script.append(("step", n)); total += n

    ctxs = []
    for r in range(K):
        c = sph.Context(prob["tank_w"], prob["tank_h"], h, 2 * prob["n_global"] + 4096, msg_capacity=msg_cap, device=r, rank=r, nranks=K,
                        halo_width=(layer if goo else 0.0) if E == 0 else layer, exchanges_per_step=1 if E else 0)
        if E > 1:
            c.set_exchange_period(E)
        t = t0.copy(); t.node_start_x, t.node_end_x = prob["slabs"][r][2], prob["slabs"][r][3]
        c.set_params(t); c.init_lattice(prob, r)
        ctxs.append(c)
    handles = [c.p2p_handle() for c in ctxs]
    for r, c in enumerate(ctxs):
        c.p2p_connect(handles[r - 1] if r > 0 else None, handles[r + 1] if r < K - 1 else None)
    one = sph.Context(p1["tank_w"], p1["tank_h"], h, p1["n_global"] + 64)
    one.set_params(t0); one.init_lattice(p1)

    errors, barrier = [], threading.Barrier(K)

    def play(c, r, jitter):
        try:
            for s in script:
                if jitter is not None:
                    time.sleep(jitter.random() * 0.004)
                if s[0] == "step":
                    c.step(s[1])
                elif s[0] == "queue":
                    t = s[1].copy()
                    if r is not None:
                        t.node_start_x, t.node_end_x = s[2][r]; t.active = bytes([1 if r < s[3] else 0])
                    c.queue_params(t)
                elif s[0] == "save":
                    if r is not None:
                        c.synchronize(); barrier.wait(timeout=120)          # all slabs together, as bench.py does it
                    c.state_save()
                elif s[0] == "restore":
                    if r is not None:
                        c.synchronize(); barrier.wait(timeout=120)
                    c.state_restore()
                    if r is not None:
                        c.synchronize(); barrier.wait(timeout=120)
            c.synchronize()
        except Exception as e:       # noqa: BLE001
            errors.append((r, repr(e)))
            barrier.abort()

    threads = [threading.Thread(target=play, args=(c, r, random.Random(1000 * seed + r))) for r, c in enumerate(ctxs)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=600)
    assert not errors, errors
    assert not any(t.is_alive() for t in threads), "a slab is still waiting for its neighbour"
    play(one, None, None)
    parts = [c.download() for c in ctxs]
    uid = np.concatenate([p[1] for p in parts]); state = np.concatenate([p[0] for p in parts])
    ref, ru = one.download()
    bad = [(c.status().capacity_overflow, c.status().msg_overflow, c.status().exchange_timeouts) for c in ctxs]
    assert len(uid) == len(ru) and np.array_equal(np.sort(uid), ru), ("lost or duplicated", len(uid), len(ru), bad)
    assert all((b[0], b[2]) == (0, 0) for b in bad) and (small or all(b[1] == 0 for b in bad)), bad
    order = np.argsort(uid)
    for fld in () if small else ("x", "y", "v_x", "v_y"):
        if not np.array_equal(state[fld][order].view("u4"), ref[fld].view("u4")):
            nbad = int((state[fld][order].view("u4") != ref[fld].view("u4")).sum())
            raise AssertionError(f"{fld} differs for {nbad} particles; K={K} E={E} n={n_req} goo={goo} elastic={elastic} script={[s[0] + (str(s[1]) if s[0] == 'step' else '') for s in script]}")
    for c in ctxs:
        c.close()
    one.close()
    return f"ok K={K} E={E} n={n_req} water={water} goo={goo} elastic={elastic} msg_cap={msg_cap} waited={sum(b[1] for b in bad)} steps={total} ops={len(script)}"


if __name__ == "__main__":
    s0, n = int(sys.argv[1]), int(sys.argv[2])
    failed = 0
    for seed in range(s0, s0 + n):
        try:
            print("seed", seed, run(seed), flush=True)
        except AssertionError as e:
            failed += 1
            print("seed", seed, "FAIL", str(e)[:700], flush=True)
        except sph.SphError as e:
            failed += 1
            print("seed", seed, "SphError", str(e)[:300], flush=True)
    sys.exit(1 if failed else 0)
