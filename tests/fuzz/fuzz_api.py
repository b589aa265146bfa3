"""TEST INFRASTRUCTURE -- randomised call sequences against the C ABI's host logic, on the kernel-source emulator.

Context A gets a random interleaving of state-changing calls (sph_step, the stage API, sph_run_frame, the asynchronous
frame / coordinate feed with up to two tickets in flight, sph_set_params, sph_queue_params, sph_state_save / _restore)
and of observers, including observers and calls that must be REFUSED between the stages of a step; context B gets only
the state-changing calls in their plainest form.  Every frame A hands out must equal B's, and both must end in the same
state bit for bit.      python tests/fuzz/fuzz_api.py FIRST_SEED COUNT
(tests/test_emu_fuzz.py runs a few fixed seeds.)"""
import ctypes as C
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(os.path.dirname(HERE)), os.path.dirname(HERE)]
import sph_b200  # noqa: E402
from emu.build_emu import build  # noqa: E402

sph_b200._lib = sph_b200._bind(C.CDLL(build()))


def mk(n=700, preset="x"):
    prob = sph_b200.make_problem(n, tank_w=15.0 * float(np.sqrt(n / 750.0)), water_frac=0.5)
    t = sph_b200.default_params(prob["h"], prob["tank_w"], prob["tank_h"], preset)
    t.mover_center_x = 0.3 * prob["tank_w"]
    c = sph_b200.Context(prob["tank_w"], prob["tank_h"], prob["h"], prob["n_global"] + 64)
    c.set_params(t)
    c.init_lattice(prob)
    return prob, t, c


def same(a, b):
    (sa, ua), (sb, ub) = a.download(), b.download()
    return np.array_equal(ua, ub) and all(np.array_equal(sa[f].view("u4"), sb[f].view("u4")) for f in ("x", "y", "v_x", "v_y", "x_prev", "y_prev"))


def run(seed, nops=40):
    rng = random.Random(seed)
    prob, t, A = mk()
    _, _, Bc = mk()
    cap = A.capacity
    bufs = [np.zeros(2 * cap, "i2") for _ in range(2)]
    sbuf = np.zeros(2 * cap, "i2")
    pending = []          # tickets in flight on A
    saved = False
    queued = False        # a queued parameter block has not landed yet (it lands inside the next step)
    log = []
    tcur = t.copy()       # the latest block handed to the library (set, queued or carried by a frame)
    tforce = t.copy()     # the block in force (a queued block is not, until a step has run)
    tq = None
    for k in range(nops):
        op = rng.choice(["step", "stages", "frame", "frame_async", "wait", "pack", "pack_async", "status", "download", "pairs",
                         "save", "restore", "set_params", "queue", "frame_none", "cells", "fwd"])
        log.append(op)
        if os.environ.get("FUZZ_LOG"):
            print(k, op, flush=True)
        if op in ("step", "stages", "frame", "frame_async", "frame_none"):
            landed = True
        else:
            landed = False
        try:
            if op == "step":
                n = rng.randint(1, 5)
                A.step(n); Bc.step(n)
            elif op == "stages":
                # observers (and refusals) between the stages of a step must leave the step intact
                def noise():
                    for _ in range(rng.randint(0, 2)):
                        o = rng.choice(["status", "pack", "download", "pairs", "cells", "fwd", "save", "restore", "set_same", "queue_mid", "step_mid", "pack_async"])
                        if os.environ.get("FUZZ_LOG"):
                            print("   noise", o, flush=True)
                        try:
                            if o == "status": A.status()
                            elif o == "pack": A.pack_coords()
                            elif o == "download": A.download()
                            elif o == "pairs": A.pairs()
                            elif o == "cells": A.cells()
                            elif o == "fwd": A.forward_counts()
                            elif o == "save": A.state_save(); raise AssertionError("state_save accepted in mid-step")
                            elif o == "restore" and saved: A.state_restore(); raise AssertionError("state_restore accepted in mid-step")
                            elif o == "set_same" and not was_queued: A.set_params(tforce)      # (a no-op only while nothing is queued)
                            elif o == "step_mid": A.step(1); raise AssertionError("sph_step accepted in mid-step")
                            elif o == "pack_async" and len(pending) < 2:
                                idx = 1 - pending[-1][2] if pending else 0
                                tk = A.pack_coords_async(bufs[idx])
                                pending.append((tk, bufs[idx], idx, None))
                        except sph_b200.SphError:
                            pass
                was_queued = queued
                A.advect(); noise(); A.sort(); noise(); A.density(); noise(); A.relax(); noise(); A.sort()
                Bc.step(1)
            elif op == "frame":
                n = A.run_frame(tcur, 4, sbuf)
                Bc.step(3); Bc.queue_params(tcur); Bc.step(1)
                ref = Bc.pack_coords().ravel()
                assert np.array_equal(sbuf[:2 * n], ref[:2 * n]), "sync frame differs"
            elif op == "frame_none":
                A.run_frame(None, 2, None); Bc.step(2)
            elif op == "frame_async":
                if len(pending) == 2:
                    tk, *_ = pending.pop(0)
                    A.coords_wait(tk)
                buf = bufs[len(pending) and (1 - pending[-1][2]) or 0]
                idx = 0 if buf is bufs[0] else 1
                tk = A.run_frame_async(tcur, 4, buf)
                Bc.step(3); Bc.queue_params(tcur); Bc.step(1)
                pending.append((tk, buf, idx, Bc.pack_coords().ravel().copy()))
            elif op == "pack_async":
                if len(pending) == 2:
                    tk, *_ = pending.pop(0)
                    A.coords_wait(tk)
                idx = 1 - pending[-1][2] if pending else 0
                tk = A.pack_coords_async(bufs[idx])
                pending.append((tk, bufs[idx], idx, Bc.pack_coords().ravel().copy()))
            elif op == "wait":
                if pending:
                    tk, buf, idx, ref = pending.pop(0)
                    n = A.coords_wait(tk)
                    assert ref is None or np.array_equal(buf[:2 * n], ref[:2 * n]), "async frame differs"
            elif op == "pack":
                a = A.pack_coords(); b = Bc.pack_coords()
                assert np.array_equal(a, b), "pack differs"
                # the synchronous pack waits for pending copies but leaves tickets pending
            elif op == "status":
                sa, sb = A.status(), Bc.status()
                assert sa.n_local == sb.n_local and sa.max_bucket == sb.max_bucket
            elif op == "download":
                assert same(A, Bc), "state differs"
            elif op == "pairs":
                assert len(A.pairs()) == len(Bc.pairs())
            elif op == "cells":
                A.cells()
            elif op == "fwd":
                A.forward_counts()
            elif op == "save":
                try:
                    A.state_save()
                except sph_b200.SphError as e:
                    assert "queued parameter block" in str(e), e
                    continue
                Bc.state_save(); saved = True
                t_saved = tforce.copy()
            elif op == "restore":
                if saved:
                    A.state_restore(); Bc.state_restore()
                    tcur = t_saved.copy(); tforce = t_saved.copy()       # the snapshot holds the parameter block in force too
                    queued = False
            elif op == "set_params":
                tcur = tcur.copy(); tcur.k = 0.2 + 0.1 * rng.random(); tcur.mover_center_x = prob["tank_w"] * rng.random()
                A.set_params(tcur); Bc.set_params(tcur)
                tforce = tcur.copy()
            elif op == "queue":
                tcur = tcur.copy(); tcur.sigma = 20.0 * rng.random(); tcur.mover_center_y = prob["tank_h"] * rng.random()
                A.queue_params(tcur); Bc.queue_params(tcur)
                queued = True; tq = tcur.copy()
        except sph_b200.SphError as e:
            print("seed", seed, "op", k, op, "SphError", e, "log", log[-8:])
            raise
        if landed:
            if queued:
                tforce = tq.copy()
            if op in ("frame", "frame_async"):
                tforce = tcur.copy()             # the frame's own block landed in its last step
            queued = False
        if os.environ.get("FUZZ_CHECK"):
            assert same(A, Bc), ("state differs after op", k, op)
    while pending:
        tk, buf, idx, ref = pending.pop(0)
        n = A.coords_wait(tk)
        assert ref is None or np.array_equal(buf[:2 * n], ref[:2 * n]), "async frame differs at the end"
    assert same(A, Bc), ("final state differs", log)
    A.close(); Bc.close()


if __name__ == "__main__":
    s0 = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    failed = 0
    for seed in range(s0, s0 + int(sys.argv[2]) if len(sys.argv) > 2 else s0 + 10):
        try:
            run(seed, nops=int(os.environ.get("FUZZ_OPS", "40")))
            print("seed", seed, "ok", flush=True)
        except AssertionError as e:
            failed += 1
            print("seed", seed, "FAIL", str(e)[:400], flush=True)
        except sph_b200.SphError:        # (already printed with the calls that led to it)
            failed += 1
    sys.exit(1 if failed else 0)
