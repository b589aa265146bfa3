"""One rank of the CPU (gloo) slab test: the multi-rank driver sph_b200.slab.SlabRunner with the gather
oracle standing in for the CUDA library.  Launched by test_slab_gloo.py, one process per rank."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch.distributed as dist

    from oracle.oracle import GatherOracle, default_tunable, make_problem
    from sph_b200.slab import SlabRunner

    out, n_req, steps, balance = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    block = len(sys.argv) > 5 and sys.argv[5] in ("block", "block_cost")
    policy = "cost" if len(sys.argv) > 5 and sys.argv[5] == "block_cost" else "count"
    goo = len(sys.argv) > 5 and "goo_stabilised" in sys.argv[5]
    emu = len(sys.argv) > 5 and sys.argv[5].startswith("emu")       # the CUDA source compiled for the host (tests/emu)
    block = block or (emu and "block" in sys.argv[5])
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    if block:   # dam-break block in the left half, mover sphere straddling a slab edge inside the water
        prob = make_problem(n_req, tank_w=15.0 * float(np.sqrt(n_req / 750.0)), water_frac=0.5, nranks=world)
        t = default_tunable(prob["h"], prob["tank_w"], prob["tank_h"], preset="y" if goo else "x")
        t.mover_center_x = 0.4 * prob["tank_w"]
    else:
        prob = make_problem(n_req, nranks=world)
        t = default_tunable(prob["h"], prob["tank_w"], prob["tank_h"], preset="y" if goo else "x")

    def backend(tw, th, h, cap, msg, r, w):
        if emu:
            sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
            from emu.backend import EmuSlab
            e = EmuSlab(tw, th, h, cap, msg, r, w)
            if goo:
                e.set_viscosity_stabilisation(0.5)
            return e
        g = GatherOracle(tw, th, h, cap, msg, r, w)
        if goo:
            g.set_viscosity_stabilisation(0.5)      # the proposal of DESIGN.md 5b (oracle only)
        return g

    sim = SlabRunner(prob, t, rank, world, backend=backend, balance=bool(balance), balance_policy=policy,
                     exchange_period=int(os.environ.get("SPH_EMU_XPERIOD", "1")),
                     msg_capacity=int(os.environ.get("SPH_EMU_MSG_CAPACITY", "0")) or None,
                     halo_width=float(os.environ.get("SPH_EMU_HALO_WIDTH", "0")) or None)
    sim.init_lattice()
    history = []
    elastic = len(sys.argv) > 5 and "elastic" in sys.argv[5]
    for s in range(steps):
        if elastic and s == 43:
            sim.remove_partition()          # last slab parked: it drains into its left neighbour
        if elastic and s == 123:
            history.append((-1, sim.ctx.status().n_local, 0, 0))
            sim.add_partition()             # and comes back with the right half of that neighbour's slab
        sim.step_once()
        if s % 20 == 19:
            st = sim.ctx.status()
            history.append((st.n_local, st.n_halo, sim.edges[rank][0], sim.edges[rank][1]))
    e2e = None
    if os.environ.get("SPH_EMU_E2E"):
        # bench.py's end-to-end leg on slabs (SlabRunner.e2e: blocks of frames that each start from the restored state,
        # synchronous and pipelined protocols), with the CUDA-only calls of torch stubbed: protocol and bookkeeping, not time
        import torch
        torch.cuda.synchronize = lambda *a: None
        torch.Tensor.pin_memory = lambda self: self
        sim.state_save()
        outs = []
        for _ in range(2):      # twice, as bench.py does it (restore, then the leg): both must end in the same state
            sim.state_restore()
            e2e = sim.e2e(4, torch.zeros(16, dtype=torch.uint8), block=2)
            outs.append(sim.ctx.download())
        same = all(np.array_equal(outs[0][0][f].view("u4"), outs[1][0][f].view("u4")) for f in ("x", "y", "v_x", "v_y")) \
            and np.array_equal(outs[0][1], outs[1][1])
        e2e = dict(e2e, repeatable=bool(same))
    a, uid = sim.ctx.download()
    st = sim.ctx.status()
    np.savez(f"{out}.rank{rank}.npz", state=a, uid=uid, history=np.array(history, "f8"),
             overflow=np.array([st.capacity_overflow, st.msg_overflow]), edges=np.array(sim.edges, "f8"),
             costs=np.array(sim.costs if sim.costs is not None else [], "f8"), exchanges=np.array([sim.exchanges]),
             n_exchanges=np.array([getattr(sim, "n_exchanges", 0)]),
             e2e=np.array([repr(e2e)]))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
