"""One rank of the multi-GPU slab test (NCCL): dam-break block, migration + ghosts + rebalancing.
Rank 0 also runs the same problem on one GPU alone for the bit-for-bit comparison."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    import sph_b200
    from sph_b200.slab import SlabRunner

    out, n_req, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    transport = sys.argv[4] if len(sys.argv) > 4 else "p2p"
    goo = len(sys.argv) > 5 and sys.argv[5] == "goo_stabilised"     # preset y with the stabilised viscosity gather
    # "onex:E:policy": one exchange per step, neighbours meet every E steps, edge policy count / cost / time
    onex = next((a for a in sys.argv[5:] if a.startswith("onex:")), None)
    period, policy = (int(onex.split(":")[1]), onex.split(":")[2]) if onex else (1, "count")
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    tank_w = 15.0 * float(np.sqrt(n_req / (1500.0 * 0.5)))
    prob = sph_b200.make_problem(n_req, tank_w=tank_w, water_frac=0.5, nranks=world)
    t = sph_b200.default_params(prob["h"], prob["tank_w"], prob["tank_h"], preset="y" if goo else "x")
    t.mover_center_x = 0.4 * prob["tank_w"]          # in the path of the collapsing block
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        # (a one-exchange build of the library needs one more h of ghost layer with the stabilised gather)
        halo = 4.5 if goo and sph_b200.lib().sph_exchanges_per_step() == 1 else None
        sim = SlabRunner(prob, t, rank, world, stream, transport=transport, halo_width=halo, exchange_period=period,
                         exchanges_per_step=1 if onex else 0, balance_policy=policy)
        if goo:
            sim.ctx.set_viscosity_stabilisation(0.5)
        sim.init_lattice()
        sim.run(steps)
        a, uid = sim.ctx.download()
        st = sim.ctx.status()
        print(f"rank {rank}: n_local={st.n_local} n_halo={st.n_halo} cap_over={st.capacity_overflow} msg_over={st.msg_overflow} "
              f"err={sim.ctx.L.sph_last_error(sim.ctx.h).decode()!r}", flush=True)
    np.savez(f"{out}.rank{rank}.npz", state=a, uid=uid, overflow=np.array([st.capacity_overflow, st.msg_overflow]),
             edges=np.array(sim.edges, "f8"))
    if rank == 0:
        p1 = sph_b200.make_problem(n_req, tank_w=tank_w, water_frac=0.5, nranks=1)
        ctx = sph_b200.Context(p1["tank_w"], p1["tank_h"], p1["h"], p1["n_global"] + 64)
        t1 = sph_b200.default_params(p1["h"], p1["tank_w"], p1["tank_h"], preset="y" if goo else "x"); t1.mover_center_x = 0.4 * p1["tank_w"]
        ctx.set_params(t1)
        if goo:
            ctx.set_viscosity_stabilisation(0.5)
        a1, u1 = sph_b200.lattice(p1)
        ctx.upload(a1, u1)
        ctx.step(steps)
        s1, su = ctx.download()
        np.savez(f"{out}.single.npz", state=s1, uid=su)
        print("edges", sim.edges, "counts", sim.counts)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
