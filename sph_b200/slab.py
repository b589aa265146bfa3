"""Multi-rank driver: one process per GPU, one x-slab per process (communication.c's design).

torch.distributed is plumbing: it bootstraps the ranks, exchanges the cudaIpc handles of the exchange blocks
and all-gathers one int per rank per frame.  The neighbour messages themselves -- packed inside the
advect/relax kernels, unpacked inside the sort (include/sph_b200.h, "slab exchange") -- move over NVLink by
peer-memory stores with device-side arrival flags (transport="p2p", one CUDA graph per step), or, as an
alternative and in the CPU tests (gloo), through torch.distributed send/recv (transport="collective").
Per step:

    advect -> exchange 0 (migrants + predicted-position ghosts) -> sort -> density -> relax
           -> exchange 1 (relaxed position + velocity ghosts)   -> sort

which replaces transferOOBParticles + 2 x start/finishHaloExchange + 2 x hash_halo
(fluid.c:310-348).  Once per frame (every `steps_per_frame` steps, at the sub-step where the
reference's MPI_Scatterv lands, fluid.c:293-294) all ranks share their particle counts and run the
reference's edge balancer (renderer.c:427-477) on identical inputs, so every rank derives the same
new edges without a coordinator.

balance_policy="cost" (optional, not the reference's): the same edge arithmetic fed with a work estimate per
slab (sum over resident entries of 14 + neighbours, formed by the density kernel, sph_copy_load) and a
tighter dead band.  In a dam-break the cost per particle follows the local density, which differs from slab
to slab while the reference equalises particle COUNTS; since the gather gives bit-identical results for any
decomposition, only the schedule changes (tests/test_slab_gloo.py).
"""
import time

import numpy as np


class _CudaBytes:
    """Zero-copy view of a device pointer for torch.as_tensor."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def keep_slabs_wider_than(old, new, min_width, nactive):
    """One-exchange build: an interior slab narrower than the ghost layer would leave its neighbours' layers
    incomplete, and the reference's balancer (renderer.c:427-477) only guarantees 2 h.  Undo every edge move that
    would shrink an interior slab below `min_width`.  Pure arithmetic on identical inputs: every rank gets the same
    answer."""
    out = [list(e) for e in new]
    # (undoing one edge's move changes the width its neighbours' moves leave of the slab between them: repeated until
    #  nothing is undone any more -- one pass let a slab through that both of its edges had been moving along with)
    for _ in range(nactive):
        undone = False
        for r in range(nactive - 1):                   # the edge between slab r and slab r+1
            if out[r][1] == old[r][1]:
                continue
            shrunk = r if out[r][1] < old[r][1] else r + 1
            interior = 0 < shrunk < nactive - 1
            # (what the move leaves of the slab's OLD extent counts as well: in the step in which the new edges land the
            #  strip its other neighbour needs as ghosts must already be this slab's -- a slab whose two edges move the
            #  same way keeps its width while its old and new extents drift apart; sph_host_balance_time has the same rule)
            kept = out[r][1] - old[r][0] if shrunk == r else old[r + 1][1] - out[r][1]
            if interior and (out[shrunk][1] - out[shrunk][0] < min_width or kept < min_width):
                out[r][1] = old[r][1]
                out[r + 1][0] = old[r + 1][0]
                undone = True
        if not undone:
            break
    return [tuple(e) for e in out]


class SlabRunner:
    def __init__(self, prob, tunable, rank, world, stream=None, capacity_factor=2.0, backend=None,
                 msg_capacity=None, steps_per_frame=4, balance=True, group=None, transport="p2p",
                 async_counts=True, balance_policy="count", cost_band_divisor=40.0, halo_width=None, exchange_period=1,
                 exchanges_per_step=0, time_proportional=True):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.prob, self.rank, self.world, self.group = prob, rank, world, group
        self.steps_per_frame, self.do_balance = steps_per_frame, balance
        assert balance_policy in ("count", "cost", "time")
        self.balance_policy, self.cost_band_divisor = balance_policy, cost_band_divisor
        self.time_band_divisor = 100.0      # "time" policy: a slab's measured time within 1 % of the mean is left alone
        self.time_proportional = time_proportional
        self.costs = None
        self.sub_step = 0
        self.n_active = world            # slabs taking part (render_state->num_compute_procs_active)
        self.async_counts = async_counts
        self.stream = stream
        self.t = tunable.copy()
        self.edges = [(s, e) for (_, _, s, e) in prob["slabs"]]
        self.t.node_start_x, self.t.node_end_x = self.edges[rank]
        n_slab = max(nc for (_, nc, _, _) in prob["slabs"]) * int(np.floor(np.float32(prob["tank_h"]) / np.float32(prob["spacing"])))
        rows = int(np.floor(np.float32(prob["tank_h"]) / np.float32(prob["spacing"])))
        self.exchange_period = int(exchange_period)
        if self.exchange_period > 1 and not halo_width:
            # 3.5 h of ghost layer per step between exchanges, 4.5 h while the stabilised viscosity gather is engaged
            halo_width = (4.5 if tunable.time_step * tunable.sigma >= 0.5 else 3.5) * self.exchange_period
        # ghost layer 2h wide = ~4 lattice columns at rest (2 per h); settle-time compression and migrants: x6
        self.msg_capacity = int(msg_capacity or max(4096, 6 * (int(2 * (halo_width or 2.0)) + 1) * rows))
        self.capacity = int(capacity_factor * n_slab) + 4 * self.msg_capacity
        if backend is None:
            import sph_b200
            self.sph = sph_b200
            # halo_width None: the build's default (2 h; 3.5 h for the one-exchange build, which needs 4.5 h with
            # the stabilised viscosity gather)
            self.ctx = sph_b200.Context(prob["tank_w"], prob["tank_h"], prob["h"], self.capacity,
                                        msg_capacity=self.msg_capacity, device=torch.cuda.current_device(),
                                        rank=rank, nranks=world, halo_width=halo_width or 0.0,
                                        stream=stream.cuda_stream if stream is not None else None,
                                        exchanges_per_step=1 if self.exchange_period > 1 else exchanges_per_step)
            self.cuda = True
        else:
            self.ctx = backend(prob["tank_w"], prob["tank_h"], prob["h"], self.capacity, self.msg_capacity, rank, world)
            self.cuda = False
        self.ctx.set_params(self.t)
        self.exchanges = int(getattr(self.ctx, "exchanges_per_step", 2))     # 1: one-exchange build of the library
        if self.exchange_period > 1:
            # neighbours meet every `exchange_period` steps; needs the one-exchange build and a layer of 3.5 h per step
            self.ctx.set_exchange_period(self.exchange_period)
        self.has_left, self.has_right = rank > 0, rank < world - 1
        # "p2p": neighbours map each other's exchange block (cudaIpc) and the kernels store messages
        # straight into it over NVLink; "collective": torch.distributed send/recv moves the buffers
        self.transport = transport if (self.cuda and world > 1) else "collective"
        if self.transport == "p2p":
            handles = [None] * world
            dist.all_gather_object(handles, self.ctx.p2p_handle(), group=group)
            self.ctx.p2p_connect(handles[rank - 1] if self.has_left else None,
                                 handles[rank + 1] if self.has_right else None)
        self._bufs = {}
        self.counts = None
        self.exchange_s = 0.0

    # -------------------------------------------------------------------------------- plumbing
    def _buffers(self, which):
        if which not in self._bufs:
            torch = self.torch
            if self.cuda:
                ptrs, nb = self.ctx.exchange_pointers(which)
                self._bufs[which] = [torch.as_tensor(_CudaBytes(p, nb), device="cuda") for p in ptrs]
            else:
                self._bufs[which] = [torch.from_numpy(a) for a in self.ctx.exchange_buffers(which)]
        return self._bufs[which]

    def exchange(self, which):
        dist = self.dist
        self.n_exchanges = getattr(self, "n_exchanges", 0) + 1
        send_l, recv_l, send_r, recv_r = self._buffers(which)
        ops = []
        if self.has_right:
            ops += [dist.P2POp(dist.isend, send_r, self.rank + 1, self.group),
                    dist.P2POp(dist.irecv, recv_r, self.rank + 1, self.group)]
        if self.has_left:
            ops += [dist.P2POp(dist.isend, send_l, self.rank - 1, self.group),
                    dist.P2POp(dist.irecv, recv_l, self.rank - 1, self.group)]
        if ops:
            for r in dist.batch_isend_irecv(ops):
                r.wait()

    def _load_now(self):
        """(local count, work estimate) of this slab, read synchronously (tests, non-CUDA backends)."""
        st = self.ctx.status()
        if self.cuda:
            buf = self.torch.zeros(2, dtype=self.torch.int32, device="cuda")
            self.ctx.copy_load(buf.data_ptr())
            return [int(v) for v in buf.tolist()]
        from .csrc_constants import COST_BASE
        return [st.n_local, COST_BASE * (st.n_local + st.n_halo) + 2 * len(self.ctx.pairs())]

    def gather_counts(self):
        """-> (counts, costs) of all slabs"""
        torch, dist = self.torch, self.dist
        dev = "cuda" if self.cuda else "cpu"
        mine = torch.tensor(self._load_now(), dtype=torch.int32, device=dev)
        out = [torch.zeros(2, dtype=torch.int32, device=dev) for _ in range(self.world)]
        dist.all_gather(out, mine, group=self.group)
        vals = [[int(v) for v in x.tolist()] for x in out]
        return [v[0] for v in vals], [v[1] for v in vals]

    def sample_counts_async(self):
        """End of a frame: all-gather the slab populations (and work figures) WITHOUT stalling the host or the compute
        stream -- the collective runs on a side stream behind an event, so it is no barrier between the slabs' steps
        (in round 1 it sat in the compute stream: every frame all ranks met there a third time).  The result is used
        one frame later, which is the reference's own timing: its render rank balances on the counts of the
        coordinate messages of the previous frame (renderer.c:268-290)."""
        torch, dist = self.torch, self.dist
        if not hasattr(self, "_cnt_mine"):
            self._cnt_mine = [torch.zeros(4, dtype=torch.int32, device="cuda") for _ in range(2)]
            self._cnt_all = [torch.zeros(4 * self.world, dtype=torch.int32, device="cuda") for _ in range(2)]
            self._cnt_host = torch.zeros(4 * self.world, dtype=torch.int32).pin_memory()
            self._cnt_event = torch.cuda.Event()
            self._cnt_ready = torch.cuda.Event()
            self._side = torch.cuda.Stream()
            self._cnt_flip = 0
        k = self._cnt_flip = 1 - self._cnt_flip
        cur = self.stream if self.stream is not None else torch.cuda.current_stream()
        self.ctx.copy_work(self._cnt_mine[k].data_ptr())
        self._cnt_ready.record(cur)
        with torch.cuda.stream(self._side):
            self._side.wait_event(self._cnt_ready)
            dist.all_gather_into_tensor(self._cnt_all[k], self._cnt_mine[k], group=self.group)
            self._cnt_host.copy_(self._cnt_all[k], non_blocking=True)
            self._cnt_event.record(self._side)
        self._cnt_pending = True

    def rebalance(self):
        """check_partition_left on identical inputs on every rank; the new edges are queued so that they
        land between prediction and migration of the coming step (fluid.c:293-310)."""
        import sph_b200
        if self.cuda and self.async_counts:
            if not getattr(self, "_cnt_pending", False):
                return                                  # first frame: nothing sampled yet
            self._cnt_event.synchronize()               # recorded a frame ago: already complete
            both = [int(c) for c in self._cnt_host.tolist()]
            counts, costs, busy, waits = both[0::4], both[1::4], both[2::4], both[3::4]
        else:
            counts, costs = self.gather_counts()
            busy = waits = [0] * self.world
        self.counts, self.costs, self.busy_us, self.wait_us = counts, costs, busy, waits
        old_edges = list(self.edges)
        active = busy[:self.n_active]
        if self.balance_policy == "time" and min(active) > 0:
            # same edge arithmetic, fed with each slab's MEASURED device time between its meetings (sph_copy_work):
            # whatever makes a slab slow -- denser fluid, more ghosts, a mover, a slower GPU -- it gives up columns
            if self.time_proportional:
                # edges move in proportion to the measured imbalance of the two slabs they separate (converges in tens
                # of frames; the reference's fixed h/8 per frame needs hundreds for a 10 % imbalance)
                layer = (self.ctx.cfg.halo_width or 2.0) if self.exchanges == 1 else 2.0
                self.edges = sph_b200.balance_time(self.edges, busy, self.prob["h"], self.n_active, gain=0.5, max_shift_h=2.0,
                                                   min_width_h=max(2.0, layer))
            else:
                self.edges = sph_b200.balance(self.edges, busy, self.prob["h"], self.n_active, band_divisor=self.time_band_divisor)
        elif self.balance_policy == "cost":
            # same edge arithmetic, fed with the work estimate (scaled to stay far from int overflow)
            self.edges = sph_b200.balance(self.edges, [c >> 4 for c in costs], self.prob["h"], self.n_active,
                                          band_divisor=self.cost_band_divisor)
        else:
            # the reference feeds coordinate counts (2 per particle) on both sides of the ratio (renderer.c:280,290)
            self.edges = sph_b200.balance(self.edges, [2 * c for c in counts], self.prob["h"], self.n_active)
        if self.exchanges == 1:
            self.edges = keep_slabs_wider_than(old_edges, self.edges, (self.ctx.cfg.halo_width or 3.5) * self.prob["h"], self.n_active)
        self._queue_edges()

    def _queue_edges(self):
        self.t.node_start_x, self.t.node_end_x = self.edges[self.rank]
        self.t.active = bytes([1 if self.rank < self.n_active else 0])
        self.ctx.queue_params(self.t)

    def remove_partition(self):
        """Park the last active slab outside the tank (controls.c:405-426); its particles drain into the left
        neighbour through the ordinary migration path, at most one message capacity per step.  Call on every
        rank at the same step."""
        import sph_b200
        self.edges, self.n_active = sph_b200.remove_partition(self.edges, self.prob["h"], self.n_active)
        self._queue_edges()

    def add_partition(self):
        """Split the last active slab in half and hand the right half to the next parked rank
        (controls.c:429-455).  Call on every rank at the same step."""
        import sph_b200
        edges, n_active = sph_b200.add_partition(self.edges, self.prob["h"], self.n_active)
        if self.exchanges == 1:
            # one exchange per step: an interior slab narrower than the ghost layer makes the library refuse to step
            # (SPH_ERR_STATE) -- on that one rank, while its neighbours wait for it.  Same arithmetic on every rank:
            # refuse here, everywhere at once.
            layer = (self.ctx.cfg.halo_width or 3.5) * self.prob["h"]
            narrow = [r for r in range(1, n_active - 1) if edges[r][1] - edges[r][0] < layer]
            if narrow:
                raise ValueError(f"add_partition: slab {narrow[0]} would be narrower than the ghost layer of this exchange mode "
                                 f"({(edges[narrow[0]][1] - edges[narrow[0]][0]) / self.prob['h']:.1f} h < {layer / self.prob['h']:.1f} h)")
        self.edges, self.n_active = edges, n_active
        self._queue_edges()

    # -------------------------------------------------------------------------------- simulation
    def init_lattice(self):
        if self.cuda:
            self.ctx.init_lattice(self.prob, self.rank)        # filled on the device: no host AoS
        else:
            import sph_b200
            a, uid = sph_b200.lattice(self.prob, self.rank)
            self.ctx.upload(a, uid)
        if self.world > 1:
            if self.cuda:
                self.torch.cuda.synchronize()
            self.dist.barrier(group=self.group)       # message sequence numbers restart together

    def refresh_ghosts(self):
        """After an upload of MOVING particles on several slabs (a restart): give every slab its ghost layer, so that
        the viscosity pass of the first step sees the neighbours across the edges (sph_refresh_ghosts).  Call on every
        rank together."""
        self.ctx.refresh_ghosts()
        if self.transport != "p2p":
            self.exchange(1)
        self.ctx.sort()

    def step_once(self):
        if self.do_balance and self.world > 1 and self.sub_step == self.steps_per_frame - 1:
            self.rebalance()
        c = self.ctx
        if self.transport == "p2p":
            c.step(1)                                  # one CUDA graph; the exchange happens inside the kernels
            if self.do_balance and self.async_counts and self.sub_step == self.steps_per_frame - 1:
                self.sample_counts_async()
            self.sub_step = (self.sub_step + 1) % self.steps_per_frame
            return
        c.advect()
        if getattr(c, "exchange_due", True):           # exchange period > 1: nothing moves between exchange steps
            self.exchange(0)
        c.sort()
        c.density()
        c.relax()
        if self.exchanges == 2:
            self.exchange(1)
        c.sort()
        if self.do_balance and self.cuda and self.async_counts and self.world > 1 and self.sub_step == self.steps_per_frame - 1:
            self.sample_counts_async()
        self.sub_step = (self.sub_step + 1) % self.steps_per_frame

    def run(self, n):
        for _ in range(n):
            self.step_once()

    def state_save(self):
        """Snapshot in device memory (sph_state_save) + this driver's own bookkeeping; every rank together."""
        self.ctx.state_save()
        self._saved = (list(self.edges), self.sub_step, self.n_active, self.t.copy(), getattr(self, "_cnt_pending", False))

    def state_restore(self):
        self.ctx.state_restore()
        edges, self.sub_step, self.n_active, t, pending = self._saved
        self.edges, self.t = list(edges), t.copy()
        self._cnt_pending = False        # counts sampled after the snapshot belong to a future that is being undone

    # -------------------------------------------------------------------------------- bench hooks
    @property
    def launches(self):
        return self.ctx.launches

    def stage_times(self, nsteps, flush_buf):
        torch = self.torch
        names = ("advect", "exchange0", "sort1", "density", "relax", "exchange1", "sort2")
        acc = {k: 0.0 for k in names}
        c = self.ctx
        x = (lambda w: None) if self.transport == "p2p" else (lambda w: self.exchange(w) if w < self.exchanges else None)
        for _ in range(nsteps):
            flush_buf.zero_()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(8)]
            ev[0].record(self.stream); c.advect()
            ev[1].record(self.stream); x(0)
            ev[2].record(self.stream); c.sort()
            ev[3].record(self.stream); c.density()
            ev[4].record(self.stream); c.relax()
            ev[5].record(self.stream); x(1)
            ev[6].record(self.stream); c.sort()
            ev[7].record(self.stream)
            torch.cuda.synchronize()
            for i, k in enumerate(names):
                acc[k] += ev[i].elapsed_time(ev[i + 1])
        out = {k: v / nsteps for k, v in acc.items()}
        self._exchange_ms = out.pop("exchange0") + out.pop("exchange1")
        out["exchange"] = self._exchange_ms
        return {k: v for k, v in out.items()}

    def kernel_name(self, stage):
        return {"advect": "k_advect", "density": "k_density", "relax": "k_relax",
                "exchange": "peer stores inside k_advect/k_relax" if self.transport == "p2p" else "nccl send/recv",
                "sort1": "k_unpack+k_scan_apply+k_scatter+k_reorder",
                "sort2": "k_scan_apply+k_scatter+k_reorder"}[stage]

    def e2e(self, frames, flush_buf, block=None):
        """(In blocks of `block` frames that each start from the state of state_save(), restored outside the clock: the
        frames then are the steps the bench's timed blocks time.)
        Frames with host buffers, pipelined like the reference's MPI_Isend of its frame (fluid.c:283-287, :354-365) and
        like the single-GPU bench: per frame a parameter block H2D (queued by the rebalancer, lands at the last sub-step),
        4 steps, the slab's int16 coordinates D2H into pinned memory; frame f is collected after frame f+1 has been
        submitted.  One host clock per rank around all frames (the bench takes the max over ranks).  The synchronous
        protocol (sph_pack_coords after every frame, L2 flushed before it) is timed too and reported beside it."""
        torch, np_ = self.torch, np
        bufs = [torch.empty(2 * self.capacity, dtype=torch.int16).pin_memory().numpy() for _ in range(2)]
        c = self.ctx
        # synchronous protocol (a few frames are enough for the comparison figure)
        block = max(1, min(block or frames, frames))
        self.state_save()                              # the state every block starts from: the one e2e() was called in
        restore = True
        sync_frames = max(3, min(frames, 12))
        secs_sync, n = 0.0, 0
        for f in range(sync_frames + 1):
            flush_buf.zero_()
            torch.cuda.synchronize()
            self.dist.barrier()
            t0 = time.perf_counter()
            self.run(self.steps_per_frame)
            n = c.L.sph_pack_coords(c.h, bufs[0].ctypes.data, self.capacity)
            if f > 0:
                secs_sync += time.perf_counter() - t0
        out = {"seconds": secs_sync * frames / sync_frames, "steps": self.steps_per_frame * frames,
               "h2d_per_step": 64 / self.steps_per_frame, "d2h_per_step": 4 * n / self.steps_per_frame, "pipelined": False}
        try:
            for f in range(2):
                self.run(self.steps_per_frame)
                c.coords_wait(c.pack_coords_async(bufs[f]))
            secs, done, copied = 0.0, 0, 0
            while done < frames:
                if restore:
                    self.state_restore()
                torch.cuda.synchronize()
                self.dist.barrier()
                tickets = []
                t0 = time.perf_counter()
                for f in range(block):
                    self.run(self.steps_per_frame)
                    tickets.append(c.pack_coords_async(bufs[f % 2]))
                    if f > 0:
                        n = c.coords_wait(tickets[f - 1])
                        copied += c.coords_copied(tickets[f - 1])
                n = c.coords_wait(tickets[-1])
                copied += c.coords_copied(tickets[-1])
                secs += time.perf_counter() - t0
                done += block
            secs *= frames / done                      # (per `frames` frames, like the synchronous figure)
            # the last frame must be what the synchronous call packs from the same state (a slab packs its particles in
            # the order an atomic cursor hands out: compare as sets of coordinate pairs)
            last = bufs[(block - 1) % 2][:2 * n].reshape(n, 2).copy()
            ref = c.pack_coords()
            key = lambda a: np_.sort(a[:, 0].astype("i4") * 65536 + a[:, 1].astype("i4"))
            if len(ref) == n and np_.array_equal(key(last), key(ref)):
                # a slab does not know its population on the host without a stall: the population of the last collected
                # frame plus an eighth travels (sph_pack_coords_async); counted from what the library copied
                out = {"seconds": secs, "steps": self.steps_per_frame * frames, "h2d_per_step": 64 / self.steps_per_frame,
                       "d2h_per_step": 4 * copied / (self.steps_per_frame * done), "pipelined": True,
                       "sync_seconds": secs_sync * frames / sync_frames}
            else:
                out["pipelined_error"] = "last frame differs from the synchronous feed"
        except Exception as e:      # the synchronous number stands
            out["pipelined_error"] = repr(e)[:200]
        return out

    def stats(self):
        s = self.ctx.status()
        npairs = self.ctx.L.sph_get_pairs(self.ctx.h, None, 0) if self.cuda else 0
        return {"n_local": s.n_local, "n_halo": s.n_halo, "max_bucket": s.max_bucket,
                "mean_neighbours": 2.0 * npairs / max(s.n_local + s.n_halo, 1),
                "capacity_overflow": s.capacity_overflow, "msg_overflow": s.msg_overflow,
                "exchange_timeouts": getattr(s, "exchange_timeouts", 0)}
