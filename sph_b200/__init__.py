"""sph_b200 -- B200-native TinySPH compute-rank timestep.

Python is plumbing here: this module binds the C ABI in include/sph_b200.h
(libsph_b200.so, hand-written CUDA for sm_100a) with ctypes so that tests, bench.py and the
multi-rank driver (sph_b200.slab) can call it.  There is NO CPU fallback: importing works
without a GPU (so the symbol table can be checked), but creating a Context without the built
library or without a CUDA device raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsph_b200.so")

# == struct FLUID_PARTICLE (fluid.h:56-70), 52 bytes
PARTICLE = np.dtype([("x_prev", "f4"), ("y_prev", "f4"), ("x", "f4"), ("y", "f4"),
                     ("v_x", "f4"), ("v_y", "f4"), ("a_x", "f4"), ("a_y", "f4"),
                     ("density", "f4"), ("density_near", "f4"),
                     ("pressure", "f4"), ("pressure_near", "f4"), ("id", "i4")])

ORDER_UID, ORDER_CELL = 0, 1
HALO_BIT = 0x80000000
UID_MASK = 0x7FFFFFFF


class Tunable(C.Structure):
    """== struct TUNABLE_PARAMETERS (fluid.h:78-97), 64 bytes."""
    _fields_ = [(n, C.c_float) for n in (
        "rest_density", "smoothing_radius", "g", "k", "k_near", "k_spring", "sigma", "beta",
        "time_step", "node_start_x", "node_end_x", "mover_center_x", "mover_center_y",
        "mover_width", "mover_height")] + [("mover_type", C.c_char), ("kill_sim", C.c_char),
                                           ("active", C.c_char)]

    def copy(self):
        t = Tunable()
        C.memmove(C.byref(t), C.byref(self), C.sizeof(Tunable))
        return t


class Config(C.Structure):
    """== sph_config."""
    _fields_ = [("tank_w", C.c_float), ("tank_h", C.c_float), ("h", C.c_float),
                ("capacity", C.c_int), ("msg_capacity", C.c_int), ("device", C.c_int),
                ("rank", C.c_int), ("nranks", C.c_int), ("halo_width", C.c_float),
                ("stream", C.c_void_p), ("exchanges_per_step", C.c_int)]


class Status(C.Structure):
    """== sph_status."""
    _fields_ = [(n, C.c_int) for n in (
        "n_local", "n_halo", "max_bucket", "bucket_overflow", "neighbor_overflow",
        "capacity_overflow", "msg_overflow", "migrated_left", "migrated_right", "exchange_timeouts")] + [("steps", C.c_longlong)]


# every symbol include/sph_b200.h declares
C_ABI_SYMBOLS = (
    "sph_create", "sph_destroy", "sph_last_error", "sph_synchronize", "sph_get_status",
    "sph_set_params", "sph_queue_params", "sph_set_edges", "sph_upload", "sph_download",
    "sph_advect", "sph_sort", "sph_density", "sph_relax", "sph_step", "sph_exchange_buffers",
    "sph_set_neighbors", "sph_get_cells", "sph_get_pairs", "sph_get_forward_counts",
    "sph_pack_coords", "sph_launch_count", "sph_run_frame", "sph_p2p_local_handle", "sph_p2p_connect", "sph_copy_n_local", "sph_copy_load", "sph_init_lattice",
    "sph_set_viscosity_stabilisation", "sph_pack_coords_async", "sph_coords_wait", "sph_coords_copied", "sph_run_frame_async",
    "sph_exchanges_per_step", "sph_refresh_ghosts", "sph_exchange_via_host", "sph_set_exchange_period", "sph_exchange_due", "sph_get_exchange_times", "sph_state_save", "sph_state_restore", "sph_copy_work", "sph_ctx_exchanges_per_step",
)

_lib = None


class SphError(RuntimeError):
    pass


def lib():
    """Load libsph_b200.so; raises when it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SphError(f"{LIB_PATH} is missing: build it with `python -m sph_b200.build` "
                           "(nvcc, sm_100a). sph_b200 has no CPU path.")
        _lib = _bind(C.CDLL(LIB_PATH))
    return _lib


def _bind(L):
    """Declare the C-ABI signatures on a loaded library."""
    L.sph_last_error.restype = C.c_char_p
    L.sph_last_error.argtypes = [C.c_void_p]
    L.sph_launch_count.restype = C.c_longlong
    L.sph_launch_count.argtypes = [C.c_void_p]
    L.sph_get_pairs.restype = C.c_longlong
    L.sph_get_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong]
    L.sph_set_edges.argtypes = [C.c_void_p, C.c_float, C.c_float]
    L.sph_set_viscosity_stabilisation.argtypes = [C.c_void_p, C.c_float, C.c_float]
    L.sph_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_void_p)]
    for name in ("sph_destroy", "sph_synchronize", "sph_advect", "sph_sort", "sph_density", "sph_relax", "sph_refresh_ghosts",
                 "sph_state_save", "sph_state_restore"):
        getattr(L, name).argtypes = [C.c_void_p]
    L.sph_step.argtypes = [C.c_void_p, C.c_int]
    L.sph_set_params.argtypes = [C.c_void_p, C.POINTER(Tunable)]
    L.sph_queue_params.argtypes = [C.c_void_p, C.POINTER(Tunable)]
    L.sph_get_status.argtypes = [C.c_void_p, C.POINTER(Status)]
    L.sph_upload.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.sph_download.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    L.sph_set_neighbors.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.sph_get_cells.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.sph_get_forward_counts.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.sph_pack_coords.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.sph_run_frame.argtypes = [C.c_void_p, C.POINTER(Tunable), C.c_int, C.c_void_p, C.c_int]
    L.sph_run_frame_async.argtypes = [C.c_void_p, C.POINTER(Tunable), C.c_int, C.c_void_p, C.c_int]
    L.sph_pack_coords_async.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.sph_coords_wait.argtypes = [C.c_void_p, C.c_int]
    L.sph_coords_copied.argtypes = [C.c_void_p, C.c_int]
    L.sph_init_lattice.argtypes = [C.c_void_p] + [C.c_float] * 4 + [C.c_int] * 3
    L.sph_copy_n_local.argtypes = [C.c_void_p, C.c_void_p]
    L.sph_copy_load.argtypes = [C.c_void_p, C.c_void_p]
    L.sph_copy_work.argtypes = [C.c_void_p, C.c_void_p]
    L.sph_ctx_exchanges_per_step.argtypes = [C.c_void_p]
    L.sph_p2p_local_handle.argtypes = [C.c_void_p, C.c_void_p]
    L.sph_p2p_connect.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.sph_exchange_buffers.argtypes = [C.c_void_p, C.c_int] + [C.POINTER(C.c_void_p)] * 4 + [C.POINTER(C.c_size_t)]
    L.sph_set_exchange_period.argtypes = [C.c_void_p, C.c_int]
    L.sph_exchange_due.argtypes = [C.c_void_p]
    L.sph_get_exchange_times.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_int]
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Context:
    """One slab of the simulation resident on one GPU (sph_ctx)."""

    def __init__(self, tank_w, tank_h, h, capacity, msg_capacity=1, device=0, rank=0, nranks=1,
                 halo_width=0.0, stream=None, exchanges_per_step=0):
        """halo_width 0: the mode's default (2 h; 3.5 h with one exchange per step).  exchanges_per_step 0: the build's default."""
        self.L = lib()
        self.capacity = int(capacity)
        self.cfg = Config(tank_w, tank_h, h, int(capacity), int(msg_capacity), device, rank, nranks,
                          halo_width, stream, int(exchanges_per_step))
        self.h = C.c_void_p()
        rc = self.L.sph_create(C.byref(self.cfg), C.byref(self.h))
        if rc != 0:
            msg = self.L.sph_last_error(self.h).decode() if self.h else "sph_create failed"
            if self.h:
                self.L.sph_destroy(self.h)
                self.h = None
            raise SphError(f"sph_create -> {rc}: {msg}")

    def close(self):
        if getattr(self, "h", None):
            self.L.sph_destroy(self.h)
            self.h = None

    __del__ = close

    def _ck(self, rc, what):
        if rc != 0:
            raise SphError(f"{what} -> {rc}: {self.L.sph_last_error(self.h).decode()}")

    def set_params(self, t): self._ck(self.L.sph_set_params(self.h, C.byref(t)), "sph_set_params")
    def queue_params(self, t): self._ck(self.L.sph_queue_params(self.h, C.byref(t)), "sph_queue_params")
    def set_viscosity_stabilisation(self, gamma, min_dt_sigma=0.0):
        """Optional stabilised viscosity gather (sph_set_viscosity_stabilisation): needed for the goo preset."""
        self._ck(self.L.sph_set_viscosity_stabilisation(self.h, gamma, min_dt_sigma), "sph_set_viscosity_stabilisation")

    def set_edges(self, s, e): self._ck(self.L.sph_set_edges(self.h, s, e), "sph_set_edges")
    def set_neighbors(self, l, r): self._ck(self.L.sph_set_neighbors(self.h, int(l), int(r)), "sph_set_neighbors")
    def synchronize(self): self._ck(self.L.sph_synchronize(self.h), "sph_synchronize")

    def upload(self, aos, uid=None):
        aos = np.ascontiguousarray(aos, PARTICLE)
        u = None if uid is None else np.ascontiguousarray(uid, "u4")
        self._ck(self.L.sph_upload(self.h, _p(aos), None if u is None else _p(u), len(aos)), "sph_upload")

    def init_lattice(self, prob, rank=0):
        """Device-side lattice fill of one slab of make_problem()'s geometry. Returns the particle count."""
        sc, nc, _, _ = prob["slabs"][rank]
        n = self.L.sph_init_lattice(self.h, 0.0, 0.0, prob["tank_h"], prob["spacing"], sc, nc, prob["total_cols"])
        if n < 0:
            self._ck(-n, "sph_init_lattice")
        return n

    def download(self, order=ORDER_UID, include_halo=False):
        a = np.zeros(self.capacity, PARTICLE)
        u = np.zeros(self.capacity, "u4")
        n = self.L.sph_download(self.h, _p(a), _p(u), order, int(include_halo))
        if n < 0:
            self._ck(-n, "sph_download")
        return a[:n].copy(), u[:n].copy()

    def advect(self): self._ck(self.L.sph_advect(self.h), "sph_advect")
    def sort(self): self._ck(self.L.sph_sort(self.h), "sph_sort")
    def density(self): self._ck(self.L.sph_density(self.h), "sph_density")
    def relax(self): self._ck(self.L.sph_relax(self.h), "sph_relax")
    def step(self, n=1): self._ck(self.L.sph_step(self.h, int(n)), "sph_step")
    def refresh_ghosts(self): self._ck(self.L.sph_refresh_ghosts(self.h), "sph_refresh_ghosts")
    def state_save(self): self._ck(self.L.sph_state_save(self.h), "sph_state_save")
    def state_restore(self): self._ck(self.L.sph_state_restore(self.h), "sph_state_restore")

    def run_frame(self, tunable, steps, coords_out):
        """One render frame (fluid.c:270-372): `steps` sub-steps, the parameter scatter landing in
        the last one, then the int16 coordinate feed into `coords_out` (host array). Returns n_local."""
        n = self.L.sph_run_frame(self.h, C.byref(tunable) if tunable is not None else None, int(steps),
                                 _p(coords_out) if coords_out is not None else None,
                                 0 if coords_out is None else coords_out.size // 2)
        if n < 0:
            self._ck(-n, "sph_run_frame")
        return n

    def run_frame_async(self, tunable, steps, coords_out):
        """run_frame without the wait at its end: returns a ticket for coords_wait (sph_run_frame_async)."""
        k = self.L.sph_run_frame_async(self.h, C.byref(tunable) if tunable is not None else None, int(steps),
                                       _p(coords_out), coords_out.size // 2)
        if k < 0:
            self._ck(-k, "sph_run_frame_async")
        return k

    def pack_coords_async(self, coords_out):
        k = self.L.sph_pack_coords_async(self.h, _p(coords_out), coords_out.size // 2)
        if k < 0:
            self._ck(-k, "sph_pack_coords_async")
        return k

    def coords_wait(self, ticket):
        """Blocks until the frame of `ticket` is in its host buffer; returns its particle count."""
        n = self.L.sph_coords_wait(self.h, int(ticket))
        if n < 0:
            self._ck(-n, "sph_coords_wait")
        return n

    def coords_copied(self, ticket):
        """Entries of that ticket's frame that crossed to the host (after coords_wait)."""
        return int(self.L.sph_coords_copied(self.h, int(ticket)))

    def status(self):
        s = Status()
        self._ck(self.L.sph_get_status(self.h, C.byref(s)), "sph_get_status")
        return s

    @property
    def exchanges_per_step(self):
        return int(self.L.sph_ctx_exchanges_per_step(self.h))

    def set_exchange_period(self, period):
        """One-exchange build: neighbours meet every `period` steps (sph_set_exchange_period)."""
        self._ck(self.L.sph_set_exchange_period(self.h, int(period)), "sph_set_exchange_period")

    def exchange_times(self, reset=True):
        """-> ({send, wait, unpack} microseconds per meeting, meetings) since the last reset (peer-memory transport)."""
        us = (C.c_double * 3)(); n = C.c_int()
        self._ck(self.L.sph_get_exchange_times(self.h, us, C.byref(n), int(reset)), "sph_get_exchange_times")
        k = max(n.value, 1)
        return {"send_us": us[0] / k, "wait_us": us[1] / k, "unpack_us": us[2] / k}, n.value

    @property
    def exchange_due(self):
        return bool(self.L.sph_exchange_due(self.h))

    def exchange_via_host(self, which, sendrecv):
        """sph_exchange_via_host with a Python callable sendrecv(send_bytes_or_None, to_side, recv_nbytes, from_side)
        -> bytes received (or None): what a host with a plain (not CUDA-aware) MPI_Sendrecv would plug in."""
        FN = C.CFUNCTYPE(None, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p)

        def cb(send, nsend, to_side, recv, nrecv, from_side, user):
            got = sendrecv(C.string_at(send, nsend) if send else None, to_side, nrecv, from_side)
            if recv and got is not None:
                C.memmove(recv, got, min(len(got), nrecv))
        fn = FN(cb)
        self.L.sph_exchange_via_host.argtypes = [C.c_void_p, C.c_int, FN, C.c_void_p]
        self._ck(self.L.sph_exchange_via_host(self.h, int(which), fn, None), "sph_exchange_via_host")

    def exchange_pointers(self, which):
        """Device pointers (send_left, recv_left, send_right, recv_right) and message bytes."""
        ptr = [C.c_void_p() for _ in range(4)]
        nb = C.c_size_t()
        self._ck(self.L.sph_exchange_buffers(self.h, which, *[C.byref(p) for p in ptr], C.byref(nb)),
                 "sph_exchange_buffers")
        return [p.value for p in ptr], nb.value

    def copy_n_local(self, device_ptr):
        self._ck(self.L.sph_copy_n_local(self.h, device_ptr), "sph_copy_n_local")

    def copy_load(self, device_ptr):
        """{n_local, work estimate} as two ints into device memory, stream-ordered (sph_copy_load)."""
        self._ck(self.L.sph_copy_load(self.h, device_ptr), "sph_copy_load")

    def copy_work(self, device_ptr):
        """{n_local, work estimate, own time [us], waits [us]} as four ints into device memory (sph_copy_work)."""
        self._ck(self.L.sph_copy_work(self.h, device_ptr), "sph_copy_work")

    def p2p_handle(self):
        """64-byte cudaIpc handle of this rank's exchange block."""
        buf = C.create_string_buffer(64)
        self._ck(self.L.sph_p2p_local_handle(self.h, buf), "sph_p2p_local_handle")
        return bytes(buf.raw)

    def p2p_connect(self, left, right):
        l = C.create_string_buffer(left, 64) if left else None
        r = C.create_string_buffer(right, 64) if right else None
        self._ck(self.L.sph_p2p_connect(self.h, l, r), "sph_p2p_connect")

    def cells(self):
        u = np.zeros(self.capacity, "u4"); c = np.zeros(self.capacity, "u4")
        n = self.L.sph_get_cells(self.h, _p(u), _p(c), self.capacity)
        if n < 0:
            self._ck(-n, "sph_get_cells")
        return u[:n].copy(), c[:n].copy()

    def pairs(self):
        n = self.L.sph_get_pairs(self.h, None, 0)
        if n < 0:
            self._ck(int(-n), "sph_get_pairs")
        out = np.zeros(max(int(n), 1), "u8")
        n = self.L.sph_get_pairs(self.h, _p(out), len(out))
        if n < 0:
            self._ck(int(-n), "sph_get_pairs")
        return np.sort(out[:n])

    def forward_counts(self):
        u = np.zeros(self.capacity, "u4"); c = np.zeros(self.capacity, "i4")
        n = self.L.sph_get_forward_counts(self.h, _p(u), _p(c), self.capacity)
        if n < 0:
            self._ck(-n, "sph_get_forward_counts")
        return u[:n].copy(), c[:n].copy()

    def pack_coords(self):
        xy = np.zeros(2 * self.capacity, "i2")
        n = self.L.sph_pack_coords(self.h, _p(xy), self.capacity)
        if n < 0:
            self._ck(-n, "sph_pack_coords")
        return xy[:2 * n].reshape(n, 2).copy()

    @property
    def launches(self):
        return int(self.L.sph_launch_count(self.h))


# ------------------------------------------------------------------------------------------------
# C host layer (include/sph_host.h): start-up geometry, parameter model, slab load balancer
# ------------------------------------------------------------------------------------------------
HOST_SYMBOLS = ("sph_host_mover_autopilot", "sph_host_mover_autopilot_ex", "sph_host_spacing", "sph_host_default_params", "sph_host_preset", "sph_host_partition",
                "sph_host_lattice", "sph_host_balance", "sph_host_balance_ex", "sph_host_balance_time", "sph_host_remove_partition", "sph_host_add_partition")


def _host():
    L = lib()
    if not getattr(L, "_host_ready", False):
        L.sph_host_spacing.restype = C.c_float
        L.sph_host_spacing.argtypes = [C.c_float, C.c_float, C.c_int]
        L.sph_host_default_params.argtypes = [C.POINTER(Tunable), C.c_float, C.c_float, C.c_float]
        L.sph_host_preset.argtypes = [C.POINTER(Tunable), C.c_char]
        L.sph_host_partition.argtypes = [C.c_float] * 6 + [C.c_int] + [C.c_void_p] * 4
        L.sph_host_lattice.argtypes = [C.c_float] * 4 + [C.c_int] * 3 + [C.c_void_p] * 2
        L.sph_host_mover_autopilot.argtypes = [C.POINTER(Tunable), C.c_float, C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_int)]
        L.sph_host_mover_autopilot_ex.argtypes = [C.POINTER(Tunable), C.c_float, C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_float]
        L.sph_host_balance.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.sph_host_remove_partition.argtypes = [C.c_void_p, C.c_int]
        L.sph_host_add_partition.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L._host_ready = True
    return L


def default_params(h, tank_w, tank_h, preset="x"):
    t = Tunable()
    L = _host()
    L.sph_host_default_params(C.byref(t), h, tank_w, tank_h)
    if L.sph_host_preset(C.byref(t), preset.encode()) != 0:
        raise ValueError(f"unknown preset {preset!r}")
    return t


def make_problem(n_request, tank_w=15.0, aspect=16.0 / 9.0, water_frac=1.0, nranks=1):
    """Geometry of start_simulation (fluid.c:116-159) for a tank scaled to n_request particles."""
    L = _host()
    tank_w = float(np.float32(tank_w))
    tank_h = float(np.float32(np.float32(tank_w) / np.float32(aspect)))
    water_w = float(np.float32(np.float32(tank_w) * np.float32(water_frac)))
    spacing = float(L.sph_host_spacing(water_w, tank_h, int(n_request)))
    sc = np.zeros(nranks, "i4"); nc = np.zeros(nranks, "i4")
    sx = np.zeros(nranks, "f4"); ex = np.zeros(nranks, "f4")
    n_global = L.sph_host_partition(tank_w, 0.0, water_w, 0.0, tank_h, spacing, nranks, _p(sc), _p(nc), _p(sx), _p(ex))
    return dict(tank_w=tank_w, tank_h=tank_h, water_w=water_w, spacing=spacing,
                h=float(np.float32(2.0) * np.float32(spacing)), n_global=int(n_global), total_cols=int(nc.sum()),
                slabs=[(int(sc[r]), int(nc[r]), float(sx[r]), float(ex[r])) for r in range(nranks)])


def lattice(prob, rank=0):
    """Initial particles of one slab (geometry.c:29-59) with persistent uids."""
    L = _host()
    sc, nc, _, _ = prob["slabs"][rank]
    rows = int(np.floor(np.float32(prob["tank_h"]) / np.float32(prob["spacing"])))
    a = np.zeros(nc * rows, PARTICLE); uid = np.zeros(nc * rows, "u4")
    n = L.sph_host_lattice(0.0, 0.0, prob["tank_h"], prob["spacing"], sc, nc, prob["total_cols"], _p(a), _p(uid))
    assert n == len(a)
    return a, uid


def _edge_blocks(edges, h):
    m = (Tunable * len(edges))()
    for r, (s, e) in enumerate(edges):
        m[r].smoothing_radius = h; m[r].node_start_x = s; m[r].node_end_x = e
    return m


def balance(edges, counts, h, nactive=None, band_divisor=15.0):
    """check_partition_left (renderer.c:427-477) on a list of (start_x, end_x); returns the new list.
    Only the first `nactive` slabs take part (render_state->num_compute_procs_active).  `band_divisor`:
    dead band = even / band_divisor (the reference's 15 unless the caller balances on a work estimate)."""
    L = _host()
    n = len(edges)
    nactive = n if nactive is None else nactive
    m = _edge_blocks(edges, h)
    c = np.asarray(counts, "i4")
    L.sph_host_balance_ex.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_float]
    L.sph_host_balance_ex(m, nactive, _p(c), int(c.sum()), float(band_divisor))
    return [(float(m[r].node_start_x), float(m[r].node_end_x)) for r in range(n)]


def balance_time(edges, busy, h, nactive=None, gain=0.5, max_shift_h=1.0, min_width_h=2.0):
    """sph_host_balance_time: edges moved in proportion to the measured imbalance of adjacent slabs (not the reference's policy)."""
    L = _host()
    n = len(edges)
    nactive = n if nactive is None else nactive
    m = _edge_blocks(edges, h)
    c = np.asarray(busy, "i4")
    L.sph_host_balance_time.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_float, C.c_float]
    L.sph_host_balance_time(m, nactive, _p(c), float(gain), float(max_shift_h), float(min_width_h))
    return [(float(m[r].node_start_x), float(m[r].node_end_x)) for r in range(n)]


def remove_partition(edges, h, nactive):
    """remove_partition (controls.c:405-426): returns (new edges, new nactive)."""
    m = _edge_blocks(edges, h)
    na = _host().sph_host_remove_partition(m, nactive)
    return [(float(m[r].node_start_x), float(m[r].node_end_x)) for r in range(len(edges))], na


def add_partition(edges, h, nactive):
    """add_partition (controls.c:429-455): returns (new edges, new nactive)."""
    m = _edge_blocks(edges, h)
    na = _host().sph_host_add_partition(m, nactive, len(edges))
    return [(float(m[r].node_start_x), float(m[r].node_end_x)) for r in range(len(edges))], na
