"""ctypes front end of the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (sph_b200) never does.

Three things live here:
  * build_oracle(): gcc -O2 -ffp-contract=off oracle/sph_oracle.c -> oracle/_build/liborc.so
  * SeqOracle / GatherOracle: the two restatements in sph_oracle.c
  * Ref: the UNMODIFIED reference compiled into oracle/_ref/libsph_ref.so by
    oracle/ref_build/Makefile (exists only where /root/reference was present at build time,
    or where the prebuilt .so travelled with the repo snapshot).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/src"

# == struct FLUID_PARTICLE (fluid.h:56-70), 52 bytes
PARTICLE = np.dtype([("x_prev", "f4"), ("y_prev", "f4"), ("x", "f4"), ("y", "f4"),
                     ("v_x", "f4"), ("v_y", "f4"), ("a_x", "f4"), ("a_y", "f4"),
                     ("density", "f4"), ("density_near", "f4"),
                     ("pressure", "f4"), ("pressure_near", "f4"), ("id", "i4")])
assert PARTICLE.itemsize == 52


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


assert C.sizeof(Tunable) == 64


class Param(C.Structure):
    """== struct PARAM (fluid.h:100-106), 80 bytes."""
    _fields_ = [("tunable_params", Tunable), ("number_fluid_particles_global", C.c_int),
                ("number_fluid_particles_local", C.c_int), ("max_fluid_particle_index", C.c_int),
                ("number_halo_particles", C.c_int)]


assert C.sizeof(Param) == 80


class Config(C.Structure):
    """== sph_config (include/sph_b200.h)."""
    _fields_ = [("tank_w", C.c_float), ("tank_h", C.c_float), ("h", C.c_float),
                ("capacity", C.c_int), ("msg_capacity", C.c_int), ("device", C.c_int),
                ("rank", C.c_int), ("nranks", C.c_int), ("halo_width", C.c_float),
                ("stream", C.c_void_p), ("exchanges_per_step", C.c_int)]


class Status(C.Structure):
    """== sph_status (include/sph_b200.h)."""
    _fields_ = [(n, C.c_int) for n in (
        "n_local", "n_halo", "max_bucket", "bucket_overflow", "neighbor_overflow",
        "capacity_overflow", "msg_overflow", "migrated_left", "migrated_right", "exchange_timeouts")] + [("steps", C.c_longlong)]


# fluid presets: fluid.c:90-97 / controls.c:344-401
PRESETS = {
    "x": dict(g=6.0, k=0.2, k_near=6.0, k_spring=10.0, sigma=5.0, beta=0.5, rest_density=30.0),
    "y": dict(g=6.0, k=0.1, k_near=3.0, k_spring=-30.0, sigma=100.0, beta=10.0, rest_density=30.0),
    "a": dict(g=0.0, k=0.2, k_near=6.0, k_spring=10.0, sigma=20.0, beta=2.0, rest_density=55.0),
    "b": dict(g=6.0, k=0.0, k_near=0.0, k_spring=115.0, sigma=20.0, beta=2.0, rest_density=0.0),
}


def default_tunable(h, tank_w, tank_h, preset="x", steps_per_frame=4):
    """Parameter block as start_simulation builds it (fluid.c:88-107,159), mover parked as in
    SURVEY.md 8(d): sphere, diameter 2/15 of the tank width, at (0.5 W, 0.35 H)."""
    t = Tunable()
    for k, v in PRESETS[preset].items():
        setattr(t, k, v)
    t.smoothing_radius = h
    t.time_step = float(np.float32(np.float32(1.0) / np.float32(30.0)) / np.float32(steps_per_frame))
    t.node_start_x = 0.0
    t.node_end_x = tank_w
    t.mover_center_x = 0.5 * tank_w
    t.mover_center_y = 0.35 * tank_h
    t.mover_width = 2.0 * tank_w / 15.0
    t.mover_height = 2.0 * tank_w / 15.0
    t.mover_type = bytes([0])
    t.kill_sim = bytes([0])
    t.active = bytes([1])
    return t


# --------------------------------------------------------------------------- build

def build_oracle(force=False):
    """Compile oracle/sph_oracle.c -> oracle/_build/liborc.so (flags: see sph_oracle.h)."""
    out_dir = os.path.join(HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, "liborc.so")
    srcs = [os.path.join(HERE, "sph_oracle.c"), os.path.join(HERE, "sph_oracle.h"),
            os.path.join(HERE, "..", "include", "sph_b200.h")]
    if force or not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        subprocess.check_call(["gcc", "-std=c99", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC",
                               "-shared", "-o", out, srcs[0], "-lm"])
    return out


def build_ref():
    """Compile the unmodified reference into oracle/_ref/ when /root/reference is present.
    Returns the .so path, or None when neither the sources nor a prebuilt .so exist."""
    so = os.path.join(HERE, "_ref", "libsph_ref.so")
    if os.path.isdir(REF_SRC):
        subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "ref_build")])
    return so if os.path.exists(so) else None


def ref_binary():
    p = os.path.join(HERE, "_ref", "sph_ref_run")
    return p if os.path.exists(p) else None


_orc = None


def orc():
    global _orc
    if _orc is None:
        L = C.CDLL(build_oracle())
        L.orc_seq_create.restype = C.c_void_p
        L.orc_seq_create.argtypes = [C.c_int, C.c_float, C.c_float, C.POINTER(Tunable)]
        L.orc_g_create.restype = C.c_void_p
        L.orc_g_create.argtypes = [C.POINTER(Config)]
        L.orc_hash_val.restype = C.c_uint
        L.orc_hash_val.argtypes = [C.c_float, C.c_float, C.c_float, C.c_uint]
        L.orc_spacing.restype = C.c_float
        L.orc_spacing.argtypes = [C.c_float, C.c_float, C.c_int]
        L.orc_g_get_pairs.restype = C.c_longlong
        L.orc_boundary.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_float, C.c_float,
                                   C.POINTER(Tunable)]
        L.orc_partition.argtypes = [C.c_float] * 6 + [C.c_int] + [C.c_void_p] * 4
        L.orc_lattice.argtypes = [C.c_float] * 4 + [C.c_int] * 3 + [C.c_void_p] * 2
        L.orc_g_set_edges.argtypes = [C.c_void_p, C.c_float, C.c_float]
        _orc = L
    return _orc


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


# --------------------------------------------------------------------------- problem set-up

def make_problem(n_request, tank_w=15.0, aspect=16.0 / 9.0, water_frac=1.0, nranks=1):
    """Geometry of start_simulation (fluid.c:116-159) for a scaled tank.
    Returns dict(tank_w, tank_h, spacing, h, n_global, slabs=[(start_col, ncols, start_x, end_x)], total_cols)."""
    L = orc()
    tank_w = float(np.float32(tank_w))
    tank_h = float(np.float32(np.float32(tank_w) / np.float32(aspect)))
    water_w = float(np.float32(np.float32(tank_w) * np.float32(water_frac)))
    spacing = L.orc_spacing(water_w, tank_h, n_request)
    sc = np.zeros(nranks, "i4"); nc = np.zeros(nranks, "i4")
    sx = np.zeros(nranks, "f4"); ex = np.zeros(nranks, "f4")
    n_global = L.orc_partition(tank_w, 0.0, water_w, 0.0, tank_h, spacing, nranks, _p(sc), _p(nc), _p(sx), _p(ex))
    return dict(tank_w=tank_w, tank_h=tank_h, water_w=water_w, spacing=spacing,
                h=float(np.float32(2.0) * np.float32(spacing)), n_global=n_global,
                total_cols=int(nc.sum()),
                slabs=[(int(sc[r]), int(nc[r]), float(sx[r]), float(ex[r])) for r in range(nranks)])


def lattice(prob, rank=0):
    """Initial particles of one rank (geometry.c:29-59), with persistent uids."""
    L = orc()
    sc, nc, _, _ = prob["slabs"][rank]
    num_y = int(np.floor(np.float32(prob["tank_h"]) / np.float32(prob["spacing"])))
    a = np.zeros(nc * num_y, PARTICLE); uid = np.zeros(nc * num_y, "u4")
    n = L.orc_lattice(0.0, 0.0, prob["tank_h"], prob["spacing"], sc, nc, prob["total_cols"], _p(a), _p(uid))
    assert n == len(a)
    return a, uid


# --------------------------------------------------------------------------- orc_seq

class SeqOracle:
    """Reference algorithm as written (Gauss-Seidel, forward lists, caps)."""

    def __init__(self, cap, tank_w, tank_h, tunable):
        self.L = orc()
        self.t = tunable.copy()
        self.h = C.c_void_p(self.L.orc_seq_create(cap, tank_w, tank_h, C.byref(self.t)))
        self.n = 0

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_seq_destroy(self.h); self.h = None

    def load(self, aos):
        aos = np.ascontiguousarray(aos, PARTICLE); self.n = len(aos)
        self.L.orc_seq_load(self.h, _p(aos), len(aos))

    def store(self):
        out = np.zeros(self.n, PARTICLE); self.L.orc_seq_store(self.h, _p(out)); return out

    def apply_gravity(self): self.L.orc_seq_apply_gravity(self.h)
    def viscosity(self): self.L.orc_seq_viscosity(self.h)
    def predict(self): self.L.orc_seq_predict(self.h)
    def hash(self, compute_density): self.L.orc_seq_hash(self.h, int(compute_density))
    def relax(self): self.L.orc_seq_relax(self.h)
    def update_velocities(self): self.L.orc_seq_update_velocities(self.h)

    def step(self, queued=None):
        self.L.orc_seq_step(self.h, C.byref(queued) if queued is not None else None)

    def lists(self):
        """(counts[n], nitems[n, 400]) views copied out of the C struct."""
        class S(C.Structure):
            _fields_ = [("n", C.c_int), ("cap", C.c_int), ("tank_w", C.c_float), ("tank_h", C.c_float),
                        ("t", Tunable)] + [(k, C.c_void_p) for k in
                        ("x", "y", "xp", "yp", "vx", "vy", "dens", "densn", "press", "pressn")] + \
                       [("spacing", C.c_float), ("size_x", C.c_uint), ("size_y", C.c_uint),
                        ("max_bucket", C.c_int), ("max_nbr", C.c_int)] + \
                       [(k, C.c_void_p) for k in ("bcount", "bitems", "ncount", "nitems")]
        s = S.from_address(self.h.value)
        n = s.n
        ncount = np.ctypeslib.as_array(C.cast(s.ncount, C.POINTER(C.c_int)), (n,)).copy()
        nitems = np.ctypeslib.as_array(C.cast(s.nitems, C.POINTER(C.c_int)), (n, s.max_nbr)).copy()
        cells = s.size_x * s.size_y
        bcount = np.ctypeslib.as_array(C.cast(s.bcount, C.POINTER(C.c_int)), (cells,)).copy()
        bitems = np.ctypeslib.as_array(C.cast(s.bitems, C.POINTER(C.c_int)), (cells, s.max_bucket)).copy()
        return ncount, nitems, bcount, bitems, (s.size_x, s.size_y)


# --------------------------------------------------------------------------- orc_g

class GatherOracle:
    """Gather (Jacobi) form; same call surface as sph_b200.Context."""

    def __init__(self, tank_w, tank_h, h, capacity, msg_capacity=1, rank=0, nranks=1, halo_width=2.0):
        self.L = orc()
        self.cfg = Config(tank_w, tank_h, h, capacity, msg_capacity, 0, rank, nranks, halo_width, None)
        self.h = C.c_void_p(self.L.orc_g_create(C.byref(self.cfg)))
        self.capacity = capacity

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_g_destroy(self.h); self.h = None

    def set_params(self, t): self.L.orc_g_set_params(self.h, C.byref(t))
    def queue_params(self, t): self.L.orc_g_queue_params(self.h, C.byref(t))
    def set_edges(self, s, e): self.L.orc_g_set_edges(self.h, s, e)

    def set_viscosity_stabilisation(self, gamma, min_dt_sigma=0.0):
        """Symmetric damping of the viscosity gather for stiff presets; default (0.5, 0.5) like sph_create."""
        self.L.orc_g_set_viscosity_stabilisation_ex.argtypes = [C.c_void_p, C.c_float, C.c_float]
        self.L.orc_g_set_viscosity_stabilisation_ex(self.h, float(gamma), float(min_dt_sigma))
    def set_neighbors(self, l, r): self.L.orc_g_set_neighbors(self.h, int(l), int(r))

    def upload(self, aos, uid=None):
        aos = np.ascontiguousarray(aos, PARTICLE)
        u = None if uid is None else np.ascontiguousarray(uid, "u4")
        rc = self.L.orc_g_upload(self.h, _p(aos), None if u is None else _p(u), len(aos))
        assert rc == 0, rc

    def download(self, order=0, include_halo=False):
        a = np.zeros(self.capacity, PARTICLE); u = np.zeros(self.capacity, "u4")
        n = self.L.orc_g_download(self.h, _p(a), _p(u), order, int(include_halo))
        return a[:n].copy(), u[:n].copy()

    def advect(self): self.L.orc_g_advect(self.h)
    def sort(self): self.L.orc_g_sort(self.h)
    def density(self): self.L.orc_g_density(self.h)
    def relax(self): self.L.orc_g_relax(self.h)
    def step(self, n=1): self.L.orc_g_step(self.h, n)

    def status(self):
        s = Status(); self.L.orc_g_get_status(self.h, C.byref(s)); return s

    def exchange_buffers(self, which):
        sl, rl, sr, rr = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        nb = C.c_size_t()
        self.L.orc_g_exchange_buffers(self.h, which, C.byref(sl), C.byref(rl), C.byref(sr), C.byref(rr), C.byref(nb))
        mk = lambda p: np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_ubyte)), (nb.value,))
        return mk(sl), mk(rl), mk(sr), mk(rr)

    def cells(self):
        u = np.zeros(self.capacity, "u4"); c = np.zeros(self.capacity, "u4")
        n = self.L.orc_g_get_cells(self.h, _p(u), _p(c), self.capacity)
        return u[:n].copy(), c[:n].copy()

    def pairs(self):
        n = self.L.orc_g_get_pairs(self.h, None, C.c_longlong(0))
        out = np.zeros(max(n, 1), "u8")
        n = self.L.orc_g_get_pairs(self.h, _p(out), C.c_longlong(len(out)))
        return np.sort(out[:n])

    def forward_counts(self):
        u = np.zeros(self.capacity, "u4"); c = np.zeros(self.capacity, "i4")
        n = self.L.orc_g_get_forward_counts(self.h, _p(u), _p(c), self.capacity)
        return u[:n].copy(), c[:n].copy()

    def pack_coords(self):
        xy = np.zeros(2 * self.capacity, "i2")
        n = self.L.orc_g_pack_coords(self.h, _p(xy), self.capacity)
        return xy[:2 * n].reshape(n, 2).copy()


# --------------------------------------------------------------------------- the real reference

class RefConfig(C.Structure):
    """== refh_config (oracle/ref_build/ref_harness.c)."""
    _fields_ = [("n_request", C.c_int), ("tank_w", C.c_float), ("tank_h", C.c_float),
                ("water_min_x", C.c_float), ("water_max_x", C.c_float),
                ("water_min_y", C.c_float), ("water_max_y", C.c_float),
                ("mover_cx", C.c_float), ("mover_cy", C.c_float), ("mover_w", C.c_float), ("mover_h", C.c_float),
                ("mover_type", C.c_int), ("steps_per_frame", C.c_int), ("cap_factor", C.c_int)]


class Ref:
    """The unmodified reference functions (oracle/_ref/libsph_ref.so), one rank in-process."""
    _lib = None

    @classmethod
    def available(cls):
        return os.path.exists(os.path.join(HERE, "_ref", "libsph_ref.so"))

    @classmethod
    def lib(cls):
        if cls._lib is None:
            L = C.CDLL(os.path.join(HERE, "_ref", "libsph_ref.so"))
            for name in ("refh_create", "refh_params", "refh_boundary", "refh_grid", "refh_pointers",
                         "refh_particles", "refh_neighbors", "refh_edges", "refh_oob"):
                getattr(L, name).restype = C.c_void_p
            L.refh_spacing.restype = C.c_float
            L.refh_get_neighbor_lists.restype = C.c_long
            L.refh_get_buckets.restype = C.c_long
            L.hash_val.restype = C.c_uint
            L.hash_val.argtypes = [C.c_float, C.c_float, C.c_void_p, C.c_void_p]
            cls._lib = L
        return cls._lib

    def __init__(self, n_request, tank_w=15.0, aspect=16.0 / 9.0, water_frac=1.0, cap_factor=2):
        L = self.lib()
        self.L = L
        tank_w = float(np.float32(tank_w)); tank_h = float(np.float32(np.float32(tank_w) / np.float32(aspect)))
        self.tank_w, self.tank_h = tank_w, tank_h
        cfg = RefConfig(n_request, tank_w, tank_h, 0.0, float(np.float32(tank_w) * np.float32(water_frac)), 0.0, tank_h,
                        0.5 * tank_w, 0.35 * tank_h, 2.0 * tank_w / 15.0, 2.0 * tank_w / 15.0, 0, 4, cap_factor)
        devnull = os.open(os.devnull, os.O_WRONLY); saved = os.dup(1)
        try:   # the reference printf()s its set-up
            os.dup2(devnull, 1)
            self.s = C.c_void_p(L.refh_create(C.byref(cfg)))
            C.CDLL(None).fflush(None)
        finally:
            os.dup2(saved, 1); os.close(saved); os.close(devnull)
        self.params = Param.from_address(L.refh_params(self.s))
        self.p_params = C.c_void_p(L.refh_params(self.s))
        self.p_boundary = C.c_void_p(L.refh_boundary(self.s))
        self.p_grid = C.c_void_p(L.refh_grid(self.s))
        self.p_pointers = C.c_void_p(L.refh_pointers(self.s))
        self.p_particles = C.c_void_p(L.refh_particles(self.s))
        self.p_neighbors = C.c_void_p(L.refh_neighbors(self.s))
        self.p_edges = C.c_void_p(L.refh_edges(self.s))
        self.p_oob = C.c_void_p(L.refh_oob(self.s))

    def __del__(self):
        if getattr(self, "s", None):
            self.L.refh_destroy(self.s); self.s = None

    @property
    def n(self): return self.L.refh_n_local(self.s)
    @property
    def tunable(self): return self.params.tunable_params
    @property
    def h(self): return float(self.params.tunable_params.smoothing_radius)
    @property
    def spacing(self): return float(self.L.refh_spacing(self.s))

    def state(self):
        out = np.zeros(self.n, PARTICLE); self.L.refh_get_state(self.s, _p(out), 0); return out

    def set_state(self, aos):
        aos = np.ascontiguousarray(aos, PARTICLE); self.L.refh_set_state(self.s, _p(aos), len(aos))

    def step(self, n=1):
        for _ in range(n): self.L.refh_step(self.s)

    def queue_params(self, t): self.L.refh_queue_params(self.s, C.byref(t))

    # the reference's own entry points, called directly (fluid.h:112-126, hash.h:50-52)
    def apply_gravity(self): self.L.apply_gravity(self.p_pointers, self.p_params)
    def viscosity_impluses(self): self.L.viscosity_impluses(self.p_pointers, self.p_neighbors, self.p_params)
    def predict_positions(self): self.L.predict_positions(self.p_pointers, self.p_boundary, self.p_params)
    def hash_fluid(self, dens): self.L.hash_fluid(self.p_pointers, self.p_grid, self.p_params, C.c_bool(dens))
    def double_density_relaxation(self): self.L.double_density_relaxation(self.p_pointers, self.p_neighbors, self.p_params)
    def updateVelocities(self): self.L.updateVelocities(self.p_pointers, self.p_edges, self.p_boundary, self.p_params)
    def hash_val(self, x, y): return self.L.hash_val(x, y, self.p_grid, self.p_params)

    def neighbor_lists(self):
        n = self.n
        counts = np.zeros(n, "i4")
        total = self.L.refh_get_neighbor_lists(self.s, _p(counts), None, C.c_long(0))
        flat = np.zeros(max(total, 1), "i4")
        self.L.refh_get_neighbor_lists(self.s, _p(counts), _p(flat), C.c_long(len(flat)))
        return counts, flat[:total]

    def buckets(self, ncells):
        counts = np.zeros(ncells, "i4")
        total = self.L.refh_get_buckets(self.s, _p(counts), None, C.c_long(0))
        flat = np.zeros(max(total, 1), "i4")
        self.L.refh_get_buckets(self.s, _p(counts), _p(flat), C.c_long(len(flat)))
        return counts, flat[:total]
