"""The reference's OWN start_simulation() (fluid.c:71-395, unmodified) against a headless render stub
(oracle/ref_build/ref_drive.c), in two links:

  sph_ref_cpu_drive   the pure reference: its frames must equal the sequential oracle's, which pins the
                      oracle's start-up path (spacing, partition, lattice, parameter block, frame packing)
                      on the reference's own driver, not only on its functions;
  sph_ref_gpu_drive   the same driver with libsph_b200.so in front of it in the lookup order: every
                      hot-path function must bind to the product library, and without a GPU it must
                      stop loudly (the GPU half of this is tests/test_gpu_ref_drive.py).

Plus the reference-named start-up entry points of include/sph_ref_api.h against the reference's golden rows."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from common import GOLDEN
from oracle.oracle import Param, SeqOracle, Tunable, lattice, make_problem

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPU_DRIVE = os.path.join(ROOT, "oracle", "_ref", "sph_ref_cpu_drive")
GPU_DRIVE = os.path.join(ROOT, "oracle", "_ref", "sph_ref_gpu_drive")
HOT = ["apply_gravity", "viscosity_impluses", "predict_positions", "identify_oob_particles", "hash_fluid",
       "hash_halo", "startHaloExchange", "finishHaloExchange", "double_density_relaxation", "updateVelocities",
       "partitionProblem", "setParticleNumbers", "initParticles"]


def read_drive(path):
    """-> (n_global, world_w, world_h, first block, [(block, coords[n, 2] int16)])"""
    raw = open(path, "rb").read()
    assert raw[:4] == b"SPHD"
    n = int(np.frombuffer(raw, "i4", 1, 4)[0])
    w, h = [float(v) for v in np.frombuffer(raw, "f4", 2, 8)]
    frames = int(np.frombuffer(raw, "i4", 1, 16)[0])
    off = 20
    first = Tunable.from_buffer_copy(raw[off:off + 64]); off += 64
    out = []
    for _ in range(frames):
        blk = Tunable.from_buffer_copy(raw[off:off + 64]); off += 64
        pairs = int(np.frombuffer(raw, "i4", 1, off)[0]); off += 4
        out.append((blk, np.frombuffer(raw, "i2", 2 * pairs, off).reshape(pairs, 2).copy())); off += 4 * pairs
    assert off == len(raw)
    return n, w, h, first, out


def pack(x, y, w, h):
    """fluid.c:360-361 in fp32, truncated like the C conversion to short"""
    f = np.float32
    return np.stack([((f(2.0) * x / f(w) - f(1.0)) * f(32767.0)).astype("i2"),
                     ((f(2.0) * y / f(h) - f(1.0)) * f(32767.0)).astype("i2")], axis=1)


WORLD_CPU = os.path.join(ROOT, "oracle", "_ref", "sph_ref_world_cpu")
WORLD_GPU = os.path.join(ROOT, "oracle", "_ref", "sph_ref_world_gpu")
RESTART = os.path.join(ROOT, "oracle", "_ref", "sph_ref_restart")


def read_world(path):
    """oracle/ref_build/render_stubs.c's record -> (K, world_w, world_h, [(edges[K, 2], mover[2], points[n, 2] f32)])"""
    raw = open(path, "rb").read()
    assert raw[:4] == b"SPHR"
    K = int(np.frombuffer(raw, "i4", 1, 4)[0])
    w, h = [float(v) for v in np.frombuffer(raw, "f4", 2, 8)]
    off, frames = 16, []
    while off < len(raw):
        n = int(np.frombuffer(raw, "i4", 1, off)[0]); off += 4
        edges = np.frombuffer(raw, "f4", 2 * K, off).reshape(K, 2).copy(); off += 8 * K
        mover = np.frombuffer(raw, "f4", 2, off).copy(); off += 8
        frames.append((edges, mover, np.frombuffer(raw, "f4", 2 * n, off).reshape(n, 2).copy())); off += 8 * n
    return K, w, h, frames


def bindings(stdout):
    return dict(line.split(": ", 1)[1].split(" -> ") for line in stdout.splitlines() if line.startswith("binding: "))


@pytest.mark.skipif(not os.path.exists(CPU_DRIVE), reason="oracle/_ref not built")
def test_unmodified_reference_driver_equals_sequential_oracle(tmp_path):
    out = str(tmp_path / "cpu.bin")
    r = subprocess.run([CPU_DRIVE, "--frames", "6", "--out", out], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-500:]
    b = bindings(r.stdout)
    assert all(b[k].endswith("libref_driver.so") for k in HOT)          # the pure reference
    n, w, h, first, frames = read_drive(out)
    prob = make_problem(1500)                                              # fluid.c:113,119 and a 1920x1080 screen
    assert (n, w, h) == (prob["n_global"], prob["tank_w"], prob["tank_h"]) and n == 1508
    assert first.smoothing_radius == np.float32(prob["h"]) and first.node_end_x == np.float32(w)
    a, _ = lattice(prob)
    o = SeqOracle(2 * n, w, h, first)
    o.load(a)                                                              # neighbour lists start empty (fluid.c:205)
    for k, (blk, coords) in enumerate(frames):
        o.step(); o.step(); o.step(); o.step(queued=blk)                   # the scatter lands in the 4th sub-step
        s = o.store()
        assert np.array_equal(pack(s["x"], s["y"], w, h), coords), k


@pytest.mark.skipif(not os.path.exists(WORLD_CPU), reason="oracle/_ref not built")
def test_whole_reference_program_runs_headless(tmp_path):
    """The reference's own main(), renderer.c and controls.c (unmodified, fake GL headers, no-op GL modules) with three
    compute ranks over the mini-MPI -- BASELINE config 1, `mpirun -n 4` -- to its clean end through the kill_sim
    scatter: the harness the GPU path is put under in tests/test_emu_ref_drive.py / test_zz_gpu_ref_drive_ranks.py."""
    out = str(tmp_path / "world.bin")
    r = subprocess.run([WORLD_CPU, "--ranks", "3", "--frames", "12", "--out", out], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-500:]
    b = bindings(r.stdout)
    assert all(v.endswith("libref_full.so") for v in b.values()), b          # the pure reference, renderer included
    K, w, h, frames = read_world(out)
    prob = make_problem(1500)
    assert (K, w, h, len(frames)) == (3, prob["tank_w"], prob["tank_h"], 12)
    for edges, mover, pts in frames:
        assert len(pts) == prob["n_global"] and np.all(np.abs(pts) <= 1.0)    # everybody drawn, inside the screen
        assert edges[0, 0] == 0.0 and edges[-1, 1] == np.float32(w) 
        assert np.all(np.abs(edges[1:, 0] - edges[:-1, 1]) < 1e-5)            # (partitionProblem's adjacent edges may differ in the last bit)
    assert frames[5][1][0] > frames[0][1][0]                                 # the stub's "user" dragged the mover


@pytest.mark.skipif(not os.path.exists(GPU_DRIVE), reason="oracle/_ref not built")
def test_driver_binds_to_product_library_and_has_no_cpu_path(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by test_gpu_ref_drive.py")
    r = subprocess.run([GPU_DRIVE, "--frames", "1", "--out", str(tmp_path / "g.bin")], capture_output=True, text=True,
                       timeout=120)
    b = bindings(r.stdout)
    assert b["start_simulation"].endswith("libref_driver.so")             # the reference's driver ...
    assert all(b[k].endswith("libsph_b200.so") for k in HOT), b           # ... on the product's functions
    assert r.returncode != 0 and "there is no CPU path" in r.stderr       # aborts instead of limping on


@pytest.mark.skipif(not os.path.exists(WORLD_GPU), reason="oracle/_ref not built")
def test_whole_program_on_the_library_without_a_gpu_stops_at_once(tmp_path):
    """Three compute ranks under the reference's renderer, product library in front, no device: every rank aborts with
    "there is no CPU path", and the launcher takes the render rank down with them instead of leaving it waiting."""
    import time
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by test_zz_gpu_ref_drive_ranks.py")
    t0 = time.time()
    r = subprocess.run([WORLD_GPU, "--ranks", "3", "--frames", "2", "--out", str(tmp_path / "w.bin")], capture_output=True,
                       text=True, timeout=120)
    b = bindings(r.stdout)
    assert b["start_renderer"].endswith("libref_full.so") and all(b[k].endswith("libsph_b200.so") for k in HOT), b
    assert r.returncode != 0 and "there is no CPU path" in r.stderr and time.time() - t0 < 30


def test_startup_entry_points_match_reference_golden(built_lib):
    """partitionProblem / setParticleNumbers / initParticles of sph_ref_api.h (geometry.c:29-160,
    fluid.c:747-768) against rows produced by the reference's own partitionProblem."""
    L = C.CDLL(built_lib)

    class AABB(C.Structure):
        _fields_ = [(k, C.c_float) for k in ("min_x", "max_x", "min_y", "max_y", "min_z", "max_z")]

    class Edge(C.Structure):            # communication.h:45-52 up to (not including) MPI_Request reqs[4]
        _fields_ = [("max_edge_particles", C.c_int), ("l", C.c_void_p), ("r", C.c_void_p), ("nl", C.c_int), ("nr", C.c_int),
                    ("reqs", C.c_int * 8)]

    class Oob(C.Structure):             # communication.h:55-63
        _fields_ = [("max_oob_particles", C.c_int), ("l", C.c_void_p), ("r", C.c_void_p), ("nl", C.c_int), ("nr", C.c_int),
                    ("vac", C.c_void_p), ("number_vacancies", C.c_int)]

    rows = np.load(os.path.join(GOLDEN, "partition.npz"))["rows"]
    for n_req, tank_w, frac, nranks, rank, spacing, xs, lx, sx, ex, n_global in rows:
        tank_h = float(np.float32(tank_w) / np.float32(16.0 / 9.0))
        b = AABB(0, tank_w, 0, tank_h, 0, 0)
        w = AABB(0, float(np.float32(tank_w) * np.float32(frac)), 0, tank_h, 0, 0)
        p = Param(); p.number_fluid_particles_global = int(n_req)
        L.sph_ref_set_rank(int(rank), int(nranks))
        x0, ln = C.c_int(), C.c_int()
        L.partitionProblem(C.byref(b), C.byref(w), C.byref(x0), C.byref(ln), C.c_float(spacing), C.byref(p))
        assert (x0.value, ln.value, p.number_fluid_particles_global) == (int(xs), int(lx), int(n_global))
        assert p.tunable_params.node_start_x == np.float32(sx) and p.tunable_params.node_end_x == np.float32(ex)
    L.sph_ref_set_rank(0, 1)

    # initParticles == the oracle's lattice (itself pinned on the reference), pointers and counts as geometry.c:52-66
    prob = make_problem(1500)
    a, _ = lattice(prob)
    cap = 2 * len(a)
    from oracle.oracle import PARTICLE
    parts = np.zeros(cap, PARTICLE); parts["v_x"] = 7.0
    ptrs = (C.c_void_p * cap)(*([1] * cap))
    e, o = Edge(nl=5, nr=5), Oob(number_vacancies=9)
    p = Param(); p.number_fluid_particles_global = prob["n_global"]
    water = AABB(0, prob["tank_w"], 0, prob["tank_h"], 0, 0)
    L.setParticleNumbers(C.byref(water), C.byref(water), C.byref(e), C.byref(o), prob["total_cols"], C.c_float(prob["spacing"]), C.byref(p))
    assert (e.max_edge_particles, o.max_oob_particles, o.number_vacancies) == (prob["n_global"],) * 2 + (0,)
    L.initParticles(ptrs, parts.ctypes.data_as(C.c_void_p), C.byref(water), 0, prob["total_cols"], C.byref(e), cap,
                    C.c_float(prob["spacing"]), C.byref(p))
    n = p.number_fluid_particles_local
    assert n == len(a) and p.max_fluid_particle_index == n - 1 and (e.nl, e.nr) == (0, 0)
    for f in ("x", "y", "v_x", "v_y", "a_x", "a_y"):
        assert np.array_equal(parts[f][:n].view("u4"), a[f].view("u4")), f
    assert np.array_equal(parts["id"][:n], np.arange(n))
    base = parts.ctypes.data
    assert [ptrs[i] for i in (0, 1, n - 1)] == [base, base + PARTICLE.itemsize, base + (n - 1) * PARTICLE.itemsize]
    assert all(ptrs[i] is None for i in (n, cap - 1))                      # fluid.c:758-759
