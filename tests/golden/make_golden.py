"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libsph_ref.so and
oracle/_ref/sph_ref_run, built by oracle/ref_build/Makefile from /root/reference/src).

Run in the build container (where /root/reference exists):  python tests/golden/make_golden.py
The reference has no test vectors of its own; these files are its outputs on
deterministic inputs, committed so that the pin holds where /root/reference is absent.
"""
import ctypes as C
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import PARTICLE, PRESETS, Ref, build_ref, ref_binary  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def pairs_of(ref):
    """Symmetric closure of the reference's forward lists as sorted (a<<32|b), a<b, local indices."""
    c, flat = ref.neighbor_lists()
    owner = np.repeat(np.arange(len(c)), c)
    a = np.minimum(owner, flat).astype("u8")
    b = np.maximum(owner, flat).astype("u8")
    return np.sort((a << np.uint64(32)) | b), c


def case(name, n_request, warm_list, preset="x", mover_type=0, water_frac=1.0, tank_w=15.0):
    out = {}
    for warm in warm_list:
        ref = Ref(n_request, tank_w=tank_w, water_frac=water_frac)
        t = ref.tunable
        for k, v in PRESETS[preset].items():
            setattr(t, k, v)
        t.mover_type = bytes([mover_type])
        if mover_type == 1:
            t.mover_height = 0.5 * t.mover_width
        ref.step(warm)
        tag = f"w{warm}"
        out[f"{tag}_state"] = ref.state()
        p, fwd = pairs_of(ref)
        out[f"{tag}_pairs"] = p
        out[f"{tag}_fwd"] = fwd
        st = out[f"{tag}_state"]
        out[f"{tag}_cells"] = np.array([ref.hash_val(float(x), float(y)) for x, y in zip(st["x"], st["y"])], "u4")
        # stage-level known answers from this snapshot, through the reference's own entry points
        ref.apply_gravity(); ref.viscosity_impluses(); ref.predict_positions()
        out[f"{tag}_advect"] = ref.state()
        ref.hash_fluid(True)
        out[f"{tag}_density"] = ref.state()
        p2, _ = pairs_of(ref)
        out[f"{tag}_pairs_pred"] = p2
        ref.double_density_relaxation(); ref.updateVelocities()
        out[f"{tag}_relaxed"] = ref.state()
        ref.hash_fluid(False)
        out[f"{tag}_after1"] = ref.state()
        ref.step(9)
        out[f"{tag}_after10"] = ref.state()
        if warm == warm_list[0]:
            tb = bytes(C.string_at(C.addressof(ref.tunable), 64))
            out["tunable"] = np.frombuffer(tb, "u1").copy()
            out["geom"] = np.array([ref.tank_w, ref.tank_h, ref.h, ref.spacing], "f4")
    # long-run statistics (time-averaged over the last 200 of 1200 steps)
    ref = Ref(n_request, tank_w=tank_w, water_frac=water_frac)
    t = ref.tunable
    for k, v in PRESETS[preset].items():
        setattr(t, k, v)
    t.mover_type = bytes([mover_type])
    if mover_type == 1:
        t.mover_height = 0.5 * t.mover_width
    ref.step(1000)
    acc = []
    for _ in range(200):
        ref.step(1)
        s = ref.state()
        acc.append([s["density"].mean(), s["density"].max(), s["y"].mean(),
                    0.5 * (s["v_x"] ** 2 + s["v_y"] ** 2).mean()])
    out["longrun_stats"] = np.array(acc, "f8").mean(axis=0)   # mean density, max density, mean height, KE/particle
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, {k: getattr(v, "shape", None) for k, v in out.items() if k.endswith("state")}, out["longrun_stats"])


def partition_golden():
    """partitionProblem (geometry.c:101-160) through the reference for several (N, ranks, tank)."""
    L = Ref.lib()
    rows = []

    class AABB(C.Structure):
        _fields_ = [(k, C.c_float) for k in ("min_x", "max_x", "min_y", "max_y", "min_z", "max_z")]

    from oracle.oracle import Param
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    os.dup2(devnull, 1)
    try:
        for n_req, tank_w, frac in ((1500, 15.0, 1.0), (1500, 15.0, 0.5), (100000, 122.47, 1.0), (1000000, 387.3, 0.5)):
            tank_h = float(np.float32(tank_w) / np.float32(16.0 / 9.0))
            for nranks in (1, 2, 3, 4, 7, 8):
                for rank in range(nranks):
                    L.mini_mpi_world_create(nranks, C.c_size_t(4096))
                    L.mini_mpi_bind(rank)
                    b = AABB(0, tank_w, 0, tank_h, 0, 0)
                    w = AABB(0, float(np.float32(tank_w) * np.float32(frac)), 0, tank_h, 0, 0)
                    p = Param()
                    p.number_fluid_particles_global = n_req
                    area = np.float32((w.max_x - w.min_x)) * np.float32((w.max_y - w.min_y))
                    spacing = float(np.float32(np.power(np.float64(area / np.float32(n_req)), 0.5)))
                    xs, lx = C.c_int(), C.c_int()
                    L.partitionProblem(C.byref(b), C.byref(w), C.byref(xs), C.byref(lx), C.c_float(spacing), C.byref(p))
                    rows.append((n_req, tank_w, frac, nranks, rank, spacing, xs.value, lx.value,
                                 p.tunable_params.node_start_x, p.tunable_params.node_end_x,
                                 p.number_fluid_particles_global))
        C.CDLL(None).fflush(None)
    finally:
        os.dup2(saved, 1); os.close(saved); os.close(devnull)
        L.mini_mpi_world_create(1, C.c_size_t(4096)); L.mini_mpi_bind(0)
    np.savez_compressed(os.path.join(OUT, "partition.npz"), rows=np.array(rows, "f8"))
    print("partition rows", len(rows))


def multirank_golden():
    """The reference's own exchange code on 3 ranks (halo, migration, rebalanced edges):
    per-uid state after 200 steps of the default problem, from sph_ref_run."""
    exe = ref_binary()
    with tempfile.TemporaryDirectory() as d:
        for ranks, steps in ((3, 200), (1, 200)):
            base = os.path.join(d, f"r{ranks}")
            subprocess.check_call([exe, "--ranks", str(ranks), "--n", "1500", "--steps", str(steps), "--dump", base],
                                  stdout=subprocess.DEVNULL)
            recs, edges, counts = [], [], []
            for r in range(ranks):
                raw = open(f"{base}.rank{r}.bin", "rb").read()
                hdr = np.frombuffer(raw[:16], "i4")
                edges.append(np.frombuffer(raw[16:24], "f4"))
                recs.append(np.frombuffer(raw[24:], PARTICLE, hdr[0]))
                counts.append(hdr[0])
            a = np.concatenate(recs)
            uid = a["a_x"].view("i4")
            order = np.argsort(uid)
            np.savez_compressed(os.path.join(OUT, f"multirank_r{ranks}.npz"), state=a[order], uid=uid[order],
                                edges=np.array(edges), counts=np.array(counts))
            print("multirank", ranks, counts, np.array(edges).ravel())


if __name__ == "__main__":
    assert build_ref(), "reference sources not available"
    case("default1508", 1500, [100, 400])
    case("goo_rect1508", 1500, [300], preset="y", mover_type=1)
    case("block3000", 3000, [150], water_frac=0.5, tank_w=21.2)
    case("zerog1508", 1500, [200], preset="a")
    case("gas1508", 1500, [200], preset="b")
    partition_golden()
    multirank_golden()
