"""TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT.

Builds tests/emu/_build/libsph_emu.so: sph_b200/csrc/sph_capi.cu (with the kernels it includes) compiled
UNCHANGED by g++ against tests/emu/fake/cuda_runtime.h, plus the C host layer.  Same exported symbols as
libsph_b200.so.  Only tests/test_emu_*.py load it; see fake/cuda_runtime.h for what it is and is not.
"""
import glob
import os
import subprocess
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "_build", "libsph_emu.so")


_lock = threading.Lock()


def _common_objects(pkg, inc, deps):
    """the objects every variant links: the C host layer and the fake runtime (built once, under a lock: prebuild()
    compiles variants from several threads)"""
    with _lock:
        objs = []
        for src in sorted(glob.glob(os.path.join(pkg, "host", "*.c"))):
            o = os.path.join(HERE, "_build", os.path.basename(src) + ".o")
            if not os.path.exists(o) or any(os.path.getmtime(d) > os.path.getmtime(o) for d in deps):
                subprocess.check_call(["gcc", "-std=gnu99", "-O2", "-ffp-contract=off", "-fPIC", "-I", inc, "-c", src, "-o", o + ".tmp"])
                os.replace(o + ".tmp", o)
            objs.append(o)
        o2 = os.path.join(HERE, "_build", "emu_runtime.o")
        if not os.path.exists(o2) or any(os.path.getmtime(d) > os.path.getmtime(o2) for d in deps):
            subprocess.check_call(CXX + ["-I", os.path.join(HERE, "fake"), "-I", inc, "-c", os.path.join(HERE, "emu_runtime.cpp"), "-o", o2 + ".tmp"])
            os.replace(o2 + ".tmp", o2)
        return objs + [o2]


CXX = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-DSPH_EMU", "-w"]


def build(force=False, defines=(), name="libsph_emu.so"):
    """defines / name: build-flag variants of the kernels (tests/test_emu_variants.py).  A library is reused when it is
    newer than every source AND was built with the same defines (kept beside it in <name>.defines)."""
    OUT = os.path.join(HERE, "_build", name)
    tag = os.path.splitext(name)[0]
    pkg = os.path.join(ROOT, "sph_b200")
    inc = os.path.join(ROOT, "include")
    cu = os.path.join(pkg, "csrc", "sph_capi.cu")
    hc = sorted(glob.glob(os.path.join(pkg, "host", "*.c")))
    deps = [cu, os.path.join(HERE, "emu_runtime.cpp"), os.path.join(HERE, "fake", "cuda_runtime.h"), os.path.join(HERE, "fake", "sph_emu_ptx.h"), __file__] + hc + \
        glob.glob(os.path.join(pkg, "csrc", "*.cuh")) + glob.glob(os.path.join(inc, "*.h"))
    want = " ".join(defines)
    side = OUT + ".defines"
    same = os.path.exists(side) and open(side).read() == want
    if not force and same and os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    common = _common_objects(pkg, inc, deps)
    o1 = os.path.join(HERE, "_build", tag + ".sph_capi.o")
    subprocess.check_call(CXX + ["-I", os.path.join(HERE, "fake"), "-I", inc] + [f"-D{d}" for d in defines] + ["-x", "c++", "-c", cu, "-o", o1])
    subprocess.check_call(["g++", "-shared", "-o", OUT + ".tmp", o1] + common + ["-lm"])
    os.replace(OUT + ".tmp", OUT)
    with open(side, "w") as f:
        f.write(want)
    return OUT


# every variant the CPU suite loads: built side by side at the start of a session that runs emulator tests
# (tests/conftest.py), ~10 s each one after the other
_PK = ("SPH_PACKED=1", "SPH_PACKED_RELAX=1")
VARIANTS = {
    "libsph_emu.so": (),
    "libsph_emu_sph_one_exchange1.so": ("SPH_ONE_EXCHANGE=1",),
    "libsph_emu_packed.so": _PK + ("SPH_RELAX_PD4=1",),
    "libsph_emu_trim0.so": ("SPH_TRIM=1",),
    "libsph_emu_trim1.so": ("SPH_TRIM=1",) + _PK,
    "libsph_emu_bf0.so": ("SPH_RELAX_BF=1",),
    "libsph_emu_bf1.so": ("SPH_RELAX_BF=1",) + _PK,
    # round 2, second half (tests/test_emu_variants.py): the build round 2 shipped first; other tile / trip sizes of the
    # shipped machinery; everything that was measured and rejected, switched on together
    "libsph_emu_r2a.so": ("SPH_SORT_SRC=0", "SPH_SCAN_FAST=0", "SPH_ASYNC=0", "SCAN_ITEMS=8", "SPH_RELAX_RARE=0"),
    "libsph_emu_sizes.so": ("SCAN_ITEMS=8", "SPH_SORT_ITEMS=2", "SPH_RELAX_TRIP=2"),
    "libsph_emu_rejected.so": ("SPH_PAIRMASK=1", "SPH_ASYNC=7", "SPH_DEFER=3", "SPH_ADVECT_PV4=1", "SPH_KEYROWS=1", "SPH_PREFETCH=7"),
}


def prebuild(extra=None, workers=None):
    from concurrent.futures import ThreadPoolExecutor
    todo = dict(VARIANTS, **(extra or {}))
    with ThreadPoolExecutor(max_workers=workers or min(len(todo), os.cpu_count() or 2)) as ex:
        return list(ex.map(lambda kv: build(defines=kv[1], name=kv[0]), todo.items()))


if __name__ == "__main__":
    print(build(force=True))
