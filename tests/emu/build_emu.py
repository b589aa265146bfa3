"""TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT.

Builds tests/emu/_build/libsph_emu.so: sph_b200/csrc/sph_capi.cu (with the kernels it includes) compiled
UNCHANGED by g++ against tests/emu/fake/cuda_runtime.h, plus the C host layer.  Same exported symbols as
libsph_b200.so.  Only tests/test_emu_*.py load it; see fake/cuda_runtime.h for what it is and is not.
"""
import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "_build", "libsph_emu.so")


def build(force=False, defines=(), name="libsph_emu.so"):
    """defines / name: build-flag variants of the kernels (tests/test_emu_variants.py)"""
    OUT = os.path.join(HERE, "_build", name)
    tag = os.path.splitext(name)[0]
    pkg = os.path.join(ROOT, "sph_b200")
    inc = os.path.join(ROOT, "include")
    cu = os.path.join(pkg, "csrc", "sph_capi.cu")
    hc = sorted(glob.glob(os.path.join(pkg, "host", "*.c")))
    deps = [cu, os.path.join(HERE, "emu_runtime.cpp"), os.path.join(HERE, "fake", "cuda_runtime.h"), os.path.join(HERE, "fake", "sph_emu_ptx.h"), __file__] + hc + \
        glob.glob(os.path.join(pkg, "csrc", "*.cuh")) + glob.glob(os.path.join(inc, "*.h"))
    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    objs = []
    for src in hc:
        o = os.path.join(HERE, "_build", os.path.basename(src) + ".o")
        subprocess.check_call(["gcc", "-std=gnu99", "-O2", "-ffp-contract=off", "-fPIC", "-I", inc, "-c", src, "-o", o])
        objs.append(o)
    cxx = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-DSPH_EMU", "-w",
           "-I", os.path.join(HERE, "fake"), "-I", inc] + [f"-D{d}" for d in defines]
    o1 = os.path.join(HERE, "_build", tag + ".sph_capi.o")
    subprocess.check_call(cxx + ["-x", "c++", "-c", cu, "-o", o1])
    o2 = os.path.join(HERE, "_build", "emu_runtime.o")
    subprocess.check_call(cxx + ["-c", os.path.join(HERE, "emu_runtime.cpp"), "-o", o2])
    subprocess.check_call(["g++", "-shared", "-o", OUT, o1, o2] + objs + ["-lm"])
    return OUT


if __name__ == "__main__":
    print(build(force=True))
