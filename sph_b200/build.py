"""Build libsph_b200.so in-tree: nvcc for sm_100a (the CUDA library + C ABI) and gcc for the C
host layer (sph_b200/host/*.c).  `python -m sph_b200.build`.  nvcc cross-compiles without a GPU."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--threads", "0"]


def _stale(out, deps):
    return not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps)


def build(force=False, verbose=False, defines=(), out_name="libsph_b200.so"):
    """defines / out_name: tuning variants for A/B runs on the GPU box (scripts/gpu_variants.sh)."""
    out = os.path.join(HERE, out_name)
    csrc = os.path.join(HERE, "csrc")
    host = os.path.join(HERE, "host")
    inc = os.path.join(HERE, "..", "include")
    cu = sorted(glob.glob(os.path.join(csrc, "*.cu")))
    hc = sorted(glob.glob(os.path.join(host, "*.c")))
    deps = cu + hc + glob.glob(os.path.join(csrc, "*.cuh")) + glob.glob(os.path.join(host, "*.h")) + \
        glob.glob(os.path.join(inc, "*.h"))
    if not force and not _stale(out, deps):
        return out
    bdir = os.path.join(HERE, "build")
    os.makedirs(bdir, exist_ok=True)
    objs = []
    for src in hc:
        o = os.path.join(bdir, os.path.basename(src) + ".o")
        subprocess.check_call(["gcc", "-std=gnu99", "-O2", "-ffp-contract=off", "-fPIC", "-Wall", "-I", inc, "-c", src, "-o", o])
        objs.append(o)
    cmd = ["nvcc"] + NVCC_FLAGS + [f"-D{d}" for d in defines] + (["-Xptxas", "-v"] if verbose else []) + \
        ["-I", inc, "-shared", "-o", out] + cu + objs
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    # python -m sph_b200.build [--force] [-v] [--variant NAME -DSPH_X=1 ...]  ->  sph_b200/variants/NAME.so
    if "--variant" in sys.argv:
        name = sys.argv[sys.argv.index("--variant") + 1]
        os.makedirs(os.path.join(HERE, "variants"), exist_ok=True)
        print(build(force=True, verbose="-v" in sys.argv, defines=[a[2:] for a in sys.argv if a.startswith("-D")],
                    out_name=os.path.join("variants", name + ".so")))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
