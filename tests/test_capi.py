"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/sph_b200.h
declares.  No compute calls here (no GPU in this container); on a box without a device the
library must refuse loudly instead of falling back to anything."""
import ctypes as C
import os
import re

import pytest


def test_library_exports_every_declared_symbol(built_lib):
    import sph_b200
    L = C.CDLL(built_lib)
    hdr = open(os.path.join(os.path.dirname(built_lib), "..", "include", "sph_b200.h")).read()
    declared = set(re.findall(r"\b(sph_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(sph_b200.C_ABI_SYMBOLS), declared ^ set(sph_b200.C_ABI_SYMBOLS)
    for name in declared:
        assert hasattr(L, name), name


def test_record_layouts_match_reference():
    import sph_b200
    assert sph_b200.PARTICLE.itemsize == 52 and sph_b200.PARTICLE.fields["id"][1] == 48       # fluid.h:56-70
    assert C.sizeof(sph_b200.Tunable) == 64                                                   # fluid.h:78-97
    assert sph_b200.Tunable.mover_type.offset == 60


def test_no_cpu_fallback(built_lib):
    """Without a CUDA device sph_create must fail with SPH_ERR_CUDA (1) and say why."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import sph_b200
    sph_b200._lib = None          # (an emulator test earlier in the same process may have rebound the module's library)
    with pytest.raises(sph_b200.SphError) as e:
        sph_b200.Context(15.0, 8.4375, 0.58, 2048)
    assert "-> 1" in str(e.value)


def test_sass_is_sm100a(built_lib):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", built_lib], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_python_constants_match_the_device_header():
    import re
    from sph_b200.csrc_constants import COST_BASE
    hdr = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "sph_b200", "csrc", "sph_device.cuh")).read()
    assert int(re.search(r"#define SPH_COST_BASE (\d+)", hdr).group(1)) == COST_BASE


def test_shipped_kernels_are_the_profiled_kernels(built_lib):
    """profiles/r2b_sass_hashes.json holds a hash of the SASS of every kernel of the build that was measured and
    profiled on the B200 at the end of round 2 (scripts/sass_hash.py; the sets of round 1 and of round 2's first build
    are kept beside it).  The default build must
    still contain exactly that machine code -- refactors and build-flag variants may not disturb the measured path
    without the profiles being redone (then refresh the hashes together with profiles/)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "scripts", "sass_hash.py"), built_lib, "--against",
                        os.path.join(root, "profiles", "r2b_sass_hashes.json")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
    assert r.stdout.count("same") >= 18 and "CHANGED" not in r.stdout and "NEW" not in r.stdout, r.stdout


def test_pdl_variant_builds_and_every_kernel_waits_before_it_reads(built_lib):
    """-DSPH_PDL=1: every kernel entry point must carry the grid-dependency wait (SASS ACQBULK) and
    the early trigger (PREEXIT), the wait before the first global load -- a kernel without it would read what its
    predecessor in the stream is still writing.  The default build has neither instruction."""
    import subprocess
    from sph_b200.build import build
    os.makedirs(os.path.join(os.path.dirname(built_lib), "variants"), exist_ok=True)
    lib = build(force=True, defines=("SPH_PDL=1",), out_name=os.path.join("variants", "pdl_test.so"))
    per_kernel, name = {}, None
    for line in subprocess.run(["cuobjdump", "-sass", lib], check=True, capture_output=True, text=True).stdout.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1); per_kernel[name] = []
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if m and name:
            per_kernel[name].append(m.group(1))
    assert len(per_kernel) >= 17
    for k, ops in per_kernel.items():
        assert "ACQBULK" in ops and "PREEXIT" in ops, k
        first_global = min(i for i, o in enumerate(ops) if o.startswith(("LDG", "STG", "ATOMG", "RED", "LD.", "ST.")))
        assert ops.index("ACQBULK") < first_global, k
    base = subprocess.run(["cuobjdump", "-sass", built_lib], check=True, capture_output=True, text=True).stdout
    assert "ACQBULK" not in base and "PREEXIT" not in base
    os.remove(lib)


def test_every_build_flag_tested_in_the_sources_has_a_default():
    """An `#if SPH_X` whose macro is defined nowhere silently reads as 0 (round 2 lost the defaults of two flags to an
    editing slip and shipped a slower kernel for a few commits): every SPH_ flag tested by the preprocessor in
    sph_b200/csrc must have an `#ifndef SPH_X / #define SPH_X default` block (or a plain #define) in those sources."""
    import glob
    import re
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    src = "".join(open(f).read() for f in glob.glob(os.path.join(root, "sph_b200", "csrc", "*.cu*")) +
                  glob.glob(os.path.join(root, "include", "*.h")))
    used = set(re.findall(r"^\s*#\s*(?:if|elif)[^\n]*?\b(SPH_[A-Z0-9_]+)", src, re.M)) | \
        set(m for line in re.findall(r"^\s*#\s*(?:if|elif)([^\n]*)", src, re.M) for m in re.findall(r"\bSPH_[A-Z0-9_]+", line))
    defined = set(re.findall(r"^\s*#\s*define\s+(SPH_[A-Z0-9_]+)", src, re.M))
    missing = sorted(u for u in used if u not in defined and u != "SPH_EMU")        # SPH_EMU: only the test suite defines it
    assert not missing, missing
