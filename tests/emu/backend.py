"""TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT.

The ctypes binding of sph_b200 pointed at tests/emu/_build/libsph_emu.so (the CUDA source compiled for the host),
as an engine for SlabRunner(backend=...) in the gloo tests: the same C-ABI calls a GPU rank makes, with the
"device" message buffers living in host memory so that torch.distributed (gloo) can move them."""
import ctypes as C

import numpy as np


def use_emulator():
    """Point sph_b200's binding at the emulator library for THIS process (tests only).  SPH_EMU_DEFINES selects a
    build-flag variant of the kernels, e.g. "SPH_ONE_EXCHANGE=1" or "SPH_PACKED=1 SPH_ONE_EXCHANGE=1"."""
    import os
    import sph_b200
    from .build_emu import build
    defines = tuple(os.environ.get("SPH_EMU_DEFINES", "").split())
    name = "libsph_emu" + "".join("_" + d.replace("=", "") for d in defines).lower() + ".so"
    sph_b200._lib = sph_b200._bind(C.CDLL(build(defines=defines, name=name)))
    return sph_b200


class EmuSlab:
    def __init__(self, tank_w, tank_h, h, capacity, msg_capacity, rank, nranks):
        sph = self.sph = use_emulator()
        import os
        # ghost-layer width: the build's default (2 h, or 3.5 h for the one-exchange build) unless the test says otherwise
        self.c = sph.Context(tank_w, tank_h, h, capacity, msg_capacity=msg_capacity, rank=rank, nranks=nranks,
                             halo_width=float(os.environ.get("SPH_EMU_HALO_WIDTH", "0")))
        self._views = {}

    def _t(self, t):
        o = self.sph.Tunable()
        C.memmove(C.byref(o), C.byref(t), 64)       # oracle.Tunable and sph_b200.Tunable share the 64-byte layout
        return o

    def set_params(self, t): self.c.set_params(self._t(t))
    def queue_params(self, t): self.c.queue_params(self._t(t))

    def exchange_buffers(self, which):
        if which not in self._views:
            ptrs, nb = self.c.exchange_pointers(which)
            self._views[which] = [np.ctypeslib.as_array((C.c_ubyte * nb).from_address(p)) for p in ptrs]
        return self._views[which]

    def __getattr__(self, k):
        return getattr(self.c, k)
