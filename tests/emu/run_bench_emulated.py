"""TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT.

Runs bench.py's own main() in a container without a GPU: the ctypes binding is pointed at the kernel-source
emulator and torch.cuda is replaced by stand-ins (events read the host clock).  The numbers it prints mean
nothing; the point is that every line of the bench's single-GPU path -- timed loop, stage times, both
end-to-end protocols, statistics, JSON assembly -- has executed before the driver runs it on a B200.
    python tests/emu/run_bench_emulated.py --particles 3000 --steps 4 --warmup 3 --preroll 8 --no-cpu-baseline
"""
import contextlib
import ctypes as C
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.dirname(HERE)]

import torch  # noqa: E402

import sph_b200  # noqa: E402
from emu.build_emu import build  # noqa: E402

sph_b200._lib = sph_b200._bind(C.CDLL(build()))


class Event:
    def __init__(self, enable_timing=False):
        self.t = None

    def record(self, stream=None):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return max((other.t - self.t) * 1e3, 1e-6)

    def synchronize(self):
        pass


class Stream:
    cuda_stream = None


_empty = torch.empty
torch.empty = lambda *a, device=None, **k: _empty(*([min(a[0], 1 << 16)] if a and isinstance(a[0], int) and a[0] > (1 << 24) else a), **k)
torch.Tensor.pin_memory = lambda self: self
torch.cuda.is_available = lambda: True
torch.cuda.set_device = lambda d: None
torch.cuda.synchronize = lambda *a: None
torch.cuda.Stream = Stream
torch.cuda.Event = Event
torch.cuda.stream = lambda s: contextlib.nullcontext()
torch.cuda.current_device = lambda: 0

_tensor = torch.tensor
torch.tensor = lambda *a, device=None, **k: _tensor(*a, **k)

if int(os.environ.get("WORLD_SIZE", "1")) > 1:
    # several ranks (python -m torch.distributed.run ... run_bench_emulated.py --gpus N): gloo instead of NCCL, every rank's
    # slab on the emulator with host message buffers that torch.distributed moves (the "collective" transport of
    # SlabRunner; the peer-memory protocol itself is exercised by tests/test_emu_p2p.py).  What this executes is
    # bench.py's own N > 1 path: parity pre-check, time-balanced edges, exchange period, reductions over ranks, the
    # per-slab table, the config-3 section, JSON assembly.
    import torch.distributed as dist

    import sph_b200.slab as slab
    from emu.backend import EmuSlab

    os.environ.setdefault("SPH_EMU_DEFINES", "SPH_ONE_EXCHANGE=1")
    _init = dist.init_process_group
    dist.init_process_group = lambda backend=None, device_id=None, **k: _init("gloo", **k)
    _slab_init = slab.SlabRunner.__init__

    def _emulated_slab(self, prob, tunable, rank, world, stream=None, capacity_factor=2.0, backend=None, **k):
        hw = k.get("halo_width")
        if not hw and int(k.get("exchange_period", 1)) > 1:
            hw = (4.5 if tunable.time_step * tunable.sigma >= 0.5 else 3.5) * int(k["exchange_period"])
        os.environ["SPH_EMU_HALO_WIDTH"] = str(hw or 0)
        k.pop("exchanges_per_step", None)          # (the emulator backend is the one-exchange build)
        _slab_init(self, prob, tunable, rank, world, stream=None, capacity_factor=capacity_factor, backend=EmuSlab, **k)

    slab.SlabRunner.__init__ = _emulated_slab

import bench  # noqa: E402

if __name__ == "__main__":
    sys.argv = ["bench.py"] + sys.argv[1:]
    sys.exit(bench.main())
