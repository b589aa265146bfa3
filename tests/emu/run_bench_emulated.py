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

import bench  # noqa: E402

if __name__ == "__main__":
    sys.argv = ["bench.py"] + sys.argv[1:]
    sys.exit(bench.main())
