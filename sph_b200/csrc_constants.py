"""Constants shared with sph_b200/csrc/sph_device.cuh (kept in step by tests/test_capi.py)."""
COST_BASE = 14      # SPH_COST_BASE: neighbour-equivalents of the per-entry part of a slab's work estimate
