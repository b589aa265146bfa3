"""TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT.  Kernel-source emulator: see fake/cuda_runtime.h."""
