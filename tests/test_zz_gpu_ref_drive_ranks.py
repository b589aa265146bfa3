"""The reference's UNMODIFIED start_simulation() as SEVERAL compute ranks on the B200 path (written after round 1's GPU
budget was spent; the same check passes on the kernel-source emulator, tests/test_emu_ref_drive.py; this file sorts
last among the GPU files on purpose).  Every rank is a process of its own with one slab; all of them share device 0
unless SPH_B200_DEVICES says how many devices to spread over.  Run on the B200 box: pytest -m gpu."""
import os

import pytest

from test_emu_ref_drive import (KEYS, check_config1_against_the_pure_reference, check_ranks_against_one_rank,
                                check_restart_from_a_moving_fluid, check_whole_program)
from test_ref_drive import GPU_DRIVE, RESTART, WORLD_GPU

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not os.path.exists(GPU_DRIVE), reason="oracle/_ref not built")
@pytest.mark.parametrize("ranks,wobble", [(3, 0), (3, 1)])
def test_unmodified_reference_driver_with_several_compute_ranks_on_the_gpu_path(built_lib, tmp_path, ranks, wobble):
    import torch
    env = dict(os.environ)
    if torch.cuda.device_count() > 1:
        env["SPH_B200_DEVICES"] = str(torch.cuda.device_count())
    check_ranks_against_one_rank(GPU_DRIVE, env, tmp_path, ranks, 10, "libsph_b200.so", wobble)


@pytest.mark.skipif(not os.path.exists(WORLD_GPU), reason="oracle/_ref not built")
@pytest.mark.parametrize("frames,script", [(14, None), (20, KEYS)])
def test_whole_reference_program_with_its_renderer_on_the_gpu_path(built_lib, tmp_path, frames, script):
    """BASELINE config 1 (mpirun -n 4: the reference's render rank + 3 compute ranks), hot path on the B200."""
    import torch
    env = dict(os.environ)
    if torch.cuda.device_count() > 1:
        env["SPH_B200_DEVICES"] = str(torch.cuda.device_count())
    check_whole_program(WORLD_GPU, env, tmp_path, 3, frames, "libsph_b200.so", script)


@pytest.mark.skipif(not os.path.exists(WORLD_GPU), reason="oracle/_ref not built")
def test_config1_whole_program_statistics_against_the_pure_reference_on_the_gpu_path(built_lib, tmp_path):
    """BASELINE config 1 end to end: the reference's whole program, pure vs with its hot path on the B200."""
    print("worst deviation (mean y, std y, mean x, std x):", check_config1_against_the_pure_reference(dict(os.environ), tmp_path))


@pytest.mark.skipif(not os.path.exists(RESTART), reason="oracle/_ref not built")
def test_restart_from_a_moving_fluid_through_the_reference_names_on_the_gpu_path(built_lib, tmp_path):
    check_restart_from_a_moving_fluid(dict(os.environ), tmp_path)
