"""A few fixed seeds of the two randomised harnesses under tests/fuzz/ (kernel-source emulator, no GPU): call sequences
against the C ABI's host logic, slab runs with a walking mover, moving edges and changing presets against one slab.
More seeds by hand: python tests/fuzz/fuzz_api.py 0 40; WALK=1 python tests/fuzz/fuzz_slabs.py 0 60; python tests/fuzz/fuzz_p2p.py 0 40."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def test_random_call_sequences_leave_the_simulation_intact(built_lib):
    r = subprocess.run([sys.executable, os.path.join(HERE, "fuzz", "fuzz_api.py"), "100", "2"], capture_output=True, text=True,
                       timeout=600, env=dict(os.environ, FUZZ_OPS="30"))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count(" ok") == 2


def test_random_slab_runs_with_a_walking_mover_equal_one_slab_bit_for_bit(built_lib):
    # seeds 5-10: 2, 3 and 4 slabs; two exchanges per step, one per step, one every two steps
    r = subprocess.run([sys.executable, os.path.join(HERE, "fuzz", "fuzz_slabs.py"), "5", "6"], capture_output=True, text=True,
                       timeout=600, env=dict(os.environ, WALK="1"))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count(" ok ") == 6, r.stdout


def test_random_scripts_through_the_peer_memory_protocol_equal_one_slab(built_lib):
    # seeds 103-108: 2 and 3 emulated devices; sph_step in pieces, queued blocks, snapshots / restores at barriers, a parked slab
    r = subprocess.run([sys.executable, os.path.join(HERE, "fuzz", "fuzz_p2p.py"), "103", "6"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count(" ok ") >= 4, r.stdout


def test_random_soups_against_the_gather_oracle(built_lib):
    # seeds 0-7: uniform and clustered soups, all presets, both mover types
    r = subprocess.run([sys.executable, os.path.join(HERE, "fuzz", "fuzz_soup.py"), "0", "8"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count(" ok ") == 8, r.stdout


def test_reference_whole_program_under_a_random_user(built_lib):
    # seeds 0-2: 2 and 3 compute ranks, presets and remove / add_partition pressed at random frames
    from test_ref_drive import WORLD_GPU
    import pytest
    if not os.path.exists(WORLD_GPU):
        pytest.skip("oracle/_ref not built")
    r = subprocess.run([sys.executable, os.path.join(HERE, "fuzz", "fuzz_world.py"), "0", "3"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count(" ok ") == 3, r.stdout
