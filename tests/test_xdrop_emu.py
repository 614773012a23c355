"""The warp-cooperative x-drop bucket replay (lastz_b200/csrc/cuda/xdrop_warp.cuh, the body of k_extend2)
compiled for the host lane emulator (tests/warp_emu) and checked bit for bit against a sequential
restatement of process_for_simple_hit + xdrop_extend_seed_hit (seed_search.c:1056-1192, :2528-2959):
candidates (coordinates, scores, entropy match counts), the bucket's final diagEnd, extension and
column counters.  The emulator also aborts on divergent or abandoned full-mask collectives.
Runs without a GPU; the same source is what nvcc compiles into liblastz_b200.so."""
import os
import subprocess

import pytest

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "warp_emu")


@pytest.mark.parametrize("cap", ["8u", "128u"])      # 8: almost every scan is finished by the warp; 128: the shipped value
def test_bucket_replay_on_lane_emulator(tmp_path, cap):
    exe = str(tmp_path / "test_xdrop")
    subprocess.run(["g++", "-O1", "-std=c++17", "-Wno-unknown-pragmas", f"-DXD_CAP={cap}", "-o", exe,
                    os.path.join(HERE, "test_xdrop.cpp"), os.path.join(HERE, "wemu.cpp")], check=True)
    p = subprocess.run([exe, "1"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert " 0 mismatching" in p.stdout.splitlines()[-1]
