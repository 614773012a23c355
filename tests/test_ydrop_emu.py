"""The Y-drop DP kernel (lastz_b200/csrc/cuda/ydrop_mw.cuh, k_ydrop_mw<8,4> -- 99 % of the bench step's kernel
time) compiled for the host block emulator (tests/warp_emu/cuda_emu.h: one coroutine per CUDA thread,
__syncthreads and the *_sync warp intrinsics as checked rendezvous) and compared with the oracle's
gapped_extend for single anchors: score, end points and the edit script column by column, including
traceback truncation, narrow bands and --noytrim.  Runs without a GPU; the same source is what nvcc
compiles into liblastz_b200.so."""
import os
import subprocess

from conftest import ROOT

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "warp_emu")


def test_ydrop_kernel_on_block_emulator(tmp_path):
    exe = str(tmp_path / "test_ydrop_mw")
    oracle_dir = os.path.join(ROOT, "oracle")
    subprocess.run(["g++", "-O2", "-std=c++17", "-Wno-unknown-pragmas", "-o", exe,
                    os.path.join(HERE, "test_ydrop_mw.cpp"), os.path.join(HERE, "cuda_emu.cpp"),
                    "-L" + oracle_dir, "-llzb_oracle", "-Wl,-rpath," + oracle_dir], check=True)
    p = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert " 0 mismatching" in p.stdout.splitlines()[-1]
