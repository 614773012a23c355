"""The gapped stage's anchor loop (lastz_b200/csrc/cuda/gapped_sched.hpp) on the host: the product's scheduler source over
the product's Y-drop kernels running on the block emulator, job completions delivered in shuffled order, compared with
the oracle's gapped_extend -- alignments op for op and the reference's counters.  Covers what single-anchor kernel tests
cannot: sweeps started speculatively against an older set of alignments, validated against later commits, resumed from
checkpoints, restarted with new neighbours, lanes taken back for the head anchor.  Runs without a GPU."""
import os
import subprocess

from conftest import ROOT

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "warp_emu")


def test_anchor_loop_on_block_emulator(tmp_path):
    exe = str(tmp_path / "test_gapped_sched")
    oracle_dir = os.path.join(ROOT, "oracle")
    subprocess.run(["g++", "-O2", "-std=c++17", "-Wno-unknown-pragmas", "-o", exe,
                    os.path.join(HERE, "test_gapped_sched.cpp"), os.path.join(HERE, "cuda_emu.cpp"),
                    "-L" + oracle_dir, "-llzb_oracle", "-Wl,-rpath," + oracle_dir], check=True)
    p = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert " 0 mismatching" in p.stdout.splitlines()[-1]
