"""Every kernel of the seed stage -- k_classify, k_index_words, k_query_words, k_count_hits, k_slot_count,
k_expand, k_bucket_bounds, k_bucket_sizes, k_extend2, the first k_extend, k_extend_alt, k_peaks -- compiled
from the product's own .cuh sources for the host block emulator (tests/warp_emu/cuda_emu.h) and driven by
a host loop that mirrors lzb_seed_hit_search; index contents, raw hit counts, HSP tables (x-drop, raw hits,
--nogfextend, --exact, --mismatch, small hashes) and anchor peaks must equal the oracle's.  No GPU needed."""
import os
import subprocess

from conftest import ROOT

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "warp_emu")


def test_seed_stage_kernels_on_block_emulator(tmp_path):
    exe = str(tmp_path / "test_seed_kernels")
    oracle_dir = os.path.join(ROOT, "oracle")
    subprocess.run(["g++", "-O1", "-std=c++17", "-Wno-unknown-pragmas", "-o", exe,
                    os.path.join(HERE, "test_seed_kernels.cpp"), os.path.join(HERE, "cuda_emu.cpp"),
                    "-L" + oracle_dir, "-llzb_oracle", "-Wl,-rpath," + oracle_dir], check=True)
    p = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert p.stdout.splitlines()[-1].startswith("0 checks failed")
