"""lastz_b200 -- B200-native (sm_100a) seed-and-extend hot path of LASTZ behind a C-ABI.

The compute lives in lastz_b200/csrc (CUDA kernels + C host front end); this package is the thin
Python binding used by tests/, bench.py and __graft_entry__.py.
"""
from . import capi  # noqa: F401
from .api import (Engine, SEED_12OF19, SEED_14OF22, default_scoring, merge_segments, parse_seed, read_fasta,  # noqa: F401
                  reduce_to_chain, revcomp)
