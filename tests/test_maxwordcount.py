"""--maxwordcount=<limit>[%] (limit_position_table pos_table.c:1763, find_position_table_limit :2000): words that occur
too often in the target leave the position table.  Front end + library against the unmodified reference: the oracle
library on the CPU, the CUDA library (lzb_target_limit: k_limit_counts, scan, k_limit_compact) on the GPU."""
import os

import pytest

from conftest import GOLDEN, ORACLE_CLI, REF_CLI, ROOT, run_cli

PRODUCT_CLI = os.path.join(ROOT, "lastz_b200", "csrc", "lastz_b200")
CASES = [
    ([os.path.join(GOLDEN, "pseudocat.fa"), os.path.join(GOLDEN, "pseudopig.fa"), "W=8", "T=0", "--nogapped", "--format=segments"], "--maxwordcount=3"),
    ([os.path.join(GOLDEN, "pseudocat.fa"), os.path.join(GOLDEN, "pseudopig.fa"), "W=8", "T=0", "--nogapped", "--format=segments"], "--maxwordcount=50%"),
    ([os.path.join(GOLDEN, "aglobin.2bit") + "/human", os.path.join(GOLDEN, "aglobin.2bit") + "/cow", "--format=general-"], "--maxwordcount=2"),
    ([os.path.join(GOLDEN, "aglobin.2bit") + "/human", os.path.join(GOLDEN, "aglobin.2bit") + "/cow", "--format=general-"], "--maxwordcount=80%"),
    ([os.path.join(GOLDEN, "aglobin.2bit") + "/human", "--self", "--nogapped", "--format=lav"], "--maxwordcount=40"),
]


def _strip(t):
    return [l for l in t.splitlines() if not l.startswith("#")]


def _check(cli, args, opt):
    if not os.path.exists(REF_CLI):
        pytest.skip("oracle/_ref/lastz has not been built")
    want, _ = run_cli(REF_CLI, args + [opt])
    got, _ = run_cli(cli, args + [opt])
    assert _strip(got) == _strip(want)
    plain, _ = run_cli(REF_CLI, args)
    return _strip(want) != _strip(plain)


def test_oracle_front_end_limits_the_table():
    changed = [_check(ORACLE_CLI, a, o) for a, o in CASES]
    assert any(changed), "none of the limits removed a word: the cases test nothing"


@pytest.mark.gpu
def test_product_limits_the_table():
    changed = [_check(PRODUCT_CLI, a, o) for a, o in CASES]
    assert any(changed)
