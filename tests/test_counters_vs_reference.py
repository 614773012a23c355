"""The two counters bench.py's rates are built on -- raw seed hits and DP cells visited -- as the reference's own
counter build (oracle/_ref/lastz_stats, -Dcollect_stats; seed_search.c:2117, gapped_extend.c:3593/3776) reports them,
against the oracle's.  The first row of every one-sided sweep counts (:3593): bench.py's self-check caught the oracle
and the kernels leaving it out."""
import os
import re
import subprocess

import pytest

from conftest import ORACLE_CLI, ROOT

COUNTER = os.path.join(ROOT, "oracle", "_ref", "lastz_stats")


def _reference_counts(args):
    p = subprocess.run([COUNTER] + args + ["--stats"], capture_output=True, text=True, check=True)
    get = lambda name: int(re.search(name + r":\s*([\d,]+)", p.stderr).group(1).replace(",", ""))
    return get("raw seed hits"), get("DP cells visited")


def _oracle_counts(args):
    p = subprocess.run([ORACLE_CLI] + args + ["--stats"], capture_output=True, text=True, check=True)
    m = re.search(r"raw_seed_hits=(\d+) hsps=\d+ dp_cells=(\d+)", p.stderr)
    return int(m.group(1)), int(m.group(2))


@pytest.mark.skipif(not os.path.exists(COUNTER), reason="oracle/_ref/lastz_stats not built (needs /root/reference at build time)")
@pytest.mark.parametrize("size,extra", [(300000, []), (600000, ["--chain"]), (300000, ["--strand=plus", "--allocate:traceback=2M"])])
def test_oracle_counters_equal_the_reference(synth, size, extra):
    t, q = synth(size)
    assert _oracle_counts([t, q] + extra) == _reference_counts([t, q] + extra)
