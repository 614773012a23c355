"""Parity at bench size: the product command line against the UNMODIFIED reference (oracle/_ref/lastz, built from
/root/reference/src by oracle/build_ref.sh; it travels to the GPU box) on the bench's own synthetic pairs.

  * the 5 Mbp pair in full (LAV, every stanza): 33 k HSPs, ~20 alignments cut by traceback truncation and by each other
  * two 1 Mbp query subranges `q.fa[a..b]` of the 50 Mbp pair against the whole 50 Mbp target -- the reference's own
    way of cutting a query (src/Makefile:536-537) and the cut the multi-GPU runs use -- HSP tables (--format=segments,
    --nogapped) and alignments (LAV)
The reference runs on the box's host cores, one process per subrange, while the product runs on the GPU."""
import os
import subprocess

import pytest

from conftest import GEN_SYNTH, REF_CLI, ROOT

pytestmark = pytest.mark.gpu
PRODUCT_CLI = os.path.join(ROOT, "lastz_b200", "csrc", "lastz_b200")


def _body(text):
    """everything below the d-stanza (it quotes the command line)"""
    lines = text.splitlines()
    for k, l in enumerate(lines):
        if l.startswith("}"):
            return lines[k + 1:]
    return lines


def _synth(tmp_path_factory, size):
    d = tmp_path_factory.mktemp(f"synth{size}")
    t, q = str(d / "t.fa"), str(d / "q.fa")
    subprocess.run([GEN_SYNTH, str(size), "20260925", t, q], check=True)
    return t, q


@pytest.fixture(scope="module")
def pair5(tmp_path_factory):
    return _synth(tmp_path_factory, 5_000_000)


@pytest.fixture(scope="module")
def pair50(tmp_path_factory):
    return _synth(tmp_path_factory, 50_000_000)


def test_5mbp_pair_in_full(pair5):
    if not os.path.exists(REF_CLI):
        pytest.skip("oracle/_ref/lastz has not been built")
    t, q = pair5
    ref = subprocess.Popen([REF_CLI, t, q], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    got = subprocess.run([PRODUCT_CLI, t, q], capture_output=True, text=True, check=True).stdout
    want = ref.communicate()[0]
    assert ref.returncode == 0
    assert _body(got) == _body(want)
    assert sum(1 for l in got.splitlines() if l.startswith("a {")) >= 15


def test_50mbp_pair_query_subranges(pair50):
    if not os.path.exists(REF_CLI):
        pytest.skip("oracle/_ref/lastz has not been built")
    t, q = pair50
    ranges = [(k * 30_000_000 + 5_000_001, k * 30_000_000 + 6_000_000) for k in range(2)]
    refs = []
    for a, b in ranges:                                     # the reference: 8 processes on the host cores
        refs.append((subprocess.Popen([REF_CLI, t, f"{q}[{a}..{b}]"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True),
                     subprocess.Popen([REF_CLI, t, f"{q}[{a}..{b}]", "--nogapped", "--format=segments"], stdout=subprocess.PIPE,
                                      stderr=subprocess.DEVNULL, text=True)))
    gots = []
    for a, b in ranges:
        lav = subprocess.run([PRODUCT_CLI, t, f"{q}[{a}..{b}]"], capture_output=True, text=True, check=True).stdout
        seg = subprocess.run([PRODUCT_CLI, t, f"{q}[{a}..{b}]", "--nogapped", "--format=segments"], capture_output=True, text=True, check=True).stdout
        gots.append((lav, seg))
    for (lav, seg), (rl, rs) in zip(gots, refs):
        want_lav, want_seg = rl.communicate()[0], rs.communicate()[0]
        assert rl.returncode == 0 and rs.returncode == 0
        assert _body(lav) == _body(want_lav)
        assert [l for l in seg.splitlines() if not l.startswith("#")] == [l for l in want_seg.splitlines() if not l.startswith("#")]
        assert sum(1 for l in lav.splitlines() if l.startswith("a {")) >= 2
