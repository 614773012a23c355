"""Seed stage parity on the GPU: liblastz_b200.so (through the C-ABI) against the oracle.

Bit-exact bar: index contents, raw hit sets, HSP coordinates + scores and their discovery order.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from lastz_b200 import Engine, default_scoring, parse_seed, read_fasta, revcomp, SEED_14OF22

pytestmark = pytest.mark.gpu

FIELDS = ["pos1", "pos2", "length", "s", "id"]


@pytest.fixture(scope="module")
def engines():
    ss = default_scoring()
    prod, orc = Engine.product(0), Engine.oracle()
    prod.set_scoring(ss)
    orc.set_scoring(ss)
    assert prod.backend == "cuda-sm_100a" and orc.backend == "oracle-cpu"
    yield prod, orc
    prod.close()
    orc.close()


def _pair(synth, size):
    t, q = synth(size)
    return read_fasta(t)[0][1], read_fasta(q)[0][1]


def _same_segments(a, b):
    assert len(a) == len(b), (len(a), len(b))
    for f in FIELDS:
        assert np.array_equal(a[f], b[f]), f


@pytest.mark.parametrize("pattern,step", [(None, 1), (None, 5), (SEED_14OF22, 1), ("11111111", 1), ("1111", 7)])
def test_index_matches_oracle(engines, synth, pattern, step):
    prod, orc = engines
    tseq, _ = _pair(synth, 100000)
    tseq = tseq[:40000] + b"NNNNNNNNNN" + tseq[40000:60000].lower() + tseq[60000:]   # invalid + soft-masked runs
    seed = parse_seed(pattern) if pattern else parse_seed()
    tp, to = prod.build_seed_position_table(tseq, seed, step), orc.build_seed_position_table(tseq, seed, step)
    cp, pp = prod.export_index(tp, seed.weight)
    co, po = orc.export_index(to, seed.weight)
    assert np.array_equal(cp, co)
    assert np.array_equal(pp, po)          # same positions in the same (decreasing) order per word
    prod.free_position_table(tp)
    orc.free_position_table(to)


def test_index_subinterval_and_short(engines):
    prod, orc = engines
    seed = parse_seed()
    for seq, start, end in [(b"ACGT" * 3, 0, 0), (b"ACGTTGCAAGGCTTAACCGGTTAAC" * 40, 13, 777)]:
        tp, to = prod.build_seed_position_table(seq, seed, 1, start, end), orc.build_seed_position_table(seq, seed, 1, start, end)
        cp, pp = prod.export_index(tp, seed.weight)
        co, po = orc.export_index(to, seed.weight)
        assert np.array_equal(cp, co) and np.array_equal(pp, po)


CONFIGS = [
    dict(),                                                   # 12of19 + 1 transition, x-drop, entropy
    dict(seed=("11111111", 0)),                               # W=8 T=0 (base_test.hsp)
    dict(seed=("111010011101", 1)),                           # base_test.seeded
    dict(seed=(None, 2), hsp_threshold=2500),                 # two transitions
    dict(entropy=False, x_drop=400, hsp_threshold=2000),
    dict(gf_extend=0),                                        # --nogfextend with the diag-hash filter
    dict(gf_extend=0, plain_hits=True, seed=("11111111", 0)),  # base_test.hits
    dict(hash_bits=10),                                       # tiny hash: collisions everywhere (diag_hash.h:43-48)
    dict(hash_bits=22),                                       # lastz_32's 4M-entry hash
    dict(gf_extend=2, hsp_threshold=14, seed=("11111111", 0)),  # --exact=14 (match_extend_seed_hit)
    dict(gf_extend=3, gf_mismatches=2, hsp_threshold=20),       # --mismatch=2,20 (mismatch_extend_seed_hit)
    dict(gf_extend=3, gf_mismatches=1, hsp_threshold=12, seed=("111111", 0), hash_bits=9),   # collisions + mismatch mode
]


@pytest.mark.parametrize("cfg", CONFIGS)
def test_hsps_match_oracle_on_fixtures(engines, cfg):
    prod, orc = engines
    cfg = dict(cfg)
    pat, wt = cfg.pop("seed", (None, 1))
    seed = parse_seed(pat, wt) if pat else parse_seed(with_trans=wt)
    tseq = read_fasta(os.path.join(GOLDEN, "pseudocat.fa"))[0][1]
    tp, to = prod.build_seed_position_table(tseq, seed), orc.build_seed_position_table(tseq, seed)
    for _, qseq in read_fasta(os.path.join(GOLDEN, "pseudopig.fa")):
        for strand, s in ((0, qseq), (3, revcomp(qseq))):
            qp, qo = prod.load_query(s), orc.load_query(s)
            a, sa = prod.seed_hit_search(tp, qp, seed, strand_id=strand, **cfg)
            b, sb = orc.seed_hit_search(to, qo, seed, strand_id=strand, **cfg)
            _same_segments(a, b)
            assert sa.rawSeedHits == sb.rawSeedHits and sa.wordsInQuery == sb.wordsInQuery
            if not cfg.get("plain_hits") and cfg.get("gf_extend", 1):
                assert sa.extensions == sb.extensions and sa.bpExtended == sb.bpExtended
            prod.free_query(qp)
            orc.free_query(qo)


@pytest.mark.parametrize("size,cfg", [
    (300000, dict()),
    (300000, dict(hash_bits=8)),
    (1000000, dict()),
])
def test_hsps_match_oracle_on_synthetic(engines, synth, size, cfg, monkeypatch):
    prod, orc = engines
    tseq, qseq = _pair(synth, size)
    seed = parse_seed()
    tp, to = prod.build_seed_position_table(tseq, seed), orc.build_seed_position_table(tseq, seed)
    for strand, s in ((0, qseq), (3, revcomp(qseq))):
        qp, qo = prod.load_query(s), orc.load_query(s)
        b, sb = orc.seed_hit_search(to, qo, seed, strand_id=strand, **cfg)
        for cap in (None, "20000"):          # second run: tiny chunks, diagEnd carried across many chunks
            if cap:
                monkeypatch.setenv("LZB_HIT_CAP", cap)
            else:
                monkeypatch.delenv("LZB_HIT_CAP", raising=False)
            a, sa = prod.seed_hit_search(tp, qp, seed, strand_id=strand, **cfg)
            _same_segments(a, b)
            assert (sa.rawSeedHits, sa.extensions, sa.bpExtended) == (sb.rawSeedHits, sb.extensions, sb.bpExtended)
        prod.free_query(qp)
        orc.free_query(qo)


def test_self_compare_and_subrange(engines):
    """--self on aglobin human-like data: hits on/below the diagonal are dropped (seed_search.c:2182)."""
    prod, orc = engines
    seed = parse_seed()
    seq = read_fasta(os.path.join(GOLDEN, "pseudopig.fa"))[0][1]
    tp, to = prod.build_seed_position_table(seq, seed), orc.build_seed_position_table(seq, seed)
    for strand, s, same in ((0, seq, True), (3, revcomp(seq), False)):
        qp, qo = prod.load_query(s), orc.load_query(s)
        kw = dict(self_compare=True, same_strand=same, strand_id=strand, start=1000, end=len(s) - 500)
        a, sa = prod.seed_hit_search(tp, qp, seed, **kw)
        b, sb = orc.seed_hit_search(to, qo, seed, **kw)
        _same_segments(a, b)
        assert sa.rawSeedHits == sb.rawSeedHits


def test_errors_are_loud(engines):
    prod, _ = engines
    seed = parse_seed()
    tp = prod.build_seed_position_table(b"ACGT" * 100, seed)
    qp = prod.load_query(b"ACGT" * 10)
    with pytest.raises(RuntimeError, match="interval end is bad"):
        prod.seed_hit_search(tp, qp, seed, end=1000)
    with pytest.raises(RuntimeError, match="interval is void"):
        prod.seed_hit_search(tp, qp, seed, start=5, end=5)
