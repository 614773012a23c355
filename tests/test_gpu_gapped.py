"""Gapped stage parity on the GPU: Y-drop DP + traceback + anchor loop against the oracle.

Bit-exact bar: alignment end points, scores, edit scripts (op for op) and the number of DP cells
the reference would have visited.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from lastz_b200 import Engine, default_scoring, parse_seed, read_fasta, revcomp

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engines():
    ss = default_scoring()
    prod, orc = Engine.product(0), Engine.oracle()
    prod.set_scoring(ss)
    orc.set_scoring(ss)
    yield prod, orc
    prod.close()
    orc.close()


def _compare(prod, orc, tseq, qseq, strand, seed, gap_kw, seed_kw=None):
    seed_kw = seed_kw or {}
    tp, to = prod.build_seed_position_table(tseq, seed), orc.build_seed_position_table(tseq, seed)
    qp, qo = prod.load_query(qseq), orc.load_query(qseq)
    hsps, _ = orc.seed_hit_search(to, qo, seed, strand_id=strand, **seed_kw)
    ap = prod.reduce_to_points(tp, qp, hsps.copy())
    ao = orc.reduce_to_points(to, qo, hsps.copy())
    for f in ("pos1", "pos2", "length", "s"):
        assert np.array_equal(ap[f], ao[f]), f                      # K4 anchor peaks
    okw = {k: v for k, v in gap_kw.items() if k != "speculation"}
    want, so, _ = orc.gapped_extend(to, qo, tseq, qseq, ao, **okw)
    got, sp, _ = prod.gapped_extend(tp, qp, tseq, qseq, ap, **gap_kw)
    assert len(got) == len(want), (len(got), len(want))
    for g, w in zip(got, want):
        assert (g["beg1"], g["end1"], g["beg2"], g["end2"], g["s"]) == (w["beg1"], w["end1"], w["beg2"], w["end2"], w["s"])
        assert np.array_equal(g["ops"], w["ops"])
    assert sp.anchorsExtended == so.anchorsExtended
    # the metric numerator: cells of the DPs whose results were used equal the reference counter
    # (gapped_extend.c:3593,3776) whatever was speculated on the side
    assert sp.dpCells == so.dpCells and sp.truncated == so.truncated
    assert sp.dpCellsComputed >= sp.dpCells
    for e, h in ((prod, (tp, qp)), (orc, (to, qo))):
        e.free_position_table(h[0])
        e.free_query(h[1])
    return got, sp


@pytest.mark.parametrize("spec", [1, 16])
def test_default_pipeline_on_fixtures(engines, spec):
    prod, orc = engines
    seed = parse_seed()
    tseq = read_fasta(os.path.join(GOLDEN, "pseudocat.fa"))[0][1]
    for _, qseq in read_fasta(os.path.join(GOLDEN, "pseudopig.fa")):
        for strand, s in ((0, qseq), (3, revcomp(qseq))):
            _compare(prod, orc, tseq, s, strand, seed, dict(speculation=spec))


@pytest.mark.parametrize("size,kw", [
    (200000, dict(speculation=1)),
    (200000, dict(speculation=16)),
    (300000, dict(speculation=1, traceback_bytes=2 * 1024 * 1024)),      # truncation, many bounded alignments
    (300000, dict(speculation=16, traceback_bytes=2 * 1024 * 1024)),
    (300000, dict(speculation=8, traceback_bytes=1024 * 1024, all_bounds=True)),
    (200000, dict(speculation=4, y_drop=3000, score_threshold=5000)),
    (200000, dict(speculation=4, trim_to_peak=False)),
    (600000, dict(speculation=256, traceback_bytes=1024 * 1024)),       # hundreds of sweeps in one launch, most of them resumed from a checkpoint
    (600000, dict(speculation=48, traceback_bytes=2 * 1024 * 1024, trim_to_peak=False)),   # fewer lanes than anchors worth starting
])
def test_synthetic_pairs(engines, synth, size, kw):
    prod, orc = engines
    t, q = synth(size)
    tseq, qseq = read_fasta(t)[0][1], read_fasta(q)[0][1]
    seed = parse_seed()
    for strand, s in ((0, qseq), (3, revcomp(qseq))):
        _compare(prod, orc, tseq, s, strand, seed, kw)


def test_sequence_ends_and_noise(engines, synth):
    """anchors near both sequence ends (N, M small), lowercase/N runs inside the band"""
    prod, orc = engines
    t, q = synth(100000)
    tseq, qseq = read_fasta(t)[0][1], read_fasta(q)[0][1]
    tseq = tseq[:30000] + tseq[30000:30400].lower() + tseq[30400:60000] + b"N" * 37 + tseq[60037:]
    qseq = qseq[200:]                                                   # alignment runs into seq2's start
    seed = parse_seed()
    _compare(prod, orc, tseq, qseq, 0, seed, dict(speculation=4))
    _compare(prod, orc, tseq[:50000], qseq[:20000], 0, seed, dict(speculation=1))


@pytest.mark.parametrize("mode", ["1", "2"])
def test_fallback_kernels(engines, synth, monkeypatch, mode):
    """the one-warp kernel and the shared-memory kernel (what a band too wide for the register window falls back to)
    under the same scheduler: no checkpoints there, so a reached sweep restarts from its first row"""
    prod, orc = engines
    monkeypatch.setenv("LZB_DP_MODE", mode)
    t, q = synth(200000)
    tseq, qseq = read_fasta(t)[0][1], read_fasta(q)[0][1]
    _compare(prod, orc, tseq, qseq, 0, parse_seed(), dict(speculation=16, traceback_bytes=1024 * 1024))


def test_checkpoint_interval(engines, synth, monkeypatch):
    """a different checkpoint spacing must not change anything (fresh context: the spacing is fixed when the lanes are made)"""
    from lastz_b200 import Engine, default_scoring
    monkeypatch.setenv("LZB_CKPT_EVERY", "96")
    prod2 = Engine.product(0)
    prod2.set_scoring(default_scoring())
    t, q = synth(300000)
    tseq, qseq = read_fasta(t)[0][1], read_fasta(q)[0][1]
    _compare(prod2, engines[1], tseq, qseq, 0, parse_seed(), dict(speculation=32, traceback_bytes=512 * 1024))
    prod2.close()


def test_identical_sequences_trivial_alignment(engines):
    prod, orc = engines
    seq = read_fasta(os.path.join(GOLDEN, "pseudopig.fa"))[0][1]
    seed = parse_seed()
    got, _ = _compare(prod, orc, seq, seq, 0, seed, dict(speculation=4, identity_check=True),
                      seed_kw=dict(self_compare=True, same_strand=True))
    assert got[0]["isTrivial"] == 1
    _compare(prod, orc, seq, seq, 0, seed, dict(speculation=4, identity_check=True, inhibit_trivial=True),
             seed_kw=dict(self_compare=True, same_strand=True))
