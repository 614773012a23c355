"""--recoverseeds and --twins on the GPU: k_extend_recover (process_for_recoverable_hit, seed_search.c:1221) against the oracle
through the C-ABI, and the product command line (with merge_segments) against the unmodified reference.

Written after the round's GPU minutes were spent: the kernel's own source runs against the oracle on the block emulator
(tests/test_seed_kernels_emu.py: four --recoverseeds modes incl. 64- and 128-bucket hashes), the oracle and the host
front end are compared with the reference binary on the CPU (tests/test_oracle_golden.py::test_oracle_recoverseeds), but
these pytest cases have no GPU run behind them yet (five of their command lines were checked on a B200 by hand, profiles/r02_hit_processors_on_gpu.txt) -- which is why the file sorts last (pytest -x stops at the first failure)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, ORACLE_CLI, PRODUCT_CLI, REF_CLI, run_cli, same_output
from lastz_b200 import Engine, default_scoring, merge_segments, parse_seed, read_fasta, revcomp

pytestmark = pytest.mark.gpu

CAT = os.path.join(GOLDEN, "pseudocat.fa")
PIG = os.path.join(GOLDEN, "pseudopig.fa")
FIELDS = ["pos1", "pos2", "length", "s", "id"]


@pytest.mark.parametrize("cfg", [dict(), dict(hash_bits=8), dict(gf_extend=0, hash_bits=10), dict(entropy=False, hsp_threshold=2200, hash_bits=6)])
def test_recoverable_processor_matches_oracle(cfg, synth, monkeypatch):
    ss = default_scoring()
    prod, orc = Engine.product(0), Engine.oracle()
    prod.set_scoring(ss)
    orc.set_scoring(ss)
    seed = parse_seed()
    t, q = synth(300000)
    pairs = [(read_fasta(CAT)[0][1], read_fasta(PIG)[0][1]), (read_fasta(t)[0][1], read_fasta(q)[0][1])]
    for tseq, qseq in pairs:
        tp, to = prod.build_seed_position_table(tseq, seed), orc.build_seed_position_table(tseq, seed)
        for strand, s in ((0, qseq), (3, revcomp(qseq))):
            qp, qo = prod.load_query(s), orc.load_query(s)
            b, sb = orc.seed_hit_search(to, qo, seed, strand_id=strand, recover_seeds=True, **cfg)
            for cap in (None, "20000"):          # second run: tiny chunks, extent and actual diagonal carried across chunks
                if cap:
                    monkeypatch.setenv("LZB_HIT_CAP", cap)
                else:
                    monkeypatch.delenv("LZB_HIT_CAP", raising=False)
                a, sa = prod.seed_hit_search(tp, qp, seed, strand_id=strand, recover_seeds=True, **cfg)
                assert len(a) == len(b)
                for f in FIELDS:
                    assert np.array_equal(a[f], b[f]), f
                assert sa.rawSeedHits == sb.rawSeedHits
                if cfg.get("gf_extend", 1):
                    assert (sa.extensions, sa.bpExtended) == (sb.extensions, sb.bpExtended)
            m = merge_segments(a)
            assert len(m) <= len(a) and (len(m) == 0 or np.all(np.diff((m["pos1"].astype(np.int64) - m["pos2"].astype(np.int64))) >= 0))
            prod.free_query(qp)
            orc.free_query(qo)
        prod.free_position_table(tp)
        orc.free_position_table(to)
    prod.close()
    orc.close()


@pytest.mark.parametrize("opts", [["--recoverseeds"], ["--recoverseeds", "--nogapped", "--format=general-"],
                                  ["--recoverseeds", "--nogfextend", "--nogapped", "--format=general-"], ["--recoverseeds", "--chain", "--format=maf-"]],
                         ids=lambda o: " ".join(o))
def test_cli_recoverseeds(synth, opts):
    ref = REF_CLI if os.path.exists(REF_CLI) else ORACLE_CLI
    for files in ([CAT, PIG], list(synth(300000))):
        same_output(run_cli(PRODUCT_CLI, files + opts)[0], run_cli(ref, files + opts)[0])


@pytest.mark.parametrize("cfg", [dict(twin_spans=(38, 88)), dict(twin_spans=(33, 68), hash_bits=8), dict(twin_spans=(48, 238), gf_extend=0, hash_bits=10)])
def test_twin_processor_matches_oracle(cfg, synth, monkeypatch):
    """k_extend_twin (process_for_twin_hit): the bucket replay with queue entries carried across chunks (second run: chunks
    of about 20000 hits) against the oracle's global queue"""
    ss = default_scoring()
    prod, orc = Engine.product(0), Engine.oracle()
    prod.set_scoring(ss)
    orc.set_scoring(ss)
    seed = parse_seed()
    t, q = synth(300000)
    tseq, qseq = read_fasta(t)[0][1], read_fasta(q)[0][1]
    tp, to = prod.build_seed_position_table(tseq, seed), orc.build_seed_position_table(tseq, seed)
    for strand, s in ((0, qseq), (3, revcomp(qseq))):
        qp, qo = prod.load_query(s), orc.load_query(s)
        b, sb = orc.seed_hit_search(to, qo, seed, strand_id=strand, **cfg)
        for cap in (None, "20000"):
            if cap:
                monkeypatch.setenv("LZB_HIT_CAP", cap)
            else:
                monkeypatch.delenv("LZB_HIT_CAP", raising=False)
            a, sa = prod.seed_hit_search(tp, qp, seed, strand_id=strand, **cfg)
            assert len(a) == len(b)
            for f in FIELDS:
                assert np.array_equal(a[f], b[f]), f
            assert sa.rawSeedHits == sb.rawSeedHits
            if cfg.get("gf_extend", 1):
                assert (sa.extensions, sa.bpExtended) == (sb.extensions, sb.bpExtended)
        prod.free_query(qp)
        orc.free_query(qo)
    prod.close()
    orc.close()


@pytest.mark.parametrize("opts", [["--twins=0..50", "--nogapped", "--format=general-"], ["--twins=-5..30"],
                                  ["--twins=10..100", "--nogfextend", "--nogapped", "--format=general-"]], ids=lambda o: " ".join(o))
def test_cli_twins(synth, opts):
    ref = REF_CLI if os.path.exists(REF_CLI) else ORACLE_CLI
    for files in ([CAT, PIG], list(synth(300000))):
        same_output(run_cli(PRODUCT_CLI, files + opts)[0], run_cli(ref, files + opts)[0])


def test_twins_refuse_a_queue_that_would_forget(synth):
    """a queue smaller than the hit density over the span window: the reference's result then depends on what its queue
    has forgotten, and the library says so instead of returning something else"""
    t, q = synth(300000)
    with pytest.raises(RuntimeError, match="seed hit queue"):
        run_cli(PRODUCT_CLI, [t, q, "--twins=0..200", "--seed=match8", "--step=2", "--nogapped", "--seedqueue=2000"])


@pytest.mark.parametrize("depth", ["0.3", "keep:0.05", "keep:0.2", "keep,nowarn:0.15"])
def test_cli_querydepth(depth):
    """gapped_extend's maxPairedBases inside the in-order-commit scheduler: the commit that takes the sum over the limit
    closes every open anchor (the scheduler's own source runs this against the oracle on the block emulator with 8 and 16
    lanes, tests/warp_emu/test_gapped_sched.cpp cases 6 and 7)"""
    ref = REF_CLI if os.path.exists(REF_CLI) else ORACLE_CLI
    aglobin = os.path.join(GOLDEN, "aglobin.2bit")
    args = [aglobin + "/human", aglobin + "/cow", "--querydepth=" + depth, "--format=general-"]
    same_output(run_cli(PRODUCT_CLI, args)[0], run_cli(ref, args)[0])


@pytest.mark.parametrize("limit", ["=8", "=keep,nowarn:33", "+=13", "=nowarn:200"])
def test_cli_queryhsplimit(limit):
    """searchLimit in the product: every hit is still enumerated and the table is cut, on the host, after the query position
    that took the number of HSPs past the limit -- the same table the reference's early stop leaves (the cut rule runs
    against the oracle's real early stop on the block emulator: tests/warp_emu/test_seed_kernels.cpp, 'searchLimit' modes)"""
    ref = REF_CLI if os.path.exists(REF_CLI) else ORACLE_CLI
    aglobin = os.path.join(GOLDEN, "aglobin.2bit")
    for extra in ([], ["--nogapped"]):
        args = [aglobin + "/human", aglobin + "/cow", "--queryhsplimit" + limit, "--format=general-"] + extra
        same_output(run_cli(PRODUCT_CLI, args)[0], run_cli(ref, args)[0])
