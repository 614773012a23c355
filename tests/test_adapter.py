"""The drop-in boundary, compiled: the reference's own lastz (its objects, built from /root/reference/src) with the five
hot-path entry points -- build_seed_position_table, free_position_table, seed_hit_search, reduce_to_points,
gapped_extend, reference signatures -- implemented by adapter/lastz_adapter.c over include/lastz_b200.h
(adapter/build.sh renames the originals inside the reference objects with objcopy, so lastz.o calls the adapter).

CPU: the adapter linked against the ORACLE library reproduces the reference's base-test goldens -- this checks the
binding itself (argument meaning, reporter replay, record layouts, ownership of the returned lists).
GPU: the same adapter source linked against liblastz_b200.so -- the reference running its hot path on the B200.
"""
import os
import subprocess

import pytest

from conftest import GOLDEN, REF_CLI, ROOT, lav_body, run_cli

ADAPTER_ORACLE = os.path.join(ROOT, "adapter", "_build", "lastz_adapter_oracle")
ADAPTER_B200 = os.path.join(ROOT, "adapter", "_build", "lastz_adapter_b200")
CAT = os.path.join(GOLDEN, "pseudocat.fa")
PIG = os.path.join(GOLDEN, "pseudopig.fa")
AGLOBIN = os.path.join(GOLDEN, "aglobin.2bit")

LAV_CASES = [
    ("base_test.default.lav", [CAT, PIG]),                                                  # src/Makefile:208
    ("base_test.hits.lav", [CAT, PIG, "W=8", "T=0", "--plus", "--nogfextend", "--nogapped"]),  # :295 (plain hits)
    ("base_test.hsp.lav", [CAT, PIG, "C=3", "W=8", "T=0"]),                                    # :306
    ("base_test.seeded.lav", [CAT, PIG, "C=3", "--seed=111010011101"]),
    ("base_test.chained.lav", [CAT, PIG, "C=1", "W=8", "T=0"]),                                # :351 (the reference's own chain.c in between)
    ("base_test.extended.lav", [CAT, PIG, "C=2", "W=8", "T=0"]),                               # :362
    ("base_test.subrange.lav", [AGLOBIN + "/human[10000,60000]", AGLOBIN + "/cow[15000#40000]"]),   # :534
]


def _build():
    if not os.path.exists(ADAPTER_ORACLE) and os.path.isdir("/root/reference/src"):
        subprocess.run([os.path.join(ROOT, "adapter", "build.sh")], check=True)
    if not os.path.exists(ADAPTER_ORACLE):
        pytest.skip("adapter/_build is not there and the reference sources are not present to build it")


def _norm(t):
    return [l.replace("../test_data/", "").replace(GOLDEN + "/", "") for l in lav_body(t)]


@pytest.mark.parametrize("golden,args", LAV_CASES)
def test_reference_with_adapter_reproduces_goldens(golden, args):
    _build()
    out, _ = run_cli(ADAPTER_ORACLE, args)
    assert _norm(out) == _norm(open(os.path.join(GOLDEN, golden)).read())


def test_adapter_anchors_file_to_maf():
    """base_test_anchors (src/Makefile:510): the gapped stage alone, from the reference's anchors reader"""
    _build()
    out, _ = run_cli(ADAPTER_ORACLE, [AGLOBIN + "/human", AGLOBIN + "/cow", "C=0", "--format=maf-",
                                      "--anchors=" + os.path.join(GOLDEN, "base_test.anchors.anchors")])
    assert out == open(os.path.join(GOLDEN, "base_test.anchors.maf")).read()


def test_adapter_recoverable_processor(synth):
    """--recoverseeds / --twins: the reference picks process_for_recoverable_hit / process_for_twin_hit, the adapter passes
    that on as recoverSeeds / the twin spans, and the reference's own merge_segments runs on what comes back"""
    _build()
    if not os.path.exists(REF_CLI):
        pytest.skip("oracle/_ref not built")
    t, q = synth(300000)
    for args in ([CAT, PIG, "--recoverseeds"], [t, q, "--recoverseeds", "--nogapped", "--format=general-"],
                 [t, q, "--twins=0..50", "--nogapped", "--format=general-"], [CAT, PIG, "--twins=-5..30"],
                 [AGLOBIN + "/human", AGLOBIN + "/cow", "--queryhsplimit=keep,nowarn:33", "--nogapped", "--format=general-"],
                 [AGLOBIN + "/human", AGLOBIN + "/cow", "--queryhsplimit=20"], [AGLOBIN + "/human", AGLOBIN + "/cow", "--querydepth=keep:0.1", "--format=general-"]):
        assert _norm(run_cli(ADAPTER_ORACLE, args)[0]) == _norm(run_cli(REF_CLI, args)[0])


def test_adapter_refuses_what_the_library_lacks():
    _build()
    for opts in (["--hspthresh=top10%"],):
        p = subprocess.run([ADAPTER_ORACLE, CAT, PIG] + opts, capture_output=True, text=True)
        assert p.returncode != 0 and "lastz_b200 adapter" in p.stderr, (opts, p.stderr[-300:])


@pytest.mark.gpu
@pytest.mark.parametrize("golden,args", LAV_CASES)
def test_reference_with_adapter_on_the_gpu(golden, args):
    if not os.path.exists(ADAPTER_B200):
        pytest.skip("adapter/_build/lastz_adapter_b200 was not built (needs /root/reference/src at build time)")
    out, _ = run_cli(ADAPTER_B200, args)
    assert _norm(out) == _norm(open(os.path.join(GOLDEN, golden)).read())
