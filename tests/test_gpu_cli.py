"""End to end through the lastz_b200 command line on the GPU: golden LAVs of the reference and,
where oracle/_ref exists, the unmodified reference binary on synthetic pairs."""
import os

import pytest

from conftest import (ALT_EXTEND_FIXTURE_CASES, ALT_EXTEND_SYNTH_CASES, GENERAL_CASES, GOLDEN, ORACLE_CLI, PRODUCT_CLI, REF_CLI, SELF_CASES,
                      lav_body, masked_query, run_cli, same_output, self_case_target)

pytestmark = pytest.mark.gpu

CAT = os.path.join(GOLDEN, "pseudocat.fa")
PIG = os.path.join(GOLDEN, "pseudopig.fa")


def norm(t):
    return [l.replace("../test_data/", "").replace(GOLDEN + "/", "") for l in lav_body(t)]


@pytest.mark.parametrize("golden,opts", [
    ("base_test.default.lav", []),
    ("base_test.hits.lav", ["W=8", "T=0", "--plus", "--nogfextend", "--nogapped"]),
    ("base_test.hsp.lav", ["C=3", "W=8", "T=0"]),
    ("base_test.seeded.lav", ["C=3", "--seed=111010011101"]),
    ("base_test.hwseeded.lav", ["C=3", "--seed=TTT0T0T0TTT00T0T"]),
    ("base_test.chained.lav", ["C=1", "W=8", "T=0"]),
    ("base_test.extended.lav", ["C=2", "W=8", "T=0"]),
])
def test_cli_reproduces_golden_lav(golden, opts):
    out, _ = run_cli(PRODUCT_CLI, [CAT, PIG] + opts)
    assert norm(out) == norm(open(os.path.join(GOLDEN, golden)).read())


AGLOBIN = os.path.join(GOLDEN, "aglobin.2bit")


def test_cli_subrange_golden():
    """base_test_subrange (Makefile:534): 2bit contigs with [a,b] and [a#len] subranges; contig ordinals in the s stanza."""
    out, _ = run_cli(PRODUCT_CLI, [AGLOBIN + "/human[10000,60000]", AGLOBIN + "/cow[15000#40000]"])
    norm = lambda t: [l.replace("../test_data/", "").replace(GOLDEN + "/", "") for l in lav_body(t)]
    assert norm(out) == norm(open(os.path.join(GOLDEN, "base_test.subrange.lav")).read())


def test_cli_anchors_golden():
    """base_test_anchors (Makefile:510): gapped stage from an external anchors file, MAF blocks out."""
    out, _ = run_cli(PRODUCT_CLI, [AGLOBIN + "/human", AGLOBIN + "/cow", "C=0", "--format=maf-",
                                  "--anchors=" + os.path.join(GOLDEN, "base_test.anchors.anchors")])
    assert out == open(os.path.join(GOLDEN, "base_test.anchors.maf")).read()


def test_cli_segments_round_trip(tmp_path):
    segs, _ = run_cli(PRODUCT_CLI, [CAT, PIG, "--nogapped", "--format=segments"])
    f = tmp_path / "hsps.segments"
    f.write_text(segs)
    out, _ = run_cli(PRODUCT_CLI, [CAT, PIG, f"--segments={f}"])
    assert norm(out) == norm(open(os.path.join(GOLDEN, "base_test.default.lav")).read())


@pytest.mark.parametrize("size,opts", [
    (1000000, []),
    (1000000, ["--nogapped", "--format=segments"]),
    (1000000, ["--allocate:traceback=8M"]),
    (500000, ["--seed=14of22", "--notransition", "--step=2", "--hspthresh=2200"]),
    (1000000, ["--chain"]),
    (1000000, ["--chain=40,30", "--nogapped"]),
])
def test_cli_matches_reference_on_synthetic(synth, size, opts):
    ref = REF_CLI if os.path.exists(REF_CLI) else ORACLE_CLI
    t, q = synth(size)
    got, _ = run_cli(PRODUCT_CLI, [t, q] + opts)
    want, _ = run_cli(ref, [t, q] + opts)
    if "--format=segments" in opts:
        assert got == want
    else:
        assert lav_body(got) == lav_body(want)


@pytest.mark.parametrize("which,opts", SELF_CASES)
def test_cli_self_alignment_matches_reference(tmp_path, which, opts):
    """BASELINE.json configs[1] (aglobin.2bit/human --self --seed=12of19 --nogapped) and the other --self shapes."""
    ref = REF_CLI if os.path.exists(REF_CLI) else ORACLE_CLI
    target = self_case_target(which, tmp_path)
    got, _ = run_cli(PRODUCT_CLI, [target] + opts)
    want, _ = run_cli(ref, [target] + opts)
    assert [l for l in got.splitlines() if "lastz.v" not in l] == [l for l in want.splitlines() if "lastz.v" not in l]


@pytest.mark.parametrize("opts", ALT_EXTEND_FIXTURE_CASES)
def test_cli_exact_and_mismatch_extension_on_fixtures(opts):
    ref = REF_CLI if os.path.exists(REF_CLI) else ORACLE_CLI
    same_output(run_cli(PRODUCT_CLI, [CAT, PIG] + opts)[0], run_cli(ref, [CAT, PIG] + opts)[0])


@pytest.mark.parametrize("opts", ALT_EXTEND_SYNTH_CASES)
def test_cli_exact_and_mismatch_extension_on_synthetic(synth, tmp_path, opts):
    ref = REF_CLI if os.path.exists(REF_CLI) else ORACLE_CLI
    t, qm = masked_query(synth, tmp_path)
    same_output(run_cli(PRODUCT_CLI, [t, qm] + opts)[0], run_cli(ref, [t, qm] + opts)[0])


@pytest.mark.parametrize("a1,a2,opts", GENERAL_CASES)
def test_cli_general_format(a1, a2, opts):
    ref = REF_CLI if os.path.exists(REF_CLI) else ORACLE_CLI
    assert run_cli(PRODUCT_CLI, [CAT + a1, PIG + a2] + opts)[0] == run_cli(ref, [CAT + a1, PIG + a2] + opts)[0]
