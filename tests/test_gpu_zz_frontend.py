"""Front-end features through the product command line on the GPU.  The first two groups (GFA output, --anyornone) ran
on a B200 in round 1; everything after them was written when the round's GPU minutes were spent:
adaptive thresholds, the transition-variant order fixture, [multi] queries and targets, field-list / SAM / PAF / blastn /
rdotplot writers, filters, FASTQ input, ragged query files.  Those are the same host C sources the CPU suite compares
with the reference binary through the oracle library; here they run over the CUDA library.  The file sorts last on
purpose: pytest -x stops at the first failure, and these are the cases without a GPU run behind them yet."""
import os

import pytest

from conftest import ADAPTIVE_CASES, ANYORNONE_CASES, FASTQ_CASES, FIELD_CASES, FILTER_CASES, GFA_CASES, GOLDEN, MULTI_QUERY_CASES, MULTI_TARGET_CASES, adaptive_case_files, ORACLE_CLI, PRODUCT_CLI, REF_CLI, run_cli, same_output

pytestmark = pytest.mark.gpu

CAT = os.path.join(GOLDEN, "pseudocat.fa")
PIG = os.path.join(GOLDEN, "pseudopig.fa")


# the writers and filters are host code already compared with the reference on the CPU: every third case goes through the GPU
@pytest.mark.parametrize("a1,a2,opts", GFA_CASES + FIELD_CASES[::3])
def test_cli_gfa_format(a1, a2, opts):
    ref = REF_CLI if os.path.exists(REF_CLI) else ORACLE_CLI
    assert run_cli(PRODUCT_CLI, [CAT + a1, PIG + a2] + opts)[0] == run_cli(ref, [CAT + a1, PIG + a2] + opts)[0]


@pytest.mark.parametrize("opts", ANYORNONE_CASES)
def test_cli_anyornone(opts):
    ref = REF_CLI if os.path.exists(REF_CLI) else ORACLE_CLI
    same_output(run_cli(PRODUCT_CLI, [CAT, PIG] + opts)[0], run_cli(ref, [CAT, PIG] + opts)[0])


@pytest.mark.parametrize("which,opts", ADAPTIVE_CASES[::2])
def test_cli_adaptive_threshold(which, opts):
    """K=top<N>%: the library returns every extension (threshold far below any score, entropy off) and the front end
    replays the reference's coverage-limited heap over them in discovery order"""
    ref = REF_CLI if os.path.exists(REF_CLI) else ORACLE_CLI
    files = adaptive_case_files(which)
    same_output(run_cli(PRODUCT_CLI, files + opts)[0], run_cli(ref, files + opts)[0])


def test_cli_variant_order_fixture():
    files = [os.path.join(GOLDEN, "aglobin.2bit") + "/human", os.path.join(GOLDEN, "aglobin.2bit") + "/cow[20000..32000]"]
    out, _ = run_cli(PRODUCT_CLI, files + ["--nogfextend", "--nogapped", "--strand=plus", "--format=general-"])
    got = ["\t".join((l.split("\t")[4], l.split("\t")[9])) for l in out.splitlines()]
    assert got == open(os.path.join(GOLDEN, "aglobin_cow_20k_32k.plus_hits.order.tsv")).read().splitlines()


@pytest.mark.parametrize("opts", MULTI_QUERY_CASES)
def test_cli_multi_query(opts):
    """[multi] query: seeds and x-drop scans stop at the NUL separators by themselves (no seed word spans one, their
    substitution score ends any scan); the gapped stage cuts each sweep at the separators around its anchor"""
    ref = REF_CLI if os.path.exists(REF_CLI) else ORACLE_CLI
    args = [CAT, PIG + "[multi]"] + opts
    same_output(run_cli(PRODUCT_CLI, args)[0], run_cli(ref, args)[0])


def test_cli_multi_subrange_golden():
    """base_test_multi_subrange (Makefile:582) through the product"""
    out, _ = run_cli(PRODUCT_CLI, [os.path.join(GOLDEN, "aglobin.2bit") + "/human", os.path.join(GOLDEN, "shorties.2bit") + "[multi,51..200]", "K=3000", "--maf-"])
    assert out == open(os.path.join(GOLDEN, "base_test.multi_subrange.maf")).read()


@pytest.mark.parametrize("target,query,opts", MULTI_TARGET_CASES)
def test_cli_multi_target(target, query, opts):
    ref = REF_CLI if os.path.exists(REF_CLI) else ORACLE_CLI
    args = [os.path.join(GOLDEN, target), os.path.join(GOLDEN, query)] + opts
    same_output(run_cli(PRODUCT_CLI, args)[0], run_cli(ref, args)[0])


@pytest.mark.parametrize("opts", FILTER_CASES[::3])
def test_cli_filters(opts):
    ref = REF_CLI if os.path.exists(REF_CLI) else ORACLE_CLI
    files = adaptive_case_files("aglobin")
    same_output(run_cli(PRODUCT_CLI, files + opts)[0], run_cli(ref, files + opts)[0])


@pytest.mark.parametrize("query,opts", FASTQ_CASES[1::3])
def test_cli_fastq_query(query, opts):
    ref = REF_CLI if os.path.exists(REF_CLI) else ORACLE_CLI
    args = [os.path.join(GOLDEN, "aglobin.2bit") + "/human", os.path.join(GOLDEN, query)] + opts
    same_output(run_cli(PRODUCT_CLI, args)[0], run_cli(ref, args)[0])


@pytest.mark.parametrize("qact,opts", [("", ["--format=general-"]), ("[multi]", ["--format=maf-"])])
def test_cli_ragged_query_file(qact, opts):
    """a query shorter than the seed, an empty record, N only, lower case only, DOS line ends"""
    ref = REF_CLI if os.path.exists(REF_CLI) else ORACLE_CLI
    args = [os.path.join(GOLDEN, "edge_target.fa"), os.path.join(GOLDEN, "edge_queries.fa") + qact] + opts
    same_output(run_cli(PRODUCT_CLI, args)[0], run_cli(ref, args)[0])
