"""The oracle (CPU restatement + host front end) against the reference's own golden files.

These are the reference's `make test` / `make base_tests` cases that pin the hot path
(src/Makefile:208-217, :295, :306, :329, :384, :465); fixtures copied verbatim from test_data/.
"""
import os
import subprocess

import pytest

from conftest import (write_inverted_repeat_fasta, ADAPTIVE_CASES, ALT_EXTEND_FIXTURE_CASES, ALT_EXTEND_SYNTH_CASES, ANYORNONE_CASES, FASTQ_CASES, FIELD_CASES, FILTER_CASES, MULTI_QUERY_CASES, MULTI_TARGET_CASES, adaptive_case_files, GENERAL_CASES, GFA_CASES, GOLDEN, ORACLE_CLI, REF_CLI, SELF_CASES, lav_body,
                      masked_query, run_cli, same_output, self_case_target)

CAT = os.path.join(GOLDEN, "pseudocat.fa")
PIG = os.path.join(GOLDEN, "pseudopig.fa")

CASES = [
    ("base_test.default.lav", []),                                                  # Makefile:208
    ("base_test.hits.lav", ["W=8", "T=0", "--plus", "--nogfextend", "--nogapped"]),  # Makefile:295
    ("base_test.hsp.lav", ["C=3", "W=8", "T=0"]),                                    # Makefile:306
    ("base_test.seeded.lav", ["C=3", "--seed=111010011101"]),
    ("base_test.hwseeded.lav", ["C=3", "--seed=TTT0T0T0TTT00T0T"]),                        # Makefile:465
    ("base_test.chained.lav", ["C=1", "W=8", "T=0"]),                                # Makefile:351 (chain only)
    ("base_test.extended.lav", ["C=2", "W=8", "T=0"]),                               # Makefile:362 (chain + gapped)
]


@pytest.mark.parametrize("golden,opts", CASES)
def test_oracle_reproduces_golden_lav(golden, opts):
    out, _ = run_cli(ORACLE_CLI, [CAT, PIG] + opts)
    want = open(os.path.join(GOLDEN, golden)).read()
    # paths differ (../test_data/...), the comparator strips them the same way lav_compare.py does
    norm = lambda t: [l.replace("../test_data/", "").replace(GOLDEN + "/", "") for l in lav_body(t)]
    assert norm(out) == norm(want)


AGLOBIN = os.path.join(GOLDEN, "aglobin.2bit")


def test_oracle_subrange_golden():
    """base_test_subrange (Makefile:534): 2bit contigs with [a,b] and [a#len] subranges; contig ordinals in the s stanza."""
    out, _ = run_cli(ORACLE_CLI, [AGLOBIN + "/human[10000,60000]", AGLOBIN + "/cow[15000#40000]"])
    norm = lambda t: [l.replace("../test_data/", "").replace(GOLDEN + "/", "") for l in lav_body(t)]
    assert norm(out) == norm(open(os.path.join(GOLDEN, "base_test.subrange.lav")).read())


def test_oracle_anchors_golden():
    """base_test_anchors (Makefile:510): gapped stage from an external anchors file, MAF blocks out."""
    out, _ = run_cli(ORACLE_CLI, [AGLOBIN + "/human", AGLOBIN + "/cow", "C=0", "--format=maf-",
                                  "--anchors=" + os.path.join(GOLDEN, "base_test.anchors.anchors")])
    assert out == open(os.path.join(GOLDEN, "base_test.anchors.maf")).read()


def test_oracle_2bit_query_golden():
    """base_test_2bit2 (Makefile:441): a 2bit file as the query, same alignments as the FASTA run."""
    out, _ = run_cli(ORACLE_CLI, [CAT, os.path.join(GOLDEN, "pseudopig.2bit"), "C=2", "W=8", "T=0"])
    out = out.replace("pig", "> pig").replace("do> pig.2bit", "dopig.fa")          # the Makefile's two seds
    norm = lambda t: [l.replace("../test_data/", "").replace(GOLDEN + "/", "") for l in lav_body(t)]
    assert norm(out) == norm(open(os.path.join(GOLDEN, "base_test.extended.lav")).read())


def test_oracle_query_from_stdin_golden():
    """base_test_stdin2 (Makefile:402): no query file => the query is read from stdin."""
    import subprocess
    p = subprocess.run([ORACLE_CLI, CAT, "C=3", "W=8", "T=0"], stdin=open(PIG, "rb"), capture_output=True)
    assert p.returncode == 0, p.stderr[-2000:]
    out = p.stdout.decode().replace("(stdin)", "../test_data/pseudopig.fa")
    norm = lambda t: [l.replace("../test_data/", "").replace(GOLDEN + "/", "") for l in lav_body(t)]
    assert norm(out) == norm(open(os.path.join(GOLDEN, "base_test.hsp.lav")).read())


def test_oracle_2bit_target_golden():
    """base_test_2bit1 (Makefile:427): one contig of a 2bit file as the TARGET."""
    out, _ = run_cli(ORACLE_CLI, [os.path.join(GOLDEN, "pseudopig.2bit") + "/pig2", CAT, "C=2", "W=8", "T=0"])
    import re
    out = out.replace("pig", "> pig").replace("do> pig.2bit", "dopig2.fa")
    out = re.sub(r"(dopig2.*) 0 2", r"\1 0 1", out)                                # the Makefile's three seds
    norm = lambda t: [l.replace("../test_data/", "").replace(GOLDEN + "/", "") for l in lav_body(t)]
    assert norm(out) == norm(open(os.path.join(GOLDEN, "base_test.pig_cat.lav")).read())


@pytest.mark.parametrize("query", ["shorties.fa", "shorties.2bit"])
def test_oracle_contigs_of_interest_golden(query):
    """base_test_coi (Makefile:553): [subset=<names file>] on a multi-sequence FASTA / 2bit query, --maf- out."""
    names = os.path.join(GOLDEN, "shorties.names")
    out, _ = run_cli(ORACLE_CLI, [AGLOBIN + "/human", os.path.join(GOLDEN, query) + f"[subset={names}]", "K=3000", "--maf-"])
    assert out == open(os.path.join(GOLDEN, "base_test.coi.maf")).read()


def test_oracle_axt_golden():
    """base_test_axt (Makefile:340); comment lines are not compared (tools/axt_compare.py:132)."""
    out, _ = run_cli(ORACLE_CLI, [CAT, PIG, "--format=axt"])
    body = lambda t: [l for l in t.splitlines() if not l.startswith("#")]
    assert body(out) == body(open(os.path.join(GOLDEN, "base_test.default.axt")).read())


def test_oracle_nib_target_golden():
    """base_test_nib1 (Makefile:414): a .nib file as the target (header = file:start-end)."""
    import re
    out, _ = run_cli(ORACLE_CLI, [os.path.join(GOLDEN, "pseudopig2.nib"), CAT, "C=2", "W=8", "T=0"])
    out = re.sub(r'"[^"]*\.nib:[^"]*"', '"> pig2"', out).replace(".nib", ".fa")          # the Makefile's two seds
    norm = lambda t: [l.replace("../test_data/", "").replace(GOLDEN + "/", "") for l in lav_body(t)]
    assert norm(out) == norm(open(os.path.join(GOLDEN, "base_test.pig_cat.lav")).read())


def test_oracle_mask_golden():
    """base_test_mask (Makefile:543): [nmask=<file>] intervals become N, N scored as ambiguous (--ambiguous=n,60)."""
    mask = os.path.join(GOLDEN, "pseudopig.n.mask")
    out, _ = run_cli(ORACLE_CLI, [CAT, PIG + f"[nmask={mask}]", "--ambiguous=n,60"])
    norm = lambda t: [l.replace("../test_data/", "").replace(GOLDEN + "/", "") for l in lav_body(t)]
    assert norm(out) == norm(open(os.path.join(GOLDEN, "base_test.mask.lav")).read())


def test_oracle_overweight_seed_hits_golden():
    """base_test_ow_seeded (Makefile:487): --justhits with a 16-bit seed and --word=12.  The reference resolves
    the overweight seed through a 12-bit table (seed_search.c:878), which changes the ORDER hits are found in, not
    the set; its own test compares order-insensitively (gfa_compare.py --sort) and so does this one."""
    out, _ = run_cli(ORACLE_CLI, [CAT, PIG, "--justhits", "--seed=111010011101", "--word=12", "--gfa"])
    hits = lambda t: sorted(l for l in t.splitlines() if l.startswith("a "))
    assert hits(out) == hits(open(os.path.join(GOLDEN, "base_test.owseeded.gfa")).read())


def test_oracle_anchors_multi_golden():
    """base_test_anchors_multi (Makefile:522): one anchors file for several query sequences (matched by name)."""
    names = os.path.join(GOLDEN, "shorties.names")
    out, _ = run_cli(ORACLE_CLI, [AGLOBIN + "/human", os.path.join(GOLDEN, "shorties.fa") + f"[subset={names}]", "C=0", "--format=maf-",
                                  "--anchors=" + os.path.join(GOLDEN, "base_test.anchors_multi.anchors")])
    assert out == open(os.path.join(GOLDEN, "base_test.anchors_multi.maf")).read()


def _maf_blocks_by_pos1(text):
    """MAF blocks ordered by (start, strand) in sequence 1, then (start, strand) in sequence 2, then lengths and
    score: the order the reference's base_test_multi puts both files in before diffing them (Makefile:571)."""
    def key(b):
        s1, s2 = [l.split() for l in b.splitlines() if l.startswith("s ")][:2]
        score = int(b.split("score=")[1].split()[0])
        return (int(s1[2]), s1[4], int(s2[2]), s2[4], int(s1[3]), int(s2[3]), score, s1[1], s2[1])
    blocks = [b for b in text.split("\n\n") if b.strip()]
    blocks.sort(key=key)
    return "\n\n".join(blocks) + "\n\n"


def test_oracle_multi_golden():
    """base_test_multi (Makefile:571): the query is ALL the named sequences of a 2bit file as one partitioned
    sequence ([multi], sequences.h:188-191); the reference's recipe sorts the blocks by position in the target."""
    names = os.path.join(GOLDEN, "shorties.names")
    out, _ = run_cli(ORACLE_CLI, [AGLOBIN + "/human", os.path.join(GOLDEN, "shorties.2bit") + f"[multi,@{names}]", "K=3000", "--maf-"])
    assert _maf_blocks_by_pos1(out) == open(os.path.join(GOLDEN, "base_test.multi.maf")).read()


def test_oracle_multi_subrange_golden():
    """base_test_multi_subrange (Makefile:582): [multi] with a subrange applied to every sequence."""
    out, _ = run_cli(ORACLE_CLI, [AGLOBIN + "/human", os.path.join(GOLDEN, "shorties.2bit") + "[multi,51..200]", "K=3000", "--maf-"])
    assert out == open(os.path.join(GOLDEN, "base_test.multi_subrange.maf")).read()


@pytest.mark.parametrize("opts", MULTI_QUERY_CASES)
def test_oracle_multi_query_matches_reference(opts):
    if not os.path.exists(REF_CLI):
        pytest.skip("oracle/_ref not built")
    args = [CAT, PIG + "[multi]"] + opts
    same_output(run_cli(ORACLE_CLI, args)[0], run_cli(REF_CLI, args)[0])


@pytest.mark.parametrize("target,query,opts", MULTI_TARGET_CASES)
def test_oracle_multi_target_matches_reference(target, query, opts):
    if not os.path.exists(REF_CLI):
        pytest.skip("oracle/_ref not built")
    args = [os.path.join(GOLDEN, target), os.path.join(GOLDEN, query)] + opts
    same_output(run_cli(ORACLE_CLI, args)[0], run_cli(REF_CLI, args)[0])


@pytest.mark.parametrize("opts", FILTER_CASES)
def test_oracle_filters_match_reference(opts):
    if not os.path.exists(REF_CLI):
        pytest.skip("oracle/_ref not built")
    files = adaptive_case_files("aglobin")
    same_output(run_cli(ORACLE_CLI, files + opts)[0], run_cli(REF_CLI, files + opts)[0])


@pytest.mark.parametrize("opts,score", [(["--format=axt"], 0), (["--format=maf-", "--nogapped"], 1)])
def test_oracle_rdotplot_side_file(tmp_path, opts, score):
    """--rdotplot[+score]=<file>: the dot plot written next to the main output (output.c:713, :928)"""
    if not os.path.exists(REF_CLI):
        pytest.skip("oracle/_ref not built")
    flag = "--rdotplot+score=" if score else "--rdotplot="
    got = run_cli(ORACLE_CLI, [CAT, PIG] + opts + [flag + str(tmp_path / "o.dots")])[0]
    want = run_cli(REF_CLI, [CAT, PIG] + opts + [flag + str(tmp_path / "r.dots")])[0]
    assert open(tmp_path / "o.dots").read() == open(tmp_path / "r.dots").read()
    assert [l for l in got.splitlines() if "lastz.v" not in l] == [l for l in want.splitlines() if "lastz.v" not in l]


@pytest.mark.parametrize("query,opts", FASTQ_CASES)
def test_oracle_fastq_query_matches_reference(query, opts):
    if not os.path.exists(REF_CLI):
        pytest.skip("oracle/_ref not built")
    args = [AGLOBIN + "/human", os.path.join(GOLDEN, query)] + opts
    same_output(run_cli(ORACLE_CLI, args)[0], run_cli(REF_CLI, args)[0])


@pytest.mark.parametrize("tact,qact", [("", ""), (",nameparse=darkspace", "[nameparse=darkspace]"), (",nameparse=alphanum", "[nameparse=alphanum]"),
                                       (",nameparse=full", "[fullname]"), ("", "[nickname=bob]")])
def test_oracle_sequence_names(tact, qact):
    """shorten_header (sequences.c:5913): names cut at blanks, '|' and ':' by default, file suffixes dropped, the
    "reverse complement of" / "positions x of" prefixes skipped; [nameparse=...] and [nickname=...] (tests/golden/names.fa
    is a hand-made file with such headers)"""
    if not os.path.exists(REF_CLI):
        pytest.skip("oracle/_ref not built")
    f = os.path.join(GOLDEN, "names.fa")
    args = [f + "[multi" + tact + "]", f + qact, "--format=general-:name1,name2,start1,end1", "K=2000", "--nogapped"]
    same_output(run_cli(ORACLE_CLI, args)[0], run_cli(REF_CLI, args)[0])


@pytest.mark.parametrize("qact,opts", [("", ["--format=general-"]), ("", []), ("[multi]", ["--format=maf-"]), ("[unmask]", ["--format=sam"]),
                                       ("", ["--format=blastn"])])
def test_oracle_ragged_query_file(qact, opts):
    """tests/golden/edge_queries.fa (hand-made): a normal record, one shorter than the seed, an EMPTY record (warned about
    and skipped, sequences.c:2429), one of N only, one in lower case, one with DOS line ends; stdout and stderr"""
    if not os.path.exists(REF_CLI):
        pytest.skip("oracle/_ref not built")
    args = [os.path.join(GOLDEN, "edge_target.fa"), os.path.join(GOLDEN, "edge_queries.fa") + qact] + opts
    got, want = run_cli(ORACLE_CLI, args), run_cli(REF_CLI, args)
    assert got[0] == want[0] and got[1] == want[1]


def test_multi_target_refusals():
    """what a partitioned target does not do yet stops with a FAILURE instead of giving other results than the reference"""
    import subprocess
    t = os.path.join(GOLDEN, "aglobin.2bit[multi]")
    for q, opts in [("aglobin.2bit[multi]", []), ("shorties.fa", ["--notrivial"]), ("shorties.fa", ["--format=lav"]),
                    ("shorties.fa", ["K=top20%"])]:
        p = subprocess.run([ORACLE_CLI, t, os.path.join(GOLDEN, q)] + opts, capture_output=True, text=True)
        assert p.returncode != 0 and "FAILURE" in p.stderr, (q, opts)


def test_oracle_segments_round_trip(tmp_path):
    """base_test_segments (Makefile:384): HSPs written, re-read as anchors, gapped stage alone."""
    segs, _ = run_cli(ORACLE_CLI, [CAT, PIG, "--nogapped", "--format=segments"])
    f = tmp_path / "hsps.segments"
    f.write_text(segs)
    out, _ = run_cli(ORACLE_CLI, [CAT, PIG, f"--segments={f}"])
    want = open(os.path.join(GOLDEN, "base_test.default.lav")).read()
    norm = lambda t: [l.replace("../test_data/", "").replace(GOLDEN + "/", "") for l in lav_body(t)]
    assert norm(out) == norm(want)


@pytest.mark.skipif(not os.path.exists(REF_CLI), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("size,opts", [
    (200000, []),
    (200000, ["--nogapped", "--format=segments"]),
    (300000, ["--allocate:traceback=2M"]),       # forces truncation + many bounded alignments
    (200000, ["--seed=14of22", "--notransition", "--step=3"]),
    (200000, ["--transition=2", "--hspthresh=2500", "--noentropy"]),
    (200000, ["--ydrop=4000", "--gappedthresh=5000", "--xdrop=400"]),
    (300000, ["--chain", "--nogapped"]),                     # chain.c: K-d tree order decides ties
    (300000, ["--chain=40,30", "--nogapped"]),               # diagonal / anti-diagonal penalties (lastz.c:3687)
    (300000, ["--chain"]),                                   # config 4: chained + gapped
])
def test_oracle_matches_reference_on_synthetic(synth, size, opts):
    t, q = synth(size)
    got, _ = run_cli(ORACLE_CLI, [t, q] + opts)
    want, _ = run_cli(REF_CLI, [t, q] + opts)
    if "--format=segments" in opts:
        assert got == want
    else:
        assert lav_body(got) == lav_body(want)


@pytest.mark.parametrize("which,opts", SELF_CASES)
def test_oracle_self_alignment_matches_reference(tmp_path, which, opts):
    """--self: below-diagonal hits dropped (seed_search.c:2182), HSPs / alignments mirrored (lastz.c:3858, :4229)."""
    if not os.path.exists(REF_CLI):
        pytest.skip("oracle/_ref not built")
    target = self_case_target(which, tmp_path)
    got, _ = run_cli(ORACLE_CLI, [target] + opts)
    want, _ = run_cli(REF_CLI, [target] + opts)
    assert [l for l in got.splitlines() if "lastz.v" not in l] == [l for l in want.splitlines() if "lastz.v" not in l]


@pytest.mark.parametrize("opts", ALT_EXTEND_FIXTURE_CASES)
def test_oracle_exact_and_mismatch_extension_on_fixtures(opts):
    if not os.path.exists(REF_CLI):
        pytest.skip("oracle/_ref not built")
    files = [os.path.join(GOLDEN, "pseudocat.fa"), os.path.join(GOLDEN, "pseudopig.fa")]
    same_output(run_cli(ORACLE_CLI, files + opts)[0], run_cli(REF_CLI, files + opts)[0])


@pytest.mark.parametrize("opts", ALT_EXTEND_SYNTH_CASES)
def test_oracle_exact_and_mismatch_extension_on_synthetic(synth, tmp_path, opts):
    if not os.path.exists(REF_CLI):
        pytest.skip("oracle/_ref not built")
    t, qm = masked_query(synth, tmp_path)
    same_output(run_cli(ORACLE_CLI, [t, qm] + opts)[0], run_cli(REF_CLI, [t, qm] + opts)[0])


@pytest.mark.parametrize("a1,a2,opts", GENERAL_CASES + GFA_CASES + FIELD_CASES)
def test_oracle_general_format(a1, a2, opts):
    if not os.path.exists(REF_CLI):
        pytest.skip("oracle/_ref not built")
    files = [os.path.join(GOLDEN, "pseudocat.fa") + a1, os.path.join(GOLDEN, "pseudopig.fa") + a2]
    assert run_cli(ORACLE_CLI, files + opts)[0] == run_cli(REF_CLI, files + opts)[0]


SCORES = """# a scoring file in the reference's grammar (dna_utilities.c:581-628)
gap_open_penalty   = 300
gap_extend_penalty = 25
hsp_threshold      = 2200
x_drop             = 600
y_drop             = 7000
     A     C     G     T
A   91   -90   -25  -100
C  -90   100  -100   -25
G  -25  -100   100   -90
T -100   -25   -90    91
"""


@pytest.mark.parametrize("opts", [["--scores={f}"], ["--scores={f}", "--nogapped"], ["Q={f}", "K=2600", "--chain"]])
def test_oracle_scoring_file(synth, tmp_path, opts):
    """--scores=<file>: substitution matrix, gap penalties and the thresholds embedded in the file."""
    if not os.path.exists(REF_CLI):
        pytest.skip("oracle/_ref not built")
    f = tmp_path / "my.scores"
    f.write_text(SCORES)
    t, qm = masked_query(synth, tmp_path)
    opts = [o.format(f=f) for o in opts]
    same_output(run_cli(ORACLE_CLI, [t, qm] + opts)[0], run_cli(REF_CLI, [t, qm] + opts)[0])


@pytest.mark.parametrize("opts", [[], ["--nogapped"]])
def test_oracle_matches_lastz_32_with_its_diag_hash(synth, tmp_path, opts):
    """lastz_32 (src/Makefile:59: 4M-entry diagonal hash) gives other HSPs than lastz; --diaghash=22 reproduces them."""
    ref32 = os.path.join(os.path.dirname(REF_CLI), "lastz_32")
    if not os.path.exists(ref32):
        pytest.skip("oracle/_ref/lastz_32 not built")
    t, qm = masked_query(synth, tmp_path)
    strip = lambda x: [l for l in x.splitlines() if "lastz" not in l]
    assert strip(run_cli(ORACLE_CLI, [t, qm, "--diaghash=22"] + opts)[0]) == strip(run_cli(ref32, [t, qm] + opts)[0])


@pytest.mark.parametrize("opts", ANYORNONE_CASES)
def test_oracle_anyornone(opts):
    if not os.path.exists(REF_CLI):
        pytest.skip("oracle/_ref not built")
    same_output(run_cli(ORACLE_CLI, [CAT, PIG] + opts)[0], run_cli(REF_CLI, [CAT, PIG] + opts)[0])
    multi = [AGLOBIN + "/human", os.path.join(GOLDEN, "shorties.fa")] + opts      # 20 query sequences
    same_output(run_cli(ORACLE_CLI, multi)[0], run_cli(REF_CLI, multi)[0])


def test_oracle_adaptive_k_golden():
    """base_test_adaptive_k (Makefile:317-327): K=top50% HSPs in GFA, compared like the reference's own test: the 'a'
    lines, sorted."""
    out, _ = run_cli(ORACLE_CLI, adaptive_case_files("aglobin") + ["C=3", "W=8", "T=0", "--noentropy", "K=top50%", "--gfa"])
    got = sorted(l for l in out.splitlines() if l.startswith("a"))
    want = sorted(open(os.path.join(GOLDEN, "base_test.adaptive_k.gfa")).read().splitlines())
    assert got == want


@pytest.mark.parametrize("which,opts", ADAPTIVE_CASES)
def test_oracle_adaptive_threshold_matches_reference(which, opts):
    """byte for byte, so the table's heap order (what --nogapped prints) is the reference's too"""
    if not os.path.exists(REF_CLI):
        pytest.skip("oracle/_ref not built")
    files = adaptive_case_files(which)
    same_output(run_cli(ORACLE_CLI, files + opts)[0], run_cli(REF_CLI, files + opts)[0])


def test_oracle_variant_order_fixture():
    """Seed hits at one query position are discovered exact word first, then the transition variants in the order of
    seed->transFlips: rightmost seed position first (seeds.c:165,603-613).  pseudocat/pseudopig never has two variants
    hitting at one position; this stretch of aglobin does (fixture: tests/golden/make_ref_fixtures.sh)."""
    files = [os.path.join(GOLDEN, "aglobin.2bit") + "/human", os.path.join(GOLDEN, "aglobin.2bit") + "/cow[20000..32000]"]
    out, _ = run_cli(ORACLE_CLI, files + ["--nogfextend", "--nogapped", "--strand=plus", "--format=general-"])
    got = ["\t".join((l.split("\t")[4], l.split("\t")[9])) for l in out.splitlines()]
    assert got == open(os.path.join(GOLDEN, "aglobin_cow_20k_32k.plus_hits.order.tsv")).read().splitlines()


@pytest.mark.parametrize("fmt", ["--format=sam", "--format=softsam", "--format=sam-"])
def test_oracle_sam_header_without_alignments(tmp_path, fmt):
    """@SQ lines appear with the first record (sam.c:213-250), so a run that finds nothing prints @HD alone."""
    if not os.path.exists(REF_CLI):
        pytest.skip("oracle/_ref not built")
    a, b = tmp_path / "a.fa", tmp_path / "b.fa"
    a.write_text(">a\n" + "ACGTTGCA" * 40 + "\n")
    b.write_text(">b\n" + "AAAAAAAAAAAAAAAACCCCCCCCCCCCCCCC" * 10 + "\n")
    for files in ([str(a), str(b)], [CAT, PIG]):
        same_output(run_cli(ORACLE_CLI, files + [fmt])[0], run_cli(REF_CLI, files + [fmt])[0])


RECOVER_CASES = [
    ["--recoverseeds"],
    ["--recoverseeds", "--nogapped", "--format=general-"],
    ["--recoverseeds", "--nogfextend", "--nogapped", "--format=general-"],
    ["--recoverseeds", "--chain", "--format=maf-"],
    ["--recoverhits", "--step=3", "--xdrop=600"],
    ["--recoverseeds", "--nogfextend", "--format=general-", "--gappedthresh=20000", "--step=20"],
    ["--recoverseeds", "--seed=match12", "--format=axt", "--strand=plus"],
    ["--recoverseeds", "--nogapped", "--format=general-", "--seed=match10", "--step=5", "--hspthresh=2000"],
    ["--recoverseeds", "--norecoverseeds", "--nogapped", "--format=segments"],
]


@pytest.mark.parametrize("opts", RECOVER_CASES, ids=lambda o: " ".join(o))
def test_oracle_recoverseeds(synth, opts):
    """process_for_recoverable_hit (seed_search.c:1221) + merge_segments (segment.c:1527): a hit on another diagonal of
    the same hash bucket is extended instead of lost, overlapping HSPs are merged per diagonal."""
    if not os.path.exists(REF_CLI):
        pytest.skip("oracle/_ref not built")
    pairs = [[CAT, PIG]] + ([list(synth(300000))] if opts in RECOVER_CASES[:3] else [])   # (the other gapped runs take 10 s to 2 min each at 300 kbp)
    for files in pairs:
        same_output(run_cli(ORACLE_CLI, files + opts)[0], run_cli(REF_CLI, files + opts)[0])


def test_oracle_recoverseeds_changes_the_table(synth):
    """the 2^16-bucket hash does lose hits on this pair, so the cases above are not vacuous"""
    t, q = synth(300000)
    base = run_cli(ORACLE_CLI, [t, q, "--nogapped", "--format=segments"])[0]
    rec = run_cli(ORACLE_CLI, [t, q, "--nogapped", "--format=segments", "--recoverseeds"])[0]
    assert base != rec


def test_oracle_anchor_file_is_merged(synth, tmp_path):
    """anchors read from a file are merged per diagonal before use (lastz.c:757, :3296): overlapping, contained and
    adjoining segments in shuffled order"""
    if not os.path.exists(REF_CLI):
        pytest.skip("oracle/_ref not built")
    import random
    t, q = synth(300000)
    rows = [l for l in run_cli(REF_CLI, [t, q, "--nogapped", "--format=segments", "--hspthresh=2500"])[0].splitlines() if not l.startswith("#")]
    rnd, out = random.Random(5), []
    for l in rows[:400]:
        out.append(l)
        f = l.split()
        a1, b1, a2, b2, sc = int(f[1]), int(f[2]), int(f[4]), int(f[5]), int(f[7])
        if f[6] != "+":
            continue
        r = rnd.random()
        if r < 0.3:
            out.append(f"{f[0]}\t{a1 + 5}\t{b1 + 9}\t{f[3]}\t{a2 + 5}\t{b2 + 9}\t+\t{sc + 7}")
        elif r < 0.5 and b1 - a1 > 20:
            out.append(f"{f[0]}\t{a1 + 3}\t{b1 - 3}\t{f[3]}\t{a2 + 3}\t{b2 - 3}\t+\t{sc - 5}")
        elif r < 0.6:
            out.append(f"{f[0]}\t{b1 + 1}\t{b1 + 30}\t{f[3]}\t{b2 + 1}\t{b2 + 30}\t+\t{sc}")
    rnd.shuffle(out)
    seg = tmp_path / "anchors.seg"
    seg.write_text("\n".join(out) + "\n")
    for opts in (["--nogapped", "--format=general-"], ["--format=lav"]):
        args = [t, q, f"--segments={seg}"] + opts
        same_output(run_cli(ORACLE_CLI, args)[0], run_cli(REF_CLI, args)[0])


TWIN_CASES = [
    ["--twins=0..50", "--nogapped", "--format=general-"],
    ["--twins=-5..30"],
    ["--twins=10..100", "--nogfextend", "--nogapped", "--format=general-"],
    ["--twins=60", "--seed=match9", "--nogapped", "--format=segments", "--hspthresh=2500"],
    ["--twins=0:40", "--chain", "--format=maf-"],
    ["--twins=0..200", "--seed=match8", "--step=2", "--nogapped", "--format=general-", "--seedqueue=2000"],   # a queue small enough to forget live entries
    ["--twins=0..50", "--notwins", "--nogapped", "--format=segments"],
]


@pytest.mark.parametrize("opts", TWIN_CASES, ids=lambda o: " ".join(o))
def test_oracle_twins(synth, opts):
    """process_for_twin_hit (seed_search.c:1814, seed-hit-queue version) + merge_segments: a hit is extended only when an
    earlier hit of its diagonal lies within the span window; the oracle keeps the reference's global queue, capacity
    included"""
    if not os.path.exists(REF_CLI):
        pytest.skip("oracle/_ref not built")
    pairs = [[CAT, PIG]] + ([list(synth(300000))] if "--chain" not in opts and opts != TWIN_CASES[1] else [])
    for files in pairs:
        same_output(run_cli(ORACLE_CLI, files + opts)[0], run_cli(REF_CLI, files + opts)[0])


@pytest.mark.parametrize("depth", ["0.05", "0.3", "keep:0.05", "keep:0.1", "keep:0.2", "keep:0.3", "keep,nowarn:0.15", "nowarn:0.2", "discard:0.1", "7"])
def test_oracle_querydepth(depth):
    """gapped_extend's maxPairedBases (gapped_extend.c:1444-1459; --querydepth=<d> = d x query length): once the alignments
    kept so far pair more bases than that, nothing further is extended; everything found is discarded unless `keep`"""
    if not os.path.exists(REF_CLI):
        pytest.skip("oracle/_ref not built")
    for fmt in ("--format=general-", "--format=maf-"):
        args = [AGLOBIN + "/human", AGLOBIN + "/cow", "--querydepth=" + depth, fmt]
        same_output(run_cli(ORACLE_CLI, args)[0], run_cli(REF_CLI, args)[0])


HSPLIMIT_SPELLINGS = ["=3", "=8", "=60", "=keep,nowarn:2", "=keep,nowarn:5", "=keep,nowarn:33", "+=4", "+=13", "+=nowarn:1", "+=warn:40", "=nowarn:15", "=warn:52", "=1K"]


@pytest.mark.parametrize("limit", HSPLIMIT_SPELLINGS)
def test_oracle_queryhsplimit(limit, tmp_path):
    """seed_hit_search's searchLimit (seed_search.c:551) and what lastz.c builds on it: both strands searched first, the second
    strand's limit reduced by what the first found (:3067), a query over the limit dropped as a whole unless `keep`
    (:3139), printing capped at the limit per query (output.c:556, :744).  Pairs with HSPs on one strand and on both."""
    if not os.path.exists(REF_CLI):
        pytest.skip("oracle/_ref not built")
    inv = str(tmp_path / "inv.fa")
    write_inverted_repeat_fasta(inv)
    for files in ([AGLOBIN + "/human", AGLOBIN + "/cow"], [inv, inv]):
        for extra in ([], ["--nogapped"], ["--nogapped", "--strand=minus"], ["--chain", "--format=maf-"]):
            args = files + ["--queryhsplimit" + limit, "--format=general-"] + extra
            same_output(run_cli(ORACLE_CLI, args)[0], run_cli(REF_CLI, args)[0])


@pytest.mark.parametrize("opts", [["--queryhsplimit+=30", "--format=rdotplot"], ["--queryhsplimit+=7", "--format=rdotplot+score"], ["--queryhsplimit+=7", "--format=rdotplot", "--nogapped"],
                                  ["--twins=0..40", "--queryhsplimit=keep,nowarn:3", "--format=general-"], ["--twins=0..40", "--queryhsplimit=5", "--format=general-", "--nogapped"],
                                  ["--recoverseeds", "--queryhsplimit+=8", "--format=general-"]], ids=lambda o: " ".join(o))
def test_oracle_queryhsplimit_corner_cases(opts):
    """found by tools/ref_sweep.py: rdotplot prints block by block, so the per-query print cap counts blocks (deGapifyOutput,
    lastz.c:7410, output.c:744); the twin processor never counts its HSPs, so the scan is not stopped for it"""
    if not os.path.exists(REF_CLI):
        pytest.skip("oracle/_ref not built")
    args = [AGLOBIN + "/human", AGLOBIN + "/cow"] + opts
    same_output(run_cli(ORACLE_CLI, args)[0], run_cli(REF_CLI, args)[0])


def test_oracle_queryhsplimit_keep_spelling_fails_like_the_reference():
    for cli in (ORACLE_CLI, REF_CLI):
        if not os.path.exists(cli):
            continue
        p = subprocess.run([cli, AGLOBIN + "/human", AGLOBIN + "/cow", "--queryhsplimit=keep:5"], capture_output=True, text=True)
        assert p.returncode != 0 and "is not an integer" in p.stderr
