import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
ORACLE_CLI = os.path.join(ROOT, "oracle", "lastz_oracle")
REF_CLI = os.path.join(ROOT, "oracle", "_ref", "lastz")
PRODUCT_CLI = os.path.join(ROOT, "lastz_b200", "csrc", "lastz_b200")
GEN_SYNTH = os.path.join(ROOT, "tools", "gen_synth")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    # the checker (oracle + synthetic generator) is plain gcc; build it on demand
    if not (os.path.exists(ORACLE_CLI) and os.path.exists(os.path.join(ROOT, "oracle", "liblzb_oracle.so"))):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "liblzb_oracle.so", "lastz_oracle"], check=True)
    if not os.path.exists(GEN_SYNTH):
        subprocess.run(["gcc", "-O2", "-o", GEN_SYNTH, os.path.join(ROOT, "tools", "gen_synth.c")], check=True)
    if not os.path.exists(os.path.join(ROOT, "lastz_b200", "csrc", "liblzb_host.so")):
        subprocess.run(["make", "-C", os.path.join(ROOT, "lastz_b200", "csrc"), "liblzb_host.so"], check=True)


def run_cli(cli, args, cwd=None):
    """stdout of one command-line run; stderr is returned too (the reference warns there)."""
    p = subprocess.run([cli] + list(args), cwd=cwd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError(f"{cli} {' '.join(args)} failed ({p.returncode}): {p.stderr[-2000:]}")
    return p.stdout, p.stderr


def lav_body(text):
    """LAV with the d-stanza command line dropped (tools/lav_compare.py:41-78 ignores it too)."""
    out, in_d, ix = [], False, 0
    for line in text.splitlines():
        s = line.rstrip()
        if s == "d {":
            in_d, ix = True, 0
            out.append(s)
            continue
        if in_d:
            ix += 1
            if s == "}":
                in_d = False
            elif ix == 1:
                continue
        out.append(s)
    return out


@pytest.fixture(scope="session")
def synth(tmp_path_factory):
    """synthetic pairs by size (SURVEY.md 8d generator), cached per session"""
    d = tmp_path_factory.mktemp("synth")
    made = {}

    def get(n, seed=20260925):
        key = (n, seed)
        if key not in made:
            t, q = str(d / f"t{n}_{seed}.fa"), str(d / f"q{n}_{seed}.fa")
            subprocess.run([GEN_SYNTH, str(n), str(seed), t, q], check=True)
            made[key] = (t, q)
        return made[key]
    return get


def write_inverted_repeat_fasta(path, seed=7):
    """A 60 kb sequence with a reverse-complement copy, a hairpin, a direct repeat and an exact palindrome:
    --self alignments that lie below, touch and cross the main diagonal on the minus strand
    (mirror_alignments lastz.c:4229, edit_script_upper_truncate edit_script.c:573)."""
    import random
    rnd = random.Random(seed)
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}

    def rc(s):
        return "".join(comp[c] for c in reversed(s))

    def mut(s, r):
        return "".join(rnd.choice("ACGT") if rnd.random() < r else c for c in s)
    s = "".join(rnd.choice("ACGT") for _ in range(60000))
    s = s[:30000] + mut(rc(s[5000:9000]), 0.05) + s[34000:]
    s = s[:45500] + "ACGTTGCA" + mut(rc(s[44000:45500]), 0.03) + s[45500 + 8 + 1500:]
    s = s[:52000] + mut(s[12000:14000], 0.04) + s[54000:]
    s = s[:20600] + rc(s[20000:20600]) + s[21200:]
    with open(path, "w") as f:
        f.write(">selfy\n" + "\n".join(s[i:i + 60] for i in range(0, len(s), 60)) + "\n")


# --self runs (BASELINE.json configs[1] is the first): target spec, options
SELF_CASES = [
    ("aglobin", ["--self", "--seed=12of19", "--nogapped"]),
    ("aglobin", ["--self", "--seed=12of19", "--nogapped", "--format=segments"]),
    ("aglobin", ["--self", "--nogapped", "--chain"]),
    ("aglobin", ["--self"]),
    ("inverted", ["--self"]),
    ("inverted", ["--self", "--nogapped"]),
    ("inverted", ["--self", "--chain"]),
    ("inverted", ["--self", "--hspthresh=2000", "--ydrop=5000"]),
]


def self_case_target(which, tmp_path):
    if which == "aglobin":
        return os.path.join(GOLDEN, "aglobin.2bit") + "/human"
    p = str(tmp_path / "selfy.fa")
    write_inverted_repeat_fasta(p)
    return p


# --exact / --mismatch extension (match_extend_seed_hit seed_search.c:3018, mismatch_extend_seed_hit :3450) and
# the scoreless-anchor rescoring that chaining / the gapped stage need (lastz.c:3336)
ALT_EXTEND_FIXTURE_CASES = [
    ["--exact=14", "--nogapped", "W=8", "T=0"],
    ["--exact=10", "--nogapped", "W=8", "T=0"],
    ["--mismatch=2,20", "--nogapped"],
    ["--2mismatch=18", "--nogapped"],
    ["--mismatch=1,12", "--nogapped", "--seed=match6"],
    ["--mismatch=5,25", "--nogapped", "--step=3"],
    ["--exact=14", "W=8", "T=0"],
]
ALT_EXTEND_SYNTH_CASES = [
    ["--exact=30", "--nogapped"],
    ["--mismatch=3,60", "--nogapped"],
    ["--mismatch=10,200", "--nogapped", "--seed=14of22"],
    ["--mismatch=2,50", "--chain", "--nogapped"],
    ["--mismatch=4,80"],
    ["--nogfextend", "--seed=match12", "--step=20"],
]


def masked_query(synth, tmp_path, size=300000):
    """synthetic pair whose query has a soft-masked stretch and a run of N (case-insensitive matching, invalid bases)"""
    t, q = synth(size)
    lines = open(q).read().split("\n")
    b = "".join(lines[1:])
    b = b[:50000] + b[50000:52000].lower() + b[52000:90000] + "N" * 300 + b[90300:]
    qm = str(tmp_path / "qm.fa")
    with open(qm, "w") as f:
        f.write(lines[0] + "\n" + "\n".join(b[i:i + 60] for i in range(0, len(b), 60)) + "\n")
    return t, qm


def same_output(got, want):
    """byte for byte, job header lines included (both programs are given the same command line)"""
    assert got == want


# --format=general (default fields, genpaf.c:595-690 / genpaf.h:117): target spec suffix, query spec suffix, options
GENERAL_CASES = [
    ("", "", ["--format=general"]),
    ("", "", ["--format=general", "--nogapped"]),
    ("", "", ["--format=general-", "--chain"]),
    ("", "", ["--format=general", "--strand=minus", "K=2200"]),
    ("[2000..15000]", "[500..20000]", ["--format=general"]),
    ("", "", ["--format=maf-"]),                                  # print_maf_align maf.c:271
    ("", "", ["--format=maf-", "--nogapped"]),
    ("[2000..15000]", "[500..20000]", ["--format=maf-", "--chain"]),
    ("", "", ["--format=axt"]),                                   # print_axt_align axt.c:96, header included
    ("", "[500..20000]", ["--format=axt", "--nogapped"]),
    ("", "", ["--format=maf", "--chain", "Y=5000"]),              # MAF with the parameter header (maf.c:96)
]

# --format=general:<fields>, mapping, cigar (genpaf.c:545-1490, cigar.c:135): every field this front end carries
FIELD_CASES = [
    ("", "", ["--format=general:name1,number1,strand1,size1,start1,zstart1,end1,length1,name2,number2,strand2,size2,start2,zstart2,start2+,zstart2+,end2,end2+,length2"]),
    ("", "", ["--format=general:score,nmatch,nmismatch,npair,ncolumn,ngap,cgap,diagonal,shingle,number,znumber,NA"]),
    ("", "", ["--format=general-:identity,idfrac,id%,blastid%,coverage,covfrac,cov%,continuity,confrac,con%,gaprate"]),
    ("", "", ["--format=general-:text1,text2,diff", "--strand=minus", "K=2200"]),
    ("", "", ["--format=general-:cigar,cigar-,cigarx,cigarx-,cigarx1,cigarx1-"]),
    ("", "", ["--format=general-:n1,s1,z1,e1,l1,n2,s2,z2,s2+,z2+,e2,e2+,l2,d,diag,s,id,ident,cov,con,gap", "--nogapped"]),
    ("[100..9000]", "[2000..18000]", ["--format=general:name1,size1,start1,end1,name2,strand2,size2,start2,end2,start2+,end2+,shingle,coverage"]),
    ("", "", ["--format=mapping"]),
    ("", "[multi]", ["--format=mapping-", "--nogapped"]),
    ("", "", ["--format=cigar"]),
    ("[100..9000]", "[2000..18000]", ["--format=cigar", "--nogapped"]),
    ("", "", ["--format=sam"]),                                   # sam.c:196-560: hard clipping, header
    ("", "", ["--format=softsam-", "--strand=minus"]),            # soft clipping: the whole read, clipped parts in lower case
    ("", "", ["--format=sam+eqx"]),                               # =/X runs instead of M
    ("", "[multi]", ["--format=softsam+eqx-", "--nogapped"]),
    ("", "[2000..18000]", ["--format=sam-"]),
    ("", "", ["--format=paf"]),                                   # genpafPafMinimap2Keys: cg:Z: with M runs
    ("", "[multi]", ["--format=paf:wfmash", "--strand=minus"]),   # ... with =/X runs
    ("", "", ["--format=blastn"]),                                # comment block per query and strand, e-value and bit score
    ("", "", ["--format=blastn-", "--nogapped"]),
    ("", "", ["--format=rdotplot"]),                              # one line segment per gap-free block, names announced once
    ("[100..9000]", "[2000..18000]", ["--format=rdotplot+score"]),
    ("", "[multi]", ["--format=rdotplot+score", "--nogapped", "--strand=minus"]),
]

# --format=gfa (gfa.c:95-330): A lines + one a line per gap-free block
GFA_CASES = [
    ("", "", ["--format=gfa"]),
    ("", "[500..20000]", ["--gfa", "--nogapped", "--transition=2", "--hspthresh=2200"]),
    ("", "", ["--gfa", "--chain", "--seed=14of22", "--notransition", "--step=4"]),
]


# [multi] queries (sequences.h:188-191: the sequences of a file joined by NULs; hits, extensions and DP sweeps stop at them)
MULTI_QUERY_CASES = [["--format=general-"], ["--format=general-", "--nogapped"], ["--format=maf-", "--strand=minus"],
                     ["--format=axt", "W=8", "T=0"], ["--format=general-", "--exact=14", "--nogapped", "W=8", "T=0"]]

# [multi] TARGETS (and both sides partitioned): (target, query, options), files relative to tests/golden
MULTI_TARGET_CASES = [
    ("pseudopig.fa[multi]", "pseudocat.fa", ["--format=general-"]),
    ("pseudopig.fa[multi]", "pseudocat.fa", ["--format=maf", "--nogapped"]),
    ("aglobin.2bit[multi]", "shorties.fa", ["--format=general-", "K=2500"]),
    ("aglobin.2bit[multi]", "shorties.fa[multi]", ["--format=axt", "K=2500"]),
    ("shorties.fa[multi]", "aglobin.2bit/human", ["--format=axt", "K=2500", "--noytrim"]),
    ("shorties.2bit[multi,51..200]", "aglobin.2bit/human", ["--format=maf-", "K=3000", "--strand=minus"]),
    ("aglobin.2bit[multi]", "shorties.fa[multi]", ["--format=general-", "K=2500", "--chain"]),          # chained per pair of partitions, chain.c:224
    ("aglobin.2bit/human", "shorties.fa[multi]", ["--format=maf-", "K=2000", "--chain=20,30", "--nogapped"]),
    # a query that IS one partition of the target: trivial self-alignment of that partition (gapped_extend.c:1185-1230)
    ("aglobin.2bit[multi]", "aglobin.2bit/cow", ["--format=general-"]),
    ("shorties.fa[multi]", "shorties.fa", ["--format=general-", "K=2000"]),        # all against all, 20 x 20
    ("names.fa[multi]", "names.fa", ["--format=axt", "K=2000"]),
    ("aglobin.2bit[multi]", "shorties.fa[multi]", ["--format=segments", "K=2500", "--chain"]),     # genpafSegmentKeys rows, partition names
]

# --filter= family (lastz.c:6672-6950; filter_aligns_by_* after the gapped stage, filter_segments_by_* on HSPs): aglobin human x cow
FILTER_CASES = [
    ["--format=general-", "--identity=60..75"],
    ["--format=general-", "--filter=identity:..72.5%"],
    ["--format=general-", "--filter=coverage:0.5..3"],
    ["--format=general-", "--filter=continuity:90..97%"],
    ["--format=general-", "--filter=nmatch:1K"],
    ["--format=general-", "--filter=nmismatch:0..200", "--filter=ngap:0..5"],
    ["--format=general-", "--filter=cgap:..30"],
    ["--format=general-", "--nogapped", "--coverage=0.1", "--matchcount=60", "K=2000"],
    ["--format=general-", "--nogapped", "--filter=nmismatch:0..20", "K=2000", "--chain"],
    ["--format=maf-", "--identity=65", "--continuity=90", "--coverage=0.8"],
    # --match=<reward>[,<penalty>]: unit scores and the thresholds/penalties derived from them (lastz.c:9168-9236)
    ["--match=1,3", "--chain", "--format=lav"],
    ["--match=1,1", "--exact=30", "--format=general-"],
    ["--match=5,4", "O=30", "E=3", "X=40", "--format=maf-"],
    ["--match=10,30", "--format=general-", "--nogapped"],
]

# FASTQ queries (load_fastq_sequence sequences.c:2540): four-line records, qualities carried to the SAM writer
FASTQ_CASES = [
    ("shorties.fq", ["--format=general-", "K=2500"]),
    ("shorties.fq", ["--format=sam", "K=2500"]),                    # QUAL column from the file, reversed on the - strand
    ("shorties.fq", ["--format=softsam+eqx-", "K=2500", "--nogapped"]),
    ("shorties.fq[multi]", ["--format=softsam", "K=2500", "--strand=minus"]),
    ("shorties.fq[20..150]", ["--format=maf-", "K=2000"]),
    # read-mapping shortcuts (expanders[] lastz.c:559-577): seeds with two transitions, unit scores, identity filter, N ambiguous
    ("shorties.fq", ["--yasra85", "--format=general-"]),
    ("shorties.fq", ["--yasra95short", "--format=softsam-"]),
    ("shorties.fa", ["--yasra90", "--format=lav"]),
]

# adaptive HSP threshold K=top<N>% / K=top<bases> (add_segment's coverage-limited min-heap segment.c:981-1180, both
# strands collected into one table and the - strand finished first, lastz.c:1426,1678-1700): target suffix, query, options
ADAPTIVE_CASES = [
    ("aglobin", ["K=top50%", "C=3", "W=8", "T=0", "--noentropy", "--gfa"]),          # base_test_adaptive_k, Makefile:317
    ("aglobin", ["K=top50%", "--nogapped", "--format=general-"]),                      # entropy decided against the table as it stands
    ("aglobin", ["K=top5K", "--format=maf-"]),                                         # gapped threshold = lowest HSP kept
    ("aglobin", ["K=top5K", "--strand=minus", "--format=maf-"]),                       # one strand: the second table is empty
    ("aglobin", ["K=top30%", "--chain", "--format=general-"]),
    ("catpig", ["--hspthresh=top2.5%", "--nogapped", "--chain"]),                      # + strand HSPs rescored after the split
    ("catpig", ["K=top20%", "--format=lav"]),                                          # header prints top<bases>
    ("catpig", ["K=top20%", "L=2500", "--format=axt"]),
    ("catpig", ["K=top1500", "--format=general-"]),
    ("catpig", ["K=top100%", "--nogapped", "--format=general-"]),                      # the limit is never met: a plain list
]


def adaptive_case_files(which):
    if which == "aglobin":
        return [os.path.join(GOLDEN, "aglobin.2bit") + "/human", os.path.join(GOLDEN, "aglobin.2bit") + "/cow"]
    return [os.path.join(GOLDEN, "pseudocat.fa"), os.path.join(GOLDEN, "pseudopig.fa")]


# --anyornone (gappily_extend_hsps gapped_extend.c:5279): first HSP, in discovery order, whose
# unconstrained gapped extension reaches the threshold; one alignment per query, both strands
ANYORNONE_CASES = [
    ["--anyornone", "--format=general-"],
    ["--anyornone"],
    ["--anyornone", "--nogapped", "--format=general-"],
    ["--anyornone", "--strand=minus", "--format=maf-"],
    ["--anyornone", "--gappedthresh=60000", "--format=general-"],
    ["--anyornone", "W=8", "T=0", "--hspthresh=2500", "--gfa"],
]
