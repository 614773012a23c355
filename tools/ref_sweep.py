"""Randomised comparison of the front end (linked with the oracle library: oracle/lastz_oracle) against the reference
binary (oracle/_ref/lastz): random pairs of fixtures, seeds, options and output formats; prints every command line whose
stdout differs.  One-sided refusals are listed too (the reference rejects more option combinations than this front end).
This is how the transition-variant order bug of round 1 would have been found earlier: run it after touching the hot
path's semantics.  Usage, from the repo root:  python tools/ref_sweep.py <seed> <iterations>"""
import os
import random
import subprocess
import sys

G = os.path.join("tests", "golden")
REF, OURS = os.path.join("oracle", "_ref", "lastz"), os.path.join("oracle", "lastz_oracle")
PAIRS = [("aglobin.2bit/human", "aglobin.2bit/cow"), ("aglobin.2bit/cow", "aglobin.2bit/human"), ("pseudocat.fa", "pseudopig.fa"),
         ("aglobin.2bit/human", "shorties.fa"), ("pseudopig2.fa", "pseudocat.fa"), ("aglobin.2bit/cow[5000..40000]", "aglobin.2bit/human[10000..60000]"),
         ("aglobin.2bit[multi]", "shorties.fa"), ("aglobin.2bit/human", "shorties.fa[multi]"), ("aglobin.2bit/human", "shorties.fq")]
SEEDS = ["", "--seed=12of19", "--seed=14of22", "--seed=111010011101", "--seed=1T1001100T010T01111", "--seed=match10", "W=9",
         "--seed=TTT1T11T1TT1T1T", "--seed=110101101"]
OPTIONS = [["T=0"], ["--transition=2"], ["--notransition"], ["--step=2"], ["--step=5"], ["--strand=plus"], ["--strand=minus"], ["K=1800"],
           ["K=2600", "L=2000"], ["X=400"], ["X=1500"], ["Y=3000"], ["Y=15000"], ["--noentropy"], ["--nogapped"], ["--chain"], ["--chain=10,20"],
           ["--noytrim"], ["--allgappedbounds"], ["O=300", "E=40"], ["--exact=18"], ["--mismatch=2,28"], ["--nogfextend"], ["--ambiguous=n"],
           ["--allocate:traceback=200K"], ["--match=1,2"], ["--identity=70"], ["K=top15%"], ["--notrivial"],
           ["--recoverseeds"], ["--twins=0..40"], ["--twins=-8..25"], ["--twins=20..150", "--seedqueue=300"], ["--maxwordcount=30"], ["--maxwordcount=90%"],
           ["--queryhsplimit=25"], ["--queryhsplimit=keep,nowarn:12"], ["--queryhsplimit+=30"], ["--querydepth=keep:0.2"], ["--querydepth=0.5"]]
FORMATS = ["--format=general-", "--format=lav", "--format=maf-", "--format=axt", "--format=sam-", "--format=cigar", "--format=paf",
           "--format=rdotplot", "--format=general-:name1,start1,end1,name2,start2+,end2+,cigarx,nmatch,ngap,diff"]


def main():
    random.seed(int(sys.argv[1]))
    diffs = 0
    for _ in range(int(sys.argv[2])):
        t, q = random.choice(PAIRS)
        opts = [s for s in [random.choice(SEEDS)] if s]
        for o in random.sample(OPTIONS, random.randint(0, 4)):
            opts += o
        opts.append(random.choice(FORMATS))
        argv = [os.path.join(G, t), os.path.join(G, q)] + opts
        try:                                             # low-weight seeds without extension make millions of anchors: skip what takes minutes
            r = subprocess.run([REF] + argv, capture_output=True, text=True, timeout=60)
            o = subprocess.run([OURS] + argv, capture_output=True, text=True, timeout=180)
        except subprocess.TimeoutExpired:
            print("TIMEOUT", " ".join(argv[2:]))
            continue
        if r.returncode != 0 and o.returncode != 0:
            continue
        if r.returncode != 0 or o.returncode != 0:
            print("ONE-SIDED", "reference refuses:" if r.returncode else "this front end refuses:", " ".join(argv[2:]), "|",
                  (r.stderr if r.returncode else o.stderr).strip().split("\n")[0][:140])
        elif r.stdout != o.stdout:
            diffs += 1
            print("DIFF", " ".join(argv))
    print("done, diffs:", diffs)
    return 1 if diffs else 0


if __name__ == "__main__":
    sys.exit(main())
