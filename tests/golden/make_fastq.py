"""shorties.fq: the 20 sequences of shorties.fa (copied from the reference's test_data) as four-line FASTQ records
with seeded random qualities; every other record repeats its name on the '+' line.  Run from the repo root."""
import os
import random

here = os.path.dirname(os.path.abspath(__file__))
random.seed(5)
recs, name, seq = [], None, []
for line in open(os.path.join(here, "shorties.fa")):
    line = line.rstrip("\n")
    if line.startswith(">"):
        if name:
            recs.append((name, "".join(seq)))
        name, seq = line[1:], []
    else:
        seq.append(line)
recs.append((name, "".join(seq)))
with open(os.path.join(here, "shorties.fq"), "w") as f:
    for i, (n, s) in enumerate(recs):
        q = "".join(chr(33 + random.randrange(40)) for _ in s)
        f.write("@%s\n%s\n+%s\n%s\n" % (n, s, n if i % 2 else "", q))
