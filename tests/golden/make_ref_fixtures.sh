#!/bin/sh
# Fixtures that are OUTPUTS OF THE REFERENCE BINARY (oracle/_ref/lastz, built from /root/reference by
# oracle/build_ref.sh), as opposed to the files copied from the reference's test_data.  Run from the repo root.
set -e
G=tests/golden
# order of raw seed hits where several transition variants hit at one query position
oracle/_ref/lastz $G/aglobin.2bit/human "$G/aglobin.2bit/cow[20000..32000]" --nogfextend --nogapped --strand=plus \
    --format=general- | cut -f5,10 > $G/aglobin_cow_20k_32k.plus_hits.order.tsv
# (shorties.fq is not a reference output: it is shorties.fa rewritten as FASTQ with seeded random qualities by
#  tests/golden/make_fastq.py; the comparisons that use it run the reference binary on the same file)
