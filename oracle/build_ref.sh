#!/bin/bash
# Build the UNMODIFIED reference lastz (C99, libm only) from the sources where they
# lie under /root/reference/src, straight with gcc (the reference's own Makefile is not
# run).  Outputs go only into oracle/_ref/ (git-ignored, travels to the GPU box).
#   lastz        stock integer build                      (src/Makefile:76,95-109)
#   lastz_32     32-bit positions, 4M-entry diag hash     (src/Makefile:59)
#   lastz_stats  -Dcollect_stats counter build (--stats prints raw seed hits, DP cells)
# This is TEST/BENCH infrastructure (the checker and the CPU baseline), never the product.
set -e
REF=${REF:-/root/reference/src}
OUT="$(cd "$(dirname "$0")" && pwd)/_ref"
if [ ! -d "$REF" ]; then
  echo "build_ref.sh: $REF not present (GPU box?) - keeping prebuilt oracle/_ref" >&2
  exit 0
fi
mkdir -p "$OUT"
SRCS="lastz infer_scores seeds pos_table quantum seed_search diag_hash chain gapped_extend tweener masking segment edit_script identity_dist coverage_dist continuity_dist output gfa lav axt maf cigar sam genpaf text_align align_diffs utilities dna_utilities sequences capsule"
VER=(-DVERSION_MAJOR='"1"' -DVERSION_MINOR='"04"' -DVERSION_SUBMINOR='"58"' -DREVISION_DATE='"20260507"' -DSUBVERSION_REV='""')
COMMON="-O3 -w -D_FILE_OFFSET_BITS=64 -D_LARGEFILE_SOURCE"
build() { # name, extra flags
  local name=$1; shift
  if [ -x "$OUT/$name" ] && [ "$OUT/$name" -nt "$REF/lastz.c" ]; then return; fi
  local tmp; tmp=$(mktemp -d)
  for s in $SRCS; do
    gcc -c $COMMON "${VER[@]}" "$@" "$REF/$s.c" -o "$tmp/$s.o" &
  done
  wait
  gcc "$tmp"/*.o -lm -o "$OUT/$name"
  rm -rf "$tmp"
  echo "built $OUT/$name"
}
build lastz       -Dscore_type=I
build lastz_32    -Dmax_sequence_index=32 -Dmax_malloc_index=40 -Ddiag_hash_size=4194304
build lastz_stats -Dscore_type=I -Dcollect_stats
