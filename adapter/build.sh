#!/bin/bash
# Build the reference lastz with its hot path bound to include/lastz_b200.h through adapter/lastz_adapter.c:
#   adapter/_build/lastz_adapter_oracle   linked against oracle/liblzb_oracle.so      (CPU check of the boundary)
#   adapter/_build/lastz_adapter_b200     linked against lastz_b200/csrc/liblastz_b200.so (the B200 drop-in)
# The reference's sources are compiled where they lie under /root/reference/src (nothing is copied); in the three objects
# that hold the hot path the five replaced entry points are renamed ref_* with objcopy so that the adapter's definitions
# are the ones lastz.o calls.  TEST/INTEGRATION infrastructure; outputs are git-ignored and travel to the GPU box.
set -eo pipefail
REF=${REF:-/root/reference/src}
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(dirname "$HERE")"
OUT="$HERE/_build"
if [ ! -d "$REF" ]; then echo "adapter/build.sh: $REF not present (GPU box?) - keeping prebuilt adapter/_build" >&2; exit 0; fi
mkdir -p "$OUT/obj"
SRCS="lastz infer_scores seeds pos_table quantum seed_search diag_hash chain gapped_extend tweener masking segment edit_script identity_dist coverage_dist continuity_dist output gfa lav axt maf cigar sam genpaf text_align align_diffs utilities dna_utilities sequences capsule"
VER=(-DVERSION_MAJOR='"1"' -DVERSION_MINOR='"04"' -DVERSION_SUBMINOR='"58"' -DREVISION_DATE='"20260507"' -DSUBVERSION_REV='""')
COMMON="-O3 -w -D_FILE_OFFSET_BITS=64 -D_LARGEFILE_SOURCE -Dscore_type=I"
for s in $SRCS; do gcc -c $COMMON "${VER[@]}" "$REF/$s.c" -o "$OUT/obj/$s.o" & done
wait
for s in pos_table seed_search gapped_extend; do
  objcopy --redefine-sym build_seed_position_table=ref_build_seed_position_table --redefine-sym free_position_table=ref_free_position_table \
          --redefine-sym seed_hit_search=ref_seed_hit_search --redefine-sym reduce_to_points=ref_reduce_to_points \
          --redefine-sym gapped_extend=ref_gapped_extend "$OUT/obj/$s.o"
done
gcc -c $COMMON -std=gnu99 -Wall -I"$REF" -I"$ROOT/include" "$HERE/lastz_adapter.c" -o "$OUT/obj/lastz_adapter.o"
gcc "$OUT"/obj/*.o -L"$ROOT/oracle" -llzb_oracle -Wl,-rpath,"$ROOT/oracle" -lm -o "$OUT/lastz_adapter_oracle"
if [ -f "$ROOT/lastz_b200/csrc/liblastz_b200.so" ]; then
  gcc "$OUT"/obj/*.o -L"$ROOT/lastz_b200/csrc" -llastz_b200 -Wl,-rpath,"$ROOT/lastz_b200/csrc" -lm -o "$OUT/lastz_adapter_b200"
fi
echo "built $OUT/lastz_adapter_oracle $( [ -f "$OUT/lastz_adapter_b200" ] && echo "$OUT/lastz_adapter_b200" )"
