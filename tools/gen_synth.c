/*
 * gen_synth.c -- deterministic synthetic sequence pairs (normative).
 *   gen_synth <L> <seed> <target.fa> <query.fa>
 * PRNG = splitmix64.  Target (stream S): base i = "ACGT"[next()>>62].  Query (stream S+1), for
 * each target base c: r = next(), u = r & 0xFFFFFF; u < 671089 (4%) -> substitution to
 * "ACGT"[(code(c)+1+((r>>24)%3))&3]; else u < 754975 (0.5%) -> deletion; else u < 838861 (0.5%)
 * -> c then the inserted base "ACGT"[(r>>24)&3]; else c.  One FASTA record each (t / q), 60 columns.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

static uint64_t st;
static uint64_t next64(void) {
    st += 0x9E3779B97F4A7C15ull; uint64_t z = st;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static void put(FILE* f, char c, uint64_t* col) { fputc(c, f); if (++*col % 60 == 0) fputc('\n', f); }

int main(int argc, char** argv) {
    if (argc != 5) { fprintf(stderr, "usage: gen_synth <L> <seed> <target.fa> <query.fa>\n"); return 1; }
    uint64_t L = strtoull(argv[1], 0, 10), S = strtoull(argv[2], 0, 10);
    char* t = malloc(L);
    st = S;
    for (uint64_t i = 0; i < L; i++) t[i] = "ACGT"[next64() >> 62];
    FILE* f = fopen(argv[3], "w"); uint64_t col = 0;
    fprintf(f, ">t\n"); for (uint64_t i = 0; i < L; i++) put(f, t[i], &col);
    if (col % 60) fputc('\n', f);
    fclose(f);
    f = fopen(argv[4], "w"); col = 0; st = S + 1;
    fprintf(f, ">q\n");
    for (uint64_t i = 0; i < L; i++) {
        char c = t[i]; int code = c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : 3;
        uint64_t r = next64(); uint32_t u = (uint32_t)(r & 0xFFFFFF);
        if (u < 671089) put(f, "ACGT"[(code + 1 + (int)((r >> 24) % 3)) & 3], &col);
        else if (u < 754975) ;
        else if (u < 838861) { put(f, c, &col); put(f, "ACGT"[(r >> 24) & 3], &col); }
        else put(f, c, &col);
    }
    if (col % 60) fputc('\n', f);
    fclose(f); free(t);
    return 0;
}
