/* c_abi_scan.c -- the drop-in boundary used from plain C: no Python, no torch, nothing but include/grafimo_b200.h.
 *
 *   gcc -O2 -I include examples/c_abi_scan.c -o c_abi_scan -L grafimo_b200 -lgrafimo_b200 -Wl,-rpath,$PWD/grafimo_b200 -lm
 *   ./c_abi_scan MOTIF.bin SEQS.txt [p-threshold]
 *
 * MOTIF.bin: int64 w | int64 min_val | int64 scale | double offset | double bg[4] (A,C,G,T) | int64 score_matrix[4][w] (rows
 * A,C,G,T) -- what the reference's Motif carries after scale_pwm (src/grafimo/motif_ops.py:1027-1111).  SEQS.txt: one sequence
 * per line (ACGTN, any case).  The program runs the score-distribution DP on the GPU (replaces comp_pval_mat,
 * src/grafimo/motif_processing.pyx:552-632), uploads the motif, scans every window of every sequence on both strands
 * (replaces the scoring, q-value and filter / sort steps of compute_results, src/grafimo/score_sequences.py:111-207) and
 * prints one line per hit: window row, strand, integer score, log-odds score, p-value, q-value.
 * tests/test_gpu_c_abi.py builds it, runs it and compares the lines with the Python binding's table for the same input. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "grafimo_b200.h"

#define CHECK(call)                                                                                        \
    do {                                                                                                   \
        int rc__ = (call);                                                                                 \
        if (rc__ != GB2_OK) {                                                                              \
            fprintf(stderr, "%s failed: %s (%s)\n", #call, gb2_error_string(rc__), ctx ? gb2_ctx_last_error(ctx) : ""); \
            return 1;                                                                                      \
        }                                                                                                  \
    } while (0)

int main(int argc, char **argv)
{
    gb2_ctx *ctx = NULL;
    if (argc < 3) {
        fprintf(stderr, "usage: %s MOTIF.bin SEQS.txt [p-threshold]\n", argv[0]);
        return 2;
    }
    const double threshold = argc > 3 ? atof(argv[3]) : 1e-3;
    /* ---- motif */
    FILE *fm = fopen(argv[1], "rb");
    if (!fm) { perror(argv[1]); return 2; }
    int64_t head[3];
    double offset, bg[4];
    if (fread(head, 8, 3, fm) != 3 || fread(&offset, 8, 1, fm) != 1 || fread(bg, 8, 4, fm) != 4) return 2;
    const int w = (int)head[0];
    int64_t *sm = (int64_t *)malloc(sizeof(int64_t) * 4 * (size_t)w);
    if (fread(sm, 8, 4 * (size_t)w, fm) != 4 * (size_t)w) return 2;
    fclose(fm);
    /* ---- sequences: one buffer, offsets and lengths (the layout gb2_scan_host_sequences takes) */
    FILE *fs = fopen(argv[2], "rb");
    if (!fs) { perror(argv[2]); return 2; }
    fseek(fs, 0, SEEK_END);
    const long n_bytes = ftell(fs);
    fseek(fs, 0, SEEK_SET);
    uint8_t *text = NULL;
    if (gb2_host_alloc((uint64_t)n_bytes + 1, (void **)&text) != GB2_OK) return 3; /* pinned: copied at the PCIe rate */
    if (fread(text, 1, (size_t)n_bytes, fs) != (size_t)n_bytes) return 2;
    fclose(fs);
    int64_t n_seqs = 0, cap_seqs = 1024;
    int64_t *off = (int64_t *)malloc(8 * (size_t)cap_seqs), *len = (int64_t *)malloc(8 * (size_t)cap_seqs);
    for (long i = 0; i < n_bytes;) {
        long j = i;
        while (j < n_bytes && text[j] != '\n') ++j;
        if (j > i) {
            if (n_seqs == cap_seqs) {
                cap_seqs *= 2;
                off = (int64_t *)realloc(off, 8 * (size_t)cap_seqs);
                len = (int64_t *)realloc(len, 8 * (size_t)cap_seqs);
            }
            off[n_seqs] = i;
            len[n_seqs] = j - i;
            ++n_seqs;
        }
        i = j + 1;
    }
    /* ---- GPU */
    if (gb2_abi_version() != GB2_ABI_VERSION) { fprintf(stderr, "header / library mismatch\n"); return 3; }
    CHECK(gb2_ctx_create(0, NULL, &ctx));
    const int32_t widths[1] = {w};
    double *pval = (double *)malloc(sizeof(double) * ((size_t)GB2_RANGE * (size_t)w + 1));
    CHECK(gb2_pval_dp_batched(ctx, 1, widths, sm, bg, pval));
    gb2_motif *motif = NULL;
    CHECK(gb2_motif_create(ctx, sm, w, pval, head[1], head[2], offset, &motif));
    uint64_t cap = 1 << 20, n_hits = 0, stats[4];
    uint64_t *row = (uint64_t *)malloc(8 * cap);
    uint8_t *strand = (uint8_t *)malloc(cap);
    int32_t *iscore = (int32_t *)malloc(4 * cap);
    double *score = (double *)malloc(8 * cap), *p = (double *)malloc(8 * cap), *q = (double *)malloc(8 * cap);
    CHECK(gb2_scan_host_sequences(ctx, motif, 0, text, NULL, n_seqs, off, len, 2, threshold, 0, 1, cap, row, strand, iscore, score,
                                  p, q, &n_hits, stats));
    uint64_t h2d = 0, d2h = 0, given = 0, packed = 0;
    CHECK(gb2_scan_last_transfer(ctx, &h2d, &d2h, &given, &packed));
    fprintf(stderr, "%lld sequences, %llu windows scored, %llu N bases, %llu hits; %llu bytes to the device (%llu chunks as text, %llu packed on the host)\n",
            (long long)n_seqs, (unsigned long long)stats[0], (unsigned long long)stats[1], (unsigned long long)n_hits,
            (unsigned long long)h2d, (unsigned long long)given, (unsigned long long)packed);
    for (uint64_t k = 0; k < n_hits; ++k)
        printf("%llu\t%c\t%d\t%.17g\t%.17g\t%.17g\n", (unsigned long long)row[k], strand[k] ? '-' : '+', iscore[k], score[k], p[k], q[k]);
    CHECK(gb2_motif_destroy(motif));
    gb2_host_free(text);
    gb2_ctx_destroy(ctx);
    return 0;
}
