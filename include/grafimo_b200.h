/*
 * grafimo_b200.h -- C ABI of the B200-native GRAFIMO motif-scanning hot path.
 *
 * The reference (pinellolab/GRAFIMO, pure Python) has no FFI; the seams this library sits behind
 * are two Python call sites (paths relative to the reference tree):
 *
 *   B1  grafimo.score_sequences.compute_results      src/grafimo/score_sequences.py:44-211
 *   B2  motif_processing.comp_pval_mat               src/grafimo/motif_processing.pyx:608-632
 *   B3  grafimo.score_sequences.compute_qvalues      src/grafimo/score_sequences.py:401-428
 *
 * Every entry point below names the reference lines it replaces.  The ctypes binding a GRAFIMO
 * maintainer would add is shown in INTEGRATION.md; grafimo_b200/_lib.py is that binding.
 *
 * Conventions
 *   - plain C, no exceptions across the boundary; every function returns an int status
 *     (GB2_OK == 0); gb2_error_string() / gb2_ctx_last_error() give the text;
 *   - pointers named d_* are DEVICE pointers (cudaMalloc / torch allocations, 16-byte aligned),
 *     pointers named h_* are HOST pointers; there are no torch types in any signature;
 *   - calls are stream-ordered on the context's stream; a context belongs to one host thread
 *     and one device.  Outputs written to d_* buffers are valid after gb2_ctx_sync() (or after
 *     any later work on the same stream); h_* outputs are valid on return;
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     GB2_ERR_CUDA.
 *
 * Data layout
 *   packed k-mer  w <= 32 ("narrow"): one uint64 per k-mer -- base i in bits [2i, 2i+1], A=0 C=1 G=2 T=3, bits >= 2w
 *                 zero.  32 < w <= 64 ("wide"; the reference has no width limit and JASPAR holds 34/35-bp profiles):
 *                 TWO consecutive uint64 per k-mer, {bases 0..31, bases 32..w-1}, i.e. arrays named d_packed / d_kmer
 *                 hold 2n words and must be 16-byte aligned.  Which form an array has follows from the width given to
 *                 (or stored in the motif / prepared query of) the call.
 *   N mask        uint32[ceil(n/32)]: bit (r & 31) of word (r >> 5) is set when row r holds a
 *                 symbol other than ACGTacgt (the reference scores such rows as `min_val`,
 *                 p-value 1: score_sequences.py:376-378).
 *   histogram     uint64[span + 1], span = hi - lo + 1 with lo/hi the smallest/largest reachable
 *                 integer score; bin k counts scored windows with integer score lo + k, bin
 *                 `span` counts N-rows (score = min_val, p = 1).
 */
#ifndef GRAFIMO_B200_H
#define GRAFIMO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GB2_ABI_VERSION 3
#define GB2_NARROW_WIDTH 32 /* widest k-mer that fits one packed word */
#define GB2_MAX_WIDTH 64
#define GB2_RANGE 1000 /* src/grafimo/utils.py:26 */
#define GB2_NCCL_ID_BYTES 128 /* sizeof(ncclUniqueId) */

enum {
    GB2_OK = 0,
    GB2_ERR_ARG = 1,      /* bad argument (null pointer, width out of range, misaligned buffer) */
    GB2_ERR_CUDA = 2,     /* CUDA runtime failure (including "no device") */
    GB2_ERR_NOMEM = 3,    /* host or device allocation failed */
    GB2_ERR_CAPACITY = 4, /* more hits than the caller's buffer holds (count still returned) */
    GB2_ERR_MOTIF = 5,    /* motif not usable: not scaled, empty p-value matrix, index out of range */
    GB2_ERR_STATE = 6
};

typedef struct gb2_ctx gb2_ctx;     /* one per (host thread, device): stream + scratch */
typedef struct gb2_motif gb2_motif; /* device-resident motif: chunk LUTs, p-value table, cut-offs */
typedef struct gb2_graph gb2_graph; /* device-resident variation graph of one chromosome (K7) */

/* One hit (a scored window that passed the p-value test). 16 bytes. */
typedef struct gb2_hit {
    uint64_t row;    /* global row index = row_base + index within the scored batch */
    int32_t score;   /* absolute integer score (sum of scaled matrix entries) */
    uint32_t strand; /* 0 = the k-mer as given ('+' for forward k-mers), 1 = its reverse complement */
} gb2_hit;

/* Facts about an uploaded motif. */
typedef struct gb2_motif_info {
    int32_t width;
    int32_t n_chunks;      /* lookup-table chunks = ceil(width / chunk_bases) */
    int32_t lut_replicas;  /* shared-memory replication factor chosen for the scoring kernel */
    int32_t monotone;      /* 1 when the p-value table is non-increasing in the score */
    int64_t lo, hi;        /* smallest / largest reachable integer score */
    int64_t span;          /* hi - lo + 1 */
    int64_t min_val;       /* matrix minimum entry (score given to N rows) */
    int64_t scale;
    double offset;
    double total;          /* sequential sum of pval_mat (denominator of every p-value) */
    int64_t smem_bytes;    /* dynamic shared memory of the scoring kernel for this motif */
    int32_t chunk_bases;   /* bases per lookup: 4 (256-entry tables) or 3 (64-entry tables: long motifs / large spans) */
    int32_t hist_global;   /* 1: the histogram did not fit shared memory and is counted with global atomics */
} gb2_motif_info;

int gb2_abi_version(void);
const char *gb2_error_string(int code);

/* ---- context --------------------------------------------------------------------------- */
/* `stream` is a cudaStream_t to run on (e.g. torch's current stream) or NULL to create one. */
int gb2_ctx_create(int device, void *stream, gb2_ctx **out);
int gb2_ctx_destroy(gb2_ctx *ctx);
int gb2_ctx_set_stream(gb2_ctx *ctx, void *stream);
int gb2_ctx_sync(gb2_ctx *ctx);
const char *gb2_ctx_last_error(const gb2_ctx *ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int64_t gb2_ctx_launch_count(const gb2_ctx *ctx);
int gb2_device_count(void);
int gb2_ctx_sm_count(const gb2_ctx *ctx);

/* ---- multi-GPU: the context owns the NCCL communicator ---------------------------------------------------- */
/* Replaces the Manager-dict funnel and the parent-side merge that precede the q-value step in the reference
 * (score_sequences.py:115-118,171-188,194-198): one process (or host thread) per GPU scores its own shard, then
 *   gb2_allreduce_hist   sums the per-GPU score histograms in place (uint64[n], ncclAllReduce on the context's stream) --
 *                        after it gb2_qvalues_from_hist gives the same, globally exact q-table on every GPU;
 *   gb2_allgather_bytes  concatenates equally sized byte blocks of every rank (d_recv holds world * bytes_per_rank bytes,
 *                        rank r's block at r * bytes_per_rank) -- used to merge the fixed-width hit columns on the device;
 *   gb2_allreduce_max_f64 element-wise maximum of doubles (device-timed durations: "max over ranks").
 * Rendezvous: rank 0 calls gb2_comm_unique_id and hands the 128 bytes to the other ranks by any means (file, MPI,
 * torch.distributed store, ...); then every rank calls gb2_comm_init (collective).  With world == 1 gb2_comm_init needs
 * no id and every collective is a no-op / local copy.  NCCL is bound at run time (libnccl.so.2 as already loaded by the
 * process, else the system library; GB2_NCCL_LIB overrides); GB2_ERR_STATE when it cannot be found. */
int gb2_comm_unique_id(uint8_t *id /* [GB2_NCCL_ID_BYTES] */);
int gb2_comm_init(gb2_ctx *ctx, const uint8_t *id, int rank, int world);
int gb2_comm_destroy(gb2_ctx *ctx);
int gb2_comm_info(const gb2_ctx *ctx, int *rank, int *world);
int gb2_allreduce_hist(gb2_ctx *ctx, uint64_t *d_hist, int64_t n);
int gb2_allreduce_max_f64(gb2_ctx *ctx, double *d_values, int64_t n);
int gb2_allgather_bytes(gb2_ctx *ctx, const void *d_send, void *d_recv, int64_t bytes_per_rank);

/* ---- K1: k-mer encoder ------------------------------------------------------------------ */
/* Replaces the per-row string handling of score_seqs (score_sequences.py:279,286,375-386):
 * n ASCII k-mers of w bytes (row stride `stride` >= w bytes) -> packed uint64 (two per k-mer when w > 32) + N mask.
 * d_counts[0] += rows flagged in the mask, d_counts[1] += rows holding a symbol that is neither
 * ACGTacgt nor 'N' (undefined in the reference; scored like N here). d_counts may be NULL. */
int gb2_encode_kmers(gb2_ctx *ctx, const uint8_t *d_ascii, int64_t n, int w, int64_t stride,
                     uint64_t *d_packed, uint32_t *d_nmask, uint64_t *d_counts);

/* ---- K1b: device-side reader of the `vg find -K w -E` k-mer TSV ---------------------------------------- */
/* Replaces the per-line parsing of score_seqs (score_sequences.py:273-293).  d_text holds the bytes of one or more
 * TSV files (each line: region, k-mer, chr:start(+|-), chr:stop(+|-), haplotype count, ref|non.ref, node path;
 * any run of blanks/tabs separates fields).
 * gb2_tsv_index_lines: byte offset of every non-blank line -> d_line_off[capacity]; *d_n_rows = number of lines found
 *   (offsets beyond `capacity` are not written: call again with a larger buffer when *d_n_rows > capacity).
 *   skip_minus != 0 drops rows whose third field ends in '-' (what --no-reverse does BEFORE scoring and counting,
 *   score_sequences.py:281-282).  n_bytes < 2^31 per call.
 * gb2_tsv_parse_rows: per line -> packed k-mer + N mask (fused K1), start, stop, strand character, haplotype
 *   count, ref code (1 = "ref", 0 = "non.ref", 2 = other), length of the region name and offset of the k-mer
 *   within the line (so the host can slice the two strings of reported rows from its copy of the text).
 *   d_counts[0] += rows masked (N or bad symbol), [1] += rows with a symbol outside ACGTacgtN, [2] += malformed
 *   lines (fewer than six fields, k-mer not exactly w symbols, non-numeric position or count). */
int gb2_tsv_index_lines(gb2_ctx *ctx, const uint8_t *d_text, int64_t n_bytes, int skip_minus, uint64_t *d_line_off,
                        uint64_t capacity, uint64_t *d_n_rows);
int gb2_tsv_parse_rows(gb2_ctx *ctx, const uint8_t *d_text, int64_t n_bytes, const uint64_t *d_line_off, int64_t n_rows,
                       int w, uint64_t *d_packed, uint32_t *d_nmask, int64_t *d_start, int64_t *d_stop, uint8_t *d_strand,
                       int64_t *d_freq, uint8_t *d_ref, uint32_t *d_name_len, uint32_t *d_seq_off, uint64_t *d_counts);

/* ---- K3: batched score-distribution DP --------------------------------------------------- */
/* Replaces comp_pval_mat (motif_processing.pyx:552-603) for n_motifs motifs at once.
 * h_widths[m] = w_m; h_score_mats = concatenated int64[4][w_m] (rows A,C,G,T); h_bgs = [A,C,G,T]
 * per motif; h_out = concatenated float64[1000*w_m+1].  Bit-exact to the reference for any
 * background (accumulation order A,C,G,T; product and sum rounded separately). Host pointers. */
int gb2_pval_dp_batched(gb2_ctx *ctx, int n_motifs, const int32_t *h_widths, const int64_t *h_score_mats,
                        const double *h_bgs, double *h_out);

/* ---- motif upload (+ K4: score -> p-value table) ------------------------------------------ */
/* Takes what Motif carries (motif.py:18): score_matrix int64[4][w] (rows A,C,G,T), pval_matrix
 * float64[1000*w+1], min_val, scale, offset.  Builds the 4-base chunk LUTs and, on the device, the
 * table p[s] = seqsum(pval_mat[s:]) / seqsum(pval_mat) for every reachable s in the reference's
 * summation order (score_sequences.py:390-391).  Host pointers. */
int gb2_motif_create(gb2_ctx *ctx, const int64_t *h_score_matrix, int w, const double *h_pval_mat,
                     int64_t min_val, int64_t scale, double offset, gb2_motif **out);
/* The same for n_motifs motifs at once (a motif collection, BASELINE config 3; replaces the per-motif loop of
 * motif_ops.py:303-335 / 971-1022 on the device side): h_widths[m]; h_score_mats[m] -> int64[4][w_m]; h_pval_mats[m] ->
 * float64[1000*w_m+1] (arrays of n_motifs host pointers: every Motif keeps its own arrays, nothing has to be
 * concatenated) plus min_val / scale / offset per motif.
 * ONE device allocation, ONE upload, TWO kernel launches (K4 for every motif) and ONE synchronisation for the whole
 * collection; out[n_motifs] receives the handles (each destroyed with gb2_motif_destroy; the shared allocation goes
 * with the last one).  On failure no handle is returned. */
int gb2_motif_create_batched(gb2_ctx *ctx, int n_motifs, const int32_t *h_widths, const int64_t *const *h_score_mats,
                             const double *const *h_pval_mats, const int64_t *h_min_vals, const int64_t *h_scales,
                             const double *h_offsets, gb2_motif **out);
int gb2_motif_destroy(gb2_motif *motif);
int gb2_motif_get_info(const gb2_motif *motif, gb2_motif_info *info);
/* copies the p-value table (span doubles, score lo..hi) to the host */
int gb2_motif_get_ptable(gb2_ctx *ctx, const gb2_motif *motif, double *h_ptable);
/* device pointer of the same table (span doubles), owned by the motif */
const double *gb2_motif_ptable_device(const gb2_motif *motif);

/* ---- K2: scoring ---------------------------------------------------------------------------- */
/* Replaces compute_score_seq (score_sequences.py:331-396) and the p-value test of
 * ResultTmp.to_df (resultsTmp.py:303-307) for a batch of packed k-mers.
 *   strands      1: score each k-mer as given; 2: also its reverse complement (what `vg find -E`
 *                emits as the '-' row) from the same 8-byte read.
 *   p_threshold  a window is a hit when p < p_threshold (strict). N rows never hit. With
 *                p_threshold > 1 every window with p <= 1 is a hit (used by tests).
 *   d_hist       uint64[span+1] or NULL (no q-values wanted): += per-score counts.
 *   d_hits / hit_capacity / d_hit_count: hit records appended at *d_hit_count (device counter,
 *                += hits found even when capacity is exceeded; excess records are dropped).
 *   d_dense      uint32[n] or NULL: per k-mer ((rc - lo) << 16 | (fwd - lo)); 0xFFFFFFFF for N rows (8-byte aligned
 *                for the fastest stores).
 * d_packed must be 16-byte aligned (uint64[n], or uint64[n][2] for a motif wider than 32).  row_base is added to the
 * row index in hit records.  Motifs whose score span does not fit shared memory next to the lookup tables (w >~ 54)
 * count their histogram with global atomics instead: same results, slower. */
int gb2_score(gb2_ctx *ctx, const gb2_motif *motif, const uint64_t *d_packed, const uint32_t *d_nmask,
              int64_t n, uint64_t row_base, int strands, double p_threshold, uint64_t *d_hist,
              gb2_hit *d_hits, uint64_t hit_capacity, uint64_t *d_hit_count, uint32_t *d_dense);

/* ---- K5: Benjamini-Hochberg from the score histogram ----------------------------------------- */
/* Replaces compute_qvalues (score_sequences.py:401-428; statsmodels fdr_bh) given the (globally
 * all-reduced) histogram: sorts bins by p ascending, C = cumulative count, q = reverse running
 * minimum of p / (C / float(N)), clipped at 1.
 *   d_qtab   double[span+1]  q-value per histogram bin
 *   d_rank   uint32[span+1]  position of the bin in p-ascending order (sort key for K6)
 *   d_total  uint64[1]       N = number of scored windows */
int gb2_qvalues_from_hist(gb2_ctx *ctx, const gb2_motif *motif, const uint64_t *d_hist, double *d_qtab,
                          uint32_t *d_rank, uint64_t *d_total);

/* The same step for MANY motifs in one launch (one CTA per motif): d_hist / d_qtab / d_rank are the per-motif arrays laid
 * back to back, motif m owning bins [h_bin_off[m], h_bin_off[m+1]) (= span_m + 1 of them); d_totals[n_motifs].  d_hist may
 * be NULL (ranks only).  Motifs with a non-monotone p-value table take the sorting form above, one at a time. */
int gb2_qvalues_from_hist_many(gb2_ctx *ctx, int32_t n_motifs, const gb2_motif *const *motifs, const int64_t *h_bin_off,
                               const uint64_t *d_hist, double *d_qtab, uint32_t *d_rank, uint64_t *d_totals);

/* Stand-alone form of the same step for an arbitrary list of p-values (B3 seam: compute_qvalues takes a
 * list and returns a list, score_sequences.py:401-428): CUB radix sort, raw = p / (k / float(n)), reverse
 * running minimum, clip at 1, scattered back to input order.  Host pointers. */
int gb2_bh_pvalues(gb2_ctx *ctx, const double *h_p, int64_t n, double *h_q);

/* ---- K6: finalize hits -------------------------------------------------------------------------- */
/* Replaces the filter + sort of ResultTmp.to_df (resultsTmp.py:303-313) and the log-odds / p / q
 * columns (score_sequences.py:393; resultsTmp.py:277-279): optional q < q_threshold filter
 * (--qvalueT), CUB radix sort by (p ascending, row ascending, strand) -- a deterministic order; the
 * reference's tie order is undefined -- and the numeric columns.  d_qtab/d_rank may be NULL when no
 * q-values were computed (then the sort uses the score and d_q is not written).
 *   row_limit     exclusive upper bound of the row indices in d_hits (0 = unknown): fewer radix passes.
 *   p_threshold   hits with p >= p_threshold are dropped (hit records produced by gb2_score already satisfy it;
 *                 records expanded from dense scores do not).
 * All outputs have room for n_hits entries; *d_n_out receives the number kept. */
int gb2_finalize_hits(gb2_ctx *ctx, const gb2_motif *motif, const gb2_hit *d_hits, uint64_t n_hits, uint64_t row_limit,
                      const double *d_qtab, const uint32_t *d_rank, double p_threshold, int q_filter,
                      double q_threshold, uint64_t *d_row, uint8_t *d_strand, int32_t *d_iscore, double *d_score,
                      double *d_p, double *d_q, uint64_t *d_n_out);

/* The same step for an unselective scan (`-t 1`, docs/paper_results/run_analysis.sh:43: every window with p < 1 is a
 * report row), from the dense scores gb2_score wrote for n_kmers consecutive k-mers (d_dense, see gb2_score) instead of
 * hit records: window i = (k-mer i / strands, strand i % strands) is already in (row, strand) order, so a stable sort on
 * the <= 16-bit p-rank alone gives the same (p, row, strand) order: two hand-written partition passes (csrc/dense_sort.cu),
 * the first reading the dense scores, the second writing the columns in place.  Row indices = row_base + k-mer index.
 * d_score, d_p and d_q may be NULL (a caller that prints the rows on the device, gb2_report_*, needs only row, strand
 * and integer score).  Outputs need room for n_kmers * strands entries; n_kmers * strands < 2^31. */
int gb2_finalize_dense(gb2_ctx *ctx, const gb2_motif *motif, const uint32_t *d_dense, uint64_t n_kmers, int strands,
                       uint64_t row_base, const double *d_qtab, const uint32_t *d_rank, double p_threshold, int q_filter,
                       double q_threshold, uint64_t *d_row, uint8_t *d_strand, int32_t *d_iscore, double *d_score,
                       double *d_p, double *d_q, uint64_t *d_n_out);

/* The same step for MANY motifs at once (a motif collection scanned over the same k-mers, BASELINE config 3): the hits of
 * all motifs are in one buffer -- gb2_score was called with row_base = (motif index << 40) and one shared hit counter --
 * and are filtered, sorted by (motif, p ascending, row, strand) and annotated by one key kernel, one radix sort and one
 * gather.  motifs / d_qtabs / d_ranks: host arrays of n_motifs pointers (d_qtabs may be NULL without q-values).
 * row_limit: exclusive bound of the row indices (< 2^40).  d_motif receives the motif index of every kept hit; the rows
 * of one motif are contiguous.  motif bits + rank bits + row bits + 1 must fit 63 bits. */
int gb2_finalize_hits_many(gb2_ctx *ctx, int32_t n_motifs, const gb2_motif *const *motifs, const double *const *d_qtabs,
                           const uint32_t *const *d_ranks, const gb2_hit *d_hits, uint64_t n_hits, uint64_t row_limit,
                           double p_threshold, int q_filter, double q_threshold, uint32_t *d_motif, uint64_t *d_row,
                           uint8_t *d_strand, int32_t *d_iscore, double *d_score, double *d_p, double *d_q, uint64_t *d_n_out);

/* ---- haplotype tally ---------------------------------------------------------------------------- */
/* Per-haplotype windows -> vg-like deduplicated rows: sorts (position, packed k-mer) pairs and
 * run-length encodes them (segmented reduction).  Gives what `vg find -E -H gbwt` reports in the
 * frequency and ref columns consumed at score_sequences.py:292-293.
 *   d_pos[n], d_packed[n]      per-haplotype windows (any order; one packed word each, i.e. w <= 32); both arrays are
 *                              sorted in place
 *   d_ref_packed               reference window per position (indexed by pos - pos_base) or NULL
 *   outputs (capacity n): d_u_pos, d_u_packed, d_u_freq (haplotype count), d_u_isref (1 when equal
 *   to the reference window); *d_n_unique = number of distinct rows. */
int gb2_tally_haplotypes(gb2_ctx *ctx, uint64_t *d_pos, uint64_t *d_packed, int64_t n,
                         const uint64_t *d_ref_packed, uint64_t pos_base, int64_t n_ref,
                         uint64_t *d_u_pos, uint64_t *d_u_packed, uint32_t *d_u_freq, uint8_t *d_u_isref,
                         uint64_t *d_n_unique);

/* ---- host-buffer convenience: the whole path in one call ------------------------------------------ */
/* compute_results' numeric core (score_sequences.py:273-321 numeric part, :194-198, resultsTmp.py:
 * 303-313) from HOST memory: h_ascii holds n k-mers of w bytes (stride bytes apart; pinned memory
 * gives full PCIe speed).  Copies in chunks overlapped with compute, encodes, scores, builds the
 * histogram, BH, finalizes, and copies the hit table back.
 *   strands 1|2, p_threshold, q_filter (0/1: threshold applies to q), want_q (0 = --no-qvalue).
 *   Outputs (host, capacity hit_capacity): h_row, h_strand, h_iscore, h_score, h_p, h_q (h_q may be
 *   NULL when want_q == 0); *h_n_hits = rows kept; h_stats[4] = {windows scored, N rows, bad rows,
 *   hits before the q filter}.  Returns GB2_ERR_CAPACITY when hit_capacity was too small. */
int gb2_scan_host(gb2_ctx *ctx, const gb2_motif *motif, const uint8_t *h_ascii, int64_t n, int w,
                  int64_t stride, int strands, double p_threshold, int q_filter, int want_q,
                  uint64_t hit_capacity, uint64_t *h_row, uint8_t *h_strand, int32_t *h_iscore,
                  double *h_score, double *h_p, double *h_q, uint64_t *h_n_hits, uint64_t *h_stats);

/* The same call for k-mers that are already 2-bit packed in host memory (8 bytes per k-mer instead of w; 16 when the
 * motif is wider than 32): h_packed as d_packed of gb2_score, h_nmask the N mask (may be NULL = no N rows).
 * h_stats[1], h_stats[2] are 0 (the caller made the mask). */
int gb2_scan_host_packed(gb2_ctx *ctx, const gb2_motif *motif, const uint64_t *h_packed, const uint32_t *h_nmask,
                         int64_t n, int strands, double p_threshold, int q_filter, int want_q, uint64_t hit_capacity,
                         uint64_t *h_row, uint8_t *h_strand, int32_t *h_iscore, double *h_score, double *h_p, double *h_q,
                         uint64_t *h_n_hits, uint64_t *h_stats);
/* pinned (page-locked) host memory for the h_* inputs of the gb2_scan_host* calls: full PCIe bandwidth, and the copies
 * overlap the kernels.  Any host memory works; pageable memory is staged by the driver. */
int gb2_host_alloc(uint64_t bytes, void **out);
int gb2_host_free(void *ptr);

/* ---- K2 over sequences: windows formed on the device ------------------------------------------------------- */
/* The reference scores one text row per window (score_seqs, score_sequences.py:273-321), so a base of a haplotype is
 * read w times.  When the caller holds the SEQUENCES (every window of every haplotype is to be scored: the workload the
 * headline metric is quoted on), these entry points take them as they are -- 2 bits or one ASCII byte per base -- and
 * form the windows in registers: window i of sequence s is bases [i, i+w) and gets the row index
 * row_base + (windows of sequences 0..s-1) + i (or row_base + h_row_base[s] + i), exactly the index it would have in the
 * expanded list of k-mers given to gb2_score / gb2_scan_host.  Sequences shorter than w have no window.
 *
 * 2-bit layout: sequence s occupies words [h_word_off[s], h_word_off[s] + ceil(h_len[s] / 32)) of d_seq2; base i sits in
 * bits [2(i & 31), 2(i & 31) + 1] of word h_word_off[s] + (i >> 5), A=0 C=1 G=2 T=3 (the packed k-mer layout).
 * d_nbits (may be NULL: no such base) has one uint32 per word: bit (i & 31) set = base i is not A/C/G/T; a window touching
 * such a base is an N row (score = min_val, p = 1, score_sequences.py:376-378; histogram bin `span`).
 *
 * gb2_encode_sequences: ASCII (any case) -> that layout.  Sequence s = bytes [h_text_off[s], h_text_off[s] + h_len[s]) of
 * d_text.  d_counts[0] += bases that are not ACGTacgt, d_counts[1] += those that are not N/n either (undefined in the
 * reference; treated like N).  d_nbits / d_counts may be NULL.
 * gb2_score_sequences: as gb2_score (same histogram, hit records, dense scores indexed by window, thresholds), motif
 * width <= 32.  *h_n_windows (may be NULL) = windows scored per strand. */
int gb2_encode_sequences(gb2_ctx *ctx, const uint8_t *d_text, int64_t text_bytes, int64_t n_seqs, const int64_t *h_text_off,
                         const int64_t *h_len, const int64_t *h_word_off, uint64_t *d_seq2, uint32_t *d_nbits,
                         uint64_t *d_counts);
int gb2_score_sequences(gb2_ctx *ctx, const gb2_motif *motif, const uint64_t *d_seq2, const uint32_t *d_nbits,
                        int64_t n_seqs, const int64_t *h_len, const int64_t *h_word_off, const int64_t *h_row_base,
                        uint64_t row_base, int strands, double p_threshold, uint64_t *d_hist, gb2_hit *d_hits,
                        uint64_t hit_capacity, uint64_t *d_hit_count, uint32_t *d_dense, uint64_t *h_n_windows);
/* gb2_scan_host for sequences in HOST memory.  format 0: h_data = ASCII bytes, h_off[s] = byte offset of sequence s
 * (h_nbits must be NULL); format 1: h_data = 2-bit words (layout above), h_off[s] = word offset, h_nbits optional.
 * Host->device traffic: one byte (format 0) or a quarter byte (format 1) per base, i.e. per window -- against w bytes
 * per window for gb2_scan_host.  The batch is copied in chunks (long sequences are cut into pieces overlapping by w-1
 * bases) while the previous chunk is encoded and scored.  Outputs as gb2_scan_host; h_row = window index as defined
 * above; h_stats = {windows scored (both strands), non-ACGT bases, bases that are neither ACGT nor N, hits before the
 * q filter} (the two base counts are 0 for format 1).
 * Transfer compression (format 0): the end-to-end rate of this call is the PCIe rate, so host threads that would idle
 * during the copy re-code part of the chunks into the 2-bit layout (csrc/host_pack.cpp; AVX-512 / AVX2 / scalar) while
 * the copy engine moves the text of the others; packed chunks cross PCIe at 0.25 byte per base (0.375 when they hold an N).  Nothing is scored on the
 * host and the result does not depend on which chunks were packed.  GB2_HOST_PACK_THREADS sets the thread count (0 = off;
 * default: the host's hardware threads, at most 16; half of them shared out when two ranks of the context's communicator use the host, and
 * none when more than two ranks share the host: the host memory system, not PCIe, limits the copies then).
 * gb2_scan_last_transfer: bytes the last gb2_scan_host* call of this context copied host->device and device->host, and
 * how many of its chunks went as given / were packed on the host (any pointer may be NULL). */
/* The host packer on its own, for callers that keep their sequences 2-bit packed (format 1 above, gb2_score_sequences): n_bases
 * ASCII bases -> ceil(n_bases / 32) words and as many N-bit words, exactly what gb2_encode_sequences writes for them (invalid
 * bases: code 0 + N bit; bases past the end: code 0, no N bit).  h_counts (may be NULL): [0] += bases that are not ACGTacgt,
 * [1] += those that are not N/n either.  No GPU involved; thread-safe (call it from as many threads as there are sequences). */
int gb2_pack_sequence_host(const uint8_t *h_text, int64_t n_bases, uint64_t *h_words, uint32_t *h_nbits, uint64_t *h_counts);
int gb2_scan_last_transfer(const gb2_ctx *ctx, uint64_t *h2d_bytes, uint64_t *d2h_bytes, uint64_t *chunks_as_given,
                           uint64_t *chunks_host_packed);
int gb2_scan_host_sequences(gb2_ctx *ctx, const gb2_motif *motif, int format, const void *h_data, const uint32_t *h_nbits,
                            int64_t n_seqs, const int64_t *h_off, const int64_t *h_len, int strands, double p_threshold,
                            int q_filter, int want_q, uint64_t hit_capacity, uint64_t *h_row, uint8_t *h_strand,
                            int32_t *h_iscore, double *h_score, double *h_p, double *h_q, uint64_t *h_n_hits,
                            uint64_t *h_stats);

/* ---- K7: k-mer extraction from a variation graph ---------------------------------- */
/* Replaces the external `vg find -p REGION -x XG -H GBWT -K w -E` call the reference issues per BED region
 * (src/grafimo/extract_regions.py:180,225,326) and the text parse that follows it (score_sequences.py:273-293).
 * The graph is what `vg construct -r REF -v VCF` + `vg index -G` hold (constructVG.py:332,394-396), as flat host arrays
 * (built by grafimo_b200/vgraph.py):
 *   h_node_off[n_nodes+1]  first base of node i in h_seq (nodes are non-empty; node index = vg node id - 1)
 *   h_seq                  base codes 0..3 = A,C,G,T, 4 = anything else
 *   h_node_a0 / h_node_clamp   a walk that STARTS at base j of node i is reported at min(a0 + j, clamp), one that ENDS
 *                          there at min(a0 + j + 1, clamp) (reference-path coordinates, 0-based, end exclusive)
 *   h_node_flags           bit 0: node lies on the reference path
 *   h_edge_off[n_nodes+1], h_edge_to[n_edges]   out-edges, CSR, targets ascending
 *   h_node_cons / h_edge_cons   row of h_cons_bits with the haplotypes through the node / along the edge, or
 *                          0xFFFFFFFF = every haplotype
 *   h_cons_bits[n_cons][words]  haplotype bit sets (bit h & 31 of word h >> 5), words a multiple of 4
 *   n_hap == 0             no haplotype index: every frequency is reported as 0 (vg find without -H)
 * Host pointers; the arrays are copied. */
int gb2_graph_create(gb2_ctx *ctx, int64_t n_nodes, const uint32_t *h_node_off, const uint8_t *h_seq,
                     const int64_t *h_node_a0, const int64_t *h_node_clamp, const uint8_t *h_node_flags,
                     const uint32_t *h_node_cons, int64_t n_edges, const uint32_t *h_edge_off, const uint32_t *h_edge_to,
                     const uint32_t *h_edge_cons, int32_t n_hap, int32_t words, int64_t n_cons,
                     const uint32_t *h_cons_bits, gb2_graph **out);
/* The same graph built by the library from the inputs of `vg construct` / `vg index -G` (constructVG.py:332,394-396):
 * reference sequence (ASCII, any case, non-ACGT = N), reduced alleles sorted by position (h_var_pos[v] 0-based start,
 * h_var_ref_len[v] reference bases replaced, alternative allele = h_alt[h_alt_off[v] .. h_alt_off[v+1]) ASCII; within a
 * position input order decides node ids and which carried allele wins), phased genotypes as bit sets
 * h_gt_bits[n_variants][words] (bit h = haplotype h carries the alternative allele; NULL = no haplotype index, every
 * frequency 0), nodes chained at max_node_len bases (vg's default 32).  Host pointers. */
int gb2_graph_build(gb2_ctx *ctx, const uint8_t *h_ref, int64_t ref_len, int64_t n_variants, const int64_t *h_var_pos,
                    const int32_t *h_var_ref_len, const int64_t *h_alt_off, const uint8_t *h_alt, int32_t n_hap,
                    int32_t words, const uint32_t *h_gt_bits, int32_t max_node_len, gb2_graph **out);
/* gb2_graph_build for several chromosomes at once (one gb2_graph_input each, fields as the arguments above): the host
 * passes run on up to n_threads worker threads (0 = one per hardware thread), every finished graph is uploaded by the
 * calling thread.  out[n_graphs]; on failure no graph is returned. */
typedef struct gb2_graph_input {
    const uint8_t *h_ref;
    int64_t ref_len, n_variants;
    const int64_t *h_var_pos;
    const int32_t *h_var_ref_len;
    const int64_t *h_alt_off;
    const uint8_t *h_alt;
    int32_t n_hap, words;
    const uint32_t *h_gt_bits;
    int32_t max_node_len, reserved;
} gb2_graph_input;
int gb2_graph_build_batch(gb2_ctx *ctx, int32_t n_graphs, const gb2_graph_input *inputs, int32_t n_threads, gb2_graph **out);
/* The host pass of gb2_graph_build alone (no GPU, no context): sizes of the graph and a 64-bit digest of every array it
 * would upload.  With n_threads > 1 the breakpoints of the chromosome are cut into independent ranges (only where no
 * allele spans or ends) that are built on worker threads and stitched together; the result -- node ids, edge order,
 * numbering of the haplotype sets, hence the digest -- does not depend on n_threads or on chunk_bps (forced range size in
 * breakpoints, 0 = automatic).  gb2_graph_build uses every hardware thread (GB2_BUILD_THREADS overrides).
 * h_stats[6] = nodes, edges, bases, haplotype-set rows, ranges used, digest. */
int gb2_graph_build_stats(const uint8_t *h_ref, int64_t ref_len, int64_t n_variants, const int64_t *h_var_pos,
                          const int32_t *h_var_ref_len, const int64_t *h_alt_off, const uint8_t *h_alt, int32_t n_hap,
                          int32_t words, const uint32_t *h_gt_bits, int32_t max_node_len, int32_t n_threads,
                          int64_t chunk_bps, uint64_t *h_stats);
typedef struct gb2_graph_info {
    int64_t n_nodes, n_edges, n_bases, n_sets; /* n_sets: stored haplotype-set rows */
    int32_t n_hap, words;
} gb2_graph_info;
int gb2_graph_get_info(const gb2_graph *graph, gb2_graph_info *info);
int gb2_graph_destroy(gb2_graph *graph);
/* Pass 1 for n_regions regions [h_start[r], h_stop[r]) at once: counts the w-base walks whose reported start and stop
 * lie inside the region and keeps the row offsets in the graph object.  *h_n_rows = rows gb2_graph_extract will
 * write.  Synchronises the stream.
 * GB2_ERR_CAPACITY: more than 2^24 walks start at one base (variants too dense for this width). */
int gb2_graph_prepare(gb2_ctx *ctx, gb2_graph *graph, int32_t n_regions, const int64_t *h_start, const int64_t *h_stop,
                      int w, uint64_t *h_n_rows);
/* Pass 2 of the prepared query (stream-ordered): rows in (region, first base, depth-first) order, forward strand only
 * (the '-' row vg prints for a walk is its reverse complement with start/stop swapped: score with strands = 2).
 *   d_packed uint64[cap] (uint64[cap][2] when w > 32; 16-byte aligned), d_nmask uint32[ceil(cap/32)], d_start/d_stop int64[cap], d_freq int32[cap]
 *   (haplotypes containing the walk's node sequence), d_isref uint8[cap] (1 = every node on the reference path; the
 *   reference rewrites it when |stop-start| != w, score_sequences.py:305-307), d_region uint32[cap] (index into the
 *   region arrays), optional d_walk uint32[cap][32] ([cap][64] when w > 32) + d_walk_len + d_walk_off (node indices of
 *   the walk, offset of the first base) for writing vg's node-path column; d_counts[0] += rows holding a non-ACGT base. */
int gb2_graph_extract(gb2_ctx *ctx, gb2_graph *graph, uint64_t capacity, uint64_t *d_packed, uint32_t *d_nmask,
                      int64_t *d_start, int64_t *d_stop, int32_t *d_freq, uint8_t *d_isref, uint32_t *d_region,
                      uint32_t *d_walk, uint8_t *d_walk_len, uint8_t *d_walk_off, uint64_t *d_counts);

/* ---- K9: device-side reader of phased VCF text (the input of the graph path) ---------------------------- */
/* Replaces, for the graph path, the VCF reading the reference leaves to `vg construct -v` / `vg index -G -v`
 * (src/grafimo/constructVG.py:332,394-396).  d_text holds whole lines of VCF text; d_line_off the byte offset of every
 * non-blank line (gb2_tsv_index_lines with skip_minus = 0).
 * gb2_vcf_parse_fields, per line: d_kind (0 = header/comment line, 1 = data line, 2 = malformed), d_chrom_len (CHROM
 *   starts at the line start), d_pos (POS as written, 1-based), byte range of REF and of the whole ALT column relative to
 *   the line start, d_n_alts (comma-separated ALT alleles; 0 for "."), d_samples_off (offset of the first sample column,
 *   -1 when the line has no samples or FORMAT does not begin with GT), d_line_len (bytes up to the end of line).
 * gb2_vcf_parse_genotypes, per data line with d_row_base[line] >= 0: rows d_row_base[line] .. +n_alts-1 of d_bits
 *   (uint32 [rows][words]) receive the haplotype bit set of each ALT allele: bit (sample * ploidy + j) is set when the
 *   j-th allele of the sample's call ("a|b", "a/b", "a"; "." = reference) is that allele.  d_counts[0] += lines with more
 *   than 16 ALT alleles (skipped, rows left untouched), d_counts[1] += calls naming an allele that does not exist or a
 *   haplotype >= n_hap. */
int gb2_vcf_parse_fields(gb2_ctx *ctx, const uint8_t *d_text, int64_t n_bytes, const uint64_t *d_line_off, int64_t n_lines,
                         uint8_t *d_kind, int32_t *d_chrom_len, int64_t *d_pos, int32_t *d_ref_off, int32_t *d_ref_len,
                         int32_t *d_alt_off, int32_t *d_alt_len, int32_t *d_n_alts, int32_t *d_samples_off,
                         int32_t *d_line_len);
int gb2_vcf_parse_genotypes(gb2_ctx *ctx, const uint8_t *d_text, int64_t n_bytes, const uint64_t *d_line_off, int64_t n_lines,
                            const int32_t *d_samples_off, const int32_t *d_line_len, const int32_t *d_n_alts,
                            const int64_t *d_row_base, int ploidy, int32_t n_hap, int32_t words, uint32_t *d_bits,
                            uint64_t *d_counts);

/* ---- K8: report files formatted on the device ---------------------------------------- */
/* Replaces `DataFrame.to_csv(sep="\t")` (src/grafimo/res_writer.py:136) and writeGFF3 (res_writer.py:213-303) for
 * device-resident hit columns; byte layout per SURVEY.md 8a (a14).  Score / p-value / q-value text comes from per-bin
 * string tables the host formats once with the reference's own formatters; the kernel writes integers, the k-mer and
 * the literals.  Strings live back to back in d_strings, string k = d_strings[d_string_off[k] .. d_string_off[k+1]).
 * Tables (index of their first string): first_score / first_p / first_q (+ d_bin[row]), first_name / first_chrom
 * (+ d_name[row]; chrom = seqname.split(':')[0]), first_const: motif id, motif name, "ref", "non.ref",
 * "\tgrafimo\tnucleotide_motif\t", "\t.\tName=", ";Alias=", ";ID=", "=-=", ";pvalue==", ";qvalue=", ";sequence==", "=;\n". */
typedef struct gb2_report {
    uint64_t n_rows, index_base;       /* index_base: value of the first row's index column (TSV) */
    int32_t width, layout, want_q, reserved; /* layout 0 = TSV rows, 1 = GFF3 rows; want_q 0 = no q-value column */
    const uint64_t *d_kmer;            /* the k-mer as reported (reverse-complemented already for '-' hits); two words
                                          per row when width > 32 */
    const uint8_t *d_strand;           /* '+' or '-' */
    const int64_t *d_start, *d_stop, *d_freq;
    const uint8_t *d_ref;              /* 1 = "ref", 0 = "non.ref" (after the |stop-start| != w rewrite) */
    const int32_t *d_bin, *d_name;
    const uint8_t *d_strings;
    const uint32_t *d_string_off;
    int32_t first_score, first_p, first_q, first_name, first_chrom, first_const;
} gb2_report;
/* Pass 1: byte length of every row -> exclusive offsets in d_row_off[n_rows + 1]; *h_total_bytes = size of the body
 * (header lines are the caller's).  Synchronises.  Pass 2 writes the rows at those offsets into d_out. */
int gb2_report_measure(gb2_ctx *ctx, const gb2_report *report, uint64_t *d_row_off, uint64_t *h_total_bytes);
int gb2_report_write(gb2_ctx *ctx, const gb2_report *report, const uint64_t *d_row_off, uint8_t *d_out, uint64_t capacity);

#ifdef __cplusplus
}
#endif
#endif /* GRAFIMO_B200_H */
