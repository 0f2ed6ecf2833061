// scan_host.cu -- the numeric core of compute_results (src/grafimo/score_sequences.py:273-321,194-198;
// src/grafimo/resultsTmp.py:303-313) from HOST buffers, three input forms:
//   gb2_scan_host          ASCII k-mers, w bytes per window        (what the reference's TSV rows hold)
//   gb2_scan_host_packed   2-bit packed k-mers, 8 bytes per window (16 for w > 32)
//   gb2_scan_host_sequences (seqscan.cu)  whole sequences, 1 byte (ASCII) or 0.25 byte (2-bit) per window
//
// The input is copied in chunks on a copy stream into a double-buffered staging area while the previous chunk is
// encoded (K1) and scored (K2) on the compute stream; then K5 (BH), K6 (finalize) and one device->host copy of the hit
// table.  Pinned host memory gives full PCIe bandwidth; pageable memory works but is staged by the driver.  The staging
// area lives in the context (grow-only), so repeated calls do not pay cudaMalloc / cudaFree.
#include <algorithm>

#include "internal.cuh"
#include "scan_tail.cuh"

static inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

size_t gb2_scan_tail_bytes(const gb2_motif *m, uint64_t hit_capacity)
{
    const int64_t nb = m->span + 1;
    return al256((size_t)nb * 8) + 256 + al256((size_t)hit_capacity * sizeof(gb2_hit)) + al256((size_t)nb * 8) +
           al256((size_t)nb * 4) + al256((size_t)hit_capacity * 8) * 4 + al256((size_t)hit_capacity * 4) + al256((size_t)hit_capacity);
}

char *gb2_scan_tail_carve(char *q, const gb2_motif *m, uint64_t hit_capacity, gb2_scan_bufs &b)
{
    const int64_t nb = m->span + 1;
    b.hist_and_cnt_bytes = al256((size_t)nb * 8) + 256;
    b.d_hist = (uint64_t *)q; q += al256((size_t)nb * 8);
    b.d_cnt = (uint64_t *)q; q += 256;  // [0]=N rows [1]=bad rows [2]=hit count [3]=kept [4]=total
    b.d_hits = (gb2_hit *)q; q += al256((size_t)hit_capacity * sizeof(gb2_hit));
    b.d_qtab = (double *)q; q += al256((size_t)nb * 8);
    b.d_rank = (uint32_t *)q; q += al256((size_t)nb * 4);
    b.o_row = (uint64_t *)q; q += al256((size_t)hit_capacity * 8);
    b.o_score = (double *)q; q += al256((size_t)hit_capacity * 8);
    b.o_p = (double *)q; q += al256((size_t)hit_capacity * 8);
    b.o_q = (double *)q; q += al256((size_t)hit_capacity * 8);
    b.o_iscore = (int32_t *)q; q += al256((size_t)hit_capacity * 4);
    b.o_strand = (uint8_t *)q; q += al256((size_t)hit_capacity);
    return q;
}

#define TAIL_CUDA(call)                                                                                   \
    do {                                                                                                  \
        cudaError_t e__ = (call);                                                                         \
        if (e__ != cudaSuccess) {                                                                         \
            GB2_SET_ERR(ctx, "%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return GB2_ERR_CUDA;                                                                          \
        }                                                                                                 \
    } while (0)

// K5 + K6 + the copy of the hit table back to the host (everything after the last chunk was scored)
int gb2_scan_tail_finish(gb2_ctx *ctx, const gb2_motif *m, const gb2_scan_bufs &b, uint64_t windows, uint64_t row_limit,
                         double p_threshold, int q_filter, int want_q, uint64_t hit_capacity, const gb2_scan_out &o)
{
    int rc = gb2_qvalues_from_hist(ctx, m, want_q ? b.d_hist : nullptr, b.d_qtab, b.d_rank, b.d_cnt + 4);
    if (rc != GB2_OK) return rc;
    TAIL_CUDA(cudaMemcpyAsync(ctx->h_mail, b.d_cnt, 5 * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    TAIL_CUDA(cudaStreamSynchronize(ctx->stream));
    const uint64_t n_hits = ctx->h_mail[2];
    if (o.h_stats) {
        o.h_stats[0] = windows;
        o.h_stats[1] = ctx->h_mail[0];
        o.h_stats[2] = ctx->h_mail[1];
        o.h_stats[3] = n_hits;
    }
    if (n_hits > hit_capacity) {
        GB2_SET_ERR(ctx, "scan: %llu hits exceed the capacity %llu", (unsigned long long)n_hits, (unsigned long long)hit_capacity);
        *o.h_n_hits = n_hits;
        return GB2_ERR_CAPACITY;
    }
    rc = gb2_finalize_hits(ctx, m, b.d_hits, n_hits, row_limit, want_q ? b.d_qtab : nullptr, b.d_rank, p_threshold, q_filter,
                           p_threshold, b.o_row, b.o_strand, b.o_iscore, b.o_score, b.o_p, want_q ? b.o_q : nullptr, b.d_cnt + 3);
    if (rc != GB2_OK) return rc;
    TAIL_CUDA(cudaMemcpyAsync(ctx->h_mail + 8, b.d_cnt + 3, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    TAIL_CUDA(cudaStreamSynchronize(ctx->stream));
    const uint64_t kept = ctx->h_mail[8];
    *o.h_n_hits = kept;
    ctx->last_d2h_bytes = 6 * sizeof(uint64_t) + kept * (8 + 1 + 4 + 8 + 8 + (want_q ? 8 : 0));
    if (kept) {
        TAIL_CUDA(cudaMemcpyAsync(o.h_row, b.o_row, kept * 8, cudaMemcpyDeviceToHost, ctx->stream));
        TAIL_CUDA(cudaMemcpyAsync(o.h_strand, b.o_strand, kept, cudaMemcpyDeviceToHost, ctx->stream));
        TAIL_CUDA(cudaMemcpyAsync(o.h_iscore, b.o_iscore, kept * 4, cudaMemcpyDeviceToHost, ctx->stream));
        TAIL_CUDA(cudaMemcpyAsync(o.h_score, b.o_score, kept * 8, cudaMemcpyDeviceToHost, ctx->stream));
        TAIL_CUDA(cudaMemcpyAsync(o.h_p, b.o_p, kept * 8, cudaMemcpyDeviceToHost, ctx->stream));
        if (want_q) TAIL_CUDA(cudaMemcpyAsync(o.h_q, b.o_q, kept * 8, cudaMemcpyDeviceToHost, ctx->stream));
        TAIL_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return GB2_OK;
}

int gb2_scan_check_args(gb2_ctx *ctx, const char *who, int strands, int q_filter, int want_q, uint64_t hit_capacity,
                        const gb2_scan_out &o)
{
    GB2_REQUIRE(ctx, o.h_n_hits != nullptr, "%s: null hit counter", who);
    *o.h_n_hits = 0;
    GB2_REQUIRE(ctx, strands == 1 || strands == 2, "%s: strands must be 1 or 2", who);
    GB2_REQUIRE(ctx, !q_filter || want_q, "%s: a q-value threshold needs q-values", who);
    GB2_REQUIRE(ctx, hit_capacity == 0 || (o.h_row && o.h_strand && o.h_iscore && o.h_score && o.h_p && (o.h_q || !want_q)),
                "%s: null output buffer", who);
    GB2_REQUIRE(ctx, hit_capacity < ((uint64_t)1 << 31), "%s: hit capacity must be below 2^31", who);
    if (o.h_stats) o.h_stats[0] = o.h_stats[1] = o.h_stats[2] = o.h_stats[3] = 0;
    return GB2_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// k-mer forms: ASCII (packed_in == 0) or packed words (packed_in == 1)
// ---------------------------------------------------------------------------------------------------------------
static int scan_host_kmers(gb2_ctx *ctx, const gb2_motif *m, const uint8_t *h_in, const uint32_t *h_nmask, int packed_in,
                           int64_t n, int w, int64_t stride, int strands, double p_threshold, int q_filter, int want_q,
                           uint64_t hit_capacity, const gb2_scan_out &o)
{
    const char *who = packed_in ? "gb2_scan_host_packed" : "gb2_scan_host";
    if (!ctx || !m) return GB2_ERR_ARG;
    int rc = gb2_scan_check_args(ctx, who, strands, q_filter, want_q, hit_capacity, o);
    if (rc != GB2_OK) return rc;
    GB2_REQUIRE(ctx, n >= 0 && w == m->w, "%s: k-mer width %d does not match the motif (%d)", who, w, m->w);
    const int64_t kbytes = w > GB2_NARROW_WIDTH ? 16 : 8;
    if (packed_in) stride = kbytes;
    GB2_REQUIRE(ctx, stride >= (packed_in ? kbytes : w) && stride <= 512, "%s: stride %lld outside [w,512]", who, (long long)stride);
    GB2_REQUIRE(ctx, n == 0 || h_in != nullptr, "%s: null k-mer buffer", who);
    if (n == 0) return GB2_OK;
    GB2_CUDA(ctx, cudaSetDevice(ctx->device));

    const int64_t chunk_rows = std::min<int64_t>(n, (int64_t)1 << 24);  // 16 Mi rows per chunk (multiple of 32)
    const size_t b_in = al256((size_t)chunk_rows * (size_t)stride + 64);
    const size_t b_packed = packed_in ? 0 : al256((size_t)chunk_rows * (size_t)kbytes);
    const size_t b_mask = al256((size_t)gb2_div_up(chunk_rows, 32) * 4);
    const size_t total = 2 * b_in + b_packed + 2 * b_mask + gb2_scan_tail_bytes(m, hit_capacity);

    char *q = nullptr;
    rc = gb2_pool_reserve(ctx, total, &q);
    if (rc != GB2_OK) return rc;
    uint8_t *d_in[2];
    d_in[0] = (uint8_t *)q; q += b_in;
    d_in[1] = (uint8_t *)q; q += b_in;
    uint64_t *d_packed = (uint64_t *)q; q += b_packed;
    uint32_t *d_mask[2];
    d_mask[0] = (uint32_t *)q; q += b_mask;
    d_mask[1] = (uint32_t *)q; q += b_mask;
    gb2_scan_bufs b;
    gb2_scan_tail_carve(q, m, hit_capacity, b);

    cudaEvent_t copied[2] = {nullptr, nullptr}, consumed[2] = {nullptr, nullptr};
    cudaError_t e = cudaSuccess;
#define SH_CUDA(call)                                                                                         \
    do {                                                                                                      \
        e = (call);                                                                                           \
        if (e != cudaSuccess) {                                                                               \
            GB2_SET_ERR(ctx, "%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e));      \
            rc = GB2_ERR_CUDA;                                                                                \
            goto done;                                                                                        \
        }                                                                                                     \
    } while (0)
    {
        for (int i = 0; i < 2; ++i) {
            SH_CUDA(cudaEventCreateWithFlags(&copied[i], cudaEventDisableTiming));
            SH_CUDA(cudaEventCreateWithFlags(&consumed[i], cudaEventDisableTiming));
        }
        SH_CUDA(cudaMemsetAsync(b.d_hist, 0, b.hist_and_cnt_bytes, ctx->stream));
        // the copy stream must not start before earlier work on the compute stream that may still use the pool
        SH_CUDA(cudaEventRecord(consumed[0], ctx->stream));
        SH_CUDA(cudaEventRecord(consumed[1], ctx->stream));
        int64_t done_rows = 0;
        int buf = 0;
        ctx->last_h2d_bytes = 0;
        ctx->last_chunks_given = 0;
        ctx->last_chunks_packed = 0;
        while (done_rows < n) {
            const int64_t rows = std::min(chunk_rows, n - done_rows);
            const size_t bytes = packed_in ? (size_t)rows * (size_t)kbytes : (size_t)((rows - 1) * stride + w);
            SH_CUDA(cudaStreamWaitEvent(ctx->copy_stream, consumed[buf], 0));
            SH_CUDA(cudaMemcpyAsync(d_in[buf], h_in + done_rows * stride, bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
            if (packed_in && h_nmask)  // chunk starts are multiples of 32 rows: whole mask words
                SH_CUDA(cudaMemcpyAsync(d_mask[buf], h_nmask + (done_rows >> 5), (size_t)gb2_div_up(rows, 32) * 4,
                                        cudaMemcpyHostToDevice, ctx->copy_stream));
            ctx->last_h2d_bytes += bytes + ((packed_in && h_nmask) ? (size_t)gb2_div_up(rows, 32) * 4 : 0);
            ctx->last_chunks_given++;
            SH_CUDA(cudaEventRecord(copied[buf], ctx->copy_stream));
            SH_CUDA(cudaStreamWaitEvent(ctx->stream, copied[buf], 0));
            const uint64_t *d_kmers = d_packed;
            const uint32_t *d_nm = d_mask[0];
            if (packed_in) {
                d_kmers = (const uint64_t *)d_in[buf];
                d_nm = h_nmask ? d_mask[buf] : nullptr;
            } else {
                rc = gb2_encode_kmers(ctx, d_in[buf], rows, w, stride, d_packed, d_mask[0], b.d_cnt);
                if (rc != GB2_OK) goto done;
                SH_CUDA(cudaEventRecord(consumed[buf], ctx->stream));  // the ASCII buffer is free once encoded
            }
            rc = gb2_score(ctx, m, d_kmers, d_nm, rows, (uint64_t)done_rows, strands, p_threshold,
                           want_q ? b.d_hist : nullptr, b.d_hits, hit_capacity, b.d_cnt + 2, nullptr);
            if (rc != GB2_OK) goto done;
            if (packed_in) SH_CUDA(cudaEventRecord(consumed[buf], ctx->stream));
            done_rows += rows;
            buf ^= 1;
        }
        rc = gb2_scan_tail_finish(ctx, m, b, (uint64_t)n * (uint64_t)strands, (uint64_t)n, p_threshold, q_filter, want_q,
                                  hit_capacity, o);
    }
done:
#undef SH_CUDA
    cudaStreamSynchronize(ctx->copy_stream);
    cudaStreamSynchronize(ctx->stream);
    for (int i = 0; i < 2; ++i) {
        if (copied[i]) cudaEventDestroy(copied[i]);
        if (consumed[i]) cudaEventDestroy(consumed[i]);
    }
    return rc;
}

extern "C" int gb2_scan_host(gb2_ctx *ctx, const gb2_motif *m, const uint8_t *h_ascii, int64_t n, int w, int64_t stride,
                             int strands, double p_threshold, int q_filter, int want_q, uint64_t hit_capacity,
                             uint64_t *h_row, uint8_t *h_strand, int32_t *h_iscore, double *h_score, double *h_p,
                             double *h_q, uint64_t *h_n_hits, uint64_t *h_stats)
{
    gb2_scan_out o = {h_row, h_strand, h_iscore, h_score, h_p, h_q, h_n_hits, h_stats};
    return scan_host_kmers(ctx, m, h_ascii, nullptr, 0, n, w, stride, strands, p_threshold, q_filter, want_q, hit_capacity, o);
}

extern "C" int gb2_scan_host_packed(gb2_ctx *ctx, const gb2_motif *m, const uint64_t *h_packed, const uint32_t *h_nmask,
                                    int64_t n, int strands, double p_threshold, int q_filter, int want_q,
                                    uint64_t hit_capacity, uint64_t *h_row, uint8_t *h_strand, int32_t *h_iscore,
                                    double *h_score, double *h_p, double *h_q, uint64_t *h_n_hits, uint64_t *h_stats)
{
    gb2_scan_out o = {h_row, h_strand, h_iscore, h_score, h_p, h_q, h_n_hits, h_stats};
    return scan_host_kmers(ctx, m, (const uint8_t *)h_packed, h_nmask, 1, n, m ? m->w : 0, 0, strands, p_threshold, q_filter,
                           want_q, hit_capacity, o);
}
