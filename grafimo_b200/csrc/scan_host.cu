// scan_host.cu -- gb2_scan_host: the numeric core of compute_results
// (src/grafimo/score_sequences.py:273-321,194-198; src/grafimo/resultsTmp.py:303-313) from HOST buffers.
//
// ASCII k-mers are copied in chunks on a copy stream into a double-buffered staging area while the
// previous chunk is encoded (K1) and scored (K2) on the compute stream; then K5 (BH), K6 (finalize) and
// one device->host copy of the hit table.  Host<->device traffic: n*stride bytes in, 33 bytes per kept hit
// out.  Pinned host memory gives full PCIe bandwidth; pageable memory works but is staged by the driver.
#include <algorithm>

#include "internal.cuh"

extern "C" int gb2_scan_host(gb2_ctx *ctx, const gb2_motif *m, const uint8_t *h_ascii, int64_t n, int w, int64_t stride,
                             int strands, double p_threshold, int q_filter, int want_q, uint64_t hit_capacity,
                             uint64_t *h_row, uint8_t *h_strand, int32_t *h_iscore, double *h_score, double *h_p,
                             double *h_q, uint64_t *h_n_hits, uint64_t *h_stats)
{
    if (!ctx || !m) return GB2_ERR_ARG;
    GB2_REQUIRE(ctx, h_n_hits != nullptr, "gb2_scan_host: null hit counter");
    *h_n_hits = 0;
    GB2_REQUIRE(ctx, n >= 0 && w == m->w, "gb2_scan_host: k-mer width %d does not match the motif (%d)", w, m->w);
    GB2_REQUIRE(ctx, stride >= w && stride <= 512, "gb2_scan_host: stride %lld outside [w,512]", (long long)stride);
    GB2_REQUIRE(ctx, strands == 1 || strands == 2, "gb2_scan_host: strands must be 1 or 2");
    GB2_REQUIRE(ctx, !q_filter || want_q, "gb2_scan_host: a q-value threshold needs q-values");
    GB2_REQUIRE(ctx, n == 0 || h_ascii != nullptr, "gb2_scan_host: null k-mer buffer");
    GB2_REQUIRE(ctx, hit_capacity == 0 || (h_row && h_strand && h_iscore && h_score && h_p && (h_q || !want_q)),
                "gb2_scan_host: null output buffer");
    GB2_REQUIRE(ctx, hit_capacity < ((uint64_t)1 << 31), "gb2_scan_host: hit capacity must be below 2^31");
    if (h_stats) h_stats[0] = h_stats[1] = h_stats[2] = h_stats[3] = 0;
    if (n == 0) return GB2_OK;
    GB2_CUDA(ctx, cudaSetDevice(ctx->device));

    const int64_t chunk_rows = std::min<int64_t>(n, (int64_t)1 << 24);  // 16 Mi rows per chunk (multiple of 32)
    const int64_t nb = m->span + 1;
    auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t b_ascii = align((size_t)chunk_rows * (size_t)stride + 64);
    const size_t b_packed = align((size_t)chunk_rows * (w > GB2_NARROW_WIDTH ? 16 : 8));
    const size_t b_mask = align((size_t)gb2_div_up(chunk_rows, 32) * 4);
    const size_t b_hist = align((size_t)nb * 8);
    const size_t b_small = 256;  // counters: [0]=N rows [1]=bad rows [2]=hit count [3]=kept [4]=total
    const size_t b_hits = align((size_t)hit_capacity * sizeof(gb2_hit));
    const size_t b_qtab = align((size_t)nb * 8), b_rank = align((size_t)nb * 4);
    const size_t b_out = align((size_t)hit_capacity * 8) * 4 + align((size_t)hit_capacity * 4) + align((size_t)hit_capacity);
    const size_t total = 2 * b_ascii + b_packed + b_mask + b_hist + b_small + b_hits + b_qtab + b_rank + b_out;

    char *pool = nullptr;
    GB2_CUDA(ctx, cudaMalloc((void **)&pool, total));
    char *q = pool;
    uint8_t *d_ascii[2];
    d_ascii[0] = (uint8_t *)q; q += b_ascii;
    d_ascii[1] = (uint8_t *)q; q += b_ascii;
    uint64_t *d_packed = (uint64_t *)q; q += b_packed;
    uint32_t *d_mask = (uint32_t *)q; q += b_mask;
    uint64_t *d_hist = (uint64_t *)q; q += b_hist;
    uint64_t *d_cnt = (uint64_t *)q; q += b_small;
    gb2_hit *d_hits = (gb2_hit *)q; q += b_hits;
    double *d_qtab = (double *)q; q += b_qtab;
    uint32_t *d_rank = (uint32_t *)q; q += b_rank;
    uint64_t *o_row = (uint64_t *)q; q += align((size_t)hit_capacity * 8);
    double *o_score = (double *)q; q += align((size_t)hit_capacity * 8);
    double *o_p = (double *)q; q += align((size_t)hit_capacity * 8);
    double *o_q = (double *)q; q += align((size_t)hit_capacity * 8);
    int32_t *o_iscore = (int32_t *)q; q += align((size_t)hit_capacity * 4);
    uint8_t *o_strand = (uint8_t *)q;

    int rc = GB2_OK;
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
        SH_CUDA(cudaMemsetAsync(d_hist, 0, b_hist + b_small, ctx->stream));
        // the copy stream must not start before earlier work on the compute stream that may still use the pool
        SH_CUDA(cudaEventRecord(consumed[0], ctx->stream));
        SH_CUDA(cudaEventRecord(consumed[1], ctx->stream));
        int64_t done_rows = 0;
        int buf = 0;
        while (done_rows < n) {
            const int64_t rows = std::min(chunk_rows, n - done_rows);
            const size_t bytes = (size_t)((rows - 1) * stride + w);
            SH_CUDA(cudaStreamWaitEvent(ctx->copy_stream, consumed[buf], 0));
            SH_CUDA(cudaMemcpyAsync(d_ascii[buf], h_ascii + done_rows * stride, bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
            SH_CUDA(cudaEventRecord(copied[buf], ctx->copy_stream));
            SH_CUDA(cudaStreamWaitEvent(ctx->stream, copied[buf], 0));
            rc = gb2_encode_kmers(ctx, d_ascii[buf], rows, w, stride, d_packed, d_mask, d_cnt);
            if (rc != GB2_OK) goto done;
            SH_CUDA(cudaEventRecord(consumed[buf], ctx->stream));
            rc = gb2_score(ctx, m, d_packed, d_mask, rows, (uint64_t)done_rows, strands, p_threshold,
                           want_q ? d_hist : nullptr, d_hits, hit_capacity, d_cnt + 2, nullptr);
            if (rc != GB2_OK) goto done;
            done_rows += rows;
            buf ^= 1;
        }
        rc = gb2_qvalues_from_hist(ctx, m, want_q ? d_hist : nullptr, d_qtab, d_rank, d_cnt + 4);
        if (rc != GB2_OK) goto done;
        SH_CUDA(cudaMemcpyAsync(ctx->h_mail, d_cnt, 5 * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
        SH_CUDA(cudaStreamSynchronize(ctx->stream));
        const uint64_t n_hits = ctx->h_mail[2];
        if (h_stats) {
            h_stats[0] = (uint64_t)n * (uint64_t)strands;
            h_stats[1] = ctx->h_mail[0];
            h_stats[2] = ctx->h_mail[1];
            h_stats[3] = n_hits;
        }
        if (n_hits > hit_capacity) {
            GB2_SET_ERR(ctx, "gb2_scan_host: %llu hits exceed the capacity %llu", (unsigned long long)n_hits,
                        (unsigned long long)hit_capacity);
            *h_n_hits = n_hits;
            rc = GB2_ERR_CAPACITY;
            goto done;
        }
        rc = gb2_finalize_hits(ctx, m, d_hits, n_hits, (uint64_t)n, want_q ? d_qtab : nullptr, d_rank, p_threshold, q_filter,
                               p_threshold, o_row,
                               o_strand, o_iscore, o_score, o_p, want_q ? o_q : nullptr, d_cnt + 3);
        if (rc != GB2_OK) goto done;
        SH_CUDA(cudaMemcpyAsync(ctx->h_mail + 8, d_cnt + 3, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
        SH_CUDA(cudaStreamSynchronize(ctx->stream));
        const uint64_t kept = ctx->h_mail[8];
        *h_n_hits = kept;
        if (kept) {
            SH_CUDA(cudaMemcpyAsync(h_row, o_row, kept * 8, cudaMemcpyDeviceToHost, ctx->stream));
            SH_CUDA(cudaMemcpyAsync(h_strand, o_strand, kept, cudaMemcpyDeviceToHost, ctx->stream));
            SH_CUDA(cudaMemcpyAsync(h_iscore, o_iscore, kept * 4, cudaMemcpyDeviceToHost, ctx->stream));
            SH_CUDA(cudaMemcpyAsync(h_score, o_score, kept * 8, cudaMemcpyDeviceToHost, ctx->stream));
            SH_CUDA(cudaMemcpyAsync(h_p, o_p, kept * 8, cudaMemcpyDeviceToHost, ctx->stream));
            if (want_q) SH_CUDA(cudaMemcpyAsync(h_q, o_q, kept * 8, cudaMemcpyDeviceToHost, ctx->stream));
            SH_CUDA(cudaStreamSynchronize(ctx->stream));
        }
    }
done:
#undef SH_CUDA
    cudaStreamSynchronize(ctx->copy_stream);
    cudaStreamSynchronize(ctx->stream);
    for (int i = 0; i < 2; ++i) {
        if (copied[i]) cudaEventDestroy(copied[i]);
        if (consumed[i]) cudaEventDestroy(consumed[i]);
    }
    cudaFree(pool);
    return rc;
}
