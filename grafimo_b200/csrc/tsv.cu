// tsv.cu -- device-side reader of the k-mer TSVs that `vg find -x XG -H GBWT -K w -E -p REGION` writes
// (7 whitespace-separated fields: region, k-mer, chr:start(+|-), chr:stop(+|-), haplotype count, ref|non.ref,
// node path).  Replaces the per-line Python of score_seqs (src/grafimo/score_sequences.py:273-293): the k-mer is
// packed to 2 bits/base right here (fused K1), the numeric metadata (start, stop, strand, frequency, ref flag)
// become side arrays on the device, and only byte offsets of the two string fields are kept so that the host can
// slice names / sequences of the (few) reported rows out of its copy of the text.
//
//   pass A  gb2_tsv_index_lines : byte offsets of the non-empty lines (CUB select over a counting iterator);
//           with `skip_minus` the '-' strand rows are dropped here, BEFORE scoring and counting, which is what
//           --no-reverse does in the reference (score_sequences.py:281-282)
//   pass B  gb2_tsv_parse_rows  : one thread per line walks the first six fields
#include <cub/cub.cuh>

#include "internal.cuh"

__device__ __forceinline__ bool is_ws(uint8_t c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }
__device__ __forceinline__ bool is_eol(uint8_t c) { return c == '\n' || c == 0; }

struct LineStartPred {
    const uint8_t *text;
    int64_t n;
    int skip_minus;
    __device__ __forceinline__ bool operator()(const int64_t &i) const
    {
        if (i > 0 && text[i - 1] != '\n') return false;
        // a line counts when it holds a non-blank character
        int64_t p = i;
        while (p < n && is_ws(text[p])) ++p;
        if (p >= n || is_eol(text[p])) return false;
        if (!skip_minus) return true;
        // strand = last character of the third field
        for (int f = 0; f < 2; ++f) {
            while (p < n && !is_ws(text[p]) && !is_eol(text[p])) ++p;
            while (p < n && is_ws(text[p])) ++p;
        }
        int64_t last = -1;
        while (p < n && !is_ws(text[p]) && !is_eol(text[p])) last = p++;
        return !(last >= 0 && text[last] == '-');
    }
};

extern "C" int gb2_tsv_index_lines(gb2_ctx *ctx, const uint8_t *d_text, int64_t n_bytes, int skip_minus,
                                   uint64_t *d_line_off, uint64_t *d_n_rows)
{
    if (!ctx) return GB2_ERR_ARG;
    GB2_REQUIRE(ctx, n_bytes >= 0 && n_bytes < ((int64_t)1 << 31), "gb2_tsv_index_lines: at most 2^31-1 bytes per call");
    GB2_REQUIRE(ctx, d_n_rows != nullptr, "gb2_tsv_index_lines: null counter");
    GB2_CUDA(ctx, cudaSetDevice(ctx->device));
    GB2_CUDA(ctx, cudaMemsetAsync(d_n_rows, 0, sizeof(uint64_t), ctx->stream));
    if (n_bytes == 0) return GB2_OK;
    GB2_REQUIRE(ctx, d_text && d_line_off, "gb2_tsv_index_lines: null buffer");
    cub::CountingInputIterator<int64_t> it(0);
    LineStartPred pred{d_text, n_bytes, skip_minus};
    size_t bytes = 0;
    cub::DeviceSelect::If(nullptr, bytes, it, (int64_t *)d_line_off, (unsigned long long *)d_n_rows, (int)n_bytes, pred, ctx->stream);
    int rc = gb2_scratch_reserve(ctx, bytes);
    if (rc != GB2_OK) return rc;
    GB2_CUDA(ctx, cub::DeviceSelect::If(ctx->scratch, bytes, it, (int64_t *)d_line_off, (unsigned long long *)d_n_rows,
                                        (int)n_bytes, pred, ctx->stream));
    ctx->launches += 1;
    return GB2_OK;
}

__device__ __forceinline__ uint32_t tsv_base_code(uint32_t c)
{
    const uint32_t u = c & 0xDFu;
    const uint32_t t = (u >> 1) & 3u;
    const uint32_t code = t ^ (t >> 1);
    const bool acgt = (u == 'A') | (u == 'C') | (u == 'G') | (u == 'T');
    return acgt ? code : (c == 'N' ? 4u : 5u);
}

// position token "chr:12345+": value = digits between the first ':' and the last character; strand = last character
__device__ __forceinline__ bool parse_pos(const uint8_t *t, int64_t &p, int64_t n, long long &value, uint8_t &strand)
{
    int64_t colon = -1, b = p;
    while (p < n && !is_ws(t[p]) && !is_eol(t[p])) {
        if (t[p] == ':' && colon < 0) colon = p;
        ++p;
    }
    const int64_t e = p;  // one past the token
    if (colon < 0 || e - colon < 3 || e == b) return false;
    strand = t[e - 1];
    long long v = 0;
    bool neg = false;
    int64_t q = colon + 1;
    if (t[q] == '-' && q + 1 < e - 1) { neg = true; ++q; }
    int64_t stop = e - 1;
    for (int64_t k = q; k < stop; ++k)
        if (t[k] == ':') { stop = k; break; }  // python: split(":")[1]
    if (stop == q) return false;
    for (int64_t k = q; k < stop; ++k) {
        const uint8_t c = t[k];
        if (c < '0' || c > '9') return false;
        v = v * 10 + (c - '0');
    }
    value = neg ? -v : v;
    return true;
}

__global__ void __launch_bounds__(128) gb2_tsv_parse_kernel(const uint8_t *__restrict__ t, int64_t n,
                                                            const unsigned long long *__restrict__ line_off, int64_t n_rows,
                                                            int w, unsigned long long *__restrict__ packed,
                                                            uint32_t *__restrict__ nmask, long long *__restrict__ start,
                                                            long long *__restrict__ stop, uint8_t *__restrict__ strand,
                                                            long long *__restrict__ freq, uint8_t *__restrict__ ref,
                                                            uint32_t *__restrict__ name_len, uint32_t *__restrict__ seq_off,
                                                            unsigned long long *__restrict__ counts)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t flag = 0;  // bit0 masked (N / bad symbol), bit1 bad symbol, bit2 malformed line
    if (r < n_rows) {
        const int64_t b = (int64_t)line_off[r];
        int64_t p = b;
        while (p < n && is_ws(t[p])) ++p;
        const int64_t name_b = p;
        while (p < n && !is_ws(t[p]) && !is_eol(t[p])) ++p;
        name_len[r] = (uint32_t)(p - name_b);
        while (p < n && is_ws(t[p])) ++p;
        seq_off[r] = (uint32_t)(p - b);
        unsigned long long x = 0;
        bool ok = true;
        for (int i = 0; i < w; ++i) {
            const uint8_t c = (p < n) ? t[p] : 0;
            if (is_ws(c) || is_eol(c)) { ok = false; break; }
            const uint32_t code = tsv_base_code(c);
            x |= (unsigned long long)(code & 3u) << (2 * i);
            flag |= (code >= 4u ? 1u : 0u) | (code == 5u ? 2u : 0u);
            ++p;
        }
        if (ok && !(p < n && is_ws(t[p]))) ok = false;  // the k-mer must be exactly w symbols long
        while (p < n && is_ws(t[p])) ++p;
        long long v0 = 0, v1 = 0, fq = 0;
        uint8_t s0 = '?', s1 = '?';
        ok = ok && parse_pos(t, p, n, v0, s0);
        while (p < n && is_ws(t[p])) ++p;
        ok = ok && parse_pos(t, p, n, v1, s1);
        while (p < n && is_ws(t[p])) ++p;
        {   // haplotype count
            const int64_t fb = p;
            while (p < n && !is_ws(t[p]) && !is_eol(t[p])) {
                const uint8_t c = t[p];
                if (c < '0' || c > '9') ok = false;
                fq = fq * 10 + (c - '0');
                ++p;
            }
            if (p == fb) ok = false;
        }
        while (p < n && is_ws(t[p])) ++p;
        uint8_t rf = 2;  // 1 = "ref", 0 = "non.ref", 2 = anything else (host slices the text)
        {
            const int64_t rb = p;
            while (p < n && !is_ws(t[p]) && !is_eol(t[p])) ++p;
            const int64_t len = p - rb;
            if (len == 3 && t[rb] == 'r' && t[rb + 1] == 'e' && t[rb + 2] == 'f') rf = 1;
            else if (len == 7 && t[rb] == 'n' && t[rb + 1] == 'o' && t[rb + 2] == 'n' && t[rb + 3] == '.' && t[rb + 4] == 'r' &&
                     t[rb + 5] == 'e' && t[rb + 6] == 'f') rf = 0;
            if (len == 0) ok = false;
        }
        if (!ok) flag |= 4u;
        packed[r] = (flag & 1u) ? 0ull : x;
        start[r] = v0;
        stop[r] = v1;
        strand[r] = s0;
        freq[r] = fq;
        ref[r] = rf;
    }
    const unsigned m = __ballot_sync(0xFFFFFFFFu, flag & 1u);
    const unsigned mb = __ballot_sync(0xFFFFFFFFu, flag & 2u);
    const unsigned mm = __ballot_sync(0xFFFFFFFFu, flag & 4u);
    if ((threadIdx.x & 31) == 0) {
        if (r < n_rows) nmask[r >> 5] = m;
        if (m) atomicAdd(counts + 0, (unsigned long long)__popc(m));
        if (mb) atomicAdd(counts + 1, (unsigned long long)__popc(mb));
        if (mm) atomicAdd(counts + 2, (unsigned long long)__popc(mm));
    }
}

extern "C" int gb2_tsv_parse_rows(gb2_ctx *ctx, const uint8_t *d_text, int64_t n_bytes, const uint64_t *d_line_off,
                                  int64_t n_rows, int w, uint64_t *d_packed, uint32_t *d_nmask, int64_t *d_start,
                                  int64_t *d_stop, uint8_t *d_strand, int64_t *d_freq, uint8_t *d_ref, uint32_t *d_name_len,
                                  uint32_t *d_seq_off, uint64_t *d_counts)
{
    if (!ctx) return GB2_ERR_ARG;
    GB2_REQUIRE(ctx, n_rows >= 0 && n_bytes >= 0, "gb2_tsv_parse_rows: negative size");
    GB2_REQUIRE(ctx, w >= 1 && w <= GB2_MAX_WIDTH, "gb2_tsv_parse_rows: width %d outside [1,%d]", w, GB2_MAX_WIDTH);
    if (n_rows == 0) return GB2_OK;
    GB2_REQUIRE(ctx, d_text && d_line_off && d_packed && d_nmask && d_start && d_stop && d_strand && d_freq && d_ref &&
                         d_name_len && d_seq_off && d_counts, "gb2_tsv_parse_rows: null buffer");
    GB2_CUDA(ctx, cudaSetDevice(ctx->device));
    const int threads = 128;
    const int64_t blocks = gb2_div_up(n_rows, threads);
    GB2_REQUIRE(ctx, blocks < ((int64_t)1 << 31), "gb2_tsv_parse_rows: too many rows for one launch");
    gb2_tsv_parse_kernel<<<(unsigned)blocks, threads, 0, ctx->stream>>>(
        d_text, n_bytes, (const unsigned long long *)d_line_off, n_rows, w, (unsigned long long *)d_packed, d_nmask,
        (long long *)d_start, (long long *)d_stop, d_strand, (long long *)d_freq, d_ref, d_name_len, d_seq_off,
        (unsigned long long *)d_counts);
    GB2_LAUNCH_CHECK(ctx);
    return GB2_OK;
}
