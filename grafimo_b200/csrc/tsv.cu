// tsv.cu -- device-side reader of the k-mer TSVs that `vg find -x XG -H GBWT -K w -E -p REGION` writes
// (7 whitespace-separated fields: region, k-mer, chr:start(+|-), chr:stop(+|-), haplotype count, ref|non.ref,
// node path).  Replaces the per-line Python of score_seqs (src/grafimo/score_sequences.py:273-293): the k-mer is
// packed to 2 bits/base right here (fused K1), the numeric metadata (start, stop, strand, frequency, ref flag)
// become side arrays on the device, and only byte offsets of the two string fields are kept so that the host can
// slice names / sequences of the (few) reported rows out of its copy of the text.
//
//   pass A  gb2_tsv_index_lines : byte offsets of the non-empty lines (count per 16 KB tile, scan, ordered write);
//           with `skip_minus` the '-' strand rows are dropped here, BEFORE scoring and counting, which is what
//           --no-reverse does in the reference (score_sequences.py:281-282)
//   pass B  gb2_tsv_parse_rows  : one thread per line walks the first six fields
#include <cub/cub.cuh>

#include "internal.cuh"

__device__ __forceinline__ bool is_ws(uint8_t c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }
__device__ __forceinline__ bool is_eol(uint8_t c) { return c == '\n' || c == 0; }

// Byte view of the text that is either the global buffer or a shared-memory copy of [lo, hi).
struct TextView {
    const uint8_t *p;  // p[i - lo] is byte i
    int64_t lo, n;     // n = total bytes of the text (absolute end)
    __device__ __forceinline__ uint8_t operator[](int64_t i) const { return p[i - lo]; }
};

// A line counts when it holds a non-blank character; with skip_minus also when its third field does not end in '-'.
__device__ __forceinline__ bool line_counts(const uint8_t *__restrict__ text, int64_t n, int64_t i, int skip_minus)
{
    int64_t p = i;
    while (p < n && is_ws(text[p])) ++p;
    if (p >= n || is_eol(text[p])) return false;
    if (!skip_minus) return true;
    for (int f = 0; f < 2; ++f) {
        while (p < n && !is_ws(text[p]) && !is_eol(text[p])) ++p;
        while (p < n && is_ws(text[p])) ++p;
    }
    int64_t last = -1;
    while (p < n && !is_ws(text[p]) && !is_eol(text[p])) last = p++;
    return !(last >= 0 && text[last] == '-');
}

#define IDX_THREADS 256
#define IDX_BYTES_PER_THREAD 64

// Line starts inside this thread's 64 contiguous bytes, as a 64-bit mask.
__device__ __forceinline__ unsigned long long line_start_mask(const uint8_t *__restrict__ text, int64_t n, int64_t base,
                                                              int skip_minus)
{
    unsigned long long mask = 0;
    if (base >= n) return 0;
    uint8_t prev = base > 0 ? text[base - 1] : (uint8_t)'\n';
    if (base + IDX_BYTES_PER_THREAD <= n && ((uintptr_t)(text + base) & 15) == 0) {
#pragma unroll
        for (int v = 0; v < IDX_BYTES_PER_THREAD / 16; ++v) {
            const uint4 q = __ldg(reinterpret_cast<const uint4 *>(text + base) + v);
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const uint8_t c = (uint8_t)(w[k >> 2] >> (8 * (k & 3)));
                if (prev == '\n' && c != '\n') mask |= 1ull << (v * 16 + k);
                prev = c;
            }
        }
    } else {
        for (int k = 0; k < IDX_BYTES_PER_THREAD && base + k < n; ++k) {
            const uint8_t c = text[base + k];
            if (prev == '\n' && c != '\n') mask |= 1ull << k;
            prev = c;
        }
    }
    // rare per-candidate check (blank lines, --no-reverse)
    unsigned long long m = mask;
    while (m) {
        const int k = __ffsll((long long)m) - 1;
        m &= m - 1;
        if (!line_counts(text, n, base + k, skip_minus)) mask &= ~(1ull << k);
    }
    return mask;
}

__global__ void __launch_bounds__(IDX_THREADS) gb2_tsv_count_kernel(const uint8_t *__restrict__ text, int64_t n, int skip_minus,
                                                                   uint32_t *__restrict__ block_counts)
{
    typedef cub::BlockReduce<uint32_t, IDX_THREADS> Reduce;
    __shared__ typename Reduce::TempStorage tmp;
    const int64_t base = ((int64_t)blockIdx.x * IDX_THREADS + threadIdx.x) * IDX_BYTES_PER_THREAD;
    const uint32_t c = (uint32_t)__popcll(line_start_mask(text, n, base, skip_minus));
    const uint32_t total = Reduce(tmp).Sum(c);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = total;
}

__global__ void __launch_bounds__(IDX_THREADS) gb2_tsv_offsets_kernel(const uint8_t *__restrict__ text, int64_t n, int skip_minus,
                                                                     const uint32_t *__restrict__ block_base, uint32_t n_blocks,
                                                                     unsigned long long *__restrict__ line_off,
                                                                     unsigned long long capacity,
                                                                     unsigned long long *__restrict__ n_rows)
{
    typedef cub::BlockScan<uint32_t, IDX_THREADS> Scan;
    __shared__ typename Scan::TempStorage tmp;
    const int64_t base = ((int64_t)blockIdx.x * IDX_THREADS + threadIdx.x) * IDX_BYTES_PER_THREAD;
    unsigned long long mask = line_start_mask(text, n, base, skip_minus);
    const uint32_t c = (uint32_t)__popcll(mask);
    uint32_t excl, total;
    Scan(tmp).ExclusiveSum(c, excl, total);
    unsigned long long dst = (unsigned long long)block_base[blockIdx.x] + excl;
    while (mask) {
        const int k = __ffsll((long long)mask) - 1;
        mask &= mask - 1;
        if (dst < capacity) line_off[dst] = (unsigned long long)(base + k);
        ++dst;
    }
    if (blockIdx.x == n_blocks - 1 && threadIdx.x == 0) *n_rows = (unsigned long long)block_base[blockIdx.x] + total;
}

extern "C" int gb2_tsv_index_lines(gb2_ctx *ctx, const uint8_t *d_text, int64_t n_bytes, int skip_minus,
                                   uint64_t *d_line_off, uint64_t capacity, uint64_t *d_n_rows)
{
    if (!ctx) return GB2_ERR_ARG;
    GB2_REQUIRE(ctx, n_bytes >= 0 && n_bytes < ((int64_t)1 << 31), "gb2_tsv_index_lines: at most 2^31-1 bytes per call");
    GB2_REQUIRE(ctx, d_n_rows != nullptr, "gb2_tsv_index_lines: null counter");
    GB2_CUDA(ctx, cudaSetDevice(ctx->device));
    GB2_CUDA(ctx, cudaMemsetAsync(d_n_rows, 0, sizeof(uint64_t), ctx->stream));
    if (n_bytes == 0) return GB2_OK;
    GB2_REQUIRE(ctx, d_text && d_line_off, "gb2_tsv_index_lines: null buffer");
    const int64_t per_block = (int64_t)IDX_THREADS * IDX_BYTES_PER_THREAD;
    const uint32_t n_blocks = (uint32_t)gb2_div_up(n_bytes, per_block);
    size_t cub_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)n_blocks, ctx->stream);
    auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
    int rc = gb2_scratch_reserve(ctx, align(cub_bytes) + 2 * align((size_t)n_blocks * 4));
    if (rc != GB2_OK) return rc;
    char *base = (char *)ctx->scratch;
    void *d_tmp = base; base += align(cub_bytes);
    uint32_t *counts = (uint32_t *)base; base += align((size_t)n_blocks * 4);
    uint32_t *bases = (uint32_t *)base;
    gb2_tsv_count_kernel<<<n_blocks, IDX_THREADS, 0, ctx->stream>>>(d_text, n_bytes, skip_minus, counts);
    GB2_LAUNCH_CHECK(ctx);
    GB2_CUDA(ctx, cub::DeviceScan::ExclusiveSum(d_tmp, cub_bytes, counts, bases, (int)n_blocks, ctx->stream));
    ctx->launches += 1;
    gb2_tsv_offsets_kernel<<<n_blocks, IDX_THREADS, 0, ctx->stream>>>(d_text, n_bytes, skip_minus, bases, n_blocks,
                                                                     (unsigned long long *)d_line_off,
                                                                     (unsigned long long)capacity,
                                                                     (unsigned long long *)d_n_rows);
    GB2_LAUNCH_CHECK(ctx);
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
__device__ __forceinline__ bool parse_pos(const TextView &t, int64_t &p, long long &value, uint8_t &strand)
{
    const int64_t n = t.n;
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

#define PARSE_ROWS 128
#define PARSE_SMEM (40 * 1024)

// One thread per line.  The CTA first copies the contiguous byte range of its 128 lines into shared memory with
// coalesced 16-byte loads (lines are ~100 bytes at arbitrary offsets: per-thread global byte loads would touch a
// different cache line per lane); ranges that do not fit (very long node paths) are parsed from global memory.
__global__ void __launch_bounds__(PARSE_ROWS) gb2_tsv_parse_kernel(const uint8_t *__restrict__ text, int64_t n,
                                                                  const unsigned long long *__restrict__ line_off,
                                                                  int64_t n_rows, int w, unsigned long long *__restrict__ packed,
                                                                  uint32_t *__restrict__ nmask, long long *__restrict__ start,
                                                                  long long *__restrict__ stop, uint8_t *__restrict__ strand,
                                                                  long long *__restrict__ freq, uint8_t *__restrict__ ref,
                                                                  uint32_t *__restrict__ name_len, uint32_t *__restrict__ seq_off,
                                                                  unsigned long long *__restrict__ counts)
{
    __shared__ __align__(16) uint8_t stage[PARSE_SMEM + 32];
    const int64_t r0 = (int64_t)blockIdx.x * PARSE_ROWS;
    const int64_t r = r0 + threadIdx.x;
    const int64_t seg_lo = (int64_t)line_off[r0];
    const int64_t r_end = min(r0 + (int64_t)PARSE_ROWS, n_rows);
    const int64_t seg_hi = r_end < n_rows ? (int64_t)line_off[r_end] : n;
    TextView t;
    t.n = n;
    const int64_t alo = seg_lo & ~(int64_t)15;
    if (seg_hi - alo <= PARSE_SMEM && ((uintptr_t)text & 15) == 0) {
        const int nvec = (int)((seg_hi - alo + 15) >> 4);
        for (int i = threadIdx.x; i < nvec; i += PARSE_ROWS) {
            const int64_t a = alo + ((int64_t)i << 4);
            uint4 v;
            if (a + 16 <= n) {
                v = __ldg(reinterpret_cast<const uint4 *>(text + a));
            } else {
                uint8_t tmp[16];
                for (int b = 0; b < 16; ++b) tmp[b] = (a + b < n) ? text[a + b] : (uint8_t)'\n';
                v = *reinterpret_cast<uint4 *>(tmp);
            }
            reinterpret_cast<uint4 *>(stage)[i] = v;
        }
        __syncthreads();
        t.p = stage;
        t.lo = alo;
        t.n = min(n, alo + ((int64_t)nvec << 4));  // bytes beyond the staged range belong to other CTAs' lines
    } else {
        t.p = text;
        t.lo = 0;
    }
    uint32_t flag = 0;  // bit0 masked (N / bad symbol), bit1 bad symbol, bit2 malformed line
    if (r < n_rows) {
        const int64_t tn = t.n;
        const int64_t b = (int64_t)line_off[r];
        int64_t p = b;
        while (p < tn && is_ws(t[p])) ++p;
        const int64_t name_b = p;
        while (p < tn && !is_ws(t[p]) && !is_eol(t[p])) ++p;
        name_len[r] = (uint32_t)(p - name_b);
        while (p < tn && is_ws(t[p])) ++p;
        seq_off[r] = (uint32_t)(p - b);
        unsigned long long x = 0, x2 = 0;  // bases 0..31 / 32..w-1 (wide k-mers: two packed words per row)
        bool ok = true;
        for (int i = 0; i < w; ++i) {
            const uint8_t c = (p < tn) ? t[p] : 0;
            if (is_ws(c) || is_eol(c)) { ok = false; break; }
            const uint32_t code = tsv_base_code(c);
            if (i < 32) x |= (unsigned long long)(code & 3u) << (2 * i);
            else x2 |= (unsigned long long)(code & 3u) << (2 * (i - 32));
            flag |= (code >= 4u ? 1u : 0u) | (code == 5u ? 2u : 0u);
            ++p;
        }
        if (ok && !(p < tn && is_ws(t[p]))) ok = false;  // the k-mer must be exactly w symbols long
        while (p < tn && is_ws(t[p])) ++p;
        long long v0 = 0, v1 = 0, fq = 0;
        uint8_t s0 = '?', s1 = '?';
        ok = ok && parse_pos(t, p, v0, s0);
        while (p < tn && is_ws(t[p])) ++p;
        ok = ok && parse_pos(t, p, v1, s1);
        while (p < tn && is_ws(t[p])) ++p;
        {   // haplotype count
            const int64_t fb = p;
            while (p < tn && !is_ws(t[p]) && !is_eol(t[p])) {
                const uint8_t c = t[p];
                if (c < '0' || c > '9') ok = false;
                fq = fq * 10 + (c - '0');
                ++p;
            }
            if (p == fb) ok = false;
        }
        while (p < tn && is_ws(t[p])) ++p;
        uint8_t rf = 2;  // 1 = "ref", 0 = "non.ref", 2 = anything else (host slices the text)
        {
            const int64_t rb = p;
            while (p < tn && !is_ws(t[p]) && !is_eol(t[p])) ++p;
            const int64_t len = p - rb;
            if (len == 3 && t[rb] == 'r' && t[rb + 1] == 'e' && t[rb + 2] == 'f') rf = 1;
            else if (len == 7 && t[rb] == 'n' && t[rb + 1] == 'o' && t[rb + 2] == 'n' && t[rb + 3] == '.' && t[rb + 4] == 'r' &&
                     t[rb + 5] == 'e' && t[rb + 6] == 'f') rf = 0;
            if (len == 0) ok = false;
        }
        if (!ok) flag |= 4u;
        if (w > GB2_NARROW_WIDTH) {
            packed[2 * r] = (flag & 1u) ? 0ull : x;
            packed[2 * r + 1] = (flag & 1u) ? 0ull : x2;
        } else {
            packed[r] = (flag & 1u) ? 0ull : x;
        }
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
    const int threads = PARSE_ROWS;
    const int64_t blocks = gb2_div_up(n_rows, threads);
    GB2_REQUIRE(ctx, blocks < ((int64_t)1 << 31), "gb2_tsv_parse_rows: too many rows for one launch");
    gb2_tsv_parse_kernel<<<(unsigned)blocks, threads, 0, ctx->stream>>>(
        d_text, n_bytes, (const unsigned long long *)d_line_off, n_rows, w, (unsigned long long *)d_packed, d_nmask,
        (long long *)d_start, (long long *)d_stop, d_strand, (long long *)d_freq, d_ref, d_name_len, d_seq_off,
        (unsigned long long *)d_counts);
    GB2_LAUNCH_CHECK(ctx);
    return GB2_OK;
}
