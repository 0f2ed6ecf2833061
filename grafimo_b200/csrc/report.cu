// report.cu -- K8: the report files formatted on the device.
//
// The reference writes its result table with pandas (`DataFrame.to_csv(sep="\t")`, src/grafimo/res_writer.py:136) and a
// per-row Python loop for GFF3 (writeGFF3, src/grafimo/res_writer.py:213-303, row list from utils.dftolist :498-575).
// For an unthresholded scan (`-t 1`, what docs/paper_results/run_analysis.sh:43 runs) every scored window is a report
// row and those writers -- and the DataFrame in front of them -- take a thousand times longer than the scan.  Here the
// bytes of both files are produced from the device-resident hit columns:
//   * score, p-value and q-value take one value per score bin, so their text comes from per-bin string tables the host
//     formats once with the reference's own formatters (repr / round(.,1) / numpy scientific) -- byte-identical output
//     without a float printer on the device;
//   * integers, the k-mer (2-bit packed -> letters) and the fixed literals are written by the kernel;
//   * pass 1 measures every row, an exclusive scan places it, pass 2 writes it.
// Layouts:
//   TSV   <index>\t<motif_id>\t<motif_name>\t<seqname>\t<start>\t<stop>\t<strand>\t<score>\t<p>\t[<q>\t]<kmer>\t<freq>\t<ref>\n
//   GFF3  <chrom>\tgrafimo\tnucleotide_motif\t<lo>\t<hi>\t<score.1>\t<strand>\t.\tName=<id>_<seqname><strand>:<ref>;
//         Alias=<name>;ID=<id>=-=<name>=-=<seqname>;pvalue==<p>;[qvalue=<q>;]sequence==<kmer>=;\n      (lo/hi: start/stop
//         swapped on the '-' strand, res_writer.py:267-272)
#include <cub/cub.cuh>

#include "internal.cuh"

struct ReportView {
    unsigned long long n, index_base;
    int w, layout, want_q;
    const unsigned long long *kmer;  // as reported (already reverse-complemented for '-' hits)
    const uint8_t *strand;           // '+' or '-'
    const long long *start, *stop, *freq;
    const uint8_t *ref;              // 1 = "ref", 0 = "non.ref"
    const int32_t *bin;              // index into the score / p / q string tables
    const int32_t *name;             // index into the sequence-name (and chromosome) table
    const uint8_t *tab;              // all strings back to back
    const uint32_t *tab_off;         // string k = tab[tab_off[k] .. tab_off[k+1])
    int32_t base_score, base_p, base_q, base_name, base_chrom, base_const;  // first string of each table
};

struct CountSink {
    unsigned long long len = 0;
    __device__ __forceinline__ void put(uint8_t) { ++len; }
    __device__ __forceinline__ void put_n(const uint8_t *, uint32_t n) { len += n; }
};

struct WriteSink {
    uint8_t *p;
    __device__ __forceinline__ void put(uint8_t c) { *p++ = c; }
    __device__ __forceinline__ void put_n(const uint8_t *s, uint32_t n)
    {
        for (uint32_t i = 0; i < n; ++i) p[i] = s[i];
        p += n;
    }
};

template <typename Sink>
__device__ __forceinline__ void put_str(Sink &s, const ReportView &v, int32_t k)
{
    const uint32_t a = v.tab_off[k], b = v.tab_off[k + 1];
    s.put_n(v.tab + a, b - a);
}

template <typename Sink>
__device__ __forceinline__ void put_int(Sink &s, long long x)
{
    uint8_t buf[20];
    int n = 0;
    unsigned long long u = x < 0 ? (unsigned long long)(-(x + 1)) + 1ull : (unsigned long long)x;
    do {
        buf[n++] = (uint8_t)('0' + (u % 10ull));
        u /= 10ull;
    } while (u);
    if (x < 0) s.put('-');
    while (n) s.put(buf[--n]);
}

// k-mer of row i as letters; rows wider than 32 bases take two packed words (bases 0..31, 32..w-1)
template <typename Sink>
__device__ __forceinline__ void put_kmer(Sink &s, const unsigned long long *kmer, unsigned long long i, int w)
{
    const bool wide = w > GB2_NARROW_WIDTH;
    unsigned long long x = wide ? kmer[2 * i] : kmer[i];
    for (int k = 0; k < w; ++k) {
        if (k == 32) x = kmer[2 * i + 1];
        const uint32_t c = (uint32_t)(x >> (2 * (k & 31))) & 3u;
        s.put((uint8_t)(c == 0 ? 'A' : c == 1 ? 'C' : c == 2 ? 'G' : 'T'));
    }
}

// constants table: [0] motif id, [1] motif name, [2] "ref", [3] "non.ref"
template <typename Sink>
__device__ __forceinline__ void format_row(Sink &s, const ReportView &v, unsigned long long i)
{
    const int32_t bin = v.bin[i], name = v.name[i];
    const uint8_t strand = v.strand[i];
    const int32_t C = v.base_const;
    if (v.layout == 0) {
        put_int(s, (long long)(v.index_base + i)); s.put('\t');
        put_str(s, v, C + 0); s.put('\t');
        put_str(s, v, C + 1); s.put('\t');
        put_str(s, v, v.base_name + name); s.put('\t');
        put_int(s, v.start[i]); s.put('\t');
        put_int(s, v.stop[i]); s.put('\t');
        s.put(strand); s.put('\t');
        put_str(s, v, v.base_score + bin); s.put('\t');
        put_str(s, v, v.base_p + bin); s.put('\t');
        if (v.want_q) { put_str(s, v, v.base_q + bin); s.put('\t'); }
        put_kmer(s, v.kmer, i, v.w); s.put('\t');
        put_int(s, v.freq[i]); s.put('\t');
        put_str(s, v, C + (v.ref[i] ? 2 : 3));
        s.put('\n');
    } else {
        const long long a = v.start[i], b = v.stop[i];
        put_str(s, v, v.base_chrom + name);
        put_str(s, v, C + 4);  // "\tgrafimo\tnucleotide_motif\t"
        put_int(s, strand == '-' ? b : a); s.put('\t');
        put_int(s, strand == '-' ? a : b); s.put('\t');
        put_str(s, v, v.base_score + bin); s.put('\t');
        s.put(strand);
        put_str(s, v, C + 5);  // "\t.\tName="
        put_str(s, v, C + 0); s.put('_');
        put_str(s, v, v.base_name + name); s.put(strand); s.put(':');
        put_str(s, v, C + (v.ref[i] ? 2 : 3));
        put_str(s, v, C + 6);  // ";Alias="
        put_str(s, v, C + 1);
        put_str(s, v, C + 7);  // ";ID="
        put_str(s, v, C + 0);
        put_str(s, v, C + 8);  // "=-="
        put_str(s, v, C + 1);
        put_str(s, v, C + 8);
        put_str(s, v, v.base_name + name);
        put_str(s, v, C + 9);  // ";pvalue=="
        put_str(s, v, v.base_p + bin);
        if (v.want_q) {
            put_str(s, v, C + 10);  // ";qvalue="
            put_str(s, v, v.base_q + bin);
        }
        put_str(s, v, C + 11);  // ";sequence=="
        put_kmer(s, v.kmer, i, v.w);
        put_str(s, v, C + 12);  // "=;\n"
    }
}

__global__ void __launch_bounds__(256) gb2_report_measure_kernel(const ReportView v, unsigned long long *__restrict__ len)
{
    const unsigned long long i = (unsigned long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= v.n) return;
    CountSink s;
    format_row(s, v, i);
    len[i] = s.len;
}

__global__ void __launch_bounds__(256) gb2_report_write_kernel(const ReportView v, const unsigned long long *__restrict__ off,
                                                               uint8_t *__restrict__ out)
{
    const unsigned long long i = (unsigned long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= v.n) return;
    WriteSink s{out + off[i]};
    format_row(s, v, i);
}

static int make_view(gb2_ctx *ctx, const gb2_report *r, ReportView &v)
{
    GB2_REQUIRE(ctx, r != nullptr, "gb2_report: null description");
    GB2_REQUIRE(ctx, r->layout == 0 || r->layout == 1, "gb2_report: layout must be 0 (TSV) or 1 (GFF3)");
    GB2_REQUIRE(ctx, r->width >= 1 && r->width <= GB2_MAX_WIDTH, "gb2_report: width %d outside [1,%d]", r->width, GB2_MAX_WIDTH);
    GB2_REQUIRE(ctx, r->n_rows == 0 || (r->d_kmer && r->d_strand && r->d_start && r->d_stop && r->d_freq && r->d_ref && r->d_bin &&
                                       r->d_name && r->d_strings && r->d_string_off), "gb2_report: null column");
    v.n = r->n_rows; v.index_base = r->index_base; v.w = r->width; v.layout = r->layout; v.want_q = r->want_q;
    v.kmer = (const unsigned long long *)r->d_kmer; v.strand = r->d_strand;
    v.start = (const long long *)r->d_start; v.stop = (const long long *)r->d_stop; v.freq = (const long long *)r->d_freq;
    v.ref = r->d_ref; v.bin = r->d_bin; v.name = r->d_name; v.tab = r->d_strings; v.tab_off = r->d_string_off;
    v.base_score = r->first_score; v.base_p = r->first_p; v.base_q = r->first_q; v.base_name = r->first_name;
    v.base_chrom = r->first_chrom; v.base_const = r->first_const;
    return GB2_OK;
}

extern "C" int gb2_report_measure(gb2_ctx *ctx, const gb2_report *r, uint64_t *d_row_off, uint64_t *h_total_bytes)
{
    if (!ctx || !h_total_bytes) return GB2_ERR_ARG;
    *h_total_bytes = 0;
    ReportView v;
    int rc = make_view(ctx, r, v);
    if (rc != GB2_OK) return rc;
    if (v.n == 0) return GB2_OK;
    GB2_REQUIRE(ctx, d_row_off != nullptr, "gb2_report_measure: null offsets");
    GB2_REQUIRE(ctx, v.n < ((unsigned long long)1 << 39), "gb2_report_measure: too many rows");
    GB2_CUDA(ctx, cudaSetDevice(ctx->device));
    unsigned long long *off = (unsigned long long *)d_row_off;
    GB2_CUDA(ctx, cudaMemsetAsync(off + v.n, 0, sizeof(unsigned long long), ctx->stream));
    gb2_report_measure_kernel<<<(unsigned)gb2_div_up((int64_t)v.n, 256), 256, 0, ctx->stream>>>(v, off);
    GB2_LAUNCH_CHECK(ctx);
    size_t cub_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, off, off, (int64_t)(v.n + 1), ctx->stream);
    rc = gb2_scratch_reserve(ctx, cub_bytes);
    if (rc != GB2_OK) return rc;
    GB2_CUDA(ctx, cub::DeviceScan::ExclusiveSum(ctx->scratch, cub_bytes, off, off, (int64_t)(v.n + 1), ctx->stream));
    ctx->launches += 1;
    GB2_CUDA(ctx, cudaMemcpyAsync(ctx->h_mail, off + v.n, 8, cudaMemcpyDeviceToHost, ctx->stream));
    GB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *h_total_bytes = ctx->h_mail[0];
    return GB2_OK;
}

extern "C" int gb2_report_write(gb2_ctx *ctx, const gb2_report *r, const uint64_t *d_row_off, uint8_t *d_out, uint64_t capacity)
{
    if (!ctx) return GB2_ERR_ARG;
    ReportView v;
    int rc = make_view(ctx, r, v);
    if (rc != GB2_OK) return rc;
    if (v.n == 0) return GB2_OK;
    GB2_REQUIRE(ctx, d_row_off && d_out, "gb2_report_write: null buffer");
    GB2_CUDA(ctx, cudaSetDevice(ctx->device));
    GB2_CUDA(ctx, cudaMemcpyAsync(ctx->h_mail, (const unsigned long long *)d_row_off + v.n, 8, cudaMemcpyDeviceToHost, ctx->stream));
    GB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->h_mail[0] > capacity) {
        GB2_SET_ERR(ctx, "gb2_report_write: %llu bytes, capacity %llu", (unsigned long long)ctx->h_mail[0], (unsigned long long)capacity);
        return GB2_ERR_CAPACITY;
    }
    gb2_report_write_kernel<<<(unsigned)gb2_div_up((int64_t)v.n, 256), 256, 0, ctx->stream>>>(v, (const unsigned long long *)d_row_off, d_out);
    GB2_LAUNCH_CHECK(ctx);
    return GB2_OK;
}
