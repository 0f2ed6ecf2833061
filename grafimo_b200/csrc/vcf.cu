// vcf.cu -- K9: device-side reader of phased VCF text (the variants + haplotypes the variation graph is built from).
//
// The reference hands the VCF to external programs (`vg construct -v VCF`, `vg index -G x.gbwt -v VCF`,
// src/grafimo/constructVG.py:332,394-396).  For the graph path (csrc/graph.cu, graph_build.cu) the same text is
// tokenised here: a 1000-Genomes chromosome is ~10^6 lines of ~10 KB (one "a|b" call per sample), far too much for
// per-line host parsing, and all of it is independent byte work.
//
//   gb2_vcf_parse_fields     one thread per line: the eight fixed columns + FORMAT -> kind (header / data / malformed),
//                            POS, byte ranges of CHROM / REF / ALT, number of ALT alleles, where the samples start;
//   gb2_vcf_parse_genotypes  one CTA per data line.  Fixed-shape lines (plain GT, every call "d|d": 1000-Genomes style)
//                            are read as 32-bit call words at computed positions; any other line goes through 4 KB
//                            tiles, tab ordinals by a block scan, every call "a|b" / "a/b" / "a" parsed where its tab is
//                            found.  Bit (sample * ploidy + j) is set in the row of allele a -- the rows are the
//                            haplotype bit sets gb2_graph_build takes (h_gt_bits).
// Line starts come from gb2_tsv_index_lines (csrc/tsv.cu).
#include <cub/cub.cuh>

#include "internal.cuh"

#define VCF_MAX_ALT 16       // ALT alleles per line whose genotype rows are built (more: line is counted and skipped)
#define GT_THREADS 128
#define GT_BYTES_PER_THREAD 32

__device__ __forceinline__ bool vcf_eol(uint8_t c) { return c == '\n' || c == '\r' || c == 0; }

__global__ void __launch_bounds__(128) gb2_vcf_fields_kernel(const uint8_t *__restrict__ text, int64_t n_bytes,
                                                             const unsigned long long *__restrict__ line_off, int64_t n_lines,
                                                             uint8_t *__restrict__ kind, int32_t *__restrict__ chrom_len,
                                                             long long *__restrict__ pos, int32_t *__restrict__ ref_off,
                                                             int32_t *__restrict__ ref_len, int32_t *__restrict__ alt_off,
                                                             int32_t *__restrict__ alt_len, int32_t *__restrict__ n_alts,
                                                             int32_t *__restrict__ samples_off, int32_t *__restrict__ line_len)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_lines) return;
    const int64_t lo = (int64_t)line_off[i];
    const int64_t hi = (i + 1 < n_lines) ? (int64_t)line_off[i + 1] : n_bytes;  // next line start bounds this one
    int64_t p = lo;
    uint8_t k = 1;
    int32_t clen = 0, roff = 0, rlen = 0, aoff = 0, alen = 0, na = 0, soff = -1;
    long long v = 0;
    if (text[lo] == '#') {
        k = 0;
    } else {
        // field f spans [fb, p); fields are TAB separated (VCF 4.x)
        bool gt_first = false;
        for (int f = 0;; ++f) {
            const int64_t fb = p;
            while (p < hi && text[p] != '\t' && !vcf_eol(text[p])) ++p;
            const int64_t fl = p - fb;
            if (f == 0) {
                clen = (int32_t)fl;
                if (fl == 0) k = 2;
            } else if (f == 1) {
                if (fl == 0 || fl > 18) k = 2;
                for (int64_t q = fb; q < p; ++q) {
                    const uint8_t c = text[q];
                    if (c < '0' || c > '9') { k = 2; break; }
                    v = v * 10 + (c - '0');
                }
            } else if (f == 3) {
                roff = (int32_t)(fb - lo); rlen = (int32_t)fl;
                if (fl == 0) k = 2;
            } else if (f == 4) {
                aoff = (int32_t)(fb - lo); alen = (int32_t)fl;
                if (fl == 0) k = 2;
                else if (!(fl == 1 && text[fb] == '.')) {
                    na = 1;
                    for (int64_t q = fb; q < p; ++q) na += text[q] == ',';
                }
            } else if (f == 8) {  // FORMAT: the genotype must be its first key (VCF: "GT must be the first field if present")
                gt_first = fl >= 2 && text[fb] == 'G' && text[fb + 1] == 'T' && (fl == 2 || text[fb + 2] == ':');
            }
            const bool tab = p < hi && text[p] == '\t';
            if (f == 8) {
                if (tab && gt_first) soff = (int32_t)(p + 1 - lo);
                break;
            }
            if (!tab) {
                if (f < 7) k = 2;  // fewer than eight columns; exactly eight = a sites-only line
                break;
            }
            ++p;
        }
    }
    int64_t e = hi;  // the line ends where the next one starts, minus its line terminator(s) and blank lines
    while (e > p && (vcf_eol(text[e - 1]) || text[e - 1] == ' ' || text[e - 1] == '\t')) --e;
    kind[i] = k;
    chrom_len[i] = clen;
    pos[i] = v;
    ref_off[i] = roff; ref_len[i] = rlen; alt_off[i] = aoff; alt_len[i] = alen;
    n_alts[i] = k == 1 ? na : 0;
    samples_off[i] = (k == 1 && soff >= 0) ? soff : -1;
    line_len[i] = (int32_t)(e - lo);
}

// One CTA per line.  Bits are accumulated in shared memory ([n_alts][words]) and written out once.
__global__ void __launch_bounds__(GT_THREADS) gb2_vcf_gt_kernel(const uint8_t *__restrict__ text, int64_t n_bytes,
                                                                const unsigned long long *__restrict__ line_off,
                                                                const int32_t *__restrict__ samples_off,
                                                                const int32_t *__restrict__ line_len,
                                                                const int32_t *__restrict__ n_alts,
                                                                const long long *__restrict__ row_base, int ploidy, int n_hap,
                                                                int words, uint32_t *__restrict__ bits,
                                                                unsigned long long *__restrict__ counts)
{
    extern __shared__ uint32_t rows_s[];  // [min(n_alts, VCF_MAX_ALT)][words]
    typedef cub::BlockScan<uint32_t, GT_THREADS> Scan;
    __shared__ typename Scan::TempStorage tmp;
    const int64_t line = blockIdx.x;
    const long long base = row_base[line];
    if (base < 0) return;
    const int na = n_alts[line];
    if (na <= 0) return;
    if (na > VCF_MAX_ALT) {
        if (threadIdx.x == 0) atomicAdd(counts, 1ull);
        return;
    }
    for (int k = threadIdx.x; k < na * words; k += GT_THREADS) rows_s[k] = 0u;
    __syncthreads();
    const int64_t lo = (int64_t)line_off[line];
    const int soff = samples_off[line];
    if (soff > 0) {
        const int64_t first = lo + soff - 1;  // the TAB in front of sample 0
        const int64_t end = lo + line_len[line];
        uint32_t carry = 0;                   // tabs seen in earlier tiles == index of the next sample
        unsigned long long bad = 0;
        // ---- fixed-shape lines: FORMAT is plain GT and every call is "d|d" (or "d/d", '.' for a digit) -- every line of a
        //      1000-Genomes-style file.  Then call s sits at first + 1 + 4 s: no TAB search, no scan, no barrier per
        //      tile; a thread checks four calls per step as 32-bit words ("0|0\t" is one compare) and only calls that
        //      carry an alternative allele touch shared memory.  Any call of another shape sends the whole line to the
        //      general path below.
        const int n_samples = ploidy == 2 ? n_hap / 2 : 0;
        if (n_samples >= 1 && end - (first + 1) == 4ll * n_samples - 1 && first + 1 + 4ll * n_samples + 16 <= n_bytes) {
            bool ok = true;
            const int64_t q0 = first + 1;
            const int sh = 8 * (int)(q0 & 3);
            const uint32_t *w32 = reinterpret_cast<const uint32_t *>(text + (q0 & ~(int64_t)3));
            auto take = [&](uint32_t rec, int smp, bool last) {
                const uint32_t c0 = rec & 0xFFu, c1 = (rec >> 8) & 0xFFu, c2 = (rec >> 16) & 0xFFu, c3 = rec >> 24;
                const bool d0 = c0 - '0' <= 9u, d2 = c2 - '0' <= 9u;
                if (!((c1 == '|' || c1 == '/') && (d0 || c0 == '.') && (d2 || c2 == '.') && (last || c3 == '\t'))) {
                    ok = false;
                    return;
                }
                const int v0 = d0 ? (int)(c0 - '0') : 0, v1 = d2 ? (int)(c2 - '0') : 0;
                const int h = 2 * smp;
                if (v0 >= 1) {
                    if (v0 <= na) atomicOr(&rows_s[(v0 - 1) * words + (h >> 5)], 1u << (h & 31));
                    else ++bad;
                }
                if (v1 >= 1) {
                    if (v1 <= na) atomicOr(&rows_s[(v1 - 1) * words + ((h + 1) >> 5)], 1u << ((h + 1) & 31));
                    else ++bad;
                }
            };
            const int n_full = n_samples - 1;  // the last call has no TAB behind it: checked on its own
            for (int r0 = 4 * (int)threadIdx.x; r0 < n_full; r0 += 4 * GT_THREADS) {
                uint32_t wv[5];
#pragma unroll
                for (int k = 0; k < 5; ++k) wv[k] = __ldg(w32 + r0 + k);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (r0 + j < n_full) {
                        const uint32_t rec = __funnelshift_r(wv[j], wv[j + 1], sh);
                        if (rec != 0x09307C30u) take(rec, r0 + j, false);  // "0|0\t": nothing to record
                    }
                }
            }
            if (threadIdx.x == 0) {
                const int64_t q = q0 + 4ll * n_full;
                take((uint32_t)text[q] | ((uint32_t)text[q + 1] << 8) | ((uint32_t)text[q + 2] << 16), n_full, true);
            }
            if (__syncthreads_and(ok ? 1 : 0)) {
                if (bad) atomicAdd(counts + 1, bad);
                for (int k = threadIdx.x; k < na * words; k += GT_THREADS) bits[(size_t)base * words + k] = rows_s[k];
                return;
            }
            bad = 0;  // another shape somewhere in the line: start again on the general path
            for (int k = threadIdx.x; k < na * words; k += GT_THREADS) rows_s[k] = 0u;
            __syncthreads();
        }
        // tiles start on a 16-byte boundary so that every thread reads its 32 bytes as two 128-bit loads
        for (int64_t tile = first & ~(int64_t)15; tile < end; tile += GT_THREADS * GT_BYTES_PER_THREAD) {
            const int64_t b0 = tile + (int64_t)threadIdx.x * GT_BYTES_PER_THREAD;
            uint32_t tabs = 0;
            if (b0 < end) {
                if (b0 + GT_BYTES_PER_THREAD <= n_bytes) {
                    const uint4 *v = reinterpret_cast<const uint4 *>(text + b0);
#pragma unroll
                    for (int h = 0; h < GT_BYTES_PER_THREAD / 16; ++h) {
                        const uint4 x = __ldg(v + h);
                        const uint32_t wv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            // per-byte compare with TAB (0xFF where equal), then the four flags gathered into a nibble
                            const uint32_t eq = __vcmpeq4(wv[k], 0x09090909u) & 0x01010101u;
                            const uint32_t m = ((eq * 0x01020408u) >> 24) & 0xFu;
                            tabs |= m << (16 * h + 4 * k);
                        }
                    }
                } else {
                    for (int k = 0; k < GT_BYTES_PER_THREAD; ++k)
                        if (b0 + k < n_bytes && text[b0 + k] == '\t') tabs |= 1u << k;
                }
                // only TABs inside [first, end) count
                if (b0 < first) tabs &= ~((1u << (int)(first - b0)) - 1u);
                if (b0 + GT_BYTES_PER_THREAD > end) tabs &= (end - b0 >= 32) ? ~0u : ((1u << (int)(end - b0)) - 1u);
            }
            uint32_t before, total;
            Scan(tmp).ExclusiveSum((uint32_t)__popc(tabs), before, total);
            uint32_t s = carry + before;
            while (tabs) {
                const int k = __ffs((int)tabs) - 1;
                tabs &= tabs - 1;
                // the call of sample s: alleles separated by '|' or '/', up to the first ':' / TAB / end of line
                int64_t q = b0 + k + 1;
                // common shape "d|d" + terminator: four independent byte loads, no dependent loop
                if (q + 4 <= end || (q + 3 == end)) {
                    const uint8_t c0 = text[q], c1 = text[q + 1], c2 = text[q + 2];
                    const uint8_t c3 = q + 3 < end ? text[q + 3] : (uint8_t)'\n';
                    const bool sep = c1 == '|' || c1 == '/';
                    const bool term = c3 == '\t' || c3 == ':' || vcf_eol(c3);
                    const bool d0 = c0 >= '0' && c0 <= '9', d2 = c2 >= '0' && c2 <= '9';
                    if (sep && term && (d0 || c0 == '.') && (d2 || c2 == '.') && ploidy >= 2) {
                        const int v0 = d0 ? c0 - '0' : 0, v1 = d2 ? c2 - '0' : 0;
                        const long long h = (long long)s * ploidy;
                        if (v0 >= 1) {
                            if (v0 <= na && h < n_hap) atomicOr(&rows_s[(v0 - 1) * words + (int)(h >> 5)], 1u << (h & 31));
                            else ++bad;
                        }
                        if (v1 >= 1) {
                            if (v1 <= na && h + 1 < n_hap) atomicOr(&rows_s[(v1 - 1) * words + (int)((h + 1) >> 5)], 1u << ((h + 1) & 31));
                            else ++bad;
                        }
                        ++s;
                        continue;
                    }
                }
                int j = 0, val = -1;
                while (true) {
                    const uint8_t c = q < end ? text[q] : (uint8_t)'\n';
                    if (c >= '0' && c <= '9') {
                        val = (val < 0 ? 0 : val) * 10 + (c - '0');
                        if (val > 1 << 20) val = 1 << 20;
                    } else {
                        if (val >= 1 && j < ploidy) {
                            const long long h = (long long)s * ploidy + j;
                            if (val <= na && h < n_hap) atomicOr(&rows_s[(val - 1) * words + (int)(h >> 5)], 1u << (h & 31));
                            else ++bad;
                        }
                        if (c != '|' && c != '/') break;
                        ++j;
                        val = -1;
                    }
                    ++q;
                }
                ++s;
            }
            carry += total;
            __syncthreads();  // tmp is reused by the next tile's scan
        }
        if (bad) atomicAdd(counts + 1, bad);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < na * words; k += GT_THREADS) bits[(size_t)base * words + k] = rows_s[k];
}

extern "C" int gb2_vcf_parse_fields(gb2_ctx *ctx, const uint8_t *d_text, int64_t n_bytes, const uint64_t *d_line_off,
                                    int64_t n_lines, uint8_t *d_kind, int32_t *d_chrom_len, int64_t *d_pos,
                                    int32_t *d_ref_off, int32_t *d_ref_len, int32_t *d_alt_off, int32_t *d_alt_len,
                                    int32_t *d_n_alts, int32_t *d_samples_off, int32_t *d_line_len)
{
    if (!ctx) return GB2_ERR_ARG;
    GB2_REQUIRE(ctx, n_lines >= 0 && n_bytes >= 0 && n_bytes < ((int64_t)1 << 31), "gb2_vcf_parse_fields: at most 2^31-1 bytes per call");
    if (n_lines == 0) return GB2_OK;
    GB2_REQUIRE(ctx, d_text && d_line_off && d_kind && d_chrom_len && d_pos && d_ref_off && d_ref_len && d_alt_off && d_alt_len &&
                    d_n_alts && d_samples_off && d_line_len, "gb2_vcf_parse_fields: null buffer");
    GB2_CUDA(ctx, cudaSetDevice(ctx->device));
    gb2_vcf_fields_kernel<<<(unsigned)gb2_div_up(n_lines, 128), 128, 0, ctx->stream>>>(
        d_text, n_bytes, (const unsigned long long *)d_line_off, n_lines, d_kind, d_chrom_len, (long long *)d_pos, d_ref_off,
        d_ref_len, d_alt_off, d_alt_len, d_n_alts, d_samples_off, d_line_len);
    GB2_LAUNCH_CHECK(ctx);
    return GB2_OK;
}

extern "C" int gb2_vcf_parse_genotypes(gb2_ctx *ctx, const uint8_t *d_text, int64_t n_bytes, const uint64_t *d_line_off, int64_t n_lines,
                                       const int32_t *d_samples_off, const int32_t *d_line_len, const int32_t *d_n_alts,
                                       const int64_t *d_row_base, int ploidy, int32_t n_hap, int32_t words,
                                       uint32_t *d_bits, uint64_t *d_counts)
{
    if (!ctx) return GB2_ERR_ARG;
    GB2_REQUIRE(ctx, n_lines >= 0 && n_lines < ((int64_t)1 << 31), "gb2_vcf_parse_genotypes: line count out of range");
    GB2_REQUIRE(ctx, ploidy >= 1 && ploidy <= 8, "gb2_vcf_parse_genotypes: ploidy %d outside [1,8]", ploidy);
    GB2_REQUIRE(ctx, n_hap >= 0 && words >= 1 && (int64_t)words * 32 >= n_hap, "gb2_vcf_parse_genotypes: %d words do not hold %d haplotypes", words, n_hap);
    if (n_lines == 0) return GB2_OK;
    GB2_REQUIRE(ctx, d_text && d_line_off && d_samples_off && d_line_len && d_n_alts && d_row_base && d_bits && d_counts,
                "gb2_vcf_parse_genotypes: null buffer");
    GB2_REQUIRE(ctx, ((uintptr_t)d_text & 15u) == 0, "gb2_vcf_parse_genotypes: the text buffer must be 16-byte aligned");
    GB2_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t smem = (size_t)VCF_MAX_ALT * words * sizeof(uint32_t);
    GB2_REQUIRE(ctx, smem <= (size_t)ctx->max_smem_optin - 4096, "gb2_vcf_parse_genotypes: too many haplotypes for one CTA (%d)", n_hap);
    GB2_CUDA(ctx, cudaFuncSetAttribute(gb2_vcf_gt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gb2_vcf_gt_kernel<<<(unsigned)n_lines, GT_THREADS, smem, ctx->stream>>>(
        d_text, n_bytes, (const unsigned long long *)d_line_off, d_samples_off, d_line_len, d_n_alts, (const long long *)d_row_base, ploidy,
        n_hap, words, d_bits, (unsigned long long *)d_counts);
    GB2_LAUNCH_CHECK(ctx);
    return GB2_OK;
}
