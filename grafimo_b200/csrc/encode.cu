// encode.cu -- K1: ASCII k-mers -> 2-bit packed uint64 + N mask.
//
// Replaces the per-row string handling of score_seqs / compute_score_seq
// (src/grafimo/score_sequences.py:279,286,375-386): letters are case-insensitive, 'N' marks the
// row as "score = min_val"; any other symbol is undefined in the reference and is flagged (and
// masked like N) here.
//
// A CTA stages the contiguous byte range of its 256 rows in shared memory with 16-byte loads
// (rows are w bytes, usually an odd stride, so per-thread byte loads from HBM would waste sectors),
// then each thread packs its own row from shared memory.  One warp ballot produces the mask word.
#include "internal.cuh"

#define ENC_ROWS 256

__device__ __forceinline__ uint32_t base_code(uint32_t c)
{
    // 0..3 = A,C,G,T (either case), 4 = 'N', 5 = anything else
    const uint32_t u = c & 0xDFu;  // upper-case
    const uint32_t t = (u >> 1) & 3u;
    const uint32_t code = t ^ (t >> 1);  // A->0 C->1 G->2 T->3
    const bool acgt = (u == 'A') | (u == 'C') | (u == 'G') | (u == 'T');
    return acgt ? code : (c == 'N' ? 4u : 5u);
}

// WIDE: 32 < w <= 64, two packed words per row (bases 0..31, 32..w-1)
template <bool WIDE>
__global__ void __launch_bounds__(ENC_ROWS) gb2_encode_kernel(const uint8_t *__restrict__ ascii, int64_t n, int w,
                                                             int64_t stride, uint64_t *__restrict__ packed,
                                                             uint32_t *__restrict__ nmask,
                                                             unsigned long long *__restrict__ counts)
{
    extern __shared__ __align__(16) uint8_t stage[];
    const int tid = threadIdx.x;
    const int64_t row0 = (int64_t)blockIdx.x * ENC_ROWS;
    const int64_t rows = min((int64_t)ENC_ROWS, n - row0);
    // byte range of this CTA's rows, widened to 16-byte alignment
    const uintptr_t gbeg = (uintptr_t)ascii + (uintptr_t)(row0 * stride);
    const uintptr_t gend = gbeg + (uintptr_t)((rows - 1) * stride + w);
    const uintptr_t abeg = gbeg & ~(uintptr_t)15;
    const int head = (int)(gbeg - abeg);
    const int nvec = (int)((gend - abeg + 15) >> 4);
    // the last vector may run past the caller's buffer by up to 15 bytes; fall back to bytes there
    const uintptr_t buf_end = (uintptr_t)ascii + (uintptr_t)((n - 1) * stride + w);
    for (int i = tid; i < nvec; i += ENC_ROWS) {
        const uintptr_t a = abeg + ((uintptr_t)i << 4);
        if (a >= (uintptr_t)ascii && a + 16 <= buf_end) {
            reinterpret_cast<uint4 *>(stage)[i] = __ldg(reinterpret_cast<const uint4 *>(a));
        } else {
            for (int b = 0; b < 16; ++b) {
                const uintptr_t q = a + b;
                stage[(i << 4) + b] = (q >= (uintptr_t)ascii && q < buf_end) ? __ldg(reinterpret_cast<const uint8_t *>(q)) : (uint8_t)'A';
            }
        }
    }
    __syncthreads();

    uint64_t x = 0;
    uint32_t flag = 0;  // bit0: masked (N or bad), bit1: bad symbol
    if (tid < rows) {
        // Four symbols at a time, SIMD within a 32-bit register: 2-bit codes from bits 1-2 of each byte
        // (A,C,G,T -> 0,1,3,2 -> 0,1,2,3), one multiply gathers the four codes into a byte, and the expected
        // letter is rebuilt from the code and compared with the upper-cased input to catch anything else.
        const uint32_t addr = (uint32_t)head + (uint32_t)tid * (uint32_t)stride;
        const uint32_t *sw = reinterpret_cast<const uint32_t *>(stage) + (addr >> 2);
        const uint32_t sel = 0x3210u + 0x1111u * (addr & 3u);
        const int nwords = (w + 3) >> 2;
        uint32_t lo = 0, hi = 0, lo2 = 0, hi2 = 0, bad = 0;
        uint32_t cur = sw[0];
#pragma unroll
        for (int j = 0; j < (WIDE ? 16 : 8); ++j) {
            if (j < nwords) {
                const uint32_t nxt = sw[j + 1];
                uint32_t v = __byte_perm(cur, nxt, sel);  // bytes 4j .. 4j+3 of the row
                cur = nxt;
                const int rem = w - 4 * j;                // symbols left, >= 1
                if (rem < 4) {
                    const uint32_t keep = 0xFFFFFFFFu >> (8 * (4 - rem));
                    v = (v & keep) | (0x41414141u & ~keep);  // pad with 'A' (code 0)
                }
                const uint32_t u = v & 0xDFDFDFDFu;
                const uint32_t t = (v >> 1) & 0x03030303u;
                const uint32_t c = t ^ ((t >> 1) & 0x01010101u);
                const uint32_t ge2 = (c >> 1) & 0x01010101u;
                const uint32_t eq3 = ge2 & c;
                const uint32_t letters = 0x41414141u + 2u * c + 2u * ge2 + 11u * eq3;  // A C G T = 41 43 47 54
                bad |= u ^ letters;
                const uint32_t four = (c * 0x01041040u) >> 24;  // codes of the 4 symbols in 8 bits
                if (j < 4) lo |= four << (8 * j);
                else if (j < 8) hi |= four << (8 * (j - 4));
                else if (j < 12) lo2 |= four << (8 * (j - 8));
                else hi2 |= four << (8 * (j - 12));
            }
        }
        x = ((uint64_t)hi << 32) | lo;
        uint64_t x2 = ((uint64_t)hi2 << 32) | lo2;
        if (bad) {  // rare: find out whether it is an N or a symbol the reference does not define
            const uint8_t *sb = stage + addr;
            for (int i = 0; i < w; ++i) {
                const uint32_t code = base_code(sb[i]);
                flag |= (code >= 4u ? 1u : 0u) | (code == 5u ? 2u : 0u);
            }
            if (flag) x = x2 = 0;
        }
        if (WIDE) reinterpret_cast<ulonglong2 *>(packed)[row0 + tid] = make_ulonglong2(x, x2);
        else packed[row0 + tid] = x;
    }
    const unsigned m = __ballot_sync(0xFFFFFFFFu, flag & 1u);
    const unsigned mb = __ballot_sync(0xFFFFFFFFu, flag & 2u);
    if ((tid & 31) == 0) {
        const int64_t wrow = row0 + tid;
        if (wrow < n) nmask[wrow >> 5] = m;
        if (counts) {
            if (m) atomicAdd(counts + 0, (unsigned long long)__popc(m));
            if (mb) atomicAdd(counts + 1, (unsigned long long)__popc(mb));
        }
    }
}

extern "C" int gb2_encode_kmers(gb2_ctx *ctx, const uint8_t *d_ascii, int64_t n, int w, int64_t stride,
                                uint64_t *d_packed, uint32_t *d_nmask, uint64_t *d_counts)
{
    if (!ctx) return GB2_ERR_ARG;
    GB2_REQUIRE(ctx, n >= 0, "gb2_encode_kmers: negative row count");
    GB2_REQUIRE(ctx, w >= 1 && w <= GB2_MAX_WIDTH, "gb2_encode_kmers: width %d outside [1,%d]", w, GB2_MAX_WIDTH);
    GB2_REQUIRE(ctx, stride >= w && stride <= 512, "gb2_encode_kmers: stride %lld outside [w,512]", (long long)stride);
    if (n == 0) return GB2_OK;
    GB2_REQUIRE(ctx, d_ascii && d_packed && d_nmask, "gb2_encode_kmers: null buffer");
    GB2_CUDA(ctx, cudaSetDevice(ctx->device));
    const bool wide = w > GB2_NARROW_WIDTH;
    GB2_REQUIRE(ctx, !wide || ((uintptr_t)d_packed & 15u) == 0, "gb2_encode_kmers: wide k-mers need a 16-byte aligned output");
    const size_t smem = (size_t)ENC_ROWS * (size_t)stride + 48 + (wide ? 32 : 0);
    auto kern = wide ? gb2_encode_kernel<true> : gb2_encode_kernel<false>;
    if (smem > 48 * 1024)
        GB2_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t blocks = gb2_div_up(n, ENC_ROWS);
    GB2_REQUIRE(ctx, blocks < ((int64_t)1 << 31), "gb2_encode_kmers: too many rows for one launch");
    kern<<<(unsigned)blocks, ENC_ROWS, smem, ctx->stream>>>(d_ascii, n, w, stride, d_packed, d_nmask,
                                                            (unsigned long long *)d_counts);
    GB2_LAUNCH_CHECK(ctx);
    return GB2_OK;
}
