// score_wide.cu -- K2 for wide k-mers (32 < w <= 64, two packed words each): one 128-bit load = one k-mer.
//
// 3-base chunks (64-entry tables, NCHUNK = ceil(w / 3) = 11..22 lookups): at 256 bytes per chunk the tables replicate 32x
// (16x / 8x for the widest motifs with the largest spans) beside a histogram of up to 33 k bins, so a lookup is one
// conflict-free wavefront.  Round 2 history (profiles/r02_wide_kernel_ncu_before.txt -> r02_configs_wide.json): with
// 4-base chunks (1 KB per chunk: R = 16 / 8 / 4 at w = 35 / 48 / 64) the shared-memory pipe was 88-96 % busy at
// 0.60 / 0.43 / 0.34 of the HBM roofline; 3-base chunks with a RUN-TIME replication factor were no faster because the
// address arithmetic (four IMAD + three LOP3 per lookup) made the kernel issue-bound; with NCHUNK and R both template
// parameters a lookup is SHF + LOP3 + LEA + LDS(immediate offset) + IADD as in the narrow kernel.
// Same packed 16-bit fields, same histogram / hit / dense semantics as the narrow kernel; when the histogram does not
// fit shared memory next to the tables (hist_in_smem == 0) it is counted with 64-bit global atomics.
#include <algorithm>

#include "score_common.cuh"

template <int Q>
__device__ __forceinline__ uint32_t pick_word(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3)
{
    return Q == 0 ? w0 : Q == 1 ? w1 : Q == 2 ? w2 : Q == 3 ? w3 : 0u;
}

template <int C, int NCHUNK, int R>
struct WideSum {
    static __device__ __forceinline__ uint32_t run(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t lut32)
    {
        constexpr int bit = 6 * C, q = bit >> 5, sh = bit & 31;
        const uint32_t lo = pick_word<q>(w0, w1, w2, w3);
        const uint32_t x = (sh + 6 <= 32) ? (lo >> sh) : __funnelshift_r(lo, pick_word<q + 1>(w0, w1, w2, w3), sh);
        return lds_u32<C * 64 * R * 4>(lut32 + (x & 63u) * (uint32_t)(R * 4)) + WideSum<C + 1, NCHUNK, R>::run(w0, w1, w2, w3, lut32);
    }
};
template <int NCHUNK, int R>
struct WideSum<NCHUNK, NCHUNK, R> {
    static __device__ __forceinline__ uint32_t run(uint32_t, uint32_t, uint32_t, uint32_t, uint32_t) { return 0u; }
};

// One tile of 1024 * U k-mers.  FULL: every row of the tile exists -- no per-row bounds test (the kernel is bound by the
// integer ALU pipe, 80 % busy in profiles/r02_wide_kernel_ncu.txt, so the remainder logic stays out of the main loop).
template <int NCHUNK, int R, int U, bool FULL>
__device__ __forceinline__ void wide_tile(const ScoreParams &p, const uint4 *src, int64_t r0, uint32_t lut32, uint32_t hist32,
                                          bool hist_smem, bool do_hist, bool two, uint32_t nsent, uint32_t cut_hi, unsigned lane)
{
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int64_t row = r0 + u * 1024;
        v[u] = (FULL || row < p.n) ? ld_stream_u4(src + row) : make_uint4(0, 0, 0, 0);
    }
    uint32_t acc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) acc[u] = WideSum<0, NCHUNK, R>::run(v[u].x, v[u].y, v[u].z, v[u].w, lut32);
    if (p.nmask != nullptr) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t row = r0 + u * 1024;
            if ((FULL || row < p.n) && ((__ldg(p.nmask + (row >> 5)) >> (row & 31)) & 1u)) acc[u] = nsent;
        }
    }
    if (hist_smem) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (FULL || r0 + u * 1024 < p.n) {
                red_shared_inc(hist32 + 4u * __byte_perm(acc[u], 0u, 0x4410u));
                if (two) red_shared_inc(hist32 + 4u * (acc[u] >> 16));
            }
        }
    } else if (do_hist) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (FULL || r0 + u * 1024 < p.n) {
                atomicAdd(p.hist + (acc[u] & 0xFFFFu), 1ull);
                if (two) atomicAdd(p.hist + (acc[u] >> 16), 1ull);
            }
        }
    }
    if (p.dense != nullptr) {
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (FULL || r0 + u * 1024 < p.n) p.dense[r0 + u * 1024] = (acc[u] == nsent) ? 0xFFFFFFFFu : acc[u];
    }
    if (p.hits != nullptr) {
        // N rows carry bin `span` in both fields: above the cut, so they pass this screen and are rejected by bin_hits
        uint32_t mx = FULL ? acc[0] : 0u;
#pragma unroll
        for (int u = FULL ? 1 : 0; u < U; ++u) mx = __vimax3_u16x2(mx, (FULL || r0 + u * 1024 < p.n) ? acc[u] : 0u, 0u);
        const bool any = ((mx & 0xFFFFu) >= p.cut) | (two & (mx >= cut_hi));
        if (__any_sync(0xFFFFFFFFu, any)) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t row = r0 + u * 1024;
                const bool ok = FULL || row < p.n;
                const uint32_t bf = acc[u] & 0xFFFFu, br = acc[u] >> 16;
                append_hits(p, ok && bin_hits(p, bf), (uint64_t)row, bf, 0u, lane);
                if (two) append_hits(p, ok && bin_hits(p, br), (uint64_t)row, br, 1u, lane);
            }
        }
    }
}

template <int NCHUNK, int R>
__global__ void __launch_bounds__(1024, 1) gb2_score_wide_kernel(const ScoreParams p, int hist_in_smem)
{
    extern __shared__ __align__(16) uint32_t smem[];
    uint32_t *lut_s = smem;                     // [NCHUNK*64][R]
    uint32_t *hist_s = smem + NCHUNK * 64 * R;  // [span+1] when hist_in_smem
    const unsigned tid = threadIdx.x, lane = tid & 31u;
    const bool do_hist = p.hist != nullptr;
    const bool hist_smem = do_hist && hist_in_smem;

    for (int i = tid; i < NCHUNK * 64 * R; i += 1024) lut_s[i] = p.lut[i / R];
    if (hist_smem)
        for (uint32_t i = tid; i <= p.span; i += 1024) hist_s[i] = 0u;
    __syncthreads();

    const uint32_t lut32 = smem_u32(lut_s) + 4u * (lane & (uint32_t)(R - 1));
    const uint32_t hist32 = smem_u32(hist_s);
    const uint4 *src = reinterpret_cast<const uint4 *>(p.packed);
    const uint32_t nsent = (p.span << 16) | p.span;
    const uint32_t cut_hi = p.cut << 16;
    const bool two = p.two_strands != 0;
    constexpr int U = 4;
    constexpr int64_t TILE = 1024 * U;
    const int64_t nfull = p.n / TILE;
    for (int64_t tile = blockIdx.x; tile < nfull; tile += gridDim.x)
        wide_tile<NCHUNK, R, U, true>(p, src, tile * TILE + tid, lut32, hist32, hist_smem, do_hist, two, nsent, cut_hi, lane);
    if (nfull * TILE < p.n && (int64_t)blockIdx.x == nfull % gridDim.x)  // the remainder: one CTA, guarded
        wide_tile<NCHUNK, R, U, false>(p, src, nfull * TILE + tid, lut32, hist32, hist_smem, do_hist, two, nsent, cut_hi, lane);
    if (hist_smem) {
        __syncthreads();
        for (uint32_t i = tid; i <= p.span; i += 1024) {
            const uint32_t c = hist_s[i];
            if (c) atomicAdd(p.hist + i, (unsigned long long)c);
        }
    }
}

template <int NCHUNK, int R>
static int launch_wide(gb2_ctx *ctx, const ScoreParams &p, size_t smem, int grid, int hist_in_smem)
{
    auto kern = gb2_score_wide_kernel<NCHUNK, R>;
    GB2_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, 1024, smem, ctx->stream>>>(p, hist_in_smem);
    GB2_LAUNCH_CHECK(ctx);
    return GB2_OK;
}

template <int NCHUNK>
static int dispatch_wide_r(gb2_ctx *ctx, int R, const ScoreParams &p, size_t smem, int grid, int hist_in_smem)
{
    switch (R) {
    case 32: return launch_wide<NCHUNK, 32>(ctx, p, smem, grid, hist_in_smem);
    case 16: return launch_wide<NCHUNK, 16>(ctx, p, smem, grid, hist_in_smem);
    default: return launch_wide<NCHUNK, 8>(ctx, p, smem, grid, hist_in_smem);
    }
}

int gb2_launch_score_wide(gb2_ctx *ctx, const gb2_motif *m, const ScoreParams &p, int64_t n)
{
    GB2_REQUIRE(ctx, m->chunk_bases == 3 && m->n_chunks >= 11 && m->n_chunks <= 22 && m->replicas >= 8,
                "gb2_score: bad chunk plan of a wide motif");
    const size_t smem = (size_t)m->smem_bytes;
    const int grid = (int)std::min<int64_t>(ctx->sm_count, std::max<int64_t>(1, gb2_div_up(n, 4096)));
    const int hs = m->hist_global ? 0 : 1;
    switch (m->n_chunks) {
    case 11: return dispatch_wide_r<11>(ctx, m->replicas, p, smem, grid, hs);
    case 12: return dispatch_wide_r<12>(ctx, m->replicas, p, smem, grid, hs);
    case 13: return dispatch_wide_r<13>(ctx, m->replicas, p, smem, grid, hs);
    case 14: return dispatch_wide_r<14>(ctx, m->replicas, p, smem, grid, hs);
    case 15: return dispatch_wide_r<15>(ctx, m->replicas, p, smem, grid, hs);
    case 16: return dispatch_wide_r<16>(ctx, m->replicas, p, smem, grid, hs);
    case 17: return dispatch_wide_r<17>(ctx, m->replicas, p, smem, grid, hs);
    case 18: return dispatch_wide_r<18>(ctx, m->replicas, p, smem, grid, hs);
    case 19: return dispatch_wide_r<19>(ctx, m->replicas, p, smem, grid, hs);
    case 20: return dispatch_wide_r<20>(ctx, m->replicas, p, smem, grid, hs);
    case 21: return dispatch_wide_r<21>(ctx, m->replicas, p, smem, grid, hs);
    default: return dispatch_wide_r<22>(ctx, m->replicas, p, smem, grid, hs);
    }
}
