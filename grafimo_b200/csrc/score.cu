// score.cu -- K2: forward + reverse-complement integer PWM scoring of packed k-mers, per-score
// histogram (for the q-values), p-value cut-off and warp-aggregated hit compaction.
//
// Replaces compute_score_seq (src/grafimo/score_sequences.py:331-396) and the p-value filter of
// ResultTmp.to_df (src/grafimo/resultsTmp.py:303-307).
//
// Design (HBM-bound integer work, no tensor cores):
//   * one 128-bit load brings two packed k-mers (8 B each); both strands are scored from that
//     register -- the reverse complement is never read or materialised;
//   * the scaled matrix is folded into ceil(w/4) lookup tables of 256 entries (one per 4-base
//     byte of the packed word).  An entry holds (rc_part << 16) | fwd_part, both relative to the
//     column minima, so ONE shared-memory lookup + ONE integer add per byte accumulates both
//     strands, and the two 16-bit fields of the sum are directly the histogram bins (score - lo);
//   * the tables are replicated R times in shared memory with the replica chosen by the lane
//     (entry e, replica r at word e*R + r): with R = 32 every lane reads its own bank, so a lookup
//     of 32 random entries costs one conflict-free shared-memory wavefront;
//   * the histogram lives in shared memory (u32, one per CTA) and is flushed once per CTA with
//     64-bit global atomics; hits are rare and are appended with one global atomic per warp;
//   * persistent grid: one 1024-thread CTA per SM (the tables take most of the 227 KB).
#include "score_common.cuh"

// dense scores of pair j (k-mers 2j, 2j+1): one 8-byte store when the buffer allows it
__device__ __forceinline__ void store_dense_pair(uint32_t *dense, int64_t j, uint2 o, bool aligned8)
{
    if (aligned8) {
        reinterpret_cast<uint2 *>(dense)[j] = o;
    } else {
        dense[2 * j] = o.x;
        dense[2 * j + 1] = o.y;
    }
}

// Rare path, entered by the whole warp: exact per-bin test + warp-aggregated append for one pair.
__device__ __forceinline__ void emit_pair_hits(const ScoreParams &p, uint32_t a0, uint32_t a1, int64_t j, bool ok, bool two,
                                               unsigned lane)
{
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const uint32_t a = h ? a1 : a0;
        const uint64_t row = (uint64_t)(2 * j + h);
        const uint32_t bf = a & 0xFFFFu, br = a >> 16;
        append_hits(p, ok && bin_hits(p, bf), row, bf, 0u, lane);
        if (two) append_hits(p, ok && bin_hits(p, br), row, br, 1u, lane);
    }
}

template <int CB, int NCHUNK, int R, int U>
__global__ void __launch_bounds__(1024, 1) gb2_score_kernel(const ScoreParams p)
{
    extern __shared__ __align__(16) uint32_t smem[];
    constexpr int LUT_WORDS = NCHUNK * ChunkGeom<CB>::ENTRIES * R;
    uint32_t *lut_s = smem;               // [NCHUNK * 4^CB][R]
    uint32_t *hist_s = smem + LUT_WORDS;  // [span+1]
    const unsigned tid = threadIdx.x, lane = tid & 31u;
    const bool do_hist = p.hist != nullptr;

    for (int i = tid; i < LUT_WORDS; i += 1024) lut_s[i] = p.lut[i / R];
    if (do_hist)
        for (uint32_t i = tid; i <= p.span; i += 1024) hist_s[i] = 0u;
    __syncthreads();

    const uint32_t lut32 = smem_u32(lut_s) + 4u * (lane & (R - 1));
    const uint32_t hist32 = smem_u32(hist_s);
    const uint4 *src = reinterpret_cast<const uint4 *>(p.packed);
    const int64_t npairs = p.n >> 1;
    const uint32_t nsent = (p.span << 16) | p.span;  // both fields -> bin `span`
    const uint32_t cut_hi = p.cut << 16;
    const bool two = p.two_strands != 0;
    const bool has_n = p.nmask != nullptr;
    const bool dense8 = (reinterpret_cast<uintptr_t>(p.dense) & 7u) == 0;  // a batch may start at an odd row of a larger buffer

    // ---- full tiles: 1024*U pairs, no bounds checks; thread t owns pairs t, t+1024, ... of the tile,
    //      so every warp-wide load covers 512 contiguous bytes and U loads are in flight per thread
    constexpr int TILE = 1024 * U;
    const int64_t nfull = npairs / TILE;
    for (int64_t tile = blockIdx.x; tile < nfull; tile += gridDim.x) {
        const int64_t j0 = tile * TILE + tid;
        const uint4 *s = src + j0;
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = ld_stream_u4(s + u * 1024);
        uint32_t acc[2 * U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            acc[2 * u] = score_word<CB, NCHUNK, R>(v[u].x, v[u].y, lut32);
            acc[2 * u + 1] = score_word<CB, NCHUNK, R>(v[u].z, v[u].w, lut32);
        }
        if (has_n) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t j = j0 + u * 1024;
                const uint32_t nb = (__ldg(p.nmask + (j >> 4)) >> ((j & 15) * 2)) & 3u;
                if (nb & 1u) acc[2 * u] = nsent;
                if (nb & 2u) acc[2 * u + 1] = nsent;
            }
        }
        if (do_hist) {
#pragma unroll
            for (int k = 0; k < 2 * U; ++k) {
                red_shared_inc(hist32 + 4u * __byte_perm(acc[k], 0u, 0x4410u));
                if (two) red_shared_inc(hist32 + 4u * (acc[k] >> 16));
            }
        }
        if (p.dense != nullptr) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                uint2 o;
                o.x = (acc[2 * u] == nsent) ? 0xFFFFFFFFu : acc[2 * u];
                o.y = (acc[2 * u + 1] == nsent) ? 0xFFFFFFFFu : acc[2 * u + 1];
                store_dense_pair(p.dense, j0 + u * 1024, o, dense8);
            }
        }
        // cheap screen: per-field maximum over the 2U k-mers (packed 16-bit max), one compare per strand
        uint32_t mx = acc[0];
#pragma unroll
        for (int k = 1; k + 1 < 2 * U; k += 2) mx = __vimax3_u16x2(mx, acc[k], acc[k + 1]);
        mx = __vimax3_u16x2(mx, acc[2 * U - 1], acc[2 * U - 1]);
        const bool any = ((mx & 0xFFFFu) >= p.cut) | (two & (mx >= cut_hi));
        if (p.hits != nullptr && __any_sync(0xFFFFFFFFu, any)) {
#pragma unroll
            for (int u = 0; u < U; ++u) emit_pair_hits(p, acc[2 * u], acc[2 * u + 1], j0 + u * 1024, true, two, lane);
        }
    }

    // ---- remainder (< TILE pairs): one CTA, guarded
    if ((int64_t)blockIdx.x == nfull % gridDim.x) {
        for (int64_t jb = nfull * TILE; jb < npairs; jb += 1024) {
            const int64_t j = jb + tid;
            const bool ok = j < npairs;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (ok) v = ld_stream_u4(src + j);
            uint32_t a0 = score_word<CB, NCHUNK, R>(v.x, v.y, lut32);
            uint32_t a1 = score_word<CB, NCHUNK, R>(v.z, v.w, lut32);
            if (has_n && ok) {
                const uint32_t nb = (__ldg(p.nmask + (j >> 4)) >> ((j & 15) * 2)) & 3u;
                if (nb & 1u) a0 = nsent;
                if (nb & 2u) a1 = nsent;
            }
            if (ok) {
                if (do_hist) {
                    red_shared_inc(hist32 + 4u * (a0 & 0xFFFFu));
                    red_shared_inc(hist32 + 4u * (a1 & 0xFFFFu));
                    if (two) {
                        red_shared_inc(hist32 + 4u * (a0 >> 16));
                        red_shared_inc(hist32 + 4u * (a1 >> 16));
                    }
                }
                if (p.dense != nullptr) {
                    uint2 o;
                    o.x = (a0 == nsent) ? 0xFFFFFFFFu : a0;
                    o.y = (a1 == nsent) ? 0xFFFFFFFFu : a1;
                    store_dense_pair(p.dense, j, o, dense8);
                }
            }
            if (p.hits != nullptr) emit_pair_hits(p, a0, a1, j, ok, two, lane);
        }
        // odd tail: the last k-mer has no pair partner; warp 0 handles it
        if ((p.n & 1) && tid < 32) {
            const int64_t row = p.n - 1;
            const bool mine = lane == 0;
            uint32_t a = 0;
            if (mine) {
                const uint64_t x = p.packed[row];
                a = score_word<CB, NCHUNK, R>((uint32_t)x, (uint32_t)(x >> 32), lut32);
                if (has_n && ((__ldg(p.nmask + (row >> 5)) >> (row & 31)) & 1u)) a = nsent;
                if (do_hist) {
                    red_shared_inc(hist32 + 4u * (a & 0xFFFFu));
                    if (two) red_shared_inc(hist32 + 4u * (a >> 16));
                }
                if (p.dense != nullptr) p.dense[row] = (a == nsent) ? 0xFFFFFFFFu : a;
            }
            if (p.hits != nullptr) {
                append_hits(p, mine && bin_hits(p, a & 0xFFFFu), (uint64_t)row, a & 0xFFFFu, 0u, lane);
                if (two) append_hits(p, mine && bin_hits(p, a >> 16), (uint64_t)row, a >> 16, 1u, lane);
            }
        }
    }

    if (do_hist) {
        __syncthreads();
        for (uint32_t i = tid; i <= p.span; i += 1024) {
            const uint32_t c = hist_s[i];
            if (c) atomicAdd(p.hist + i, (unsigned long long)c);
        }
    }
}

// ---------------------------------------------------------------------------------------------
template <int CB, int NCHUNK, int R>
static int launch_score(gb2_ctx *ctx, const ScoreParams &p, size_t smem, int grid)
{
    auto kern = gb2_score_kernel<CB, NCHUNK, R, 4>;
    GB2_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, 1024, smem, ctx->stream>>>(p);
    GB2_LAUNCH_CHECK(ctx);
    return GB2_OK;
}

template <int NCHUNK>
static int dispatch_r(gb2_ctx *ctx, int R, const ScoreParams &p, size_t smem, int grid)
{
    switch (R) {
    case 32: return launch_score<4, NCHUNK, 32>(ctx, p, smem, grid);
    case 16: return launch_score<4, NCHUNK, 16>(ctx, p, smem, grid);
    case 8: return launch_score<4, NCHUNK, 8>(ctx, p, smem, grid);
    case 4: return launch_score<4, NCHUNK, 4>(ctx, p, smem, grid);
    case 2: return launch_score<4, NCHUNK, 2>(ctx, p, smem, grid);
    default: return launch_score<4, NCHUNK, 1>(ctx, p, smem, grid);
    }
}

// 3-base chunks (narrow motifs of 19..32 bp whose 4-base tables would not replicate 32x): R = 32, 16 or 8
template <int NCHUNK>
static int dispatch_r3(gb2_ctx *ctx, int R, const ScoreParams &p, size_t smem, int grid)
{
    switch (R) {
    case 32: return launch_score<3, NCHUNK, 32>(ctx, p, smem, grid);
    case 16: return launch_score<3, NCHUNK, 16>(ctx, p, smem, grid);
    default: return launch_score<3, NCHUNK, 8>(ctx, p, smem, grid);
    }
}

// p-value cut-off as an integer test: hit <=> ptab[bin] < threshold (strict, resultsTmp.py:305-307)
int gb2_fill_score_params(gb2_ctx *ctx, const gb2_motif *m, double p_threshold, ScoreParams &p)
{
    p.lut = m->d_lut;
    p.span = (uint32_t)m->span;
    p.lo = (int32_t)m->lo;
    const std::vector<double> &pt = m->h_ptab;
    int64_t cut = m->span;
    for (int64_t k = 0; k < m->span; ++k)
        if (pt[(size_t)k] < p_threshold) { cut = k; break; }
    bool simple = true;
    for (int64_t k = cut; k < m->span; ++k)
        if (!(pt[(size_t)k] < p_threshold)) { simple = false; break; }
    p.cut = (uint32_t)cut;
    p.bitmap = nullptr;
    if (!simple) {  // p-value table not monotone across the cut: exact per-bin bitmap
        std::vector<uint32_t> bm((size_t)gb2_div_up(m->span + 1, 32), 0u);
        for (int64_t k = cut; k < m->span; ++k)
            if (pt[(size_t)k] < p_threshold) bm[(size_t)(k >> 5)] |= 1u << (k & 31);
        GB2_CUDA(ctx, cudaMemcpyAsync(m->d_bitmap, bm.data(), bm.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
        GB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // bm is a stack-lifetime host buffer
        p.bitmap = m->d_bitmap;
    }
    return GB2_OK;
}

extern "C" int gb2_score(gb2_ctx *ctx, const gb2_motif *m, const uint64_t *d_packed, const uint32_t *d_nmask, int64_t n,
                         uint64_t row_base, int strands, double p_threshold, uint64_t *d_hist, gb2_hit *d_hits,
                         uint64_t hit_capacity, uint64_t *d_hit_count, uint32_t *d_dense)
{
    if (!ctx || !m) return GB2_ERR_ARG;
    GB2_REQUIRE(ctx, n >= 0, "gb2_score: negative row count");
    GB2_REQUIRE(ctx, strands == 1 || strands == 2, "gb2_score: strands must be 1 or 2");
    GB2_REQUIRE(ctx, !(p_threshold != p_threshold) && p_threshold > 0.0, "gb2_score: threshold must be > 0");
    GB2_REQUIRE(ctx, m->device == ctx->device, "gb2_score: motif lives on device %d, context on %d", m->device, ctx->device);
    if (n == 0) return GB2_OK;
    GB2_REQUIRE(ctx, d_packed != nullptr, "gb2_score: null k-mer buffer");
    GB2_REQUIRE(ctx, ((uintptr_t)d_packed & 15u) == 0, "gb2_score: packed k-mers must be 16-byte aligned");
    GB2_REQUIRE(ctx, d_hits == nullptr || d_hit_count != nullptr, "gb2_score: hit buffer without a counter");
    GB2_REQUIRE(ctx, d_dense == nullptr || ((uintptr_t)d_dense & 3u) == 0, "gb2_score: dense buffer must be 4-byte aligned");
    GB2_CUDA(ctx, cudaSetDevice(ctx->device));

    ScoreParams p;
    p.packed = d_packed;
    p.nmask = d_nmask;
    p.n = n;
    p.row_base = row_base;
    p.two_strands = strands == 2;
    p.hist = (unsigned long long *)d_hist;
    p.hits = d_hits;
    p.hit_capacity = hit_capacity;
    p.hit_count = (unsigned long long *)d_hit_count;
    p.dense = d_dense;
    {
        const int rc = gb2_fill_score_params(ctx, m, p_threshold, p);
        if (rc != GB2_OK) return rc;
    }

    const size_t smem = (size_t)m->smem_bytes;
    if (m->w > GB2_NARROW_WIDTH) return gb2_launch_score_wide(ctx, m, p, n);  // two packed words per k-mer: score_wide.cu
    const int64_t npairs = n >> 1;
    int grid = (int)std::min<int64_t>(ctx->sm_count, std::max<int64_t>(1, gb2_div_up(npairs, 4096)));
    if (m->chunk_bases == 3) {
        switch (m->n_chunks) {
        case 7: return dispatch_r3<7>(ctx, m->replicas, p, smem, grid);
        case 8: return dispatch_r3<8>(ctx, m->replicas, p, smem, grid);
        case 9: return dispatch_r3<9>(ctx, m->replicas, p, smem, grid);
        case 10: return dispatch_r3<10>(ctx, m->replicas, p, smem, grid);
        default: return dispatch_r3<11>(ctx, m->replicas, p, smem, grid);
        }
    }
    switch (m->n_chunks) {
    case 1: return dispatch_r<1>(ctx, m->replicas, p, smem, grid);
    case 2: return dispatch_r<2>(ctx, m->replicas, p, smem, grid);
    case 3: return dispatch_r<3>(ctx, m->replicas, p, smem, grid);
    case 4: return dispatch_r<4>(ctx, m->replicas, p, smem, grid);
    case 5: return dispatch_r<5>(ctx, m->replicas, p, smem, grid);
    case 6: return dispatch_r<6>(ctx, m->replicas, p, smem, grid);
    case 7: return dispatch_r<7>(ctx, m->replicas, p, smem, grid);
    default: return dispatch_r<8>(ctx, m->replicas, p, smem, grid);
    }
}
