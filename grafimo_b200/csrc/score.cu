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
#include "internal.cuh"

struct ScoreParams {
    const uint64_t *packed;
    const uint32_t *nmask;
    int64_t n;
    uint64_t row_base;
    const uint32_t *lut;     // [n_chunks][256]
    const uint32_t *bitmap;  // hit bitmap over bins, or nullptr when "bin >= cut" is the whole test
    uint32_t span;           // number of score bins; bin `span` collects N rows
    uint32_t cut;            // smallest bin that can be a hit
    int32_t lo;              // absolute score of bin 0
    int32_t two_strands;
    unsigned long long *hist;  // global [span+1] or nullptr
    gb2_hit *hits;
    unsigned long long hit_capacity;
    unsigned long long *hit_count;
    uint32_t *dense;
};

__device__ __forceinline__ uint4 ld_stream_u4(const uint4 *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

template <int NCHUNK, int R>
__device__ __forceinline__ uint32_t score_word(uint32_t w0, uint32_t w1, const uint32_t *my)
{
    // my = lut_s + (lane & (R-1)); entry (c, b) at my[(c*256 + b) * R]
    uint32_t acc = 0;
#pragma unroll
    for (int c = 0; c < NCHUNK; ++c) {
        const uint32_t word = (c < 4) ? w0 : w1;
        const uint32_t b = (word >> (8 * (c & 3))) & 0xFFu;
        acc += my[(c * 256 + b) * R];
    }
    return acc;
}

// Appends the hits of one (k-mer, strand) slot across the warp: one ballot, one global atomic.
__device__ __forceinline__ void append_hits(const ScoreParams &p, bool pred, uint64_t row, uint32_t bin, uint32_t strand,
                                            unsigned lane)
{
    const unsigned m = __ballot_sync(0xFFFFFFFFu, pred);
    if (m == 0) return;
    unsigned long long base = 0;
    if (lane == (unsigned)(__ffs(m) - 1)) base = atomicAdd(p.hit_count, (unsigned long long)__popc(m));
    base = __shfl_sync(0xFFFFFFFFu, base, __ffs(m) - 1);
    if (pred) {
        const unsigned long long slot = base + __popc(m & ((1u << lane) - 1u));
        if (slot < p.hit_capacity) {
            uint4 rec;
            const uint64_t grow = p.row_base + row;
            rec.x = (uint32_t)grow;
            rec.y = (uint32_t)(grow >> 32);
            rec.z = (uint32_t)(p.lo + (int32_t)bin);
            rec.w = strand;
            reinterpret_cast<uint4 *>(p.hits)[slot] = rec;
        }
    }
}

__device__ __forceinline__ bool bin_hits(const ScoreParams &p, uint32_t bin)
{
    if (bin < p.cut || bin >= p.span) return false;  // bin == span: N row, never a hit
    if (p.bitmap == nullptr) return true;
    return (p.bitmap[bin >> 5] >> (bin & 31)) & 1u;
}

template <int NCHUNK, int R, int U>
__global__ void __launch_bounds__(1024, 1) gb2_score_kernel(const ScoreParams p)
{
    extern __shared__ __align__(16) uint32_t smem[];
    uint32_t *lut_s = smem;                      // [NCHUNK*256][R]
    uint32_t *hist_s = smem + NCHUNK * 256 * R;  // [span+1]
    const unsigned tid = threadIdx.x, lane = tid & 31u;
    const bool do_hist = p.hist != nullptr;

    for (int i = tid; i < NCHUNK * 256 * R; i += blockDim.x) lut_s[i] = p.lut[i / R];
    if (do_hist)
        for (uint32_t i = tid; i <= p.span; i += blockDim.x) hist_s[i] = 0u;
    __syncthreads();

    const uint32_t *my = lut_s + (lane & (R - 1));
    const uint4 *src = reinterpret_cast<const uint4 *>(p.packed);
    const int64_t npairs = p.n >> 1;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const uint32_t nsent = (p.span << 16) | p.span;  // both fields -> bin `span`
    const uint32_t cut_hi = p.cut << 16;
    const bool two = p.two_strands != 0;

    // every lane of a warp runs the same trip count (ballots below need the full warp)
    const int64_t first = (int64_t)blockIdx.x * blockDim.x + tid;
    const int64_t warp_first = first - lane;
    for (int64_t base = warp_first; base < npairs; base += stride * U) {
        uint4 v[U];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t j = base + lane + (int64_t)u * stride;
            ok[u] = j < npairs;
            v[u] = ok[u] ? ld_stream_u4(src + j) : make_uint4(0, 0, 0, 0);
        }
        uint32_t acc[2 * U];
        bool any = false;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t j = base + lane + (int64_t)u * stride;
            uint32_t a0 = score_word<NCHUNK, R>(v[u].x, v[u].y, my);
            uint32_t a1 = score_word<NCHUNK, R>(v[u].z, v[u].w, my);
            if (p.nmask != nullptr && ok[u]) {
                const int64_t row = 2 * j;
                const uint32_t nb = (__ldg(p.nmask + (row >> 5)) >> (row & 31)) & 3u;
                if (nb & 1u) a0 = nsent;
                if (nb & 2u) a1 = nsent;
            }
            acc[2 * u] = a0;
            acc[2 * u + 1] = a1;
            if (ok[u]) {
                if (do_hist) {
                    atomicAdd(&hist_s[a0 & 0xFFFFu], 1u);
                    atomicAdd(&hist_s[a1 & 0xFFFFu], 1u);
                    if (two) {
                        atomicAdd(&hist_s[a0 >> 16], 1u);
                        atomicAdd(&hist_s[a1 >> 16], 1u);
                    }
                }
                if (p.dense != nullptr) {
                    uint2 d;
                    d.x = (a0 == nsent) ? 0xFFFFFFFFu : a0;
                    d.y = (a1 == nsent) ? 0xFFFFFFFFu : a1;
                    reinterpret_cast<uint2 *>(p.dense)[j] = d;
                }
                any |= ((a0 & 0xFFFFu) >= p.cut) | ((a1 & 0xFFFFu) >= p.cut);
                if (two) any |= (a0 >= cut_hi) | (a1 >= cut_hi);
            }
        }
        if (p.hits != nullptr && __any_sync(0xFFFFFFFFu, any)) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t j = base + lane + (int64_t)u * stride;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t a = acc[2 * u + h];
                    const uint64_t row = (uint64_t)(2 * j + h);
                    const uint32_t bf = a & 0xFFFFu, br = a >> 16;
                    append_hits(p, ok[u] && bin_hits(p, bf), row, bf, 0u, lane);
                    if (two) append_hits(p, ok[u] && bin_hits(p, br), row, br, 1u, lane);
                }
            }
        }
    }

    // odd tail: the last k-mer has no pair partner; warp 0 of block 0 handles it
    if ((p.n & 1) && blockIdx.x == 0 && tid < 32) {
        const int64_t row = p.n - 1;
        const bool mine = lane == 0;
        uint32_t a = 0;
        if (mine) {
            const uint64_t x = p.packed[row];
            a = score_word<NCHUNK, R>((uint32_t)x, (uint32_t)(x >> 32), my);
            if (p.nmask != nullptr && ((__ldg(p.nmask + (row >> 5)) >> (row & 31)) & 1u)) a = nsent;
            if (do_hist) {
                atomicAdd(&hist_s[a & 0xFFFFu], 1u);
                if (two) atomicAdd(&hist_s[a >> 16], 1u);
            }
            if (p.dense != nullptr) p.dense[row] = (a == nsent) ? 0xFFFFFFFFu : a;
        }
        if (p.hits != nullptr) {
            append_hits(p, mine && bin_hits(p, a & 0xFFFFu), (uint64_t)row, a & 0xFFFFu, 0u, lane);
            if (two) append_hits(p, mine && bin_hits(p, a >> 16), (uint64_t)row, a >> 16, 1u, lane);
        }
    }

    if (do_hist) {
        __syncthreads();
        for (uint32_t i = tid; i <= p.span; i += blockDim.x) {
            const uint32_t c = hist_s[i];
            if (c) atomicAdd(p.hist + i, (unsigned long long)c);
        }
    }
}

// ---------------------------------------------------------------------------------------------
template <int NCHUNK, int R>
static int launch_score(gb2_ctx *ctx, const ScoreParams &p, size_t smem, int grid)
{
    auto kern = gb2_score_kernel<NCHUNK, R, 2>;
    GB2_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, 1024, smem, ctx->stream>>>(p);
    GB2_LAUNCH_CHECK(ctx);
    return GB2_OK;
}

template <int NCHUNK>
static int dispatch_r(gb2_ctx *ctx, int R, const ScoreParams &p, size_t smem, int grid)
{
    switch (R) {
    case 32: return launch_score<NCHUNK, 32>(ctx, p, smem, grid);
    case 16: return launch_score<NCHUNK, 16>(ctx, p, smem, grid);
    case 8: return launch_score<NCHUNK, 8>(ctx, p, smem, grid);
    case 4: return launch_score<NCHUNK, 4>(ctx, p, smem, grid);
    case 2: return launch_score<NCHUNK, 2>(ctx, p, smem, grid);
    default: return launch_score<NCHUNK, 1>(ctx, p, smem, grid);
    }
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
    GB2_REQUIRE(ctx, d_dense == nullptr || ((uintptr_t)d_dense & 7u) == 0, "gb2_score: dense buffer must be 8-byte aligned");
    GB2_CUDA(ctx, cudaSetDevice(ctx->device));

    ScoreParams p;
    p.packed = d_packed;
    p.nmask = d_nmask;
    p.n = n;
    p.row_base = row_base;
    p.lut = m->d_lut;
    p.span = (uint32_t)m->span;
    p.lo = (int32_t)m->lo;
    p.two_strands = strands == 2;
    p.hist = (unsigned long long *)d_hist;
    p.hits = d_hits;
    p.hit_capacity = hit_capacity;
    p.hit_count = (unsigned long long *)d_hit_count;
    p.dense = d_dense;

    // p-value cut-off as an integer test: hit <=> ptab[bin] < threshold (strict, resultsTmp.py:305-307)
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

    const int64_t npairs = n >> 1;
    int grid = (int)std::min<int64_t>(ctx->sm_count, std::max<int64_t>(1, gb2_div_up(npairs, 1024)));
    const size_t smem = (size_t)m->smem_bytes;
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
