// dense_sort.cu -- K6, dense form, without a library sort: the report rows of an unselective scan (`-t 1`) from K2's dense
// scores by two fused stable partition passes.
//
// Replaces the filter + sort of ResultTmp.to_df (src/grafimo/resultsTmp.py:303-313) for the case where every window is a
// report row.  Round 1 did: key kernel (p-rank + index per window) -> cub::DeviceRadixSort::SortPairs (histogram kernel +
// two Onesweep passes over 8-byte pairs) -> gather kernel; 1.75 of the 1.93 ms of a 60 M-row step, 1.07 ms of it in the library
// sort (profiles/r01_c5_dense_vs_hits_launch_list.txt).  The order wanted is (p-rank, window index) and the windows arrive in
// index order, so a STABLE least-significant-digit sort on the <= 16-bit p-rank is enough: two passes of 8 bits.  Written by
// hand the two passes fuse with their neighbours:
//   pass A  reads K2's dense scores directly (no key array: rank = rank_table[bin]; windows that fail the threshold are
//           dropped here) and writes (rank u16, window u32) -- 6 bytes per kept window -- partitioned by the LOW digit;
//   pass B  partitions by the HIGH digit and writes the six report columns at their final position (no gather pass).
// Each pass is count -> exclusive scan -> scatter over tiles of THREADS x ITEMS consecutive elements: the count kernel leaves
// counts[digit][tile], one exclusive scan over that array (digit-major) IS the global offset of every (digit, tile) run, and
// the scatter kernel ranks its tile stably in shared memory -- every warp owns a contiguous slice of the tile and private
// digit counters, the 32 lanes of an iteration find their peers (same digit) with eight ballots, a per-digit scan over the
// warps orders the slices -- stages the tile sorted by digit and writes every run with consecutive threads on consecutive
// addresses.  In pass B the elements of a tile share their low digit (pass A sorted by it), so a run has ONE rank: the table
// lookups of the columns are warp-uniform.  Traffic per 60 M windows: 0.12 GB (count A) + 0.12 + 0.36 (scatter A) + 0.12
// (count B) + 0.36 + 2.2 GB (scatter B: the columns themselves) against ~5.3 GB before.
// The rank of a bin is static per motif (position in p-ascending order, K5) and the reported bins are a PREFIX in rank order
// (p is non-decreasing in the rank by construction and so is the Benjamini-Hochberg q-value: a reverse running minimum), so
// "is this window reported" is one compare with the number of kept bins.  Pass A maps bin -> rank arithmetically when the
// p-value table is monotone in the score (rank = span - 1 - bin) and through a u16 table in shared memory otherwise (a
// global-memory table costs one L1 tag lookup per lane: measured 104 us for the count kernel alone); its kernels are
// persistent so that the table is loaded once per CTA.
// GB2_DENSE_CUB=1 selects the library-sort form (gb2_finalize_dense_cub, qvalue.cu) -- kept as the checker of this one.
#include <stdlib.h>

#include <cub/cub.cuh>

#include "internal.cuh"

#define DS_NONE 0xFFFFFFFFu

// per-bin values of the report columns (score_sequences.py:393 for the score), the inverse rank table, the u16 rank table of
// pass A and the number of reported bins (they are the first `n_keep` in rank order)
__global__ void gb2_ds_tables_kernel(const double *__restrict__ ptab, const double *__restrict__ qtab,
                                     const uint32_t *__restrict__ rank, uint32_t span, double p_thr, int q_filter, double q_thr,
                                     int32_t lo_score, int w, double scale, double offset, uint16_t *__restrict__ rank16,
                                     uint32_t *__restrict__ bin_of_rank, double *__restrict__ score_tab, uint32_t *__restrict__ n_keep)
{
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > span) return;
    const double p = b < span ? ptab[b] : 1.0;  // bin `span`: N rows, p = 1
    const bool keep = p < p_thr && (!q_filter || qtab[b] < q_thr);  // strict, resultsTmp.py:305-307
    const uint32_t r = rank[b];
    rank16[b] = (uint16_t)r;
    bin_of_rank[r] = b;
    const int32_t sc = lo_score + (int32_t)b;
    score_tab[b] = __dadd_rn(__ddiv_rn((double)sc, scale), __dmul_rn((double)w, offset));
    if (keep) atomicAdd(n_keep, 1u);
}

// MODE 0: pass A (input = dense scores, digit = low byte of the rank, windows whose bin is not reported are dropped)
// MODE 1: pass B (input = the (rank, window) pairs pass A wrote, digit = high byte of the rank)
// RANK 0: rank = span - 1 - bin (monotone p-value table), RANK 1: rank = rank16[bin] from shared memory.  Pass A only.
template <int RANK>
__device__ __forceinline__ uint32_t ds_rank_of(uint32_t d, uint32_t bin, uint32_t span, const uint16_t *s_rank16, uint32_t n_keep);

template <int MODE, int RANK>
__device__ __forceinline__ void ds_load(const uint32_t *__restrict__ dense, const uint16_t *s_rank16, uint32_t n_keep, int strands,
                                        uint32_t span, const uint16_t *__restrict__ in_rank, const uint32_t *__restrict__ in_idx,
                                        uint32_t i, uint32_t n, uint32_t &rk, uint32_t &idx)
{
    rk = DS_NONE;
    idx = i;
    if (i < n) {
        if (MODE == 0) {
            const uint32_t d = __ldg(dense + (strands == 2 ? (i >> 1) : i));
            rk = ds_rank_of<RANK>(d, (strands == 2 && (i & 1u)) ? (d >> 16) : (d & 0xFFFFu), span, s_rank16, n_keep);
        } else {
            rk = in_rank[i];
            idx = in_idx[i];
        }
    }
}

template <int MODE>
__device__ __forceinline__ uint32_t ds_digit(uint32_t rk)
{
    return MODE == 0 ? (rk & 255u) : ((rk >> 8) & 255u);
}

struct DsIn {
    const uint32_t *dense;    // pass A
    const uint16_t *rank16;   // pass A, RANK 1: global copy of the table
    const uint32_t *n_keep;   // pass A: reported bins
    uint64_t n_windows;
    int strands;
    uint32_t span;
    const uint16_t *in_rank;  // pass B
    const uint32_t *in_idx;
    const uint32_t *n_in;     // pass B: windows pass A kept
    uint32_t n_tiles;
};

template <int RANK>
__device__ __forceinline__ const uint16_t *ds_stage_table(const DsIn &in, uint16_t *s_tab)
{
    if (RANK == 1) {
        for (uint32_t b = threadIdx.x; b <= in.span; b += blockDim.x) s_tab[b] = in.rank16[b];
        __syncthreads();
    }
    return s_tab;
}

// bin of a dense word's half -> rank (DS_NONE if not reported)
template <int RANK>
__device__ __forceinline__ uint32_t ds_rank_of(uint32_t d, uint32_t bin, uint32_t span, const uint16_t *s_rank16, uint32_t n_keep)
{
    bin = d == 0xFFFFFFFFu ? span : min(bin, span);  // all ones: N row
    const uint32_t r = RANK == 0 ? (bin < span ? span - 1u - bin : span) : (uint32_t)s_rank16[bin];
    return r < n_keep ? r : DS_NONE;
}

// digit counts of every tile.  The order inside a tile does not matter here, so every thread takes 16 contiguous bytes of
// each of CT consecutive tiles (all loads first: the kernel lives on memory-level parallelism).
template <int MODE, int RANK, int THREADS, int ITEMS, int CT>
__global__ void __launch_bounds__(THREADS) gb2_ds_count_kernel(DsIn in, uint32_t *__restrict__ table)
{
    constexpr uint32_t TILE = THREADS * ITEMS;
    static_assert(TILE == THREADS * 8, "one 16-byte load per thread and tile: 8 windows (two strands) or 8 ranks");
    extern __shared__ uint16_t s_dyn[];
    __shared__ uint32_t cnt[CT][256];
    const uint16_t *s_tab = ds_stage_table<(MODE == 0 ? RANK : 0)>(in, s_dyn);
    const uint32_t n = MODE == 0 ? (uint32_t)in.n_windows : *in.n_in;
    const uint32_t n_keep = MODE == 0 ? *in.n_keep : 0u;
    const uint32_t n_groups = (in.n_tiles + CT - 1) / CT;
    for (uint32_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        for (int i = threadIdx.x; i < CT * 256; i += THREADS) (&cnt[0][0])[i] = 0u;
        __syncthreads();
        if (MODE == 0) {
            // dense words: one per k-mer.  Two strands: 4 words = the 8 windows of this thread; one strand: 8 words.
            const uint32_t nw = in.strands == 2 ? 4u : 8u;
            uint32_t d[CT][8];
            uint32_t cntw[CT];
#pragma unroll
            for (int c = 0; c < CT; ++c) {
                const uint32_t e0 = (grp * CT + c) * TILE + 8u * threadIdx.x;  // first window of this thread in tile c
                const uint32_t left = e0 < n ? n - e0 : 0u;
                cntw[c] = min(left, 8u) / (in.strands == 2 ? 2u : 1u);
                const uint32_t *src = in.dense + e0 / (in.strands == 2 ? 2u : 1u);
                if (cntw[c] == nw && (reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
                    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(src));
                    d[c][0] = v.x; d[c][1] = v.y; d[c][2] = v.z; d[c][3] = v.w;
                    if (in.strands != 2) {
                        const uint4 u = __ldg(reinterpret_cast<const uint4 *>(src) + 1);
                        d[c][4] = u.x; d[c][5] = u.y; d[c][6] = u.z; d[c][7] = u.w;
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k) d[c][k] = (uint32_t)k < cntw[c] ? __ldg(src + k) : 0u;
                }
            }
#pragma unroll
            for (int c = 0; c < CT; ++c) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if ((uint32_t)k < cntw[c]) {
                        const uint32_t r0 = ds_rank_of<RANK>(d[c][k], d[c][k] & 0xFFFFu, in.span, s_tab, n_keep);
                        if (r0 != DS_NONE) atomicAdd(&cnt[c][r0 & 255u], 1u);
                        if (in.strands == 2) {
                            const uint32_t r1 = ds_rank_of<RANK>(d[c][k], d[c][k] >> 16, in.span, s_tab, n_keep);
                            if (r1 != DS_NONE) atomicAdd(&cnt[c][r1 & 255u], 1u);
                        }
                    }
                }
            }
        } else {
            uint4 v[CT];
            uint32_t left[CT];
#pragma unroll
            for (int c = 0; c < CT; ++c) {
                const uint32_t e0 = (grp * CT + c) * TILE + 8u * threadIdx.x;  // in_rank is 256-byte aligned, e0 a multiple of 8
                left[c] = e0 < n ? min(n - e0, 8u) : 0u;
                v[c] = make_uint4(0u, 0u, 0u, 0u);
                if (left[c] == 8u) {
                    v[c] = __ldg(reinterpret_cast<const uint4 *>(in.in_rank + e0));
                } else {
                    uint32_t q[4] = {0u, 0u, 0u, 0u};
                    for (uint32_t k = 0; k < left[c]; ++k) q[k >> 1] |= (uint32_t)in.in_rank[e0 + k] << (16u * (k & 1u));
                    v[c] = make_uint4(q[0], q[1], q[2], q[3]);
                }
            }
#pragma unroll
            for (int c = 0; c < CT; ++c) {
                const uint32_t q[4] = {v[c].x, v[c].y, v[c].z, v[c].w};
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if ((uint32_t)k < left[c]) atomicAdd(&cnt[c][(q[k >> 1] >> (8u + 16u * (k & 1u))) & 255u], 1u);
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < CT * 256; i += THREADS) {
            const uint32_t tile = grp * CT + (i >> 8);
            if (tile < in.n_tiles) table[(i & 255u) * in.n_tiles + tile] = cnt[i >> 8][i & 255];
        }
        __syncthreads();
    }
}

struct DsColumns {
    const uint32_t *bin_of_rank;
    const double *ptab, *qtab, *score_tab;
    uint64_t row_base;
    int32_t lo_score;
    uint64_t *o_row;
    uint8_t *o_strand;
    int32_t *o_iscore;
    double *o_score, *o_p, *o_q;
};

// "Which lanes hold my digit": every lane ORs its bit into the digit's mask word in shared memory and reads the word back.
// Measured on B200 (tools/ubench_match.cu, 32 warps/SM, random digits, scheduler cycles per warp-item): MATCH.ANY 184,
// eight ballots 71, atomicOr + read back + clear 37.
template <int MODE, int RANK, int THREADS, int ITEMS, int OCC>
__global__ void __launch_bounds__(THREADS, OCC) gb2_ds_scatter_kernel(DsIn in, const uint32_t *__restrict__ table,  // exclusive offsets [256][n_tiles]
                                                                uint16_t *__restrict__ out_rank, uint32_t *__restrict__ out_idx,
                                                                DsColumns col)
{
    constexpr int WARPS = THREADS / 32, TILE = THREADS * ITEMS;
    static_assert(TILE <= 65536 && ITEMS * 32 * 8 >= 256 * 4, "u16 tile positions; the lane masks of a warp fit its part of s_pair");
    extern __shared__ uint16_t s_dyn[];
    __shared__ uint16_t wcnt[WARPS][256];        // per-warp digit counters, then the tile position of the warp's first element of the digit
    __shared__ unsigned long long s_pair[TILE];  // the tile sorted by digit: window << 16 | rank
    __shared__ uint32_t gbase[256];              // global start of the digit's run minus its tile-local start (wraps; only sums are used)
    __shared__ uint32_t wsum[8];
    __shared__ uint32_t tile_kept;
    const uint16_t *s_tab = ds_stage_table<(MODE == 0 ? RANK : 0)>(in, s_dyn);
    const uint32_t n = MODE == 0 ? (uint32_t)in.n_windows : *in.n_in;
    const uint32_t n_keep = MODE == 0 ? *in.n_keep : 0u;
    const unsigned lane = threadIdx.x & 31u, wp = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u, me = 1u << lane;
    uint32_t *mrow = reinterpret_cast<uint32_t *>(s_pair) + wp * 256;  // lane masks of the current iteration (s_pair is free until the staging)
    uint16_t *crow = wcnt[wp];
    for (uint32_t tile = blockIdx.x; tile < in.n_tiles; tile += gridDim.x) {
        const uint32_t base = tile * TILE;
        if (base >= n) break;
        for (int d = lane; d < 256; d += 32) {
            crow[d] = 0;
            mrow[d] = 0u;  // every item leaves its mask words zero again
        }
        uint32_t rk[ITEMS], idx[ITEMS];
        uint16_t pre[ITEMS];
#pragma unroll
        for (int k = 0; k < ITEMS; ++k)  // warp wp owns the contiguous slice [wp * 32 * ITEMS, (wp + 1) * 32 * ITEMS) of the tile
            ds_load<MODE, RANK>(in.dense, s_tab, n_keep, in.strands, in.span, in.in_rank, in.in_idx,
                                base + wp * (32 * ITEMS) + k * 32 + lane, n, rk[k], idx[k]);
        __syncwarp();
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            // peers = the lanes of this iteration with the same digit; the first of them moves the digit's counter
            const bool valid = rk[k] != DS_NONE;
            const uint32_t digit = ds_digit<MODE>(rk[k]);
            if (valid) atomicOr(mrow + digit, me);
            __syncwarp();
            const uint32_t peers = valid ? mrow[digit] : 0u;
            const uint32_t c = valid ? crow[digit] : 0u;
            pre[k] = (uint16_t)(c + __popc(peers & lt));
            __syncwarp();
            if (valid && (peers & lt) == 0u) {
                crow[digit] = (uint16_t)(c + __popc(peers));
                mrow[digit] = 0u;
            }
            __syncwarp();
        }
        __syncthreads();
        // per digit: exclusive scan of the warp counts (slices are in index order), then of the digit totals
        uint32_t total = 0;
        if (threadIdx.x < 256) {
#pragma unroll
            for (int v = 0; v < WARPS; ++v) total += wcnt[v][threadIdx.x];
        }
        uint32_t incl = total;
        if (threadIdx.x < 256) {
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if (lane >= (unsigned)o) incl += up;
            }
            if (lane == 31) wsum[wp] = incl;
        }
        __syncthreads();
        if (threadIdx.x < 256) {
            uint32_t before = 0;
            for (unsigned v = 0; v < wp; ++v) before += wsum[v];
            const uint32_t start = before + incl - total;
            uint32_t run = start;
#pragma unroll
            for (int v = 0; v < WARPS; ++v) {
                const uint32_t c = wcnt[v][threadIdx.x];
                wcnt[v][threadIdx.x] = (uint16_t)run;
                run += c;
            }
            gbase[threadIdx.x] = __ldg(table + threadIdx.x * in.n_tiles + tile) - start;
            if (threadIdx.x == 255) tile_kept = start + total;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            if (rk[k] != DS_NONE) {
                const uint32_t pos = (uint32_t)crow[ds_digit<MODE>(rk[k])] + pre[k];
                s_pair[pos] = ((unsigned long long)idx[k] << 16) | rk[k];
            }
        }
        __syncthreads();
        const uint32_t kept = tile_kept;
        constexpr int BATCH = 4;  // loads of a batch first, then its stores: the table reads of one row do not wait for the stores of the previous one
#pragma unroll
        for (int k0 = 0; k0 < ITEMS; k0 += BATCH) {
            if ((uint32_t)k0 * THREADS >= kept) break;
            uint32_t r[BATCH], i[BATCH], pos[BATCH];
#pragma unroll
            for (int k = 0; k < BATCH; ++k) {
                const uint32_t t = (uint32_t)(k0 + k) * THREADS + threadIdx.x;
                const bool ok = t < kept;
                const unsigned long long pr = ok ? s_pair[t] : 0ull;
                r[k] = ok ? (uint32_t)(pr & 0xFFFFull) : DS_NONE;
                i[k] = (uint32_t)(pr >> 16);
                pos[k] = ok ? gbase[ds_digit<MODE>(r[k])] + t : 0u;
            }
            if (MODE == 0) {
#pragma unroll
                for (int k = 0; k < BATCH; ++k)
                    if (r[k] != DS_NONE) {
                        out_rank[pos[k]] = (uint16_t)r[k];
                        out_idx[pos[k]] = i[k];
                    }
            } else {
                uint32_t bin[BATCH];
                double pv[BATCH], qv[BATCH], sv[BATCH];
#pragma unroll
                for (int k = 0; k < BATCH; ++k) bin[k] = r[k] != DS_NONE ? __ldg(col.bin_of_rank + r[k]) : 0u;  // kept windows have bin < span
#pragma unroll
                for (int k = 0; k < BATCH; ++k) {
                    pv[k] = col.o_p != nullptr ? __ldg(col.ptab + bin[k]) : 0.0;
                    sv[k] = col.o_score != nullptr ? __ldg(col.score_tab + bin[k]) : 0.0;
                    qv[k] = (col.o_q != nullptr && col.qtab != nullptr) ? __ldg(col.qtab + bin[k]) : 0.0;
                }
#pragma unroll
                for (int k = 0; k < BATCH; ++k)
                    if (r[k] != DS_NONE) {
                        col.o_row[pos[k]] = col.row_base + (in.strands == 2 ? (uint64_t)(i[k] >> 1) : (uint64_t)i[k]);
                        col.o_strand[pos[k]] = (uint8_t)(in.strands == 2 ? (i[k] & 1u) : 0u);
                        col.o_iscore[pos[k]] = col.lo_score + (int32_t)bin[k];
                        if (col.o_score != nullptr) col.o_score[pos[k]] = sv[k];
                        if (col.o_p != nullptr) col.o_p[pos[k]] = pv[k];
                        if (col.o_q != nullptr && col.qtab != nullptr) col.o_q[pos[k]] = qv[k];
                    }
            }
        }
        __syncthreads();  // the staging array is reused by the next tile
    }
}

__global__ void gb2_ds_total_kernel(const uint32_t *__restrict__ total, uint64_t *__restrict__ n_out)
{
    *n_out = (uint64_t)(*total);
}

int gb2_finalize_dense_cub(gb2_ctx *ctx, const gb2_motif *m, const uint32_t *d_dense, uint64_t n_kmers, int strands,
                           uint64_t row_base, const double *d_qtab, const uint32_t *d_rank, double p_threshold, int q_filter,
                           double q_threshold, uint64_t *d_row, uint8_t *d_strand, int32_t *d_iscore, double *d_score,
                           double *d_p, double *d_q, uint64_t *d_n_out);

#define DS_ITEMS 8
#define DS_CT 4

// Occupancy the scatter kernels are compiled for (measured at 60 M windows): pass A 356 us at 2 CTAs per SM (64 registers),
// 319 us at 3 (40 registers, no spills); pass B with all six columns 599 us at 2, 766 us at 3 (spills).
#define DS_OCC_A 3
#define DS_OCC_B 2

template <int RANK, int THREADS>
static int ds_pass_a(gb2_ctx *ctx, const DsIn &in, uint32_t *tab_a, size_t n_table, void *d_tmp, size_t cub_bytes, uint16_t *mid_rank,
                     uint32_t *mid_idx, const DsColumns &col)
{
    cudaStream_t st = ctx->stream;
    const size_t dyn = RANK == 1 ? (((size_t)in.span + 1) * 2 + 15) & ~(size_t)15 : 0;
    auto kc = gb2_ds_count_kernel<0, RANK, THREADS, DS_ITEMS, DS_CT>;
    auto ks = gb2_ds_scatter_kernel<0, RANK, THREADS, DS_ITEMS, DS_OCC_A>;
    if (dyn) {
        GB2_CUDA(ctx, cudaFuncSetAttribute(kc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        GB2_CUDA(ctx, cudaFuncSetAttribute(ks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    }
    int occ_c = 1, occ_s = 1;
    GB2_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_c, kc, THREADS, dyn));
    GB2_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_s, ks, THREADS, dyn));
    GB2_REQUIRE(ctx, occ_c >= 1 && occ_s >= 1, "gb2_finalize_dense: the rank table does not fit shared memory");
    // persistent grids where the table is staged once per CTA; the arithmetic form has nothing to stage
    const unsigned n_groups = (in.n_tiles + DS_CT - 1) / DS_CT;
    const unsigned grid_c = RANK == 1 ? std::min<unsigned>(n_groups, (unsigned)(ctx->sm_count * occ_c)) : n_groups;
    const unsigned grid_s = RANK == 1 ? std::min<unsigned>(in.n_tiles, (unsigned)(ctx->sm_count * occ_s)) : in.n_tiles;
    kc<<<grid_c, THREADS, dyn, st>>>(in, tab_a);
    GB2_LAUNCH_CHECK(ctx);
    GB2_CUDA(ctx, cub::DeviceScan::ExclusiveSum(d_tmp, cub_bytes, tab_a, tab_a, (int)n_table, st));
    ctx->launches += 1;
    ks<<<grid_s, THREADS, dyn, st>>>(in, tab_a, mid_rank, mid_idx, col);
    GB2_LAUNCH_CHECK(ctx);
    return GB2_OK;
}

template <int THREADS>
static int ds_run(gb2_ctx *ctx, const gb2_motif *m, const uint32_t *d_dense, uint64_t n_windows, int strands, uint64_t row_base,
                  const double *d_qtab, const uint32_t *d_rank, double p_threshold, int q_filter, double q_threshold, uint64_t *d_row,
                  uint8_t *d_strand, int32_t *d_iscore, double *d_score, double *d_p, double *d_q, uint64_t *d_n_out)
{
    constexpr uint32_t TILE = THREADS * DS_ITEMS;
    const uint32_t n_tiles = (uint32_t)gb2_div_up((int64_t)n_windows, TILE);
    const size_t n_table = (size_t)256 * n_tiles + 1;  // the extra element becomes the total after the exclusive scan
    size_t cub_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, (uint32_t *)nullptr, (uint32_t *)nullptr, (int)n_table, ctx->stream);
    auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const uint32_t nbins = (uint32_t)m->span + 1;
    const size_t need = align(cub_bytes) + 2 * align(n_table * 4) + align((size_t)n_windows * 2) + align((size_t)n_windows * 4) +
                        align((size_t)nbins * 2) + align((size_t)nbins * 4) + align((size_t)nbins * 8) + 256;
    int rc = gb2_scratch_reserve(ctx, need);
    if (rc != GB2_OK) return rc;
    char *base = (char *)ctx->scratch;
    void *d_tmp = base; base += align(cub_bytes);
    uint32_t *tab_a = (uint32_t *)base; base += align(n_table * 4);
    uint32_t *tab_b = (uint32_t *)base; base += align(n_table * 4);
    uint16_t *mid_rank = (uint16_t *)base; base += align((size_t)n_windows * 2);
    uint32_t *mid_idx = (uint32_t *)base; base += align((size_t)n_windows * 4);
    uint16_t *rank16 = (uint16_t *)base; base += align((size_t)nbins * 2);
    uint32_t *bin_of_rank = (uint32_t *)base; base += align((size_t)nbins * 4);
    double *score_tab = (double *)base; base += align((size_t)nbins * 8);
    uint32_t *n_keep = (uint32_t *)base;
    cudaStream_t st = ctx->stream;
    GB2_CUDA(ctx, cudaMemsetAsync(n_keep, 0, 4, st));
    GB2_CUDA(ctx, cudaMemsetAsync(tab_a + n_table - 1, 0, 4, st));
    GB2_CUDA(ctx, cudaMemsetAsync(tab_b + n_table - 1, 0, 4, st));
    gb2_ds_tables_kernel<<<(nbins + 255) / 256, 256, 0, st>>>(m->d_ptab, d_qtab, d_rank, (uint32_t)m->span, p_threshold, q_filter,
                                                             q_threshold, (int32_t)m->lo, m->w, (double)m->scale, m->offset, rank16,
                                                             bin_of_rank, score_tab, n_keep);
    GB2_LAUNCH_CHECK(ctx);
    DsColumns col{bin_of_rank, m->d_ptab, d_qtab, score_tab, row_base, (int32_t)m->lo, d_row, d_strand, d_iscore, d_score, d_p, d_q};
    DsIn in{d_dense, rank16, n_keep, n_windows, strands, (uint32_t)m->span, nullptr, nullptr, nullptr, n_tiles};
    // pass A: dense scores -> (rank, window) pairs ordered by (low digit, window)
    const char *force_table = getenv("GB2_DENSE_RANK_TABLE");
    if (m->monotone && !(force_table && force_table[0] == '1'))
        rc = ds_pass_a<0, THREADS>(ctx, in, tab_a, n_table, d_tmp, cub_bytes, mid_rank, mid_idx, col);
    else
        rc = ds_pass_a<1, THREADS>(ctx, in, tab_a, n_table, d_tmp, cub_bytes, mid_rank, mid_idx, col);
    if (rc != GB2_OK) return rc;
    const uint32_t *n_mid = tab_a + n_table - 1;  // kept windows
    gb2_ds_total_kernel<<<1, 1, 0, st>>>(n_mid, d_n_out);
    GB2_LAUNCH_CHECK(ctx);
    // pass B: ordered by (high digit, low digit, window) = (p-rank, row, strand); the report columns are written in place
    in.in_rank = mid_rank;
    in.in_idx = mid_idx;
    in.n_in = n_mid;
    gb2_ds_count_kernel<1, 0, THREADS, DS_ITEMS, DS_CT><<<(n_tiles + DS_CT - 1) / DS_CT, THREADS, 0, st>>>(in, tab_b);
    GB2_LAUNCH_CHECK(ctx);
    GB2_CUDA(ctx, cub::DeviceScan::ExclusiveSum(d_tmp, cub_bytes, tab_b, tab_b, (int)n_table, st));
    ctx->launches += 1;
    gb2_ds_scatter_kernel<1, 0, THREADS, DS_ITEMS, DS_OCC_B><<<n_tiles, THREADS, 0, st>>>(in, tab_b, nullptr, nullptr, col);
    GB2_LAUNCH_CHECK(ctx);
    return GB2_OK;
}

// d_score, d_p and d_q may be null: a caller that prints the report on the device (K8, report.cu) needs only the row, the
// strand and the integer score of every row -- 13 instead of 37 bytes written per row.
extern "C" int gb2_finalize_dense(gb2_ctx *ctx, const gb2_motif *m, const uint32_t *d_dense, uint64_t n_kmers, int strands,
                                  uint64_t row_base, const double *d_qtab, const uint32_t *d_rank, double p_threshold,
                                  int q_filter, double q_threshold, uint64_t *d_row, uint8_t *d_strand, int32_t *d_iscore,
                                  double *d_score, double *d_p, double *d_q, uint64_t *d_n_out)
{
    if (!ctx || !m) return GB2_ERR_ARG;
    const char *force_cub = getenv("GB2_DENSE_CUB");
    if (force_cub && force_cub[0] == '1')
        return gb2_finalize_dense_cub(ctx, m, d_dense, n_kmers, strands, row_base, d_qtab, d_rank, p_threshold, q_filter,
                                      q_threshold, d_row, d_strand, d_iscore, d_score, d_p, d_q, d_n_out);
    GB2_REQUIRE(ctx, d_n_out != nullptr && d_rank != nullptr, "gb2_finalize_dense: null counter or rank table");
    GB2_REQUIRE(ctx, strands == 1 || strands == 2, "gb2_finalize_dense: strands must be 1 or 2");
    GB2_REQUIRE(ctx, !q_filter || d_qtab != nullptr, "gb2_finalize_dense: q filter needs the q table");
    GB2_REQUIRE(ctx, m->span + 1 <= 65536, "gb2_finalize_dense: rank table wider than 16 bits");
    const uint64_t n_windows = n_kmers * (uint64_t)strands;
    GB2_REQUIRE(ctx, n_windows < ((uint64_t)1 << 31), "gb2_finalize_dense: at most 2^31-1 windows per call");
    GB2_CUDA(ctx, cudaSetDevice(ctx->device));
    GB2_CUDA(ctx, cudaMemsetAsync(d_n_out, 0, sizeof(uint64_t), ctx->stream));
    if (n_windows == 0) return GB2_OK;
    GB2_REQUIRE(ctx, d_dense && d_row && d_strand && d_iscore, "gb2_finalize_dense: null buffer");
    return ds_run<512>(ctx, m, d_dense, n_windows, strands, row_base, d_qtab, d_rank, p_threshold, q_filter, q_threshold, d_row,
                       d_strand, d_iscore, d_score, d_p, d_q, d_n_out);
}
