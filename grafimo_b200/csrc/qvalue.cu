// qvalue.cu -- K5 (Benjamini-Hochberg from the score histogram), K6 (finalize hits: filter, CUB radix
// sort, numeric columns) and the haplotype tally (sort + run-length segmented reduction).
//
//   K5  src/grafimo/score_sequences.py:401-428 (statsmodels fdr_bh on every scored row).  The p-value
//       is a function of the integer score, so the multiset of p-values IS the histogram:
//       sort bins by p ascending, C = running count, raw = p / (C / float(N)), q = reverse running
//       minimum clipped at 1.  Tied p-values (several bins, or many rows in one bin) take the value of
//       the last tied rank, exactly like the row-wise formula.  Integer prefix sums and min-scans are
//       exact in any order, so they run as block scans; the two fp64 divisions are IEEE (div.rn).
//   K6  src/grafimo/resultsTmp.py:303-313 (+ score_sequences.py:393 for the log-odds column).
#include <math_constants.h>

#include <algorithm>
#include <vector>

#include <cub/cub.cuh>
#include <thrust/iterator/reverse_iterator.h>

#include "internal.cuh"

// ---------------------------------------------------------------------------------------------
// K5
// ---------------------------------------------------------------------------------------------
__global__ void gb2_bh_keys_kernel(const double *__restrict__ ptab, uint32_t span, double *__restrict__ keys,
                                   uint32_t *__restrict__ bins)
{
    // input order = descending score (ascending p for a monotone table; ties stay in this order)
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > span) return;
    if (i == span) {  // N rows: score = min_val, p = 1
        keys[i] = 1.0;
        bins[i] = span;
    } else {
        const uint32_t b = span - 1u - i;
        keys[i] = ptab[b];
        bins[i] = b;
    }
}

#define BH_THREADS 1024
// K5 runs as three small kernels so that the 2 x (span + 1) fp64 divisions -- the bulk of the work -- spread over the
// whole GPU instead of queueing on one SM (a single-CTA version took 26 us at span 7,424 and 90 us at 22,805):
//   count  one CTA: rank of every bin, cumulative row count C at every sorted position (block scan, exact integers)
//   raw    grid:    raw = p / (C / float(N)) with IEEE div.rn, +inf for empty bins
//   min    one CTA: reverse running minimum (block min-scan, exact), clip at 1, scatter to the bins
// The one-CTA kernels walk the sorted positions in chunks of 1,024 CONSECUTIVE elements (thread t takes element
// chunk * 1024 + t, the chunk total is carried): a thread-owns-a-contiguous-range layout makes every warp access touch 32
// sectors and the single SM's load/store unit then bounds the kernel (measured 43 + 30 us at 22,806 bins; loads in flight
// were not the limit -- batching them changed nothing).
// inclusive scan over the 1,024 threads of the CTA (warp shuffles + one pass over the 32 warp totals), `carry` = what
// came before this chunk; cub::BlockScan's default raking form costs ~2.5 us per call at 1,024 threads
template <typename T, typename Op>
__device__ __forceinline__ T bh_block_scan(T v, T identity, Op op, T *s_warp, T &carry)
{
    const unsigned lane = threadIdx.x & 31u, wp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const T up = __shfl_up_sync(0xFFFFFFFFu, v, o);
        if (lane >= (unsigned)o) v = op(up, v);
    }
    if (lane == 31) s_warp[wp] = v;
    __syncthreads();
    if (wp == 0) {
        T w = s_warp[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const T up = __shfl_up_sync(0xFFFFFFFFu, w, o);
            if (lane >= (unsigned)o) w = op(up, w);
        }
        s_warp[lane] = w;
    }
    __syncthreads();
    const T before = op(carry, wp ? s_warp[wp - 1] : identity);
    const T total = s_warp[BH_THREADS / 32 - 1];
    __syncthreads();  // s_warp is reused by the next chunk
    carry = op(carry, total);
    return op(before, v);
}

struct BhAdd {
    __device__ __forceinline__ unsigned long long operator()(unsigned long long a, unsigned long long b) const { return a + b; }
};
struct BhMin {
    __device__ __forceinline__ double operator()(double a, double b) const { return fmin(a, b); }
};

__global__ void __launch_bounds__(BH_THREADS) gb2_bh_count_kernel(const uint32_t *__restrict__ sorted_bins,
                                                                  const unsigned long long *__restrict__ hist, uint32_t nbins,
                                                                  uint32_t *__restrict__ rank,
                                                                  unsigned long long *__restrict__ cum,
                                                                  unsigned long long *__restrict__ total_out)
{
    static_assert(BH_THREADS == 1024, "bh_block_scan scans 32 warp totals with one warp");
    __shared__ unsigned long long s_warp[32];
    const uint32_t tid = threadIdx.x;
    unsigned long long carry = 0ull;
    // software pipeline: the bin indices run two chunks ahead and the counts they address one chunk ahead, so that neither of
    // the two dependent loads of an element waits inside the chunk that scans it
    uint32_t b_cur = tid < nbins ? sorted_bins[tid] : 0u;
    uint32_t b_next = tid + BH_THREADS < nbins ? sorted_bins[tid + BH_THREADS] : 0u;
    unsigned long long h_cur = (hist != nullptr && tid < nbins) ? hist[b_cur] : 0ull;
    for (uint32_t base = 0; base < nbins; base += BH_THREADS) {
        const uint32_t i = base + tid, b = b_cur;
        const unsigned long long h = h_cur;
        const uint32_t i1 = i + BH_THREADS, i2 = i + 2 * BH_THREADS;
        h_cur = (hist != nullptr && i1 < nbins) ? hist[b_next] : 0ull;
        b_cur = b_next;
        b_next = i2 < nbins ? sorted_bins[i2] : 0u;
        if (i < nbins) rank[b] = i;  // position of every bin in p-ascending order
        if (hist == nullptr) continue;  // rank-only call (no q-values wanted)
        const unsigned long long c = bh_block_scan(h, 0ull, BhAdd(), s_warp, carry);
        if (i < nbins) cum[i] = c;
    }
    if (hist != nullptr && tid == 0) *total_out = carry;
}

__global__ void gb2_bh_raw_kernel(const double *__restrict__ sorted_p, const unsigned long long *__restrict__ cum,
                                  uint32_t nbins, double *__restrict__ raw)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nbins) return;
    const unsigned long long c = cum[i], before = i ? cum[i - 1] : 0ull;
    const double dn = (double)cum[nbins - 1];
    raw[i] = c != before ? __ddiv_rn(sorted_p[i], __ddiv_rn((double)c, dn)) : CUDART_INF;
}

__global__ void __launch_bounds__(BH_THREADS) gb2_bh_min_kernel(const double *__restrict__ raw,
                                                                const uint32_t *__restrict__ sorted_bins, uint32_t nbins,
                                                                double *__restrict__ qtab)
{
    __shared__ double s_warp[32];
    const uint32_t tid = threadIdx.x;
    // right to left: thread t of a chunk takes the element t places before the chunk's right end, so an inclusive
    // min-scan over the threads IS the running minimum from the right; the minimum of the chunks already done is carried
    double carry = CUDART_INF;
    double r_next = tid < nbins ? raw[nbins - 1u - tid] : CUDART_INF;
    uint32_t b_next = tid < nbins ? sorted_bins[nbins - 1u - tid] : 0u;
    for (uint32_t base = 0; base < nbins; base += BH_THREADS) {
        const uint32_t k = base + tid;  // distance from the right end
        const double r = r_next;
        const uint32_t b = b_next;
        const uint32_t kn = k + BH_THREADS;
        r_next = kn < nbins ? raw[nbins - 1u - kn] : CUDART_INF;
        b_next = kn < nbins ? sorted_bins[nbins - 1u - kn] : 0u;
        const double m = bh_block_scan(r, (double)CUDART_INF, BhMin(), s_warp, carry);
        if (k < nbins) qtab[b] = m > 1.0 ? 1.0 : m;
    }
}

extern "C" int gb2_qvalues_from_hist(gb2_ctx *ctx, const gb2_motif *m, const uint64_t *d_hist, double *d_qtab,
                                     uint32_t *d_rank, uint64_t *d_total)
{
    if (!ctx || !m) return GB2_ERR_ARG;
    GB2_REQUIRE(ctx, d_rank != nullptr, "gb2_qvalues_from_hist: null rank buffer");
    GB2_REQUIRE(ctx, d_hist == nullptr || (d_qtab && d_total), "gb2_qvalues_from_hist: histogram given without q/total buffers");
    GB2_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint32_t span = (uint32_t)m->span, nb = span + 1;
    size_t cub_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const double *)nullptr, (double *)nullptr,
                                    (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)nb, 0, 64, ctx->stream);
    auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t need = align(cub_bytes) + 4 * align(nb * sizeof(double)) + 2 * align(nb * sizeof(uint32_t));
    int rc = gb2_scratch_reserve(ctx, need);
    if (rc != GB2_OK) return rc;
    char *base = (char *)ctx->scratch;
    void *d_tmp = base; base += align(cub_bytes);
    double *k_in = (double *)base; base += align(nb * sizeof(double));
    double *k_out = (double *)base; base += align(nb * sizeof(double));
    unsigned long long *cum = (unsigned long long *)base; base += align(nb * sizeof(double));
    double *raw = (double *)base; base += align(nb * sizeof(double));
    uint32_t *v_in = (uint32_t *)base; base += align(nb * sizeof(uint32_t));
    uint32_t *v_out = (uint32_t *)base;
    if (m->monotone) {
        // p is non-increasing in the score: descending score IS ascending p (N bin, p = 1, last) -- no sort needed
        gb2_bh_keys_kernel<<<(nb + 255) / 256, 256, 0, ctx->stream>>>(m->d_ptab, span, k_out, v_out);
        GB2_LAUNCH_CHECK(ctx);
    } else {
        gb2_bh_keys_kernel<<<(nb + 255) / 256, 256, 0, ctx->stream>>>(m->d_ptab, span, k_in, v_in);
        GB2_LAUNCH_CHECK(ctx);
        GB2_CUDA(ctx, cub::DeviceRadixSort::SortPairs(d_tmp, cub_bytes, k_in, k_out, v_in, v_out, (int)nb, 0, 64, ctx->stream));
        ctx->launches += 1;
    }
    gb2_bh_count_kernel<<<1, BH_THREADS, 0, ctx->stream>>>(v_out, (const unsigned long long *)d_hist, nb, d_rank, cum,
                                                           (unsigned long long *)d_total);
    GB2_LAUNCH_CHECK(ctx);
    if (d_hist != nullptr) {
        gb2_bh_raw_kernel<<<(nb + 255) / 256, 256, 0, ctx->stream>>>(k_out, cum, nb, raw);
        GB2_LAUNCH_CHECK(ctx);
        gb2_bh_min_kernel<<<1, BH_THREADS, 0, ctx->stream>>>(raw, v_out, nb, d_qtab);
        GB2_LAUNCH_CHECK(ctx);
    }
    return GB2_OK;
}

// K5 for MANY motifs in one launch (a motif collection scanned over the same k-mers, BASELINE config 3): one CTA per
// motif does what the three kernels above do -- the per-motif form costs three launches, two of them single-CTA, per
// motif, and with 800 motifs the GPU is filled by the motifs themselves.  Motifs whose p-value table is monotone in the
// score (all but pathological ones) need no sort: sorted position i is bin span-1-i, the N bin (p = 1) last.
struct BhMotif {
    const double *ptab;
    long long bin_off;  // first bin of the motif in the concatenated hist / qtab / rank arrays
    uint32_t span;
    uint32_t pad;
};

__global__ void __launch_bounds__(BH_THREADS) gb2_bh_many_kernel(const BhMotif *__restrict__ ms,
                                                                 const unsigned long long *__restrict__ hist_all,
                                                                 double *__restrict__ qtab_all, uint32_t *__restrict__ rank_all,
                                                                 unsigned long long *__restrict__ totals,
                                                                 double *__restrict__ raw_all)
{
    typedef cub::BlockScan<unsigned long long, BH_THREADS> ScanU64;
    typedef cub::BlockScan<double, BH_THREADS> ScanF64;
    __shared__ union {
        typename ScanU64::TempStorage u;
        typename ScanF64::TempStorage f;
    } tmp;
    __shared__ double s_run[BH_THREADS];
    const BhMotif m = ms[blockIdx.x];
    const uint32_t span = m.span, nbins = span + 1u, tid = threadIdx.x;
    const uint32_t per = (nbins + BH_THREADS - 1) / BH_THREADS;
    const uint32_t beg = min(tid * per, nbins), end = min(beg + per, nbins);
    uint32_t *rank = rank_all + m.bin_off;
    auto bin_of = [&](uint32_t i) { return i < span ? span - 1u - i : span; };
    for (uint32_t i = beg; i < end; ++i) rank[bin_of(i)] = i;
    if (hist_all == nullptr) return;  // rank-only call (no q-values wanted)
    const unsigned long long *hist = hist_all + m.bin_off;
    double *qtab = qtab_all + m.bin_off, *raw = raw_all + m.bin_off;
    unsigned long long local = 0;
    for (uint32_t i = beg; i < end; ++i) local += hist[bin_of(i)];
    unsigned long long prefix, total;
    ScanU64(tmp.u).ExclusiveSum(local, prefix, total);
    if (tid == 0) totals[blockIdx.x] = total;
    const double dn = (double)total;
    unsigned long long c = prefix;
    double run_min = CUDART_INF;
    for (uint32_t i = beg; i < end; ++i) {
        const uint32_t b = bin_of(i);
        const unsigned long long h = hist[b];
        c += h;
        const double p = i < span ? m.ptab[b] : 1.0;
        const double r = h ? __ddiv_rn(p, __ddiv_rn((double)c, dn)) : CUDART_INF;  // raw = p / (C / float(N))
        raw[i] = r;
        run_min = fmin(run_min, r);
    }
    // suffix-min over threads: reverse the thread order and take an inclusive min-scan
    double scanned;
    s_run[BH_THREADS - 1 - tid] = run_min;
    __syncthreads();
    const double rev = s_run[tid];
    ScanF64(tmp.f).InclusiveScan(rev, scanned, cub::Min());
    __syncthreads();
    s_run[tid] = scanned;  // s_run[j] = min over original threads >= BH_THREADS-1-j
    __syncthreads();
    double mn = (tid + 1 < BH_THREADS) ? s_run[BH_THREADS - 2 - tid] : CUDART_INF;
    for (uint32_t i = end; i > beg; --i) {  // the thread re-reads only what it wrote itself
        mn = fmin(mn, raw[i - 1]);
        qtab[bin_of(i - 1)] = mn > 1.0 ? 1.0 : mn;
    }
}

extern "C" int gb2_qvalues_from_hist_many(gb2_ctx *ctx, int32_t n_motifs, const gb2_motif *const *motifs,
                                          const int64_t *h_bin_off, const uint64_t *d_hist, double *d_qtab, uint32_t *d_rank,
                                          uint64_t *d_totals)
{
    if (!ctx) return GB2_ERR_ARG;
    GB2_REQUIRE(ctx, n_motifs >= 0, "gb2_qvalues_from_hist_many: negative motif count");
    if (n_motifs == 0) return GB2_OK;
    GB2_REQUIRE(ctx, motifs && h_bin_off && d_rank, "gb2_qvalues_from_hist_many: null argument");
    GB2_REQUIRE(ctx, d_hist == nullptr || (d_qtab && d_totals), "gb2_qvalues_from_hist_many: histogram given without q/total buffers");
    GB2_CUDA(ctx, cudaSetDevice(ctx->device));
    std::vector<BhMotif> fast;
    std::vector<int> fast_idx, slow_idx;
    for (int i = 0; i < n_motifs; ++i) {
        const gb2_motif *m = motifs[i];
        GB2_REQUIRE(ctx, m != nullptr && m->device == ctx->device, "gb2_qvalues_from_hist_many: bad motif %d", i);
        GB2_REQUIRE(ctx, h_bin_off[i + 1] - h_bin_off[i] == m->span + 1, "gb2_qvalues_from_hist_many: motif %d needs %lld bins", i,
                    (long long)(m->span + 1));
        if (m->monotone) {
            BhMotif b;
            b.ptab = m->d_ptab; b.bin_off = h_bin_off[i]; b.span = (uint32_t)m->span; b.pad = 0;
            fast.push_back(b);
            fast_idx.push_back(i);
        } else {
            slow_idx.push_back(i);
        }
    }
    for (int i : slow_idx) {  // p-value table not monotone: the sorting form, one motif at a time
        const int64_t o = h_bin_off[i];
        int rc = gb2_qvalues_from_hist(ctx, motifs[i], d_hist ? d_hist + o : nullptr, d_qtab ? d_qtab + o : nullptr, d_rank + o,
                                       d_totals ? d_totals + i : nullptr);
        if (rc != GB2_OK) return rc;
    }
    if (fast.empty()) return GB2_OK;
    // per-motif totals of the fast group land in a compact array first (the kernel indexes by CTA), then go to their slots
    auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t b_desc = align(fast.size() * sizeof(BhMotif));
    const size_t b_raw = align((size_t)h_bin_off[n_motifs] * sizeof(double));
    const size_t b_tot = align(fast.size() * sizeof(uint64_t));
    int rc = gb2_scratch_reserve(ctx, b_desc + b_raw + b_tot);
    if (rc != GB2_OK) return rc;
    char *base = (char *)ctx->scratch;
    BhMotif *d_desc = (BhMotif *)base;
    double *d_raw = (double *)(base + b_desc);
    unsigned long long *d_tot = (unsigned long long *)(base + b_desc + b_raw);
    GB2_CUDA(ctx, cudaMemcpyAsync(d_desc, fast.data(), fast.size() * sizeof(BhMotif), cudaMemcpyHostToDevice, ctx->stream));
    const bool compact = (int)fast.size() != n_motifs;
    gb2_bh_many_kernel<<<(unsigned)fast.size(), BH_THREADS, 0, ctx->stream>>>(
        d_desc, (const unsigned long long *)d_hist, d_qtab, d_rank, compact ? d_tot : (unsigned long long *)d_totals, d_raw);
    GB2_LAUNCH_CHECK(ctx);
    if (compact && d_hist != nullptr) {
        for (size_t k = 0; k < fast_idx.size(); ++k)
            GB2_CUDA(ctx, cudaMemcpyAsync(d_totals + fast_idx[k], d_tot + k, sizeof(uint64_t), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    return GB2_OK;
}

// ---------------------------------------------------------------------------------------------
// B3: stand-alone Benjamini-Hochberg of an arbitrary p-value list (row-wise form of the same formula)
// ---------------------------------------------------------------------------------------------
__global__ void gb2_bh_iota_kernel(uint32_t *idx, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) idx[i] = i;
}

__global__ void gb2_bh_raw_kernel(const double *__restrict__ sorted_p, uint32_t n, double *__restrict__ raw)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) raw[i] = __ddiv_rn(sorted_p[i], __ddiv_rn((double)(i + 1u), (double)n));  // p / (k / float(n))
}

__global__ void gb2_bh_scatter_kernel(const double *__restrict__ cummin, const uint32_t *__restrict__ order, uint32_t n,
                                      double *__restrict__ q)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) q[order[i]] = cummin[i] > 1.0 ? 1.0 : cummin[i];
}

extern "C" int gb2_bh_pvalues(gb2_ctx *ctx, const double *h_p, int64_t n, double *h_q)
{
    if (!ctx) return GB2_ERR_ARG;
    GB2_REQUIRE(ctx, n >= 0 && n < ((int64_t)1 << 31), "gb2_bh_pvalues: row count must be below 2^31");
    if (n == 0) return GB2_OK;
    GB2_REQUIRE(ctx, h_p && h_q, "gb2_bh_pvalues: null buffer");
    GB2_CUDA(ctx, cudaSetDevice(ctx->device));
    const int ni = (int)n;
    size_t cub_sort = 0, cub_scan = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cub_sort, (const double *)nullptr, (double *)nullptr, (const uint32_t *)nullptr,
                                    (uint32_t *)nullptr, ni, 0, 64, ctx->stream);
    {
        thrust::reverse_iterator<const double *> rin((const double *)nullptr);
        thrust::reverse_iterator<double *> rout((double *)nullptr);
        cub::DeviceScan::InclusiveScan(nullptr, cub_scan, rin, rout, cub::Min(), ni, ctx->stream);
    }
    auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t cub_bytes = std::max(cub_sort, cub_scan);
    const size_t need = align(cub_bytes) + 3 * align((size_t)n * 8) + 2 * align((size_t)n * 4);
    int rc = gb2_scratch_reserve(ctx, need);
    if (rc != GB2_OK) return rc;
    char *base = (char *)ctx->scratch;
    void *d_tmp = base; base += align(cub_bytes);
    double *a = (double *)base; base += align((size_t)n * 8);
    double *b = (double *)base; base += align((size_t)n * 8);
    double *c = (double *)base; base += align((size_t)n * 8);
    uint32_t *i_in = (uint32_t *)base; base += align((size_t)n * 4);
    uint32_t *i_out = (uint32_t *)base;
    const int threads = 256;
    const unsigned blocks = (unsigned)gb2_div_up(n, threads);
    GB2_CUDA(ctx, cudaMemcpyAsync(a, h_p, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    gb2_bh_iota_kernel<<<blocks, threads, 0, ctx->stream>>>(i_in, (uint32_t)n);
    GB2_LAUNCH_CHECK(ctx);
    GB2_CUDA(ctx, cub::DeviceRadixSort::SortPairs(d_tmp, cub_sort, a, b, i_in, i_out, ni, 0, 64, ctx->stream));
    gb2_bh_raw_kernel<<<blocks, threads, 0, ctx->stream>>>(b, (uint32_t)n, a);
    GB2_LAUNCH_CHECK(ctx);
    {
        thrust::reverse_iterator<const double *> rin((const double *)a + n);
        thrust::reverse_iterator<double *> rout(c + n);
        GB2_CUDA(ctx, cub::DeviceScan::InclusiveScan(d_tmp, cub_scan, rin, rout, cub::Min(), ni, ctx->stream));
    }
    ctx->launches += 2;
    gb2_bh_scatter_kernel<<<blocks, threads, 0, ctx->stream>>>(c, i_out, (uint32_t)n, b);
    GB2_LAUNCH_CHECK(ctx);
    GB2_CUDA(ctx, cudaMemcpyAsync(h_q, b, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    GB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GB2_OK;
}

// ---------------------------------------------------------------------------------------------
// K6
// ---------------------------------------------------------------------------------------------
// Kept-row count of a key kernel: one global atomic per CTA (a per-warp atomic on the single counter serialises in L2:
// 1.5 ms for the 1.9 M warps of a 60 M-row unselective scan, ncu launch list of round 1).
__device__ __forceinline__ void block_count_add(unsigned keep_count, unsigned long long *__restrict__ n_kept)
{
    __shared__ unsigned s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    unsigned c = keep_count;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, d);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_cnt, c);
    __syncthreads();
    if (threadIdx.x == 0 && s_cnt) atomicAdd(n_kept, (unsigned long long)s_cnt);
}

__global__ void gb2_hit_keys_kernel(const gb2_hit *__restrict__ hits, uint64_t n, int32_t lo, uint32_t span,
                                    const double *__restrict__ ptab, const double *__restrict__ qtab,
                                    const uint32_t *__restrict__ rank, double p_thr, int q_filter, double q_thr,
                                    int row_bits, unsigned long long *__restrict__ keys, uint32_t *__restrict__ idx,
                                    unsigned long long *__restrict__ n_kept)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool keep = false;
    if (i < n) {
        const gb2_hit h = hits[i];
        const uint32_t bin = (uint32_t)(h.score - lo);
        const double p = bin < span ? ptab[bin] : 1.0;  // bin == span: N row (score = min_val, p = 1)
        keep = p < p_thr && (!q_filter || qtab[bin] < q_thr);  // strict, resultsTmp.py:305-307
        // (p rank, row, strand): p ascending, then a deterministic order among equal p
        const unsigned long long row = h.row & ((1ull << row_bits) - 1ull);
        unsigned long long key = ((unsigned long long)rank[bin] << (row_bits + 1)) | (row << 1) | (h.strand & 1u);
        keys[i] = keep ? key : ~0ull;
        idx[i] = (uint32_t)i;
    }
    block_count_add(keep ? 1u : 0u, n_kept);
}

__global__ void gb2_hit_gather_kernel(const gb2_hit *__restrict__ hits, const uint32_t *__restrict__ order,
                                      const unsigned long long *__restrict__ n_kept, int32_t lo, uint32_t span, int w,
                                      double scale, double offset, const double *__restrict__ ptab, const double *__restrict__ qtab,
                                      uint64_t *__restrict__ o_row, uint8_t *__restrict__ o_strand,
                                      int32_t *__restrict__ o_iscore, double *__restrict__ o_score,
                                      double *__restrict__ o_p, double *__restrict__ o_q)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *n_kept) return;
    const gb2_hit h = hits[order[i]];
    const uint32_t bin = (uint32_t)(h.score - lo);
    o_row[i] = h.row;
    o_strand[i] = (uint8_t)h.strand;
    o_iscore[i] = h.score;
    // logodds = (score / scale) + (width * offset)            score_sequences.py:393
    o_score[i] = __dadd_rn(__ddiv_rn((double)h.score, scale), __dmul_rn((double)w, offset));
    o_p[i] = bin < span ? ptab[bin] : 1.0;
    if (o_q != nullptr && qtab != nullptr) o_q[i] = qtab[bin];
}

extern "C" int gb2_finalize_hits(gb2_ctx *ctx, const gb2_motif *m, const gb2_hit *d_hits, uint64_t n_hits,
                                 uint64_t row_limit, const double *d_qtab, const uint32_t *d_rank, double p_threshold,
                                 int q_filter, double q_threshold, uint64_t *d_row, uint8_t *d_strand, int32_t *d_iscore,
                                 double *d_score, double *d_p, double *d_q, uint64_t *d_n_out)
{
    if (!ctx || !m) return GB2_ERR_ARG;
    GB2_REQUIRE(ctx, d_n_out != nullptr && d_rank != nullptr, "gb2_finalize_hits: null counter or rank table");
    GB2_REQUIRE(ctx, !q_filter || d_qtab != nullptr, "gb2_finalize_hits: q filter needs the q table");
    GB2_REQUIRE(ctx, n_hits < ((uint64_t)1 << 31), "gb2_finalize_hits: at most 2^31-1 hits per call");
    GB2_CUDA(ctx, cudaSetDevice(ctx->device));
    GB2_CUDA(ctx, cudaMemsetAsync(d_n_out, 0, sizeof(uint64_t), ctx->stream));
    if (n_hits == 0) return GB2_OK;
    GB2_REQUIRE(ctx, d_hits && d_row && d_strand && d_iscore && d_score && d_p, "gb2_finalize_hits: null buffer");
    const int n = (int)n_hits;
    // sort key = [p rank | row | strand], using only the bits that can be set (fewer radix passes)
    int rank_bits = 1, row_bits = 47;
    while ((1ll << rank_bits) < m->span + 1) ++rank_bits;
    if (row_limit > 0) {
        row_bits = 1;
        while (row_bits < 47 && (1ull << row_bits) < row_limit) ++row_bits;
    }
    const int end_bit = rank_bits + row_bits + 1;  // <= 16 + 47 + 1
    size_t cub_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                    (const uint32_t *)nullptr, (uint32_t *)nullptr, n, 0, 64, ctx->stream);
    auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t need = align(cub_bytes) + 2 * align((size_t)n * 8) + 2 * align((size_t)n * 4);
    int rc = gb2_scratch_reserve(ctx, need);
    if (rc != GB2_OK) return rc;
    char *base = (char *)ctx->scratch;
    void *d_tmp = base; base += align(cub_bytes);
    unsigned long long *k_in = (unsigned long long *)base; base += align((size_t)n * 8);
    unsigned long long *k_out = (unsigned long long *)base; base += align((size_t)n * 8);
    uint32_t *v_in = (uint32_t *)base; base += align((size_t)n * 4);
    uint32_t *v_out = (uint32_t *)base;
    const int threads = 256;
    const unsigned blocks = (unsigned)gb2_div_up(n, threads);
    gb2_hit_keys_kernel<<<blocks, threads, 0, ctx->stream>>>(d_hits, n_hits, (int32_t)m->lo, (uint32_t)m->span, m->d_ptab,
                                                            d_qtab, d_rank, p_threshold, q_filter, q_threshold, row_bits,
                                                            k_in, v_in, (unsigned long long *)d_n_out);
    GB2_LAUNCH_CHECK(ctx);
    // dropped hits carry the all-ones key: they sort last for any end_bit because kept keys are < 2^end_bit
    GB2_CUDA(ctx, cub::DeviceRadixSort::SortPairs(d_tmp, cub_bytes, k_in, k_out, v_in, v_out, n, 0,
                                                  end_bit < 64 ? end_bit + 1 : 64, ctx->stream));
    ctx->launches += 1;
    gb2_hit_gather_kernel<<<blocks, threads, 0, ctx->stream>>>(d_hits, v_out, (const unsigned long long *)d_n_out,
                                                              (int32_t)m->lo, (uint32_t)m->span, m->w, (double)m->scale,
                                                              m->offset, m->d_ptab, d_qtab, d_row, d_strand, d_iscore, d_score,
                                                              d_p, d_q);
    GB2_LAUNCH_CHECK(ctx);
    return GB2_OK;
}

// ---------------------------------------------------------------------------------------------
// K6, dense form: the report rows of an unselective scan (`-t 1`, what docs/paper_results/run_analysis.sh runs: every
// scored window with p < 1 is reported) straight from K2's dense scores -- no hit records, no 64-bit sort keys.
// Window i = k-mer i / strands, strand i % strands, i.e. the windows are already in (row, strand) order, so ONE
// stable radix sort on the p-rank alone (<= 17 bits, 2-3 passes of 4-byte keys) gives the (p, row, strand) order the
// hit path gets from sorting 41-bit keys.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t dense_bin(const uint32_t *__restrict__ dense, uint64_t i, int strands, uint32_t span)
{
    const uint32_t d = __ldg(dense + (strands == 2 ? (i >> 1) : i));
    if (d == 0xFFFFFFFFu) return span;  // N row: score = min_val, p = 1
    return (strands == 2 && (i & 1ull)) ? (d >> 16) : (d & 0xFFFFu);
}

__global__ void __launch_bounds__(256) gb2_dense_keys_kernel(const uint32_t *__restrict__ dense, uint64_t n_windows, int strands,
                                                             uint32_t span, const double *__restrict__ ptab,
                                                             const double *__restrict__ qtab, const uint32_t *__restrict__ rank,
                                                             double p_thr, int q_filter, double q_thr, uint32_t drop_key,
                                                             uint32_t *__restrict__ keys, uint32_t *__restrict__ idx,
                                                             unsigned long long *__restrict__ n_kept)
{
    unsigned kept = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_windows; i += stride) {
        const uint32_t bin = dense_bin(dense, i, strands, span);
        const double p = bin < span ? ptab[bin] : 1.0;
        const bool keep = p < p_thr && (!q_filter || qtab[bin] < q_thr);  // strict, resultsTmp.py:305-307
        keys[i] = keep ? rank[bin] : drop_key;
        idx[i] = (uint32_t)i;
        kept += keep ? 1u : 0u;
    }
    block_count_add(kept, n_kept);
}

__global__ void gb2_rank_inverse_kernel(const uint32_t *__restrict__ rank, uint32_t nbins, uint32_t *__restrict__ bin_of_rank)
{
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nbins) bin_of_rank[rank[b]] = b;
}

// The sorted key of a kept window IS its p-rank, so its bin comes from the (tiny) inverse rank table: every read of this
// kernel is coalesced or cache-resident, nothing goes back to the dense scores.
__global__ void gb2_dense_gather_kernel(const uint32_t *__restrict__ sorted_rank, const uint32_t *__restrict__ order,
                                        const uint32_t *__restrict__ bin_of_rank,
                                        const unsigned long long *__restrict__ n_kept, int strands, uint64_t row_base,
                                        int32_t lo, int w, double scale, double offset,
                                        const double *__restrict__ ptab, const double *__restrict__ qtab,
                                        uint64_t *__restrict__ o_row, uint8_t *__restrict__ o_strand,
                                        int32_t *__restrict__ o_iscore, double *__restrict__ o_score,
                                        double *__restrict__ o_p, double *__restrict__ o_q)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= *n_kept) return;
    const uint64_t i = order[t];
    const uint32_t bin = bin_of_rank[sorted_rank[t]];  // kept windows have bin < span
    const int32_t sc = lo + (int32_t)bin;
    o_row[t] = row_base + (strands == 2 ? (i >> 1) : i);
    o_strand[t] = (uint8_t)(strands == 2 ? (i & 1ull) : 0ull);
    o_iscore[t] = sc;
    if (o_score != nullptr) o_score[t] = __dadd_rn(__ddiv_rn((double)sc, scale), __dmul_rn((double)w, offset));  // score_sequences.py:393
    if (o_p != nullptr) o_p[t] = ptab[bin];
    if (o_q != nullptr && qtab != nullptr) o_q[t] = qtab[bin];
}

// library-sort form; gb2_finalize_dense (dense_sort.cu) is the product path, this one its checker (GB2_DENSE_CUB=1)
int gb2_finalize_dense_cub(gb2_ctx *ctx, const gb2_motif *m, const uint32_t *d_dense, uint64_t n_kmers, int strands,
                                  uint64_t row_base, const double *d_qtab, const uint32_t *d_rank, double p_threshold,
                                  int q_filter, double q_threshold, uint64_t *d_row, uint8_t *d_strand, int32_t *d_iscore,
                                  double *d_score, double *d_p, double *d_q, uint64_t *d_n_out)
{
    if (!ctx || !m) return GB2_ERR_ARG;
    GB2_REQUIRE(ctx, d_n_out != nullptr && d_rank != nullptr, "gb2_finalize_dense_cub: null counter or rank table");
    GB2_REQUIRE(ctx, strands == 1 || strands == 2, "gb2_finalize_dense_cub: strands must be 1 or 2");
    GB2_REQUIRE(ctx, !q_filter || d_qtab != nullptr, "gb2_finalize_dense_cub: q filter needs the q table");
    const uint64_t n_windows = n_kmers * (uint64_t)strands;
    GB2_REQUIRE(ctx, n_windows < ((uint64_t)1 << 31), "gb2_finalize_dense_cub: at most 2^31-1 windows per call");
    GB2_CUDA(ctx, cudaSetDevice(ctx->device));
    GB2_CUDA(ctx, cudaMemsetAsync(d_n_out, 0, sizeof(uint64_t), ctx->stream));
    if (n_windows == 0) return GB2_OK;
    GB2_REQUIRE(ctx, d_dense && d_row && d_strand && d_iscore, "gb2_finalize_dense_cub: null buffer");
    const int n = (int)n_windows;
    int rank_bits = 1;
    while ((1ll << rank_bits) < m->span + 1) ++rank_bits;  // ranks are < span + 1 <= 2^rank_bits
    const uint32_t drop_key = 1u << rank_bits;             // sorts after every kept window
    size_t cub_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr, (const uint32_t *)nullptr,
                                    (uint32_t *)nullptr, n, 0, rank_bits + 1, ctx->stream);
    auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const uint32_t nbins = (uint32_t)m->span + 1;
    const size_t need = align(cub_bytes) + 4 * align((size_t)n * 4) + align((size_t)nbins * 4);
    int rc = gb2_scratch_reserve(ctx, need);
    if (rc != GB2_OK) return rc;
    char *base = (char *)ctx->scratch;
    void *d_tmp = base; base += align(cub_bytes);
    uint32_t *k_in = (uint32_t *)base; base += align((size_t)n * 4);
    uint32_t *k_out = (uint32_t *)base; base += align((size_t)n * 4);
    uint32_t *v_in = (uint32_t *)base; base += align((size_t)n * 4);
    uint32_t *v_out = (uint32_t *)base; base += align((size_t)n * 4);
    uint32_t *bin_of_rank = (uint32_t *)base;
    const int threads = 256;
    const unsigned blocks = (unsigned)gb2_div_up(n, threads);
    gb2_rank_inverse_kernel<<<(nbins + 255) / 256, 256, 0, ctx->stream>>>(d_rank, nbins, bin_of_rank);
    GB2_LAUNCH_CHECK(ctx);
    const unsigned key_blocks = (unsigned)std::min<int64_t>(blocks, (int64_t)ctx->sm_count * 16);  // grid-stride: one atomic per CTA
    gb2_dense_keys_kernel<<<key_blocks, threads, 0, ctx->stream>>>(d_dense, n_windows, strands, (uint32_t)m->span, m->d_ptab, d_qtab,
                                                              d_rank, p_threshold, q_filter, q_threshold, drop_key, k_in, v_in,
                                                              (unsigned long long *)d_n_out);
    GB2_LAUNCH_CHECK(ctx);
    GB2_CUDA(ctx, cub::DeviceRadixSort::SortPairs(d_tmp, cub_bytes, k_in, k_out, v_in, v_out, n, 0, rank_bits + 1, ctx->stream));
    ctx->launches += 1;
    gb2_dense_gather_kernel<<<blocks, threads, 0, ctx->stream>>>(k_out, v_out, bin_of_rank, (const unsigned long long *)d_n_out,
                                                                strands, row_base, (int32_t)m->lo, m->w, (double)m->scale,
                                                                m->offset, m->d_ptab, d_qtab, d_row, d_strand, d_iscore,
                                                                d_score, d_p, d_q);
    GB2_LAUNCH_CHECK(ctx);
    return GB2_OK;
}

// ---------------------------------------------------------------------------------------------
// K6 over many motifs at once (a JASPAR-sized collection scanned over the same k-mers): the hits of all motifs sit in
// ONE buffer -- gb2_score was given row_base = motif index << 40 -- and are filtered, sorted and annotated by one key
// kernel, one radix sort keyed by motif | p-rank | row | strand and one gather, instead of a sort and two host
// round trips per motif (which cost more than the scoring itself when a motif has ~10^4 hits).
// ---------------------------------------------------------------------------------------------
struct ManyDesc {
    const double *ptab, *qtab;
    const uint32_t *rank;
    int32_t lo;
    uint32_t span;
    int32_t w;
    int32_t pad;
    double scale, offset;
};

#define GB2_MANY_ROW_BITS 40

__global__ void gb2_hit_keys_many_kernel(const gb2_hit *__restrict__ hits, uint64_t n, const ManyDesc *__restrict__ desc,
                                         int n_motifs, double p_thr, int q_filter, double q_thr, int rank_bits, int row_bits,
                                         unsigned long long *__restrict__ keys, uint32_t *__restrict__ idx,
                                         unsigned long long *__restrict__ n_kept)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool keep = false;
    if (i < n) {
        const gb2_hit h = hits[i];
        const uint32_t m = (uint32_t)(h.row >> GB2_MANY_ROW_BITS);
        unsigned long long key = ~0ull;
        if (m < (uint32_t)n_motifs) {
            const ManyDesc d = desc[m];
            const uint32_t bin = (uint32_t)(h.score - d.lo);
            const double p = bin < d.span ? d.ptab[bin] : 1.0;
            keep = p < p_thr && (!q_filter || d.qtab[bin] < q_thr);
            const unsigned long long row = h.row & ((1ull << row_bits) - 1ull);
            key = ((unsigned long long)m << (rank_bits + row_bits + 1)) | ((unsigned long long)d.rank[bin] << (row_bits + 1)) |
                  (row << 1) | (h.strand & 1u);
        }
        keys[i] = keep ? key : ~0ull;
        idx[i] = (uint32_t)i;
    }
    block_count_add(keep ? 1u : 0u, n_kept);
}

__global__ void gb2_hit_gather_many_kernel(const gb2_hit *__restrict__ hits, const uint32_t *__restrict__ order,
                                           const unsigned long long *__restrict__ n_kept, const ManyDesc *__restrict__ desc,
                                           uint32_t *__restrict__ o_motif, uint64_t *__restrict__ o_row,
                                           uint8_t *__restrict__ o_strand, int32_t *__restrict__ o_iscore,
                                           double *__restrict__ o_score, double *__restrict__ o_p, double *__restrict__ o_q)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *n_kept) return;
    const gb2_hit h = hits[order[i]];
    const uint32_t m = (uint32_t)(h.row >> GB2_MANY_ROW_BITS);
    const ManyDesc d = desc[m];
    const uint32_t bin = (uint32_t)(h.score - d.lo);
    o_motif[i] = m;
    o_row[i] = h.row & ((1ull << GB2_MANY_ROW_BITS) - 1ull);
    o_strand[i] = (uint8_t)h.strand;
    o_iscore[i] = h.score;
    o_score[i] = __dadd_rn(__ddiv_rn((double)h.score, d.scale), __dmul_rn((double)d.w, d.offset));  // score_sequences.py:393
    o_p[i] = bin < d.span ? d.ptab[bin] : 1.0;
    if (o_q != nullptr && d.qtab != nullptr) o_q[i] = d.qtab[bin];
}

extern "C" int gb2_finalize_hits_many(gb2_ctx *ctx, int32_t n_motifs, const gb2_motif *const *motifs, const double *const *d_qtabs,
                                      const uint32_t *const *d_ranks, const gb2_hit *d_hits, uint64_t n_hits, uint64_t row_limit,
                                      double p_threshold, int q_filter, double q_threshold, uint32_t *d_motif, uint64_t *d_row,
                                      uint8_t *d_strand, int32_t *d_iscore, double *d_score, double *d_p, double *d_q,
                                      uint64_t *d_n_out)
{
    if (!ctx || !motifs || !d_ranks) return GB2_ERR_ARG;
    GB2_REQUIRE(ctx, d_n_out != nullptr, "gb2_finalize_hits_many: null counter");
    GB2_REQUIRE(ctx, n_motifs >= 1 && n_motifs <= (1 << 20), "gb2_finalize_hits_many: motif count out of range");
    GB2_REQUIRE(ctx, !q_filter || d_qtabs != nullptr, "gb2_finalize_hits_many: q filter needs the q tables");
    GB2_REQUIRE(ctx, n_hits < ((uint64_t)1 << 31), "gb2_finalize_hits_many: at most 2^31-1 hits per call");
    GB2_REQUIRE(ctx, row_limit > 0 && row_limit <= ((uint64_t)1 << GB2_MANY_ROW_BITS), "gb2_finalize_hits_many: rows must be below 2^40");
    GB2_CUDA(ctx, cudaSetDevice(ctx->device));
    GB2_CUDA(ctx, cudaMemsetAsync(d_n_out, 0, sizeof(uint64_t), ctx->stream));
    if (n_hits == 0) return GB2_OK;
    GB2_REQUIRE(ctx, d_hits && d_motif && d_row && d_strand && d_iscore && d_score && d_p, "gb2_finalize_hits_many: null buffer");
    std::vector<ManyDesc> desc((size_t)n_motifs);
    int64_t max_span = 1;
    for (int m = 0; m < n_motifs; ++m) {
        const gb2_motif *mo = motifs[m];
        GB2_REQUIRE(ctx, mo != nullptr && d_ranks[m] != nullptr, "gb2_finalize_hits_many: null motif or rank table (%d)", m);
        GB2_REQUIRE(ctx, mo->device == ctx->device, "gb2_finalize_hits_many: motif %d lives on another device", m);
        desc[(size_t)m] = ManyDesc{mo->d_ptab, d_qtabs ? d_qtabs[m] : nullptr, d_ranks[m], (int32_t)mo->lo, (uint32_t)mo->span,
                                   mo->w, 0, (double)mo->scale, mo->offset};
        GB2_REQUIRE(ctx, !q_filter || desc[(size_t)m].qtab != nullptr, "gb2_finalize_hits_many: q filter needs the q table of motif %d", m);
        max_span = std::max(max_span, mo->span);
    }
    int rank_bits = 1, row_bits = 1, motif_bits = 1;
    while ((1ll << rank_bits) < max_span + 1) ++rank_bits;
    while (row_bits < GB2_MANY_ROW_BITS && (1ull << row_bits) < row_limit) ++row_bits;
    while ((1ll << motif_bits) < n_motifs) ++motif_bits;
    const int end_bit = motif_bits + rank_bits + row_bits + 1;
    GB2_REQUIRE(ctx, end_bit <= 63, "gb2_finalize_hits_many: %d motifs x %lld rows do not fit one 64-bit sort key", n_motifs, (long long)row_limit);
    const int n = (int)n_hits;
    size_t cub_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                    (const uint32_t *)nullptr, (uint32_t *)nullptr, n, 0, 64, ctx->stream);
    auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t b_desc = align(desc.size() * sizeof(ManyDesc));
    const size_t need = align(cub_bytes) + b_desc + 2 * align((size_t)n * 8) + 2 * align((size_t)n * 4);
    int rc = gb2_scratch_reserve(ctx, need);
    if (rc != GB2_OK) return rc;
    char *base = (char *)ctx->scratch;
    void *d_tmp = base; base += align(cub_bytes);
    ManyDesc *d_desc = (ManyDesc *)base; base += b_desc;
    unsigned long long *k_in = (unsigned long long *)base; base += align((size_t)n * 8);
    unsigned long long *k_out = (unsigned long long *)base; base += align((size_t)n * 8);
    uint32_t *v_in = (uint32_t *)base; base += align((size_t)n * 4);
    uint32_t *v_out = (uint32_t *)base;
    GB2_CUDA(ctx, cudaMemcpyAsync(d_desc, desc.data(), desc.size() * sizeof(ManyDesc), cudaMemcpyHostToDevice, ctx->stream));
    GB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // desc is a stack-lifetime host buffer
    const int threads = 256;
    const unsigned blocks = (unsigned)gb2_div_up(n, threads);
    gb2_hit_keys_many_kernel<<<blocks, threads, 0, ctx->stream>>>(d_hits, n_hits, d_desc, n_motifs, p_threshold, q_filter, q_threshold,
                                                                 rank_bits, row_bits, k_in, v_in, (unsigned long long *)d_n_out);
    GB2_LAUNCH_CHECK(ctx);
    GB2_CUDA(ctx, cub::DeviceRadixSort::SortPairs(d_tmp, cub_bytes, k_in, k_out, v_in, v_out, n, 0, end_bit + 1, ctx->stream));
    ctx->launches += 1;
    gb2_hit_gather_many_kernel<<<blocks, threads, 0, ctx->stream>>>(d_hits, v_out, (const unsigned long long *)d_n_out, d_desc, d_motif,
                                                                   d_row, d_strand, d_iscore, d_score, d_p, d_q);
    GB2_LAUNCH_CHECK(ctx);
    return GB2_OK;
}

// ---------------------------------------------------------------------------------------------
// haplotype tally
// ---------------------------------------------------------------------------------------------
__global__ void gb2_tally_heads_kernel(const unsigned long long *__restrict__ pos,
                                       const unsigned long long *__restrict__ packed, int64_t n,
                                       uint32_t *__restrict__ flags)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    flags[i] = (i == 0 || pos[i] != pos[i - 1] || packed[i] != packed[i - 1]) ? 1u : 0u;
}

__global__ void gb2_tally_emit_kernel(const unsigned long long *__restrict__ pos,
                                      const unsigned long long *__restrict__ packed, int64_t n,
                                      const uint32_t *__restrict__ flags, const unsigned long long *__restrict__ scan,
                                      const unsigned long long *__restrict__ ref, unsigned long long pos_base,
                                      int64_t n_ref, unsigned long long *__restrict__ u_pos,
                                      unsigned long long *__restrict__ u_packed, unsigned long long *__restrict__ heads,
                                      uint8_t *__restrict__ u_isref, unsigned long long *__restrict__ n_unique)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (flags[i]) {
        const unsigned long long u = scan[i] - 1ull;
        u_pos[u] = pos[i];
        u_packed[u] = packed[i];
        heads[u] = (unsigned long long)i;
        uint8_t isref = 0;
        if (ref != nullptr && pos[i] >= pos_base && (int64_t)(pos[i] - pos_base) < n_ref)
            isref = ref[pos[i] - pos_base] == packed[i];
        u_isref[u] = isref;
    }
    if (i == n - 1) *n_unique = scan[i];
}

__global__ void gb2_tally_freq_kernel(const unsigned long long *__restrict__ heads,
                                      const unsigned long long *__restrict__ n_unique, int64_t n,
                                      uint32_t *__restrict__ freq)
{
    const unsigned long long u = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long nu = *n_unique;
    if (u >= nu) return;
    const unsigned long long next = (u + 1 < nu) ? heads[u + 1] : (unsigned long long)n;
    freq[u] = (uint32_t)(next - heads[u]);
}

extern "C" int gb2_tally_haplotypes(gb2_ctx *ctx, uint64_t *d_pos, uint64_t *d_packed, int64_t n,
                                    const uint64_t *d_ref_packed, uint64_t pos_base, int64_t n_ref, uint64_t *d_u_pos,
                                    uint64_t *d_u_packed, uint32_t *d_u_freq, uint8_t *d_u_isref, uint64_t *d_n_unique)
{
    if (!ctx) return GB2_ERR_ARG;
    GB2_REQUIRE(ctx, n >= 0 && n < ((int64_t)1 << 31), "gb2_tally_haplotypes: row count must be below 2^31 per call");
    GB2_REQUIRE(ctx, d_n_unique != nullptr, "gb2_tally_haplotypes: null counter");
    GB2_CUDA(ctx, cudaSetDevice(ctx->device));
    GB2_CUDA(ctx, cudaMemsetAsync(d_n_unique, 0, sizeof(uint64_t), ctx->stream));
    if (n == 0) return GB2_OK;
    GB2_REQUIRE(ctx, d_pos && d_packed && d_u_pos && d_u_packed && d_u_freq && d_u_isref, "gb2_tally_haplotypes: null buffer");
    const int ni = (int)n;
    size_t cub_sort = 0, cub_scan = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cub_sort, (const unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                    (const unsigned long long *)nullptr, (unsigned long long *)nullptr, ni, 0, 64, ctx->stream);
    cub::DeviceScan::InclusiveSum(nullptr, cub_scan, (const uint32_t *)nullptr, (unsigned long long *)nullptr, ni, ctx->stream);
    auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t cub_bytes = std::max(cub_sort, cub_scan);
    const size_t need = align(cub_bytes) + 3 * align((size_t)n * 8) + align((size_t)n * 4);
    int rc = gb2_scratch_reserve(ctx, need);
    if (rc != GB2_OK) return rc;
    char *base = (char *)ctx->scratch;
    void *d_tmp = base; base += align(cub_bytes);
    unsigned long long *a = (unsigned long long *)base; base += align((size_t)n * 8);
    unsigned long long *b = (unsigned long long *)base; base += align((size_t)n * 8);
    unsigned long long *scan = (unsigned long long *)base; base += align((size_t)n * 8);
    uint32_t *flags = (uint32_t *)base;
    unsigned long long *pos = (unsigned long long *)d_pos, *pk = (unsigned long long *)d_packed;
    // LSD: stable sort by k-mer, then by position  =>  ordered by (position, k-mer)
    GB2_CUDA(ctx, cub::DeviceRadixSort::SortPairs(d_tmp, cub_sort, pk, a, pos, b, ni, 0, 64, ctx->stream));
    GB2_CUDA(ctx, cub::DeviceRadixSort::SortPairs(d_tmp, cub_sort, b, pos, a, pk, ni, 0, 64, ctx->stream));
    ctx->launches += 2;
    const int threads = 256;
    const unsigned blocks = (unsigned)gb2_div_up(n, threads);
    gb2_tally_heads_kernel<<<blocks, threads, 0, ctx->stream>>>(pos, pk, n, flags);
    GB2_LAUNCH_CHECK(ctx);
    GB2_CUDA(ctx, cub::DeviceScan::InclusiveSum(d_tmp, cub_scan, flags, scan, ni, ctx->stream));
    ctx->launches += 1;
    gb2_tally_emit_kernel<<<blocks, threads, 0, ctx->stream>>>(pos, pk, n, flags, scan, (const unsigned long long *)d_ref_packed,
                                                              pos_base, n_ref, (unsigned long long *)d_u_pos,
                                                              (unsigned long long *)d_u_packed, a, d_u_isref,
                                                              (unsigned long long *)d_n_unique);
    GB2_LAUNCH_CHECK(ctx);
    gb2_tally_freq_kernel<<<blocks, threads, 0, ctx->stream>>>(a, (const unsigned long long *)d_n_unique, n, d_u_freq);
    GB2_LAUNCH_CHECK(ctx);
    return GB2_OK;
}
