// pval.cu -- K3 (batched Staden score-distribution DP) and K4 (score -> p-value table).
//
// Both kernels reproduce the reference's fp64 results bit for bit, which fixes the ORDER of every
// addition; parallelism therefore comes from independent output bins, never from splitting a sum.
//   K3  motif_processing.pyx:588-602   per target bin: (((0 + tA) + tC) + tG) + tT, t = source*bg
//                                      (product rounded, then the add rounded -- no FMA)
//   K4  score_sequences.py:390-391     p[s] = seqsum(pv[s:]) / seqsum(pv), numba's sequential
//                                      ascending sum (numba/np/arraymath.py:163-170)
#include <algorithm>
#include <numeric>

#include "internal.cuh"

// ---------------------------------------------------------------------------------------------
// K4
// ---------------------------------------------------------------------------------------------
// One thread per start score s; each walks pv[s..span) in ascending order.  Lanes of a warp read
// consecutive addresses at every step (coalesced, L1-resident), the adds form one dependent chain
// per thread as the contract requires.  Work is O(span^2/2) adds, latency-bound by the longest chain.
// K4 for MANY motifs in two launches (a JASPAR-sized collection used to cost two launches, a D2H copy and a stream
// synchronisation PER MOTIF).  The reachable slices of all p-value matrices lie back to back (motif m = elements
// [off[m], off[m+1])); a CTA takes 128 consecutive start scores of one motif (tile -> motif by binary search over the
// per-motif tile prefix).  Same dependent add chain per start score as above: bit-identical tables.
struct PtabBatch {
    const int64_t *off;       // [n+1] element offsets of the inputs (and of the cumulative sums)
    const int64_t *tile_off;  // [n+1] tiles before motif m
    const int64_t *out_off;   // [n]   first element of motif m's p-value table in the output array
    int n;
};

__device__ __forceinline__ int ptab_find_motif(const PtabBatch &b, int64_t tile)
{
    int lo = 0, hi = b.n;  // tile_off[lo] <= tile < tile_off[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (b.tile_off[mid] <= tile) lo = mid;
        else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(128) gb2_ctab_batched_kernel(const PtabBatch b, const double *__restrict__ pv_all,
                                                               double *__restrict__ ctab_all)
{
    __shared__ int s_m;
    if (threadIdx.x == 0) s_m = ptab_find_motif(b, blockIdx.x);
    __syncthreads();
    const int m = s_m;
    const int64_t base = b.off[m], span = b.off[m + 1] - base;
    const int64_t s = ((int64_t)blockIdx.x - b.tile_off[m]) * 128 + threadIdx.x;
    if (s >= span) return;
    const double *pv = pv_all + base;
    double c = 0.0;
    int64_t k = s;
#pragma unroll 1
    for (; k + 8 <= span; k += 8) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = pv[k + u];
#pragma unroll
        for (int u = 0; u < 8; ++u) c = __dadd_rn(c, v[u]);
    }
    for (; k < span; ++k) c = __dadd_rn(c, pv[k]);
    ctab_all[base + s] = c;
}

__global__ void __launch_bounds__(128) gb2_ptab_div_batched_kernel(const PtabBatch b, const double *__restrict__ ctab_all,
                                                                   double *__restrict__ ptab_all, double *__restrict__ totals)
{
    __shared__ int s_m;
    if (threadIdx.x == 0) s_m = ptab_find_motif(b, blockIdx.x);
    __syncthreads();
    const int m = s_m;
    const int64_t base = b.off[m], span = b.off[m + 1] - base;
    const int64_t s = ((int64_t)blockIdx.x - b.tile_off[m]) * 128 + threadIdx.x;
    if (s >= span) return;
    const double tot = ctab_all[base];  // seqsum over the whole matrix (zeros outside [lo,hi])
    ptab_all[b.out_off[m] + s] = __ddiv_rn(ctab_all[base + s], tot);
    if (s == 0) totals[m] = tot;
}

int gb2_launch_ptable_batched(gb2_ctx *ctx, int n, const int64_t *h_off, const int64_t *h_out_off, const double *d_pm,
                              double *d_ctab, double *d_ptab, double *d_totals)
{
    std::vector<int64_t> host((size_t)3 * (n + 1));
    int64_t tiles = 0;
    for (int m = 0; m <= n; ++m) {
        host[(size_t)m] = h_off[m];
        host[(size_t)(n + 1 + m)] = tiles;
        host[(size_t)(2 * (n + 1) + m)] = m < n ? h_out_off[m] : 0;
        if (m < n) tiles += gb2_div_up(h_off[m + 1] - h_off[m], 128);
    }
    GB2_REQUIRE(ctx, tiles > 0 && tiles < ((int64_t)1 << 31), "gb2_motif_create: too many score bins for one launch");
    int64_t *d_idx = nullptr;  // tiny, lives until the kernels are done (stream-ordered free)
    GB2_CUDA(ctx, cudaMallocAsync((void **)&d_idx, host.size() * sizeof(int64_t), ctx->stream));
    GB2_CUDA(ctx, cudaMemcpyAsync(d_idx, host.data(), host.size() * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
    PtabBatch b;
    b.off = d_idx;
    b.tile_off = d_idx + (n + 1);
    b.out_off = d_idx + 2 * (n + 1);
    b.n = n;
    gb2_ctab_batched_kernel<<<(unsigned)tiles, 128, 0, ctx->stream>>>(b, d_pm, d_ctab);
    GB2_LAUNCH_CHECK(ctx);
    gb2_ptab_div_batched_kernel<<<(unsigned)tiles, 128, 0, ctx->stream>>>(b, d_ctab, d_ptab, d_totals);
    GB2_LAUNCH_CHECK(ctx);
    GB2_CUDA(ctx, cudaFreeAsync(d_idx, ctx->stream));
    return GB2_OK;
}

// ---------------------------------------------------------------------------------------------
// K3
// ---------------------------------------------------------------------------------------------
struct DpDesc {
    int32_t w;
    int32_t pad;
    int64_t sm_off;    // into the concatenated int32 score matrices ([4][w] each)
    int64_t work_off;  // into the work buffer (2 * max_span doubles per motif)
    int64_t out_off;   // into the concatenated output (1000*w+1 doubles per motif)
    int64_t max_span;  // sum_j (maxcol_j - mincol_j) + 1
    int64_t lo;        // sum_j mincol_j
    double bg[4];
};

// One CTA per motif.  Rows are stored relative to the running minimum score, so row `pos` occupies
// [0, span_pos) and the recurrence is  cur[r] = sum_n prev[r - (sm[n][pos] - mincol_pos)] * bg[n].
// The two rows ping-pong in global memory (L2-resident: <= 480 KB per motif); __syncthreads() orders
// the row hand-over inside the CTA.
__global__ void __launch_bounds__(1024, 1) gb2_dp_kernel(const DpDesc *__restrict__ descs,
                                                         const int32_t *__restrict__ sm_all,
                                                         double *work, double *out)
{
    __shared__ int s_delta[GB2_MAX_WIDTH][4];
    __shared__ int s_span[GB2_MAX_WIDTH];
    const DpDesc d = descs[blockIdx.x];
    const int w = d.w;
    const int32_t *sm = sm_all + d.sm_off;
    const int tid = threadIdx.x, nt = blockDim.x;
    if (tid == 0) {
        int span = 1;
        for (int j = 0; j < w; ++j) {
            int mn = sm[j], mx = sm[j];
            for (int n = 1; n < 4; ++n) {
                int v = sm[n * w + j];
                mn = min(mn, v);
                mx = max(mx, v);
            }
            for (int n = 0; n < 4; ++n) s_delta[j][n] = sm[n * w + j] - mn;
            span += mx - mn;
            s_span[j] = span;  // support width after position j
        }
    }
    // clear the output row (the support is written at the end)
    const int64_t L = (int64_t)GB2_RANGE * w + 1;
    double *o = out + d.out_off;
    for (int64_t k = tid; k < L; k += nt) o[k] = 0.0;
    __syncthreads();

    double *prev = work + d.work_off;
    double *cur = prev + d.max_span;
    const double bg0 = d.bg[0], bg1 = d.bg[1], bg2 = d.bg[2], bg3 = d.bg[3];
    {   // position 0: pv[0, sm[n,0]] += 1*bg[n], n = A,C,G,T
        const int sp = s_span[0];
        const int d0 = s_delta[0][0], d1 = s_delta[0][1], d2 = s_delta[0][2], d3 = s_delta[0][3];
        for (int r = tid; r < sp; r += nt) {
            double acc = 0.0;
            if (r == d0) acc = __dadd_rn(acc, bg0);
            if (r == d1) acc = __dadd_rn(acc, bg1);
            if (r == d2) acc = __dadd_rn(acc, bg2);
            if (r == d3) acc = __dadd_rn(acc, bg3);
            prev[r] = acc;
        }
    }
    __syncthreads();
    for (int pos = 1; pos < w; ++pos) {
        const int sp_prev = s_span[pos - 1], sp = s_span[pos];
        const int d0 = s_delta[pos][0], d1 = s_delta[pos][1], d2 = s_delta[pos][2], d3 = s_delta[pos][3];
        for (int r = tid; r < sp; r += nt) {
            const int i0 = r - d0, i1 = r - d1, i2 = r - d2, i3 = r - d3;
            const double v0 = (i0 >= 0 && i0 < sp_prev) ? prev[i0] : 0.0;
            const double v1 = (i1 >= 0 && i1 < sp_prev) ? prev[i1] : 0.0;
            const double v2 = (i2 >= 0 && i2 < sp_prev) ? prev[i2] : 0.0;
            const double v3 = (i3 >= 0 && i3 < sp_prev) ? prev[i3] : 0.0;
            double acc = 0.0;
            if (v0 > 0.0) acc = __dadd_rn(acc, __dmul_rn(v0, bg0));
            if (v1 > 0.0) acc = __dadd_rn(acc, __dmul_rn(v1, bg1));
            if (v2 > 0.0) acc = __dadd_rn(acc, __dmul_rn(v2, bg2));
            if (v3 > 0.0) acc = __dadd_rn(acc, __dmul_rn(v3, bg3));
            cur[r] = acc;
        }
        __syncthreads();
        double *t = prev; prev = cur; cur = t;
    }
    const int spf = s_span[w - 1];
    for (int r = tid; r < spf; r += nt) o[d.lo + r] = prev[r];
}

extern "C" int gb2_pval_dp_batched(gb2_ctx *ctx, int n_motifs, const int32_t *h_widths, const int64_t *h_sm,
                                   const double *h_bgs, double *h_out)
{
    if (!ctx) return GB2_ERR_ARG;
    GB2_REQUIRE(ctx, n_motifs >= 0 && (n_motifs == 0 || (h_widths && h_sm && h_bgs && h_out)), "gb2_pval_dp_batched: null argument");
    if (n_motifs == 0) return GB2_OK;
    GB2_CUDA(ctx, cudaSetDevice(ctx->device));
    std::vector<DpDesc> descs((size_t)n_motifs);
    std::vector<int32_t> sm32;
    int64_t sm_off = 0, work_off = 0, out_off = 0;
    for (int m = 0; m < n_motifs; ++m) {
        const int w = h_widths[m];
        GB2_REQUIRE(ctx, w >= 1 && w <= GB2_MAX_WIDTH, "gb2_pval_dp_batched: motif %d width %d outside [1,%d]", m, w, GB2_MAX_WIDTH);
        DpDesc &d = descs[(size_t)m];
        d.w = w; d.pad = 0;
        d.sm_off = sm_off; d.work_off = work_off; d.out_off = out_off;
        int64_t lo = 0, hi = 0;
        for (int j = 0; j < w; ++j) {
            int64_t mn = h_sm[sm_off + j], mx = mn;
            for (int n = 0; n < 4; ++n) {
                int64_t v = h_sm[sm_off + (int64_t)n * w + j];
                if (v < 0 || v > 60000) {
                    GB2_SET_ERR(ctx, "gb2_pval_dp_batched: motif %d scaled score %lld out of range", m, (long long)v);
                    return GB2_ERR_MOTIF;
                }
                mn = std::min(mn, v); mx = std::max(mx, v);
            }
            lo += mn; hi += mx;
        }
        if (hi > (int64_t)GB2_RANGE * w) {
            GB2_SET_ERR(ctx, "gb2_pval_dp_batched: motif %d max score %lld exceeds RANGE*w", m, (long long)hi);
            return GB2_ERR_MOTIF;
        }
        d.lo = lo;
        d.max_span = hi - lo + 1;
        for (int n = 0; n < 4; ++n) {
            d.bg[n] = h_bgs[(size_t)m * 4 + n];
            if (!(d.bg[n] > 0.0)) {
                GB2_SET_ERR(ctx, "gb2_pval_dp_batched: motif %d background[%d] must be > 0", m, n);
                return GB2_ERR_MOTIF;
            }
        }
        for (int64_t k = 0; k < 4 * (int64_t)w; ++k) sm32.push_back((int32_t)h_sm[sm_off + k]);
        sm_off += 4 * (int64_t)w;
        work_off += 2 * d.max_span;
        out_off += (int64_t)GB2_RANGE * w + 1;
    }
    // longest motifs first: one CTA per motif, list-scheduled over the SMs
    std::vector<int> order((size_t)n_motifs);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        return descs[(size_t)a].max_span * descs[(size_t)a].w > descs[(size_t)b].max_span * descs[(size_t)b].w;
    });
    std::vector<DpDesc> sorted((size_t)n_motifs);
    for (int i = 0; i < n_motifs; ++i) sorted[(size_t)i] = descs[(size_t)order[(size_t)i]];

    const size_t bytes_desc = sorted.size() * sizeof(DpDesc);
    const size_t bytes_sm = sm32.size() * sizeof(int32_t);
    const size_t bytes_work = (size_t)work_off * sizeof(double);
    const size_t bytes_out = (size_t)out_off * sizeof(double);
    auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t total = align(bytes_desc) + align(bytes_sm) + align(bytes_work) + align(bytes_out);
    int rc = gb2_scratch_reserve(ctx, total);
    if (rc != GB2_OK) return rc;
    char *base = (char *)ctx->scratch;
    DpDesc *d_desc = (DpDesc *)base;
    int32_t *d_sm = (int32_t *)(base + align(bytes_desc));
    double *d_work = (double *)(base + align(bytes_desc) + align(bytes_sm));
    double *d_out = (double *)(base + align(bytes_desc) + align(bytes_sm) + align(bytes_work));
    GB2_CUDA(ctx, cudaMemcpyAsync(d_desc, sorted.data(), bytes_desc, cudaMemcpyHostToDevice, ctx->stream));
    GB2_CUDA(ctx, cudaMemcpyAsync(d_sm, sm32.data(), bytes_sm, cudaMemcpyHostToDevice, ctx->stream));
    gb2_dp_kernel<<<n_motifs, 1024, 0, ctx->stream>>>(d_desc, d_sm, d_work, d_out);
    GB2_LAUNCH_CHECK(ctx);
    GB2_CUDA(ctx, cudaMemcpyAsync(h_out, d_out, bytes_out, cudaMemcpyDeviceToHost, ctx->stream));
    GB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GB2_OK;
}
