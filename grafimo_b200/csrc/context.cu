// context.cu -- context / motif lifetime of the grafimo_b200 C ABI (include/grafimo_b200.h).
#include <stdlib.h>

#include <math.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <memory>
#include <new>
#include <thread>

#include "internal.cuh"

extern "C" int gb2_abi_version(void) { return GB2_ABI_VERSION; }

extern "C" const char *gb2_error_string(int code)
{
    switch (code) {
    case GB2_OK: return "ok";
    case GB2_ERR_ARG: return "bad argument";
    case GB2_ERR_CUDA: return "CUDA runtime failure (no CPU fallback exists)";
    case GB2_ERR_NOMEM: return "out of memory";
    case GB2_ERR_CAPACITY: return "hit buffer too small";
    case GB2_ERR_MOTIF: return "motif not usable";
    case GB2_ERR_STATE: return "bad state";
    default: return "unknown error";
    }
}

extern "C" int gb2_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" int gb2_ctx_create(int device, void *stream, gb2_ctx **out)
{
    if (!out) return GB2_ERR_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return GB2_ERR_CUDA;
    }
    if (device < 0 || device >= n) return GB2_ERR_ARG;
    gb2_ctx *ctx = new (std::nothrow) gb2_ctx();
    if (!ctx) return GB2_ERR_NOMEM;
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return GB2_ERR_CUDA; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return GB2_ERR_CUDA; }
    ctx->sm_count = prop.multiProcessorCount;
    ctx->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    if (stream) {
        ctx->stream = (cudaStream_t)stream;
    } else {
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return GB2_ERR_CUDA; }
        ctx->own_stream = true;
    }
    if (cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return GB2_ERR_CUDA; }
    if (cudaMallocHost((void **)&ctx->h_mail, 64 * sizeof(uint64_t)) != cudaSuccess) { delete ctx; return GB2_ERR_NOMEM; }
    *out = ctx;
    return GB2_OK;
}

extern "C" int gb2_ctx_destroy(gb2_ctx *ctx)
{
    if (!ctx) return GB2_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    gb2_comm_release(ctx);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->pool) cudaFree(ctx->pool);
    for (auto &d : ctx->desc)
        if (d.d_ptr) cudaFree(d.d_ptr);
    if (ctx->h_mail) cudaFreeHost(ctx->h_mail);
    if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return GB2_OK;
}

extern "C" int gb2_ctx_set_stream(gb2_ctx *ctx, void *stream)
{
    if (!ctx) return GB2_ERR_ARG;
    if (ctx->own_stream) {
        cudaStreamSynchronize(ctx->stream);
        cudaStreamDestroy(ctx->stream);
        ctx->own_stream = false;
    }
    ctx->stream = (cudaStream_t)stream;
    return GB2_OK;
}

extern "C" int gb2_ctx_sync(gb2_ctx *ctx)
{
    if (!ctx) return GB2_ERR_ARG;
    GB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GB2_OK;
}

extern "C" const char *gb2_ctx_last_error(const gb2_ctx *ctx) { return ctx ? ctx->err : "null context"; }
extern "C" int64_t gb2_ctx_launch_count(const gb2_ctx *ctx) { return ctx ? ctx->launches : 0; }
extern "C" int gb2_ctx_sm_count(const gb2_ctx *ctx) { return ctx ? ctx->sm_count : 0; }

int gb2_scratch_reserve(gb2_ctx *ctx, size_t bytes)
{
    if (bytes <= ctx->scratch_bytes) return GB2_OK;
    GB2_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ctx->scratch) {
        GB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        GB2_CUDA(ctx, cudaFree(ctx->scratch));
        ctx->scratch = nullptr;
        ctx->scratch_bytes = 0;
    }
    size_t want = std::max(bytes + bytes / 4, (size_t)1 << 20);
    GB2_CUDA(ctx, cudaMalloc(&ctx->scratch, want));
    ctx->scratch_bytes = want;
    return GB2_OK;
}

int gb2_pool_reserve(gb2_ctx *ctx, size_t bytes, char **out)
{
    *out = nullptr;
    if (bytes > ctx->pool_bytes) {
        GB2_CUDA(ctx, cudaSetDevice(ctx->device));
        if (ctx->pool) {
            GB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            GB2_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
            GB2_CUDA(ctx, cudaFree(ctx->pool));
            ctx->pool = nullptr;
            ctx->pool_bytes = 0;
        }
        GB2_CUDA(ctx, cudaMalloc(&ctx->pool, bytes));
        ctx->pool_bytes = bytes;
    }
    *out = (char *)ctx->pool;
    return GB2_OK;
}

extern "C" int gb2_scan_last_transfer(const gb2_ctx *ctx, uint64_t *h2d_bytes, uint64_t *d2h_bytes, uint64_t *chunks_as_given,
                                      uint64_t *chunks_host_packed)
{
    if (!ctx) return GB2_ERR_ARG;
    if (h2d_bytes) *h2d_bytes = ctx->last_h2d_bytes;
    if (d2h_bytes) *d2h_bytes = ctx->last_d2h_bytes;
    if (chunks_as_given) *chunks_as_given = ctx->last_chunks_given;
    if (chunks_host_packed) *chunks_host_packed = ctx->last_chunks_packed;
    return GB2_OK;
}

int gb2_pinned_reserve(gb2_ctx *ctx, size_t bytes, char **out)
{
    *out = nullptr;
    if (bytes > ctx->h_pin_bytes) {
        GB2_CUDA(ctx, cudaSetDevice(ctx->device));
        if (ctx->h_pin) {
            GB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            GB2_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
            GB2_CUDA(ctx, cudaFreeHost(ctx->h_pin));
            ctx->h_pin = nullptr;
            ctx->h_pin_bytes = 0;
        }
        GB2_CUDA(ctx, cudaHostAlloc(&ctx->h_pin, bytes, cudaHostAllocDefault));
        ctx->h_pin_bytes = bytes;
    }
    *out = (char *)ctx->h_pin;
    return GB2_OK;
}

extern "C" int gb2_host_alloc(uint64_t bytes, void **out)
{
    if (!out) return GB2_ERR_ARG;
    *out = nullptr;
    if (cudaHostAlloc(out, (size_t)std::max<uint64_t>(bytes, 1), cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        return GB2_ERR_NOMEM;
    }
    return GB2_OK;
}

extern "C" int gb2_host_free(void *ptr)
{
    if (ptr && cudaFreeHost(ptr) != cudaSuccess) {
        cudaGetLastError();
        return GB2_ERR_CUDA;
    }
    return GB2_OK;
}

// ---------------------------------------------------------------------------------------------
// motif
// ---------------------------------------------------------------------------------------------
// Device memory of one or many motifs (p-value tables, chunk LUTs, hit bitmaps) is ONE allocation shared by the motifs
// created together; it is freed when the last of them is destroyed.
struct gb2_motif_block {
    void *d_ptr = nullptr;
    int refs = 0;
    int device = 0;
};

extern "C" int gb2_motif_destroy(gb2_motif *m)
{
    if (!m) return GB2_OK;
    if (m->block && --m->block->refs == 0) {
        cudaSetDevice(m->block->device);
        if (m->block->d_ptr) cudaFree(m->block->d_ptr);
        delete m->block;
    }
    delete m;
    return GB2_OK;
}

namespace {
inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// Host-side facts of one motif: reachable score range and the chunk lookup tables.
struct MotifPlan {
    int w = 0, cb = 4, n_chunks = 0;
    int64_t lo = 0, hi = 0, span = 0;
    std::vector<uint32_t> lut;  // [n_chunks][4^cb]: (rc_rel << 16) | fwd_rel
};

// Expected shared-memory wavefronts of one lookup when the table is replicated R times (32 lanes, random entries):
// R = 32 -> every lane has its own bank; below that, lanes sharing a replica collide (measured / simulated).
double lookup_wavefronts(int R)
{
    switch (R) {
    case 32: return 1.0;
    case 16: return 2.0;
    case 8: return 3.0;
    case 4: return 5.0;
    case 2: return 8.0;
    default: return 12.0;
    }
}

// Chunk size (bases per lookup) and replication for a motif: tables of 4^cb entries per chunk, ceil(w / cb) lookups per
// k-mer; the smaller tables of cb = 3 replicate 32x where the 4-base tables no longer fit beside the histogram, which
// makes every lookup one conflict-free wavefront (long motifs, large score spans).  Picks the smaller expected number of
// shared-memory wavefronts per k-mer.  -> false when not even one copy of the tables fits with the histogram in shared
// memory (then hist_global for wide motifs).
bool plan_smem(int w, int64_t span, int64_t budget, bool allow_global_hist, int &cb, int &R, int &hist_global, int64_t &smem)
{
    const int64_t hist_bytes = (span + 1) * 4;
    double best = 1e30;
    bool found = false;
    for (int c : {4, 3}) {
        if (c == 3 && w < 19) continue;  // never wins below that width (and keeps the number of kernel variants down)
        if (c == 4 && w > GB2_NARROW_WIDTH) continue;  // the wide kernel exists for 3-base chunks only (they always win there)
        const int nch = (w + c - 1) / c;
        const int64_t copy = (int64_t)nch * (c == 4 ? 1024 : 256);
        for (int hg = 0; hg <= (allow_global_hist ? 1 : 0); ++hg) {
            const int64_t hb = hg ? 0 : hist_bytes;
            if (copy + hb > budget) continue;
            int r = 32;
            while (r > 1 && copy * r + hb > budget) r >>= 1;
            if (c == 3 && r < 8) continue;  // the 3-base kernels exist for R = 32, 16, 8 (wide: then the global histogram)
            // global-memory histogram updates are ~35x slower than shared-memory ones: only when nothing else fits
            const double cost = nch * lookup_wavefronts(r) + (hg ? 300.0 : 0.0);
            if (cost < best - 1e-9) { best = cost; cb = c; R = r; hist_global = hg; smem = copy * r + hb; found = true; }
        }
    }
    return found;
}

// K4 without the quadratic chain when floating-point addition cannot round on this data: every value of pv[0..span) is a
// non-negative integer multiple of 2^L (L = the lowest set bit over all values) and the sum of all of them is at most
// 2^53 such units.  Then every partial sum of ANY order is exactly representable, each fp64 addition of the reference's
// sequential sums (score_sequences.py:390-391) is exact, and the integer suffix sums below are those sums, bit for bit.
// -> false when the condition does not hold (the caller uses the order-preserving kernels instead).
bool exact_ptable(const double *pv, int64_t span, double *ptab, double *total_out)
{
    int min_lsb = INT32_MAX;
    for (int64_t k = 0; k < span; ++k) {
        const double v = pv[k];
        if (v == 0.0) continue;
        uint64_t bits;
        memcpy(&bits, &v, sizeof(bits));
        const int e = (int)((bits >> 52) & 0x7FFu);
        if ((bits >> 63) || e == 0 || e == 0x7FF) return false;  // negative, subnormal, inf/nan: not here
        const uint64_t mant = (bits & ((1ull << 52) - 1ull)) | (1ull << 52);
        min_lsb = std::min(min_lsb, e - 1075 + __builtin_ctzll(mant));  // v = mant * 2^(e-1075)
    }
    if (min_lsb == INT32_MAX) return false;
    if (min_lsb < -1000 || min_lsb > 1000) return false;  // the two scale factors below must be normal numbers
    const double limit = 9007199254740992.0;                // 2^53
    const double to_units = ldexp(1.0, -min_lsb), from_units = ldexp(1.0, min_lsb);  // powers of two: exact scaling
    uint64_t acc = 0;
    for (int64_t k = span - 1; k >= 0; --k) {
        const double u = pv[k] * to_units;  // an integer by construction
        if (!(u <= limit)) return false;
        acc += (uint64_t)u;
        if (acc > (1ull << 53)) return false;
        ptab[k] = (double)acc;  // exact: acc <= 2^53
    }
    const double total = (double)acc * from_units;
    for (int64_t k = 0; k < span; ++k) ptab[k] = (ptab[k] * from_units) / total;  // IEEE division, as K4's div kernel
    *total_out = total;
    return true;
}

// the per-motif host work of a collection (plans, checks, exact p-tables) is independent per motif: a few worker threads
template <typename F>
void parallel_for(int n, F &&fn)
{
    const int nt = std::max(1, std::min<int>({n / 16, (int)std::thread::hardware_concurrency(), 16}));
    if (nt <= 1) {
        for (int i = 0; i < n; ++i) fn(i);
        return;
    }
    std::atomic<int> next(0);
    std::vector<std::thread> pool;
    for (int t = 0; t < nt; ++t)
        pool.emplace_back([&]() {
            for (int i = next.fetch_add(1); i < n; i = next.fetch_add(1)) fn(i);
        });
    for (auto &th : pool) th.join();
}

int motif_plan(gb2_ctx *ctx, const int64_t *sm, int w, int cb, MotifPlan &pl, const char *who)
{
    int64_t mincol[GB2_MAX_WIDTH], lo = 0, hi = 0;
    for (int j = 0; j < w; ++j) {
        int64_t mn = sm[j], mx = sm[j];
        for (int n = 0; n < 4; ++n) {
            int64_t v = sm[(int64_t)n * w + j];
            if (v < 0 || v > 60000) {
                GB2_SET_ERR(ctx, "%s: scaled score %lld out of range", who, (long long)v);
                return GB2_ERR_MOTIF;
            }
            mn = std::min(mn, v);
            mx = std::max(mx, v);
        }
        mincol[j] = mn;
        lo += mn;
        hi += mx;
    }
    pl.w = w; pl.cb = cb; pl.n_chunks = (w + cb - 1) / cb;
    pl.lo = lo; pl.hi = hi; pl.span = hi - lo + 1;
    // chunk LUT: entry = (rc_rel << 16) | fwd_rel, both relative to the column minima so that the accumulated fields
    // are directly the histogram bins (score - lo).  Index bits beyond the motif's last base are ignored.
    const int entries = 1 << (2 * cb);
    pl.lut.assign((size_t)pl.n_chunks * entries, 0u);
    for (int c = 0; c < pl.n_chunks; ++c) {
        for (int idx = 0; idx < entries; ++idx) {
            uint32_t fwd = 0, rc = 0;
            for (int j = 0; j < cb; ++j) {
                int p = cb * c + j;
                if (p >= w) break;
                int b = (idx >> (2 * j)) & 3;
                fwd += (uint32_t)(sm[(int64_t)b * w + p] - mincol[p]);
                int q = w - 1 - p;  // base at position p of x sits at position q of the reverse complement
                rc += (uint32_t)(sm[(int64_t)(3 - b) * w + q] - mincol[q]);
            }
            pl.lut[(size_t)c * entries + idx] = (rc << 16) | fwd;
        }
    }
    return GB2_OK;
}
}  // namespace

// K4 over many motifs at once (pval.cu)
int gb2_launch_ptable_batched(gb2_ctx *ctx, int n, const int64_t *h_in_off, const int64_t *h_out_off, const double *d_pm,
                              double *d_ctab, double *d_ptab, double *d_totals);

extern "C" int gb2_motif_create_batched(gb2_ctx *ctx, int n_motifs, const int32_t *h_widths, const int64_t *const *h_score_mats,
                                        const double *const *h_pval_mats, const int64_t *h_min_vals, const int64_t *h_scales,
                                        const double *h_offsets, gb2_motif **out)
{
    if (!ctx || !out) return GB2_ERR_ARG;
    for (int i = 0; i < n_motifs; ++i) out[i] = nullptr;
    GB2_REQUIRE(ctx, n_motifs >= 0, "gb2_motif_create_batched: negative motif count");
    if (n_motifs == 0) return GB2_OK;
    GB2_REQUIRE(ctx, h_widths && h_score_mats && h_pval_mats && h_min_vals && h_scales && h_offsets,
                "gb2_motif_create_batched: null argument");
    GB2_CUDA(ctx, cudaSetDevice(ctx->device));
    const bool timing = getenv("GB2_MOTIF_TIMING") != nullptr;  // phase times of this call on stderr
    const bool force_chain = getenv("GB2_K4_FORCE_CHAIN") != nullptr;  // tests: every motif through the device kernels
    const auto t_begin = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (timing) fprintf(stderr, "gb2_motif_create_batched[%d]: %-28s %.2f ms since entry\n", n_motifs, what,
                            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count());
    };
    const int64_t budget = (int64_t)ctx->max_smem_optin - 1024;
    std::vector<MotifPlan> plans((size_t)n_motifs);
    std::vector<gb2_motif *> ms((size_t)n_motifs, nullptr);
    std::vector<int64_t> sp_off((size_t)n_motifs + 1, 0);
    std::vector<size_t> lut_off((size_t)n_motifs + 1, 0), bm_off((size_t)n_motifs + 1, 0);
    int rc = GB2_OK;
    auto fail = [&](int code) {
        for (gb2_motif *m : ms) delete m;
        return code;
    };
    for (int i = 0; i < n_motifs; ++i) {
        const int w = h_widths[i];
        if (w < 1 || w > GB2_MAX_WIDTH) {
            GB2_SET_ERR(ctx, "gb2_motif_create: width %d outside [1,%d] (motif %d)", w, GB2_MAX_WIDTH, i);
            return fail(GB2_ERR_ARG);
        }
        if (!h_score_mats[i] || !h_pval_mats[i]) {
            GB2_SET_ERR(ctx, "gb2_motif_create: null matrix (motif %d)", i);
            return fail(GB2_ERR_ARG);
        }
        if (h_scales[i] <= 0) {
            GB2_SET_ERR(ctx, "gb2_motif_create: motif is not scaled (scale=%lld, motif %d)", (long long)h_scales[i], i);
            return fail(GB2_ERR_MOTIF);
        }
        const int64_t *sm = h_score_mats[i];
        MotifPlan probe;
        if ((rc = motif_plan(ctx, sm, w, 4, probe, "gb2_motif_create")) != GB2_OK) return fail(rc);
        const int64_t L = (int64_t)GB2_RANGE * w + 1;
        if (probe.hi >= L) {
            GB2_SET_ERR(ctx, "gb2_motif_create: max score %lld exceeds the p-value matrix (%lld bins)", (long long)probe.hi, (long long)L);
            return fail(GB2_ERR_MOTIF);
        }
        if (probe.span > 65535) {  // the two strands travel as 16-bit fields of one register
            GB2_SET_ERR(ctx, "gb2_motif_create: score span %lld exceeds 65535", (long long)probe.span);
            return fail(GB2_ERR_MOTIF);
        }
        int cb = 4, R = 1, hg = 0;
        int64_t smem = 0;
        if (!plan_smem(w, probe.span, budget, w > GB2_NARROW_WIDTH, cb, R, hg, smem)) {
            GB2_SET_ERR(ctx, "gb2_motif_create: score span %lld does not fit shared memory", (long long)probe.span);
            return fail(GB2_ERR_MOTIF);
        }
        if (cb == 4) plans[(size_t)i] = std::move(probe);
        else if ((rc = motif_plan(ctx, sm, w, cb, plans[(size_t)i], "gb2_motif_create")) != GB2_OK) return fail(rc);
        const MotifPlan &pl = plans[(size_t)i];
        gb2_motif *m = new (std::nothrow) gb2_motif();
        if (!m) return fail(GB2_ERR_NOMEM);
        ms[(size_t)i] = m;
        m->device = ctx->device;
        m->w = w; m->chunk_bases = cb; m->n_chunks = pl.n_chunks;
        m->lo = pl.lo; m->hi = pl.hi; m->span = pl.span;
        m->min_val = h_min_vals[i]; m->scale = h_scales[i]; m->offset = h_offsets[i];
        m->replicas = R; m->hist_global = hg; m->smem_bytes = smem;
        sp_off[(size_t)i + 1] = sp_off[(size_t)i] + pl.span;
        lut_off[(size_t)i + 1] = lut_off[(size_t)i] + align256(pl.lut.size() * sizeof(uint32_t));
        bm_off[(size_t)i + 1] = bm_off[(size_t)i] + align256((size_t)gb2_div_up(pl.span + 1, 32) * sizeof(uint32_t));
    }
    {   // mass outside the reachable range means the matrix and the p-value matrix do not belong together
        std::vector<int64_t> bad_at((size_t)n_motifs, -1);
        parallel_for(n_motifs, [&](int i) {
            const MotifPlan &pl = plans[(size_t)i];
            const double *pm = h_pval_mats[i];
            const int64_t L = (int64_t)GB2_RANGE * pl.w + 1;
            for (int64_t k = 0; k < pl.lo; ++k)
                if (pm[k] != 0.0) { bad_at[(size_t)i] = k; return; }
            for (int64_t k = pl.hi + 1; k < L; ++k)
                if (pm[k] != 0.0) { bad_at[(size_t)i] = k; return; }
        });
        for (int i = 0; i < n_motifs; ++i)
            if (bad_at[(size_t)i] >= 0) {
                GB2_SET_ERR(ctx, "gb2_motif_create: p-value matrix has mass at unreachable score %lld", (long long)bad_at[(size_t)i]);
                return fail(GB2_ERR_MOTIF);
            }
    }
    lap("plans + checks (host)");
    // ---- one allocation: [p-tables of all motifs][LUTs][bitmaps]
    const int64_t total_span = sp_off[(size_t)n_motifs];
    const size_t b_ptab = align256((size_t)total_span * sizeof(double));
    const size_t b_lut = lut_off[(size_t)n_motifs], b_bm = bm_off[(size_t)n_motifs];
    gb2_motif_block *blk = new (std::nothrow) gb2_motif_block();
    if (!blk) return fail(GB2_ERR_NOMEM);
    blk->device = ctx->device;
    if (cudaMalloc(&blk->d_ptr, b_ptab + b_lut + b_bm) != cudaSuccess) {
        GB2_SET_ERR(ctx, "gb2_motif_create: cudaMalloc failed: %s", cudaGetErrorString(cudaGetLastError()));
        delete blk;
        return fail(GB2_ERR_NOMEM);
    }
    auto drop = [&](int code) {
        cudaFree(blk->d_ptr);
        delete blk;
        return fail(code);
    };
    char *base = (char *)blk->d_ptr;
    // K4, p[s] = seqsum(pv[s:]) / seqsum(pv) in the reference's summation order (score_sequences.py:390-391).
    //  * EXACT motifs -- every partial sum is an integer multiple of one power of two and fits 53 bits (uniform background,
    //    the reference's default, up to 26 bp: all values are multiples of 4^-w) -- are done here on the host with integer
    //    suffix sums: fp64 addition cannot round on such data, so the order of the additions does not matter and the
    //    O(span) suffix sum IS the reference's O(span^2) result, bit for bit (exact_ptable).
    //  * the others keep the order-preserving device kernels (one dependent add chain per start score), batched.
    std::unique_ptr<uint8_t[]> h_lut(new (std::nothrow) uint8_t[b_lut]);
    std::unique_ptr<double[]> h_ptab_all(new (std::nothrow) double[(size_t)total_span]);
    if (!h_lut || !h_ptab_all) return drop(GB2_ERR_NOMEM);
    std::vector<int> slow;
    std::vector<int64_t> slow_in((size_t)1, 0), slow_out;
    std::vector<uint8_t> is_exact((size_t)n_motifs, 0);
    parallel_for(n_motifs, [&](int i) {
        const MotifPlan &pl = plans[(size_t)i];
        memcpy(h_lut.get() + lut_off[(size_t)i], pl.lut.data(), pl.lut.size() * sizeof(uint32_t));
        gb2_motif *m = ms[(size_t)i];
        m->d_ptab = (double *)base + sp_off[(size_t)i];
        m->d_lut = (uint32_t *)(base + b_ptab + lut_off[(size_t)i]);
        m->d_bitmap = (uint32_t *)(base + b_ptab + b_lut + bm_off[(size_t)i]);
        is_exact[(size_t)i] = !force_chain && exact_ptable(h_pval_mats[i] + pl.lo, pl.span, h_ptab_all.get() + sp_off[(size_t)i], &m->total);
    });
    for (int i = 0; i < n_motifs; ++i) {
        if (is_exact[(size_t)i]) continue;
        slow.push_back(i);
        slow_out.push_back(sp_off[(size_t)i]);
        slow_in.push_back(slow_in.back() + plans[(size_t)i].span);
    }
    lap("LUT staging + exact p-tables (host)");
    cudaError_t e = cudaMemcpyAsync(base + b_ptab, h_lut.get(), b_lut, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && !slow.empty()) {
        const int ns = (int)slow.size();
        const int64_t slow_span = slow_in.back();
        const size_t b_work = align256((size_t)slow_span * sizeof(double));
        const size_t b_tot = align256((size_t)ns * sizeof(double));
        if ((rc = gb2_scratch_reserve(ctx, 2 * b_work + b_tot)) != GB2_OK) return drop(rc);
        double *d_pm = (double *)ctx->scratch;
        double *d_ctab = (double *)((char *)ctx->scratch + b_work);
        double *d_tot = (double *)((char *)ctx->scratch + 2 * b_work);
        std::unique_ptr<double[]> h_pm(new (std::nothrow) double[(size_t)slow_span]);
        std::vector<double> h_tot((size_t)ns);
        if (!h_pm) return drop(GB2_ERR_NOMEM);
        for (int k = 0; k < ns; ++k) {
            const MotifPlan &pl = plans[(size_t)slow[(size_t)k]];
            memcpy(h_pm.get() + slow_in[(size_t)k], h_pval_mats[slow[(size_t)k]] + pl.lo, (size_t)pl.span * sizeof(double));
        }
        e = cudaMemcpyAsync(d_pm, h_pm.get(), (size_t)slow_span * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) {
            rc = gb2_launch_ptable_batched(ctx, ns, slow_in.data(), slow_out.data(), d_pm, d_ctab, (double *)base, d_tot);
            if (rc != GB2_OK) return drop(rc);
            for (int k = 0; k < ns && e == cudaSuccess; ++k)
                e = cudaMemcpyAsync(h_ptab_all.get() + slow_out[(size_t)k], (double *)base + slow_out[(size_t)k],
                                    (size_t)(slow_in[(size_t)k + 1] - slow_in[(size_t)k]) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
            if (e == cudaSuccess) e = cudaMemcpyAsync(h_tot.data(), d_tot, (size_t)ns * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);  // h_pm / h_tot are read by the copies above
            for (int k = 0; k < ns; ++k) ms[(size_t)slow[(size_t)k]]->total = h_tot[(size_t)k];
        }
    }
    // the host-made tables go up in one copy per run of consecutive exact motifs (one copy when all are exact)
    for (int i = 0; i < n_motifs && e == cudaSuccess;) {
        if (std::binary_search(slow.begin(), slow.end(), i)) { ++i; continue; }
        int j = i;
        while (j + 1 < n_motifs && !std::binary_search(slow.begin(), slow.end(), j + 1)) ++j;
        e = cudaMemcpyAsync((double *)base + sp_off[(size_t)i], h_ptab_all.get() + sp_off[(size_t)i],
                            (size_t)(sp_off[(size_t)j + 1] - sp_off[(size_t)i]) * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
        i = j + 1;
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);  // h_lut / h_ptab_all are pageable: done with them here
    if (e != cudaSuccess) {
        GB2_SET_ERR(ctx, "gb2_motif_create: upload / p-table failed: %s", cudaGetErrorString(e));
        return drop(GB2_ERR_CUDA);
    }
    lap("uploads + K4 of the inexact motifs");
    for (int i = 0; i < n_motifs; ++i)
        if (!(ms[(size_t)i]->total > 0.0)) {
            GB2_SET_ERR(ctx, "gb2_motif_create: empty p-value matrix (motif %d)", i);
            return drop(GB2_ERR_MOTIF);
        }
    parallel_for(n_motifs, [&](int i) {
        gb2_motif *m = ms[(size_t)i];
        m->h_ptab.assign(h_ptab_all.get() + sp_off[(size_t)i], h_ptab_all.get() + sp_off[(size_t)i + 1]);
        m->monotone = 1;
        for (int64_t k = 1; k < m->span; ++k)
            if (m->h_ptab[(size_t)k] > m->h_ptab[(size_t)k - 1]) { m->monotone = 0; break; }
        m->ptab_exact_host = is_exact[(size_t)i];
    });
    for (int i = 0; i < n_motifs; ++i) {
        ms[(size_t)i]->block = blk;
        blk->refs++;
        out[i] = ms[(size_t)i];
    }
    lap("monotone checks (host)");
    return GB2_OK;
}

extern "C" int gb2_motif_create(gb2_ctx *ctx, const int64_t *sm, int w, const double *h_pval_mat, int64_t min_val,
                                int64_t scale, double offset, gb2_motif **out)
{
    if (!ctx || !out) return GB2_ERR_ARG;
    *out = nullptr;
    GB2_REQUIRE(ctx, sm && h_pval_mat, "gb2_motif_create: null matrix");
    const int32_t w32 = w;
    return gb2_motif_create_batched(ctx, 1, &w32, &sm, &h_pval_mat, &min_val, &scale, &offset, out);
}

extern "C" int gb2_motif_get_info(const gb2_motif *m, gb2_motif_info *info)
{
    if (!m || !info) return GB2_ERR_ARG;
    info->width = m->w;
    info->n_chunks = m->n_chunks;
    info->chunk_bases = m->chunk_bases;
    info->hist_global = m->hist_global;
    info->lut_replicas = m->replicas;
    info->monotone = m->monotone;
    info->lo = m->lo; info->hi = m->hi; info->span = m->span;
    info->min_val = m->min_val; info->scale = m->scale; info->offset = m->offset;
    info->total = m->total;
    info->smem_bytes = m->smem_bytes;
    return GB2_OK;
}

extern "C" int gb2_motif_get_ptable(gb2_ctx *ctx, const gb2_motif *m, double *h_ptable)
{
    if (!ctx || !m || !h_ptable) return GB2_ERR_ARG;
    memcpy(h_ptable, m->h_ptab.data(), (size_t)m->span * sizeof(double));
    return GB2_OK;
}

extern "C" const double *gb2_motif_ptable_device(const gb2_motif *m) { return m ? m->d_ptab : nullptr; }
