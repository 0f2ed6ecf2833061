// context.cu -- context / motif lifetime of the grafimo_b200 C ABI (include/grafimo_b200.h).
#include <stdlib.h>

#include <algorithm>
#include <new>

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
extern "C" int gb2_motif_destroy(gb2_motif *m)
{
    if (!m) return GB2_OK;
    cudaSetDevice(m->device);
    if (m->d_block) cudaFree(m->d_block);  // d_lut, d_ptab and d_bitmap live in this one allocation
    delete m;
    return GB2_OK;
}

extern "C" int gb2_motif_create(gb2_ctx *ctx, const int64_t *sm, int w, const double *h_pval_mat, int64_t min_val,
                                int64_t scale, double offset, gb2_motif **out)
{
    if (!ctx || !out) return GB2_ERR_ARG;
    *out = nullptr;
    GB2_REQUIRE(ctx, sm && h_pval_mat, "gb2_motif_create: null matrix");
    GB2_REQUIRE(ctx, w >= 1 && w <= GB2_MAX_WIDTH, "gb2_motif_create: width %d outside [1,%d]", w, GB2_MAX_WIDTH);
    if (scale <= 0) {
        GB2_SET_ERR(ctx, "gb2_motif_create: motif is not scaled (scale=%lld)", (long long)scale);
        return GB2_ERR_MOTIF;
    }
    const int64_t L = (int64_t)GB2_RANGE * w + 1;
    int64_t mincol[GB2_MAX_WIDTH], lo = 0, hi = 0;
    for (int j = 0; j < w; ++j) {
        int64_t mn = sm[j], mx = sm[j];
        for (int n = 0; n < 4; ++n) {
            int64_t v = sm[(int64_t)n * w + j];
            if (v < 0 || v > 60000) {
                GB2_SET_ERR(ctx, "gb2_motif_create: scaled score %lld out of range", (long long)v);
                return GB2_ERR_MOTIF;
            }
            mn = std::min(mn, v);
            mx = std::max(mx, v);
        }
        mincol[j] = mn;
        lo += mn;
        hi += mx;
    }
    if (hi >= L) {
        GB2_SET_ERR(ctx, "gb2_motif_create: max score %lld exceeds the p-value matrix (%lld bins)", (long long)hi, (long long)L);
        return GB2_ERR_MOTIF;
    }
    const int64_t span = hi - lo + 1;
    GB2_CUDA(ctx, cudaSetDevice(ctx->device));

    gb2_motif *m = new (std::nothrow) gb2_motif();
    if (!m) return GB2_ERR_NOMEM;
    m->device = ctx->device;
    m->w = w;
    m->n_chunks = (w + 3) / 4;
    m->lo = lo; m->hi = hi; m->span = span;
    m->min_val = min_val; m->scale = scale; m->offset = offset;

    // ---- 4-base chunk LUT: entry = (rc_rel << 16) | fwd_rel, both relative to the column minima so
    //      that the accumulated fields are directly the histogram bins (score - lo).
    std::vector<uint32_t> lut((size_t)m->n_chunks * 256);
    for (int c = 0; c < m->n_chunks; ++c) {
        for (int idx = 0; idx < 256; ++idx) {
            uint32_t fwd = 0, rc = 0;
            for (int j = 0; j < 4; ++j) {
                int p = 4 * c + j;
                if (p >= w) break;
                int b = (idx >> (2 * j)) & 3;
                fwd += (uint32_t)(sm[(int64_t)b * w + p] - mincol[p]);
                int q = w - 1 - p;  // base at position p of x sits at position q of the reverse complement
                rc += (uint32_t)(sm[(int64_t)(3 - b) * w + q] - mincol[q]);
            }
            lut[(size_t)c * 256 + idx] = (rc << 16) | fwd;
        }
    }
    int rc_ = GB2_OK;
    double *d_pm = nullptr, *d_ctab = nullptr;
    auto align256 = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t b_ptab = align256((size_t)span * sizeof(double));
    const size_t b_lut = align256(lut.size() * sizeof(uint32_t));
    const size_t b_bitmap = align256((size_t)gb2_div_up(span + 1, 32) * sizeof(uint32_t));
    do {
        // one allocation for what the motif keeps; the two work arrays of K4 come from the context's scratch buffer
        // (no cudaMalloc / cudaFree pair per motif: an 800-motif collection is uploaded in a fraction of a second)
        if ((rc_ = gb2_scratch_reserve(ctx, 2 * b_ptab)) != GB2_OK) break;
        d_pm = (double *)ctx->scratch;
        d_ctab = (double *)((char *)ctx->scratch + b_ptab);
        if (cudaMalloc((void **)&m->d_block, b_ptab + b_lut + b_bitmap) != cudaSuccess) {
            GB2_SET_ERR(ctx, "gb2_motif_create: cudaMalloc failed: %s", cudaGetErrorString(cudaGetLastError()));
            rc_ = GB2_ERR_NOMEM;
            break;
        }
        m->d_ptab = (double *)m->d_block;
        m->d_lut = (uint32_t *)((char *)m->d_block + b_ptab);
        m->d_bitmap = (uint32_t *)((char *)m->d_block + b_ptab + b_lut);
        cudaError_t e = cudaMemcpyAsync(m->d_lut, lut.data(), lut.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(d_pm, h_pval_mat + lo, (size_t)span * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) { GB2_SET_ERR(ctx, "gb2_motif_create: upload failed: %s", cudaGetErrorString(e)); rc_ = GB2_ERR_CUDA; break; }
        // K4: p[s] = seqsum(pv[s:]) / seqsum(pv) -- mass outside [lo,hi] is exactly +0.0 and does not change a sum
        rc_ = gb2_launch_ptable(ctx, d_pm, lo, span, d_ctab, m->d_ptab);
        if (rc_ != GB2_OK) break;
        m->h_ptab.resize((size_t)span);
        e = cudaMemcpyAsync(m->h_ptab.data(), m->d_ptab, (size_t)span * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(&m->total, d_ctab, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { GB2_SET_ERR(ctx, "gb2_motif_create: p-table failed: %s", cudaGetErrorString(e)); rc_ = GB2_ERR_CUDA; break; }
        // mass outside the reachable range means the matrix and the p-value matrix do not belong together
        for (int64_t k = 0; k < L; ++k) {
            if ((k < lo || k > hi) && h_pval_mat[k] != 0.0) {
                GB2_SET_ERR(ctx, "gb2_motif_create: p-value matrix has mass at unreachable score %lld", (long long)k);
                rc_ = GB2_ERR_MOTIF;
                break;
            }
        }
        if (rc_ != GB2_OK) break;
        if (!(m->total > 0.0)) { GB2_SET_ERR(ctx, "gb2_motif_create: empty p-value matrix"); rc_ = GB2_ERR_MOTIF; break; }
    } while (0);
    if (rc_ != GB2_OK) { gb2_motif_destroy(m); return rc_; }

    m->monotone = 1;
    for (int64_t k = 1; k < span; ++k)
        if (m->h_ptab[(size_t)k] > m->h_ptab[(size_t)k - 1]) { m->monotone = 0; break; }

    // ---- shared-memory plan of the scoring kernel: replicated LUT + u32 histogram.  Wide motifs whose span does not
    //      fit next to even one copy of the tables keep the histogram in global memory (hist_global).
    const int64_t budget = (int64_t)ctx->max_smem_optin - 1024;
    int64_t hist_bytes = (span + 1) * 4;
    if (span > 65535) {  // the two strands travel as 16-bit fields of one register
        GB2_SET_ERR(ctx, "gb2_motif_create: score span %lld exceeds 65535", (long long)span);
        gb2_motif_destroy(m);
        return GB2_ERR_MOTIF;
    }
    m->hist_global = 0;
    if ((int64_t)m->n_chunks * 1024 + hist_bytes > budget) {
        if (w <= GB2_NARROW_WIDTH) {
            GB2_SET_ERR(ctx, "gb2_motif_create: score span %lld does not fit shared memory", (long long)span);
            gb2_motif_destroy(m);
            return GB2_ERR_MOTIF;
        }
        m->hist_global = 1;
        hist_bytes = 0;
    }
    int R = 32;
    while (R > 1 && (int64_t)m->n_chunks * 1024 * R + hist_bytes > budget) R >>= 1;
    m->replicas = R;
    m->smem_bytes = (int64_t)m->n_chunks * 1024 * R + hist_bytes;
    *out = m;
    return GB2_OK;
}

extern "C" int gb2_motif_get_info(const gb2_motif *m, gb2_motif_info *info)
{
    if (!m || !info) return GB2_ERR_ARG;
    info->width = m->w;
    info->n_chunks = m->n_chunks;
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
