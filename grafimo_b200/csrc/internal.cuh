// internal.cuh -- shared definitions of the grafimo_b200 CUDA library (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "grafimo_b200.h"

struct gb2_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t copy_stream = nullptr;  // H2D staging stream of gb2_scan_host
    int sm_count = 0;
    int max_smem_optin = 0;
    int64_t launches = 0;
    char err[512] = {0};
    // reusable device scratch (CUB temp storage, sort keys, ...)
    void *scratch = nullptr;
    size_t scratch_bytes = 0;
    // pinned host mailbox for small device->host reads
    uint64_t *h_mail = nullptr;
    // grow-only device pool of the host-buffer entry points (gb2_scan_host*): staging + outputs
    void *pool = nullptr;
    size_t pool_bytes = 0;
    // descriptor arrays of the sequence kernels (seqscan.cu), reused while the layout stays the same
    struct DescSlot {
        void *d_ptr = nullptr;
        size_t bytes = 0;
        uint64_t hash = 0;
        int64_t n_seqs = -1, total_units = 0;
    } desc[2];
    // grow-only pinned host staging of the host-side packer (gb2_scan_host_sequences, host_pack.cpp)
    void *h_pin = nullptr;
    size_t h_pin_bytes = 0;
    // what the last gb2_scan_host* call moved (gb2_scan_last_transfer)
    uint64_t last_h2d_bytes = 0, last_d2h_bytes = 0, last_chunks_given = 0, last_chunks_packed = 0;
    // NCCL communicator of a multi-GPU run (comm.cu); null on one GPU
    void *nccl_comm = nullptr;
    int comm_rank = 0, comm_world = 1;
};

struct gb2_motif_block;  // shared device allocation of the motifs created together (context.cu)

struct gb2_motif {
    int device = 0;
    int w = 0;
    int chunk_bases = 4;           // bases per lookup-table chunk: 4 (256 entries) or 3 (64 entries)
    int n_chunks = 0;
    int replicas = 0;
    int monotone = 0;
    int hist_global = 0;           // 1: span too large for shared memory, K2 counts with global atomics
    int ptab_exact_host = 0;       // 1: the p-value table came from the exact integer suffix sums on the host (context.cu)
    int64_t lo = 0, hi = 0, span = 0;
    int64_t min_val = 0, scale = 0;
    double offset = 0.0, total = 0.0;
    int64_t smem_bytes = 0;
    gb2_motif_block *block = nullptr;  // owner of the device memory below (shared, reference-counted)
    uint32_t *d_lut = nullptr;     // [n_chunks][4^chunk_bases]  (rc_rel << 16 | fwd_rel)
    double *d_ptab = nullptr;      // [span]  p-value of score lo+k
    uint32_t *d_bitmap = nullptr;  // [ceil(span/32)] hit bitmap for non-monotone tables
    std::vector<double> h_ptab;    // host copy of d_ptab
};

#define GB2_SET_ERR(ctx, ...)                                          \
    do {                                                               \
        if (ctx) snprintf((ctx)->err, sizeof((ctx)->err), __VA_ARGS__); \
    } while (0)

#define GB2_CUDA(ctx, call)                                                                         \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            GB2_SET_ERR(ctx, "%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return e__ == cudaErrorMemoryAllocation ? GB2_ERR_NOMEM : GB2_ERR_CUDA;                 \
        }                                                                                           \
    } while (0)

#define GB2_LAUNCH_CHECK(ctx)                \
    do {                                     \
        (ctx)->launches++;                   \
        GB2_CUDA(ctx, cudaGetLastError());   \
    } while (0)

#define GB2_REQUIRE(ctx, cond, ...)      \
    do {                                 \
        if (!(cond)) {                   \
            GB2_SET_ERR(ctx, __VA_ARGS__); \
            return GB2_ERR_ARG;          \
        }                                \
    } while (0)

// grows the context scratch buffer (stream-ordered free of the old one)
int gb2_scratch_reserve(gb2_ctx *ctx, size_t bytes);
// grows the pool of the host-buffer entry points; *out = its base
int gb2_pool_reserve(gb2_ctx *ctx, size_t bytes, char **out);
int gb2_pinned_reserve(gb2_ctx *ctx, size_t bytes, char **out);  // grow-only pinned host buffer of the context
void gb2_comm_release(gb2_ctx *ctx);  // comm.cu
// host_pack.cpp: ASCII bases -> 2-bit words + N bits, the device encoder's output bit for bit
void gb2_host_pack_bases(const uint8_t *text, int64_t n_bases, uint64_t *words, uint32_t *nbits, uint64_t *n_invalid, uint64_t *n_other);
int gb2_host_pack_simd();

static inline int64_t gb2_div_up(int64_t a, int64_t b) { return (a + b - 1) / b; }
