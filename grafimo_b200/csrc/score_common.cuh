// score_common.cuh -- device helpers shared by the scoring kernels (K2 over packed k-mers: score.cu; K2 over 2-bit
// sequences with the windows formed in registers: seqscan.cu).  Not part of the ABI.
#pragma once
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

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int IMM>
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr)
{
    uint32_t v;
    asm("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(IMM));
    return v;
}

// compiles to ATOMS.POPC.INC (same cost as ATOMS.ADD on B200: tools/ubench_atoms.cu, and merges same-address lanes)
__device__ __forceinline__ void red_shared_inc(uint32_t addr)
{
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr) : "memory");
}

// Both strands of one packed k-mer (w0 = bases 0..15, w1 = bases 16..31): per CB-base chunk one index extract, one
// address add and one conflict-free LDS.  lut32 = shared-window address of the replicated table + 4 * (lane & (R-1));
// entry (c, b) sits at lut32 + (c * 4^CB + b) * R * 4.
//   CB = 4: the chunk is a byte of the k-mer (PRMT), 256-entry tables;
//   CB = 3: six bits at a time (shift / funnel shift + mask), 64-entry tables -- a quarter of the memory per chunk, so long
//           motifs still replicate 32x and keep every lookup at one wavefront (context.cu plan_smem).
template <int CB>
struct ChunkGeom {
    static constexpr int ENTRIES = 1 << (2 * CB);
};

template <int CB, int C, int NCHUNK, int R>
struct ChunkSum {
    static __device__ __forceinline__ uint32_t run(uint32_t w0, uint32_t w1, uint32_t lut32)
    {
        uint32_t b;
        if (CB == 4) {
            const uint32_t word = (C < 4) ? w0 : w1;
            b = __byte_perm(word, 0u, 0x4440u + (uint32_t)(C & 3));  // PRMT: byte C of the k-mer
        } else {
            constexpr int bit = 2 * CB * C;
            uint32_t v;
            if (bit + 2 * CB <= 32) v = w0 >> bit;
            else if (bit >= 32) v = w1 >> (bit - 32);
            else v = __funnelshift_r(w0, w1, bit);
            b = v & (uint32_t)(ChunkGeom<CB>::ENTRIES - 1);
        }
        // table offset of chunk C rides in the LDS immediate; the entry offset is one IMAD/LEA
        return lds_u32<C * ChunkGeom<CB>::ENTRIES * R * 4>(lut32 + b * (uint32_t)(R * 4)) +
               ChunkSum<CB, C + 1, NCHUNK, R>::run(w0, w1, lut32);
    }
};
template <int CB, int NCHUNK, int R>
struct ChunkSum<CB, NCHUNK, NCHUNK, R> {
    static __device__ __forceinline__ uint32_t run(uint32_t, uint32_t, uint32_t) { return 0u; }
};

template <int CB, int NCHUNK, int R>
__device__ __forceinline__ uint32_t score_word(uint32_t w0, uint32_t w1, uint32_t lut32)
{
    return ChunkSum<CB, 0, NCHUNK, R>::run(w0, w1, lut32);
}

// Appends the hits of one (k-mer, strand) slot across the warp: one ballot, one global atomic.
__device__ __forceinline__ void append_hits(const ScoreParams &p, bool pred, uint64_t row, uint32_t bin, uint32_t strand,
                                            unsigned lane)
{
    const unsigned m = __ballot_sync(0xFFFFFFFFu, pred);
    if (m == 0) return;
    const int leader = __ffs(m) - 1;
    unsigned long long base = 0;
    if ((int)lane == leader) base = atomicAdd(p.hit_count, (unsigned long long)__popc(m));
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
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


// Host side: the p-value cut-off of a launch as an integer test (hit <=> ptab[bin] < threshold, strict:
// src/grafimo/resultsTmp.py:305-307) -> p.cut / p.bitmap, and the motif's tables -> p.lut / span / lo.
int gb2_fill_score_params(gb2_ctx *ctx, const gb2_motif *m, double p_threshold, ScoreParams &p);
// K2 for motifs wider than 32 bp (score_wide.cu)
int gb2_launch_score_wide(gb2_ctx *ctx, const gb2_motif *m, const ScoreParams &p, int64_t n);
