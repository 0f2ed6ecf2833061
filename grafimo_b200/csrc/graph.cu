// graph.cu -- K7: k-mer extraction from a variation graph, on the device.
//
// Replaces the external program the reference shells out to for every BED region,
//     vg find -p REGION -x XG -H GBWT -K w -E > width_w/REGION.tsv      (src/grafimo/extract_regions.py:180,225,326)
// and with it the text round trip into score_seqs (src/grafimo/score_sequences.py:273-293): the walks of every
// region are enumerated here straight into the packed k-mers + side arrays the scoring kernel (K2) consumes.
// First "next" row of SURVEY.md 8(f).  The graph arrays are built on the host by grafimo_b200/vgraph.py.
//
//   gb2_graph_prepare  pass 1: one thread per candidate first base (all regions in one launch) walks the graph
//                      depth-first and counts its w-base walks that lie inside the region; exclusive scan -> row
//                      offsets and the total (returned to the host so that it can size the outputs);
//   gb2_graph_extract  pass 2: the same traversal writes the rows at those offsets: packed k-mer, N flag, start,
//                      stop, ref flag, region index (optionally the node walk) and the walk's haplotype-set ids;
//                      pass 3: eight lanes per row AND those sets and count the haplotypes (coalesced 16-byte loads).
//
// Haplotype frequency of a walk = popcount(AND of the haplotype bit sets of its edges) -- in a DAG a haplotype
// contains the node sequence n1..nk exactly when it takes every edge n_i -> n_i+1 -- which is what
// `vg find -H` asks the GBWT; walks that no haplotype follows are emitted with frequency 0, like vg does.
// Rows come out in (region, first base, depth-first over ascending target node) order: deterministic.
#include <cub/cub.cuh>

#include <stdlib.h>

#include <algorithm>
#include <memory>
#include <new>
#include <thread>
#include <type_traits>

#include "internal.cuh"

#define GB2_NO_CONS 0xFFFFFFFFu
#define WALK_THREADS 32   // one warp per CTA: 8 % faster than 128 (measured), a slow walk holds fewer idle lanes
#define WALK_LIMIT (1u << 24)  // walks from one first base; beyond this the region is reported as too dense
#define FREQ_MAX_CONS 8        // walks with more haplotype sets than this are counted by their own thread (rare)
#define FREQ_GROUP 8           // lanes that share one row in the frequency pass
#define FREQ_IN_POOL 0xFFu     // ncons marker: the row's set list lives in the overflow pool

struct GraphView {
    int64_t n_nodes;
    const uint32_t *node_off;   // [n_nodes+1] first base of the node in seq
    const uint32_t *blk_node;   // [ceil(n_bases/32)] node holding base 32*i: base -> node without a binary search
    const uint8_t *seq;         // base codes 0..3, 4 = anything else
    const uint2 *node_bits;     // per node: its bases 2-bit packed (x = bases 0..15, y = 16..31), or nullptr when a node
    const uint32_t *node_nbits; //           is longer than 32 bases; node_nbits: bit j = base j is not ACGT
    const int64_t *node_a0;     // reference coordinate of base 0 (before clamping)
    const int64_t *node_clamp;  // coordinates are clamped to this (end of the allele's reference span)
    const uint8_t *node_flags;  // bit 0: on the reference path
    const uint32_t *node_cons;  // haplotype-set row of the node or GB2_NO_CONS (every haplotype)
    const uint32_t *edge_off;   // [n_nodes+1] CSR
    const uint32_t *edge_to;
    const uint32_t *edge_cons;  // haplotype-set row of the edge or GB2_NO_CONS
    const uint4 *edge_rec;      // [n_edges][3]: everything the walk needs about the edge's TARGET node, so that a step
                                // along an edge is ONE dependent load instead of two (edge target, then node arrays):
                                //   [0] = {target, haplotype-set row of the edge, node word x, node word y}
                                //   [1] = {non-ACGT bits, length, first out-edge, out-degree | on-reference-path << 31}
                                //   [2] = {a0 lo, a0 hi, clamp lo, clamp hi}
    const uint32_t *cons_bits;  // [n_cons][words]; a row that holds more than half of the haplotypes is stored COMPLEMENTED
    const uint2 *cons_meta;     // [n_cons] {x, y} = 64-bit mask of the 16-byte pieces of the stored row that are not all zero
    const uint8_t *cons_neg;    // [n_cons] 1 = the stored row is the complement of the set
    int32_t n_hap, words;
};

struct QueryView {
    int32_t n_regions, w;
    const int64_t *rs, *re;          // region [start, stop) on the reference path
    const int64_t *node_lo;          // candidate first nodes [node_lo, node_hi)
    const int64_t *node_hi;
    const unsigned long long *tprefix;  // [n_regions+1] threads before region r
};

struct RowsOut {
    uint64_t *packed;
    uint32_t *nmask;
    int64_t *start, *stop;
    int32_t *freq;
    uint8_t *isref;
    uint32_t *region;
    uint32_t *walk;      // [capacity][32] ([capacity][64] when w > 32) node indices, or nullptr
    uint8_t *walk_len;   // nodes in the walk
    uint8_t *walk_off;   // offset of the first base in the first node
    unsigned long long capacity;
    unsigned long long *counts;  // [0] += rows with a non-ACGT base
    uint32_t *cons8;     // scratch [capacity][FREQ_MAX_CONS]: haplotype-set rows of the walk, for the frequency pass
    uint8_t *ncons;      // scratch [capacity]: how many (0 = frequency already final, FREQ_IN_POOL = list is in `pool`)
    uint32_t *pool;      // scratch: haplotype-set rows of the walks with more than FREQ_MAX_CONS of them, back to back;
    unsigned long long pool_cap;    // such a row keeps {offset lo, offset hi, count} in its cons8 slots
    unsigned long long *pool_used;
};

struct gb2_graph {
    int device = 0;
    GraphView v{};
    void *blocks[24] = {nullptr};
    int n_blocks = 0;
    std::vector<uint32_t> h_node_off;  // host copy: threads per region without a device round trip
    // region -> candidate first nodes: two non-decreasing envelopes of the coordinates a node can report as a start
    std::vector<int64_t> h_low_key;    // [i] = min over nodes >= i of their smallest start coordinate
    std::vector<int64_t> h_high_key;   // [i] = max over nodes <= i of their largest start coordinate
    int64_t n_edges = 0, n_bases = 0, n_cons = 0;
    // prepared query
    bool q_valid = false;
    QueryView q{};
    int64_t q_threads = 0;
    unsigned long long q_total = 0;
    void *q_mem = nullptr;             // region arrays
    uint32_t *d_counts = nullptr;      // per thread
    unsigned long long *d_offsets = nullptr;
    int64_t q_cap_threads = 0;
    uint32_t *d_cons8 = nullptr;       // frequency-pass scratch, per row
    uint8_t *d_ncons = nullptr;
    uint32_t *d_pool = nullptr;        // overflow lists of the frequency pass + their fill counter
    unsigned long long *d_pool_used = nullptr;
    unsigned long long q_cap_rows = 0;
    uint32_t *d_flag = nullptr;        // [0] != 0: a first base exceeded WALK_LIMIT
};

template <typename T>
__device__ __forceinline__ int64_t upper_bound_dev(const T *a, int64_t lo, int64_t hi, T x)
{  // first index in [lo, hi) with a[i] > x
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (a[mid] <= x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// Haplotypes that follow a walk = |AND of its haplotype sets|.  Most sets are very sparse (the carriers of a rare
// alternative allele) or the complement of a sparse set (everybody else, on the reference side of the site), so every row
// is stored in its sparser polarity with a 64-bit mask of its non-zero 16-byte pieces, and only the pieces that can
// contribute are read:
//   some set stored as is   -> pieces where ALL such sets are non-zero;          elsewhere the AND is empty
//   only complemented sets  -> pieces where ANY complement is non-zero;         elsewhere every haplotype of the piece counts
// Lanes `sub`, `sub + step`, ... of a group share the pieces; the return value is this lane's share of the count
// (all-complemented lists: of the haplotypes to SUBTRACT from n_hap -- *subtract is set).
__device__ __forceinline__ uint32_t piece_valid_word(int n_hap, int word)
{
    const int left = n_hap - 32 * word;
    return left >= 32 ? 0xFFFFFFFFu : (left <= 0 ? 0u : ((1u << left) - 1u));
}

__device__ __forceinline__ int32_t set_list_count(const GraphView &g, const uint32_t *ids, int n, int sub, int step, bool *subtract)
{
    const int nq = g.words >> 2;
    unsigned long long all_pos = ~0ull, any_neg = 0ull;
    bool has_pos = false;
    for (int c = 0; c < n; ++c) {
        const uint2 m = __ldg(g.cons_meta + ids[c]);
        const unsigned long long mask = ((unsigned long long)m.y << 32) | m.x;
        if (__ldg(g.cons_neg + ids[c])) any_neg |= mask;
        else { all_pos &= mask; has_pos = true; }
    }
    *subtract = !has_pos;
    const unsigned long long visit = has_pos ? all_pos : any_neg;
    int32_t total = 0;
    for (int q = sub; q < nq; q += step) {
        if (q < 64 && !((visit >> q) & 1ull)) continue;  // pieces beyond 64 (more than 8,192 haplotypes) are always read
        uint4 acc = make_uint4(piece_valid_word(g.n_hap, 4 * q), piece_valid_word(g.n_hap, 4 * q + 1),
                               piece_valid_word(g.n_hap, 4 * q + 2), piece_valid_word(g.n_hap, 4 * q + 3));
        const uint4 valid = acc;
        for (int c = 0; c < n; ++c) {
            uint4 x = __ldg(reinterpret_cast<const uint4 *>(g.cons_bits + (size_t)ids[c] * g.words) + q);
            if (__ldg(g.cons_neg + ids[c])) { x.x = ~x.x; x.y = ~x.y; x.z = ~x.z; x.w = ~x.w; }
            acc.x &= x.x; acc.y &= x.y; acc.z &= x.z; acc.w &= x.w;
        }
        const int32_t kept = __popc(acc.x) + __popc(acc.y) + __popc(acc.z) + __popc(acc.w);
        total += has_pos ? kept : (__popc(valid.x) + __popc(valid.y) + __popc(valid.z) + __popc(valid.w)) - kept;
    }
    return total;
}

// one thread counts a whole list (walks whose list fits neither the per-row slots nor the overflow pool)
__device__ __forceinline__ int32_t walk_frequency(const GraphView &g, const uint32_t *cons, int n_cons)
{
    if (g.n_hap == 0) return 0;
    if (n_cons == 0) return g.n_hap;
    bool subtract;
    const int32_t c = set_list_count(g, cons, n_cons, 0, 1, &subtract);
    return subtract ? g.n_hap - c : c;
}

// MAXW = 32: k-mers of one packed word; MAXW = 64: wide k-mers (two words per row, 64-deep stacks)
template <bool WRITE, int MAXW>
__global__ void __launch_bounds__(WALK_THREADS) gb2_graph_walk_kernel(const GraphView g, const QueryView q, int64_t n_threads,
                                                                      uint32_t *__restrict__ counts,
                                                                      const unsigned long long *__restrict__ offsets,
                                                                      const RowsOut out, uint32_t *__restrict__ flag)
{
    const int64_t t = (int64_t)blockIdx.x * WALK_THREADS + threadIdx.x;
    if (t >= n_threads) return;
    const int r = (int)upper_bound_dev<unsigned long long>(q.tprefix, 0, (int64_t)q.n_regions + 1, (unsigned long long)t) - 1;
    const int64_t nlo = q.node_lo[r];
    const uint32_t gbase = g.node_off[nlo] + (uint32_t)((unsigned long long)t - q.tprefix[r]);
    int64_t node0 = g.blk_node[gbase >> 5];  // node of the 32-base block's first base, then a short scan forward
    while (g.node_off[node0 + 1] <= gbase) ++node0;
    const int off0 = (int)(gbase - g.node_off[node0]);
    const int64_t rs = q.rs[r], re = q.re[r];
    const int w = q.w;
    const int64_t start = min(g.node_a0[node0] + off0, g.node_clamp[node0]);
    uint32_t n_found = 0;
    if (start >= rs && start < re) {
        // depth-first traversal; every node on the stack contributes at least one base, so depth < w <= MAXW
        constexpr bool WIDE = MAXW > 32;
        typedef typename std::conditional<WIDE, unsigned long long, uint32_t>::type mask_t;
        uint32_t st_node[MAXW], st_edge[MAXW], st_cons[MAXW];
        uint8_t st_have[MAXW];           // bases collected before the node at this depth
        unsigned long long packed = 0;   // bases 0..31
        unsigned long long packed_hi = 0;  // bases 32..63 (WIDE)
        mask_t nbits = 0;                // bit i: base i of the k-mer is not ACGT
        mask_t nonref = 0;               // bit d: node at depth d is off the reference path
        int depth = 0, have = 0;
        uint32_t st_eend[MAXW];  // one past the last out-edge of the node at this depth
        st_node[0] = (uint32_t)node0;
        st_have[0] = 0;
        st_cons[0] = g.node_cons[node0];
        bool entering = true;
        const unsigned long long row0 = WRITE ? offsets[t] : 0ull;
        // the node being entered: from the node arrays for the start node, from the edge record (already loaded) below
        uint32_t cur_len = g.node_off[node0 + 1] - g.node_off[node0], cur_flags = g.node_flags[node0];
        uint32_t cur_eb = g.edge_off[node0], cur_ee = g.edge_off[node0 + 1], cur_nbits = 0;
        unsigned long long cur_bits = 0;
        long long cur_a0 = g.node_a0[node0], cur_clamp = g.node_clamp[node0];
        if (WRITE && g.node_bits != nullptr) {
            const uint2 nb = __ldg(g.node_bits + node0);
            cur_bits = ((unsigned long long)nb.y << 32) | nb.x;
            cur_nbits = __ldg(g.node_nbits + node0);
        }
        uint32_t cur_edge = 0;  // edge that led to the node being entered (its record holds a0 / clamp)
        while (true) {
            bool pop = false;
            if (entering) {
                const uint32_t n = st_node[depth];
                const int o = depth == 0 ? off0 : 0;
                const int len = (int)cur_len;
                const int take = min(len - o, w - have);
                if (WRITE) {
                    if (g.node_bits != nullptr) {  // the node's bases, shifted into place
                        unsigned long long bits = cur_bits >> (2 * o);
                        if (take < 32) bits &= (1ull << (2 * take)) - 1ull;
                        if (!WIDE || have < 32) {
                            packed |= bits << (2 * have);
                            if (WIDE && have > 0) packed_hi |= bits >> (64 - 2 * have);
                        } else {
                            packed_hi |= bits << (2 * (have - 32));
                        }
                        uint32_t bad = cur_nbits >> o;
                        if (take < 32) bad &= (1u << take) - 1u;
                        nbits |= (mask_t)bad << have;
                    } else {
                        const uint32_t b0 = g.node_off[n];
                        for (int k = 0; k < take; ++k) {
                            const uint32_t c = g.seq[b0 + o + k];
                            const int at = have + k;
                            if (c >= 4u) nbits |= (mask_t)1 << at;
                            else if (!WIDE || at < 32) packed |= (unsigned long long)c << (2 * at);
                            else packed_hi |= (unsigned long long)c << (2 * (at - 32));
                        }
                    }
                }
                if (cur_flags & 1u) nonref &= ~((mask_t)1 << depth); else nonref |= (mask_t)1 << depth;
                have += take;
                if (have < w) {  // node used up: go on through its edges
                    st_edge[depth] = cur_eb;
                    st_eend[depth] = cur_ee;
                    entering = false;
                    continue;
                }
                // a complete walk ending at base o + take - 1 of node n
                if (depth > 0) {
                    const uint4 r2 = __ldg(g.edge_rec + 3ull * cur_edge + 2);
                    cur_a0 = (long long)(((unsigned long long)r2.y << 32) | r2.x);
                    cur_clamp = (long long)(((unsigned long long)r2.w << 32) | r2.z);
                }
                const int64_t stop = min((int64_t)cur_a0 + (o + take), (int64_t)cur_clamp);
                if (stop <= re) {
                    if (WRITE) {
                        const unsigned long long row = row0 + n_found;
                        if (row < out.capacity) {
                            uint32_t cons[MAXW];
                            int nc = 0;
                            if (depth == 0) {
                                if (st_cons[0] != GB2_NO_CONS) cons[nc++] = st_cons[0];
                            } else {
                                for (int d = 1; d <= depth; ++d)
                                    if (st_cons[d] != GB2_NO_CONS) cons[nc++] = st_cons[d];
                            }
                            if (WIDE) reinterpret_cast<ulonglong2 *>(out.packed)[row] = make_ulonglong2(packed, packed_hi);
                            else out.packed[row] = packed;
                            out.start[row] = start;
                            out.stop[row] = stop;
                            unsigned long long pool_at = ~0ull;
                            if (nc > FREQ_MAX_CONS && g.n_hap != 0) {
                                // walks through dense variant clusters: too many sets for the fixed slots.  Counting them
                                // here, one thread reading nc x words, made a few CTAs run several times longer than
                                // the rest of the grid (ncu: one SM busy for the whole kernel, the average SM for 25 %).
                                pool_at = atomicAdd(out.pool_used, (unsigned long long)nc);
                                if (pool_at + nc > out.pool_cap) pool_at = ~0ull;
                            }
                            if (pool_at != ~0ull) {
                                for (int c = 0; c < nc; ++c) out.pool[pool_at + c] = cons[c];
                                out.cons8[row * FREQ_MAX_CONS + 0] = (uint32_t)pool_at;
                                out.cons8[row * FREQ_MAX_CONS + 1] = (uint32_t)(pool_at >> 32);
                                out.cons8[row * FREQ_MAX_CONS + 2] = (uint32_t)nc;
                                out.ncons[row] = (uint8_t)FREQ_IN_POOL;
                            } else if (nc == 0 || nc > FREQ_MAX_CONS || g.n_hap == 0) {
                                out.freq[row] = walk_frequency(g, cons, nc);
                                out.ncons[row] = 0;
                            } else {  // counted by gb2_graph_freq_kernel, FREQ_GROUP lanes per row, coalesced
                                for (int c = 0; c < nc; ++c) out.cons8[row * FREQ_MAX_CONS + c] = cons[c];
                                out.ncons[row] = (uint8_t)nc;
                            }
                            out.isref[row] = (nonref & (((mask_t)2 << depth) - (mask_t)1)) == 0 ? 1 : 0;
                            out.region[row] = (uint32_t)r;
                            if (nbits) {
                                atomicOr(out.nmask + (row >> 5), 1u << (row & 31));
                                atomicAdd(out.counts, 1ull);
                            }
                            if (out.walk != nullptr) {
                                for (int d = 0; d <= depth; ++d) out.walk[row * MAXW + d] = st_node[d];
                                out.walk_len[row] = (uint8_t)(depth + 1);
                                out.walk_off[row] = (uint8_t)off0;
                            }
                        }
                    }
                    if (++n_found >= WALK_LIMIT) {
                        atomicExch(flag, 1u);
                        break;
                    }
                }
                pop = true;
            } else {  // next edge of the node at `depth` (its bases are already in the k-mer)
                const uint32_t e = st_edge[depth];
                if (e < st_eend[depth]) {
                    st_edge[depth] = e + 1;
                    ++depth;
                    const uint4 r1 = __ldg(g.edge_rec + 3ull * e + 1);
                    cur_nbits = r1.x;
                    cur_len = r1.y;
                    cur_eb = r1.z;
                    cur_ee = r1.z + (r1.w & 0x7FFFFFFFu);
                    cur_flags = r1.w >> 31;
                    cur_edge = e;
                    if (WRITE) {
                        const uint4 r0 = __ldg(g.edge_rec + 3ull * e);
                        st_node[depth] = r0.x;
                        st_cons[depth] = r0.y;
                        cur_bits = ((unsigned long long)r0.w << 32) | r0.z;
                    }
                    st_have[depth] = (uint8_t)have;
                    entering = true;
                } else {
                    pop = true;
                }
            }
            if (pop) {
                if (depth == 0) break;
                have = st_have[depth];  // the bases of the node being left go away
                --depth;
                entering = false;
                if (WRITE) {
                    if (!WIDE || have < 32) {  // have < w here
                        packed &= (1ull << (2 * have)) - 1ull;
                        packed_hi = 0;
                    } else {
                        packed_hi &= (1ull << (2 * (have - 32))) - 1ull;
                    }
                    nbits &= ((mask_t)1 << have) - (mask_t)1;
                }
            }
        }
    }
    if (!WRITE) counts[t] = n_found;
}

// Frequency pass: FREQ_GROUP lanes per row AND the row's haplotype sets 16 bytes at a time (neighbouring lanes read
// neighbouring 16-byte pieces of the same set) and add up the population counts.
__global__ void __launch_bounds__(256) gb2_graph_freq_kernel(const GraphView g, unsigned long long n_rows,
                                                             const uint32_t *__restrict__ cons8,
                                                             const uint8_t *__restrict__ ncons,
                                                             const uint32_t *__restrict__ pool, int32_t *__restrict__ freq)
{
    const unsigned long long t = (unsigned long long)blockIdx.x * 256 + threadIdx.x;
    const unsigned long long row = t / FREQ_GROUP;
    const int sub = (int)(t % FREQ_GROUP);
    int nc = 0;
    if (row < n_rows) nc = ncons[row];
    int32_t total = 0;
    bool subtract = false;
    if (nc == (int)FREQ_IN_POOL) {  // long list in the overflow pool
        const unsigned long long at = ((unsigned long long)cons8[row * FREQ_MAX_CONS + 1] << 32) | cons8[row * FREQ_MAX_CONS];
        total = set_list_count(g, pool + at, (int)cons8[row * FREQ_MAX_CONS + 2], sub, FREQ_GROUP, &subtract);
    } else if (nc) {
        uint32_t ids[FREQ_MAX_CONS];
#pragma unroll
        for (int c = 0; c < FREQ_MAX_CONS; ++c) ids[c] = c < nc ? cons8[row * FREQ_MAX_CONS + c] : 0u;
        total = set_list_count(g, ids, nc, sub, FREQ_GROUP, &subtract);
    }
    // the FREQ_GROUP lanes of a row are neighbours inside one warp
#pragma unroll
    for (int d = FREQ_GROUP / 2; d >= 1; d >>= 1) total += __shfl_xor_sync(0xFFFFFFFFu, total, d);
    if (nc && sub == 0) freq[row] = subtract ? g.n_hap - total : total;
}

// ---------------------------------------------------------------------------------------------------------
// host-side preparation loops of gb2_graph_create over nodes / edges / set rows: independent items, split over threads
template <typename F>
static void parallel_for(int64_t n, int64_t grain, F fn)
{
    const char *e = getenv("GB2_BUILD_THREADS");
    int nt = e ? atoi(e) : (int)std::thread::hardware_concurrency();
    nt = (int)std::max<int64_t>(1, std::min<int64_t>(nt, n / std::max<int64_t>(grain, 1)));
    if (nt <= 1) { fn((int64_t)0, n); return; }
    std::vector<std::thread> pool;
    const int64_t per = (n + nt - 1) / nt;
    for (int t = 0; t < nt; ++t) {
        const int64_t lo = t * per, hi = std::min(n, lo + per);
        if (lo < hi) pool.emplace_back([=] { fn(lo, hi); });
    }
    for (auto &t : pool) t.join();
}

template <typename T>
static int upload(gb2_ctx *ctx, gb2_graph *g, const T *h, size_t n, const T **d_out)
{
    void *d = nullptr;
    const size_t bytes = std::max<size_t>(n, 1) * sizeof(T);
    GB2_CUDA(ctx, cudaMalloc(&d, bytes + 16));
    g->blocks[g->n_blocks++] = d;
    if (n) GB2_CUDA(ctx, cudaMemcpyAsync(d, h, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    *d_out = (const T *)d;
    return GB2_OK;
}

extern "C" int gb2_graph_destroy(gb2_graph *g)
{
    if (!g) return GB2_OK;
    cudaSetDevice(g->device);
    for (int i = 0; i < g->n_blocks; ++i)
        if (g->blocks[i]) cudaFree(g->blocks[i]);
    if (g->q_mem) cudaFree(g->q_mem);
    if (g->d_counts) cudaFree(g->d_counts);
    if (g->d_offsets) cudaFree(g->d_offsets);
    if (g->d_flag) cudaFree(g->d_flag);
    if (g->d_cons8) cudaFree(g->d_cons8);
    if (g->d_ncons) cudaFree(g->d_ncons);
    if (g->d_pool) cudaFree(g->d_pool);
    if (g->d_pool_used) cudaFree(g->d_pool_used);
    delete g;
    return GB2_OK;
}

extern "C" int gb2_graph_create(gb2_ctx *ctx, int64_t n_nodes, const uint32_t *h_node_off, const uint8_t *h_seq,
                                const int64_t *h_node_a0, const int64_t *h_node_clamp, const uint8_t *h_node_flags,
                                const uint32_t *h_node_cons, int64_t n_edges, const uint32_t *h_edge_off,
                                const uint32_t *h_edge_to, const uint32_t *h_edge_cons, int32_t n_hap, int32_t words,
                                int64_t n_cons, const uint32_t *h_cons_bits, gb2_graph **out)
{
    if (!ctx || !out) return GB2_ERR_ARG;
    *out = nullptr;
    GB2_REQUIRE(ctx, n_nodes >= 1 && n_nodes < ((int64_t)1 << 31), "gb2_graph_create: node count out of range");
    GB2_REQUIRE(ctx, h_node_off && h_seq && h_node_a0 && h_node_clamp && h_node_flags && h_node_cons && h_edge_off,
                "gb2_graph_create: null array");
    GB2_REQUIRE(ctx, n_edges >= 0 && (n_edges == 0 || (h_edge_to && h_edge_cons)), "gb2_graph_create: null edge array");
    GB2_REQUIRE(ctx, n_hap >= 0 && words >= 4 && (words & 3) == 0 && (int64_t)words * 32 >= n_hap,
                "gb2_graph_create: haplotype bit sets need a multiple of 4 words covering %d haplotypes", n_hap);
    GB2_REQUIRE(ctx, n_cons >= 0 && (n_cons == 0 || h_cons_bits), "gb2_graph_create: null haplotype sets");
    GB2_REQUIRE(ctx, h_edge_off[n_nodes] == (uint32_t)n_edges, "gb2_graph_create: edge offsets do not end at n_edges");
    for (int64_t i = 0; i < n_nodes; ++i) {
        GB2_REQUIRE(ctx, h_node_off[i + 1] > h_node_off[i], "gb2_graph_create: node %lld is empty", (long long)i);
        GB2_REQUIRE(ctx, h_edge_off[i + 1] >= h_edge_off[i], "gb2_graph_create: edge offsets not monotone");
        GB2_REQUIRE(ctx, h_node_cons[i] == GB2_NO_CONS || (int64_t)h_node_cons[i] < n_cons, "gb2_graph_create: bad node set index");
    }
    for (int64_t e = 0; e < n_edges; ++e) {
        GB2_REQUIRE(ctx, (int64_t)h_edge_to[e] < n_nodes, "gb2_graph_create: edge target out of range");
        GB2_REQUIRE(ctx, h_edge_cons[e] == GB2_NO_CONS || (int64_t)h_edge_cons[e] < n_cons, "gb2_graph_create: bad edge set index");
    }
    GB2_CUDA(ctx, cudaSetDevice(ctx->device));
    gb2_graph *g = new (std::nothrow) gb2_graph();
    if (!g) return GB2_ERR_NOMEM;
    g->device = ctx->device;
    const size_t n_bases = h_node_off[n_nodes];
    int rc = GB2_OK;
    do {
        if ((rc = upload(ctx, g, h_node_off, (size_t)n_nodes + 1, &g->v.node_off)) != GB2_OK) break;
        if ((rc = upload(ctx, g, h_seq, n_bases, &g->v.seq)) != GB2_OK) break;
        // nodes of at most 32 bases (vg's default): their bases as one 64-bit word each (also copied into the edge records)
        bool fits = true;
        for (int64_t i = 0; i < n_nodes && fits; ++i) fits = h_node_off[i + 1] - h_node_off[i] <= 32u;
        std::vector<uint2> nb(fits ? (size_t)n_nodes : 0);
        std::vector<uint32_t> bad(fits ? (size_t)n_nodes : 0);
        if (fits) {
            parallel_for(n_nodes, 1 << 16, [&](int64_t lo, int64_t hi) {
                for (int64_t i = lo; i < hi; ++i) {
                    unsigned long long bits = 0;
                    uint32_t m = 0;
                    const uint32_t len = h_node_off[i + 1] - h_node_off[i];
                    for (uint32_t j = 0; j < len; ++j) {
                        const uint8_t c = h_seq[h_node_off[i] + j];
                        if (c < 4) bits |= (unsigned long long)c << (2 * j); else m |= 1u << j;
                    }
                    nb[(size_t)i] = make_uint2((uint32_t)bits, (uint32_t)(bits >> 32));
                    bad[(size_t)i] = m;
                }
            });
            if ((rc = upload(ctx, g, nb.data(), (size_t)n_nodes, &g->v.node_bits)) != GB2_OK) break;
            if ((rc = upload(ctx, g, bad.data(), (size_t)n_nodes, &g->v.node_nbits)) != GB2_OK) break;
        }
        {
            std::vector<uint32_t> blk((n_bases + 31) / 32);
            int64_t nd = 0;
            for (size_t b = 0; b < blk.size(); ++b) {
                while (h_node_off[nd + 1] <= b * 32) ++nd;
                blk[b] = (uint32_t)nd;
            }
            if ((rc = upload(ctx, g, blk.data(), blk.size(), &g->v.blk_node)) != GB2_OK) break;
            if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { rc = GB2_ERR_CUDA; break; }
        }
        if ((rc = upload(ctx, g, h_node_a0, (size_t)n_nodes, &g->v.node_a0)) != GB2_OK) break;
        if ((rc = upload(ctx, g, h_node_clamp, (size_t)n_nodes, &g->v.node_clamp)) != GB2_OK) break;
        if ((rc = upload(ctx, g, h_node_flags, (size_t)n_nodes, &g->v.node_flags)) != GB2_OK) break;
        if ((rc = upload(ctx, g, h_node_cons, (size_t)n_nodes, &g->v.node_cons)) != GB2_OK) break;
        if ((rc = upload(ctx, g, h_edge_off, (size_t)n_nodes + 1, &g->v.edge_off)) != GB2_OK) break;
        if ((rc = upload(ctx, g, h_edge_to, (size_t)n_edges, &g->v.edge_to)) != GB2_OK) break;
        if ((rc = upload(ctx, g, h_edge_cons, (size_t)n_edges, &g->v.edge_cons)) != GB2_OK) break;
        {   // edge records: what a walk needs about the edge's target, in one place (see GraphView::edge_rec)
            std::vector<uint4> rec((size_t)n_edges * 3);
            parallel_for(n_edges, 1 << 16, [&](int64_t lo, int64_t hi) {
                for (int64_t e = lo; e < hi; ++e) {
                    const uint32_t t = h_edge_to[e];
                    const uint32_t len = h_node_off[t + 1] - h_node_off[t];
                    const uint2 bits = fits ? nb[(size_t)t] : make_uint2(0u, 0u);
                    const uint32_t nbad = fits ? bad[(size_t)t] : 0u;
                    const uint32_t deg = h_edge_off[t + 1] - h_edge_off[t];
                    rec[(size_t)3 * e] = make_uint4(t, h_edge_cons[e], bits.x, bits.y);
                    rec[(size_t)3 * e + 1] = make_uint4(nbad, len, h_edge_off[t], deg | ((uint32_t)(h_node_flags[t] & 1u) << 31));
                    const unsigned long long a0 = (unsigned long long)h_node_a0[t], cl = (unsigned long long)h_node_clamp[t];
                    rec[(size_t)3 * e + 2] = make_uint4((uint32_t)a0, (uint32_t)(a0 >> 32), (uint32_t)cl, (uint32_t)(cl >> 32));
                }
            });
            if ((rc = upload(ctx, g, rec.data(), rec.size(), &g->v.edge_rec)) != GB2_OK) break;
            if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { rc = GB2_ERR_CUDA; break; }  // rec goes out of scope
        }
        {   // every set in its sparser polarity + the mask of its non-zero 16-byte pieces (see set_list_count)
            std::unique_ptr<uint32_t[]> bits_mem(new (std::nothrow) uint32_t[std::max<size_t>((size_t)n_cons * words, 1)]);
            if (!bits_mem) { rc = GB2_ERR_NOMEM; break; }
            uint32_t *bits = bits_mem.get();  // filled by the worker threads below (no single-threaded copy or zero fill)
            std::vector<uint2> meta((size_t)n_cons);
            std::vector<uint8_t> neg((size_t)n_cons, 0);
            parallel_for(n_cons, 4096, [&](int64_t lo, int64_t hi) {
                for (int64_t k = lo; k < hi; ++k) {
                    uint32_t *row = bits + (size_t)k * words;
                    memcpy(row, h_cons_bits + (size_t)k * words, (size_t)words * sizeof(uint32_t));
                    int64_t members = 0;
                    for (int wd = 0; wd < words; ++wd) members += __builtin_popcount(row[wd]);
                    if (2 * members > n_hap) {
                        neg[(size_t)k] = 1;
                        for (int wd = 0; wd < words; ++wd) {
                            const int left = n_hap - 32 * wd;
                            const uint32_t valid = left >= 32 ? 0xFFFFFFFFu : (left <= 0 ? 0u : ((1u << left) - 1u));
                            row[wd] = ~row[wd] & valid;
                        }
                    }
                    unsigned long long mask = 0;
                    for (int q = 0; q < words / 4; ++q) {
                        const bool nz = (row[4 * q] | row[4 * q + 1] | row[4 * q + 2] | row[4 * q + 3]) != 0u;
                        if (q < 64 && nz) mask |= 1ull << q;  // pieces beyond 64 are always read
                    }
                    meta[(size_t)k] = make_uint2((uint32_t)mask, (uint32_t)(mask >> 32));
                }
            });
            if ((rc = upload(ctx, g, (const uint32_t *)bits, (size_t)n_cons * words, &g->v.cons_bits)) != GB2_OK) break;
            if ((rc = upload(ctx, g, meta.data(), (size_t)n_cons, &g->v.cons_meta)) != GB2_OK) break;
            if ((rc = upload(ctx, g, neg.data(), (size_t)n_cons, &g->v.cons_neg)) != GB2_OK) break;
            if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { rc = GB2_ERR_CUDA; break; }  // the vectors go out of scope
        }
        if (cudaMalloc((void **)&g->d_flag, sizeof(uint32_t)) != cudaSuccess) { rc = GB2_ERR_NOMEM; break; }
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { rc = GB2_ERR_CUDA; break; }  // host arrays may go away
    } while (0);
    if (rc != GB2_OK) {
        if (!ctx->err[0]) GB2_SET_ERR(ctx, "gb2_graph_create: upload failed");
        gb2_graph_destroy(g);
        return rc;
    }
    g->h_node_off.assign(h_node_off, h_node_off + n_nodes + 1);
    g->h_low_key.resize((size_t)n_nodes);
    g->h_high_key.resize((size_t)n_nodes);
    for (int64_t i = 0; i < n_nodes; ++i) {
        const int64_t len = (int64_t)h_node_off[i + 1] - h_node_off[i];
        const int64_t hi = std::min(h_node_a0[i] + len - 1, h_node_clamp[i]);
        g->h_high_key[(size_t)i] = i ? std::max(g->h_high_key[(size_t)i - 1], hi) : hi;
    }
    for (int64_t i = n_nodes - 1; i >= 0; --i) {
        const int64_t lo = std::min(h_node_a0[i], h_node_clamp[i]);
        g->h_low_key[(size_t)i] = i + 1 < n_nodes ? std::min(g->h_low_key[(size_t)i + 1], lo) : lo;
    }
    g->n_edges = n_edges;
    g->n_bases = (int64_t)n_bases;
    g->n_cons = n_cons;
    g->v.n_nodes = n_nodes;
    g->v.n_hap = n_hap;
    g->v.words = words;
    *out = g;
    return GB2_OK;
}

struct U32ToU64 {
    __host__ __device__ unsigned long long operator()(uint32_t x) const { return (unsigned long long)x; }
};

extern "C" int gb2_graph_get_info(const gb2_graph *g, gb2_graph_info *info)
{
    if (!g || !info) return GB2_ERR_ARG;
    info->n_nodes = g->v.n_nodes; info->n_edges = g->n_edges; info->n_bases = g->n_bases; info->n_sets = g->n_cons;
    info->n_hap = g->v.n_hap; info->words = g->v.words;
    return GB2_OK;
}

extern "C" int gb2_graph_prepare(gb2_ctx *ctx, gb2_graph *g, int32_t n_regions, const int64_t *h_start,
                                 const int64_t *h_stop, int w, uint64_t *h_n_rows)
{
    if (!ctx || !g || !h_n_rows) return GB2_ERR_ARG;
    *h_n_rows = 0;
    g->q_valid = false;
    GB2_REQUIRE(ctx, g->device == ctx->device, "gb2_graph_prepare: graph lives on device %d, context on %d", g->device, ctx->device);
    GB2_REQUIRE(ctx, w >= 1 && w <= GB2_MAX_WIDTH, "gb2_graph_prepare: width %d outside [1,%d]", w, GB2_MAX_WIDTH);
    GB2_REQUIRE(ctx, n_regions >= 0 && (n_regions == 0 || (h_start && h_stop)), "gb2_graph_prepare: null region array");
    GB2_CUDA(ctx, cudaSetDevice(ctx->device));
    // candidate first nodes of a region [rs, re): those that can report a start in it.  Nodes up to the last one whose
    // running-maximum start is < rs are out, and so are the nodes from the first one whose running-minimum start is >= re.
    std::vector<int64_t> node_lo((size_t)n_regions), node_hi((size_t)n_regions);
    std::vector<unsigned long long> tprefix((size_t)n_regions + 1, 0ull);
    for (int r = 0; r < n_regions; ++r) {
        GB2_REQUIRE(ctx, h_start[r] <= h_stop[r], "gb2_graph_prepare: region %d has start > stop", r);
        const int64_t lo = std::lower_bound(g->h_high_key.begin(), g->h_high_key.end(), h_start[r]) - g->h_high_key.begin();
        const int64_t hi = std::lower_bound(g->h_low_key.begin(), g->h_low_key.end(), h_stop[r]) - g->h_low_key.begin();
        node_lo[(size_t)r] = lo;
        node_hi[(size_t)r] = std::max(lo, hi);
        tprefix[(size_t)r + 1] = tprefix[(size_t)r] + (g->h_node_off[(size_t)node_hi[(size_t)r]] - g->h_node_off[(size_t)lo]);
    }
    const int64_t *h_node_lo = node_lo.data(), *h_node_hi = node_hi.data();
    const int64_t T = (int64_t)tprefix[(size_t)n_regions];
    g->q_threads = T;
    g->q_total = 0;
    g->q.n_regions = n_regions;
    g->q.w = w;
    if (T == 0) {
        g->q_valid = true;
        return GB2_OK;
    }
    // region arrays on the device
    if (g->q_mem) { GB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); cudaFree(g->q_mem); g->q_mem = nullptr; }
    const size_t nr = (size_t)n_regions;
    GB2_CUDA(ctx, cudaMalloc(&g->q_mem, (4 * nr + nr + 1) * 8 + 64));
    int64_t *base = (int64_t *)g->q_mem;
    GB2_CUDA(ctx, cudaMemcpyAsync(base, h_start, nr * 8, cudaMemcpyHostToDevice, ctx->stream));
    GB2_CUDA(ctx, cudaMemcpyAsync(base + nr, h_stop, nr * 8, cudaMemcpyHostToDevice, ctx->stream));
    GB2_CUDA(ctx, cudaMemcpyAsync(base + 2 * nr, h_node_lo, nr * 8, cudaMemcpyHostToDevice, ctx->stream));
    GB2_CUDA(ctx, cudaMemcpyAsync(base + 3 * nr, h_node_hi, nr * 8, cudaMemcpyHostToDevice, ctx->stream));
    GB2_CUDA(ctx, cudaMemcpyAsync(base + 4 * nr, tprefix.data(), (nr + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    g->q.rs = base; g->q.re = base + nr; g->q.node_lo = base + 2 * nr; g->q.node_hi = base + 3 * nr;
    g->q.tprefix = (const unsigned long long *)(base + 4 * nr);
    if (T > g->q_cap_threads) {
        GB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (g->d_counts) cudaFree(g->d_counts);
        if (g->d_offsets) cudaFree(g->d_offsets);
        g->d_counts = nullptr; g->d_offsets = nullptr; g->q_cap_threads = 0;
        GB2_CUDA(ctx, cudaMalloc((void **)&g->d_counts, (size_t)(T + 1) * sizeof(uint32_t)));
        GB2_CUDA(ctx, cudaMalloc((void **)&g->d_offsets, (size_t)(T + 1) * sizeof(unsigned long long)));
        g->q_cap_threads = T;
    }
    GB2_CUDA(ctx, cudaMemsetAsync(g->d_flag, 0, sizeof(uint32_t), ctx->stream));
    GB2_CUDA(ctx, cudaMemsetAsync(g->d_counts + T, 0, sizeof(uint32_t), ctx->stream));
    const int64_t grid = gb2_div_up(T, WALK_THREADS);
    GB2_REQUIRE(ctx, grid < ((int64_t)1 << 31), "gb2_graph_prepare: too many candidate bases (%lld)", (long long)T);
    RowsOut none{};
    if (w > GB2_NARROW_WIDTH)
        gb2_graph_walk_kernel<false, 64><<<(unsigned)grid, WALK_THREADS, 0, ctx->stream>>>(g->v, g->q, T, g->d_counts, nullptr, none, g->d_flag);
    else
        gb2_graph_walk_kernel<false, 32><<<(unsigned)grid, WALK_THREADS, 0, ctx->stream>>>(g->v, g->q, T, g->d_counts, nullptr, none, g->d_flag);
    GB2_LAUNCH_CHECK(ctx);
    // exclusive scan over T+1 counts (the extra zero makes offsets[T] the total)
    size_t cub_bytes = 0;
    cub::TransformInputIterator<unsigned long long, U32ToU64, const uint32_t *> it(g->d_counts, U32ToU64());
    cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, it, g->d_offsets, (int64_t)(T + 1), ctx->stream);
    int rc = gb2_scratch_reserve(ctx, cub_bytes);
    if (rc != GB2_OK) return rc;
    GB2_CUDA(ctx, cub::DeviceScan::ExclusiveSum(ctx->scratch, cub_bytes, it, g->d_offsets, (int64_t)(T + 1), ctx->stream));
    ctx->launches += 2;
    GB2_CUDA(ctx, cudaMemcpyAsync(ctx->h_mail, g->d_offsets + T, 8, cudaMemcpyDeviceToHost, ctx->stream));
    GB2_CUDA(ctx, cudaMemcpyAsync(ctx->h_mail + 1, g->d_flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
    GB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if ((uint32_t)ctx->h_mail[1] != 0) {
        GB2_SET_ERR(ctx, "gb2_graph_prepare: more than %u walks start at one base (variants too dense for width %d)", WALK_LIMIT, w);
        return GB2_ERR_CAPACITY;
    }
    g->q_total = ctx->h_mail[0];
    g->q_valid = true;
    *h_n_rows = g->q_total;
    return GB2_OK;
}

extern "C" int gb2_graph_extract(gb2_ctx *ctx, gb2_graph *g, uint64_t capacity, uint64_t *d_packed, uint32_t *d_nmask,
                                 int64_t *d_start, int64_t *d_stop, int32_t *d_freq, uint8_t *d_isref, uint32_t *d_region,
                                 uint32_t *d_walk, uint8_t *d_walk_len, uint8_t *d_walk_off, uint64_t *d_counts)
{
    if (!ctx || !g) return GB2_ERR_ARG;
    if (!g->q_valid) {
        GB2_SET_ERR(ctx, "gb2_graph_extract: call gb2_graph_prepare first");
        return GB2_ERR_STATE;
    }
    if (g->q_total > capacity) {
        GB2_SET_ERR(ctx, "gb2_graph_extract: %llu rows, capacity %llu", (unsigned long long)g->q_total, (unsigned long long)capacity);
        return GB2_ERR_CAPACITY;
    }
    if (g->q_total == 0) return GB2_OK;
    GB2_REQUIRE(ctx, d_packed && d_nmask && d_start && d_stop && d_freq && d_isref && d_region && d_counts, "gb2_graph_extract: null output");
    GB2_REQUIRE(ctx, d_walk == nullptr || (d_walk_len && d_walk_off), "gb2_graph_extract: walk output needs its length/offset arrays");
    GB2_CUDA(ctx, cudaSetDevice(ctx->device));
    RowsOut o;
    o.packed = d_packed; o.nmask = d_nmask; o.start = d_start; o.stop = d_stop; o.freq = d_freq; o.isref = d_isref;
    o.region = d_region; o.walk = d_walk; o.walk_len = d_walk_len; o.walk_off = d_walk_off;
    o.capacity = capacity; o.counts = (unsigned long long *)d_counts;
    if (g->q_total > g->q_cap_rows) {
        GB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (g->d_cons8) cudaFree(g->d_cons8);
        if (g->d_ncons) cudaFree(g->d_ncons);
        if (g->d_pool) cudaFree(g->d_pool);
        g->d_cons8 = nullptr; g->d_ncons = nullptr; g->d_pool = nullptr; g->q_cap_rows = 0;
        GB2_CUDA(ctx, cudaMalloc((void **)&g->d_cons8, (size_t)g->q_total * FREQ_MAX_CONS * sizeof(uint32_t)));
        GB2_CUDA(ctx, cudaMalloc((void **)&g->d_ncons, (size_t)g->q_total));
        GB2_CUDA(ctx, cudaMalloc((void **)&g->d_pool, ((size_t)g->q_total * 4 + 1024) * sizeof(uint32_t)));
        g->q_cap_rows = g->q_total;
    }
    if (!g->d_pool_used) GB2_CUDA(ctx, cudaMalloc((void **)&g->d_pool_used, sizeof(unsigned long long)));
    GB2_CUDA(ctx, cudaMemsetAsync(g->d_pool_used, 0, sizeof(unsigned long long), ctx->stream));
    o.cons8 = g->d_cons8;
    o.ncons = g->d_ncons;
    o.pool = g->d_pool;
    o.pool_cap = (unsigned long long)g->q_cap_rows * 4 + 1024;  // beyond it a walk is counted by its own thread (correct, slower)
    o.pool_used = g->d_pool_used;
    GB2_CUDA(ctx, cudaMemsetAsync(d_nmask, 0, (size_t)gb2_div_up((int64_t)g->q_total, 32) * sizeof(uint32_t), ctx->stream));
    const int64_t T = g->q_threads;
    const int64_t grid = gb2_div_up(T, WALK_THREADS);
    if (g->q.w > GB2_NARROW_WIDTH) {
        GB2_REQUIRE(ctx, ((uintptr_t)d_packed & 15u) == 0, "gb2_graph_extract: wide k-mers need a 16-byte aligned output");
        gb2_graph_walk_kernel<true, 64><<<(unsigned)grid, WALK_THREADS, 0, ctx->stream>>>(g->v, g->q, T, nullptr, g->d_offsets, o, g->d_flag);
    } else {
        gb2_graph_walk_kernel<true, 32><<<(unsigned)grid, WALK_THREADS, 0, ctx->stream>>>(g->v, g->q, T, nullptr, g->d_offsets, o, g->d_flag);
    }
    GB2_LAUNCH_CHECK(ctx);
    if (g->v.n_hap > 0) {
        const int64_t fgrid = gb2_div_up((int64_t)g->q_total * FREQ_GROUP, 256);
        GB2_REQUIRE(ctx, fgrid < ((int64_t)1 << 31), "gb2_graph_extract: too many rows for one launch (%llu)", (unsigned long long)g->q_total);
        gb2_graph_freq_kernel<<<(unsigned)fgrid, 256, 0, ctx->stream>>>(g->v, g->q_total, g->d_cons8, g->d_ncons, g->d_pool, d_freq);
        GB2_LAUNCH_CHECK(ctx);
    }
    return GB2_OK;
}
