// graph_build.cu -- host-side construction of the variation graph K7 walks (gb2_graph_build).
//
// Replaces, for the graph path, what the reference delegates to external programs: `vg construct -r REF -v VCF` and
// `vg index -G gbwt -x xg` (src/grafimo/constructVG.py:332,394-396).  Input is the reference sequence plus reduced,
// position-sorted alleles and the phased genotypes as bit sets; output is the flat graph of csrc/graph.cu, created
// directly on the device.  Same model as grafimo_b200/vgraph.py (which stays as the readable second implementation the
// tests compare this one with) and oracle/graph_oracle.py:
//   * the reference is cut at every allele boundary; per breakpoint the non-empty alternative alleles (input order),
//     then the reference segment; items longer than max_node_len are chained; a deletion is an edge;
//   * everything that ends at a breakpoint is joined to everything that starts there, an insertion sits in between;
//   * a haplotype follows its alleles; one that is inside an allele it took never arrives at the breakpoints under it;
//   * one bit set per node (haplotypes through it) and per edge (haplotypes along it); identical sets share a row and
//     "every haplotype" is not stored.
// One pass over the breakpoints, bit-set work proportional to (variants x words): ~10 ms per Mb at 2,504 haplotypes.
#include <stdlib.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <map>
#include <thread>
#include <unordered_map>
#include <utility>

#include "internal.cuh"

#define GB2_NO_CONS 0xFFFFFFFFu

namespace {

// std::vector whose resize() leaves new elements uninitialised: the big arrays are sized once and then filled by worker
// threads (a value-initialising resize would zero 1 GB on one thread first)
template <typename T>
struct NoInitAlloc : std::allocator<T> {
    template <typename U> struct rebind { typedef NoInitAlloc<U> other; };
    NoInitAlloc() = default;
    template <typename U> NoInitAlloc(const NoInitAlloc<U> &) {}
    template <typename U> void construct(U *p) noexcept { ::new ((void *)p) U; }
    template <typename U, typename... A> void construct(U *p, A &&...a) { ::new ((void *)p) U(std::forward<A>(a)...); }
};
typedef std::vector<uint32_t, NoInitAlloc<uint32_t>> RowVec;

typedef std::vector<uint32_t> Bits;

struct SetTable {
    int words = 0;
    bool on = false;
    Bits full;
    RowVec flat;                 // rows back to back
    std::unordered_multimap<uint64_t, uint32_t> index;

    static uint64_t hash(const uint32_t *p, int n)
    {
        uint64_t h = 1469598103934665603ull;
        for (int i = 0; i < n; ++i) {
            h ^= p[i];
            h *= 1099511628211ull;
            h ^= h >> 29;
        }
        return h;
    }
    std::vector<uint64_t> hashes;  // per row: what `index` was keyed with (reused when ranges are stitched together)
    bool index_stale = false;      // rows were placed without their index entries: rebuild before the next lookup
    uint32_t id(const uint32_t *p)
    {
        if (!on) return GB2_NO_CONS;
        if (memcmp(p, full.data(), (size_t)words * 4) == 0) return GB2_NO_CONS;
        if (index_stale) {
            index.clear();
            index.reserve(hashes.size());
            for (size_t k = 0; k < hashes.size(); ++k) index.emplace(hashes[k], (uint32_t)k);
            index_stale = false;
        }
        return id_hashed(p, hash(p, words));
    }
    // the row is known not to be the full set and h == hash(p, words)
    uint32_t id_hashed(const uint32_t *p, uint64_t h)
    {
        auto range = index.equal_range(h);
        for (auto it = range.first; it != range.second; ++it)
            if (memcmp(p, flat.data() + (size_t)it->second * words, (size_t)words * 4) == 0) return it->second;
        const uint32_t k = (uint32_t)(flat.size() / (size_t)words);
        flat.insert(flat.end(), p, p + words);
        hashes.push_back(h);
        index.emplace(h, k);
        return k;
    }
    const uint32_t *row(uint32_t k) const { return flat.data() + (size_t)k * words; }
};

struct Source {
    uint32_t node;
    Bits set;  // haplotypes whose last node is `node` when they reach the breakpoint (empty vector when sets are off)
};

struct Edge {
    uint32_t src, dst, cons;
};

inline uint8_t code_of(uint8_t c)
{
    switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 4;
    }
}

inline void and_into(Bits &dst, const Bits &a, const uint32_t *b)
{
    for (size_t i = 0; i < dst.size(); ++i) dst[i] = a[i] & b[i];
}

// what the host pass produces: the flat arrays gb2_graph_create uploads
struct HostGraph {
    std::vector<uint32_t> node_off, node_cons, edge_off, edge_to, edge_cons;
    RowVec cons_bits;
    std::vector<uint8_t> seq, flags;
    std::vector<int64_t> a0, clamp;
    int32_t n_hap = 0, words = 4;
    int64_t n_cons = 0;
    int64_t n_ranges = 1;  // independent breakpoint ranges the host pass was cut into
};

// error text of a host pass running on a worker thread (GB2_REQUIRE / GB2_SET_ERR only need an `err` array)
struct ErrSink {
    char err[512] = {0};
};

}  // namespace

extern "C" int gb2_graph_create(gb2_ctx *ctx, int64_t n_nodes, const uint32_t *h_node_off, const uint8_t *h_seq,
                                const int64_t *h_node_a0, const int64_t *h_node_clamp, const uint8_t *h_node_flags,
                                const uint32_t *h_node_cons, int64_t n_edges, const uint32_t *h_edge_off,
                                const uint32_t *h_edge_to, const uint32_t *h_edge_cons, int32_t n_hap, int32_t words,
                                int64_t n_cons, const uint32_t *h_cons_bits, gb2_graph **out);

#define GB2_PREV_NODE 0xFFFFFFFFu  // chunk-local edge source: the last node of the previous chunk

struct BuildInputs {
    const uint8_t *h_ref;
    int64_t ref_len, n_variants;
    const int64_t *h_var_pos;
    const int32_t *h_var_ref_len;
    const int64_t *h_alt_off;
    const uint8_t *h_alt;
    int32_t n_hap, words;
    const uint32_t *h_gt_bits;
    int32_t max_node_len;
};

// What one range of breakpoints produces, with chunk-local node and set numbering.
struct Chunk {
    std::vector<uint32_t> node_off, node_cons;
    std::vector<uint8_t> seq, flags;
    std::vector<int64_t> a0, clamp;
    std::vector<Edge> edges;
    SetTable sets;
};

// Breakpoints [i_lo, i_hi) of the graph.  A range may start at any breakpoint that no allele spans or ends at: there every
// haplotype arrives and the only thing that ends is the reference segment before it -- the canonical state this function
// starts from (`first` = the range begins at breakpoint 0, where nothing ends).  Ranges are therefore independent and run
// on several threads; build_host stitches them together in order.
static void build_range(const BuildInputs &in, const std::vector<int64_t> &bps, int64_t i_lo, int64_t i_hi, bool first, Chunk &ck)
{
    const int words = in.words;
    const int64_t nv = in.n_variants;
    const int64_t *h_var_pos = in.h_var_pos;
    const int32_t *h_var_ref_len = in.h_var_ref_len;
    const int64_t *h_alt_off = in.h_alt_off;
    const uint8_t *h_alt = in.h_alt, *h_ref = in.h_ref;
    const uint32_t *h_gt_bits = in.h_gt_bits;
    const int32_t max_node_len = in.max_node_len;
    SetTable &sets = ck.sets;
    sets.words = words;
    sets.on = h_gt_bits != nullptr && in.n_hap > 0;
    if (sets.on) {
        sets.full.assign((size_t)words, 0u);
        for (int h = 0; h < in.n_hap; ++h) sets.full[(size_t)(h >> 5)] |= 1u << (h & 31);
    }
    const Bits none;  // stands for "no set" when sets are off
    auto bp_index = [&](int64_t pos) { return (int64_t)(std::lower_bound(bps.begin(), bps.end(), pos) - bps.begin()); };

    std::vector<uint32_t> &node_off = ck.node_off, &node_cons = ck.node_cons;
    std::vector<uint8_t> &seq = ck.seq, &flags = ck.flags;
    std::vector<int64_t> &a0 = ck.a0, &clamp = ck.clamp;
    std::vector<Edge> &edges = ck.edges;
    node_off.assign(1, 0u);
    seq.reserve((size_t)(bps[(size_t)i_hi] - bps[(size_t)i_lo]) + 64);

    // appends the chain of nodes of one item -> (first, last) node index; the haplotype set of the item is filled in by
    // set_item_cons once it is known (node ids must follow input order, the sets are computed insertions first)
    struct Item { uint32_t first, last; size_t edge_lo, edge_hi; };
    auto add_chain = [&](const uint8_t *bases, int64_t len, int64_t start, int64_t clamp_at, bool isref) {
        Item it{0, 0, edges.size(), edges.size()};
        for (int64_t c0 = 0; c0 < len; c0 += max_node_len) {
            const int64_t n = std::min<int64_t>(max_node_len, len - c0);
            const uint32_t id = (uint32_t)a0.size();
            for (int64_t k = 0; k < n; ++k) seq.push_back(code_of(bases[c0 + k]));
            node_off.push_back((uint32_t)seq.size());
            a0.push_back(start + c0);
            clamp.push_back(isref ? start + c0 + n : clamp_at);
            flags.push_back(isref ? 1 : 0);
            node_cons.push_back(GB2_NO_CONS);
            if (c0 == 0) it.first = id; else edges.push_back(Edge{id - 1, id, GB2_NO_CONS});
            it.last = id;
        }
        it.edge_hi = edges.size();
        return it;
    };
    auto set_item_cons = [&](const Item &it, uint32_t cons) {
        for (uint32_t n = it.first; n <= it.last; ++n) node_cons[n] = cons;
        for (size_t e = it.edge_lo; e < it.edge_hi; ++e) edges[e].cons = cons;
    };

    std::map<int64_t, Bits> arrive;                    // breakpoint index -> haplotypes arriving there
    std::map<int64_t, std::vector<Source>> sources;    // breakpoint index -> what ends there
    if (sets.on) arrive[i_lo] = sets.full;
    if (!first) sources[i_lo].push_back(Source{GB2_PREV_NODE, sets.on ? sets.full : none});
    Bits A, left, took, tmp;
    if (sets.on) { A.resize((size_t)words); left.resize((size_t)words); took.resize((size_t)words); tmp.resize((size_t)words); }
    int64_t v_next = std::lower_bound(h_var_pos, h_var_pos + nv, bps[(size_t)i_lo]) - h_var_pos;

    for (int64_t i = i_lo; i < i_hi; ++i) {
        const int64_t b = bps[(size_t)i];
        const int64_t v_lo = v_next;
        while (v_next < nv && h_var_pos[v_next] == b) ++v_next;
        const int64_t v_hi = v_next;
        if (sets.on) {
            auto it = arrive.find(i);
            if (it != arrive.end()) { A = it->second; arrive.erase(it); } else std::fill(A.begin(), A.end(), 0u);
        }
        std::vector<Source> src;
        {
            auto it = sources.find(i);
            if (it != sources.end()) { src.swap(it->second); sources.erase(it); }
        }
        auto edge_from_sources = [&](uint32_t target, const Bits &sel) {
            for (const Source &s : src) {
                uint32_t c = GB2_NO_CONS;
                if (sets.on) { and_into(tmp, s.set, sel.data()); c = sets.id(tmp.data()); }
                edges.push_back(Edge{s.node, target, c});
            }
        };
        // ---- nodes of this breakpoint, in id order: alternative alleles (input order), then the reference segment
        std::vector<Item> alt_item((size_t)(v_hi - v_lo));
        bool any_ins = false;
        for (int64_t v = v_lo; v < v_hi; ++v) {
            const int64_t alen = h_alt_off[v + 1] - h_alt_off[v];
            any_ins |= h_var_ref_len[v] == 0;
            if (alen > 0) alt_item[(size_t)(v - v_lo)] = add_chain(h_alt + h_alt_off[v], alen, b, b + h_var_ref_len[v], false);
        }
        const Item ref_item = add_chain(h_ref + b, bps[(size_t)i + 1] - b, b, 0, true);
        // ---- insertions sit between what ends here and what starts here
        if (any_ins) {
            if (sets.on) left = A;
            std::vector<Source> added;
            for (int64_t v = v_lo; v < v_hi; ++v) {
                if (h_var_ref_len[v] != 0) continue;
                if (sets.on) {
                    and_into(took, left, h_gt_bits + (size_t)v * words);
                    for (int k = 0; k < words; ++k) left[(size_t)k] &= ~took[(size_t)k];
                }
                const Item &it = alt_item[(size_t)(v - v_lo)];
                set_item_cons(it, sets.on ? sets.id(took.data()) : GB2_NO_CONS);
                edge_from_sources(it.first, took);
                added.push_back(Source{it.last, sets.on ? took : none});
            }
            if (sets.on)
                for (Source &s : src)
                    for (int k = 0; k < words; ++k) s.set[(size_t)k] &= left[(size_t)k];
            for (Source &s : added) src.push_back(std::move(s));
        }
        // ---- replacements and deletions that start here (first carried one wins), then the reference segment
        if (sets.on) left = A;
        for (int64_t v = v_lo; v < v_hi; ++v) {
            const int64_t r = h_var_ref_len[v];
            if (r == 0) continue;
            if (sets.on) {
                and_into(took, left, h_gt_bits + (size_t)v * words);
                for (int k = 0; k < words; ++k) left[(size_t)k] &= ~took[(size_t)k];
            }
            const int64_t j = bp_index(b + r);
            if (sets.on) {
                auto it = arrive.find(j);
                if (it == arrive.end()) arrive[j] = took;
                else for (int k = 0; k < words; ++k) it->second[(size_t)k] |= took[(size_t)k];
            }
            if (h_alt_off[v + 1] - h_alt_off[v] > 0) {
                const Item &it = alt_item[(size_t)(v - v_lo)];
                set_item_cons(it, sets.on ? sets.id(took.data()) : GB2_NO_CONS);
                edge_from_sources(it.first, took);
                sources[j].push_back(Source{it.last, sets.on ? took : none});
            } else {  // deletion: whatever ended here now ends at its far side
                std::vector<Source> &dst = sources[j];
                for (const Source &s : src) {
                    Source t{s.node, none};
                    if (sets.on) { t.set.resize((size_t)words); and_into(t.set, s.set, took.data()); }
                    dst.push_back(std::move(t));
                }
            }
        }
        set_item_cons(ref_item, sets.on ? sets.id(left.data()) : GB2_NO_CONS);
        edge_from_sources(ref_item.first, left);
        sources[i + 1].push_back(Source{ref_item.last, sets.on ? left : none});
        if (sets.on) {
            auto it = arrive.find(i + 1);
            if (it == arrive.end()) arrive[i + 1] = left;
            else for (int k = 0; k < words; ++k) it->second[(size_t)k] |= left[(size_t)k];
        }
    }
    // what is left in `arrive` / `sources` belongs to breakpoint i_hi: the canonical state the next range starts from
}

// The host pass: pure CPU work on caller-owned arrays, no CUDA call.  n_threads > 1 cuts the breakpoints of ONE chromosome
// into independent ranges (see build_range) that run on worker threads; the result does not depend on the cut points --
// node ids, edge order and the numbering of the haplotype sets are those of the single-range pass.  chunk_bps > 0 forces
// the range size (tests).
static int build_host(ErrSink *ctx, const uint8_t *h_ref, int64_t ref_len, int64_t n_variants, const int64_t *h_var_pos,
                      const int32_t *h_var_ref_len, const int64_t *h_alt_off, const uint8_t *h_alt, int32_t n_hap,
                      int32_t words, const uint32_t *h_gt_bits, int32_t max_node_len, HostGraph &hg, int n_threads = 1,
                      int64_t chunk_bps = 0)
{
    GB2_REQUIRE(ctx, h_ref && ref_len >= 1, "gb2_graph_build: empty reference");
    GB2_REQUIRE(ctx, n_variants >= 0 && (n_variants == 0 || (h_var_pos && h_var_ref_len && h_alt_off && h_alt)),
                "gb2_graph_build: null variant array");
    GB2_REQUIRE(ctx, max_node_len >= 1, "gb2_graph_build: max_node_len must be positive");
    GB2_REQUIRE(ctx, n_hap >= 0 && words >= 4 && (words & 3) == 0 && (int64_t)words * 32 >= n_hap,
                "gb2_graph_build: haplotype bit sets need a multiple of 4 words covering %d haplotypes", n_hap);
    const int64_t L = ref_len, nv = n_variants;
    for (int64_t v = 0; v < nv; ++v) {
        const int64_t s = h_var_pos[v], r = h_var_ref_len[v], a = h_alt_off[v + 1] - h_alt_off[v];
        GB2_REQUIRE(ctx, s >= 0 && r >= 0 && a >= 0 && s + r <= L, "gb2_graph_build: variant %lld outside the reference", (long long)v);
        GB2_REQUIRE(ctx, r > 0 || a > 0, "gb2_graph_build: variant %lld has two empty alleles", (long long)v);
        GB2_REQUIRE(ctx, v == 0 || h_var_pos[v - 1] <= s, "gb2_graph_build: variants must be sorted by position");
    }
    const BuildInputs in{h_ref, ref_len, n_variants, h_var_pos, h_var_ref_len, h_alt_off, h_alt, n_hap, words, h_gt_bits, max_node_len};

    // breakpoints
    std::vector<int64_t> bps;
    bps.reserve((size_t)(2 * nv + 2));
    bps.push_back(0);
    bps.push_back(L);
    for (int64_t v = 0; v < nv; ++v) {
        bps.push_back(h_var_pos[v]);
        bps.push_back(h_var_pos[v] + h_var_ref_len[v]);
    }
    std::sort(bps.begin(), bps.end());
    bps.erase(std::unique(bps.begin(), bps.end()), bps.end());
    const int64_t nb = (int64_t)bps.size() - 1;

    // ---- ranges: cut only at breakpoints that no allele spans or ends at (the farthest end of the alleles that start
    //      before the breakpoint lies before it)
    std::vector<int64_t> cuts(1, 0);
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(n_threads, nb));
    int64_t want = chunk_bps > 0 ? chunk_bps : (nt > 1 ? std::max<int64_t>(4096, nb / (4 * (int64_t)nt)) : nb + 1);
    if (want <= nb) {
        int64_t reach = -1, v = 0;
        for (int64_t i = 1; i < nb; ++i) {
            const int64_t b = bps[(size_t)i];
            while (v < nv && h_var_pos[v] < b) { reach = std::max(reach, h_var_pos[v] + h_var_ref_len[v]); ++v; }
            if (reach < b && i - cuts.back() >= want) cuts.push_back(i);
        }
    }
    cuts.push_back(nb);
    const int64_t n_chunks = (int64_t)cuts.size() - 1;
    hg.n_ranges = n_chunks;
    const bool timing = getenv("GB2_BUILD_TIMING") != nullptr;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_ranges = now();
    std::vector<Chunk> chunks((size_t)n_chunks);
    if (n_chunks == 1 || nt == 1) {
        for (int64_t k = 0; k < n_chunks; ++k) build_range(in, bps, cuts[(size_t)k], cuts[(size_t)k + 1], k == 0, chunks[(size_t)k]);
    } else {
        std::atomic<int64_t> next(0);
        auto worker = [&]() {
            for (;;) {
                const int64_t k = next.fetch_add(1);
                if (k >= n_chunks) return;
                build_range(in, bps, cuts[(size_t)k], cuts[(size_t)k + 1], k == 0, chunks[(size_t)k]);
            }
        };
        std::vector<std::thread> pool;
        for (int t = 0; t < std::min<int64_t>(nt, n_chunks); ++t) pool.emplace_back(worker);
        for (auto &t : pool) t.join();
    }

    const double t_stitch = now();
    // ---- stitch the ranges together in order: node ids and base offsets shift, set ids go through the global table in
    //      order of first appearance (what one pass over all breakpoints would have assigned)
    SetTable sets;
    sets.words = words;
    sets.on = h_gt_bits != nullptr && n_hap > 0;
    if (sets.on) sets.full = chunks[0].sets.full;
    std::vector<uint32_t> &node_off = hg.node_off, &node_cons = hg.node_cons;
    std::vector<uint8_t> &seq = hg.seq, &flags = hg.flags;
    std::vector<int64_t> &a0 = hg.a0, &clamp = hg.clamp;
    std::vector<Edge> edges;
    if (n_chunks == 1) {  // nothing to renumber
        Chunk &c = chunks[0];
        node_off.swap(c.node_off); node_cons.swap(c.node_cons); seq.swap(c.seq); flags.swap(c.flags);
        a0.swap(c.a0); clamp.swap(c.clamp); edges.swap(c.edges);
        sets.flat.swap(c.sets.flat); sets.index.swap(c.sets.index); sets.hashes.swap(c.sets.hashes);
    } else {
        // phase A (one thread, cheap): where every range goes, and the global id of every set -- looked up by the hash the
        // range already computed; the rows themselves stay where they are until phase B
        std::vector<size_t> node_base((size_t)n_chunks + 1, 0), seq_base((size_t)n_chunks + 1, 0), edge_base((size_t)n_chunks + 1, 0);
        for (int64_t k = 0; k < n_chunks; ++k) {
            node_base[(size_t)k + 1] = node_base[(size_t)k] + chunks[(size_t)k].a0.size();
            seq_base[(size_t)k + 1] = seq_base[(size_t)k] + chunks[(size_t)k].seq.size();
            edge_base[(size_t)k + 1] = edge_base[(size_t)k] + chunks[(size_t)k].edges.size();
        }
        std::vector<std::vector<uint32_t>> remap((size_t)n_chunks);
        std::vector<const uint32_t *> row_ptr;  // global id -> the row, inside the range that saw it first
        if (sets.on) {
            // Global ids in order of first appearance over (range, local id), without a hash-table insert per row: sort
            // (hash, order), split every run of equal hashes into classes of equal rows (almost always one), the first
            // member of a class represents it, ids count the representatives in order.
            struct Ent { uint64_t hash; uint32_t ord; };
            std::vector<size_t> set_base((size_t)n_chunks + 1, 0);
            for (int64_t k = 0; k < n_chunks; ++k) set_base[(size_t)k + 1] = set_base[(size_t)k] + chunks[(size_t)k].sets.hashes.size();
            const size_t tot_sets = set_base[(size_t)n_chunks];
            std::vector<Ent> ents(tot_sets);
            std::vector<const uint32_t *> local_ptr(tot_sets);
            for (int64_t k = 0; k < n_chunks; ++k) {
                const SetTable &ls = chunks[(size_t)k].sets;
                for (size_t r = 0; r < ls.hashes.size(); ++r) {
                    const size_t o = set_base[(size_t)k] + r;
                    ents[o] = Ent{ls.hashes[r], (uint32_t)o};
                    local_ptr[o] = ls.row((uint32_t)r);
                }
            }
            std::sort(ents.begin(), ents.end(), [](const Ent &a, const Ent &b) { return a.hash != b.hash ? a.hash < b.hash : a.ord < b.ord; });
            std::vector<uint32_t> rep(tot_sets);  // order index of the representative of every row
            for (size_t lo = 0; lo < tot_sets;) {
                size_t hi = lo + 1;
                while (hi < tot_sets && ents[hi].hash == ents[lo].hash) ++hi;
                for (size_t a = lo; a < hi; ++a) {  // runs are tiny: first earlier member with the same bytes, else itself
                    uint32_t r0 = ents[a].ord;
                    for (size_t b = lo; b < a; ++b)
                        if (rep[ents[b].ord] == ents[b].ord &&
                            memcmp(local_ptr[ents[b].ord], local_ptr[ents[a].ord], (size_t)words * 4) == 0) { r0 = ents[b].ord; break; }
                    rep[ents[a].ord] = r0;
                }
                lo = hi;
            }
            std::vector<uint32_t> gid_of(tot_sets);
            for (size_t o = 0; o < tot_sets; ++o) {
                if (rep[o] == (uint32_t)o) {
                    gid_of[o] = (uint32_t)row_ptr.size();
                    row_ptr.push_back(local_ptr[o]);
                } else {
                    gid_of[o] = gid_of[rep[o]];  // the representative comes earlier in order
                }
            }
            sets.hashes.resize(row_ptr.size());
            for (int64_t k = 0; k < n_chunks; ++k) {
                std::vector<uint32_t> &rm = remap[(size_t)k];
                const SetTable &ls = chunks[(size_t)k].sets;
                rm.resize(ls.hashes.size());
                for (size_t r = 0; r < rm.size(); ++r) {
                    rm[r] = gid_of[set_base[(size_t)k] + r];
                    sets.hashes[rm[r]] = ls.hashes[r];
                }
            }
            sets.index_stale = true;  // rebuilt from `hashes` only if the CSR step has parallel edges to merge
            sets.flat.resize(row_ptr.size() * (size_t)words);
        }
        const size_t tot_nodes = node_base[(size_t)n_chunks], tot_seq = seq_base[(size_t)n_chunks], tot_edges = edge_base[(size_t)n_chunks];
        node_off.resize(tot_nodes + 1); node_cons.resize(tot_nodes); flags.resize(tot_nodes); a0.resize(tot_nodes);
        clamp.resize(tot_nodes); seq.resize(tot_seq); edges.resize(tot_edges);
        node_off[0] = 0u;
        // phase B (worker threads): every range copies itself -- and the set rows it saw first -- into place
        auto place = [&](int64_t k) {
            Chunk &c = chunks[(size_t)k];
            const uint32_t nb0 = (uint32_t)node_base[(size_t)k], sb0 = (uint32_t)seq_base[(size_t)k];
            const std::vector<uint32_t> &rm = remap[(size_t)k];
            auto gid = [&](uint32_t cons) { return cons == GB2_NO_CONS ? GB2_NO_CONS : rm[cons]; };
            if (!c.seq.empty()) memcpy(seq.data() + sb0, c.seq.data(), c.seq.size());
            for (size_t n = 1; n < c.node_off.size(); ++n) node_off[nb0 + n] = c.node_off[n] + sb0;
            for (size_t n = 0; n < c.node_cons.size(); ++n) node_cons[nb0 + n] = gid(c.node_cons[n]);
            if (!c.flags.empty()) memcpy(flags.data() + nb0, c.flags.data(), c.flags.size());
            if (!c.a0.empty()) memcpy(a0.data() + nb0, c.a0.data(), c.a0.size() * sizeof(int64_t));
            if (!c.clamp.empty()) memcpy(clamp.data() + nb0, c.clamp.data(), c.clamp.size() * sizeof(int64_t));
            Edge *eo = edges.data() + edge_base[(size_t)k];
            for (size_t e = 0; e < c.edges.size(); ++e) {
                const Edge &x = c.edges[e];
                eo[e] = Edge{x.src == GB2_PREV_NODE ? nb0 - 1u : x.src + nb0, x.dst + nb0, gid(x.cons)};
            }
            for (size_t r = 0; r < rm.size(); ++r)
                if (row_ptr[rm[r]] == c.sets.row((uint32_t)r))  // this range saw the set first: it owns the copy
                    memcpy(sets.flat.data() + (size_t)rm[r] * words, c.sets.row((uint32_t)r), (size_t)words * 4);
        };
        if (nt == 1) {
            for (int64_t k = 0; k < n_chunks; ++k) place(k);
        } else {
            std::atomic<int64_t> next(0);
            auto worker = [&]() {
                for (;;) {
                    const int64_t k = next.fetch_add(1);
                    if (k >= n_chunks) return;
                    place(k);
                }
            };
            std::vector<std::thread> pool;
            for (int t = 0; t < std::min<int64_t>(nt, n_chunks); ++t) pool.emplace_back(worker);
            for (auto &t : pool) t.join();
        }
        chunks.clear();
    }
    if (seq.size() >= ((size_t)1 << 32) || a0.size() >= ((size_t)1 << 31)) {
        GB2_SET_ERR(ctx, "gb2_graph_build: graph too large for 32-bit base offsets");
        return GB2_ERR_ARG;
    }

    const double t_csr = now();
    // ---- CSR: by (source, target); repeated structural edges (two deletions with the same ends) are merged
    const int64_t n_nodes = (int64_t)a0.size();
    std::sort(edges.begin(), edges.end(), [](const Edge &x, const Edge &y) { return x.src != y.src ? x.src < y.src : x.dst < y.dst; });
    std::vector<uint32_t> &edge_off = hg.edge_off, &edge_to = hg.edge_to, &edge_cons = hg.edge_cons;
    edge_off.assign((size_t)n_nodes + 1, 0u);
    edge_to.reserve(edges.size());
    edge_cons.reserve(edges.size());
    for (size_t e = 0; e < edges.size(); ++e) {
        if (!edge_to.empty() && e > 0 && edges[e].src == edges[e - 1].src && edges[e].dst == edges[e - 1].dst) {
            uint32_t &c = edge_cons.back();
            if (c != GB2_NO_CONS && edges[e].cons != GB2_NO_CONS) {
                Bits u(sets.row(c), sets.row(c) + words);
                const uint32_t *o = sets.row(edges[e].cons);
                for (int k = 0; k < words; ++k) u[(size_t)k] |= o[k];
                c = sets.id(u.data());
            } else {
                c = GB2_NO_CONS;
            }
            continue;
        }
        edge_to.push_back(edges[e].dst);
        edge_cons.push_back(edges[e].cons);
        edge_off[(size_t)edges[e].src + 1]++;
    }
    for (int64_t n = 0; n < n_nodes; ++n) edge_off[(size_t)n + 1] += edge_off[(size_t)n];
    std::vector<Edge>().swap(edges);
    hg.n_cons = sets.on ? (int64_t)(sets.flat.size() / (size_t)words) : 0;
    hg.n_hap = sets.on ? n_hap : 0;
    hg.words = words;
    hg.cons_bits.swap(sets.flat);
    if (timing)
        fprintf(stderr, "gb2_graph_build host pass: %lld ranges on %d threads %.3f s, stitch %.3f s, CSR %.3f s\n", (long long)n_chunks, nt,
                t_stitch - t_ranges, t_csr - t_stitch, now() - t_csr);
    return GB2_OK;
}

static int build_threads_default()
{
    const char *e = getenv("GB2_BUILD_THREADS");
    const int n = e ? atoi(e) : (int)std::thread::hardware_concurrency();
    return std::max(1, n);
}

static int upload_host_graph(gb2_ctx *ctx, const HostGraph &hg, gb2_graph **out)
{
    static const uint32_t dummy[4] = {0, 0, 0, 0};
    return gb2_graph_create(ctx, (int64_t)hg.a0.size(), hg.node_off.data(), hg.seq.data(), hg.a0.data(), hg.clamp.data(),
                            hg.flags.data(), hg.node_cons.data(), (int64_t)hg.edge_to.size(), hg.edge_off.data(),
                            hg.edge_to.data(), hg.edge_cons.data(), hg.n_hap, hg.words, hg.n_cons,
                            hg.n_cons ? hg.cons_bits.data() : dummy, out);
}

extern "C" int gb2_graph_build(gb2_ctx *ctx, const uint8_t *h_ref, int64_t ref_len, int64_t n_variants,
                               const int64_t *h_var_pos, const int32_t *h_var_ref_len, const int64_t *h_alt_off,
                               const uint8_t *h_alt, int32_t n_hap, int32_t words, const uint32_t *h_gt_bits,
                               int32_t max_node_len, gb2_graph **out)
{
    if (!ctx || !out) return GB2_ERR_ARG;
    *out = nullptr;
    HostGraph hg;
    ErrSink es;
    const int rc = build_host(&es, h_ref, ref_len, n_variants, h_var_pos, h_var_ref_len, h_alt_off, h_alt, n_hap, words,
                              h_gt_bits, max_node_len, hg, build_threads_default());
    if (rc != GB2_OK) {
        GB2_SET_ERR(ctx, "%s", es.err);
        return rc;
    }
    return upload_host_graph(ctx, hg, out);
}

// The host pass alone (no GPU, no context): sizes of the graph and a 64-bit digest of every array it would upload.  The
// digest does not depend on n_threads / chunk_bps (how many worker threads cut the breakpoints into how large ranges):
// tests/test_graph_cpu.py checks exactly that.  h_stats[6] = nodes, edges, bases, haplotype-set rows, ranges used, digest.
extern "C" int gb2_graph_build_stats(const uint8_t *h_ref, int64_t ref_len, int64_t n_variants, const int64_t *h_var_pos,
                                     const int32_t *h_var_ref_len, const int64_t *h_alt_off, const uint8_t *h_alt,
                                     int32_t n_hap, int32_t words, const uint32_t *h_gt_bits, int32_t max_node_len,
                                     int32_t n_threads, int64_t chunk_bps, uint64_t *h_stats)
{
    if (!h_stats) return GB2_ERR_ARG;
    HostGraph hg;
    ErrSink es;
    const int rc = build_host(&es, h_ref, ref_len, n_variants, h_var_pos, h_var_ref_len, h_alt_off, h_alt, n_hap, words,
                              h_gt_bits, max_node_len, hg, n_threads > 0 ? n_threads : build_threads_default(), chunk_bps);
    if (rc != GB2_OK) return rc;
    uint64_t h = 1469598103934665603ull;
    auto mix = [&](const void *p, size_t bytes) {
        const uint8_t *b = (const uint8_t *)p;
        for (size_t i = 0; i < bytes; ++i) { h ^= b[i]; h *= 1099511628211ull; }
        h ^= bytes; h *= 1099511628211ull;
    };
    mix(hg.node_off.data(), hg.node_off.size() * 4); mix(hg.seq.data(), hg.seq.size());
    mix(hg.a0.data(), hg.a0.size() * 8); mix(hg.clamp.data(), hg.clamp.size() * 8);
    mix(hg.flags.data(), hg.flags.size()); mix(hg.node_cons.data(), hg.node_cons.size() * 4);
    mix(hg.edge_off.data(), hg.edge_off.size() * 4); mix(hg.edge_to.data(), hg.edge_to.size() * 4);
    mix(hg.edge_cons.data(), hg.edge_cons.size() * 4); mix(hg.cons_bits.data(), hg.cons_bits.size() * 4);
    h_stats[0] = hg.a0.size(); h_stats[1] = hg.edge_to.size(); h_stats[2] = hg.seq.size(); h_stats[3] = (uint64_t)hg.n_cons;
    h_stats[4] = (uint64_t)hg.n_ranges; h_stats[5] = h;
    return GB2_OK;
}

// Several chromosomes at once: the host passes run on up to n_threads worker threads (0 = one per hardware thread), each
// finished graph is uploaded by the calling thread on the context's stream while the others are still being built.
extern "C" int gb2_graph_build_batch(gb2_ctx *ctx, int32_t n_graphs, const gb2_graph_input *inputs, int32_t n_threads,
                                     gb2_graph **out)
{
    if (!ctx || !out || n_graphs < 0 || (n_graphs > 0 && !inputs)) return GB2_ERR_ARG;
    for (int i = 0; i < n_graphs; ++i) out[i] = nullptr;
    if (n_graphs == 0) return GB2_OK;
    int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    const int budget = std::max(1, nt);
    nt = std::max(1, std::min(nt, (int)n_graphs));
    const int inner_threads = std::max(1, budget / nt);  // fewer chromosomes than threads: the rest cut each chromosome into ranges
    std::vector<HostGraph> hgs((size_t)n_graphs);
    std::vector<ErrSink> errs((size_t)n_graphs);
    std::vector<int> rcs((size_t)n_graphs, GB2_OK);
    std::vector<std::atomic<int>> ready((size_t)n_graphs);
    for (auto &r : ready) r.store(0);
    // largest inputs first, so that the longest chromosome does not start last
    std::vector<int> order((size_t)n_graphs);
    for (int i = 0; i < n_graphs; ++i) order[(size_t)i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        return inputs[a].ref_len + 64 * inputs[a].n_variants > inputs[b].ref_len + 64 * inputs[b].n_variants;
    });
    std::atomic<int> next(0), uploaded(0);
    auto worker = [&]() {
        for (;;) {
            const int k = next.fetch_add(1);
            if (k >= n_graphs) return;
            // bound the host copies alive at once: do not run more than nt graphs ahead of the uploader
            while (k > uploaded.load(std::memory_order_acquire) + nt) std::this_thread::yield();
            const int i = order[(size_t)k];
            const gb2_graph_input &in = inputs[i];
            rcs[(size_t)i] = build_host(&errs[(size_t)i], in.h_ref, in.ref_len, in.n_variants, in.h_var_pos, in.h_var_ref_len,
                                        in.h_alt_off, in.h_alt, in.n_hap, in.words, in.h_gt_bits, in.max_node_len, hgs[(size_t)i],
                                        inner_threads);
            ready[(size_t)i].store(1, std::memory_order_release);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 0; t < nt; ++t) pool.emplace_back(worker);
    int rc = GB2_OK;
    for (int k = 0; k < n_graphs; ++k) {  // upload in the order the workers take them
        const int i = order[(size_t)k];
        while (!ready[(size_t)i].load(std::memory_order_acquire)) std::this_thread::yield();
        if (rc == GB2_OK && rcs[(size_t)i] != GB2_OK) {
            rc = rcs[(size_t)i];
            GB2_SET_ERR(ctx, "graph %d: %s", i, errs[(size_t)i].err);
        } else if (rc == GB2_OK) {
            rc = upload_host_graph(ctx, hgs[(size_t)i], &out[i]);
        }
        hgs[(size_t)i] = HostGraph();  // free the host copy
        uploaded.store(k + 1, std::memory_order_release);  // after a failure the rest is still waited for, then dropped
    }
    for (auto &t : pool) t.join();
    if (rc != GB2_OK)
        for (int i = 0; i < n_graphs; ++i)
            if (out[i]) { gb2_graph_destroy(out[i]); out[i] = nullptr; }
    return rc;
}
