// seqscan.cu -- K2 over SEQUENCES: every motif-width window of a batch of 2-bit packed sequences is formed in
// registers (funnel shifts over two adjacent 64-bit words) and scored on both strands; nothing per window is read
// from or written to HBM.  Plus the sequence encoder (ASCII -> 2-bit words + N bits).
//
// Replaces the row loop of score_seqs (src/grafimo/score_sequences.py:273-321) + compute_score_seq (:331-396) for the
// workload the headline metric is quoted on -- every window of every haplotype -- where the reference's input (one text
// row per window) repeats every base w times: a window costs 0.25 B of HBM (2-bit) or 1 B of PCIe (ASCII) here
// instead of 8 B (packed k-mer) / 19 B (ASCII k-mer).
//
// Layout: sequence s occupies words [word_off[s], word_off[s] + ceil(len[s] / 32)) of d_seq2; base i sits in bits
// [2(i & 31), 2(i & 31) + 1] of word word_off[s] + (i >> 5) (A=0 C=1 G=2 T=3, the k-mer layout of grafimo_b200.h, so a
// window is a 2w-bit slice of two adjacent words); d_nbits holds one uint32 per word: bit (i & 31) = base i is not
// A/C/G/T (a window touching such a base is scored as `min_val`, p = 1: score_sequences.py:376-378).
//
// Work decomposition: a UNIT is 32 consecutive words (1024 window start positions) of one sequence; every warp owns a
// contiguous range of units, lane l of the warp owns word l of the unit = 32 windows, scored 8 at a time exactly like
// K2 scores the 8 k-mers of its four 128-bit loads (same replicated chunk LUT, same packed-u16 accumulators, same
// shared-memory histogram and warp-aggregated hit append).  The only additions per window are the two funnel shifts.
#include <stdlib.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>

#include "score_common.cuh"

struct SeqDesc {
    int64_t word_off;  // first word of the sequence in d_seq2 / d_nbits
    int64_t len;       // bases
    int64_t row0;      // global index of the sequence's first window
    int64_t unit0;     // units before this sequence (desc[n_seqs].unit0 = total)
};

struct SeqScoreParams {
    ScoreParams sp;  // packed / nmask / n unused
    const uint64_t *seq2;
    const uint32_t *nbits;
    const SeqDesc *desc;
    int64_t n_seqs;
    int64_t total_units;
    int w;
};

// largest s with desc[s].unit0 <= u and a non-empty unit range (desc[n_seqs].unit0 = total_units > u)
__device__ __forceinline__ int64_t find_seq(const SeqDesc *desc, int64_t n_seqs, int64_t u)
{
    int64_t lo = 0, hi = n_seqs;  // invariant: desc[lo].unit0 <= u < desc[hi].unit0
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(&desc[mid].unit0) <= u) lo = mid;
        else hi = mid;
    }
    return lo;
}

// window `POS` (0..31) of the 64 bases held in (L0,L1,H0,H1): its low 32 bases as two 32-bit words
template <int POS>
__device__ __forceinline__ void window_words(uint32_t L0, uint32_t L1, uint32_t H0, uint32_t H1, uint32_t &w0, uint32_t &w1)
{
    if (POS == 0) {
        w0 = L0; w1 = L1;
    } else if (POS < 16) {
        w0 = __funnelshift_r(L0, L1, 2 * POS);
        w1 = __funnelshift_r(L1, H0, 2 * POS);
    } else if (POS == 16) {
        w0 = L1; w1 = H0;
    } else {
        w0 = __funnelshift_r(L1, H0, 2 * POS - 32);
        w1 = __funnelshift_r(H0, H1, 2 * POS - 32);
    }
}

template <int CB, int NCHUNK, int R, int ROUND, int K>
struct RoundScore {
    static __device__ __forceinline__ void run(uint32_t L0, uint32_t L1, uint32_t H0, uint32_t H1, uint32_t lut32, uint32_t *acc)
    {
        uint32_t w0, w1;
        window_words<ROUND * 8 + K>(L0, L1, H0, H1, w0, w1);
        acc[K] = score_word<CB, NCHUNK, R>(w0, w1, lut32);
        RoundScore<CB, NCHUNK, R, ROUND, K + 1>::run(L0, L1, H0, H1, lut32, acc);
    }
};
template <int CB, int NCHUNK, int R, int ROUND>
struct RoundScore<CB, NCHUNK, R, ROUND, 8> {
    static __device__ __forceinline__ void run(uint32_t, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t *) {}
};

// Rare path, entered by the whole warp and kept out of line so that it does not cost the scoring loop registers:
// exact per-bin test + warp-aggregated append for the 8 windows of a round.
__device__ __noinline__ void emit_round_hits(const ScoreParams &p, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                             uint32_t a4, uint32_t a5, uint32_t a6, uint32_t a7, uint32_t okmask,
                                             int64_t row_first, bool two, unsigned lane)
{
    const uint32_t acc[8] = {a0, a1, a2, a3, a4, a5, a6, a7};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint32_t bf = acc[k] & 0xFFFFu, br = acc[k] >> 16;
        const uint64_t row = (uint64_t)(row_first + k);
        const bool ok = (okmask >> k) & 1u;
        append_hits(p, ok && bin_hits(p, bf), row, bf, 0u, lane);
        if (two) append_hits(p, ok && bin_hits(p, br), row, br, 1u, lane);
    }
}

// One round = windows 8*ROUND .. 8*ROUND+7 of this lane's word.  GUARD: windows may be invalid (sequence tail) or
// touch an N base; without it the round is straight-line code with no per-window predicate.
template <int CB, int NCHUNK, int R, int ROUND, bool GUARD>
__device__ __forceinline__ void score_round(const ScoreParams &p, uint32_t L0, uint32_t L1, uint32_t H0, uint32_t H1,
                                            uint64_t nn, uint64_t wmask, int nvalid, int64_t row_first, uint32_t lut32,
                                            uint32_t hist32, unsigned lane, bool do_hist, bool two, uint32_t nsent,
                                            uint32_t cut_hi)
{
    uint32_t acc[8];
    RoundScore<CB, NCHUNK, R, ROUND, 0>::run(L0, L1, H0, H1, lut32, acc);
    bool ok[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        ok[k] = true;
        if (GUARD) {
            const int pos = ROUND * 8 + k;
            ok[k] = pos < nvalid;
            if ((nn >> pos) & wmask) acc[k] = nsent;
        }
    }
    if (do_hist) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (!GUARD || ok[k]) {
                red_shared_inc(hist32 + 4u * __byte_perm(acc[k], 0u, 0x4410u));
                if (two) red_shared_inc(hist32 + 4u * (acc[k] >> 16));
            }
        }
    }
    if (p.dense != nullptr) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (!GUARD || ok[k]) p.dense[row_first + ROUND * 8 + k] = (acc[k] == nsent) ? 0xFFFFFFFFu : acc[k];
    }
    if (p.hits != nullptr) {
        uint32_t mx = 0;
        if (GUARD) {
#pragma unroll
            for (int k = 0; k < 8; ++k) mx = __vimax3_u16x2(mx, (ok[k] && acc[k] != nsent) ? acc[k] : 0u, 0u);
        } else {
            mx = __vimax3_u16x2(acc[0], acc[1], acc[2]);
            mx = __vimax3_u16x2(mx, acc[3], acc[4]);
            mx = __vimax3_u16x2(mx, acc[5], acc[6]);
            mx = __vimax3_u16x2(mx, acc[7], acc[7]);
        }
        const bool any = ((mx & 0xFFFFu) >= p.cut) | (two & (mx >= cut_hi));
        if (__any_sync(0xFFFFFFFFu, any)) {
            uint32_t okmask = 0xFFu;
            if (GUARD) {
                okmask = 0u;
#pragma unroll
                for (int k = 0; k < 8; ++k) okmask |= ok[k] ? (1u << k) : 0u;
            }
            emit_round_hits(p, acc[0], acc[1], acc[2], acc[3], acc[4], acc[5], acc[6], acc[7], okmask, row_first + ROUND * 8, two, lane);
        }
    }
}

template <int CB, int NCHUNK, int R, bool GUARD>
__device__ __forceinline__ void score_unit_body(const ScoreParams &p, uint64_t lo, uint64_t hi, uint64_t nn, uint64_t wmask,
                                           int nvalid, int64_t row_first, uint32_t lut32, uint32_t hist32, unsigned lane,
                                           bool do_hist, bool two, uint32_t nsent, uint32_t cut_hi)
{
    const uint32_t L0 = (uint32_t)lo, L1 = (uint32_t)(lo >> 32), H0 = (uint32_t)hi, H1 = (uint32_t)(hi >> 32);
    score_round<CB, NCHUNK, R, 0, GUARD>(p, L0, L1, H0, H1, nn, wmask, nvalid, row_first, lut32, hist32, lane, do_hist, two, nsent, cut_hi);
    score_round<CB, NCHUNK, R, 1, GUARD>(p, L0, L1, H0, H1, nn, wmask, nvalid, row_first, lut32, hist32, lane, do_hist, two, nsent, cut_hi);
    score_round<CB, NCHUNK, R, 2, GUARD>(p, L0, L1, H0, H1, nn, wmask, nvalid, row_first, lut32, hist32, lane, do_hist, two, nsent, cut_hi);
    score_round<CB, NCHUNK, R, 3, GUARD>(p, L0, L1, H0, H1, nn, wmask, nvalid, row_first, lut32, hist32, lane, do_hist, two, nsent, cut_hi);
}

// the guarded form (sequence tails, units with N bases) is rare: out of line, its registers are its own
template <int CB, int NCHUNK, int R>
__device__ __noinline__ void score_unit_guarded(const ScoreParams &p, uint64_t lo, uint64_t hi, uint64_t nn, uint64_t wmask,
                                                int nvalid, int64_t row_first, uint32_t lut32, uint32_t hist32, unsigned lane,
                                                bool do_hist, bool two, uint32_t nsent, uint32_t cut_hi)
{
    score_unit_body<CB, NCHUNK, R, true>(p, lo, hi, nn, wmask, nvalid, row_first, lut32, hist32, lane, do_hist, two, nsent, cut_hi);
}

template <int CB, int NCHUNK, int R, int NT>
__global__ void __launch_bounds__(NT, 1) gb2_score_seq_kernel(const SeqScoreParams q)
{
    extern __shared__ __align__(16) uint32_t smem[];
    const ScoreParams &p = q.sp;
    constexpr int LUT_WORDS = NCHUNK * ChunkGeom<CB>::ENTRIES * R;
    uint32_t *lut_s = smem;               // [NCHUNK * 4^CB][R]
    uint32_t *hist_s = smem + LUT_WORDS;  // [span+1]
    const unsigned tid = threadIdx.x, lane = tid & 31u;
    const bool do_hist = p.hist != nullptr;

    for (int i = tid; i < LUT_WORDS; i += NT) lut_s[i] = p.lut[i / R];
    if (do_hist)
        for (uint32_t i = tid; i <= p.span; i += NT) hist_s[i] = 0u;
    __syncthreads();

    const uint32_t lut32 = smem_u32(lut_s) + 4u * (lane & (R - 1));
    const uint32_t hist32 = smem_u32(hist_s);
    const uint32_t nsent = (p.span << 16) | p.span;
    const uint32_t cut_hi = p.cut << 16;
    const bool two = p.two_strands != 0;
    const bool has_n = q.nbits != nullptr;
    const int w = q.w;
    const uint64_t wmask = (w >= 64) ? ~0ull : ((1ull << w) - 1ull);
    const uint32_t nhi_mask = (w >= 33) ? 0xFFFFFFFFu : (uint32_t)((1ull << (w - 1)) - 1ull);  // bases of the next word a window can touch

    // contiguous unit range of this warp
    const int64_t nwarps = (int64_t)gridDim.x * (NT / 32);
    const int64_t gw = (int64_t)blockIdx.x * (NT / 32) + (tid >> 5);
    const int64_t u_beg = q.total_units / nwarps * gw + min(gw, q.total_units % nwarps);
    const int64_t u_end = u_beg + q.total_units / nwarps + (gw < q.total_units % nwarps ? 1 : 0);

    // No software prefetch: a unit is ~14,000 cycles of shared-memory-pipe time per warp (the pipe is shared by all
    // the warps of the SM), so the ~1 us of its two loads is hidden by the other warps.  The descriptor of the current
    // sequence is re-read (L1) per unit instead of being carried in registers.
    if (u_beg < u_end) {
        int s = (int)find_seq(q.desc, q.n_seqs, u_beg);
        for (int64_t u = u_beg; u < u_end; ++u) {
            while (u >= __ldg(&q.desc[s + 1].unit0)) ++s;  // next sequence (empty ones are skipped)
            const int64_t word_off = __ldg(&q.desc[s].word_off), len = __ldg(&q.desc[s].len);
            const int64_t nwords = (len + 31) >> 5;
            const int64_t wi = (u - __ldg(&q.desc[s].unit0)) * 32 + lane;
            const uint64_t lo = wi < nwords ? __ldg(q.seq2 + word_off + wi) : 0ull;
            uint64_t hi = __shfl_down_sync(0xFFFFFFFFu, lo, 1);
            if (lane == 31) hi = wi + 1 < nwords ? __ldg(q.seq2 + word_off + wi + 1) : 0ull;
            uint64_t nn = 0ull;
            if (has_n) {
                const uint32_t nlo = wi < nwords ? __ldg(q.nbits + word_off + wi) : 0u;
                uint32_t nhi = __shfl_down_sync(0xFFFFFFFFu, nlo, 1);
                if (lane == 31) nhi = wi + 1 < nwords ? __ldg(q.nbits + word_off + wi + 1) : 0u;
                nn = (uint64_t)nlo | ((uint64_t)(nhi & nhi_mask) << 32);
            }
            const int64_t left = len - w + 1 - wi * 32;
            const int nvalid = (int)max((int64_t)0, min((int64_t)32, left));
            const int64_t row_first = __ldg(&q.desc[s].row0) + wi * 32;
            const bool plain = (nvalid == 32) & (nn == 0ull);
            if (__all_sync(0xFFFFFFFFu, plain))
                score_unit_body<CB, NCHUNK, R, false>(p, lo, hi, 0ull, wmask, 32, row_first, lut32, hist32, lane, do_hist, two, nsent, cut_hi);
            else
                score_unit_guarded<CB, NCHUNK, R>(p, lo, hi, nn, wmask, nvalid, row_first, lut32, hist32, lane, do_hist, two, nsent, cut_hi);
        }
    }

    if (do_hist) {
        __syncthreads();
        for (uint32_t i = tid; i <= p.span; i += NT) {
            const uint32_t c = hist_s[i];
            if (c) atomicAdd(p.hist + i, (unsigned long long)c);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Sequence encoder: ASCII text -> 2-bit words + N bits.  One warp per unit (32 words = 1024 bytes of one sequence):
// the warp stages the unit's bytes in shared memory with aligned 16-byte loads, then lane l packs word l (32 symbols,
// four per 32-bit register as in K1, csrc/encode.cu).  Bytes past the end of a sequence read as 'A' with no N bit.
// counts[0] += N-or-other bases, counts[1] += bases that are neither ACGTacgt nor N/n.
// ---------------------------------------------------------------------------------------------------------------
struct SeqEncDesc {
    int64_t text_off;  // first byte of the sequence in d_text
    int64_t len;
    int64_t word_off;
    int64_t unit0;
};

#define SEQENC_WARPS 8

__global__ void __launch_bounds__(SEQENC_WARPS * 32) gb2_seq_encode_kernel(const uint8_t *__restrict__ text, int64_t text_bytes,
                                                                           const SeqEncDesc *__restrict__ desc, int64_t n_seqs,
                                                                           int64_t total_units, uint64_t *__restrict__ seq2,
                                                                           uint32_t *__restrict__ nbits,
                                                                           unsigned long long *__restrict__ counts)
{
    __shared__ __align__(16) uint8_t stage_all[SEQENC_WARPS][1024 + 32];
    const unsigned lane = threadIdx.x & 31u, wp = threadIdx.x >> 5;
    const int64_t u = (int64_t)blockIdx.x * SEQENC_WARPS + wp;
    if (u >= total_units) return;
    uint8_t *stage = stage_all[wp];
    // find the sequence of this unit (binary search over unit0; same on every lane)
    int64_t a = 0, b = n_seqs;
    while (b - a > 1) {
        const int64_t mid = (a + b) >> 1;
        if (__ldg(&desc[mid].unit0) <= u) a = mid;
        else b = mid;
    }
    const int64_t text_off = __ldg(&desc[a].text_off), len = __ldg(&desc[a].len), word_off = __ldg(&desc[a].word_off);
    const int64_t base0 = (u - __ldg(&desc[a].unit0)) * 1024;  // first base of the unit within the sequence
    const int64_t nbytes = min((int64_t)1024, len - base0);    // >= 1
    const uintptr_t gbeg = (uintptr_t)text + (uintptr_t)(text_off + base0);
    const uintptr_t abeg = gbeg & ~(uintptr_t)15;
    const int head = (int)(gbeg - abeg);
    const int nvec = (int)((head + nbytes + 15) >> 4);  // <= 66
    const uintptr_t buf_end = (uintptr_t)text + (uintptr_t)text_bytes;
    for (int i = lane; i < nvec; i += 32) {
        const uintptr_t g = abeg + ((uintptr_t)i << 4);
        uint4 v;
        if (g >= (uintptr_t)text && g + 16 <= buf_end) {
            v = __ldg(reinterpret_cast<const uint4 *>(g));
        } else {
            uint8_t tmp[16];
            for (int k = 0; k < 16; ++k) {
                const uintptr_t x = g + k;
                tmp[k] = (x >= (uintptr_t)text && x < buf_end) ? __ldg(reinterpret_cast<const uint8_t *>(x)) : (uint8_t)'A';
            }
            v = *reinterpret_cast<uint4 *>(tmp);
        }
        reinterpret_cast<uint4 *>(stage)[i] = v;
    }
    __syncwarp();
    const int64_t first = (int64_t)lane * 32;  // first base of this lane's word within the unit
    const int rem_total = (int)max((int64_t)0, min((int64_t)32, nbytes - first));
    if (rem_total > 0) {
        const uint32_t addr = (uint32_t)head + (uint32_t)first;
        const uint32_t *sw = reinterpret_cast<const uint32_t *>(stage) + (addr >> 2);
        const uint32_t sel = 0x3210u + 0x1111u * (addr & 3u);
        uint32_t lo = 0, hi = 0, bad = 0;
        uint32_t cur = sw[0];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int rem = rem_total - 4 * j;
            if (rem > 0) {
                const uint32_t nxt = sw[j + 1];
                uint32_t v = __byte_perm(cur, nxt, sel);
                cur = nxt;
                if (rem < 4) {
                    const uint32_t keep = 0xFFFFFFFFu >> (8 * (4 - rem));
                    v = (v & keep) | (0x41414141u & ~keep);
                }
                const uint32_t uu = v & 0xDFDFDFDFu;
                const uint32_t t = (v >> 1) & 0x03030303u;
                const uint32_t c = t ^ ((t >> 1) & 0x01010101u);
                const uint32_t ge2 = (c >> 1) & 0x01010101u;
                const uint32_t eq3 = ge2 & c;
                const uint32_t letters = 0x41414141u + 2u * c + 2u * ge2 + 11u * eq3;
                bad |= uu ^ letters;
                const uint32_t four = (c * 0x01041040u) >> 24;
                if (j < 4) lo |= four << (8 * j);
                else hi |= four << (8 * (j - 4));
            }
        }
        uint32_t nb = 0;
        if (bad) {  // rare: per-symbol pass; flagged bases are stored as code 0
            const uint8_t *sb = stage + addr;
            uint32_t other = 0;
            for (int i = 0; i < rem_total; ++i) {
                const uint32_t c = sb[i], uu = c & 0xDFu;
                const bool acgt = (uu == 'A') | (uu == 'C') | (uu == 'G') | (uu == 'T');
                if (!acgt) {
                    nb |= 1u << i;
                    if (uu != 'N') ++other;
                    if (i < 16) lo &= ~(3u << (2 * i));
                    else hi &= ~(3u << (2 * (i - 16)));
                }
            }
            if (counts) {
                atomicAdd(counts + 0, (unsigned long long)__popc(nb));
                if (other) atomicAdd(counts + 1, (unsigned long long)other);
            }
        }
        const int64_t wi = word_off + (base0 >> 5) + lane;
        seq2[wi] = ((uint64_t)hi << 32) | lo;
        if (nbits) nbits[wi] = nb;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
static inline size_t al256_seq(size_t x) { return (x + 255) & ~(size_t)255; }

static uint64_t hash_i64(const int64_t *a, int64_t n, uint64_t h)
{
    for (int64_t i = 0; i < n; ++i) {
        h ^= (uint64_t)a[i] + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
        h *= 0xFF51AFD7ED558CCDull;
    }
    return h;
}

// Uploads (or reuses) the descriptor array of a batch into the context's descriptor buffer `slot` (0 = scorer,
// 1 = encoder).  The previous array of a slot is reused when the layout is the same (a step loop scores the same
// layout again and again).  *total_units_out = units of the batch.
static int upload_desc(gb2_ctx *ctx, int slot, int w_for_units, int64_t n_seqs, const int64_t *h_a, const int64_t *h_len,
                       const int64_t *h_word_off, const int64_t *h_row_base, const void **d_out,
                       int64_t *total_units_out)
{
    uint64_t h = 1469598103934665603ull ^ (uint64_t)n_seqs ^ ((uint64_t)slot << 56) ^ ((uint64_t)w_for_units << 48);
    if (h_a) h = hash_i64(h_a, n_seqs, h);
    h = hash_i64(h_len, n_seqs, h);
    h = hash_i64(h_word_off, n_seqs, h);
    if (h_row_base) h = hash_i64(h_row_base, n_seqs, h ^ 0xABCDull);
    gb2_ctx::DescSlot &ds = ctx->desc[slot];
    if (ds.d_ptr && ds.hash == h && ds.n_seqs == n_seqs) {
        *d_out = ds.d_ptr;
        *total_units_out = ds.total_units;
        return GB2_OK;
    }
    std::vector<int64_t> host((size_t)(n_seqs + 1) * 4);
    int64_t units = 0, row = 0;
    for (int64_t s = 0; s < n_seqs; ++s) {
        const int64_t len = h_len[s];
        const int64_t nwords = (len + 31) >> 5;
        int64_t *e = &host[(size_t)s * 4];
        if (slot == 0) {
            e[0] = h_word_off[s]; e[1] = len; e[2] = h_row_base ? h_row_base[s] : row; e[3] = units;
            row += std::max<int64_t>(0, len - w_for_units + 1);
            // units cover the window start positions only (a sequence shorter than the motif has none)
            const int64_t nwin = std::max<int64_t>(0, len - w_for_units + 1);
            units += (nwin + 1023) >> 10;
        } else {
            e[0] = h_a[s]; e[1] = len; e[2] = h_word_off[s]; e[3] = units;
            units += (nwords + 31) >> 5;
        }
    }
    int64_t *e = &host[(size_t)n_seqs * 4];
    e[0] = 0; e[1] = 0; e[2] = row; e[3] = units;
    const size_t bytes = host.size() * sizeof(int64_t);
    if (ds.bytes < bytes) {
        if (ds.d_ptr) {
            GB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            GB2_CUDA(ctx, cudaFree(ds.d_ptr));
            ds.d_ptr = nullptr; ds.bytes = 0;
        }
        GB2_CUDA(ctx, cudaMalloc(&ds.d_ptr, bytes + bytes / 2 + 256));
        ds.bytes = bytes + bytes / 2 + 256;
    }
    // pageable source: the call returns once the bytes are staged, so `host` may go out of scope
    GB2_CUDA(ctx, cudaMemcpyAsync(ds.d_ptr, host.data(), bytes, cudaMemcpyHostToDevice, ctx->stream));
    ds.hash = h; ds.n_seqs = n_seqs; ds.total_units = units;
    *d_out = ds.d_ptr;
    *total_units_out = units;
    return GB2_OK;
}

template <int CB, int NCHUNK, int R>
static int launch_seq(gb2_ctx *ctx, const SeqScoreParams &q, size_t smem, int grid)
{
    // 32 warps per SM: measured 3.83 ms per 2.5e9 windows (CTCF, both strands) against 4.62 ms with 16 warps
    auto kern = gb2_score_seq_kernel<CB, NCHUNK, R, 1024>;
    GB2_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, 1024, smem, ctx->stream>>>(q);
    GB2_LAUNCH_CHECK(ctx);
    return GB2_OK;
}

template <int CB, int NCHUNK>
static int dispatch_seq_r(gb2_ctx *ctx, int R, const SeqScoreParams &q, size_t smem, int grid)
{
    switch (R) {
    case 32: return launch_seq<CB, NCHUNK, 32>(ctx, q, smem, grid);
    case 16: return launch_seq<CB, NCHUNK, 16>(ctx, q, smem, grid);
    default: return launch_seq<CB, NCHUNK, 8>(ctx, q, smem, grid);
    }
}

// launches the sequence kernel for a filled parameter block (descriptors already on the device)
static int gb2_launch_score_seq(gb2_ctx *ctx, const gb2_motif *m, const SeqScoreParams &q)
{
    const size_t smem = (size_t)m->smem_bytes;
    const int64_t per_cta = 32;  // warps per CTA
    const int grid = (int)std::min<int64_t>(ctx->sm_count, std::max<int64_t>(1, gb2_div_up(q.total_units, per_cta)));
    if (m->chunk_bases == 3) {
        switch (m->n_chunks) {
        case 7: return dispatch_seq_r<3, 7>(ctx, m->replicas, q, smem, grid);
        case 8: return dispatch_seq_r<3, 8>(ctx, m->replicas, q, smem, grid);
        case 9: return dispatch_seq_r<3, 9>(ctx, m->replicas, q, smem, grid);
        case 10: return dispatch_seq_r<3, 10>(ctx, m->replicas, q, smem, grid);
        default: return dispatch_seq_r<3, 11>(ctx, m->replicas, q, smem, grid);
        }
    }
    switch (m->n_chunks) {
    case 1: return dispatch_seq_r<4, 1>(ctx, m->replicas, q, smem, grid);
    case 2: return dispatch_seq_r<4, 2>(ctx, m->replicas, q, smem, grid);
    case 3: return dispatch_seq_r<4, 3>(ctx, m->replicas, q, smem, grid);
    case 4: return dispatch_seq_r<4, 4>(ctx, m->replicas, q, smem, grid);
    case 5: return dispatch_seq_r<4, 5>(ctx, m->replicas, q, smem, grid);
    case 6: return dispatch_seq_r<4, 6>(ctx, m->replicas, q, smem, grid);
    case 7: return dispatch_seq_r<4, 7>(ctx, m->replicas, q, smem, grid);
    default: return dispatch_seq_r<4, 8>(ctx, m->replicas, q, smem, grid);
    }
}

extern "C" int gb2_score_sequences(gb2_ctx *ctx, const gb2_motif *m, const uint64_t *d_seq2, const uint32_t *d_nbits,
                                   int64_t n_seqs, const int64_t *h_len, const int64_t *h_word_off,
                                   const int64_t *h_row_base, uint64_t row_base, int strands, double p_threshold,
                                   uint64_t *d_hist, gb2_hit *d_hits, uint64_t hit_capacity, uint64_t *d_hit_count,
                                   uint32_t *d_dense, uint64_t *h_n_windows)
{
    if (!ctx || !m) return GB2_ERR_ARG;
    if (h_n_windows) *h_n_windows = 0;
    GB2_REQUIRE(ctx, n_seqs >= 0, "gb2_score_sequences: negative sequence count");
    GB2_REQUIRE(ctx, strands == 1 || strands == 2, "gb2_score_sequences: strands must be 1 or 2");
    GB2_REQUIRE(ctx, !(p_threshold != p_threshold) && p_threshold > 0.0, "gb2_score_sequences: threshold must be > 0");
    GB2_REQUIRE(ctx, m->device == ctx->device, "gb2_score_sequences: motif lives on device %d, context on %d", m->device, ctx->device);
    GB2_REQUIRE(ctx, m->w <= GB2_NARROW_WIDTH, "gb2_score_sequences: motifs wider than %d take the k-mer form (gb2_score)", GB2_NARROW_WIDTH);
    GB2_REQUIRE(ctx, m->replicas >= 8, "gb2_score_sequences: score span too large for the sequence kernel");
    if (n_seqs == 0) return GB2_OK;
    GB2_REQUIRE(ctx, d_seq2 && h_len && h_word_off, "gb2_score_sequences: null buffer");
    GB2_REQUIRE(ctx, ((uintptr_t)d_seq2 & 7u) == 0, "gb2_score_sequences: sequence words must be 8-byte aligned");
    GB2_REQUIRE(ctx, d_hits == nullptr || d_hit_count != nullptr, "gb2_score_sequences: hit buffer without a counter");
    int64_t n_windows = 0;
    for (int64_t s = 0; s < n_seqs; ++s) {
        GB2_REQUIRE(ctx, h_len[s] >= 0 && h_word_off[s] >= 0, "gb2_score_sequences: negative length or offset (sequence %lld)", (long long)s);
        n_windows += std::max<int64_t>(0, h_len[s] - m->w + 1);
    }
    if (h_n_windows) *h_n_windows = (uint64_t)n_windows;
    if (n_windows == 0) return GB2_OK;
    GB2_CUDA(ctx, cudaSetDevice(ctx->device));

    SeqScoreParams q;
    memset(&q, 0, sizeof(q));
    q.sp.two_strands = strands == 2;
    q.sp.hist = (unsigned long long *)d_hist;
    q.sp.hits = d_hits;
    q.sp.hit_capacity = hit_capacity;
    q.sp.hit_count = (unsigned long long *)d_hit_count;
    q.sp.dense = d_dense;
    q.sp.row_base = row_base;  // added to the row of every hit record; dense scores are indexed without it
    int rc = gb2_fill_score_params(ctx, m, p_threshold, q.sp);
    if (rc != GB2_OK) return rc;
    q.seq2 = d_seq2;
    q.nbits = d_nbits;
    q.n_seqs = n_seqs;
    q.w = m->w;
    const void *d_desc = nullptr;
    rc = upload_desc(ctx, 0, m->w, n_seqs, nullptr, h_len, h_word_off, h_row_base, &d_desc, &q.total_units);
    if (rc != GB2_OK) return rc;
    q.desc = (const SeqDesc *)d_desc;

    return gb2_launch_score_seq(ctx, m, q);
}

extern "C" int gb2_encode_sequences(gb2_ctx *ctx, const uint8_t *d_text, int64_t text_bytes, int64_t n_seqs,
                                    const int64_t *h_text_off, const int64_t *h_len, const int64_t *h_word_off,
                                    uint64_t *d_seq2, uint32_t *d_nbits, uint64_t *d_counts)
{
    if (!ctx) return GB2_ERR_ARG;
    GB2_REQUIRE(ctx, n_seqs >= 0 && text_bytes >= 0, "gb2_encode_sequences: negative size");
    if (n_seqs == 0) return GB2_OK;
    GB2_REQUIRE(ctx, d_text && h_text_off && h_len && h_word_off && d_seq2, "gb2_encode_sequences: null buffer");
    for (int64_t s = 0; s < n_seqs; ++s)
        GB2_REQUIRE(ctx, h_len[s] >= 0 && h_text_off[s] >= 0 && h_text_off[s] + h_len[s] <= text_bytes && h_word_off[s] >= 0,
                    "gb2_encode_sequences: sequence %lld lies outside the text buffer", (long long)s);
    GB2_CUDA(ctx, cudaSetDevice(ctx->device));
    const void *d_desc = nullptr;
    int64_t total_units = 0;
    int rc = upload_desc(ctx, 1, 0, n_seqs, h_text_off, h_len, h_word_off, nullptr, &d_desc, &total_units);
    if (rc != GB2_OK) return rc;
    if (total_units == 0) return GB2_OK;
    const int64_t blocks = gb2_div_up(total_units, SEQENC_WARPS);
    GB2_REQUIRE(ctx, blocks < ((int64_t)1 << 31), "gb2_encode_sequences: too many bases for one launch");
    gb2_seq_encode_kernel<<<(unsigned)blocks, SEQENC_WARPS * 32, 0, ctx->stream>>>(
        d_text, text_bytes, (const SeqEncDesc *)d_desc, n_seqs, total_units, d_seq2, d_nbits, (unsigned long long *)d_counts);
    GB2_LAUNCH_CHECK(ctx);
    return GB2_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// gb2_scan_host_sequences: whole sequences in HOST memory -> hit table in host memory.
// The batch is cut into chunks of at most CHUNK_BASES bases; a sequence that does not fit the rest of a chunk is cut
// into PIECES that overlap by w-1 bases (every window belongs to exactly one piece; piece starts are multiples of 32
// bases so that 2-bit input is copied as whole words).  Chunk c+1 is copied while chunk c is encoded and scored.
// ---------------------------------------------------------------------------------------------------------------
#include "scan_tail.cuh"

namespace {
struct Piece {
    int64_t src;   // first byte (ASCII) or first word (2-bit) in the host buffer
    int64_t len;   // bases
    int64_t row0;  // global index of its first window
};
struct Run {  // one host->device copy: `count` bytes / words from `src` to device offset `dst`
    int64_t src, dst, count;
};
struct Chunk {
    size_t piece0 = 0, n_pieces = 0;
    size_t run0 = 0, n_runs = 0;
    int64_t words = 0, text_bytes = 0, units_score = 0, units_enc = 0;
};
}  // namespace

extern "C" int gb2_scan_host_sequences(gb2_ctx *ctx, const gb2_motif *m, int format, const void *h_data,
                                       const uint32_t *h_nbits, int64_t n_seqs, const int64_t *h_off, const int64_t *h_len,
                                       int strands, double p_threshold, int q_filter, int want_q, uint64_t hit_capacity,
                                       uint64_t *h_row, uint8_t *h_strand, int32_t *h_iscore, double *h_score, double *h_p,
                                       double *h_q, uint64_t *h_n_hits, uint64_t *h_stats)
{
    if (!ctx || !m) return GB2_ERR_ARG;
    gb2_scan_out o = {h_row, h_strand, h_iscore, h_score, h_p, h_q, h_n_hits, h_stats};
    int rc = gb2_scan_check_args(ctx, "gb2_scan_host_sequences", strands, q_filter, want_q, hit_capacity, o);
    if (rc != GB2_OK) return rc;
    GB2_REQUIRE(ctx, format == 0 || format == 1, "gb2_scan_host_sequences: format must be 0 (ASCII) or 1 (2-bit words)");
    GB2_REQUIRE(ctx, n_seqs >= 0, "gb2_scan_host_sequences: negative sequence count");
    GB2_REQUIRE(ctx, m->w <= GB2_NARROW_WIDTH, "gb2_scan_host_sequences: motifs wider than %d take the k-mer form", GB2_NARROW_WIDTH);
    GB2_REQUIRE(ctx, m->replicas >= 8, "gb2_scan_host_sequences: score span too large for the sequence kernel");
    GB2_REQUIRE(ctx, n_seqs == 0 || (h_data && h_off && h_len), "gb2_scan_host_sequences: null buffer");
    GB2_REQUIRE(ctx, format == 1 || h_nbits == nullptr, "gb2_scan_host_sequences: N bits are derived from ASCII input");
    const int w = m->w;
    const bool ascii = format == 0;
    // Host-side packing (host_pack.cpp): worker threads re-code chunks from the BACK of the chunk list into 2-bit words in
    // pinned staging while the copy engine moves the text of the chunks at the FRONT; both meet in the middle.
    // GB2_HOST_PACK_THREADS overrides the thread count (0 = off); default: the host threads this rank can count on.
    int pack_threads = 0;
    if (ascii) {
        const int hw = (int)std::thread::hardware_concurrency();
        // measured on 16- and 24-thread hosts (tools/bench_e2e.py): the rate grows up to ~12 threads and is flat beyond
        pack_threads = std::max(0, std::min(16, ctx->comm_world <= 1 ? hw : hw / (2 * ctx->comm_world)));  // 2 ranks: 6 each of 24 (measured)
        // Packing pays while PCIe is the limit of this GPU's copies.  Measured (profiles/r02_e2e_packers.json): per step of the
        // headline workload 46.6 -> 25 ms on one GPU, 46.5 -> 39.7 ms with two ranks on the host, but 51 -> 67 ms with four
        // and 110 -> 130 ms with eight: there the HOST memory system is the limit (a packed base costs it ~1.6 bytes of traffic
        // instead of 1; on the 8-GPU box raw copies reach 23 GB/s per GPU instead of 55, tools/h2d_probe.py).
        if (ctx->comm_world > 2) pack_threads = 0;
        if (const char *t = getenv("GB2_HOST_PACK_THREADS")) pack_threads = std::max(0, std::min(64, atoi(t)));
    }
    // chunk size in bases (GB2_SEQ_CHUNK_BASES overrides it: the tests use small chunks to exercise the piece logic).  With the
    // packers a finer grain wins -- the two sides meet with less idle time at the end (measured per step, 16 threads: 27.1 /
    // 27.2 / 25.1 / 25.0 ms at 64 / 32 / 16 / 8 Mi bases); the copy engine alone prefers the coarse one (46.5 / 46.8 / 47.2 / 47.7).
    int64_t CHUNK_BASES = pack_threads ? (int64_t)1 << 24 : (int64_t)1 << 26;
    if (const char *t = getenv("GB2_SEQ_CHUNK_BASES")) CHUNK_BASES = std::max<int64_t>(1024, atoll(t));
    const int64_t MIN_PIECE = std::max<int64_t>(64 + w, std::min<int64_t>((int64_t)1 << 16, CHUNK_BASES / 4));

    const bool timing = getenv("GB2_SCAN_TIMING") != nullptr;  // phase times of this call on stderr (events with timing)
    const auto t_host0 = std::chrono::steady_clock::now();
    auto host_ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_host0).count(); };
    double ms_plan = 0, ms_issue = 0, ms_tail0 = 0;
    cudaEvent_t tev[4] = {nullptr, nullptr, nullptr, nullptr};  // loop start, first copy done, last copy done, last kernel done
    bool tev_first_done = false;
    uint64_t host_flagged[2] = {0, 0};  // N-or-other / other bases counted by the host packers
    // ---- plan: pieces, chunks, copies
    std::vector<Piece> pieces;
    std::vector<Chunk> chunks;
    std::vector<Run> runs;
    int64_t n_windows = 0;
    uint64_t overlap_n = 0, overlap_other = 0;  // non-ACGT bases inside the w-1 bases two pieces of a sequence share
    {
        Chunk cur;
        int64_t fill = 0;
        auto close = [&]() {
            if (cur.n_pieces) chunks.push_back(cur);
            cur = Chunk();
            cur.piece0 = pieces.size();
            fill = 0;
        };
        for (int64_t s = 0; s < n_seqs; ++s) {
            const int64_t len = h_len[s];
            GB2_REQUIRE(ctx, len >= 0 && h_off[s] >= 0, "gb2_scan_host_sequences: negative length or offset (sequence %lld)", (long long)s);
            const int64_t nwin = len - w + 1;
            if (nwin <= 0) continue;
            int64_t a = 0;
            while (a < nwin) {
                const int64_t rem = len - a, room = CHUNK_BASES - fill;
                int64_t take;
                if (rem <= room) {
                    take = rem;
                } else if (room >= MIN_PIECE) {
                    take = (room - (w - 1)) / 32 * 32 + (w - 1);  // the next piece starts at a multiple of 32
                } else {
                    close();
                    continue;
                }
                Piece p;
                p.src = ascii ? h_off[s] + a : h_off[s] + (a >> 5);
                p.len = take;
                p.row0 = n_windows + a;
                pieces.push_back(p);
                cur.n_pieces++;
                fill += (take + 31) / 32 * 32;
                if (ascii && take < rem) {  // the next piece re-reads the last w-1 bases of this one: count them once
                    const uint8_t *t = (const uint8_t *)h_data + h_off[s] + a + take - (w - 1);
                    for (int k = 0; k < w - 1; ++k) {
                        const uint8_t u = t[k] & 0xDFu;
                        if (!(u == 'A' || u == 'C' || u == 'G' || u == 'T')) {
                            ++overlap_n;
                            if (u != 'N') ++overlap_other;
                        }
                    }
                }
                a += take - w + 1;
                if (fill + MIN_PIECE > CHUNK_BASES) close();
            }
            n_windows += nwin;
        }
        close();
    }
    if (n_windows == 0) return GB2_OK;
    GB2_CUDA(ctx, cudaSetDevice(ctx->device));

    // per chunk: device word offsets, copy runs, unit counts; descriptor arrays for every chunk back to back
    std::vector<int64_t> desc_s, desc_e;  // SeqDesc / SeqEncDesc as 4 x int64 (+ one sentinel per chunk)
    std::vector<int64_t> piece_words(pieces.size(), 0);  // first device word of every piece inside its chunk (ASCII input)
    std::vector<size_t> desc_pos(chunks.size());
    int64_t max_words = 0, max_text = 0;
    for (size_t c = 0; c < chunks.size(); ++c) {
        Chunk &ch = chunks[c];
        ch.run0 = runs.size();
        desc_pos[c] = desc_s.size() / 4;
        int64_t words = 0, dst = 0, us = 0, ue = 0;
        Run run = {0, 0, 0};
        bool open = false;
        for (size_t k = ch.piece0; k < ch.piece0 + ch.n_pieces; ++k) {
            const Piece &p = pieces[k];
            const int64_t pw = (p.len + 31) >> 5;            // device words of the piece
            const int64_t cnt = ascii ? p.len : pw;          // host elements to copy
            const int64_t gap_max = ascii ? 4096 : 512;
            if (open && p.src >= run.src && p.src <= run.src + run.count + gap_max) {
                run.count = std::max(run.count, p.src + cnt - run.src);
            } else {
                if (open) { runs.push_back(run); dst = run.dst + ((run.count + 15) & ~(int64_t)15); }
                run.src = p.src; run.dst = dst; run.count = cnt;
                open = true;
            }
            const int64_t in_dev = run.dst + (p.src - run.src);  // where the piece starts in the device copy
            const int64_t nwin = p.len - w + 1;
            if (ascii) {
                piece_words[k] = words;
                desc_s.insert(desc_s.end(), {words, p.len, p.row0, us});
                desc_e.insert(desc_e.end(), {in_dev, p.len, words, ue});
            } else {
                desc_s.insert(desc_s.end(), {in_dev, p.len, p.row0, us});
            }
            us += (nwin + 1023) >> 10;
            ue += (pw + 31) >> 5;
            words += pw;
        }
        if (open) { runs.push_back(run); dst = run.dst + ((run.count + 15) & ~(int64_t)15); }
        desc_s.insert(desc_s.end(), {0, 0, 0, us});
        if (ascii) desc_e.insert(desc_e.end(), {0, 0, 0, ue});
        ch.n_runs = runs.size() - ch.run0;
        ch.words = ascii ? words : dst;
        ch.text_bytes = ascii ? dst : 0;
        ch.units_score = us;
        ch.units_enc = ue;
        max_words = std::max(max_words, ch.words);
        max_text = std::max(max_text, ch.text_bytes);
    }

    const size_t b_text = ascii ? al256_seq((size_t)max_text + 64) : 0;
    const size_t b_words = al256_seq((size_t)max_words * 8 + 64);
    const size_t b_nbits = al256_seq((size_t)max_words * 4 + 64);
    const size_t b_desc = al256_seq(desc_s.size() * 8) + al256_seq(desc_e.size() * 8);
    if (chunks.size() < 4) pack_threads = 0;  // nothing to overlap
    const int n_slots = pack_threads ? 8 : 0;  // pinned staging slots of one packed chunk each (a slot is busy from the first
                                               // block packed until its copy has left the queue behind up to two text chunks)
    const int nwordbuf = ascii ? (pack_threads ? 3 : 1) : 2;  // ASCII: [0] the device encoder's output, [1], [2] host-packed chunks
    const size_t total = 2 * b_text + nwordbuf * (b_words + b_nbits) + b_desc + gb2_scan_tail_bytes(m, hit_capacity);
    char *q = nullptr;
    rc = gb2_pool_reserve(ctx, total, &q);
    if (rc != GB2_OK) return rc;
    uint8_t *d_text[2] = {nullptr, nullptr};
    uint64_t *d_words[3] = {nullptr, nullptr, nullptr};
    uint32_t *d_nb[3] = {nullptr, nullptr, nullptr};
    for (int i = 0; i < 2 && ascii; ++i) { d_text[i] = (uint8_t *)q; q += b_text; }
    for (int i = 0; i < nwordbuf; ++i) {
        d_words[i] = (uint64_t *)q; q += b_words;
        d_nb[i] = (uint32_t *)q; q += b_nbits;
    }
    int64_t *d_desc_s = (int64_t *)q; q += al256_seq(desc_s.size() * 8);
    int64_t *d_desc_e = (int64_t *)q; q += al256_seq(desc_e.size() * 8);
    gb2_scan_bufs b;
    gb2_scan_tail_carve(q, m, hit_capacity, b);

    cudaEvent_t copied[2] = {nullptr, nullptr}, consumed[2] = {nullptr, nullptr};
    cudaError_t e = cudaSuccess;
#define SH_CUDA(call)                                                                                         \
    do {                                                                                                      \
        e = (call);                                                                                           \
        if (e != cudaSuccess) {                                                                               \
            GB2_SET_ERR(ctx, "%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e));      \
            rc = GB2_ERR_CUDA;                                                                                \
            goto done;                                                                                        \
        }                                                                                                     \
    } while (0)
#define SH_CUDA_RET(call)                                                                                     \
    do {                                                                                                      \
        cudaError_t e2__ = (call);                                                                            \
        if (e2__ != cudaSuccess) {                                                                            \
            GB2_SET_ERR(ctx, "%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e2__));   \
            return GB2_ERR_CUDA;                                                                              \
        }                                                                                                     \
    } while (0)
    {
        for (int i = 0; i < 2; ++i) {
            SH_CUDA(cudaEventCreateWithFlags(&copied[i], cudaEventDisableTiming));
            SH_CUDA(cudaEventCreateWithFlags(&consumed[i], cudaEventDisableTiming));
        }
        SH_CUDA(cudaMemsetAsync(b.d_hist, 0, b.hist_and_cnt_bytes, ctx->stream));
        SH_CUDA(cudaMemcpyAsync(d_desc_s, desc_s.data(), desc_s.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
        if (ascii) SH_CUDA(cudaMemcpyAsync(d_desc_e, desc_e.data(), desc_e.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
        SH_CUDA(cudaEventRecord(consumed[0], ctx->stream));
        SH_CUDA(cudaEventRecord(consumed[1], ctx->stream));

        SeqScoreParams sq;
        memset(&sq, 0, sizeof(sq));
        sq.sp.two_strands = strands == 2;
        sq.sp.hist = want_q ? (unsigned long long *)b.d_hist : nullptr;
        sq.sp.hits = b.d_hits;
        sq.sp.hit_capacity = hit_capacity;
        sq.sp.hit_count = (unsigned long long *)(b.d_cnt + 2);
        sq.sp.dense = nullptr;
        sq.sp.row_base = 0;
        rc = gb2_fill_score_params(ctx, m, p_threshold, sq.sp);
        if (rc != GB2_OK) goto done;
        sq.w = w;

        if (timing) {
            for (int i = 0; i < 4; ++i) SH_CUDA(cudaEventCreate(&tev[i]));
            ms_plan = host_ms();
            SH_CUDA(cudaEventRecord(tev[0], ctx->stream));
        }
        // ---- one chunk from the caller's buffer as it is (text or 2-bit words): copy, [encode], score
        int buf = 0;
        uint64_t h2d_bytes = desc_s.size() * 8 + (ascii ? desc_e.size() * 8 : 0);
        size_t n_given = 0, n_packed = 0;
        auto issue_raw = [&](size_t c) -> int {
            const Chunk &ch = chunks[c];
            ++n_given;
            SH_CUDA_RET(cudaStreamWaitEvent(ctx->copy_stream, consumed[buf], 0));
            for (size_t r = ch.run0; r < ch.run0 + ch.n_runs; ++r) {
                const Run &run = runs[r];
                h2d_bytes += ascii ? (uint64_t)run.count : (uint64_t)run.count * (h_nbits ? 12 : 8);
                if (ascii) {
                    SH_CUDA_RET(cudaMemcpyAsync(d_text[buf] + run.dst, (const uint8_t *)h_data + run.src, (size_t)run.count,
                                                cudaMemcpyHostToDevice, ctx->copy_stream));
                } else {
                    SH_CUDA_RET(cudaMemcpyAsync(d_words[buf] + run.dst, (const uint64_t *)h_data + run.src, (size_t)run.count * 8,
                                                cudaMemcpyHostToDevice, ctx->copy_stream));
                    if (h_nbits)
                        SH_CUDA_RET(cudaMemcpyAsync(d_nb[buf] + run.dst, h_nbits + run.src, (size_t)run.count * 4,
                                                    cudaMemcpyHostToDevice, ctx->copy_stream));
                }
            }
            SH_CUDA_RET(cudaEventRecord(copied[buf], ctx->copy_stream));
            if (timing && !tev_first_done) { SH_CUDA_RET(cudaEventRecord(tev[1], ctx->copy_stream)); tev_first_done = true; }
            SH_CUDA_RET(cudaStreamWaitEvent(ctx->stream, copied[buf], 0));
            const int wb = ascii ? 0 : buf;
            if (ascii) {
                const int64_t blocks = gb2_div_up(ch.units_enc, SEQENC_WARPS);
                gb2_seq_encode_kernel<<<(unsigned)blocks, SEQENC_WARPS * 32, 0, ctx->stream>>>(
                    d_text[buf], ch.text_bytes, (const SeqEncDesc *)(d_desc_e + 4 * desc_pos[c]), (int64_t)ch.n_pieces,
                    ch.units_enc, d_words[0], d_nb[0], (unsigned long long *)b.d_cnt);
                ctx->launches++;
                SH_CUDA_RET(cudaGetLastError());
                SH_CUDA_RET(cudaEventRecord(consumed[buf], ctx->stream));  // the text buffer is free once encoded
            }
            sq.seq2 = d_words[wb];
            sq.nbits = (ascii || h_nbits) ? d_nb[wb] : nullptr;
            sq.desc = (const SeqDesc *)(d_desc_s + 4 * desc_pos[c]);
            sq.n_seqs = (int64_t)ch.n_pieces;
            sq.total_units = ch.units_score;
            int r2 = gb2_launch_score_seq(ctx, m, sq);
            if (r2 != GB2_OK) return r2;
            if (!ascii) SH_CUDA_RET(cudaEventRecord(consumed[buf], ctx->stream));
            buf ^= 1;
            return GB2_OK;
        };

        if (!pack_threads) {
            for (size_t c = 0; c < chunks.size() && rc == GB2_OK; ++c) rc = issue_raw(c);
            if (rc != GB2_OK) goto done;
        } else {
            // ---- hybrid: raw chunks from the front, host-packed chunks from the back -----------------------------------
            char *pin = nullptr;
            const size_t slot_bytes = b_words + b_nbits;
            rc = gb2_pinned_reserve(ctx, (size_t)n_slots * slot_bytes, &pin);
            if (rc != GB2_OK) goto done;
            struct Block { const uint8_t *src; int64_t n_bases; int64_t word; };
            struct Job {
                size_t chunk = 0;
                int slot = 0;
                std::vector<Block> blocks;
                size_t next = 0;                 // guarded by mu
                std::atomic<size_t> done{0};
                std::atomic<uint64_t> flagged{0};  // bases of this chunk that are not A/C/G/T: 0 = its N bits need not travel
            };
            std::mutex mu;
            std::condition_variable cv;
            size_t front = 0, back = chunks.size();  // chunks [front, back) are unclaimed
            std::vector<int> free_slots;
            for (int i = 0; i < n_slots; ++i) free_slots.push_back(i);
            std::deque<Job *> ready;
            Job *open = nullptr;
            size_t jobs_claimed = 0, jobs_issued = 0;
            bool stop = false;
            std::atomic<uint64_t> host_invalid{0}, host_other{0};
            const int64_t BLOCK_BASES = (int64_t)1 << 19;  // work unit of one packer thread (a multiple of 32)
            std::atomic<bool> worker_failed{false};
            auto worker_body = [&]() {
                for (;;) {
                    Job *job = nullptr;
                    Block blk{};
                    size_t n_blocks = 0;
                    {
                        std::unique_lock<std::mutex> lk(mu);
                        for (;;) {
                            if (stop) return;
                            if (open == nullptr) {
                                if (front >= back) return;  // every chunk is claimed; jobs under way are finished by their workers
                                if (free_slots.empty()) { cv.wait_for(lk, std::chrono::milliseconds(1)); continue; }
                                Job *j = new Job();
                                j->chunk = --back;
                                j->slot = free_slots.back();
                                free_slots.pop_back();
                                const Chunk &ch = chunks[j->chunk];
                                for (size_t k = ch.piece0; k < ch.piece0 + ch.n_pieces; ++k)
                                    for (int64_t a = 0; a < pieces[k].len; a += BLOCK_BASES)
                                        j->blocks.push_back(Block{(const uint8_t *)h_data + pieces[k].src + a,
                                                                  std::min(BLOCK_BASES, pieces[k].len - a), piece_words[k] + (a >> 5)});
                                ++jobs_claimed;
                                open = j;
                            }
                            job = open;
                            n_blocks = job->blocks.size();
                            blk = job->blocks[job->next++];
                            if (job->next == n_blocks) open = nullptr;
                            break;
                        }
                    }
                    uint64_t inv = 0, oth = 0;
                    uint64_t *hw_ = reinterpret_cast<uint64_t *>(pin + (size_t)job->slot * slot_bytes);
                    uint32_t *hn_ = reinterpret_cast<uint32_t *>(pin + (size_t)job->slot * slot_bytes + b_words);
                    gb2_host_pack_bases(blk.src, blk.n_bases, hw_ + blk.word, hn_ + blk.word, &inv, &oth);
                    if (inv) {
                        host_invalid.fetch_add(inv, std::memory_order_relaxed);
                        job->flagged.fetch_add(inv, std::memory_order_relaxed);
                    }
                    if (oth) host_other.fetch_add(oth, std::memory_order_relaxed);
                    if (job->done.fetch_add(1, std::memory_order_acq_rel) + 1 == n_blocks) {  // the main thread deletes the job once it is ready
                        std::lock_guard<std::mutex> lk(mu);
                        ready.push_back(job);
                        cv.notify_all();
                    }
                }
            };
            auto worker = [&]() {  // an exception must not leave a thread (std::terminate): report it and stop everybody
                try {
                    worker_body();
                } catch (...) {
                    worker_failed.store(true);
                    std::lock_guard<std::mutex> lk(mu);
                    stop = true;
                    cv.notify_all();
                }
            };
            std::vector<std::thread> pool;
            size_t raw_issued = 0, packed_chunks = 0;
            for (; front < 2 && rc == GB2_OK; ++front, ++raw_issued) rc = issue_raw(front);  // the copy engine starts before the threads do
            for (int t = 0; t < pack_threads && rc == GB2_OK; ++t) {
                try {
                    pool.emplace_back(worker);
                } catch (...) {  // no more threads to be had: the chunks nobody packs go through the copy engine as text
                    break;
                }
            }

            cudaEvent_t slot_copied[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}, pw_consumed[2] = {nullptr, nullptr};
            std::vector<int> slots_in_flight;
            int pbuf = 0;
            auto fail_hybrid = [&](int code) {
                {
                    std::lock_guard<std::mutex> lk(mu);
                    stop = true;
                    cv.notify_all();
                }
                for (auto &t : pool) t.join();
                pool.clear();
                return code;
            };
            for (int i = 0; i < n_slots && rc == GB2_OK; ++i)
                if (cudaEventCreateWithFlags(&slot_copied[i], cudaEventDisableTiming) != cudaSuccess) rc = GB2_ERR_CUDA;
            for (int i = 0; i < 2 && rc == GB2_OK; ++i) {
                if (cudaEventCreateWithFlags(&pw_consumed[i], cudaEventDisableTiming) != cudaSuccess) rc = GB2_ERR_CUDA;
                else if (cudaEventRecord(pw_consumed[i], ctx->stream) != cudaSuccess) rc = GB2_ERR_CUDA;
            }
            auto issue_packed = [&](Job *job) -> int {
                const Chunk &ch = chunks[job->chunk];
                const int wb = 1 + pbuf;
                const char *src = pin + (size_t)job->slot * slot_bytes;
                ++n_packed;
                SH_CUDA_RET(cudaStreamWaitEvent(ctx->copy_stream, pw_consumed[pbuf], 0));
                // a chunk without a single N (the usual case) travels as 0.25 byte per base: its N-bit words are all zero
                const bool has_n = job->flagged.load(std::memory_order_acquire) != 0;
                SH_CUDA_RET(cudaMemcpyAsync(d_words[wb], src, (size_t)ch.words * 8, cudaMemcpyHostToDevice, ctx->copy_stream));
                if (has_n)
                    SH_CUDA_RET(cudaMemcpyAsync(d_nb[wb], src + b_words, (size_t)ch.words * 4, cudaMemcpyHostToDevice, ctx->copy_stream));
                SH_CUDA_RET(cudaEventRecord(slot_copied[job->slot], ctx->copy_stream));
                SH_CUDA_RET(cudaStreamWaitEvent(ctx->stream, slot_copied[job->slot], 0));
                h2d_bytes += (uint64_t)ch.words * (has_n ? 12 : 8);
                sq.seq2 = d_words[wb];
                sq.nbits = has_n ? d_nb[wb] : nullptr;
                sq.desc = (const SeqDesc *)(d_desc_s + 4 * desc_pos[job->chunk]);
                sq.n_seqs = (int64_t)ch.n_pieces;
                sq.total_units = ch.units_score;
                int r2 = gb2_launch_score_seq(ctx, m, sq);
                if (r2 != GB2_OK) return r2;
                SH_CUDA_RET(cudaEventRecord(pw_consumed[pbuf], ctx->stream));
                slots_in_flight.push_back(job->slot);
                pbuf ^= 1;
                return GB2_OK;
            };
            // raw copies are paced (at most two chunks queued on the copy engine) so that the packers get their share
            while (rc == GB2_OK) {
                if (worker_failed.load()) {
                    GB2_SET_ERR(ctx, "gb2_scan_host_sequences: a host packer thread failed (out of memory?)");
                    rc = GB2_ERR_NOMEM;
                    break;
                }
                Job *job = nullptr;
                bool do_raw = false;
                size_t raw_c = 0;
                {
                    std::unique_lock<std::mutex> lk(mu);
                    // staging slots whose copy has finished go back to the packers
                    for (size_t i = 0; i < slots_in_flight.size();) {
                        if (cudaEventQuery(slot_copied[slots_in_flight[i]]) == cudaSuccess) {
                            free_slots.push_back(slots_in_flight[i]);
                            slots_in_flight.erase(slots_in_flight.begin() + (long)i);
                            cv.notify_all();
                        } else {
                            ++i;
                        }
                    }
                    if (!ready.empty()) {
                        job = ready.front();
                        ready.pop_front();
                    } else {
                        // a raw chunk goes into text buffer `buf`: at most two are queued on the copy engine
                        if (front < back && cudaEventQuery(copied[buf]) == cudaSuccess) {
                            raw_c = front++;
                            do_raw = true;
                        } else if (front >= back && jobs_issued == jobs_claimed && open == nullptr) {
                            break;  // every chunk is on the device queues
                        } else {
                            cv.wait_for(lk, std::chrono::microseconds(50));
                            continue;
                        }
                    }
                }
                if (job) {
                    rc = issue_packed(job);
                    delete job;
                    ++jobs_issued;
                    ++packed_chunks;
                } else if (do_raw) {
                    rc = issue_raw(raw_c);
                    ++raw_issued;
                }
            }
            rc = fail_hybrid(rc);  // stops and joins the packers (on success they have nothing left to do)
            for (Job *j : ready) delete j;  // only after an error: jobs nobody issued
            ready.clear();
            if (open) { delete open; open = nullptr; }
            host_flagged[0] = host_invalid.load();  // the device encoder counts the flagged bases of the raw chunks, the packers theirs
            host_flagged[1] = host_other.load();
            if (timing) fprintf(stderr, "gb2_scan_host_sequences: %d packer threads (simd %d): %zu chunks as text, %zu packed on the host\n",
                                pack_threads, gb2_host_pack_simd(), raw_issued, packed_chunks);
            cudaStreamSynchronize(ctx->copy_stream);
            for (int i = 0; i < 8; ++i)
                if (slot_copied[i]) cudaEventDestroy(slot_copied[i]);
            for (int i = 0; i < 2; ++i)
                if (pw_consumed[i]) cudaEventDestroy(pw_consumed[i]);
            if (rc != GB2_OK) goto done;
        }
        if (timing) SH_CUDA(cudaEventRecord(tev[2], ctx->copy_stream));
        ctx->last_h2d_bytes = h2d_bytes;
        ctx->last_chunks_given = n_given;
        ctx->last_chunks_packed = n_packed;
        if (timing) {
            SH_CUDA(cudaEventRecord(tev[3], ctx->stream));
            ms_issue = host_ms();
            SH_CUDA(cudaStreamSynchronize(ctx->stream));
            ms_tail0 = host_ms();
        }
        rc = gb2_scan_tail_finish(ctx, m, b, (uint64_t)n_windows * (uint64_t)strands, (uint64_t)n_windows, p_threshold,
                                  q_filter, want_q, hit_capacity, o);
        if (timing && tev[2]) {
            float first = 0, copies = 0, kernels = 0;
            cudaEventElapsedTime(&first, tev[0], tev[1]);
            cudaEventElapsedTime(&copies, tev[0], tev[2]);
            cudaEventElapsedTime(&kernels, tev[0], tev[3]);
            fprintf(stderr, "gb2_scan_host_sequences: %zu chunks, %zu copies; plan %.2f ms | all launches issued at %.2f ms | device: first copy done "
                    "+%.2f ms, last copy done +%.2f ms, last kernel done +%.2f ms | host: loop drained at %.2f ms, K5 + K6 + table back "
                    "%.2f ms, total %.2f ms\n", chunks.size(), runs.size(), ms_plan, ms_issue, first, copies, kernels, ms_tail0,
                    host_ms() - ms_tail0, host_ms());
        }
        if (o.h_stats && (rc == GB2_OK || rc == GB2_ERR_CAPACITY)) {
            o.h_stats[1] += host_flagged[0];
            o.h_stats[2] += host_flagged[1];
            o.h_stats[1] -= std::min<uint64_t>(o.h_stats[1], overlap_n);
            o.h_stats[2] -= std::min<uint64_t>(o.h_stats[2], overlap_other);
        }
    }
done:
#undef SH_CUDA
#undef SH_CUDA_RET
    cudaStreamSynchronize(ctx->copy_stream);
    cudaStreamSynchronize(ctx->stream);
    for (int i = 0; i < 2; ++i) {
        if (copied[i]) cudaEventDestroy(copied[i]);
        if (consumed[i]) cudaEventDestroy(consumed[i]);
    }
    for (int i = 0; i < 4; ++i)
        if (tev[i]) cudaEventDestroy(tev[i]);
    return rc;
}
