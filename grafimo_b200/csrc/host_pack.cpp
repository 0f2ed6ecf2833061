// host_pack.cpp -- transfer compression for gb2_scan_host_sequences: ASCII bases -> 2-bit words + N bits on host threads.
//
// The end-to-end rate of a scan over text in host memory is the PCIe rate (one byte per window at ~55 GB/s; the GPU scores
// those windows 12x faster than they arrive).  Host cores that would idle during the copy re-code part of the text into the
// layout the device encoder (gb2_seq_encode_kernel, seqscan.cu) would produce from it -- 2 bits per base + one N bit per base
// = 0.375 bytes per window -- so those chunks cross PCIe 2.7x smaller.  Nothing is scored here: the words go to the same
// scoring kernel as the device-encoded ones, and the output is bit-identical to the device encoder's (tests/test_gpu_sequences.py).
// Same rules as the device encoder: A/a=0 C/c=1 G/g=2 T/t=3; anything else is stored as code 0 with its N bit set; bases past
// the end of the piece are code 0 without N bit; *n_invalid counts flagged bases, *n_other those that are not N/n.
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace {

inline void pack_word_scalar(const uint8_t *t, int n, uint64_t &word, uint32_t &nb, uint64_t &invalid, uint64_t &other)
{
    uint64_t w = 0;
    uint32_t b = 0;
    for (int i = 0; i < n; ++i) {
        const uint32_t c = t[i], u = c & 0xDFu;
        uint64_t code = 0;
        switch (u) {
        case 'A': code = 0; break;
        case 'C': code = 1; break;
        case 'G': code = 2; break;
        case 'T': code = 3; break;
        default:
            b |= 1u << i;
            ++invalid;
            if (u != 'N') ++other;
        }
        w |= code << (2 * i);
    }
    word = w;
    nb = b;
}

#if defined(__x86_64__)
__attribute__((target("avx2,bmi2,popcnt"))) void pack_words_avx2(const uint8_t *t, int64_t n_words, uint64_t *words, uint32_t *nbits,
                                                                 uint64_t &invalid, uint64_t &other)
{
    const __m256i m_df = _mm256_set1_epi8((char)0xDF), m_3 = _mm256_set1_epi8(3), m_1 = _mm256_set1_epi8(1);
    const __m256i cA = _mm256_set1_epi8('A'), cC = _mm256_set1_epi8('C'), cG = _mm256_set1_epi8('G'), cT = _mm256_set1_epi8('T'),
                  cN = _mm256_set1_epi8('N');
    const uint64_t sel = 0x0303030303030303ull;
    uint64_t inv = 0, oth = 0;
    for (int64_t k = 0; k < n_words; ++k) {
        const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(t + 32 * k));
        const __m256i u = _mm256_and_si256(v, m_df);
        const __m256i x = _mm256_and_si256(_mm256_srli_epi16(v, 1), m_3);                    // A 0, C 1, G 3, T 2
        __m256i c = _mm256_xor_si256(x, _mm256_and_si256(_mm256_srli_epi16(x, 1), m_1));     // A 0, C 1, G 2, T 3
        const __m256i ok = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(u, cA), _mm256_cmpeq_epi8(u, cC)),
                                           _mm256_or_si256(_mm256_cmpeq_epi8(u, cG), _mm256_cmpeq_epi8(u, cT)));
        const uint32_t bad = ~(uint32_t)_mm256_movemask_epi8(ok);
        if (bad) {
            c = _mm256_and_si256(c, ok);
            const uint32_t is_n = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(u, cN));
            inv += (uint64_t)_mm_popcnt_u32(bad);
            oth += (uint64_t)_mm_popcnt_u32(bad & ~is_n);
        }
        const uint64_t l0 = (uint64_t)_mm256_extract_epi64(c, 0), l1 = (uint64_t)_mm256_extract_epi64(c, 1);
        const uint64_t l2 = (uint64_t)_mm256_extract_epi64(c, 2), l3 = (uint64_t)_mm256_extract_epi64(c, 3);
        words[k] = _pext_u64(l0, sel) | (_pext_u64(l1, sel) << 16) | (_pext_u64(l2, sel) << 32) | (_pext_u64(l3, sel) << 48);
        nbits[k] = bad;
    }
    invalid += inv;
    other += oth;
}

// 64 bases per step: the 2-bit codes of four neighbouring bytes are folded into one byte with two multiply-adds
// (c0 + 4 c1 per 16-bit lane, then lo + 16 hi per 32-bit lane) and the 16 result bytes leave through vpmovdb.
struct Avx512Step {
    __m128i words;  // 64 bases
    uint64_t bad;   // their N bits
};

__attribute__((target("avx512f,avx512bw,popcnt"), always_inline)) inline Avx512Step pack_step_avx512(const uint8_t *t, uint64_t &inv, uint64_t &oth)
{
    const __m512i m_df = _mm512_set1_epi8((char)0xDF), m_3 = _mm512_set1_epi8(3), m_1 = _mm512_set1_epi8(1);
    const __m512i cA = _mm512_set1_epi8('A'), cC = _mm512_set1_epi8('C'), cG = _mm512_set1_epi8('G'), cT = _mm512_set1_epi8('T'),
                  cN = _mm512_set1_epi8('N');
    const __m512i f1 = _mm512_set1_epi16(0x0401), f2 = _mm512_set1_epi32(0x00100001);
    const __m512i v = _mm512_loadu_si512(reinterpret_cast<const void *>(t));
    const __m512i u = _mm512_and_si512(v, m_df);
    const __m512i x = _mm512_and_si512(_mm512_srli_epi16(v, 1), m_3);                  // A 0, C 1, G 3, T 2
    __m512i c = _mm512_xor_si512(x, _mm512_and_si512(_mm512_srli_epi16(x, 1), m_1));   // A 0, C 1, G 2, T 3
    const __mmask64 ok = _mm512_cmpeq_epi8_mask(u, cA) | _mm512_cmpeq_epi8_mask(u, cC) | _mm512_cmpeq_epi8_mask(u, cG) |
                         _mm512_cmpeq_epi8_mask(u, cT);
    const uint64_t bad = ~(uint64_t)ok;
    if (bad) {
        c = _mm512_maskz_mov_epi8(ok, c);
        const uint64_t is_n = (uint64_t)_mm512_cmpeq_epi8_mask(u, cN);
        inv += (uint64_t)_mm_popcnt_u64(bad);
        oth += (uint64_t)_mm_popcnt_u64(bad & ~is_n);
    }
    const __m512i q = _mm512_madd_epi16(_mm512_maddubs_epi16(c, f1), f2);  // one byte of four codes per 32-bit lane
    return Avx512Step{_mm512_cvtepi32_epi8(q), bad};
}

// The output is written once and read by the copy engine (or a later pass), never by this core again: whole cache lines
// leave through streaming stores -- 512 bases = 16 words (two lines) + 16 N-bit words (one line) per round -- which spares
// the read-for-ownership of every line (a quarter of the packer's memory traffic; the packers are bound by that traffic).
__attribute__((target("avx512f,avx512bw,popcnt"))) void pack_words_avx512(const uint8_t *t, int64_t n_words, uint64_t *words,
                                                                          uint32_t *nbits, uint64_t &invalid, uint64_t &other)
{
    uint64_t inv = 0, oth = 0;
    int64_t k = 0;
    // head: up to the first word whose two output arrays both start a cache line
    while (k + 2 <= n_words && (((uintptr_t)(words + k) & 63u) != 0 || ((uintptr_t)(nbits + k) & 63u) != 0)) {
        const Avx512Step r = pack_step_avx512(t + 32 * k, inv, oth);
        _mm_storeu_si128(reinterpret_cast<__m128i *>(words + k), r.words);
        nbits[k] = (uint32_t)r.bad;
        nbits[k + 1] = (uint32_t)(r.bad >> 32);
        k += 2;
    }
    for (; k + 16 <= n_words; k += 16) {
        Avx512Step r[8];
#pragma GCC unroll 8
        for (int j = 0; j < 8; ++j) r[j] = pack_step_avx512(t + 32 * k + 64 * j, inv, oth);
        __m512i w0 = _mm512_castsi128_si512(r[0].words), w1 = _mm512_castsi128_si512(r[4].words);
        w0 = _mm512_inserti32x4(w0, r[1].words, 1); w0 = _mm512_inserti32x4(w0, r[2].words, 2); w0 = _mm512_inserti32x4(w0, r[3].words, 3);
        w1 = _mm512_inserti32x4(w1, r[5].words, 1); w1 = _mm512_inserti32x4(w1, r[6].words, 2); w1 = _mm512_inserti32x4(w1, r[7].words, 3);
        const __m512i nb = _mm512_set_epi64((long long)r[7].bad, (long long)r[6].bad, (long long)r[5].bad, (long long)r[4].bad,
                                            (long long)r[3].bad, (long long)r[2].bad, (long long)r[1].bad, (long long)r[0].bad);
        _mm512_stream_si512(reinterpret_cast<__m512i *>(words + k), w0);
        _mm512_stream_si512(reinterpret_cast<__m512i *>(words + k + 8), w1);
        _mm512_stream_si512(reinterpret_cast<__m512i *>(nbits + k), nb);
    }
    _mm_sfence();
    for (; k + 2 <= n_words; k += 2) {
        const Avx512Step r = pack_step_avx512(t + 32 * k, inv, oth);
        _mm_storeu_si128(reinterpret_cast<__m128i *>(words + k), r.words);
        nbits[k] = (uint32_t)r.bad;
        nbits[k + 1] = (uint32_t)(r.bad >> 32);
    }
    invalid += inv;
    other += oth;
    if (k < n_words) pack_words_avx2(t + 32 * k, n_words - k, words + k, nbits + k, invalid, other);
}

bool have_avx512()
{
    static const bool ok = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx2") &&
                           __builtin_cpu_supports("bmi2") && __builtin_cpu_supports("popcnt");
    return ok;
}

bool have_avx2()
{
    static const bool ok = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("bmi2") && __builtin_cpu_supports("popcnt");
    return ok;
}
#endif

}  // namespace

// 2: AVX-512BW, 1: AVX2 + BMI2, 0: scalar.  GB2_HOST_PACK_ISA=scalar|avx2 caps it (the tests walk all the paths the CPU has).
static int isa_level()
{
    int level = 0;
#if defined(__x86_64__)
    level = have_avx512() ? 2 : have_avx2() ? 1 : 0;
    if (const char *cap = getenv("GB2_HOST_PACK_ISA")) {
        if (!strcmp(cap, "scalar")) level = 0;
        else if (!strcmp(cap, "avx2")) level = level < 1 ? level : 1;
    }
#endif
    return level;
}

// `n_bases` bases at `text` -> ceil(n_bases / 32) words and N-bit words.
void gb2_host_pack_bases(const uint8_t *text, int64_t n_bases, uint64_t *words, uint32_t *nbits, uint64_t *n_invalid, uint64_t *n_other)
{
    uint64_t invalid = 0, other = 0;
    const int64_t full = n_bases >> 5;
    int64_t done = 0;
#if defined(__x86_64__)
    const int level = isa_level();
    if (level == 2) {
        pack_words_avx512(text, full, words, nbits, invalid, other);
        done = full;
    } else if (level == 1) {
        pack_words_avx2(text, full, words, nbits, invalid, other);
        done = full;
    }
#endif
    for (int64_t k = done; k < full; ++k) pack_word_scalar(text + 32 * k, 32, words[k], nbits[k], invalid, other);
    const int tail = (int)(n_bases & 31);
    if (tail) pack_word_scalar(text + 32 * full, tail, words[full], nbits[full], invalid, other);
    *n_invalid += invalid;
    *n_other += other;
}

// C ABI (include/grafimo_b200.h): the same packer for callers that keep their sequences 2-bit packed (format 1 input)
extern "C" int gb2_pack_sequence_host(const uint8_t *h_text, int64_t n_bases, uint64_t *h_words, uint32_t *h_nbits, uint64_t *h_counts)
{
    if (n_bases < 0 || (n_bases > 0 && (!h_text || !h_words || !h_nbits))) return 1;  // GB2_ERR_ARG
    uint64_t invalid = 0, other = 0;
    gb2_host_pack_bases(h_text, n_bases, h_words, h_nbits, &invalid, &other);
    if (h_counts) {
        h_counts[0] += invalid;
        h_counts[1] += other;
    }
    return 0;
}

int gb2_host_pack_simd() { return isa_level(); }
