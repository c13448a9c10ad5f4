// Host-side accelerator for the ONE piece of the path that must follow PyTorch's CPU generator bit for bit:
// the dropout / DropBlock keep-masks of the train-mode epoch (reference models/resnet_language.py:292-299, 311-325).
// The reference draws them with torch's CPU mt19937 (F.dropout -> bernoulli_(1-p), Bernoulli(gamma).sample -> bernoulli_(p
// tensor)), one serial draw per element (~10 ns each, ~5e7 elements per forward).  This file replays exactly that stream
// from torch's generator state blob (torch.get_rng_state()), vectorised, and writes the advanced state back:
//   kind 0  bernoulli_(double p):  r = (hi << 32 | lo) two 32-bit draws,  keep = (r & (2^53-1)) * 2^-53 < p
//   kind 1  bernoulli_(float p tensor): one 32-bit draw,                  keep = float((w & (2^24-1)) * 2^-24) < p
// The Python side verifies the replay against torch itself on a sample before trusting it (other torch builds may route
// bernoulli_ through MKL), and otherwise calls torch.
#include <stdint.h>
#include <string.h>
#include <immintrin.h>
#include <stdlib.h>
#include <atomic>
#include <thread>
#include <vector>
#include "../../include/srb200.h"

namespace {
constexpr int N = 624, M = 397;

struct Blob {  // at::CPUGeneratorImplStateLegacy (little endian, 5056 bytes with the trailing float-normal cache)
    uint64_t seed;
    int32_t left;
    int32_t seeded;
    uint64_t next;
    uint64_t state[N];
};

inline uint32_t twist(uint32_t u, uint32_t v) {
    return (((u & 0x80000000u) | (v & 0x7fffffffu)) >> 1) ^ ((v & 1u) ? 0x9908b0dfu : 0u);
}

__attribute__((target_clones("arch=skylake-avx512", "avx2", "default"))) void regenerate(uint32_t* s) {
#pragma GCC ivdep
    for (int j = 0; j < N - M; ++j) s[j] = s[j + M] ^ twist(s[j], s[j + 1]);
#pragma GCC ivdep
    for (int j = N - M; j < N - 1; ++j) s[j] = s[j + M - N] ^ twist(s[j], s[j + 1]);
    s[N - 1] = s[M - 1] ^ twist(s[N - 1], s[0]);
}

__attribute__((target_clones("arch=skylake-avx512", "avx2", "default"))) void temper(const uint32_t* s, uint32_t* out, int n) {
    for (int i = 0; i < n; ++i) {
        uint32_t y = s[i];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        out[i] = y;
    }
}

// x = k * 2^-53 < p  <=>  k < ceil(p * 2^53) on integers (k < 2^53 and the scaling by a power of two are exact).  Two
// consecutive 32-bit draws (hi first) read as one little-endian 64-bit word hold hi in the LOW half: rotate by 32.
__attribute__((target_clones("arch=skylake-avx512", "avx2", "default"))) int64_t keep_from_pairs(
    const uint32_t* w, int64_t n, uint64_t thresh, uint8_t* out) {
    const uint64_t* w64 = reinterpret_cast<const uint64_t*>(w);
    const int64_t t = (int64_t)thresh;   // <= 2^53: signed compares are safe
    uint32_t ones = 0;                   // n <= 65536 per call
    for (int64_t i = 0; i < n; ++i) {
        const uint64_t v = w64[i];
        const int64_t r = (int64_t)(((v << 32) | (v >> 32)) & ((1ull << 53) - 1));
        const uint8_t k = r < t;
        out[i] = k;
        ones += k;
    }
    return ones;
}

__attribute__((target_clones("arch=skylake-avx512", "avx2", "default"))) int64_t keep_from_words(
    const uint32_t* w, int64_t n, uint32_t thresh, uint8_t* out) {
    uint32_t ones = 0;
    for (int64_t i = 0; i < n; ++i) {
        const uint32_t k = (w[i] & 0xffffffu) < thresh;
        out[i] = (uint8_t)k;
        ones += k;
    }
    return ones;
}

// ---- hand-written AVX-512 versions (runtime-dispatched; the auto-vectorised clones above are the fallback) ----
#define SRB_AVX512 __attribute__((target("avx512f,avx512bw,avx512vl,avx512dq")))

bool have_avx512() {
    static const bool ok = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") &&
                           __builtin_cpu_supports("avx512vl") && __builtin_cpu_supports("avx512dq");
    return ok;
}

// One state transition, 16 lanes at a time.  Ascending j: s[j+1 .. j+16] are read before s[j .. j+15] are written, and the
// already regenerated operand of the second loop sits 227 >= 16 words behind.
SRB_AVX512 void regenerate512(uint32_t* s) {
    const __m512i up = _mm512_set1_epi32((int)0x80000000u), lo = _mm512_set1_epi32(0x7fffffff), one = _mm512_set1_epi32(1),
                  mag = _mm512_set1_epi32((int)0x9908b0dfu);
#define SRB_MT_STEP(j, src)                                                                                         \
    do {                                                                                                            \
        const __m512i a_ = _mm512_loadu_si512(s + (j)), b_ = _mm512_loadu_si512(s + (j) + 1);                       \
        const __m512i c_ = _mm512_loadu_si512(s + (src));                                                           \
        const __m512i y_ = _mm512_or_si512(_mm512_and_si512(a_, up), _mm512_and_si512(b_, lo));                     \
        __m512i r_ = _mm512_xor_si512(c_, _mm512_srli_epi32(y_, 1));                                                \
        r_ = _mm512_mask_xor_epi32(r_, _mm512_test_epi32_mask(b_, one), r_, mag);                                   \
        _mm512_storeu_si512(s + (j), r_);                                                                           \
    } while (0)
    int j = 0;
    for (; j + 16 <= N - M; j += 16) SRB_MT_STEP(j, j + M);
    for (; j < N - M; ++j) s[j] = s[j + M] ^ twist(s[j], s[j + 1]);
    for (; j + 16 <= N - 1; j += 16) SRB_MT_STEP(j, j + M - N);
    for (; j < N - 1; ++j) s[j] = s[j + M - N] ^ twist(s[j], s[j + 1]);
    s[N - 1] = s[M - 1] ^ twist(s[N - 1], s[0]);
#undef SRB_MT_STEP
}

SRB_AVX512 inline __m512i temper512(__m512i y) {
    y = _mm512_xor_si512(y, _mm512_srli_epi32(y, 11));
    y = _mm512_xor_si512(y, _mm512_and_si512(_mm512_slli_epi32(y, 7), _mm512_set1_epi32((int)0x9d2c5680u)));
    y = _mm512_xor_si512(y, _mm512_and_si512(_mm512_slli_epi32(y, 15), _mm512_set1_epi32((int)0xefc60000u)));
    return _mm512_xor_si512(y, _mm512_srli_epi32(y, 18));
}

// raw (untempered) words -> keep bytes, tempering in registers.  kind 0: 2 words per element, kind 1: one.
SRB_AVX512 int64_t finish512(const uint32_t* w, int64_t ne, int kind, uint64_t t53, uint32_t t24, uint8_t* out) {
    int64_t ones = 0, i = 0;
    if (kind == 0) {
        const __m512i m53 = _mm512_set1_epi64((long long)((1ull << 53) - 1)), th = _mm512_set1_epi64((long long)t53);
        for (; i + 16 <= ne; i += 16) {   // 32 words -> 16 elements
            const __m512i a = temper512(_mm512_loadu_si512(w + 2 * i)), b = temper512(_mm512_loadu_si512(w + 2 * i + 16));
            const __mmask8 ka = _mm512_cmplt_epu64_mask(_mm512_and_si512(_mm512_rol_epi64(a, 32), m53), th);
            const __mmask8 kb = _mm512_cmplt_epu64_mask(_mm512_and_si512(_mm512_rol_epi64(b, 32), m53), th);
            const __mmask16 k = (__mmask16)((unsigned)ka | ((unsigned)kb << 8));
            _mm_storeu_si128(reinterpret_cast<__m128i*>(out + i), _mm_maskz_set1_epi8(k, 1));
            ones += __builtin_popcount((unsigned)k);
        }
    } else {
        const __m512i m24 = _mm512_set1_epi32(0xffffff), th = _mm512_set1_epi32((int)t24);
        for (; i + 16 <= ne; i += 16) {
            const __m512i a = temper512(_mm512_loadu_si512(w + i));
            const __mmask16 k = _mm512_cmplt_epu32_mask(_mm512_and_si512(a, m24), th);
            _mm_storeu_si128(reinterpret_cast<__m128i*>(out + i), _mm_maskz_set1_epi8(k, 1));
            ones += __builtin_popcount((unsigned)k);
        }
    }
    for (; i < ne; ++i) {   // tail, scalar
        uint32_t y[2];
        const int per = kind == 0 ? 2 : 1;
        for (int q = 0; q < per; ++q) {
            uint32_t v = w[per * i + q];
            v ^= (v >> 11);
            v ^= (v << 7) & 0x9d2c5680u;
            v ^= (v << 15) & 0xefc60000u;
            v ^= (v >> 18);
            y[q] = v;
        }
        uint8_t k;
        if (kind == 0) k = ((((uint64_t)(y[0] & 0x1fffffu)) << 32) | y[1]) < t53;
        else k = (y[0] & 0xffffffu) < t24;
        out[i] = k;
        ones += k;
    }
    return ones;
}

uint64_t threshold(double p, int bits) {  // ceil(p * 2^bits), clamped to [0, 2^bits]
    if (!(p > 0.0)) return 0;
    if (p >= 1.0) return 1ull << bits;
    const double t = __builtin_ceil(__builtin_ldexp(p, bits));
    return (uint64_t)t;
}
}  // namespace

// state_blob: the bytes of torch.get_rng_state() (updated in place).  Returns the number of ones written, or -1.
extern "C" int64_t sr_host_bernoulli(void* state_blob, int64_t blob_bytes, int32_t kind, double p, int64_t n, uint8_t* out) {
    if (!state_blob || n < 0 || blob_bytes < (int64_t)sizeof(Blob) || kind < 0 || kind > 2 || (kind != 2 && !out)) return -1;
    Blob* b = static_cast<Blob*>(state_blob);
    if (!b->seeded || b->left < 1 || b->left > N || b->next > (uint64_t)N) return -1;
    alignas(64) uint32_t s[N + 16];   // + slack: the 16-lane loads of regenerate512 stay inside the array
    for (int i = 0; i < N; ++i) s[i] = (uint32_t)b->state[i];
    for (int i = N; i < N + 16; ++i) s[i] = 0;
    const bool v512 = have_avx512() && !getenv("SRB_RNG_NO_AVX512");
    auto regen = [&]() {
        if (v512) regenerate512(s);
        else regenerate(s);
    };
    int64_t remaining = b->left - 1;  // words left in the current block
    int64_t pos = (int64_t)b->next;
    if (kind == 2) {  // skip n 32-bit draws (what a consumer with an as yet unknown probability will use up)
        int64_t need = n;
        while (need > 0) {
            if (remaining == 0) {
                regen();
                pos = 0;
                remaining = N;
            }
            const int64_t take = need < remaining ? need : remaining;
            pos += take;
            remaining -= take;
            need -= take;
        }
        for (int i = 0; i < N; ++i) b->state[i] = s[i];
        b->left = (int32_t)(remaining + 1);
        b->next = (uint64_t)pos;
        return 0;
    }
    const int per = kind == 0 ? 2 : 1;
    const uint64_t t53 = threshold(p, 53);
    const uint64_t t24 = threshold((double)(float)p, 24);  // kind 1 compares against the float32 probability
    int64_t ones = 0;

    // Raw (untempered) generator words in stream order: the state recurrence is the only serial part of the job.
    auto raw_fill = [&](uint32_t* dst, int64_t nwords) {
        int64_t got = 0;
        while (got < nwords) {
            if (remaining == 0) {
                regen();
                pos = 0;
                remaining = N;
            }
            const int64_t take = (nwords - got) < remaining ? (nwords - got) : remaining;
            memcpy(dst + got, s + pos, (size_t)take * sizeof(uint32_t));
            got += take;
            pos += take;
            remaining -= take;
        }
    };
    // Tempering + threshold compare of `ne` elements whose raw words start at w (in place), 64 Ki elements at a time.
    auto finish = [&](uint32_t* w, int64_t ne, uint8_t* dst) -> int64_t {
        if (v512) return finish512(w, ne, kind, t53, (uint32_t)t24, dst);
        int64_t cnt = 0;
        constexpr int64_t kBlk = 1 << 16;
        for (int64_t e0 = 0; e0 < ne; e0 += kBlk) {
            const int64_t m = (ne - e0) < kBlk ? (ne - e0) : kBlk;
            temper(w + e0 * per, w + e0 * per, (int)(m * per));
            cnt += kind == 0 ? keep_from_pairs(w + e0 * per, m, t53, dst + e0) : keep_from_words(w + e0 * per, m, (uint32_t)t24, dst + e0);
        }
        return cnt;
    };

    int workers = 4;
    if (const char* e = getenv("SRB_RNG_THREADS")) workers = atoi(e);
    const int hw = (int)std::thread::hardware_concurrency();
    if (hw > 0 && workers > hw - 2) workers = hw - 2;
    if (workers > 12) workers = 12;
    constexpr int64_t kSlabElems = 1 << 16;   // elements per slab: 512 KB of raw words for kind 0 (stays in L2)
    if (workers < 2 || n < 16 * kSlabElems) {
        // small draw: one thread
        std::vector<uint64_t> words64((size_t)kSlabElems);   // 8-byte aligned: pairs are read as 64-bit words
        uint32_t* const words = reinterpret_cast<uint32_t*>(words64.data());
        for (int64_t e0 = 0; e0 < n; e0 += kSlabElems) {
            const int64_t ne = (n - e0) < kSlabElems ? (n - e0) : kSlabElems;
            raw_fill(words, ne * per);
            ones += finish(words, ne, out + e0);
        }
    } else {
        // large draw: the state recurrence is serial but cheap (a fifth of the arithmetic), and shipping raw words between
        // cores costs as much as computing them.  So EVERY worker walks the whole recurrence on its own copy of the state
        // and tempers + compares only every `workers`-th slab; nothing but the output bytes leaves a core.
        const int64_t nslabs = (n + kSlabElems - 1) / kSlabElems;
        struct Walker {
            alignas(64) uint32_t st[N + 16];
            int64_t pos, remaining;
        };
        std::vector<int64_t> counts((size_t)workers, 0);
        std::vector<Walker> walkers((size_t)workers);
        auto walk = [&](int k) {
            Walker& wk = walkers[(size_t)k];
            memcpy(wk.st, s, sizeof(wk.st));
            wk.pos = pos;
            wk.remaining = remaining;
            std::vector<uint64_t> words64((size_t)kSlabElems);
            uint32_t* const words = reinterpret_cast<uint32_t*>(words64.data());
            int64_t cnt = 0;
            for (int64_t i = 0; i < nslabs; ++i) {
                const int64_t e0 = i * kSlabElems;
                const int64_t ne = (n - e0) < kSlabElems ? (n - e0) : kSlabElems;
                const bool mine = (i % workers) == k;
                int64_t need = ne * per, got = 0;
                while (need > 0) {
                    if (wk.remaining == 0) {
                        if (v512) regenerate512(wk.st);
                        else regenerate(wk.st);
                        wk.pos = 0;
                        wk.remaining = N;
                    }
                    const int64_t take = need < wk.remaining ? need : wk.remaining;
                    if (mine) memcpy(words + got, wk.st + wk.pos, (size_t)take * sizeof(uint32_t));
                    got += take;
                    wk.pos += take;
                    wk.remaining -= take;
                    need -= take;
                }
                if (mine) cnt += finish(words, ne, out + e0);
            }
            counts[(size_t)k] = cnt;
        };
        std::vector<std::thread> pool;
        for (int k = 1; k < workers; ++k) pool.emplace_back(walk, k);
        walk(0);
        for (auto& t : pool) t.join();
        for (int k = 0; k < workers; ++k) ones += counts[(size_t)k];
        memcpy(s, walkers[0].st, sizeof(uint32_t) * N);   // every walker ends in the same state
        pos = walkers[0].pos;
        remaining = walkers[0].remaining;
    }
    for (int i = 0; i < N; ++i) b->state[i] = s[i];
    b->left = (int32_t)(remaining + 1);
    b->next = (uint64_t)pos;
    return ones;
}

extern "C" int64_t sr_host_dropblock(const uint8_t* seeds, int64_t planes, int32_t hs, int32_t ws, int32_t bs, uint8_t* keep) {
    if (!seeds || !keep || planes < 0 || hs < 1 || ws < 1 || bs < 1) return -1;
    const int ho = hs + bs - 1, wo = ws + bs - 1;
    const int64_t plane_in = (int64_t)hs * ws, plane_out = (int64_t)ho * wo;
    // one pass per plane: a plane without seeds (the common case: gamma is a few percent of a 6x6 / 3x3 seed map) is a
    // single memset; planes are independent, so large masks are split over a few threads
    auto run = [&](int64_t p0, int64_t p1) -> int64_t {
        int64_t kept = 0;
        for (int64_t pl = p0; pl < p1; ++pl) {
            const uint8_t* sp = seeds + pl * plane_in;
            uint8_t* kp = keep + pl * plane_out;
            memset(kp, 1, (size_t)plane_out);
            const uint8_t* q = static_cast<const uint8_t*>(memchr(sp, 1, (size_t)plane_in));
            if (!q) {
                kept += plane_out;
                continue;
            }
            for (; q; q = static_cast<const uint8_t*>(memchr(q + 1, 1, (size_t)(sp + plane_in - (q + 1))))) {
                const int r = (int)(q - sp);
                const int y = r / ws, x = r - y * ws;
                for (int i = 0; i < bs; ++i) memset(kp + (int64_t)(y + i) * wo + x, 0, (size_t)bs);
                if (q + 1 >= sp + plane_in) break;
            }
            int64_t ones = 0;
            for (int64_t k = 0; k < plane_out; ++k) ones += kp[k];
            kept += ones;
        }
        return kept;
    };
    int workers = 4;
    if (const char* e = getenv("SRB_RNG_THREADS")) workers = atoi(e);
    const int hw = (int)std::thread::hardware_concurrency();
    if (hw > 0 && workers > hw - 1) workers = hw - 1;
    if (workers > 8) workers = 8;
    if (workers < 2 || planes * plane_out < (1 << 20)) return run(0, planes);
    std::vector<int64_t> part((size_t)workers, 0);
    std::vector<std::thread> pool;
    const int64_t share = (planes + workers - 1) / workers;
    for (int k = 1; k < workers; ++k) {
        const int64_t p0 = k * share, p1 = p0 + share < planes ? p0 + share : planes;
        if (p0 < p1) pool.emplace_back([&, k, p0, p1] { part[(size_t)k] = run(p0, p1); });
    }
    part[0] = run(0, share < planes ? share : planes);
    for (auto& t : pool) t.join();
    int64_t kept = 0;
    for (int k = 0; k < workers; ++k) kept += part[(size_t)k];
    return kept;
}
