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

__attribute__((target_clones("avx2", "default"))) void regenerate(uint32_t* s) {
#pragma GCC ivdep
    for (int j = 0; j < N - M; ++j) s[j] = s[j + M] ^ twist(s[j], s[j + 1]);
#pragma GCC ivdep
    for (int j = N - M; j < N - 1; ++j) s[j] = s[j + M - N] ^ twist(s[j], s[j + 1]);
    s[N - 1] = s[M - 1] ^ twist(s[N - 1], s[0]);
}

__attribute__((target_clones("avx2", "default"))) void temper(const uint32_t* s, uint32_t* out, int n) {
    for (int i = 0; i < n; ++i) {
        uint32_t y = s[i];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        out[i] = y;
    }
}

// x = k * 2^-53 < p  <=>  k < ceil(p * 2^53) on integers (k < 2^53 and the scaling by a power of two are exact), so the
// comparison runs on 32-bit lanes: (hi21, lo32) < (thi, tlo) lexicographically.
__attribute__((target_clones("avx2", "default"))) int64_t keep_from_pairs(const uint32_t* w, int64_t n, uint64_t thresh,
                                                                            uint8_t* out) {
    const uint32_t thi = (uint32_t)(thresh >> 32), tlo = (uint32_t)thresh;
    uint32_t ones = 0;  // n <= 65536 per call
    for (int64_t i = 0; i < n; ++i) {
        const uint32_t hi = w[2 * i] & 0x1fffffu, lo = w[2 * i + 1];
        const uint32_t k = (hi < thi) | ((hi == thi) & (lo < tlo));
        out[i] = (uint8_t)k;
        ones += k;
    }
    return ones;
}

__attribute__((target_clones("avx2", "default"))) int64_t keep_from_words(const uint32_t* w, int64_t n, uint32_t thresh,
                                                                            uint8_t* out) {
    uint32_t ones = 0;
    for (int64_t i = 0; i < n; ++i) {
        const uint32_t k = (w[i] & 0xffffffu) < thresh;
        out[i] = (uint8_t)k;
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
    uint32_t s[N];
    for (int i = 0; i < N; ++i) s[i] = (uint32_t)b->state[i];
    int64_t remaining = b->left - 1;  // words left in the current block
    int64_t pos = (int64_t)b->next;
    if (kind == 2) {  // skip n 32-bit draws (what a consumer with an as yet unknown probability will use up)
        int64_t need = n;
        while (need > 0) {
            if (remaining == 0) {
                regenerate(s);
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
    constexpr int64_t kChunkElems = 1 << 16;
    std::vector<uint32_t> words((size_t)kChunkElems * 2 + N);
    uint32_t tempered[N];
    int64_t ones = 0;
    const uint64_t t53 = threshold(p, 53);
    const uint64_t t24 = threshold((double)(float)p, 24);  // kind 1 compares against the float32 probability
    for (int64_t e0 = 0; e0 < n; e0 += kChunkElems) {
        const int64_t ne = (n - e0) < kChunkElems ? (n - e0) : kChunkElems;
        const int64_t need = ne * per;
        int64_t got = 0;
        while (got < need) {
            if (remaining == 0) {
                regenerate(s);
                pos = 0;
                remaining = N;
            }
            const int64_t take = (need - got) < remaining ? (need - got) : remaining;
            temper(s + pos, tempered, (int)take);
            memcpy(words.data() + got, tempered, (size_t)take * sizeof(uint32_t));
            got += take;
            pos += take;
            remaining -= take;
        }
        ones += kind == 0 ? keep_from_pairs(words.data(), ne, t53, out + e0)
                          : keep_from_words(words.data(), ne, (uint32_t)t24, out + e0);
    }
    for (int i = 0; i < N; ++i) b->state[i] = s[i];
    b->left = (int32_t)(remaining + 1);
    b->next = (uint64_t)pos;
    return ones;
}
