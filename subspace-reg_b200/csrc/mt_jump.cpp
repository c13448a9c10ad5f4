// Jump-ahead for PyTorch's CPU generator (mt19937), host side.
//
// The keep-masks of a train-mode forward consume 5e7-9e7 consecutive 32-bit words of the generator (reference
// models/resnet_language.py:292-299, 311-325 through F.dropout / Bernoulli.sample).  The recurrence
//     x[k+624] = x[k+397] ^ twist(x[k], x[k+1])
// is serial, so producing those words on the GPU needs many walkers that each START somewhere inside the stream: the
// window (x[n] .. x[n+623]) at an arbitrary distance n.  The state transition is linear over GF(2) with a primitive
// characteristic polynomial phi of degree 19937, hence with  t^n mod phi(t) = sum_k g_k t^k  (k < 19937)
//     window(n) = XOR over { k : g_k = 1 } of window(k),
// i.e. a window 2^18 * w words ahead is an XOR of ~10^4 of the first 19937 windows - 6 M word operations, done by one
// CTA per walker (csrc/mask.cu) or by sr_host_mt_advance below (which keeps torch's HOST generator in step without
// drawing anything).  This file computes phi (Berlekamp-Massey on the generator's own output, so nothing is taken on
// faith), the polynomials g^(wJ) for J = SR_MT_JUMP_WORDS, and the host-side advance.
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <mutex>
#include <thread>
#include <vector>
#include "../../include/srb200.h"

namespace {
constexpr int N = 624, M = 397;
constexpr int DEG = 19937;
constexpr int PW = 312;                 // 64-bit words of a polynomial of degree < 19968
constexpr int64_t J = SR_MT_JUMP_WORDS;

struct Blob {  // at::CPUGeneratorImplStateLegacy (see host_rng.cpp)
    uint64_t seed;
    int32_t left;
    int32_t seeded;
    uint64_t next;
    uint64_t state[N];
};

inline uint32_t twist(uint32_t u, uint32_t v) {
    return (((u & 0x80000000u) | (v & 0x7fffffffu)) >> 1) ^ ((v & 1u) ? 0x9908b0dfu : 0u);
}

// next window: out[0..623] = x[k+624 .. k+1247] from in[0..623] = x[k .. k+623] (only the top bit of in[0] is used)
// (in and out never alias; the second loop reads out[] 227 words behind the one it writes: safe for any vector width <= 227)
__attribute__((target_clones("arch=skylake-avx512", "avx2", "default"))) void next_window(const uint32_t* __restrict__ in,
                                                                                         uint32_t* __restrict__ out) {
#pragma GCC ivdep
    for (int j = 0; j < N - M; ++j) out[j] = in[j + M] ^ twist(in[j], in[j + 1]);
#pragma GCC ivdep
    for (int j = N - M; j < N - 1; ++j) out[j] = out[j + M - N] ^ twist(in[j], in[j + 1]);
    out[N - 1] = out[M - 1] ^ twist(in[N - 1], out[0]);
}

__attribute__((target_clones("arch=skylake-avx512", "avx2", "default"))) void xor_window(uint32_t* __restrict__ acc,
                                                                                        const uint32_t* __restrict__ src) {
    for (int j = 0; j < N; ++j) acc[j] ^= src[j];
}

using Poly = std::vector<uint64_t>;   // PW words

inline int get_bit(const uint64_t* w, int64_t i) { return (int)((w[i >> 6] >> (i & 63)) & 1u); }

// ---- phi: minimal polynomial of the bit sequence b_n = lsb(x[n]) (Berlekamp-Massey over GF(2)) ----
// The transition matrix's characteristic polynomial is irreducible, so every non-zero output bit sequence has it as its
// minimal polynomial.  Returns phi as PW words (bit i = coefficient of t^i), or an empty vector if the degree is not 19937.
Poly compute_phi() {
    const int64_t NB = 2 * (int64_t)DEG + 64;
    // generator words from an arbitrary non-zero state (init_genrand(5489))
    std::vector<uint32_t> st(N), nx(N);
    st[0] = 5489u;
    for (int i = 1; i < N; ++i) st[i] = 1812433253u * (st[i - 1] ^ (st[i - 1] >> 30)) + (uint32_t)i;
    const int64_t sw = (NB + 63) / 64 + 2;
    std::vector<uint64_t> rev((size_t)sw, 0);   // rev bit j = s[NB-1-j]: a window of the PAST is then a forward read
    {
        int64_t n = 0;
        while (n < NB) {
            next_window(st.data(), nx.data());
            st.swap(nx);
            for (int i = 0; i < N && n < NB; ++i, ++n)
                if (st[i] & 1u) { const int64_t j = NB - 1 - n; rev[(size_t)(j >> 6)] |= 1ull << (j & 63); }
        }
    }
    const int cw = (DEG + 64) / 64 + 1;   // words of the connection polynomials
    std::vector<uint64_t> C((size_t)cw, 0), B((size_t)cw, 0), T((size_t)cw, 0);
    C[0] = 1; B[0] = 1;
    int64_t L = 0, m = 1;
    for (int64_t n = 0; n < NB; ++n) {
        // d = sum_{i=0..L} c_i s[n-i] = <C, rev >> (NB-1-n)>
        const int64_t off = NB - 1 - n;
        const int64_t wo = off >> 6;
        const int bo = (int)(off & 63);
        const int nwords = (int)(L >> 6) + 1;
        uint64_t acc = 0;
        for (int i = 0; i < nwords; ++i) {
            uint64_t v = rev[(size_t)(wo + i)] >> bo;
            if (bo) v |= rev[(size_t)(wo + i + 1)] << (64 - bo);
            acc ^= v & C[(size_t)i];
        }
        const int d = __builtin_parityll(acc);
        if (!d) { ++m; continue; }
        const bool grow = 2 * L <= n;
        if (grow) T = C;
        // C ^= B << m
        {
            const int ws = (int)(m >> 6), bs = (int)(m & 63);
            for (int i = cw - 1; i >= ws; --i) {
                uint64_t v = B[(size_t)(i - ws)] << bs;
                if (bs && i - ws - 1 >= 0) v |= B[(size_t)(i - ws - 1)] >> (64 - bs);
                C[(size_t)i] ^= v;
            }
        }
        if (grow) { L = n + 1 - L; B = T; m = 1; } else { ++m; }
    }
    if (L != DEG) return Poly();
    Poly phi((size_t)PW, 0);     // phi(t) = t^L C(1/t): coefficient of t^i is c_{L-i}
    for (int i = 0; i <= DEG; ++i)
        if (get_bit(C.data(), DEG - i)) phi[(size_t)(i >> 6)] |= 1ull << (i & 63);
    return phi;
}

// ---- arithmetic mod phi ----
struct Field {
    Poly phi;
    std::vector<uint64_t> shifted;   // [64][PW+1]: phi << b, b = 0..63 (for aligned XORs in the reduction)
    bool ok = false;
    void init() {
        phi = compute_phi();
        if (phi.empty()) return;
        shifted.assign((size_t)64 * (PW + 1), 0);
        for (int b = 0; b < 64; ++b) {
            uint64_t* d = &shifted[(size_t)b * (PW + 1)];
            for (int i = 0; i < PW; ++i) {
                d[i] |= phi[(size_t)i] << b;
                if (b) d[i + 1] |= phi[(size_t)i] >> (64 - b);
            }
        }
        ok = true;
    }
    // r = a * b mod phi
    void mulmod(const Poly& a, const Poly& b, Poly& r) const {
        std::vector<uint64_t> prod((size_t)2 * PW + 2, 0);
        for (int i = 0; i < PW; ++i) {
            uint64_t aw = a[(size_t)i];
            while (aw) {
                const int bit = __builtin_ctzll(aw);
                aw &= aw - 1;
                uint64_t* d = &prod[(size_t)i];
                if (bit == 0) {
                    for (int k = 0; k < PW; ++k) d[k] ^= b[(size_t)k];
                } else {
                    uint64_t carry = 0;
                    for (int k = 0; k < PW; ++k) {
                        const uint64_t v = b[(size_t)k];
                        d[k] ^= (v << bit) | carry;
                        carry = v >> (64 - bit);
                    }
                    d[PW] ^= carry;
                }
            }
        }
        // reduce: clear bits 2*DEG .. DEG from the top
        for (int64_t i = 2 * (int64_t)PW * 64 - 1; i >= DEG; --i) {
            if (!get_bit(prod.data(), i)) continue;
            const int64_t off = i - DEG;
            const uint64_t* s = &shifted[(size_t)(off & 63) * (PW + 1)];
            uint64_t* d = &prod[(size_t)(off >> 6)];
            for (int k = 0; k <= PW; ++k) d[k] ^= s[k];
        }
        r.assign(prod.begin(), prod.begin() + PW);
    }
};

struct Table {
    std::mutex mu;
    Field f;
    bool tried = false;
    Poly g1;                      // t^J mod phi
    std::vector<Poly> g;          // g[w-1] = t^(wJ) mod phi
    bool ensure(int n) {
        std::lock_guard<std::mutex> lk(mu);
        if (!tried) {
            tried = true;
            f.init();
            if (f.ok) {
                Poly t((size_t)PW, 0), r;
                t[0] = 2;                                   // t
                int64_t e = 1;
                while (e < J) { f.mulmod(t, t, r); t.swap(r); e *= 2; }   // J is a power of two
                g1 = t;
            }
        }
        if (!f.ok) return false;
        // t^((m+j+1)J) = t^(mJ) * t^((j+1)J): with m entries known the next m are independent products - computed by a few
        // threads (one product is ~4 ms; a sweep's forwards need ~330 entries, i.e. 1.3 s if done one after the other)
        while ((int)g.size() < n) {
            if (g.empty()) { g.push_back(g1); continue; }
            const int m = (int)g.size();
            const int batch = std::min(m, n - m);
            std::vector<Poly> fresh((size_t)batch);
            int workers = (int)std::thread::hardware_concurrency() - 1;
            workers = std::max(1, std::min(std::min(workers, 8), batch));
            std::atomic<int> next(0);
            auto work = [&] {
                for (int j = next.fetch_add(1); j < batch; j = next.fetch_add(1)) f.mulmod(g[(size_t)m - 1], g[(size_t)j], fresh[(size_t)j]);
            };
            std::vector<std::thread> pool;
            for (int k = 1; k < workers; ++k) pool.emplace_back(work);
            work();
            for (auto& t : pool) t.join();
            for (auto& r : fresh) g.push_back(std::move(r));
        }
        return true;
    }
};
Table& table() {
    static Table t;
    return t;
}

// window(n) for n = w*J from the window at 0, by polynomial evaluation.  win: in/out [624]; only the top bit of win[0] is
// meaningful on input and on output.
void jump_window(uint32_t* win, const uint32_t* poly32) {
    std::vector<uint32_t> y((size_t)DEG + N + N);
    memcpy(y.data(), win, sizeof(uint32_t) * N);
    for (int64_t k = N; k < (int64_t)y.size(); k += N) {
        const int64_t take = (int64_t)y.size() - k < N ? (int64_t)y.size() - k : N;
        uint32_t tmp[N];
        next_window(&y[(size_t)(k - N)], tmp);
        memcpy(&y[(size_t)k], tmp, sizeof(uint32_t) * (size_t)take);
    }
    uint32_t acc[N];
    memset(acc, 0, sizeof(acc));
    for (int wi = 0; wi < N; ++wi) {
        uint32_t gw = poly32[wi];
        while (gw) {
            const int b = __builtin_ctz(gw);
            gw &= gw - 1;
            const int64_t k = (int64_t)wi * 32 + b;
            if (k >= DEG) break;
            xor_window(acc, &y[(size_t)k]);
        }
    }
    memcpy(win, acc, sizeof(acc));
}
}  // namespace

extern "C" int64_t sr_mt_jump_table_bytes(int32_t n_polys) { return n_polys < 0 ? -1 : (int64_t)n_polys * N * 4; }

extern "C" int32_t sr_mt_jump_table(uint32_t* table_host, int32_t n_polys) {
    if (n_polys < 0 || (n_polys > 0 && !table_host)) return SR_E_ARG;
    Table& t = table();
    if (!t.ensure(n_polys)) return SR_E_ARG;
    std::lock_guard<std::mutex> lk(t.mu);
    for (int w = 0; w < n_polys; ++w) memcpy(table_host + (size_t)w * N, t.g[(size_t)w].data(), (size_t)N * 4);
    return SR_OK;
}

// Advance the generator in `state_blob` by n_words 32-bit draws without producing them; the resulting blob is byte-identical
// to the one torch holds after drawing that many words one by one.
extern "C" int32_t sr_host_mt_advance(void* state_blob, int64_t blob_bytes, int64_t n_words, const uint32_t* table_host,
                                      int32_t n_polys) {
    if (!state_blob || blob_bytes < (int64_t)sizeof(Blob) || n_words < 0) return SR_E_ARG;
    Blob* b = static_cast<Blob*>(state_blob);
    if (!b->seeded || b->left < 1 || b->left > N || b->next > (uint64_t)N) return SR_E_ARG;
    if (n_words == 0) return SR_OK;
    const int64_t remaining = b->left - 1;
    const int64_t pos = remaining == 0 ? N : (int64_t)b->next;   // next output = x[pos], x[0..623] = the current block
    const int64_t c = pos + n_words;                             // words consumed counted from the block's start
    if (c <= N) {
        b->next = (uint64_t)c;
        b->left = (int32_t)(N - c + 1);
        return SR_OK;
    }
    const int64_t q = (c - 1) / N;          // regenerations torch would have done
    const int64_t next = c - q * N;         // 1..624
    const int64_t target = q * N - 1;       // window(target) = (top bit of x[qN-1], x[qN] .. x[qN+622])
    uint32_t win[N], nx[N];
    for (int i = 0; i < N; ++i) win[i] = (uint32_t)b->state[i];
    int64_t at = 0;
    const int64_t a = target / J;
    if (a >= 1) {
        if (a > n_polys || !table_host) return SR_E_ARG;
        jump_window(win, table_host + (size_t)(a - 1) * N);
        at = a * J;
    }
    while (target - at >= N) {
        next_window(win, nx);
        memcpy(win, nx, sizeof(win));
        at += N;
    }
    const int rem = (int)(target - at);     // 0..623
    next_window(win, nx);
    // the block x[qN .. qN+623] = words rem+1 .. rem+624 of the concatenation [win | nx]
    for (int i = 0; i < N; ++i) {
        const int k = rem + 1 + i;
        b->state[i] = k < N ? win[k] : nx[k - N];
    }
    b->next = (uint64_t)next;
    b->left = (int32_t)(N - next + 1);
    return SR_OK;
}
