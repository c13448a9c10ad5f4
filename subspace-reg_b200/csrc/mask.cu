// Keep-masks of the train-mode forward drawn ON THE DEVICE, bit-identical to PyTorch's CPU generator.
//
// The reference draws the dropout / DropBlock masks of epoch 1 with torch's CPU mt19937 (models/resnet_language.py:292-299,
// 311-325): 5e7-9e7 consecutive 32-bit words per forward.  Round 1 replayed that stream on the host (host_rng.cpp) and
// shipped 43 MB of masks per forward over PCIe; with 8 ranks (or several runs per GPU) the host cores ran out.  Here the
// device produces the same words from the HOST generator's state (2.5 KB, passed as a kernel argument):
//   1. mt_base_kernel   one CTA: the first 19937 + 624 words after the generator's position (y[k])
//   2. mt_jump_kernel   one CTA per walker w: the 624-word window w * 2^18 words further on, as the XOR of the windows
//                       y[k .. k+623] selected by the bits of t^(w 2^18) mod phi (mt_jump.cpp computes the table)
//   3. mt_mask_kernel   one CTA per walker: walks its 2^18 words (three dependent 227-wide steps per 624 words), tempers,
//                       compares with the Bernoulli threshold exactly like at::bernoulli_ does, writes keep bytes (NCHW order)
// and sr_host_mt_advance moves the host generator past the words without drawing them.  DropBlock's block dilation
// (_compute_block_mask) and its numel / kept scale run on the device too (dropblock_kernel), so nothing is read back.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include "common.h"

namespace {
using namespace srb;

constexpr int N = 624, M = 397;
constexpr int DEG = 19937;
constexpr int kY = DEG + N;              // words of the base sequence the jump needs: y[k + j], k < 19937, j < 624
constexpr int kYPad = kY + 64;
constexpr int64_t J = SR_MT_JUMP_WORDS;
constexpr int kT = 320;                  // threads per CTA: 312 word pairs per 624-word frame
constexpr int kMaxRegions = 8;
// mt_jump_kernel: base sequence, polynomial, popcount prefix, compacted coefficient list (16-bit positions)
constexpr int kListCap = 12288;   // set coefficients of a jump polynomial: 9970 +- 71 (binomial); more than this traps
constexpr int kJumpSmem = (kYPad + 3 + 2 * N) * 4 + (kListCap + 16) * 2;   // 112 KB: two CTAs per SM

struct BaseParams {
    uint32_t state[N];                   // the generator's current block x[0..623]
    int32_t pos;                         // next output = x[pos], 1 <= pos <= 624
    uint32_t* y;                         // [kYPad]: y[k] = x[k + pos - 1]
};

struct Region {
    int64_t t0, t1;                      // word range of the stream (t = 0 is the generator's next output)
    uint64_t thresh;                     // kind 0: ceil(p 2^53); kind 1: ceil(float(p) 2^24)
    uint8_t* out;
    int32_t kind;
    int32_t pad;
};

struct MaskParams {
    Region r[kMaxRegions];
    int32_t n_regions;
    int64_t total;                       // words in all regions
    const uint32_t* windows;             // [W][624]
};

__device__ __forceinline__ uint32_t twist(uint32_t u, uint32_t v) {
    return (((u & 0x80000000u) | (v & 0x7fffffffu)) >> 1) ^ ((v & 1u) ? 0x9908b0dfu : 0u);
}
__device__ __forceinline__ uint32_t temper(uint32_t y) {
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

// out[0..623] = the 624 words after in[0..623] (only the top bit of in[0] is used).  Three dependent steps of <= 227 words.
// Ends with a barrier; the caller must have made `in` visible.
__device__ __forceinline__ void next_window(const uint32_t* in, uint32_t* out, int tid) {
    if (tid < N - M) out[tid] = in[tid + M] ^ twist(in[tid], in[tid + 1]);
    __syncthreads();
    if (tid < N - M) {
        const int j = tid + (N - M);
        out[j] = out[tid] ^ twist(in[j], in[j + 1]);
    }
    __syncthreads();
    if (tid < N - 1 - 2 * (N - M)) {   // 454 .. 622
        const int j = tid + 2 * (N - M);
        out[j] = out[j - (N - M)] ^ twist(in[j], in[j + 1]);
    } else if (tid == kT - 1) {
        out[N - 1] = out[M - 1] ^ twist(in[N - 1], out[0]);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kT) mt_base_kernel(const __grid_constant__ BaseParams p) {
    __shared__ uint32_t A[N], B[N];
    const int tid = threadIdx.x;
    for (int i = tid; i < N; i += kT) A[i] = p.state[i];
    __syncthreads();
    uint32_t* cur = A;
    uint32_t* nxt = B;
    const int shift = p.pos - 1;                      // y[k] = x[k + shift]
    for (int x0 = 0; x0 < shift + kYPad; x0 += N) {   // block x[x0 .. x0+623] is in `cur`
        for (int i = tid; i < N; i += kT) {
            const int k = x0 + i - shift;
            if (k >= 0 && k < kYPad) p.y[k] = cur[i];
        }
        next_window(cur, nxt, tid);
        uint32_t* t = cur; cur = nxt; nxt = t;
    }
}

// windows[w] = window (y[wJ] .. y[wJ + 623]) = XOR_{k : bit k of poly[w-1]} (y[k] .. y[k+623]); windows[0] = y[0..623].
__global__ void __launch_bounds__(kT) mt_jump_kernel(const uint32_t* __restrict__ y, const uint32_t* __restrict__ table,
                                                     uint32_t* __restrict__ windows) {
    extern __shared__ __align__(16) uint32_t sm[];
    __shared__ int count_s;
    uint32_t* ys = sm;                 // [kYPad]
    uint32_t* poly = sm + kYPad + 3;   // [624] (+3: kYPad is odd-sized, keep the lists 16-byte aligned), then offs, list
    const int tid = threadIdx.x;
    const int w = blockIdx.x;
    if (w == 0) {
        for (int i = tid; i < N; i += kT) windows[i] = y[i];
        return;
    }
    for (int i = tid; i < kYPad; i += kT) ys[i] = y[i];
    for (int i = tid; i < N; i += kT) poly[i] = table[(size_t)(w - 1) * N + i];
    __syncthreads();
    // Compact the polynomial into the list of its set coefficients (~10^4 of 19937): the XOR loop below then has a known
    // trip count and can keep sixteen independent shared-memory loads in flight per thread (a loop that peels one bit at a
    // time off the mask is bound by the latency of that dependent chain: measured 4x slower).
    uint16_t* list = reinterpret_cast<uint16_t*>(poly + 2 * N);   // [kListCap + 16]
    uint32_t* offs = poly + N;                                    // [624] exclusive prefix of the popcounts
    if (tid < 32) {
        uint32_t run = 0;
        for (int base = 0; base < N; base += 32) {
            const int i = base + tid;
            const uint32_t c = i < N ? (uint32_t)__popc(poly[i]) : 0u;
            uint32_t inc = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, inc, d);
                if (tid >= d) inc += v;
            }
            if (i < N) offs[i] = run + inc - c;
            run += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (tid == 0) count_s = (int)run;
    }
    __syncthreads();
    if (count_s > kListCap) __trap();
    for (int i = tid; i < N; i += kT) {
        uint32_t gw = poly[i];
        uint32_t o = offs[i];
        while (gw) {
            list[o++] = (uint16_t)(i * 32 + __ffs(gw) - 1);
            gw &= gw - 1;
        }
    }
    const int count = count_s;
    __syncthreads();
    uint32_t a0 = 0, a1 = 0;
    const uint32_t* y0 = ys + tid;
    const uint32_t* y1 = ys + tid + kT;    // j = tid + 320 (< 624 for tid < 304; reads stay inside the padding otherwise)
    const int full = count & ~7;
    for (int i = 0; i < full; i += 8) {
        const uint4 q = *reinterpret_cast<const uint4*>(list + i);   // eight 16-bit positions (broadcast)
        const uint32_t k0 = q.x & 0xffffu, k1 = q.x >> 16, k2 = q.y & 0xffffu, k3 = q.y >> 16;
        const uint32_t k4 = q.z & 0xffffu, k5 = q.z >> 16, k6 = q.w & 0xffffu, k7 = q.w >> 16;
        const uint32_t u0 = y0[k0], u1 = y0[k1], u2 = y0[k2], u3 = y0[k3], u4 = y0[k4], u5 = y0[k5], u6 = y0[k6], u7 = y0[k7];
        const uint32_t v0 = y1[k0], v1 = y1[k1], v2 = y1[k2], v3 = y1[k3], v4 = y1[k4], v5 = y1[k5], v6 = y1[k6], v7 = y1[k7];
        a0 ^= ((u0 ^ u1) ^ (u2 ^ u3)) ^ ((u4 ^ u5) ^ (u6 ^ u7));
        a1 ^= ((v0 ^ v1) ^ (v2 ^ v3)) ^ ((v4 ^ v5) ^ (v6 ^ v7));
    }
    for (int i = full; i < count; ++i) {
        a0 ^= y0[list[i]];
        a1 ^= y1[list[i]];
    }
    uint32_t* out = windows + (size_t)w * N;
    out[tid] = a0;
    if (tid + kT < N) out[tid + kT] = a1;
}

__global__ void __launch_bounds__(kT) mt_mask_kernel(const __grid_constant__ MaskParams p) {
    __shared__ uint32_t A[N], B[N];
    const int tid = threadIdx.x;
    const int w = blockIdx.x;
    const uint32_t* win = p.windows + (size_t)w * N;
    for (int i = tid; i < N; i += kT) A[i] = win[i];
    __syncthreads();
    uint32_t* cur = A;
    uint32_t* nxt = B;
    const int64_t t_begin = (int64_t)w * J;
    const int64_t t_end = t_begin + J < p.total ? t_begin + J : p.total;
    // frame f: stream words t_begin + 624 f + i, i = 0..623 = cur[1..623], nxt[0]
    for (int64_t tf = t_begin; tf < t_end; tf += N) {
        next_window(cur, nxt, tid);
        const int64_t t = tf + 2 * tid;
        if (tid < N / 2 && t < t_end) {
            const uint32_t w0 = temper(cur[2 * tid + 1]);
            const uint32_t w1 = temper(2 * tid + 2 < N ? cur[2 * tid + 2] : nxt[0]);
            int ri = 0;
#pragma unroll
            for (int i = 0; i < kMaxRegions - 1; ++i) ri += (i < p.n_regions - 1 && t >= p.r[i].t1) ? 1 : 0;
            const Region& r = p.r[ri];
            if (r.kind == 0) {
                const uint64_t v = ((uint64_t)(w0 & 0x1fffffu) << 32) | (uint64_t)w1;
                r.out[(t - r.t0) >> 1] = v < r.thresh ? 1 : 0;
            } else if (r.kind == 1) {
                const uint32_t th = (uint32_t)r.thresh;
                const uchar2 k = make_uchar2((w0 & 0xffffffu) < th ? 1 : 0, (w1 & 0xffffffu) < th ? 1 : 0);
                uint8_t* o = r.out + (t - r.t0);
                if (t + 1 < r.t1) *reinterpret_cast<uchar2*>(o) = k;   // (t - t0 is even and `out` 2-byte aligned)
                else o[0] = k.x;
            }
        }
        __syncthreads();   // every read of `cur` is done before it becomes the next frame's output buffer
        uint32_t* tmp = cur; cur = nxt; nxt = tmp;
    }
}

// DropBlock._compute_block_mask (resnet_language.py:327-352): keep[pl, y, x] = 0 iff a seed sits at (y - i, x - j) for some
// 0 <= i, j < bs.  One thread per output element; kept elements are counted exactly (integer atomics).
__global__ void dropblock_kernel(const uint8_t* __restrict__ seeds, int64_t planes, int hs, int ws, int bs,
                                 uint8_t* __restrict__ keep, unsigned long long* __restrict__ kept) {
    const int ho = hs + bs - 1, wo = ws + bs - 1;
    const int64_t total = planes * ho * wo;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int k = 0;
    if (i < total) {
        const int x = (int)(i % wo);
        const int64_t r = i / wo;
        const int y = (int)(r % ho);
        const int64_t pl = r / ho;
        const uint8_t* sp = seeds + pl * hs * ws;
        k = 1;
        for (int dy = 0; dy < bs; ++dy) {
            const int sy = y - dy;
            if (sy < 0 || sy >= hs) continue;
            for (int dx = 0; dx < bs; ++dx) {
                const int sx = x - dx;
                if (sx < 0 || sx >= ws) continue;
                if (sp[sy * ws + sx]) k = 0;
            }
        }
        keep[i] = (uint8_t)k;
    }
    // one atomic per CTA (a warp-level count per atomic put 1.3e5 atomics on one address: 75 us for a 4 M-element mask)
    __shared__ unsigned s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const unsigned m = __ballot_sync(0xffffffffu, k != 0);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&s_cnt, (unsigned)__popc(m));
    __syncthreads();
    if (threadIdx.x == 0 && s_cnt) atomicAdd(kept, (unsigned long long)s_cnt);
}

// countM / count_ones of resnet_language.py:321-323: python ints turned into fp32 0-d tensors, fp32 division.
__global__ void dropblock_scale_kernel(const unsigned long long* kept, int64_t numel, float* scale) {
    scale[0] = __fdiv_rn((float)(double)numel, (float)(double)(*kept));
}

struct Blob {  // at::CPUGeneratorImplStateLegacy (host_rng.cpp)
    uint64_t seed;
    int32_t left;
    int32_t seeded;
    uint64_t next;
    uint64_t state[N];
};

uint64_t threshold(double p, int bits) {  // ceil(p * 2^bits), clamped to [0, 2^bits]
    if (!(p > 0.0)) return 0;
    if (p >= 1.0) return 1ull << bits;
    return (uint64_t)__builtin_ceil(__builtin_ldexp(p, bits));
}

int64_t region_words(const sr_mask_region& r) { return r.kind == 0 ? 2 * r.n : r.n; }
}  // namespace

extern "C" int64_t sr_device_bernoulli_workspace_bytes(int64_t total_words) {
    if (total_words < 0) return -1;
    const int64_t W = (total_words + J - 1) / J;
    return srb::align_up((int64_t)kYPad * 4, 256) + srb::align_up(std::max<int64_t>(W, 1) * N * 4, 256);
}

extern "C" int32_t sr_device_bernoulli(void* state_blob, int64_t blob_bytes, const sr_mask_region* regions, int32_t n_regions,
                                       const uint32_t* table_dev, const uint32_t* table_host, int32_t n_polys,
                                       void* workspace, int64_t workspace_bytes, void* stream) {
    using namespace srb;
    if (!state_blob || blob_bytes < (int64_t)sizeof(Blob) || !regions || n_regions < 1 || n_regions > kMaxRegions)
        return fail(SR_E_ARG, "sr_device_bernoulli: bad arguments (1..%d regions)", kMaxRegions);
    Blob* b = static_cast<Blob*>(state_blob);
    if (!b->seeded || b->left < 1 || b->left > N || b->next > (uint64_t)N)
        return fail(SR_E_ARG, "sr_device_bernoulli: unexpected generator state layout");
    MaskParams mp;
    memset(&mp, 0, sizeof(mp));
    int64_t t = 0;
    for (int i = 0; i < n_regions; ++i) {
        const sr_mask_region& r = regions[i];
        if (r.kind < 0 || r.kind > 2 || r.n < 0 || (r.kind != 2 && r.n > 0 && !r.out))
            return fail(SR_E_ARG, "sr_device_bernoulli: region %d is malformed", i);
        const int64_t words = region_words(r);
        if ((words & 1) && i + 1 < n_regions)
            return fail(SR_E_ARG, "sr_device_bernoulli: region %d has an odd word count (only the last one may)", i);
        if (r.kind == 1 && (reinterpret_cast<uintptr_t>(r.out) & 1))
            return fail(SR_E_ARG, "sr_device_bernoulli: region %d output must be 2-byte aligned", i);
        mp.r[i].t0 = t;
        mp.r[i].t1 = t + words;
        mp.r[i].kind = r.kind;
        mp.r[i].thresh = r.kind == 0 ? threshold(r.p, 53) : threshold((double)(float)r.p, 24);
        mp.r[i].out = r.out;
        t += words;
    }
    mp.n_regions = n_regions;
    mp.total = t;
    if (t == 0) return SR_OK;
    const int64_t W = (t + J - 1) / J;
    if (W - 1 > n_polys || (W > 1 && (!table_dev || !table_host)))
        return fail(SR_E_ARG, "sr_device_bernoulli: %lld words need %lld jump polynomials, the table has %d", (long long)t,
                    (long long)(W - 1), n_polys);
    const int64_t need = sr_device_bernoulli_workspace_bytes(t);
    if (!workspace || workspace_bytes < need)
        return fail(SR_E_SMALLWS, "sr_device_bernoulli: workspace %lld < %lld", (long long)workspace_bytes, (long long)need);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uint32_t* y = static_cast<uint32_t*>(workspace);
    uint32_t* windows = reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(workspace) + align_up((int64_t)kYPad * 4, 256));
    mp.windows = windows;

    BaseParams bp;
    for (int i = 0; i < N; ++i) bp.state[i] = (uint32_t)b->state[i];
    bp.pos = (b->left - 1) == 0 ? N : (int32_t)b->next;
    if (bp.pos < 1) return fail(SR_E_ARG, "sr_device_bernoulli: unexpected generator position");
    bp.y = y;
    static PerDeviceOnce once;
    once_per_device(once, [] {
        cudaFuncSetAttribute(mt_jump_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kJumpSmem);
    });
    mt_base_kernel<<<1, kT, 0, st>>>(bp);
    mt_jump_kernel<<<(unsigned)W, kT, kJumpSmem, st>>>(y, table_dev, windows);
    mt_mask_kernel<<<(unsigned)W, kT, 0, st>>>(mp);
    SR_CUDA_OK(cudaGetLastError());
    // the host generator moves past the words the device draws
    const int32_t rc = sr_host_mt_advance(state_blob, blob_bytes, t, table_host, n_polys);
    if (rc != SR_OK) return fail(rc, "sr_device_bernoulli: sr_host_mt_advance failed");
    return SR_OK;
}

extern "C" int32_t sr_dropblock_keep(const uint8_t* seeds, int64_t planes, int32_t hs, int32_t ws, int32_t bs, uint8_t* keep,
                                     float* scale_out, void* stream) {
    using namespace srb;
    if (!seeds || !keep || !scale_out || planes < 0 || hs < 1 || ws < 1 || bs < 1)
        return fail(SR_E_ARG, "sr_dropblock_keep: bad arguments");
    if (reinterpret_cast<uintptr_t>(scale_out) & 15)
        return fail(SR_E_ARG, "sr_dropblock_keep: scale_out must be 16-byte aligned (4 floats)");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned long long* kept = reinterpret_cast<unsigned long long*>(scale_out + 2);
    SR_CUDA_OK(cudaMemsetAsync(scale_out, 0, 16, st));
    const int64_t total = planes * (hs + bs - 1) * (ws + bs - 1);
    if (total > 0) dropblock_kernel<<<(unsigned)((total + 1023) / 1024), 1024, 0, st>>>(seeds, planes, hs, ws, bs, keep, kept);
    dropblock_scale_kernel<<<1, 1, 0, st>>>(kept, total, scale_out);
    SR_CUDA_OK(cudaGetLastError());
    return SR_OK;
}
