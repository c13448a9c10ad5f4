// Data-layout and train-mode BatchNorm kernels around the tcgen05 convolution (HBM-bound elementwise work:
// coalesced 128-bit accesses, grids sized from the element count).
#include <cuda_bf16.h>
#include <algorithm>
#include "common.h"

namespace {

using namespace srb;

__device__ __forceinline__ float lrelu(float x, float slope) { return x > 0.f ? x : x * slope; }

// Error-compensated operands (see sr_conv_panel): hi = rn(x), lo = rn(x - hi).
__device__ __forceinline__ void store8_split(const float* v, __nv_bfloat16* hi, __nv_bfloat16* lo) {
    __align__(16) __nv_bfloat16 h[8];
    __align__(16) __nv_bfloat16 l[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        h[j] = __float2bfloat16_rn(v[j]);
        l[j] = __float2bfloat16_rn(v[j] - __bfloat162float(h[j]));
    }
    *reinterpret_cast<uint4*>(hi) = *reinterpret_cast<const uint4*>(h);
    if (lo) *reinterpret_cast<uint4*>(lo) = *reinterpret_cast<const uint4*>(l);
}

// ---- NCHW fp32 -> NHWC bf16, channels zero-padded to cpad ----------------------------------------------------
__global__ void pack_input_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                  __nv_bfloat16* __restrict__ y_lo, int64_t npix_total, int C, int HW, int cpad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // pixel index over (n, h, w)
    if (i >= npix_total) return;
    const int64_t n = i / HW;
    const int64_t hw = i - n * HW;
    const float* src = x + n * C * HW + hw;
    __nv_bfloat16* dst = y + i * cpad;
    for (int c0 = 0; c0 < cpad; c0 += 8) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = c0 + j;
            v[j] = c < C ? src[(int64_t)c * HW] : 0.f;
        }
        store8_split(v, dst + c0, y_lo ? y_lo + i * cpad + c0 : nullptr);
    }
}

// ---- uint8 HWC image store -> ToTensor (x / 255) -> Normalize ((x - mean) / std) -> NHWC bf16, channels zero-padded ----
// Same fp32 operations, in the same order, as torchvision's ToTensor + Normalize followed by pack_input_kernel.
struct NormParams {
    float mean[4], stdev[4];
};
__global__ void pack_input_u8_kernel(const uint8_t* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                     __nv_bfloat16* __restrict__ y_lo, int64_t npix_total, int C, int H, int W, int cpad,
                                     NormParams np, const int32_t* __restrict__ crop_ij, const uint8_t* __restrict__ flip,
                                     int pad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // output pixel index over (n, h, w)
    if (i >= npix_total) return;
    // RandomCrop(size, padding=pad) + RandomHorizontalFlip of the reference's support transform, applied while reading:
    // output (h, w) of image n comes from input (h + i0 - pad, w' + j0 - pad), w' = W-1-w when flipped; outside the image
    // the zero padding (a 0 byte BEFORE ToTensor / Normalize, like PIL's constant fill).
    const int64_t hw = (int64_t)H * W;
    const int64_t n = i / hw;
    const int r = (int)(i - n * hw);
    int h = r / W, w = r - h * W;
    bool inside = true;
    if (crop_ij != nullptr || flip != nullptr) {
        if (flip != nullptr && flip[n]) w = W - 1 - w;
        if (crop_ij != nullptr) {
            h += crop_ij[2 * n] - pad;
            w += crop_ij[2 * n + 1] - pad;
        }
        inside = h >= 0 && h < H && w >= 0 && w < W;
    }
    const uint8_t* src = x + (n * hw + (int64_t)h * W + w) * C;
    __nv_bfloat16* dst = y + i * cpad;
    for (int c0 = 0; c0 < cpad; c0 += 8) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = c0 + j;
            float f = 0.f;
            if (c < C) {
                f = __fdiv_rn(inside ? (float)src[c] : 0.f, 255.0f);
                f = __fdiv_rn(__fsub_rn(f, np.mean[c]), np.stdev[c]);
            }
            v[j] = f;
        }
        store8_split(v, dst + c0, y_lo ? y_lo + i * cpad + c0 : nullptr);
    }
}

__global__ void bn_fold_kernel(const float* gamma, const float* beta, const float* rm, const float* rv, float eps,
                               float* scale, float* shift, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float s = gamma[c] / sqrtf(rv[c] + eps);
    scale[c] = s;
    shift[c] = beta[c] - rm[c] * s;
}

// ---- OIHW fp32 -> [cout][taps][cin_pad] bf16 (optionally scaled per output channel) ---------------------------
__global__ void pack_weight_kernel(const float* __restrict__ w, const float* __restrict__ scale,
                                   __nv_bfloat16* __restrict__ out, __nv_bfloat16* __restrict__ out_lo, int cout, int cin,
                                   int taps, int cin_pad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)cout * taps * cin_pad;
    if (i >= total) return;
    const int ci = (int)(i % cin_pad);
    const int tap = (int)((i / cin_pad) % taps);
    const int co = (int)(i / ((int64_t)cin_pad * taps));
    float v = 0.f;
    if (ci < cin) {
        v = w[((int64_t)co * cin + ci) * taps + tap];
        if (scale) v *= scale[co];
    }
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    out[i] = h;
    if (out_lo) out_lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
}

// ---- train-mode BN statistics -> mean / invstd, running-stat EMA ---------------------------------------------
__global__ void bn_finalize_kernel(const double* stats, double count, float eps, float momentum, float* rm, float* rv,
                                   float* mean, float* invstd, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double mu = stats[c] / count;
    double var = stats[C + c] / count - mu * mu;
    if (var < 0.0) var = 0.0;
    mean[c] = (float)mu;
    invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    rm[c] = momentum * (float)mu + (1.f - momentum) * rm[c];
    rv[c] = momentum * (float)unbiased + (1.f - momentum) * rv[c];
}

struct BnApplyParams {
    sr_bn_apply_args a;
    int Ho, Wo;
};

__device__ __forceinline__ void load8(const float* p, float* v) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// y = lrelu(bn(raw) [+ bn(res_raw) | + res_act]) at one pixel, 8 consecutive channels starting at c0.
__device__ __forceinline__ void bn_pixel(const sr_bn_apply_args& a, int64_t pix, int c0, const float* al, const float* be,
                                         const float* ral, const float* rbe, float* y) {
    float x[8];
    load8(a.raw + pix * a.channels + c0, x);
#pragma unroll
    for (int j = 0; j < 8; ++j) y[j] = x[j] * al[j] + be[j];
    if (a.res_raw) {
        load8(a.res_raw + pix * a.channels + c0, x);
#pragma unroll
        for (int j = 0; j < 8; ++j) y[j] += x[j] * ral[j] + rbe[j];
    } else if (a.res_act) {
        const uint4 r = __ldg(reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(a.res_act) +
                                                              pix * a.channels + c0));
        const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162*>(&rr[k]);
            y[2 * k] += __bfloat162float(b2.x);
            y[2 * k + 1] += __bfloat162float(b2.y);
        }
        if (a.res_act_lo) {
            const uint4 q = __ldg(reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(a.res_act_lo) +
                                                                  pix * a.channels + c0));
            const uint32_t qq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162*>(&qq[k]);
                y[2 * k] += __bfloat162float(b2.x);
                y[2 * k + 1] += __bfloat162float(b2.y);
            }
        }
    }
    if (a.lrelu) {
#pragma unroll
        for (int j = 0; j < 8; ++j) y[j] = lrelu(y[j], a.slope);
    }
}

__device__ __forceinline__ void bn_coeffs(const float* mean, const float* invstd, const float* gamma, const float* beta,
                                          int c0, float* al, float* be) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float s = invstd[c0 + j] * gamma[c0 + j];
        al[j] = s;
        be[j] = beta[c0 + j] - mean[c0 + j] * s;
    }
}

// pool 0 / 2: one thread per (n, ho, wo, 8-channel group) -> bf16 NHWC
__global__ void bn_apply_kernel(const BnApplyParams p) {
    const sr_bn_apply_args& a = p.a;
    const int cg = a.channels >> 3;
    const int64_t total = (int64_t)a.batch * p.Ho * p.Wo * cg;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int g = (int)(i % cg);
    int64_t r = i / cg;
    const int wo = (int)(r % p.Wo);
    r /= p.Wo;
    const int ho = (int)(r % p.Ho);
    const int n = (int)(r / p.Ho);
    const int c0 = g * 8;
    float al[8], be[8], ral[8], rbe[8];
    bn_coeffs(a.mean, a.invstd, a.gamma, a.beta, c0, al, be);
    if (a.res_raw) bn_coeffs(a.res_mean, a.res_invstd, a.res_gamma, a.res_beta, c0, ral, rbe);
    float y[8];
    if (a.pool == 2) {
        float t[8];
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                const int64_t pix = ((int64_t)n * a.height + (2 * ho + dy)) * a.width + (2 * wo + dx);
                bn_pixel(a, pix, c0, al, be, ral, rbe, t);
#pragma unroll
                for (int j = 0; j < 8; ++j) y[j] = (dy == 0 && dx == 0) ? t[j] : fmaxf(y[j], t[j]);
            }
    } else {
        const int64_t pix = ((int64_t)n * a.height + ho) * a.width + wo;
        bn_pixel(a, pix, c0, al, be, ral, rbe, y);
    }
    if (a.keep) {
        const float ks = a.keep_scale_dev ? __ldg(a.keep_scale_dev) : a.keep_scale;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint8_t k = a.keep[(((int64_t)n * a.channels + c0 + j) * p.Ho + ho) * p.Wo + wo];
            y[j] = k ? y[j] * ks : 0.f;
        }
    }
    const int64_t o_off = (((int64_t)n * p.Ho + ho) * p.Wo + wo) * a.channels + c0;
    store8_split(y, static_cast<__nv_bfloat16*>(a.out) + o_off,
                 a.out_lo ? static_cast<__nv_bfloat16*>(a.out_lo) + o_off : nullptr);
}

// pool -1: one thread per (n, 8-channel group): mask, then mean over H x W -> fp32 [B, C]
__global__ void bn_apply_avg_kernel(const BnApplyParams p) {
    const sr_bn_apply_args& a = p.a;
    const int cg = a.channels >> 3;
    const int64_t total = (int64_t)a.batch * cg;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int g = (int)(i % cg);
    const int n = (int)(i / cg);
    const int c0 = g * 8;
    float al[8], be[8], ral[8], rbe[8];
    bn_coeffs(a.mean, a.invstd, a.gamma, a.beta, c0, al, be);
    if (a.res_raw) bn_coeffs(a.res_mean, a.res_invstd, a.res_gamma, a.res_beta, c0, ral, rbe);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const int HW = a.height * a.width;
    const float ks = (a.keep && a.keep_scale_dev) ? __ldg(a.keep_scale_dev) : a.keep_scale;
    for (int q = 0; q < HW; ++q) {
        float y[8];
        bn_pixel(a, (int64_t)n * HW + q, c0, al, be, ral, rbe, y);
        if (a.keep) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint8_t k = a.keep[((int64_t)n * a.channels + c0 + j) * HW + q];
                y[j] = k ? y[j] * ks : 0.f;
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += y[j];
    }
    float* o = static_cast<float*>(a.out) + (int64_t)n * a.channels + c0;
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = acc[j] / (float)HW;
}

// bf16 NHWC [B,H,W,C] -> fp32 [B,C] mean over H x W (AdaptiveAvgPool2d(1) after a pooled last block: resnet12)
__global__ void global_avg_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ x_lo,
                                  float* __restrict__ y, int B, int HW, int C) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * C) return;
    const int c = (int)(i % C);
    const int64_t n = i / C;
    float acc = 0.f;
    for (int q = 0; q < HW; ++q) {
        float v = __bfloat162float(x[(n * HW + q) * C + c]);
        if (x_lo) v += __bfloat162float(x_lo[(n * HW + q) * C + c]);
        acc += v;
    }
    y[i] = acc / (float)HW;
}

}  // namespace

extern "C" int32_t sr_global_avg(const void* x_nhwc_bf16, const void* x_lo, float* y, int32_t batch, int32_t height,
                                 int32_t width, int32_t channels, void* stream_v) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    if (!x_nhwc_bf16 || !y || batch < 1 || height < 1 || width < 1 || channels < 1)
        return fail(SR_E_ARG, "sr_global_avg: bad arguments");
    const int64_t total = (int64_t)batch * channels;
    global_avg_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(x_nhwc_bf16), static_cast<const __nv_bfloat16*>(x_lo), y, batch, height * width,
        channels);
    SR_CUDA_OK(cudaGetLastError());
    return SR_OK;
}

extern "C" int32_t sr_pack_input(const float* x, void* y, void* y_lo, int32_t batch, int32_t channels, int32_t height,
                                 int32_t width, int32_t cpad, void* stream_v) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    if (!x || !y || batch < 1 || channels < 1 || cpad < channels || cpad % 8)
        return fail(SR_E_ARG, "sr_pack_input: bad arguments");
    const int64_t npix = (int64_t)batch * height * width;
    const int threads = 256;
    pack_input_kernel<<<(unsigned)((npix + threads - 1) / threads), threads, 0, stream>>>(
        x, static_cast<__nv_bfloat16*>(y), static_cast<__nv_bfloat16*>(y_lo), npix, channels, height * width, cpad);
    SR_CUDA_OK(cudaGetLastError());
    return SR_OK;
}

extern "C" int32_t sr_pack_input_u8(const uint8_t* x, void* y, void* y_lo, int32_t batch, int32_t channels, int32_t height,
                                    int32_t width, const float* mean_host, const float* std_host, int32_t cpad,
                                    const int32_t* crop_ij, const uint8_t* flip, int32_t pad, void* stream_v) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    if (!x || !y || !mean_host || !std_host || batch < 1 || channels < 1 || channels > 4 || cpad < channels || cpad % 8)
        return fail(SR_E_ARG, "sr_pack_input_u8: bad arguments");
    NormParams np;
    for (int c = 0; c < 4; ++c) {
        np.mean[c] = c < channels ? mean_host[c] : 0.f;
        np.stdev[c] = c < channels ? std_host[c] : 1.f;
        if (c < channels && !(np.stdev[c] > 0.f)) return fail(SR_E_ARG, "sr_pack_input_u8: std must be positive");
    }
    const int64_t npix = (int64_t)batch * height * width;
    const int threads = 256;
    if (pad < 0 || (crop_ij && pad > 64)) return fail(SR_E_ARG, "sr_pack_input_u8: bad padding");
    pack_input_u8_kernel<<<(unsigned)((npix + threads - 1) / threads), threads, 0, stream>>>(
        x, static_cast<__nv_bfloat16*>(y), static_cast<__nv_bfloat16*>(y_lo), npix, channels, height, width, cpad, np, crop_ij,
        flip, pad);
    SR_CUDA_OK(cudaGetLastError());
    return SR_OK;
}

extern "C" int32_t sr_bn_fold(const float* gamma, const float* beta, const float* rm, const float* rv, float eps,
                              float* scale, float* shift, int32_t channels, void* stream_v) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    if (!gamma || !beta || !rm || !rv || !scale || !shift || channels < 1) return fail(SR_E_ARG, "sr_bn_fold: bad arguments");
    bn_fold_kernel<<<(channels + 127) / 128, 128, 0, stream>>>(gamma, beta, rm, rv, eps, scale, shift, channels);
    SR_CUDA_OK(cudaGetLastError());
    return SR_OK;
}

extern "C" int32_t sr_pack_weight(const float* w, const float* scale, void* out, void* out_lo, int32_t cout, int32_t cin,
                                  int32_t kh, int32_t kw, int32_t cin_pad, void* stream_v) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    if (!w || !out || cout < 1 || cin < 1 || cin_pad < cin || kh != kw || (kh != 1 && kh != 3))
        return fail(SR_E_ARG, "sr_pack_weight: bad arguments");
    const int64_t total = (int64_t)cout * kh * kw * cin_pad;
    const int threads = 256;
    pack_weight_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, stream>>>(
        w, scale, static_cast<__nv_bfloat16*>(out), static_cast<__nv_bfloat16*>(out_lo), cout, cin, kh * kw, cin_pad);
    SR_CUDA_OK(cudaGetLastError());
    return SR_OK;
}

extern "C" int32_t sr_bn_finalize(const double* stats, int64_t count, float eps, float momentum, float* rm, float* rv,
                                  float* mean, float* invstd, int32_t channels, void* stream_v) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    if (!stats || !rm || !rv || !mean || !invstd || channels < 1 || count < 1)
        return fail(SR_E_ARG, "sr_bn_finalize: bad arguments");
    bn_finalize_kernel<<<(channels + 127) / 128, 128, 0, stream>>>(stats, (double)count, eps, momentum, rm, rv, mean,
                                                                    invstd, channels);
    SR_CUDA_OK(cudaGetLastError());
    return SR_OK;
}

extern "C" int32_t sr_bn_apply(const sr_bn_apply_args* a, void* stream_v);

// ---- whole eval-mode backbone pass -------------------------------------------------------------------------------
namespace {
// Four activation buffers (conv1 out, conv2 out, and two alternating block outputs: a block's input stays live until its
// conv3 has read it as residual / downsample panel), each sized for the largest layer, x2 in the error-compensated tier.
int64_t eval_buffer_bytes(const sr_backbone_eval_args* a) {
    int64_t mx = 0;
    int h = a->height, w = a->width;
    for (int i = 0; i < a->n_blocks; ++i) {
        mx = std::max<int64_t>(mx, (int64_t)a->batch * h * w * a->blocks[i].cout * 2);
        if (a->blocks[i].pool == 2) { h /= 2; w /= 2; }
    }
    return align_up(mx * (a->x_lo ? 2 : 1), 256);
}
}  // namespace

extern "C" int64_t sr_backbone_eval_workspace_bytes(const sr_backbone_eval_args* a) {
    if (!a || !a->blocks || a->n_blocks < 1 || a->batch < 1) return 0;
    return 4 * eval_buffer_bytes(a);
}

extern "C" int32_t sr_backbone_eval(const sr_backbone_eval_args* a, void* stream_v) {
    if (!a || !a->blocks || a->n_blocks < 1 || !a->x || !a->features || !a->workspace)
        return fail(SR_E_ARG, "sr_backbone_eval: null pointer");
    if (reinterpret_cast<uintptr_t>(a->workspace) & 255) return fail(SR_E_ARG, "sr_backbone_eval: workspace must be 256-byte aligned");
    const int64_t buf = eval_buffer_bytes(a);
    if (a->workspace_bytes < 4 * buf) return fail(SR_E_SMALLWS, "sr_backbone_eval: workspace %lld < %lld", (long long)a->workspace_bytes, (long long)(4 * buf));
    const bool precise = a->x_lo != nullptr;
    uint8_t* ws = static_cast<uint8_t*>(a->workspace);
    // plane pointers of buffer k: hi at the start, lo in the second half
    auto hi = [&](int k) { return static_cast<void*>(ws + k * buf); };
    auto lo = [&](int k) { return precise ? static_cast<void*>(ws + k * buf + buf / 2) : nullptr; };
    const void* in = a->x;
    const void* in_lo = a->x_lo;
    int cin_pad = a->cin_pad, h = a->height, w = a->width;
    for (int i = 0; i < a->n_blocks; ++i) {
        const sr_eval_block& b = a->blocks[i];
        const bool last = i == a->n_blocks - 1;
        if (!b.w1 || !b.w2 || !b.w3 || (b.downsample && !b.wd)) return fail(SR_E_ARG, "sr_backbone_eval: block %d lacks weights", i);
        sr_conv_args c;
        auto conv = [&](const void* act, const void* act_lo, int cin, const void* wgt, const void* wgt_lo, const float* shift,
                        int out_k, int epi, const void* res, const void* res_lo, const void* ds_act, const void* ds_act_lo,
                        int ds_cin) -> int32_t {
            memset(&c, 0, sizeof(c));
            c.batch = a->batch; c.height = h; c.width = w; c.cout = b.cout;
            c.n_panels = ds_act ? 2 : 1;
            c.panel[0].act = act; c.panel[0].act_lo = precise ? act_lo : nullptr;
            c.panel[0].wgt = wgt; c.panel[0].wgt_lo = precise ? wgt_lo : nullptr;
            c.panel[0].cin_pad = cin; c.panel[0].taps = 9;
            if (ds_act) {
                c.panel[1].act = ds_act; c.panel[1].act_lo = precise ? ds_act_lo : nullptr;
                c.panel[1].wgt = b.wd; c.panel[1].wgt_lo = precise ? b.wd_lo : nullptr;
                c.panel[1].cin_pad = ds_cin; c.panel[1].taps = 1;
            }
            c.shift = shift; c.residual = res; c.residual_lo = precise ? res_lo : nullptr;
            c.slope = a->slope; c.epilogue = epi;
            if (epi == SR_EPI_ACT_AVG) {
                c.out = a->features;
            } else {
                c.out = hi(out_k); c.out_lo = lo(out_k);
            }
            return sr_conv(&c, stream_v);
        };
        int32_t rc = conv(in, in_lo, cin_pad, b.w1, b.w1_lo, b.s1, 0, SR_EPI_ACT, nullptr, nullptr, nullptr, nullptr, 0);
        if (rc != SR_OK) return rc;
        rc = conv(hi(0), lo(0), b.cout, b.w2, b.w2_lo, b.s2, 1, SR_EPI_ACT, nullptr, nullptr, nullptr, nullptr, 0);
        if (rc != SR_OK) return rc;
        const int out_k = 2 + (i & 1);
        int epi = b.pool == 2 ? SR_EPI_ACT_POOL2 : SR_EPI_ACT;
        const bool fused_avg = last && b.pool != 2;      // resnet18: the last block averages inside the conv epilogue
        if (fused_avg) epi = SR_EPI_ACT_AVG;
        if (b.downsample)
            rc = conv(hi(1), lo(1), b.cout, b.w3, b.w3_lo, b.s3, out_k, epi, nullptr, nullptr, in, in_lo, cin_pad);
        else
            rc = conv(hi(1), lo(1), b.cout, b.w3, b.w3_lo, b.s3, out_k, epi, in, in_lo, nullptr, nullptr, 0);
        if (rc != SR_OK) return rc;
        if (b.pool == 2) { h /= 2; w /= 2; }
        if (last && !fused_avg) return sr_global_avg(hi(out_k), lo(out_k), a->features, a->batch, h, w, b.cout, stream_v);
        in = hi(out_k);
        in_lo = lo(out_k);
        cin_pad = b.cout;
    }
    return SR_OK;
}

extern "C" int32_t sr_train_block(const sr_train_block_args* a, void* stream_v) {
    if (!a || !a->x || !a->stats || !a->mean_invstd || !a->h1 || !a->h2) return fail(SR_E_ARG, "sr_train_block: null pointer");
    const int n_convs = a->downsample ? 4 : 3;
    const bool precise = a->x_lo != nullptr;
    for (int i = 0; i < n_convs; ++i)
        if (!a->w[i] || !a->raw[i] || !a->running_mean[i] || !a->running_var[i] || (precise && !a->w_lo[i]))
            return fail(SR_E_ARG, "sr_train_block: null pointer for conv %d", i);
    if (precise && (!a->h1_lo || !a->h2_lo)) return fail(SR_E_ARG, "sr_train_block: the x3 tier needs h1_lo / h2_lo");
    const int64_t count = (int64_t)a->batch * a->height * a->width;
    const void* in[4] = {a->x, a->h1, a->h2, a->x};
    const void* in_lo[4] = {a->x_lo, a->h1_lo, a->h2_lo, a->x_lo};
    void* act[2] = {a->h1, a->h2};
    void* act_lo[2] = {a->h1_lo, a->h2_lo};
    for (int i = 0; i < n_convs; ++i) {
        sr_conv_args c;
        memset(&c, 0, sizeof(c));
        c.batch = a->batch; c.height = a->height; c.width = a->width; c.cout = a->cout; c.n_panels = 1;
        c.panel[0].act = in[i]; c.panel[0].act_lo = precise ? in_lo[i] : nullptr;
        c.panel[0].wgt = a->w[i]; c.panel[0].wgt_lo = precise ? a->w_lo[i] : nullptr;
        c.panel[0].cin_pad = (i == 0 || i == 3) ? a->cin_pad : a->cout;
        c.panel[0].taps = i == 3 ? 1 : 9;
        c.epilogue = SR_EPI_RAW_STATS;
        c.out = a->raw[i];
        c.stats = a->stats + (int64_t)i * 2 * a->cout;
        int32_t rc = sr_conv(&c, stream_v);
        if (rc != SR_OK) return rc;
        float* mean = a->mean_invstd + (int64_t)i * 2 * a->cout;
        rc = sr_bn_finalize(c.stats, count, a->eps, a->momentum, a->running_mean[i], a->running_var[i], mean, mean + a->cout,
                            a->cout, stream_v);
        if (rc != SR_OK) return rc;
        if (i < 2) {   // bn1 / bn2 + LeakyReLU -> the next conv's input
            sr_bn_apply_args b;
            memset(&b, 0, sizeof(b));
            b.batch = a->batch; b.height = a->height; b.width = a->width; b.channels = a->cout;
            b.raw = a->raw[i]; b.mean = mean; b.invstd = mean + a->cout; b.gamma = a->gamma[i]; b.beta = a->beta[i];
            b.lrelu = 1; b.slope = a->slope; b.pool = 0; b.keep_scale = 1.f;
            b.out = act[i]; b.out_lo = precise ? act_lo[i] : nullptr;
            rc = sr_bn_apply(&b, stream_v);
            if (rc != SR_OK) return rc;
        }
    }
    return SR_OK;
}

extern "C" int32_t sr_bn_apply(const sr_bn_apply_args* a, void* stream_v) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    if (!a || !a->raw || !a->mean || !a->invstd || !a->gamma || !a->beta || !a->out)
        return fail(SR_E_ARG, "sr_bn_apply: null pointer");
    if (a->channels % 8) return fail(SR_E_ARG, "sr_bn_apply: channels must be a multiple of 8");
    if (a->pool != 0 && a->pool != 2 && a->pool != -1) return fail(SR_E_ARG, "sr_bn_apply: pool must be 0, 2 or -1");
    if (a->res_raw && (!a->res_mean || !a->res_invstd || !a->res_gamma || !a->res_beta))
        return fail(SR_E_ARG, "sr_bn_apply: res_raw needs its BN parameters");
    BnApplyParams p;
    p.a = *a;
    p.Ho = a->pool == 2 ? a->height / 2 : a->height;
    p.Wo = a->pool == 2 ? a->width / 2 : a->width;
    const int threads = 256;
    if (a->pool == -1) {
        const int64_t total = (int64_t)a->batch * (a->channels / 8);
        bn_apply_avg_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, stream>>>(p);
    } else {
        const int64_t total = (int64_t)a->batch * p.Ho * p.Wo * (a->channels / 8);
        bn_apply_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, stream>>>(p);
    }
    SR_CUDA_OK(cudaGetLastError());
    return SR_OK;
}
