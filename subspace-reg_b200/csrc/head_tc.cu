// Tensor-core classifier head for LARGE problems (BASELINE config 5: 10^4 rows x 1100 classes x 512 features).
//
// Same arithmetic as head_kernel (head.cu) - CE over support + memory rows, un-squared drift norms, subspace / fixed
// pullers, weight decay, SGD-momentum / Adam, the reference's stopping rule (language_eval.py:242-318) - but the two
// GEMMs that carry all the work run on tcgen05 through the implicit-GEMM kernel of conv.cu, as 1x1 convolutions in its
// error-compensated mode (operands split into bf16 hi / lo planes, hi*hi + hi*lo + lo*hi into one fp32 TMEM accumulator:
// ~2^-17 per product, which keeps the 1e-5 tier of the fp32 head):
//
//   logits   Z[n, c]  = sum_k X[n, k] W[c, k]          "image" = sample, 1x1 pixels, cin = d, cout = C (padded to 256)
//   dW       dW[c, k] = sum_n dZ[n, c] X[n, k]         split-K: the sample range is cut into S slices; slice s is an
//                                                      "image" of 2 x (Cp/2) pixels (= classes) with cin = slice samples
//                                                      and ITS OWN weights X^T[s] (sr_conv_args.weights_per_image), so
//                                                      S x (Cp/256) x (d/256) CTAs fill the machine; the S partial
//                                                      products are summed in a fixed order by the update kernel.
//
// One epoch = six launches enqueued back to back by the host loop below (no host synchronisation; a device flag turns the
// launches after the stopping rule has fired into no-ops):
//   conv(Z) -> row statistics (lse, CE, hits) -> dZ^T as hi / lo planes -> pullers -> conv(dW partials) -> loss + rule
//   -> update (regulariser gradients, optimiser, new W in fp32 and as hi / lo planes, drift norms of the new W).
#include <algorithm>
#include <cuda_bf16.h>
#include "common.h"
#include "head_common.cuh"
#include "ptx.cuh"

namespace {
using namespace srb;

constexpr int kT = 256;
constexpr int kUpdElems = 1024;   // classifier elements per CTA of the update kernel (4 per thread)
constexpr int kMaxEpochsPerCall = 1024;

struct TcGeom {
    int N, d, C, Cp, Cz, S, slice, n_upd;   // Cp: class padding of the dW GEMM (256), Cz: of the logits GEMM / Z pitch (128)
    int n_rowctas;
    int64_t ctrl, start, Xp, XpT, Wp, Z, dZt, part, rowlse, rowpart, pull_sq, gpull, nbp, nnp, total;
};

struct TcStart {
    int epoch0, step0, stable_count0;
    float prev_loss;
    int stopflag[2];   // [e & 1] = the stopping rule has fired by the end of epoch e (read by the update CTAs of epoch e + 1)
};

bool tc_geometry(const sr_head_args* a, TcGeom* g) {
    g->N = a->n_support + a->n_memory;
    g->d = a->dim;
    g->C = a->n_classes;
    g->Cp = (int)align_up(a->n_classes, 256);
    g->Cz = (int)align_up(a->n_classes, 128);
    // (the dW GEMM maps classes to a 2 x Cp/2 feature map whose tiles must hold one slice: Cp >= 1024)
    if (a->dim % 64 != 0 || a->dim > 2048 || g->Cp > 4096 || g->Cp < 1024) return false;
    // dW tiles are 256 classes x 128 features: half the fp32 epilogue per CTA and half as many split-K partials as 256 x 256
    if (a->dim % 128 != 0) return false;
    const int ns_d = a->dim / 128;
    const int ctas_per_slice = (g->Cp / 256) * ns_d;
    g->S = std::max(1, std::min(current_device_sms() / std::max(ctas_per_slice, 1), (g->N + 63) / 64));
    g->slice = (int)align_up((g->N + g->S - 1) / g->S, 64);
    g->S = (g->N + g->slice - 1) / g->slice;          // no empty slices
    g->n_upd = (int)(((int64_t)g->C * g->d + kUpdElems - 1) / kUpdElems);
    int64_t off = 0;
    auto take = [&](int64_t bytes) { const int64_t o = off; off += align_up(bytes, 256); return o; };
    g->ctrl = take(sizeof(HeadCtrl));
    g->start = take(sizeof(TcStart));
    g->Xp = take(2ll * g->N * g->d * 2);
    g->XpT = take(2ll * g->S * g->d * g->slice * 2);
    g->Wp = take(2ll * g->Cp * g->d * 2);
    g->Z = take((int64_t)g->N * g->Cz * 4);
    g->dZt = take(2ll * g->S * g->Cp * g->slice * 2);
    g->part = take((int64_t)g->S * g->Cp * g->d * 4);
    g->n_rowctas = (g->N + 7) / 8;
    g->rowlse = take((int64_t)g->N * 4);
    g->rowpart = take((int64_t)g->n_rowctas * 4 * 8);
    g->pull_sq = take((int64_t)std::max(a->n_new, 1) * 8);
    g->gpull = take((int64_t)std::max(a->n_new, 1) * g->d * 4);
    g->nbp = take(2ll * g->n_upd * 8);   // [2][n_upd]: parity e & 1 holds the partials of W_e
    g->nnp = take(2ll * g->n_upd * 8);
    g->total = off;
    return true;
}

__device__ __forceinline__ int64_t feat_row(const sr_head_args& a, int r) {
    return r < a.n_support ? (int64_t)a.support_row0 + r : (int64_t)a.memory_row0 + (r - a.n_support);
}

__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(v);
    lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

struct TcParams {
    sr_head_args a;
    TcGeom g;
    HeadCtrl* ctrl;
    TcStart* start;
    __nv_bfloat16 *Xp, *XpT, *Wp, *dZt;
    float *Z, *part, *rowlse, *gpull;
    double *rowpart, *pull_sq, *nbp, *nnp;
};

__global__ void tc_init_kernel(const TcParams p) {
    const HeadStart st = head_start(p.a);
    p.start->epoch0 = st.epoch0;
    p.start->step0 = st.step0;
    p.start->stable_count0 = st.stable_count0;
    p.start->prev_loss = st.prev_loss;
    p.ctrl->stop = st.already_stopped ? 1 : 0;
    p.start->stopflag[0] = p.start->stopflag[1] = st.already_stopped ? 1 : 0;
    p.ctrl->epochs_done = 0;
    p.ctrl->stable_count = st.stable_count0;
    p.ctrl->prev_loss = st.prev_loss;
}

// X rows (through the support / memory row map) -> hi / lo planes [2][N][d]
__global__ void __launch_bounds__(kT) tc_split_x_kernel(const TcParams p) {
    const int d = p.g.d;
    const int64_t i = ((int64_t)blockIdx.x * kT + threadIdx.x) * 4;
    if (i >= (int64_t)p.g.N * d) return;
    const int n = (int)(i / d), k = (int)(i % d);
    const float4 v = *reinterpret_cast<const float4*>(p.a.feat + feat_row(p.a, n) * d + k);
    const float x[4] = {v.x, v.y, v.z, v.w};
    __align__(8) __nv_bfloat16 hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split_bf16(x[j], hi[j], lo[j]);
    *reinterpret_cast<uint2*>(p.Xp + i) = *reinterpret_cast<const uint2*>(hi);
    *reinterpret_cast<uint2*>(p.Xp + (int64_t)p.g.N * d + i) = *reinterpret_cast<const uint2*>(lo);
}

// X^T per sample slice -> hi / lo planes [2][S][d][slice] (zero beyond N): 32 x 32 tiles through shared memory
__global__ void __launch_bounds__(kT) tc_transpose_x_kernel(const TcParams p) {
    __shared__ float tile[32][33];
    const int d = p.g.d, slice = p.g.slice;
    const int k0 = blockIdx.x * 32, j0 = blockIdx.y * 32, s = blockIdx.z;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const int n = s * slice + j0 + r;
        tile[r][tx] = n < p.g.N ? p.a.feat[feat_row(p.a, n) * d + k0 + tx] : 0.f;
    }
    __syncthreads();
    const int64_t plane = (int64_t)p.g.S * d * slice;
    for (int r = ty; r < 32; r += 8) {
        __nv_bfloat16 hi, lo;
        split_bf16(tile[tx][r], hi, lo);
        const int64_t o = ((int64_t)s * d + k0 + r) * slice + j0 + tx;
        p.XpT[o] = hi;
        p.XpT[plane + o] = lo;
    }
}

// W -> hi / lo planes [2][Cp][d] (rows >= C zero) + the drift-norm partials of the initial W (update-kernel partition)
__global__ void __launch_bounds__(kT) tc_prep_w_kernel(const TcParams p) {
    __shared__ double red[32];
    const sr_head_args& a = p.a;
    const int d = p.g.d;
    const int64_t i = (int64_t)blockIdx.x * kUpdElems + threadIdx.x * 4;
    const bool has_base = a.base_weight != nullptr;
    const bool has_prev = a.reserve_weight != nullptr && a.n_prev_novel > 0;
    double nb = 0.0, nn = 0.0, unused = 0.0;
    if (i < (int64_t)p.g.Cp * d) {
        const int c = (int)(i / d);
        float w[4] = {0.f, 0.f, 0.f, 0.f};
        if (c < p.g.C) {
            const float4 v = *reinterpret_cast<const float4*>(a.weight + i);
            w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
            if (has_base && c < a.n_base) {
                const float4 r = *reinterpret_cast<const float4*>(a.base_weight + i);
                const float rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) { const float dl = w[j] - rr[j]; nb += (double)dl * dl; }
            }
            if (has_prev && c >= a.n_base && c < a.n_base + a.n_prev_novel) {
                const float4 r = *reinterpret_cast<const float4*>(a.reserve_weight + i - (int64_t)a.n_base * d);
                const float rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) { const float dl = w[j] - rr[j]; nn += (double)dl * dl; }
            }
        }
        __align__(8) __nv_bfloat16 hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) split_bf16(w[j], hi[j], lo[j]);
        *reinterpret_cast<uint2*>(p.Wp + i) = *reinterpret_cast<const uint2*>(hi);
        *reinterpret_cast<uint2*>(p.Wp + (int64_t)p.g.Cp * d + i) = *reinterpret_cast<const uint2*>(lo);
    }
    block_sum3(nb, nn, unused, red);
    if (threadIdx.x == 0 && (int)blockIdx.x < p.g.n_upd) { p.nbp[blockIdx.x] = nb; p.nnp[blockIdx.x] = nn; }
}

// Row statistics of Z: one warp per sample; the row is brought into shared memory by ONE bulk asynchronous copy (few
// registers, so six CTAs per SM keep ~200 KB of loads in flight), then log-sum-exp, CE and the top-1 / top-5 rank of the
// label are computed from there; per-CTA partial sums of the losses and hits go to the loss kernel.
constexpr int kRowsPerCta = kT / 32;

__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

__global__ void __launch_bounds__(kT) tc_rowstats_kernel(const TcParams p) {
    if (p.ctrl->stop) return;
    extern __shared__ __align__(128) float s_rows[];   // [kRowsPerCta][Cz]
    __shared__ uint64_t s_bar[kRowsPerCta];
    __shared__ double s_part[kRowsPerCta][4];
    const sr_head_args& a = p.a;
    const int C = p.g.C, Cz = p.g.Cz, N = p.g.N;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = blockIdx.x * kRowsPerCta + warp;
    double ls = 0.0, lm = 0.0, h1 = 0.0, h5 = 0.0;
    if (n < N) {
        const float* z = s_rows + warp * Cz;
        const uint32_t bar = smem_u32(&s_bar[warp]);
        if (lane == 0) {
            mbar_init(&s_bar[warp], 1);
            mbar_fence_init();
            mbar_expect_tx_a(bar, (uint32_t)Cz * 4u);
            bulk_load_1d(smem_u32(z), p.Z + (int64_t)n * Cz, (uint32_t)Cz * 4u, bar);
        }
        __syncwarp();
        const bool is_sup = n < a.n_support;
        const int y = (int)(is_sup ? a.labels_support[n] : a.labels_memory[n - a.n_support]);
        mbar_wait_a(bar, 0u);
        float mx = -INFINITY;
        for (int c = lane; c < C; c += 32) mx = fmaxf(mx, z[c]);
        mx = warp_max(mx);
        float se = 0.f;
        for (int c = lane; c < C; c += 32) se += expf(z[c] - mx);
        se = warp_sum(se);
        const float lse = mx + logf(se);
        const float zy = z[y];
        int greater = 0, tie_before = 0;
        for (int c = lane; c < C; c += 32) {
            const float zc = z[c];
            greater += zc > zy ? 1 : 0;
            tie_before += (zc == zy && c < y) ? 1 : 0;
        }
        greater = __reduce_add_sync(0xffffffffu, greater);
        tie_before = __reduce_add_sync(0xffffffffu, tie_before);
        if (lane == 0) {
            p.rowlse[n] = lse;
            const int rank = greater + tie_before;
            if (is_sup) { ls = (double)(lse - zy); h1 = rank == 0 ? 1.0 : 0.0; h5 = rank < 5 ? 1.0 : 0.0; }
            else lm = (double)(lse - zy);
        }
    }
    if (lane == 0) { s_part[warp][0] = ls; s_part[warp][1] = lm; s_part[warp][2] = h1; s_part[warp][3] = h5; }
    __syncthreads();
    if (threadIdx.x < 4) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kRowsPerCta; ++w) t += s_part[w][threadIdx.x];
        p.rowpart[(int64_t)blockIdx.x * 4 + threadIdx.x] = t;
    }
}

// dZ^T = (softmax - onehot) / n_rows for a 64-sample x 64-class tile, written class-major as hi / lo planes
// [2][S][Cp][slice] (what the dW GEMM streams as its activations): reads coalesced along classes, writes coalesced along
// samples (two bf16 per thread and store).
__global__ void __launch_bounds__(kT) tc_dz_kernel(const TcParams p) {
    if (p.ctrl->stop) return;
    __shared__ float tile[64][65];
    __shared__ float s_lse[64], s_inv[64];
    __shared__ int s_y[64];
    const sr_head_args& a = p.a;
    const int C = p.g.C, Cp = p.g.Cp, Cz = p.g.Cz, N = p.g.N;
    const int tid = threadIdx.x;
    const int n0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
    const int col = tid & 63, rq = tid >> 6;
    // all 16 loads of this thread in flight before anything depends on them
    float z[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int n = n0 + i * 4 + rq, c = c0 + col;
        z[i] = (n < N && c < C) ? __ldg(p.Z + (int64_t)n * Cz + c) : 0.f;
    }
    if (tid < 64) {
        const int n = n0 + tid;
        const bool ok = n < N, is_sup = n < a.n_support;
        s_lse[tid] = ok ? p.rowlse[n] : 0.f;
        s_y[tid] = ok ? (int)(is_sup ? a.labels_support[n] : a.labels_memory[n - a.n_support]) : -1;
        s_inv[tid] = ok ? 1.f / (float)(is_sup ? a.n_support : a.n_memory) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int r = i * 4 + rq, c = c0 + col;
        float v = 0.f;
        if (n0 + r < N && c < C) v = (expf(z[i] - s_lse[r]) - (c == s_y[r] ? 1.f : 0.f)) * s_inv[r];
        tile[r][col] = v;
    }
    __syncthreads();
    const int s = n0 / p.g.slice, j0 = n0 - s * p.g.slice;   // 64-sample blocks never straddle a slice (slice % 64 == 0)
    const int64_t plane = (int64_t)p.g.S * Cp * p.g.slice;
    const int sp = (tid & 31) * 2, cq = tid >> 5;            // this thread: samples sp, sp + 1 of classes cq, cq + 8, ...
#pragma unroll 4
    for (int i = 0; i < 8; ++i) {
        const int cc = i * 8 + cq, c = c0 + cc;
        if (c < C) {
            __nv_bfloat16 h0, l0, h1, l1;
            split_bf16(tile[sp][cc], h0, l0);
            split_bf16(tile[sp + 1][cc], h1, l1);
            const int64_t o = ((int64_t)s * Cp + c) * p.g.slice + j0 + sp;
            __nv_bfloat162 hh, ll;
            hh.x = h0; hh.y = h1; ll.x = l0; ll.y = l1;
            *reinterpret_cast<__nv_bfloat162*>(p.dZt + o) = hh;
            *reinterpret_cast<__nv_bfloat162*>(p.dZt + plane + o) = ll;
        }
    }
}

// Projection residual / fixed-puller residual of one new class row and its gradient (same maths as head.cu pull_task).
__global__ void __launch_bounds__(kT) tc_pull_kernel(const TcParams p) {
    if (p.ctrl->stop) return;
    extern __shared__ float dyn[];
    __shared__ double red[32];
    const sr_head_args& a = p.a;
    const int d = a.dim, i = blockIdx.x;
    float* sw = dyn;
    float* sr = sw + d;
    float* su = sr + d;
    const float* w = a.weight + (int64_t)(a.n_classes - a.n_new + i) * d;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    double local = 0.0;
    if (a.pull_mode == SR_PULL_FIXED) {
        const float* pl = a.pull + (int64_t)i * d;
        for (int k = tid; k < d; k += blockDim.x) {
            const float r = pl[k] - w[k];
            local += (double)r * (double)r;
            p.gpull[(int64_t)i * d + k] = -2.f * a.gamma * r;
        }
    } else if (a.q_rows >= d) {   // span(base) = R^d: P = I, the regulariser vanishes identically
        for (int k = tid; k < d; k += blockDim.x) p.gpull[(int64_t)i * d + k] = 0.f;
    } else {
        const float* Q = a.pull;
        const int q = a.q_rows;
        for (int k = tid; k < d; k += blockDim.x) sw[k] = w[k];
        __syncthreads();
        for (int j = warp; j < q; j += nwarps) {
            float s = 0.f;
            for (int k = lane; k < d; k += 32) s = fmaf(Q[(int64_t)j * d + k], sw[k], s);
            s = warp_sum(s);
            if (lane == 0) su[j] = s;
        }
        __syncthreads();
        for (int k = tid; k < d; k += blockDim.x) {
            float s = 0.f;
            for (int j = 0; j < q; ++j) s = fmaf(su[j], Q[(int64_t)j * d + k], s);
            const float r = s - sw[k];
            sr[k] = r;
            local += (double)r * (double)r;
        }
        __syncthreads();
        for (int j = warp; j < q; j += nwarps) {
            float s = 0.f;
            for (int k = lane; k < d; k += 32) s = fmaf(Q[(int64_t)j * d + k], sr[k], s);
            s = warp_sum(s);
            if (lane == 0) su[j] = s;
        }
        __syncthreads();
        for (int k = tid; k < d; k += blockDim.x) {
            float s = 0.f;
            for (int j = 0; j < q; ++j) s = fmaf(su[j], Q[(int64_t)j * d + k], s);
            p.gpull[(int64_t)i * d + k] = 2.f * a.gamma * (s - sr[k]);
        }
    }
    const double tot = block_sum(local, red);
    if (tid == 0) p.pull_sq[i] = tot;
}

// Loss of epoch e (pre-update weights) + the reference's stopping rule (language_eval.py:298-318).  One CTA (the extra
// CTA of the update kernel: it runs next to the update CTAs instead of in a launch of its own).
__device__ void tc_assemble_loss(const TcParams& p, int e) {
    const sr_head_args& a = p.a;
    const int tid = threadIdx.x;
    const bool has_base = a.base_weight != nullptr;
    const bool has_prev = a.reserve_weight != nullptr && a.n_prev_novel > 0;
    const int n_pull = a.pull_mode == SR_PULL_NONE ? 0 : a.n_new;
    double ls = 0.0, lm = 0.0, h1 = 0.0, h5 = 0.0, nb = 0.0, nn = 0.0, ps = 0.0;
    for (int i = tid; i < p.g.n_rowctas; i += kT) {
        ls += p.rowpart[(int64_t)i * 4];
        lm += p.rowpart[(int64_t)i * 4 + 1];
        h1 += p.rowpart[(int64_t)i * 4 + 2];
        h5 += p.rowpart[(int64_t)i * 4 + 3];
    }
    for (int i = tid; i < p.g.n_upd; i += kT) { nb += p.nbp[(e & 1) * p.g.n_upd + i]; nn += p.nnp[(e & 1) * p.g.n_upd + i]; }
    for (int i = tid; i < n_pull; i += kT) ps += p.pull_sq[i];
    {   // the seven sums in one pass: warp shuffles, then warp totals added in warp order (deterministic)
        __shared__ double s7[kT / 32][7];
        double v[7] = {ls, lm, h1, h5, nb, nn, ps};
#pragma unroll
        for (int k = 0; k < 7; ++k) v[k] = warp_sum(v[k]);
        if ((tid & 31) == 0)
#pragma unroll
            for (int k = 0; k < 7; ++k) s7[tid >> 5][k] = v[k];
        __syncthreads();
        if (tid == 0) {
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                double t = 0.0;
                for (int w = 0; w < kT / 32; ++w) t += s7[w][k];
                v[k] = t;
            }
            ls = v[0]; lm = v[1]; h1 = v[2]; h5 = v[3]; nb = v[4]; nn = v[5]; ps = v[6];
        }
    }
    if (tid == 0) {
        p.ctrl->norm_base_sq = nb;
        p.ctrl->norm_prev_sq = nn;
        const float ce_s = (float)(ls / (double)a.n_support);
        const float ce_m = a.n_memory > 0 ? (float)(lm / (double)a.n_memory) : 0.f;
        const float reg_b = has_base ? a.lmbd_base * (float)sqrt(nb) : 0.f;
        const float reg_n = has_prev ? a.lmbd_novel * (float)sqrt(nn) : 0.f;
        const float pull = n_pull ? a.gamma * (float)ps : 0.f;
        float loss = ce_s;
        if (a.n_memory > 0) loss += ce_m;
        if (has_base) loss += reg_b;
        if (has_prev) loss += reg_n;
        if (n_pull) loss += pull;
        float* tr = a.loss_trace + (int64_t)e * SR_TRACE_COLS;
        tr[0] = loss; tr[1] = ce_s; tr[2] = ce_m; tr[3] = reg_b; tr[4] = reg_n; tr[5] = pull; tr[6] = (float)h1; tr[7] = (float)h5;
        int stop = 0;
        int sc = p.ctrl->stable_count;
        const float prev = p.ctrl->prev_loss;
        if (a.stable) {
            if (fabs((double)loss - (double)prev) < a.convergence_epsilon) sc += 1; else sc = 0;
            if (sc == a.stable_epochs) stop = 1;
        }
        const int epoch = p.start->epoch0 + e + 1;
        if (epoch >= a.max_novel_epochs || ((double)loss <= a.target_train_loss && epoch >= a.min_novel_epochs + 1)) stop = 1;
        p.ctrl->stable_count = sc;
        p.ctrl->prev_loss = loss;
        p.ctrl->epochs_done = e + 1;
        p.start->stopflag[e & 1] = stop;   // for the update CTAs of epoch e + 1 (this epoch's read the other slot)
        p.ctrl->stop = stop;               // for every other kernel of the later epochs (launched after this one ends)
    }
}

// dW = sum of the S split-K partials (fixed order) + regulariser gradients + weight decay -> optimiser step; the new W
// in fp32, as hi / lo planes for the next logits GEMM, and the drift-norm partials of the new W.
__global__ void __launch_bounds__(kT) tc_update_kernel(const TcParams p, int e) {
    if (p.start->stopflag[(e + 1) & 1]) {   // the rule fired in an earlier epoch (slot written by epoch e - 1 / init)
        if ((int)blockIdx.x == p.g.n_upd && threadIdx.x == 0) p.start->stopflag[e & 1] = 1;   // keep it set for epoch e + 1
        return;
    }
    if ((int)blockIdx.x == p.g.n_upd) {            // the extra CTA: loss of this epoch + stopping rule
        tc_assemble_loss(p, e);
        return;
    }
    __shared__ double red[32];
    const sr_head_args& a = p.a;
    const int d = p.g.d, C = p.g.C, Cp = p.g.Cp;
    const int64_t i = (int64_t)blockIdx.x * kUpdElems + threadIdx.x * 4;
    const bool has_base = a.base_weight != nullptr;
    const bool has_prev = a.reserve_weight != nullptr && a.n_prev_novel > 0;
    const int n_pull = a.pull_mode == SR_PULL_NONE ? 0 : a.n_new;
    // drift norms of W_e: every CTA adds the same partials in the same order (no dependency on the loss CTA)
    double nbs = 0.0, nns = 0.0, unused0 = 0.0;
    for (int t = threadIdx.x; t < p.g.n_upd; t += kT) {
        nbs += p.nbp[(e & 1) * p.g.n_upd + t];
        nns += p.nnp[(e & 1) * p.g.n_upd + t];
    }
    block_sum3(nbs, nns, unused0, red);
    const float nb = has_base ? (float)sqrt(nbs) : 0.f;
    const float np_ = has_prev ? (float)sqrt(nns) : 0.f;
    const float sb = nb > 0.f ? a.lmbd_base / nb : 0.f;    // d||x||/dx = x/||x||, 0 at x = 0 (as torch)
    const float sn = np_ > 0.f ? a.lmbd_novel / np_ : 0.f;
    const int step = p.start->step0 + e;
    float bc1 = 1.f, bc2s = 1.f;
    if (a.optimizer == SR_OPT_ADAM) {
        bc1 = (float)(1.0 - pow((double)a.beta1, (double)(step + 1)));
        bc2s = (float)sqrt(1.0 - pow((double)a.beta2, (double)(step + 1)));
    }
    double nbp = 0.0, nnp = 0.0, unused = 0.0;
    if (i < (int64_t)C * d) {
        const int c = (int)(i / d), k = (int)(i % d);
        float g[4] = {0.f, 0.f, 0.f, 0.f};
        for (int s = 0; s < p.g.S; ++s) {
            const float4 v = *reinterpret_cast<const float4*>(p.part + ((int64_t)s * Cp + c) * d + k);
            g[0] += v.x; g[1] += v.y; g[2] += v.z; g[3] += v.w;
        }
        const float4 w4 = *reinterpret_cast<const float4*>(a.weight + i);
        float w[4] = {w4.x, w4.y, w4.z, w4.w};
        float w0[4] = {0.f, 0.f, 0.f, 0.f}, wr[4] = {0.f, 0.f, 0.f, 0.f};
        const bool in_base = has_base && c < a.n_base;
        const bool in_prev = has_prev && c >= a.n_base && c < a.n_base + a.n_prev_novel;
        if (in_base) {
            const float4 r = *reinterpret_cast<const float4*>(a.base_weight + i);
            w0[0] = r.x; w0[1] = r.y; w0[2] = r.z; w0[3] = r.w;
#pragma unroll
            for (int j = 0; j < 4; ++j) g[j] += sb * (w[j] - w0[j]);
        }
        if (in_prev) {
            const float4 r = *reinterpret_cast<const float4*>(a.reserve_weight + i - (int64_t)a.n_base * d);
            wr[0] = r.x; wr[1] = r.y; wr[2] = r.z; wr[3] = r.w;
#pragma unroll
            for (int j = 0; j < 4; ++j) g[j] += sn * (w[j] - wr[j]);
        }
        if (n_pull && c >= C - a.n_new) {
            const float4 r = *reinterpret_cast<const float4*>(p.gpull + (int64_t)(c - (C - a.n_new)) * d + k);
            g[0] += r.x; g[1] += r.y; g[2] += r.z; g[3] += r.w;
        }
        float wn[4];
        if (a.optimizer == SR_OPT_SGD) {
            const float4 v4 = *reinterpret_cast<const float4*>(a.opt_state + i);
            float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float gj = fmaf(a.weight_decay, w[j], g[j]);
                v[j] = step == 0 ? gj : fmaf(a.momentum, v[j], gj);
                wn[j] = w[j] - a.lr * v[j];
            }
            *reinterpret_cast<float4*>(a.opt_state + i) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
            const int64_t wsize = (int64_t)C * d;
            const float4 a4 = *reinterpret_cast<const float4*>(a.opt_state + i);
            const float4 b4 = *reinterpret_cast<const float4*>(a.opt_state + wsize + i);
            float m1[4] = {a4.x, a4.y, a4.z, a4.w}, m2[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float gj = fmaf(a.weight_decay, w[j], g[j]);
                m1[j] = m1[j] + (1.f - a.beta1) * (gj - m1[j]);             // exp_avg.lerp_(grad, 1 - beta1)
                m2[j] = a.beta2 * m2[j] + (1.f - a.beta2) * gj * gj;
                wn[j] = w[j] - (a.lr / bc1) * (m1[j] / (sqrtf(m2[j]) / bc2s + a.adam_eps));
            }
            *reinterpret_cast<float4*>(a.opt_state + i) = make_float4(m1[0], m1[1], m1[2], m1[3]);
            *reinterpret_cast<float4*>(a.opt_state + wsize + i) = make_float4(m2[0], m2[1], m2[2], m2[3]);
        }
        *reinterpret_cast<float4*>(a.weight + i) = make_float4(wn[0], wn[1], wn[2], wn[3]);
        __align__(8) __nv_bfloat16 hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) split_bf16(wn[j], hi[j], lo[j]);
        *reinterpret_cast<uint2*>(p.Wp + i) = *reinterpret_cast<const uint2*>(hi);
        *reinterpret_cast<uint2*>(p.Wp + (int64_t)Cp * d + i) = *reinterpret_cast<const uint2*>(lo);
        if (in_base)
#pragma unroll
            for (int j = 0; j < 4; ++j) { const float dl = wn[j] - w0[j]; nbp += (double)dl * dl; }
        if (in_prev)
#pragma unroll
            for (int j = 0; j < 4; ++j) { const float dl = wn[j] - wr[j]; nnp += (double)dl * dl; }
    }
    block_sum3(nbp, nnp, unused, red);
    if (threadIdx.x == 0) {
        p.nbp[((e + 1) & 1) * p.g.n_upd + blockIdx.x] = nbp;
        p.nnp[((e + 1) & 1) * p.g.n_upd + blockIdx.x] = nnp;
    }
}

__global__ void tc_finish_kernel(const TcParams p) {
    HeadStart st;
    st.epoch0 = p.start->epoch0;
    st.step0 = p.start->step0;
    st.stable_count0 = p.start->stable_count0;
    st.prev_loss = p.start->prev_loss;
    st.already_stopped = false;
    head_write_status(p.a, st, p.ctrl->epochs_done, p.ctrl->stop, p.ctrl->stable_count, p.ctrl->prev_loss, p.ctrl->error);
}

// ---- scoring (sr_eval_logits): fp32 rows -> hi / lo planes, optionally zero-padded to `rows_out` rows ----
__global__ void __launch_bounds__(kT) tc_split_rows_kernel(const float* __restrict__ src, int rows, int rows_out, int d,
                                                           __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
    const int64_t i = ((int64_t)blockIdx.x * kT + threadIdx.x) * 4;
    if (i >= (int64_t)rows_out * d) return;
    float x[4] = {0.f, 0.f, 0.f, 0.f};
    if (i < (int64_t)rows * d) {
        const float4 v = *reinterpret_cast<const float4*>(src + i);
        x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
    }
    __align__(8) __nv_bfloat16 h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split_bf16(x[j], h[j], l[j]);
    *reinterpret_cast<uint2*>(hi + i) = *reinterpret_cast<const uint2*>(h);
    *reinterpret_cast<uint2*>(lo + i) = *reinterpret_cast<const uint2*>(l);
}

}  // namespace

namespace srb {

// Query / base scoring of large problems (BASELINE config 5: 16 384 x 1100 x 512): the logits GEMM on tcgen05, same
// error-compensated 1x1 convolution as the head's.  Workspace: X planes | W planes (rows padded to 128) | Z [n][Cz].
int64_t eval_tc_workspace_bytes(int n, int dim, int n_classes) {
    if ((int64_t)n * n_classes < (1ll << 20) || dim % 64 != 0 || dim > 2048 || n_classes > 4096) return 0;
    const int64_t Cz = align_up(n_classes, 128);
    return align_up(2ll * n * dim * 2, 256) + align_up(2ll * Cz * dim * 2, 256) + align_up((int64_t)n * Cz * 4, 256);
}

int32_t eval_tc_logits(const sr_eval_args* a, cudaStream_t stream, const float** z_out, int* pitch) {
    const int n = a->n, d = a->dim, C = a->n_classes;
    const int Cz = (int)align_up(C, 128);
    uint8_t* ws = static_cast<uint8_t*>(a->workspace);
    if (reinterpret_cast<uintptr_t>(ws) & 255) return fail(SR_E_ARG, "sr_eval_logits: workspace must be 256-byte aligned");
    __nv_bfloat16* Xp = reinterpret_cast<__nv_bfloat16*>(ws);
    __nv_bfloat16* Wp = reinterpret_cast<__nv_bfloat16*>(ws + align_up(2ll * n * d * 2, 256));
    float* Z = reinterpret_cast<float*>(ws + align_up(2ll * n * d * 2, 256) + align_up(2ll * Cz * d * 2, 256));
    tc_split_rows_kernel<<<(unsigned)(((int64_t)n * d / 4 + kT - 1) / kT), kT, 0, stream>>>(a->feat, n, n, d, Xp, Xp + (int64_t)n * d);
    tc_split_rows_kernel<<<(unsigned)(((int64_t)Cz * d / 4 + kT - 1) / kT), kT, 0, stream>>>(a->weight, C, Cz, d, Wp,
                                                                                             Wp + (int64_t)Cz * d);
    SR_CUDA_OK(cudaGetLastError());
    sr_conv_args g;
    memset(&g, 0, sizeof(g));
    g.batch = n; g.height = 1; g.width = 1; g.cout = Cz; g.n_panels = 1;
    g.max_cout_per_cta = 128;
    g.panel[0].act = Xp; g.panel[0].act_lo = Xp + (int64_t)n * d;
    g.panel[0].wgt = Wp; g.panel[0].wgt_lo = Wp + (int64_t)Cz * d;
    g.panel[0].cin_pad = d; g.panel[0].taps = 1;
    g.epilogue = SR_EPI_RAW_STATS; g.out = Z; g.stats = nullptr;
    const int32_t rc = sr_conv(&g, stream);
    if (rc != SR_OK) return rc;
    *z_out = Z;
    *pitch = Cz;
    return SR_OK;
}

// Large problems whose shapes fit the two GEMM mappings; everything else stays on the SIMT head_kernel.
bool head_tc_applicable(const sr_head_args* a) {
    if (a->logits_support != nullptr) return false;
    if ((int64_t)(a->n_support + a->n_memory) * a->n_classes < (1ll << 20)) return false;   // small: latency-bound kernels win
    if (a->dim > 1024) return false;                                                         // pull kernel's shared buffers
    TcGeom g;
    return tc_geometry(a, &g);
}

int64_t head_tc_workspace_bytes(const sr_head_args* a) {
    TcGeom g;
    return tc_geometry(a, &g) ? g.total : 0;
}

int32_t head_tc_run(const sr_head_args* a, cudaStream_t stream) {
    TcParams p;
    p.a = *a;
    if (!tc_geometry(a, &p.g)) return fail(SR_E_ARG, "sr_head_run: shapes do not fit the tensor-core head");
    const TcGeom& g = p.g;
    if (a->workspace_bytes < g.total)
        return fail(SR_E_SMALLWS, "sr_head_run: workspace %lld < %lld", (long long)a->workspace_bytes, (long long)g.total);
    uint8_t* ws = static_cast<uint8_t*>(a->workspace);
    p.ctrl = reinterpret_cast<HeadCtrl*>(ws + g.ctrl);
    p.start = reinterpret_cast<TcStart*>(ws + g.start);
    p.Xp = reinterpret_cast<__nv_bfloat16*>(ws + g.Xp);
    p.XpT = reinterpret_cast<__nv_bfloat16*>(ws + g.XpT);
    p.Wp = reinterpret_cast<__nv_bfloat16*>(ws + g.Wp);
    p.Z = reinterpret_cast<float*>(ws + g.Z);
    p.dZt = reinterpret_cast<__nv_bfloat16*>(ws + g.dZt);
    p.part = reinterpret_cast<float*>(ws + g.part);
    p.rowlse = reinterpret_cast<float*>(ws + g.rowlse);
    p.rowpart = reinterpret_cast<double*>(ws + g.rowpart);
    p.pull_sq = reinterpret_cast<double*>(ws + g.pull_sq);
    p.gpull = reinterpret_cast<float*>(ws + g.gpull);
    p.nbp = reinterpret_cast<double*>(ws + g.nbp);
    p.nnp = reinterpret_cast<double*>(ws + g.nnp);

    SR_CUDA_OK(cudaMemsetAsync(ws + g.ctrl, 0, (size_t)(g.Xp - g.ctrl), stream));            // control blocks
    SR_CUDA_OK(cudaMemsetAsync(p.dZt, 0, (size_t)(2ll * g.S * g.Cp * g.slice * 2), stream));   // padded classes / samples stay 0
    SR_CUDA_OK(cudaMemsetAsync(p.pull_sq, 0, (size_t)(g.total - g.pull_sq), stream));
    tc_init_kernel<<<1, 1, 0, stream>>>(p);
    tc_split_x_kernel<<<(unsigned)(((int64_t)g.N * g.d / 4 + kT - 1) / kT), kT, 0, stream>>>(p);
    tc_transpose_x_kernel<<<dim3(g.d / 32, g.slice / 32, g.S), kT, 0, stream>>>(p);
    tc_prep_w_kernel<<<(unsigned)(((int64_t)g.Cp * g.d + kUpdElems - 1) / kUpdElems), kT, 0, stream>>>(p);
    SR_CUDA_OK(cudaGetLastError());

    const size_t row_smem = (size_t)kRowsPerCta * g.Cz * 4;
    if (row_smem > 48 * 1024)
        SR_CUDA_OK(cudaFuncSetAttribute(tc_rowstats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)row_smem));
    const int32_t* stop_flag = &p.ctrl->stop;
    sr_conv_args gz;   // logits
    memset(&gz, 0, sizeof(gz));
    gz.batch = g.N; gz.height = 1; gz.width = 1; gz.cout = g.Cz; gz.n_panels = 1;
    gz.max_cout_per_cta = 128;   // 2 accumulator stages: the fp32 epilogue of a tile overlaps the next tile's main loop
    gz.panel[0].act = p.Xp; gz.panel[0].act_lo = p.Xp + (int64_t)g.N * g.d;
    gz.panel[0].wgt = p.Wp; gz.panel[0].wgt_lo = p.Wp + (int64_t)g.Cp * g.d;
    gz.panel[0].cin_pad = g.d; gz.panel[0].taps = 1;
    gz.epilogue = SR_EPI_RAW_STATS; gz.out = p.Z; gz.stats = nullptr; gz.skip_if_nonzero = stop_flag;
    sr_conv_args gw;   // dW split-K partials
    memset(&gw, 0, sizeof(gw));
    gw.batch = g.S; gw.height = 2; gw.width = g.Cp / 2; gw.cout = g.d; gw.n_panels = 1;
    gw.panel[0].act = p.dZt; gw.panel[0].act_lo = p.dZt + (int64_t)g.S * g.Cp * g.slice;
    gw.panel[0].wgt = p.XpT; gw.panel[0].wgt_lo = p.XpT + (int64_t)g.S * g.d * g.slice;
    gw.panel[0].cin_pad = g.slice; gw.panel[0].taps = 1;
    gw.epilogue = SR_EPI_RAW_STATS; gw.out = p.part; gw.stats = nullptr; gw.skip_if_nonzero = stop_flag;
    gw.weights_per_image = 1;
    gw.max_cout_per_cta = 128;

    const int n_pull = a->pull_mode == SR_PULL_NONE ? 0 : a->n_new;
    const int epochs = std::min(a->max_epochs, kMaxEpochsPerCall);
    for (int e = 0; e < epochs; ++e) {
        int32_t rc = sr_conv(&gz, stream);
        if (rc != SR_OK) return rc;
        tc_rowstats_kernel<<<g.n_rowctas, kT, row_smem, stream>>>(p);
        tc_dz_kernel<<<dim3((g.N + 63) / 64, (g.C + 63) / 64), kT, 0, stream>>>(p);
        if (n_pull) tc_pull_kernel<<<n_pull, kT, 3 * g.d * sizeof(float), stream>>>(p);
        rc = sr_conv(&gw, stream);
        if (rc != SR_OK) return rc;
        tc_update_kernel<<<g.n_upd + 1, kT, 0, stream>>>(p, e);   // n_upd update CTAs + the loss / stopping-rule CTA
        SR_CUDA_OK(cudaGetLastError());
    }
    tc_finish_kernel<<<1, 1, 0, stream>>>(p);
    SR_CUDA_OK(cudaGetLastError());
    return SR_OK;
}

}  // namespace srb
