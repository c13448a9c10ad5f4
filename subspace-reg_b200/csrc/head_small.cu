// Latency-optimised persistent head kernel for the paper-sized problems (C <= 128 classes, a few hundred rows):
// the step is ~1.5 MB / 30-90 MFLOP, far below what one kernel launch costs, so everything that does not change
// between epochs stays in shared memory for the whole session and an epoch is TWO grid phases:
//
//   phase 1  row CTAs (R rows of X resident, k-major): stream W^T (k-major copy kept by phase 2) through a cp.async ring,
//            logits as 4-class x R-row register tiles per thread with the k range split over the warps, softmax,
//            CE / top-k per row, dlogits -> DL[n][c] (sample-major);
//            pull CTA (Q^T resident, 153 KB): u = W_new Q  (projection coefficients of the newest rows);
//            CTA 0: assembles the loss of the PREVIOUS epoch and applies the stopping rule (language_eval.py:298-318)
//   phase 2  column CTAs (8 feature columns of X, W, momentum, W0, reserve, Q^T resident): stream DL, dW = dZ^T X as
//            4-class x 8-column register tiles with the sample range split over the warps,
//            every regulariser gradient, weight decay, SGD-momentum / Adam, write W; partial sums of
//            ||W - W0||^2, ||W_prev - W_res||^2 for the next epoch and ||P w - w||^2 for this one.
//
// Same arithmetic as head_kernel (head.cu) except that the projection gradient uses the closed form 2*gamma*(w - Pw)
// (P is an orthogonal projector; the reference's autograd expression 2*gamma*(rP - r) differs by O(1e-8)).
#include <algorithm>
#include <cstdlib>
#include <mutex>
#include "common.h"
#include "head_common.cuh"
#include "ptx.cuh"

namespace {
using namespace srb;

constexpr int kT = 256;
// feature columns per column CTA (template parameter DC): 8 by default (16 measured ~10 % slower per epoch: the update and
// the tile arithmetic double); 16 when the caller caps the launch (sr_head_args.cta_budget) so that several runs' head
// loops are co-resident on one GPU - a cooperative launch needs all its CTAs resident at once, and two 82-92-CTA launches do
// not fit on 148 SMs, three 47-CTA launches do.
constexpr int KC = 64;    // k-chunk of the W^T stream (phase 1) / n-chunk of the DL stream (phase 2): KC rows of CP floats
constexpr int NS = 4;     // stages of the W^T / DL stream ring (one region, the two phases never overlap in a CTA).  Measured on
                          // B200: a runtime depth of 6-8 stages, 128-row chunks, 512 threads per CTA, 16 columns per column CTA,
                          // a per-CTA rotation of the chunk order and cluster-multicast copies are all 5-20 % slower than this.

struct SmallParams {
    sr_head_args a;
    int n_total, ldn;      // ldn: row pitch of DLt (n_total rounded up to 32)
    int R, GA, GC, G;      // rows per row CTA, #row CTAs, #column CTAs, grid (= max(GA, GC) + loss CTA + pull CTA)
    int CP;                // classes padded to 64 or 128 (thread mapping of phase 1)
    HeadCtrl* ctrl;
    float* DL;             // [ldn][cw]  dlogits, sample-major, cw = classes rounded up to 4 (padding stays zero)
    float* Wt;             // [d][cw]    W^T, kept in step with `weight` by the column CTAs
    float* rowloss;        // [2][n_total]
    int* rowhit;           // [2][n_total]
    double* nb_part;       // [2][GC]
    double* nn_part;       // [2][GC]
    double* pull_part;     // [GC]
    float* u;              // [n_new][q_rows]
};

// One bulk asynchronous copy (TMA, 1-D) of `bytes` contiguous bytes into shared memory, completion on an mbarrier.
// Measured: a cp.async (LDGSTS) ring sustains only ~21 B/clk per SM here (outstanding-request limit x L2 latency); the bulk
// engine has no such limit.
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// Loads that stay where they are written (issued before the stream loop, consumed after it: the L2 latency hides behind it).
__device__ __forceinline__ double ldg_early_f64(const double* p) {
    double v;
    asm volatile("ld.global.cg.f64 %0, [%1];\n" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ldg_early_f32(const float* p) {
    float v;
    asm volatile("ld.global.cg.f32 %0, [%1];\n" : "=f"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

__device__ __forceinline__ int64_t feat_row(const sr_head_args& a, int r) {
    return r < a.n_support ? (int64_t)a.support_row0 + r : (int64_t)a.memory_row0 + (r - a.n_support);
}

// Loss of epoch `e` from the partial results in the workspace + the reference's stopping rule.  One CTA.
__device__ void assemble_loss(const SmallParams& p, const HeadStart& st, int e, double* red) {
    const sr_head_args& a = p.a;
    const int par = e & 1;
    const int tid = threadIdx.x;
    double ls = 0.0, lm = 0.0, h1 = 0.0, h5 = 0.0, nb = 0.0, nn = 0.0, ps = 0.0;
    for (int r = tid; r < p.n_total; r += blockDim.x) {
        const float l = p.rowloss[par * p.n_total + r];
        const int h = p.rowhit[par * p.n_total + r];
        if (r < a.n_support) { ls += (double)l; h1 += (double)(h & 1); h5 += (double)((h >> 1) & 1); }
        else lm += (double)l;
    }
    for (int i = tid; i < p.GC; i += blockDim.x) {
        nb += p.nb_part[par * p.GC + i];
        nn += p.nn_part[par * p.GC + i];
        ps += p.pull_part[i];
    }
    ls = block_sum(ls, red); lm = block_sum(lm, red); h1 = block_sum(h1, red); h5 = block_sum(h5, red);
    nb = block_sum(nb, red); nn = block_sum(nn, red); ps = block_sum(ps, red);
    if (tid == 0) {
        const bool has_base = a.base_weight != nullptr;
        const bool has_prev = a.reserve_weight != nullptr && a.n_prev_novel > 0;
        const bool has_pull = a.pull_mode != SR_PULL_NONE;
        const float ce_s = (float)(ls / (double)a.n_support);
        const float ce_m = a.n_memory > 0 ? (float)(lm / (double)a.n_memory) : 0.f;
        const float reg_b = has_base ? a.lmbd_base * (float)sqrt(nb) : 0.f;
        const float reg_n = has_prev ? a.lmbd_novel * (float)sqrt(nn) : 0.f;
        const float pull = has_pull ? a.gamma * (float)ps : 0.f;
        float loss = ce_s;
        if (a.n_memory > 0) loss += ce_m;
        if (has_base) loss += reg_b;
        if (has_prev) loss += reg_n;
        if (has_pull) loss += pull;
        float* tr = a.loss_trace + (int64_t)e * SR_TRACE_COLS;
        tr[0] = loss; tr[1] = ce_s; tr[2] = ce_m; tr[3] = reg_b; tr[4] = reg_n; tr[5] = pull; tr[6] = (float)h1; tr[7] = (float)h5;
        int stop = 0;
        int sc = e == 0 ? st.stable_count0 : p.ctrl->stable_count;
        const float prev = e == 0 ? st.prev_loss : p.ctrl->prev_loss;
        if (a.stable) {
            if (fabs((double)loss - (double)prev) < a.convergence_epsilon) sc += 1; else sc = 0;
            if (sc == a.stable_epochs) stop = 1;
        }
        const int epoch = st.epoch0 + e + 1;
        if (epoch >= a.max_novel_epochs || ((double)loss <= a.target_train_loss && epoch >= a.min_novel_epochs + 1)) stop = 1;
        p.ctrl->stable_count = sc;
        p.ctrl->prev_loss = loss;
        p.ctrl->epochs_done = e + 1;
        p.ctrl->stop = stop;
    }
    __syncthreads();
}

template <int R, int CP, int DC>
__global__ void __launch_bounds__(kT, 1) head_small_kernel(const SmallParams p) {
    extern __shared__ __align__(16) uint8_t dyn[];
    __shared__ double red[32];
    __shared__ int s_stop;
    const sr_head_args& a = p.a;
    const int C = a.n_classes, d = a.dim, NT = p.n_total, q = a.q_rows;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cta = blockIdx.x;
    const bool is_row = cta < p.GA, is_col = cta < p.GC, is_pull = cta == p.G - 1, is_loss = cta == p.G - 2;
    const bool has_base = a.base_weight != nullptr;
    const bool has_prev = a.reserve_weight != nullptr && a.n_prev_novel > 0;
    const bool proj = a.pull_mode == SR_PULL_PROJECT && q < d;   // q >= d: P = I, the term vanishes
    const bool fixed = a.pull_mode == SR_PULL_FIXED;
    const int new0 = C - a.n_new;
    const int n_opt = a.optimizer == SR_OPT_ADAM ? 2 : 1;
    const HeadStart st = head_start(a);
    if (st.already_stopped) {   // chained launch after the stopping rule fired: nothing to do (every CTA takes this exit)
        if (is_loss && tid == 0) head_write_status(a, st, 0, 1, st.stable_count0, st.prev_loss, 0);
        return;
    }

    // ---- shared memory carve-up ----
    float* sp = reinterpret_cast<float*>(dyn);
    float* Xt = sp;              sp += is_row ? R * d : 0;                 // [d][R]  this CTA's rows of X, k-major
    float* Zs = sp;              sp += is_row ? R * (CP + 1) : 0;        // [R][CP+1] logits
    float* Xn = sp;              sp += is_col ? p.ldn * DC : 0;            // [ldn][DC]  this CTA's columns of X, sample-major
    float* Wc = sp;              sp += is_col ? C * DC : 0;                // [C][DC] master copy of this CTA's W columns
    float* Vc = sp;              sp += is_col ? n_opt * C * DC : 0;        // optimiser state columns
    float* W0c = sp;             sp += (is_col && has_base) ? a.n_base * DC : 0;
    float* Rc = sp;              sp += (is_col && has_prev) ? a.n_prev_novel * DC : 0;
    float* Pc = sp;              sp += (is_col && (proj || fixed)) ? (proj ? q : a.n_new) * DC : 0;  // Q^T or puller columns
    float* Us = sp;              sp += (is_col && proj) ? ((a.n_new * q + 3) & ~3) : 0;      // [n_new][q] projection coefficients
    float* Stg = sp;             sp += (is_row || is_col) ? NS * KC * CP : 0;      // NS x [KC][CP] stream stages
    float* Qs = sp;              sp += (is_pull && proj) ? q * d : 0;      // [q][d]
    float* wn = sp;              sp += (is_pull && proj) ? a.n_new * d : 0;

    __shared__ uint64_t stream_bar[NS];   // "chunk landed" barriers of the W^T / DL stream ring
    if (tid == 0) {
        for (int i = 0; i < NS; ++i) mbar_init(&stream_bar[i], 1);
        mbar_fence_init();
    }
    const uint32_t a_bar = smem_u32(&stream_bar[0]), a_stg = smem_u32(Stg);
    constexpr uint32_t kChunkBytes = KC * CP * 4;          // stage stride
    const int cw = (C + 3) & ~3;                            // row pitch (floats) of W^T / DL: classes rounded up to 4, not to CP
    const uint32_t chunk_bytes = (uint32_t)(KC * cw * 4);   // bytes actually streamed per chunk
    uint32_t gchunk = 0;   // chunks consumed so far by this CTA (both streams): stage = gchunk % NS, parity = (gchunk / NS) & 1

    // ---- one-time loads of everything that is constant over the session ----
    const int r0 = cta * R;
    if (is_row) {
        for (int i = tid; i < R * (d / 4); i += kT) {
            const int r = i / (d / 4), k4 = i % (d / 4);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r0 + r < NT) v = *reinterpret_cast<const float4*>(a.feat + feat_row(a, r0 + r) * d + k4 * 4);
            Xt[(k4 * 4 + 0) * R + r] = v.x;
            Xt[(k4 * 4 + 1) * R + r] = v.y;
            Xt[(k4 * 4 + 2) * R + r] = v.z;
            Xt[(k4 * 4 + 3) * R + r] = v.w;
        }
    }
    const int j0 = cta * DC;
    if (is_col) {
        for (int i = tid; i < p.ldn * DC; i += kT) {
            const int n = i / DC, j = i % DC;
            Xn[n * DC + j] = n < NT ? a.feat[feat_row(a, n) * d + j0 + j] : 0.f;
        }
        for (int i = tid; i < C * DC; i += kT) {
            const int c = i / DC, j = i % DC;
            Wc[i] = a.weight[(int64_t)c * d + j0 + j];
            p.Wt[(int64_t)(j0 + j) * cw + c] = Wc[i];
            for (int s = 0; s < n_opt; ++s) Vc[s * C * DC + i] = a.opt_state[(int64_t)s * C * d + (int64_t)c * d + j0 + j];
        }
        if (has_base)
            for (int i = tid; i < a.n_base * DC; i += kT) W0c[i] = a.base_weight[(int64_t)(i / DC) * d + j0 + i % DC];
        if (has_prev)
            for (int i = tid; i < a.n_prev_novel * DC; i += kT) Rc[i] = a.reserve_weight[(int64_t)(i / DC) * d + j0 + i % DC];
        if (proj)
            for (int i = tid; i < q * DC; i += kT) Pc[i] = a.pull[(int64_t)(i / DC) * d + j0 + i % DC];
        if (fixed)
            for (int i = tid; i < a.n_new * DC; i += kT) Pc[i] = a.pull[(int64_t)(i / DC) * d + j0 + i % DC];
    }
    if (is_pull && proj)
        for (int i = tid; i < q * (d / 4); i += kT)
            *reinterpret_cast<float4*>(Qs + i * 4) = *reinterpret_cast<const float4*>(a.pull + (int64_t)i * 4);
    __syncthreads();
    // norm partials of W_0 (buffer parity 0)
    if (is_col) {
        double nb = 0.0, nn = 0.0;
        if (has_base)
            for (int i = tid; i < a.n_base * DC; i += kT) { const float dl = Wc[i] - W0c[i]; nb += (double)dl * dl; }
        if (has_prev)
            for (int i = tid; i < a.n_prev_novel * DC; i += kT) { const float dl = Wc[a.n_base * DC + i] - Rc[i]; nn += (double)dl * dl; }
        nb = block_sum(nb, red);
        nn = block_sum(nn, red);
        if (tid == 0) { p.nb_part[cta] = nb; p.nn_part[cta] = nn; p.pull_part[cta] = 0.0; }
    }
    unsigned int bar_target = 0;
    grid_barrier(p.ctrl, bar_target);

    constexpr int CG = CP / 4;       // 4-class groups (16 or 32)
    constexpr int KS = kT / CG;      // slices of a chunk's KC rows, one per (warp, half-warp) (16 or 8)
    static_assert(KC % KS == 0 && R % 4 == 0, "tile shapes");
    const int cg4 = tid % CG, ks = tid / CG;
    int e = 0;
    bool stopped = false;
    for (; e < a.max_epochs; ++e) {
        const int par = e & 1;
        unsigned long long ts0 = 0, ts1 = 0, ts2 = 0, ts3 = 0;
        if ((cta == 0 || is_loss || is_pull) && tid == 0) ts0 = global_ns();
        // ======================= phase 1 =======================
        if (is_loss && e > 0) assemble_loss(p, st, e - 1, red);   // a CTA of its own: off the row CTAs' critical path
        if (is_row) {
            float acc[R][4];
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
            const int nchunk = d / KC;
            // (all row CTAs read the same chunk at the same time on purpose: rotating the chunk order per CTA measured 18 %
            // slower - concurrent readers of one line are served together by L2)
            auto issue = [&](int ck) {   // a chunk is KC rows of W^T; called by thread 0 only
                const uint32_t st = (gchunk + (uint32_t)ck) % NS;
                mbar_expect_tx_a(a_bar + 8u * st, chunk_bytes);
                bulk_load(a_stg + st * kChunkBytes, p.Wt + (int64_t)ck * KC * cw, chunk_bytes, a_bar + 8u * st);
            };
            if (tid == 0) {
                fence_proxy_async();   // the stages were last touched through the generic proxy (reduction buffers)
                for (int s = 0; s < NS - 1 && s < nchunk; ++s) issue(s);
            }
            for (int ck = 0; ck < nchunk; ++ck) {
                const uint32_t g = gchunk + (uint32_t)ck;
                mbar_wait_a(a_bar + 8u * (g % NS), (g / NS) & 1u);   // chunk ck has landed
                __syncthreads();                                      // everyone is done with chunk ck-1
                if (tid == 0 && ck + NS - 1 < nchunk) issue(ck + NS - 1);   // refills the stage chunk ck-1 used
                const float* wsrc = Stg + (g % NS) * KC * CP + cg4 * 4;
                const float* xsrc = Xt + (ck * KC) * R;
                if (cg4 * 4 < cw)
#pragma unroll
                for (int kk = 0; kk < KC / KS; ++kk) {
                    const int k = ks + kk * KS;
                    const float4 w4 = *reinterpret_cast<const float4*>(wsrc + k * cw);
#pragma unroll
                    for (int r4 = 0; r4 < R / 4; ++r4) {
                        const float4 x4 = *reinterpret_cast<const float4*>(xsrc + k * R + r4 * 4);
                        const float xs[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            acc[r4 * 4 + i][0] = fmaf(xs[i], w4.x, acc[r4 * 4 + i][0]);
                            acc[r4 * 4 + i][1] = fmaf(xs[i], w4.y, acc[r4 * 4 + i][1]);
                            acc[r4 * 4 + i][2] = fmaf(xs[i], w4.z, acc[r4 * 4 + i][2]);
                            acc[r4 * 4 + i][3] = fmaf(xs[i], w4.w, acc[r4 * 4 + i][3]);
                        }
                    }
                }
            }
            gchunk += (uint32_t)nchunk;
            __syncthreads();
            unsigned long long tf0 = 0, tf1 = 0;
            if (cta == 0 && tid == 0) { tf0 = global_ns(); p.ctrl->t_ns[6] += tf0 - ts0; }
            // reduce the KS partial tiles through the (now idle) stage memory: red[ks][r][c]
            {
                float* red_t = Stg;
#pragma unroll
                for (int r = 0; r < R; ++r)
                    *reinterpret_cast<float4*>(red_t + (ks * R + r) * CP + cg4 * 4) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
                __syncthreads();
                for (int o = tid; o < R * CP; o += kT) {
                    const int r = o / CP, c = o % CP;
                    float z = 0.f;
#pragma unroll
                    for (int s2 = 0; s2 < KS; ++s2) z += red_t[(s2 * R + r) * CP + c];
                    Zs[r * (CP + 1) + c] = z;
                }
            }
            __syncthreads();
            if (cta == 0 && tid == 0) { tf1 = global_ns(); p.ctrl->t_ns[7] += tf1 - tf0; }
            // softmax per row: warp w handles rows w, w + 8, ...
            for (int r = warp; r < R; r += kT / 32) {
                const int n = r0 + r;
                if (n >= NT) continue;
                const float* z = Zs + r * (CP + 1);
                const bool is_sup = n < a.n_support;
                const int y = (int)(is_sup ? a.labels_support[n] : a.labels_memory[n - a.n_support]);
                const float inv_n = 1.f / (float)(is_sup ? a.n_support : a.n_memory);
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
                for (int c = lane; c < C; c += 32) {
                    const float pr = expf(z[c] - lse);
                    p.DL[(int64_t)n * cw + c] = (pr - (c == y ? 1.f : 0.f)) * inv_n;
                }
                if (lane == 0) {
                    p.rowloss[par * NT + n] = lse - zy;
                    const int rank = greater + tie_before;
                    p.rowhit[par * NT + n] = (rank == 0 ? 1 : 0) | (rank < 5 ? 2 : 0);
                }
            }
            __syncthreads();
            if (cta == 0 && tid == 0) p.ctrl->t_ns[8] += global_ns() - tf1;
        }
        if (is_pull && proj) {   // u = W_new Q  ([n_new][q])
            for (int i = tid; i < a.n_new * (d / 4); i += kT)
                *reinterpret_cast<float4*>(wn + i * 4) = *reinterpret_cast<const float4*>(a.weight + (int64_t)new0 * d + (int64_t)i * 4);
            __syncthreads();
            if (a.n_new == 5 && d == 640) {
                // paper shape: each lane keeps its 20 k-values of the five new rows in registers; per Q row that leaves
                // 20 shared loads for 100 independent FMAs (one dependent chain per output was latency-bound at 18 us)
                float wr[5][20];
#pragma unroll
                for (int i = 0; i < 5; ++i)
#pragma unroll
                    for (int t = 0; t < 20; ++t) wr[i][t] = wn[i * 640 + t * 32 + lane];
                for (int jq = warp; jq < q; jq += kT / 32) {
                    float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int t = 0; t < 20; ++t) {
                        const float qv = Qs[jq * 640 + t * 32 + lane];
#pragma unroll
                        for (int i = 0; i < 5; ++i) acc[i] = fmaf(qv, wr[i][t], acc[i]);
                    }
#pragma unroll
                    for (int i = 0; i < 5; ++i) {
                        const float tsum = warp_sum(acc[i]);
                        if (lane == 0) p.u[i * q + jq] = tsum;
                    }
                }
            } else {
                for (int o = warp; o < a.n_new * q; o += kT / 32) {
                    const int i = o / q, jq = o % q;
                    float sacc = 0.f;
                    for (int k = lane; k < d; k += 32) sacc = fmaf(Qs[jq * d + k], wn[i * d + k], sacc);
                    sacc = warp_sum(sacc);
                    if (lane == 0) p.u[o] = sacc;
                }
            }
        }
        if (cta == 0 && tid == 0) ts1 = global_ns();
        if (is_loss && tid == 0) p.ctrl->t_ns[4] += global_ns() - ts0;
        if (is_pull && tid == 0) p.ctrl->t_ns[5] += global_ns() - ts0;
        grid_barrier(p.ctrl, bar_target);
        if (cta == 0 && tid == 0) ts2 = global_ns();
        if (tid == 0) s_stop = p.ctrl->stop;
        __syncthreads();
        if (e > 0 && s_stop) { stopped = true; break; }

        // ======================= phase 2 =======================
        if (is_col) {
            // norms of W_e - anchors (partials written at the end of the previous epoch / init) and the projection
            // coefficients of this epoch (written by the pull CTA in phase 1): loaded NOW into registers, reduced / staged
            // after the stream loop, so that their L2 latency and the block sum are off the path to the first chunk
            double nbs = 0.0, nns = 0.0;
            if (tid < p.GC) { nbs = ldg_early_f64(p.nb_part + par * p.GC + tid); nns = ldg_early_f64(p.nn_part + par * p.GC + tid); }
            const int n_u = proj ? a.n_new * q : 0;
            float u_early[2] = {0.f, 0.f};
            if (tid < n_u) u_early[0] = ldg_early_f32(p.u + tid);
            if (tid + kT < n_u) u_early[1] = ldg_early_f32(p.u + tid + kT);
            constexpr int CGU = kT / DC;   // class groups of the update mapping
            constexpr int CPT = CP / CGU;  // classes per thread in the update: cg, cg + CGU, ...
            const int j = tid % DC, cg = tid / DC;
            unsigned long long tg0 = 0, tg1 = 0;
            if (cta == 0 && tid == 0) { tg0 = global_ns(); p.ctrl->t_ns[9] += tg0 - ts2; }
            float acc[4][DC];
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int jj = 0; jj < DC; ++jj) acc[c][jj] = 0.f;
            const int nchunk = p.ldn / KC;
            auto issue = [&](int ck) {   // a chunk is KC rows of DL; called by thread 0 only
                const uint32_t st = (gchunk + (uint32_t)ck) % NS;
                mbar_expect_tx_a(a_bar + 8u * st, chunk_bytes);
                bulk_load(a_stg + st * kChunkBytes, p.DL + (int64_t)ck * KC * cw, chunk_bytes, a_bar + 8u * st);
            };
            if (tid == 0) {
                fence_proxy_async();
                for (int s = 0; s < NS - 1 && s < nchunk; ++s) issue(s);
            }
            for (int ck = 0; ck < nchunk; ++ck) {
                const uint32_t g = gchunk + (uint32_t)ck;
                mbar_wait_a(a_bar + 8u * (g % NS), (g / NS) & 1u);
                __syncthreads();
                if (tid == 0 && ck + NS - 1 < nchunk) issue(ck + NS - 1);
                const float* dsrc = Stg + (g % NS) * KC * CP + cg4 * 4;
                const float* xsrc = Xn + (ck * KC) * DC;
                if (cg4 * 4 < cw)
#pragma unroll
                for (int kk = 0; kk < KC / KS; ++kk) {
                    const int n = ks + kk * KS;
                    const float4 d4 = *reinterpret_cast<const float4*>(dsrc + n * cw);
                    const float ds[4] = {d4.x, d4.y, d4.z, d4.w};
                    float xs[DC];
#pragma unroll
                    for (int v = 0; v < DC / 4; ++v) {
                        const float4 x4 = *reinterpret_cast<const float4*>(xsrc + n * DC + v * 4);
                        xs[4 * v] = x4.x; xs[4 * v + 1] = x4.y; xs[4 * v + 2] = x4.z; xs[4 * v + 3] = x4.w;
                    }
#pragma unroll
                    for (int c = 0; c < 4; ++c)
#pragma unroll
                        for (int jj = 0; jj < DC; ++jj) acc[c][jj] = fmaf(ds[c], xs[jj], acc[c][jj]);
                }
            }
            gchunk += (uint32_t)nchunk;
            __syncthreads();
            // partial tiles of the KS sample slices -> stage memory: red[ks][c][j]; the update below sums them
            float* red_t = Stg;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
#pragma unroll
                for (int v = 0; v < DC / 4; ++v)
                    *reinterpret_cast<float4*>(red_t + (ks * CP + cg4 * 4 + c) * DC + v * 4) =
                        make_float4(acc[c][4 * v], acc[c][4 * v + 1], acc[c][4 * v + 2], acc[c][4 * v + 3]);
            }
            __syncthreads();
            if (cta == 0 && tid == 0) { tg1 = global_ns(); p.ctrl->t_ns[10] += tg1 - tg0; }
            if (tid < n_u) Us[tid] = u_early[0];
            if (tid + kT < n_u) Us[tid + kT] = u_early[1];
            for (int i = tid + 2 * kT; i < n_u; i += kT) Us[i] = p.u[i];
            {
                double unused = 0.0;
                block_sum3(nbs, nns, unused, red);   // (its barriers also publish Us)
            }
            const float nb = has_base ? (float)sqrt(nbs) : 0.f, nn = has_prev ? (float)sqrt(nns) : 0.f;
            const float sb = nb > 0.f ? a.lmbd_base / nb : 0.f;
            const float sn = nn > 0.f ? a.lmbd_novel / nn : 0.f;
            const int step = st.step0 + e;
            float bc1 = 1.f, bc2s = 1.f;
            if (a.optimizer == SR_OPT_ADAM) {
                bc1 = (float)(1.0 - pow((double)a.beta1, (double)(step + 1)));
                bc2s = (float)sqrt(1.0 - pow((double)a.beta2, (double)(step + 1)));
            }
            double nbp = 0.0, nnp = 0.0, pp = 0.0;
#pragma unroll
            for (int i = 0; i < CPT; ++i) {
                const int c = cg + CGU * i;
                if (c >= C) continue;
                const int idx = c * DC + j;
                const float w = Wc[idx];
                float g = 0.f;
#pragma unroll
                for (int s2 = 0; s2 < KS; ++s2) g += red_t[(s2 * CP + c) * DC + j];
                if (has_base && c < a.n_base) g += sb * (w - W0c[idx]);
                if (has_prev && c >= a.n_base && c < a.n_base + a.n_prev_novel) g += sn * (w - Rc[idx - a.n_base * DC]);
                if (c >= new0 && (proj || fixed)) {
                    float r;
                    if (proj) {
                        // four independent partial sums: one dependent chain of q FMAs sat on the epoch's critical path
                        float pw0 = 0.f, pw1 = 0.f, pw2 = 0.f, pw3 = 0.f;
                        const float* ui = Us + (c - new0) * q;
                        int jq = 0;
                        for (; jq + 4 <= q; jq += 4) {
                            pw0 = fmaf(ui[jq], Pc[jq * DC + j], pw0);
                            pw1 = fmaf(ui[jq + 1], Pc[(jq + 1) * DC + j], pw1);
                            pw2 = fmaf(ui[jq + 2], Pc[(jq + 2) * DC + j], pw2);
                            pw3 = fmaf(ui[jq + 3], Pc[(jq + 3) * DC + j], pw3);
                        }
                        for (; jq < q; ++jq) pw0 = fmaf(ui[jq], Pc[jq * DC + j], pw0);
                        r = ((pw0 + pw1) + (pw2 + pw3)) - w;
                    } else {
                        r = Pc[(c - new0) * DC + j] - w;
                    }
                    pp += (double)r * (double)r;
                    g += -2.f * a.gamma * r;
                }
                g = fmaf(a.weight_decay, w, g);
                float wnew;
                if (a.optimizer == SR_OPT_SGD) {
                    float v = Vc[idx];
                    v = step == 0 ? g : fmaf(a.momentum, v, g);
                    Vc[idx] = v;
                    wnew = w - a.lr * v;
                } else {
                    float m1 = Vc[idx], m2 = Vc[C * DC + idx];
                    m1 = m1 + (1.f - a.beta1) * (g - m1);
                    m2 = a.beta2 * m2 + (1.f - a.beta2) * g * g;
                    Vc[idx] = m1;
                    Vc[C * DC + idx] = m2;
                    wnew = w - (a.lr / bc1) * (m1 / (sqrtf(m2) / bc2s + a.adam_eps));
                }
                Wc[idx] = wnew;
                a.weight[(int64_t)c * d + j0 + j] = wnew;
                p.Wt[(int64_t)(j0 + j) * cw + c] = wnew;
                if (has_base && c < a.n_base) { const float dl = wnew - W0c[idx]; nbp += (double)dl * dl; }
                if (has_prev && c >= a.n_base && c < a.n_base + a.n_prev_novel) {
                    const float dl = wnew - Rc[idx - a.n_base * DC];
                    nnp += (double)dl * dl;
                }
            }
            block_sum3(nbp, nnp, pp, red);
            if (tid == 0) {
                p.nb_part[(par ^ 1) * p.GC + cta] = nbp;
                p.nn_part[(par ^ 1) * p.GC + cta] = nnp;
                p.pull_part[cta] = pp;
            }
            if (cta == 0 && tid == 0) p.ctrl->t_ns[11] += global_ns() - tg1;
        }
        if (cta == 0 && tid == 0) ts3 = global_ns();
        grid_barrier(p.ctrl, bar_target);
        if (cta == 0 && tid == 0) {
            p.ctrl->t_ns[0] += ts1 - ts0;
            p.ctrl->t_ns[1] += ts2 - ts1;
            p.ctrl->t_ns[2] += ts3 - ts2;
            p.ctrl->t_ns[3] += global_ns() - ts3;
        }
    }
    // ---- tail: loss of the last applied epoch when the loop ran out of epochs; write back optimiser state ----
    if (!stopped && is_loss && e > 0) assemble_loss(p, st, e - 1, red);
    if (is_col) {
        for (int i = tid; i < C * DC; i += kT) {
            const int c = i / DC, j = i % DC;
            for (int s = 0; s < n_opt; ++s) a.opt_state[(int64_t)s * C * d + (int64_t)c * d + j0 + j] = Vc[s * C * DC + i];
        }
    }
    if (is_loss) {
        __syncthreads();
        if (tid == 0) {
            const int done = p.ctrl->epochs_done;
            head_write_status(a, st, done, p.ctrl->stop, done > 0 ? p.ctrl->stable_count : st.stable_count0,
                              done > 0 ? p.ctrl->prev_loss : st.prev_loss, p.ctrl->error);
        }
    }
}

// Rows of X per row CTA: as few as keeps the grid within one wave (every row CTA streams all of W each epoch, so
// fewer rows per CTA buy parallelism with L2 traffic).
int pick_rows(int nt, int budget) {
    if (const char* e = getenv("SRB_HEAD_ROWS")) {   // A/B timing only
        const int v = atoi(e);
        if (v == 4 || v == 8 || v == 16) return v;
    }
    int r = nt <= 384 ? 4 : (nt <= 768 ? 8 : 16);
    if (budget > 0)
        while (r < 16 && (nt + r - 1) / r + 2 > budget) r *= 2;
    return r;
}
// Feature columns per column CTA: 8 unless the caller's CTA budget asks for fewer, fatter CTAs.
int pick_cols(int d, int budget) {
    int dc = 8;
    if (budget > 0)
        while (dc < 16 && d / dc + 2 > budget) dc *= 2;
    return dc;
}
// SRB_HEAD_CTAS overrides sr_head_args.cta_budget (A/B timing).
int cta_budget(const sr_head_args* a) {
    if (const char* e = getenv("SRB_HEAD_CTAS")) return atoi(e);
    return a->cta_budget;
}

struct SmallLayout {
    int64_t ctrl, DL, Wt, rowloss, rowhit, nb, nn, pull, u, total;
};

SmallLayout small_layout(const sr_head_args* a, int GC) {
    SmallLayout L;
    const int64_t nt = (int64_t)a->n_support + a->n_memory;
    const int64_t ldn = align_up(nt, KC);
    int64_t off = 0;
    L.ctrl = off;    off += align_up(sizeof(HeadCtrl), 256);
    const int64_t cp = a->n_classes <= 64 ? 64 : 128;
    L.DL = off;      off += align_up(ldn * cp * 4, 256);
    L.Wt = off;      off += align_up((int64_t)a->dim * cp * 4, 256);
    L.rowloss = off; off += align_up(2 * nt * 4, 256);
    L.rowhit = off;  off += align_up(2 * nt * 4, 256);
    L.nb = off;      off += align_up(2 * (int64_t)GC * 8, 256);
    L.nn = off;      off += align_up(2 * (int64_t)GC * 8, 256);
    L.pull = off;    off += align_up((int64_t)GC * 8, 256);
    L.u = off;       off += align_up((int64_t)std::max(a->n_new, 1) * std::max(a->q_rows, 1) * 4, 256);
    L.total = off;
    return L;
}

template <int R, int CP, int DC>
int32_t launch_small(const SmallParams& p, size_t dyn, cudaStream_t stream) {
    static PerDeviceOnce once;
    once_per_device(once, [] {
        cudaFuncSetAttribute(head_small_kernel<R, CP, DC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    });
    void* args[] = {const_cast<SmallParams*>(&p)};
    SR_CUDA_OK(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(head_small_kernel<R, CP, DC>), dim3(p.G), dim3(kT), args,
                                           dyn, stream));
    return SR_OK;
}
template <int CP, int DC>
int32_t launch_rows(const SmallParams& p, size_t dyn, cudaStream_t stream) {
    if (p.R == 4) return launch_small<4, CP, DC>(p, dyn, stream);
    return p.R == 8 ? launch_small<8, CP, DC>(p, dyn, stream) : launch_small<16, CP, DC>(p, dyn, stream);
}

size_t small_smem_bytes(const sr_head_args* a, int R, int CP, int DC, int ldn, bool pull_cta) {
    const int C = a->n_classes, d = a->dim;
    const bool proj = a->pull_mode == SR_PULL_PROJECT && a->q_rows < d;
    const bool fixed = a->pull_mode == SR_PULL_FIXED;
    const int n_opt = a->optimizer == SR_OPT_ADAM ? 2 : 1;
    size_t f = 0;
    f += (size_t)R * d + (size_t)R * (CP + 1);                                       // row role
    f += (size_t)NS * KC * CP;                                                       // stream stages (shared by both roles)
    f += (size_t)DC * ldn + (size_t)C * DC * (1 + n_opt) + (size_t)a->n_base * DC + (size_t)a->n_prev_novel * DC +
         (size_t)(proj ? a->q_rows : (fixed ? a->n_new : 0)) * DC + (size_t)(proj ? a->n_new * a->q_rows : 0) + 8;   // column role
    size_t g = pull_cta && proj ? (size_t)a->q_rows * d + (size_t)a->n_new * d : 0;    // pull role (its own CTA)
    return std::max(f, g) * sizeof(float);
}
}  // namespace

namespace srb {

// Shapes the resident-operand kernel handles; anything else goes to the tiled head_kernel.
bool head_small_applicable(const sr_head_args* a) {
    const int nt = a->n_support + a->n_memory;
    if (a->logits_support != nullptr) return false;
    if (a->n_classes > 128 || a->dim % 64 != 0 || a->dim > 1024 || nt > 1024) return false;
    if (a->pull_mode == SR_PULL_PROJECT && a->q_rows < a->dim && (a->q_rows > 256 || a->n_new > 16)) return false;
    const int budget = cta_budget(a);
    int R = pick_rows(nt, budget), DC = pick_cols(a->dim, budget);
    const int CP = a->n_classes <= 64 ? 64 : 128;
    const int ldn = (int)align_up(nt, KC);
    if (small_smem_bytes(a, R, CP, DC, ldn, true) > 200 * 1024) {   // the budget's shape does not fit: the default one decides
        R = pick_rows(nt, 0);
        DC = pick_cols(a->dim, 0);
    }
    if ((nt + R - 1) / R > 146 || a->dim / DC > 146) return false;
    return small_smem_bytes(a, R, CP, DC, ldn, true) <= 200 * 1024;
}

// (sized for the widest grid: 8 columns per column CTA)
int64_t head_small_workspace_bytes(const sr_head_args* a) { return small_layout(a, a->dim / 8).total; }

// The launch shape head_small_run will use (host only; sr_head_plan).
void head_small_shape(const sr_head_args* a, int* rows_per_cta, int* cols_per_cta, int* ctas) {
    const int nt = a->n_support + a->n_memory;
    const int CP = a->n_classes <= 64 ? 64 : 128;
    const int ldn = (int)align_up(nt, KC);
    const int budget = cta_budget(a);
    int R = pick_rows(nt, budget), DCc = pick_cols(a->dim, budget);
    if (small_smem_bytes(a, R, CP, DCc, ldn, true) > 200 * 1024) {
        R = pick_rows(nt, 0);
        DCc = pick_cols(a->dim, 0);
    }
    *rows_per_cta = R;
    *cols_per_cta = DCc;
    *ctas = std::max((nt + R - 1) / R, a->dim / DCc) + 2;
}

int32_t head_small_run(const sr_head_args* a, cudaStream_t stream) {
    SmallParams p;
    p.a = *a;
    p.n_total = a->n_support + a->n_memory;
    p.ldn = (int)align_up(p.n_total, KC);
    p.CP = a->n_classes <= 64 ? 64 : 128;
    const int budget = cta_budget(a);
    p.R = pick_rows(p.n_total, budget);
    int DC = pick_cols(a->dim, budget);
    if (small_smem_bytes(a, p.R, p.CP, DC, p.ldn, true) > 200 * 1024) {   // as in head_small_applicable
        p.R = pick_rows(p.n_total, 0);
        DC = pick_cols(a->dim, 0);
    }
    p.GA = (p.n_total + p.R - 1) / p.R;
    p.GC = a->dim / DC;
    p.G = std::max(p.GA, p.GC) + 2;   // + one CTA for the loss / stopping rule, one for the projection coefficients
    const SmallLayout L = small_layout(a, p.GC);
    if (a->workspace_bytes < L.total) return fail(SR_E_SMALLWS, "sr_head_run: workspace %lld < %lld",
                                                  (long long)a->workspace_bytes, (long long)L.total);
    uint8_t* ws = static_cast<uint8_t*>(a->workspace);
    p.ctrl = reinterpret_cast<HeadCtrl*>(ws + L.ctrl);
    p.DL = reinterpret_cast<float*>(ws + L.DL);
    p.Wt = reinterpret_cast<float*>(ws + L.Wt);
    p.rowloss = reinterpret_cast<float*>(ws + L.rowloss);
    p.rowhit = reinterpret_cast<int*>(ws + L.rowhit);
    p.nb_part = reinterpret_cast<double*>(ws + L.nb);
    p.nn_part = reinterpret_cast<double*>(ws + L.nn);
    p.pull_part = reinterpret_cast<double*>(ws + L.pull);
    p.u = reinterpret_cast<float*>(ws + L.u);
    SR_CUDA_OK(cudaMemsetAsync(ws, 0, (size_t)L.total, stream));   // control block, DL / W^T padding, partials
    const size_t dyn = small_smem_bytes(a, p.R, p.CP, DC, p.ldn, true);
    if (p.CP == 64) {
        return DC == 8 ? launch_rows<64, 8>(p, dyn, stream) : launch_rows<64, 16>(p, dyn, stream);
    }
    return DC == 8 ? launch_rows<128, 8>(p, dyn, stream) : launch_rows<128, 16>(p, dyn, stream);
}

}  // namespace srb
