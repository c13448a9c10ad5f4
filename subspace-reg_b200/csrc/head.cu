// Fused classifier-head fine-tuning, scoring and the subspace factor (all fp32 / fp64 SIMT: this is the
// exact-parity tier, 1e-5 against the reference's fp32 autograd path).
//
// sr_head_run      one persistent cooperative kernel per session; every epoch of the reference loop
//                  (eval/language_eval.py:242-318) is three grid-wide phases separated by a global barrier:
//                    A  Z = X W^T for support+memory rows (tiled SGEMM)   |  ||W[:nb]-W0||^2, ||W_prev-W_res||^2,
//                       projection residual r = P w - w and its gradient for the new rows  (independent tasks)
//                    B  row softmax -> per-row CE, dlogits (in place), top-1/top-5 hits
//                    C  loss assembly + stopping rule (one thread), dW = dlogits^T X fused with all regulariser
//                       gradients, weight decay and the SGD-momentum / Adam update of W
// sr_eval_logits   Z = X W^T then per-row argmax / top-5 / CE / confusion counts.
// sr_subspace_factor  Gram (fp64, warp-shuffle dot products) -> single-CTA Cholesky -> Qt = L^-1 B.
#include <cooperative_groups.h>
#include <algorithm>
#include <mutex>
#include "common.h"
#include "head_common.cuh"

namespace cg = cooperative_groups;

namespace {

using namespace srb;

constexpr int kHeadThreads = 256;
constexpr int BK = 16;

// ------------------------------------------------------------------------------------------------
// SGEMM tiles: 256 threads as 16 x 16, each owning a (BM/16) x (BN/16) micro-tile.
// ------------------------------------------------------------------------------------------------
struct RowMap {  // virtual row r of [support ; memory] -> row of the feature cache
    int n_support, support_row0, memory_row0;
    __device__ __forceinline__ int64_t operator()(int r) const {
        return r < n_support ? (int64_t)support_row0 + r : (int64_t)memory_row0 + (r - n_support);
    }
};

template <int BM, int BN>
struct TileSmem {
    float a[BK][BM + 4];
    float b[BK][BN + 4];
};

// acc[i][j] = sum_k A[rowmap(m0 + ty*TM + i), k] * Bm[n0 + tx*TN + j, k]   (both operands K-contiguous)
template <int BM, int BN>
__device__ __forceinline__ void gemm_nt_tile(const float* __restrict__ A, RowMap rm, int M, const float* __restrict__ Bm,
                                             int N, int K, int m0, int n0, TileSmem<BM, BN>& s,
                                             float (&acc)[BM / 16][BN / 16]) {
    constexpr int TM = BM / 16, TN = BN / 16;
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < K; k0 += BK) {
        // each float4 covers 4 consecutive k of one row
        for (int idx = tid; idx < BM * (BK / 4); idx += kHeadThreads) {
            const int r = idx / (BK / 4), kq = (idx % (BK / 4)) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m0 + r < M && k0 + kq < K) v = *reinterpret_cast<const float4*>(A + rm(m0 + r) * K + k0 + kq);
            s.a[kq][r] = v.x; s.a[kq + 1][r] = v.y; s.a[kq + 2][r] = v.z; s.a[kq + 3][r] = v.w;
        }
        for (int idx = tid; idx < BN * (BK / 4); idx += kHeadThreads) {
            const int r = idx / (BK / 4), kq = (idx % (BK / 4)) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n0 + r < N && k0 + kq < K) v = *reinterpret_cast<const float4*>(Bm + (int64_t)(n0 + r) * K + k0 + kq);
            s.b[kq][r] = v.x; s.b[kq + 1][r] = v.y; s.b[kq + 2][r] = v.z; s.b[kq + 3][r] = v.w;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float av[TM], bv[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) av[i] = s.a[kk][ty * TM + i];
#pragma unroll
            for (int j = 0; j < TN; ++j) bv[j] = s.b[kk][tx * TN + j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
}

// acc[i][j] = sum_r DL[r, m0 + ty*TM + i] * X[rowmap(r), n0 + tx*TN + j]     (reduction over rows r < R)
template <int BM, int BN>
__device__ __forceinline__ void gemm_tn_tile(const float* __restrict__ DL, int ldd, int M, const float* __restrict__ X,
                                             RowMap rm, int N, int R, int m0, int n0, TileSmem<BM, BN>& s,
                                             float (&acc)[BM / 16][BN / 16]) {
    constexpr int TM = BM / 16, TN = BN / 16;
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    for (int r0 = 0; r0 < R; r0 += BK) {
        for (int idx = tid; idx < BK * BM; idx += kHeadThreads) {
            const int kk = idx / BM, c = idx % BM;
            float v = 0.f;
            if (r0 + kk < R && m0 + c < M) v = DL[(int64_t)(r0 + kk) * ldd + m0 + c];
            s.a[kk][c] = v;
        }
        for (int idx = tid; idx < BK * (BN / 4); idx += kHeadThreads) {
            const int kk = idx / (BN / 4), cq = (idx % (BN / 4)) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r0 + kk < R && n0 + cq < N) v = *reinterpret_cast<const float4*>(X + rm(r0 + kk) * N + n0 + cq);
            *reinterpret_cast<float4*>(&s.b[kk][cq]) = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float av[TM], bv[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) av[i] = s.a[kk][ty * TM + i];
#pragma unroll
            for (int j = 0; j < TN; ++j) bv[j] = s.b[kk][tx * TN + j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// Persistent head kernel
// ------------------------------------------------------------------------------------------------
struct HeadParams {
    sr_head_args a;
    int n_total;
    HeadCtrl* ctrl;
    float* rowloss;   // [n_total]
    int* rowhit;      // [n_total] bit0 = top-1 hit, bit1 = top-5 hit
    double* pull_sq;  // [n_new]
    float* gpull;     // [n_new, dim]
    float* Z;         // [n_total, n_classes]
};

// Projection residual and gradient of gamma*||pull - w||^2 for new-class row `i` (one CTA).
__device__ void pull_task(const HeadParams& p, int i, float* sw, float* sr, float* su, double* red) {
    const sr_head_args& a = p.a;
    const int d = a.dim;
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
    } else if (a.q_rows >= d) {  // span(base) = R^d: P = I, the regulariser vanishes identically
        for (int k = tid; k < d; k += blockDim.x) p.gpull[(int64_t)i * d + k] = 0.f;
    } else {
        const float* Q = a.pull;  // [q_rows, d], orthonormal rows
        const int q = a.q_rows;
        for (int k = tid; k < d; k += blockDim.x) sw[k] = w[k];
        __syncthreads();
        for (int j = warp; j < q; j += nwarps) {  // u = Q w
            float s = 0.f;
            for (int k = lane; k < d; k += 32) s = fmaf(Q[(int64_t)j * d + k], sw[k], s);
            s = warp_sum(s);
            if (lane == 0) su[j] = s;
        }
        __syncthreads();
        for (int k = tid; k < d; k += blockDim.x) {  // r = Q^T u - w
            float s = 0.f;
            for (int j = 0; j < q; ++j) s = fmaf(su[j], Q[(int64_t)j * d + k], s);
            const float r = s - sw[k];
            sr[k] = r;
            local += (double)r * (double)r;
        }
        __syncthreads();
        for (int j = warp; j < q; j += nwarps) {  // u2 = Q r
            float s = 0.f;
            for (int k = lane; k < d; k += 32) s = fmaf(Q[(int64_t)j * d + k], sr[k], s);
            s = warp_sum(s);
            if (lane == 0) su[j] = s;
        }
        __syncthreads();
        for (int k = tid; k < d; k += blockDim.x) {  // g = 2 gamma (P r - r)   (autograd through both uses of w)
            float s = 0.f;
            for (int j = 0; j < q; ++j) s = fmaf(su[j], Q[(int64_t)j * d + k], s);
            p.gpull[(int64_t)i * d + k] = 2.f * a.gamma * (s - sr[k]);
        }
    }
    const double tot = block_sum(local, red);
    if (tid == 0) p.pull_sq[i] = tot;
    __syncthreads();
}

__device__ void diffnorm_task(const float* w, const float* ref, int64_t n, double* out, double* red) {
    double local = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        const float dlt = w[i] - ref[i];
        local += (double)dlt * (double)dlt;
    }
    const double tot = block_sum(local, red);
    if (threadIdx.x == 0) *out = tot;
    __syncthreads();
}

template <int BT>
__global__ void __launch_bounds__(kHeadThreads) head_kernel(const HeadParams p) {
    extern __shared__ __align__(16) uint8_t dyn_smem[];
    __shared__ TileSmem<BT, BT> tile;
    __shared__ double red[32];
    __shared__ int s_flag;
    const sr_head_args& a = p.a;
    const int C = a.n_classes, d = a.dim, NT = p.n_total;
    const RowMap rm{a.n_support, a.support_row0, a.memory_row0};
    float* sw = reinterpret_cast<float*>(dyn_smem);
    float* sr = sw + d;
    float* su = sr + d;
    unsigned int bar_target = 0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int TM = BT / 16;

    const int tilesA_m = (NT + BT - 1) / BT, tilesA_n = (C + BT - 1) / BT;
    const int n_gemmA = tilesA_m * tilesA_n;
    const int n_pull = a.pull_mode == SR_PULL_NONE ? 0 : a.n_new;
    const int has_base = a.base_weight != nullptr ? 1 : 0;
    const int has_prev = (a.reserve_weight != nullptr && a.n_prev_novel > 0) ? 1 : 0;
    const int n_tasksA = n_gemmA + n_pull + has_base + has_prev;
    const int tilesC_m = (C + BT - 1) / BT, tilesC_n = (d + BT - 1) / BT;
    const int n_tasksC = tilesC_m * tilesC_n;
    const HeadStart st = head_start(a);
    if (st.already_stopped) {   // chained launch after the stopping rule fired: nothing to do (every CTA takes this exit)
        if (blockIdx.x == 0 && tid == 0) head_write_status(a, st, 0, 1, st.stable_count0, st.prev_loss, 0);
        return;
    }

    for (int e = 0; e < a.max_epochs; ++e) {
        // ============================ phase A ============================
        // Small independent tasks first (they are the longest single-CTA chains), spread from the END of the grid.
        for (int task = (int)(gridDim.x - 1 - blockIdx.x); task < n_tasksA; task += gridDim.x) {
            if (task < n_pull) {
                pull_task(p, task, sw, sr, su, red);
            } else if (task < n_pull + has_base) {
                diffnorm_task(a.weight, a.base_weight, (int64_t)a.n_base * d, &p.ctrl->norm_base_sq, red);
            } else if (task < n_pull + has_base + has_prev) {
                diffnorm_task(a.weight + (int64_t)a.n_base * d, a.reserve_weight, (int64_t)a.n_prev_novel * d,
                              &p.ctrl->norm_prev_sq, red);
            } else {
                const int t = task - (n_pull + has_base + has_prev);
                const int m0 = (t / tilesA_n) * BT, n0 = (t % tilesA_n) * BT;
                float acc[TM][TM];
                gemm_nt_tile<BT, BT>(a.feat, rm, NT, a.weight, C, d, m0, n0, tile, acc);
                const int tx = tid & 15, ty = tid >> 4;
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TM; ++j) {
                        const int r = m0 + ty * TM + i, c = n0 + tx * TM + j;
                        if (r < NT && c < C) p.Z[(int64_t)r * C + c] = acc[i][j];
                    }
            }
        }
        grid_barrier(p.ctrl, bar_target);

        // ============================ phase B: row softmax ============================
        {
            const int warps_total = gridDim.x * (kHeadThreads / 32);
            for (int r = blockIdx.x * (kHeadThreads / 32) + warp; r < NT; r += warps_total) {
                float* z = p.Z + (int64_t)r * C;
                const bool is_sup = r < a.n_support;
                const int y = (int)(is_sup ? a.labels_support[r] : a.labels_memory[r - a.n_support]);
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
                if (is_sup && a.logits_support != nullptr)
                    for (int c = lane; c < C; c += 32) a.logits_support[(int64_t)r * C + c] = z[c];
                for (int c = lane; c < C; c += 32) {
                    const float pr = expf(z[c] - lse);
                    z[c] = (pr - (c == y ? 1.f : 0.f)) * inv_n;
                }
                if (lane == 0) {
                    p.rowloss[r] = lse - zy;
                    const int rank = greater + tie_before;
                    p.rowhit[r] = (rank == 0 ? 1 : 0) | (rank < 5 ? 2 : 0);
                }
            }
        }
        grid_barrier(p.ctrl, bar_target);

        // ============================ phase C ============================
        if (blockIdx.x == 0) {  // loss assembly + stopping rule (language_eval.py:298-318)
            double ls = 0.0, lm = 0.0, h1 = 0.0, h5 = 0.0;
            for (int r = tid; r < NT; r += blockDim.x) {
                if (r < a.n_support) {
                    ls += (double)p.rowloss[r];
                    h1 += (double)(p.rowhit[r] & 1);
                    h5 += (double)((p.rowhit[r] >> 1) & 1);
                } else {
                    lm += (double)p.rowloss[r];
                }
            }
            ls = block_sum(ls, red);
            lm = block_sum(lm, red);
            h1 = block_sum(h1, red);
            h5 = block_sum(h5, red);
            double ps = 0.0;
            for (int i = tid; i < n_pull; i += blockDim.x) ps += p.pull_sq[i];
            ps = block_sum(ps, red);
            if (tid == 0) {
                const float ce_s = (float)(ls / (double)a.n_support);
                const float ce_m = a.n_memory > 0 ? (float)(lm / (double)a.n_memory) : 0.f;
                const float reg_b = has_base ? a.lmbd_base * (float)sqrt(p.ctrl->norm_base_sq) : 0.f;
                const float reg_n = has_prev ? a.lmbd_novel * (float)sqrt(p.ctrl->norm_prev_sq) : 0.f;
                const float pull = n_pull ? a.gamma * (float)ps : 0.f;
                float loss = ce_s;
                if (a.n_memory > 0) loss += ce_m;
                if (has_base) loss += reg_b;
                if (has_prev) loss += reg_n;
                if (n_pull) loss += pull;
                float* tr = a.loss_trace + (int64_t)e * SR_TRACE_COLS;
                tr[0] = loss; tr[1] = ce_s; tr[2] = ce_m; tr[3] = reg_b; tr[4] = reg_n; tr[5] = pull;
                tr[6] = (float)h1; tr[7] = (float)h5;
                int stop = 0;
                int sc = e == 0 ? st.stable_count0 : p.ctrl->stable_count;
                const float prev = e == 0 ? st.prev_loss : p.ctrl->prev_loss;
                if (a.stable) {
                    if (fabs((double)loss - (double)prev) < a.convergence_epsilon) sc += 1; else sc = 0;
                    if (sc == a.stable_epochs) stop = 1;
                }
                const int epoch = st.epoch0 + e + 1;
                if (epoch >= a.max_novel_epochs ||
                    ((double)loss <= a.target_train_loss && epoch >= a.min_novel_epochs + 1))
                    stop = 1;
                p.ctrl->stable_count = sc;
                p.ctrl->prev_loss = loss;
                p.ctrl->epochs_done = e + 1;
                p.ctrl->stop = stop;
            }
            __syncthreads();
        }
        {
            const float nb = has_base ? (float)sqrt(p.ctrl->norm_base_sq) : 0.f;
            const float np_ = has_prev ? (float)sqrt(p.ctrl->norm_prev_sq) : 0.f;
            const float sb = nb > 0.f ? a.lmbd_base / nb : 0.f;    // d||x||/dx = x/||x||, 0 at x = 0 (as torch)
            const float sn = np_ > 0.f ? a.lmbd_novel / np_ : 0.f;
            const int step = st.step0 + e;  // optimiser steps already taken
            float bc1 = 1.f, bc2s = 1.f;
            if (a.optimizer == SR_OPT_ADAM) {
                bc1 = (float)(1.0 - pow((double)a.beta1, (double)(step + 1)));
                bc2s = (float)sqrt(1.0 - pow((double)a.beta2, (double)(step + 1)));
            }
            const int64_t wsize = (int64_t)C * d;
            for (int t = blockIdx.x; t < n_tasksC; t += gridDim.x) {
                const int m0 = (t / tilesC_n) * BT, n0 = (t % tilesC_n) * BT;
                float acc[TM][TM];
                gemm_tn_tile<BT, BT>(p.Z, C, C, a.feat, rm, d, NT, m0, n0, tile, acc);
                const int tx = tid & 15, ty = tid >> 4;
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TM; ++j) {
                        const int c = m0 + ty * TM + i, k = n0 + tx * TM + j;
                        if (c >= C || k >= d) continue;
                        const int64_t idx = (int64_t)c * d + k;
                        const float w = a.weight[idx];
                        float g = acc[i][j];
                        if (has_base && c < a.n_base) g += sb * (w - a.base_weight[idx]);
                        if (has_prev && c >= a.n_base && c < a.n_base + a.n_prev_novel)
                            g += sn * (w - a.reserve_weight[idx - (int64_t)a.n_base * d]);
                        if (n_pull && c >= C - a.n_new) g += p.gpull[(int64_t)(c - (C - a.n_new)) * d + k];
                        g = fmaf(a.weight_decay, w, g);
                        if (a.optimizer == SR_OPT_SGD) {
                            float v = a.opt_state[idx];
                            v = step == 0 ? g : fmaf(a.momentum, v, g);
                            a.opt_state[idx] = v;
                            a.weight[idx] = w - a.lr * v;
                        } else {
                            float m1 = a.opt_state[idx], m2 = a.opt_state[wsize + idx];
                            m1 = m1 + (1.f - a.beta1) * (g - m1);             // exp_avg.lerp_(grad, 1 - beta1)
                            m2 = a.beta2 * m2 + (1.f - a.beta2) * g * g;
                            a.opt_state[idx] = m1;
                            a.opt_state[wsize + idx] = m2;
                            const float denom = sqrtf(m2) / bc2s + a.adam_eps;
                            a.weight[idx] = w - (a.lr / bc1) * (m1 / denom);
                        }
                    }
            }
        }
        grid_barrier(p.ctrl, bar_target);
        if (tid == 0) s_flag = p.ctrl->stop;
        __syncthreads();
        if (s_flag) break;
    }
    if (blockIdx.x == 0 && tid == 0) {
        const int done = p.ctrl->epochs_done;
        head_write_status(a, st, done, p.ctrl->stop, done > 0 ? p.ctrl->stable_count : st.stable_count0,
                          done > 0 ? p.ctrl->prev_loss : st.prev_loss, p.ctrl->error);
    }
}

struct HeadLayout {
    int64_t ctrl, rowloss, rowhit, pull_sq, gpull, Z, total;
};

HeadLayout head_layout(const sr_head_args* a) {
    HeadLayout L;
    const int64_t nt = (int64_t)a->n_support + a->n_memory;
    int64_t off = 0;
    L.ctrl = off;    off += align_up(sizeof(HeadCtrl), 256);
    L.rowloss = off; off += align_up(nt * 4, 256);
    L.rowhit = off;  off += align_up(nt * 4, 256);
    L.pull_sq = off; off += align_up((int64_t)std::max(a->n_new, 1) * 8, 256);
    L.gpull = off;   off += align_up((int64_t)std::max(a->n_new, 1) * a->dim * 4, 256);
    L.Z = off;       off += align_up(nt * a->n_classes * 4, 256);
    L.total = off;
    return L;
}

// ------------------------------------------------------------------------------------------------
// Scoring
// ------------------------------------------------------------------------------------------------
template <int BT>
__global__ void __launch_bounds__(kHeadThreads) logits_kernel(const float* X, const float* W, float* Z, int n, int C,
                                                              int d) {
    __shared__ TileSmem<BT, BT> tile;
    constexpr int TM = BT / 16;
    const int tiles_n = (C + BT - 1) / BT;
    const int m0 = (blockIdx.x / tiles_n) * BT, n0 = (blockIdx.x % tiles_n) * BT;
    const RowMap rm{n, 0, 0};
    float acc[TM][TM];
    gemm_nt_tile<BT, BT>(X, rm, n, W, C, d, m0, n0, tile, acc);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TM; ++j) {
            const int r = m0 + ty * TM + i, c = n0 + tx * TM + j;
            if (r < n && c < C) Z[(int64_t)r * C + c] = acc[i][j];
        }
}

// `src` / `pitch`: logits computed elsewhere with a padded row pitch (the tensor-core path); they are copied to a.logits.
__global__ void __launch_bounds__(256) score_rows_kernel(const sr_eval_args a, const float* src, int pitch) {
    __shared__ int s_cnt[2];
    __shared__ float s_loss;
    if (threadIdx.x == 0) { s_cnt[0] = 0; s_cnt[1] = 0; s_loss = 0.f; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r = blockIdx.x * 8 + warp;
    if (r < a.n) {
        const int C = a.n_classes;
        const float* z = src != nullptr ? src + (int64_t)r * pitch : a.logits + (int64_t)r * C;
        const int y = (int)a.labels[r];
        float mx = -INFINITY;
        int arg = 0x7fffffff;
        for (int c = lane; c < C; c += 32) {
            const float v = z[c];
            if (src != nullptr) a.logits[(int64_t)r * C + c] = v;
            if (v > mx) { mx = v; arg = c; }  // strided ascending: first maximum per lane
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const float om = __shfl_xor_sync(0xffffffffu, mx, o);
            const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
            if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
        }
        float se = 0.f;
        for (int c = lane; c < C; c += 32) se += expf(z[c] - mx);
        se = warp_sum(se);
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
            const int rank = greater + tie_before;
            a.pred[r] = arg;
            if (rank == 0) atomicAdd(&s_cnt[0], 1);
            if (rank < 5) atomicAdd(&s_cnt[1], 1);
            atomicAdd(&s_loss, (mx + logf(se)) - zy);
            if (a.confusion != nullptr && y < a.conf_dim && arg < a.conf_dim)
                atomicAdd(reinterpret_cast<unsigned long long*>(a.confusion + (int64_t)y * a.conf_dim + arg), 1ull);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_cnt[0]) atomicAdd(&a.counts[0], s_cnt[0]);
        if (s_cnt[1]) atomicAdd(&a.counts[1], s_cnt[1]);
        atomicAdd(a.loss_sum, s_loss);
    }
}

// ------------------------------------------------------------------------------------------------
// Subspace factor
// ------------------------------------------------------------------------------------------------
__global__ void gram_kernel(const float* __restrict__ B, int n, int d, double* __restrict__ G) {
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int64_t pairs = (int64_t)n * (n + 1) / 2;
    for (int64_t pidx = gw; pidx < pairs; pidx += warps) {
        // unrank (i >= j) from pidx = i(i+1)/2 + j
        int i = (int)((sqrt(8.0 * (double)pidx + 1.0) - 1.0) * 0.5);
        while ((int64_t)(i + 1) * (i + 2) / 2 <= pidx) ++i;
        while ((int64_t)i * (i + 1) / 2 > pidx) --i;
        const int j = (int)(pidx - (int64_t)i * (i + 1) / 2);
        double s = 0.0;
        for (int k = lane; k < d; k += 32) s += (double)B[(int64_t)i * d + k] * (double)B[(int64_t)j * d + k];
        s = warp_sum(s);
        if (lane == 0) { G[(int64_t)i * n + j] = s; G[(int64_t)j * n + i] = s; }
    }
}

// Right-looking Cholesky of the n x n fp64 matrix (lower triangle), one CTA.  `Gs` is shared memory when it fits.
__global__ void __launch_bounds__(1024) chol_kernel(double* __restrict__ Gg, int n, int use_smem, int* info) {
    extern __shared__ __align__(16) uint8_t dyn[];
    double* G = use_smem ? reinterpret_cast<double*>(dyn) : Gg;
    const int tid = threadIdx.x, nt = blockDim.x;
    if (use_smem) {
        for (int i = tid; i < n * n; i += nt) G[i] = Gg[i];
        __syncthreads();
    }
    __shared__ int s_bad;
    if (tid == 0) s_bad = 0;
    __syncthreads();
    for (int j = 0; j < n; ++j) {
        const double piv = G[(int64_t)j * n + j];
        if (!(piv > 0.0)) {
            if (tid == 0) s_bad = j + 1;
            __syncthreads();
            break;
        }
        const double l = sqrt(piv);
        __syncthreads();
        for (int i = j + tid; i < n; i += nt) G[(int64_t)i * n + j] = (i == j) ? l : G[(int64_t)i * n + j] / l;
        __syncthreads();
        const int rem = n - j - 1;  // trailing update on the lower triangle
        for (int idx = tid; idx < rem * rem; idx += nt) {
            const int i = j + 1 + idx / rem, k = j + 1 + idx % rem;
            if (k <= i) G[(int64_t)i * n + k] -= G[(int64_t)i * n + j] * G[(int64_t)k * n + j];
        }
        __syncthreads();
    }
    if (use_smem) {
        for (int i = tid; i < n * n; i += nt) Gg[i] = G[i];
    }
    if (tid == 0) info[2] = s_bad;
}

// Qt = L^-1 B: forward substitution, one thread per column of B.
__global__ void trisolve_kernel(const double* __restrict__ L, const float* __restrict__ B, int n, int d,
                                double* __restrict__ Qd, float* __restrict__ Qt) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= d) return;
    for (int i = 0; i < n; ++i) {
        double s = (double)B[(int64_t)i * d + k];
        for (int j = 0; j < i; ++j) s -= L[(int64_t)i * n + j] * Qd[(int64_t)j * d + k];
        s /= L[(int64_t)i * n + i];
        Qd[(int64_t)i * d + k] = s;
        Qt[(int64_t)i * d + k] = (float)s;
    }
}

__global__ void set_info_kernel(int* info, int a, int b, int c) { info[0] = a; info[1] = b; info[2] = c; info[3] = 0; }

int num_sms() { return current_device_sms(); }

template <int BT>
int32_t launch_head(const HeadParams& p, int grid, size_t dyn, cudaStream_t stream) {
    void* args[] = {const_cast<HeadParams*>(&p)};
    SR_CUDA_OK(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(head_kernel<BT>), dim3(grid), dim3(kHeadThreads), args,
                                           dyn, stream));
    return SR_OK;
}

}  // namespace

extern "C" int64_t sr_head_workspace_bytes(const sr_head_args* a) {
    if (!a) return 0;
    int64_t n = head_layout(a).total;
    if (a->dim >= 64 && a->dim % 64 == 0) n = std::max<int64_t>(n, srb::head_small_workspace_bytes(a));
    if (a->dim >= 64 && a->dim % 64 == 0 && a->n_classes <= 128) n = std::max<int64_t>(n, srb::head_cluster_workspace_bytes(a));
    if (srb::head_tc_applicable(a)) n = std::max<int64_t>(n, srb::head_tc_workspace_bytes(a));
    return n;
}

// Host-only view of the dispatch: which kernel sr_head_run picks for this problem and, for the paper-size kernel, its launch
// shape under a->cta_budget.  out[0] = kernel (0 = fp32 SIMT head_kernel, 1 = head_small, 2 = head_cluster, 3 = tensor-core
// head_tc), out[1] = rows per row CTA, out[2] = feature columns per column CTA, out[3] = CTAs of the cooperative launch.
extern "C" int32_t sr_head_plan(const sr_head_args* a, int32_t* out4) {
    if (!a || !out4) return fail(SR_E_ARG, "sr_head_plan: null pointer");
    out4[0] = out4[1] = out4[2] = out4[3] = 0;
    if (a->dim < 4 || a->dim % 4 || a->n_support < 1 || a->n_classes < 1) return fail(SR_E_ARG, "sr_head_plan: empty problem");
    if (srb::head_small_applicable(a)) {
        out4[0] = srb::head_cluster_applicable(a) ? 2 : 1;
        int r = 0, c = 0, g = 0;
        srb::head_small_shape(a, &r, &c, &g);
        out4[1] = r; out4[2] = c; out4[3] = g;
    } else if (srb::head_tc_applicable(a) && !getenv("SRB_HEAD_SIMT")) {
        out4[0] = 3;
    }
    return SR_OK;
}

extern "C" int32_t sr_head_run(const sr_head_args* a, void* stream_v) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    if (!a) return fail(SR_E_ARG, "sr_head_run: null args");
    if (!a->feat || !a->weight || !a->opt_state || !a->labels_support || !a->loss_trace || !a->status || !a->workspace)
        return fail(SR_E_ARG, "sr_head_run: null pointer");
    if (a->dim < 4 || a->dim % 4) return fail(SR_E_ARG, "sr_head_run: dim must be a multiple of 4 (16-byte rows)");
    if (a->n_support < 1 || a->n_classes < 1 || a->max_epochs < 1) return fail(SR_E_ARG, "sr_head_run: empty problem");
    if (a->n_memory > 0 && !a->labels_memory) return fail(SR_E_ARG, "sr_head_run: memory rows without labels");
    if (a->n_new < 0 || a->n_new > a->n_classes) return fail(SR_E_ARG, "sr_head_run: bad n_new");
    if (a->pull_mode != SR_PULL_NONE && (!a->pull || a->n_new < 1)) return fail(SR_E_ARG, "sr_head_run: pull mode without data");
    if (a->pull_mode == SR_PULL_PROJECT && a->q_rows < 1) return fail(SR_E_ARG, "sr_head_run: q_rows < 1");
    if (a->base_weight && a->n_base > a->n_classes) return fail(SR_E_ARG, "sr_head_run: n_base > n_classes");
    if (a->reserve_weight && a->n_base + a->n_prev_novel > a->n_classes)
        return fail(SR_E_ARG, "sr_head_run: reserve rows exceed n_classes");
    if (a->optimizer != SR_OPT_SGD && a->optimizer != SR_OPT_ADAM) return fail(SR_E_ARG, "sr_head_run: bad optimizer");
    if (reinterpret_cast<uintptr_t>(a->workspace) & 255) return fail(SR_E_ARG, "sr_head_run: workspace must be 256-byte aligned");
    if (a->resume_status && a->resume_status == a->status)
        return fail(SR_E_ARG, "sr_head_run: resume_status must be a different block than status");
    // Paper-sized problems: everything constant over the session stays in shared memory (head_small.cu).
    // ... or, opt-in (SRB_HEAD_CLUSTER=1), the variant that splits both reductions over thread-block clusters (head_cluster.cu)
    if (srb::head_small_applicable(a) && srb::head_cluster_applicable(a)) return srb::head_cluster_run(a, stream);
    if (srb::head_small_applicable(a)) return srb::head_small_run(a, stream);
    // Large problems: the two GEMMs on tcgen05 (head_tc.cu).  SRB_HEAD_SIMT=1 keeps the fp32 SIMT kernel below (A/B checks).
    if (srb::head_tc_applicable(a) && !getenv("SRB_HEAD_SIMT")) return srb::head_tc_run(a, stream);
    const HeadLayout L = head_layout(a);
    if (a->workspace_bytes < L.total) return fail(SR_E_SMALLWS, "sr_head_run: workspace %lld < %lld",
                                                  (long long)a->workspace_bytes, (long long)L.total);

    HeadParams p;
    p.a = *a;
    p.n_total = a->n_support + a->n_memory;
    uint8_t* ws = static_cast<uint8_t*>(a->workspace);
    p.ctrl = reinterpret_cast<HeadCtrl*>(ws + L.ctrl);
    p.rowloss = reinterpret_cast<float*>(ws + L.rowloss);
    p.rowhit = reinterpret_cast<int*>(ws + L.rowhit);
    p.pull_sq = reinterpret_cast<double*>(ws + L.pull_sq);
    p.gpull = reinterpret_cast<float*>(ws + L.gpull);
    p.Z = reinterpret_cast<float*>(ws + L.Z);
    SR_CUDA_OK(cudaMemsetAsync(p.ctrl, 0, sizeof(HeadCtrl), stream));

    const size_t dyn = (size_t)(2 * a->dim + std::max(a->q_rows, 1)) * sizeof(float);
    if (dyn > 160 * 1024) return fail(SR_E_ARG, "sr_head_run: dim/q_rows too large for the projection task");
    const int sms = num_sms();
    const int64_t work64 = ((int64_t)(p.n_total + 63) / 64) * ((a->n_classes + 63) / 64);
    const bool big = work64 >= 2 * (int64_t)sms;
    const int bt = big ? 64 : 32;
    const int64_t tasksA = ((int64_t)(p.n_total + bt - 1) / bt) * ((a->n_classes + bt - 1) / bt) + a->n_new + 2;
    const int64_t tasksC = ((int64_t)(a->n_classes + bt - 1) / bt) * ((a->dim + bt - 1) / bt);
    int grid = (int)std::min<int64_t>(sms, std::max<int64_t>(std::max(tasksA, tasksC), 1));
    {
        static PerDeviceOnce once;
        once_per_device(once, [] {
            cudaFuncSetAttribute(head_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
            cudaFuncSetAttribute(head_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        });
    }
    return big ? launch_head<64>(p, grid, dyn, stream) : launch_head<32>(p, grid, dyn, stream);
}

extern "C" int32_t sr_eval_logits(const sr_eval_args* a, void* stream_v) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    if (!a || !a->feat || !a->weight || !a->labels || !a->logits || !a->pred || !a->counts || !a->loss_sum)
        return fail(SR_E_ARG, "sr_eval_logits: null pointer");
    if (a->n < 1 || a->n_classes < 1 || a->dim < 4 || a->dim % 4) return fail(SR_E_ARG, "sr_eval_logits: bad sizes");
    if (a->workspace != nullptr && !getenv("SRB_HEAD_SIMT")) {
        const int64_t need = srb::eval_tc_workspace_bytes(a->n, a->dim, a->n_classes);
        if (need > 0 && a->workspace_bytes >= need) {
            const float* z = nullptr;
            int pitch = 0;
            const int32_t rc = srb::eval_tc_logits(a, stream, &z, &pitch);
            if (rc != SR_OK) return rc;
            score_rows_kernel<<<(a->n + 7) / 8, 256, 0, stream>>>(*a, z, pitch);
            SR_CUDA_OK(cudaGetLastError());
            return SR_OK;
        }
    }
    const int64_t t64 = ((int64_t)(a->n + 63) / 64) * ((a->n_classes + 63) / 64);
    if (t64 >= 2 * (int64_t)num_sms()) {
        logits_kernel<64><<<(unsigned)t64, kHeadThreads, 0, stream>>>(a->feat, a->weight, a->logits, a->n, a->n_classes, a->dim);
    } else {
        const int64_t t32 = ((int64_t)(a->n + 31) / 32) * ((a->n_classes + 31) / 32);
        logits_kernel<32><<<(unsigned)t32, kHeadThreads, 0, stream>>>(a->feat, a->weight, a->logits, a->n, a->n_classes, a->dim);
    }
    SR_CUDA_OK(cudaGetLastError());
    score_rows_kernel<<<(a->n + 7) / 8, 256, 0, stream>>>(*a, nullptr, 0);
    SR_CUDA_OK(cudaGetLastError());
    return SR_OK;
}

extern "C" int64_t sr_eval_workspace_bytes(int32_t n, int32_t dim, int32_t n_classes) {
    return srb::eval_tc_workspace_bytes(n, dim, n_classes);
}

extern "C" int32_t sr_score_logits(const sr_eval_args* a, void* stream_v) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    if (!a || !a->labels || !a->logits || !a->pred || !a->counts || !a->loss_sum)
        return fail(SR_E_ARG, "sr_score_logits: null pointer");
    if (a->n < 1 || a->n_classes < 1) return fail(SR_E_ARG, "sr_score_logits: bad sizes");
    score_rows_kernel<<<(a->n + 7) / 8, 256, 0, stream>>>(*a, nullptr, 0);
    SR_CUDA_OK(cudaGetLastError());
    return SR_OK;
}

extern "C" int64_t sr_subspace_factor_workspace_bytes(int32_t n_base, int32_t dim) {
    if (n_base < 1 || dim < 1) return 0;
    const int64_t n = std::min(n_base, dim);
    return align_up(n * n * 8, 256) + align_up(n * (int64_t)dim * 8, 256);
}

extern "C" int32_t sr_subspace_factor(const float* base, int32_t n_base, int32_t dim, float* qt, int32_t* info,
                                      void* workspace, int64_t workspace_bytes, void* stream_v) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    if (!base || !qt || !info || n_base < 1 || dim < 1) return fail(SR_E_ARG, "sr_subspace_factor: bad arguments");
    if (n_base >= dim) {  // generic position: the rows span R^dim and the projector is the identity
        set_info_kernel<<<1, 1, 0, stream>>>(info, dim, 1, 0);
        SR_CUDA_OK(cudaGetLastError());
        return SR_OK;
    }
    if (!workspace || workspace_bytes < sr_subspace_factor_workspace_bytes(n_base, dim))
        return fail(SR_E_SMALLWS, "sr_subspace_factor: workspace too small");
    const int n = n_base;
    double* G = static_cast<double*>(workspace);
    double* Qd = reinterpret_cast<double*>(static_cast<uint8_t*>(workspace) + align_up((int64_t)n * n * 8, 256));
    set_info_kernel<<<1, 1, 0, stream>>>(info, n, 0, 0);
    const int64_t pairs = (int64_t)n * (n + 1) / 2;
    const int gblocks = (int)std::min<int64_t>((pairs + 7) / 8, 4 * num_sms());
    gram_kernel<<<gblocks, 256, 0, stream>>>(base, n, dim, G);
    SR_CUDA_OK(cudaGetLastError());
    const size_t gbytes = (size_t)n * n * 8;
    const int use_smem = gbytes <= 200 * 1024 ? 1 : 0;
    if (use_smem) {
        static PerDeviceOnce once;
        once_per_device(once, [] { cudaFuncSetAttribute(chol_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); });
    }
    chol_kernel<<<1, 1024, use_smem ? gbytes : 0, stream>>>(G, n, use_smem, info);
    SR_CUDA_OK(cudaGetLastError());
    trisolve_kernel<<<(dim + 127) / 128, 128, 0, stream>>>(G, base, n, dim, Qd, qt);
    SR_CUDA_OK(cudaGetLastError());
    return SR_OK;
}
