/*
 * srb200 — C ABI of the B200-native (sm_100a) kernels behind subspace-reg's incremental-session path.
 *
 * The reference (feyzaakyurek/subspace-reg) has no FFI layer: its "operator interface" is the set of
 * PyTorch library calls made from models/resnet_language.py and eval/language_eval.py.  Each entry
 * point below replaces one group of those calls; the cited file:line is the reference call site.
 *
 * Conventions (all entry points):
 *   - plain pointers + sizes, no torch types; every pointer is a DEVICE pointer unless stated;
 *   - the caller owns all memory (including workspaces); the library never allocates device memory,
 *     never frees and never keeps a pointer past the call;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*), no host synchronisation
 *     unless stated;
 *   - return 0 on success, a negative SR_E_* code otherwise; sr_last_error() returns a thread-local
 *     human readable message for the last failure on the calling thread;
 *   - sm_100a only: any other device is refused with SR_E_DEVICE (there is no fallback path).
 */
#ifndef SRB200_H_
#define SRB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SR_OK 0
#define SR_E_ARG (-1)     /* bad argument / unsupported shape */
#define SR_E_DEVICE (-2)  /* not an sm_100 device, or driver entry point missing */
#define SR_E_CUDA (-3)    /* CUDA runtime error (message in sr_last_error) */
#define SR_E_SMALLWS (-4) /* workspace too small */

const char* sr_last_error(void);
/* version = major*10000 + minor*100 + patch */
int32_t sr_version(void);
/* 0 if device `dev` can run this library (compute capability 10.x), SR_E_DEVICE otherwise. */
int32_t sr_check_device(int32_t dev);

/* ------------------------------------------------------------------------------------------------
 * Backbone: data-layout kernels
 * ---------------------------------------------------------------------------------------------- */

/* NCHW fp32 images -> NHWC bf16 with the channel dimension zero-padded to `cpad` (multiple of 16).
 * Replaces the implicit layout PyTorch's conv2d consumes (resnet_language.py:170-171). */
int32_t sr_pack_input(const float* x_nchw, void* y_nhwc_bf16, void* y_lo, int32_t batch, int32_t channels,
                      int32_t height, int32_t width, int32_t cpad, void* stream);
/* y_lo / w_lo / x_lo / out_lo below: the low bf16 plane of an error-compensated operand pair (see sr_conv_panel), same
 * shape as the main output; NULL = plain bf16. */

/* uint8 HWC images (the reference's image store, dataset/mini_imagenet.py:278-350) -> NHWC bf16 with the channel
 * dimension zero-padded to `cpad`, applying ToTensor (x / 255) and Normalize ((x - mean) / std, transform_cfg.py:8-10,
 * 42-45) on the way: the same fp32 operations in the same order, so the result is bit-identical to the reference's
 * preprocessing followed by sr_pack_input.  mean / std: HOST float[channels].
 * Optional augmentation of the support transform (transform_cfg.py:32-40), with parameters drawn by the caller:
 * crop_ij: DEVICE int32 [batch,2] = top-left corner of the height x width crop inside the image zero-padded by `pad`
 * on every side (RandomCrop(size, padding=pad)), or NULL; flip: DEVICE uint8 [batch] (RandomHorizontalFlip, applied
 * after the crop), or NULL. */
int32_t sr_pack_input_u8(const uint8_t* x_nhwc_u8, void* y_nhwc_bf16, void* y_lo, int32_t batch, int32_t channels,
                         int32_t height, int32_t width, const float* mean_host, const float* std_host, int32_t cpad,
                         const int32_t* crop_ij, const uint8_t* flip, int32_t pad, void* stream);

/* BatchNorm eval-mode fold (resnet_language.py:250-255,148; nn.BatchNorm2d eval semantics):
 *   scale[c] = gamma[c] / sqrt(running_var[c] + eps),  shift[c] = beta[c] - running_mean[c] * scale[c]. */
int32_t sr_bn_fold(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                   float eps, float* scale, float* shift, int32_t channels, void* stream);

/* Conv weight repack: OIHW fp32 [cout,cin,kh,kw] -> bf16 [cout][kh*kw][cin_pad] (zero padded), each output
 * channel optionally multiplied by scale[cout] (NULL = 1) before rounding (folded BN scale). */
int32_t sr_pack_weight(const float* w_oihw, const float* scale, void* w_packed_bf16, void* w_lo, int32_t cout,
                       int32_t cin, int32_t kh, int32_t kw, int32_t cin_pad, void* stream);

/* AdaptiveAvgPool2d(1) on a bf16 NHWC map -> fp32 [batch, channels] (resnet_language.py:179-181 when the last block is
 * pooled, i.e. resnet12; resnet18's last block averages inside sr_conv / sr_bn_apply). */
int32_t sr_global_avg(const void* x_nhwc_bf16, const void* x_lo, float* y, int32_t batch, int32_t height, int32_t width,
                      int32_t channels, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Backbone: implicit-GEMM convolution on tcgen05 / TMEM fed by TMA
 * Replaces nn.Conv2d + nn.BatchNorm2d(eval) + LeakyReLU + residual add + MaxPool2d / AdaptiveAvgPool2d
 * of BasicBlock.forward (resnet_language.py:268-301) and ResNet.forward (:170-181).
 * ---------------------------------------------------------------------------------------------- */
#define SR_EPI_ACT 0       /* y = lrelu(acc + shift [+ residual])            -> bf16 NHWC [B,H,W,cout]       */
#define SR_EPI_ACT_POOL2 1 /* ... then MaxPool2d(2) (floor)                  -> bf16 NHWC [B,H/2,W/2,cout]   */
#define SR_EPI_ACT_AVG 2   /* ... then mean over H x W (AdaptiveAvgPool2d(1)) -> fp32 [B,cout]               */
#define SR_EPI_RAW_STATS 3 /* raw accumulator -> fp32 NHWC [B,H,W,cout]; per-channel sum / sum of squares    */
                           /*   are ADDED to stats[0..cout) / stats[cout..2cout) (fp64, caller zeroes them)  */

typedef struct sr_conv_panel {
    const void* act;  /* bf16 NHWC [batch,height,width,cin_pad]                                   */
    const void* wgt;  /* bf16 [cout][taps][cin_pad] from sr_pack_weight                           */
    int32_t cin_pad;  /* channel pitch, multiple of 16                                            */
    int32_t taps;     /* 9 = 3x3 stride 1 pad 1,  1 = 1x1 stride 1                                */
    /* Error-compensated ("bf16x3") mode, selected by act_lo != NULL on every panel: each operand is a PAIR of bf16
     * tensors, hi = rn(x) and lo = rn(x - hi) (~17 mantissa bits together); the kernel accumulates hi*hi + hi*lo +
     * lo*hi in fp32.  This is the parity tier for north_star's "bit-exact class predictions" against the reference's
     * fp32 nn.Conv2d (resnet_language.py:249-254); 3x the tensor work of the plain bf16 mode. */
    const void* act_lo; /* same shape as act, or NULL                                             */
    const void* wgt_lo; /* same shape as wgt, or NULL                                             */
} sr_conv_panel;

typedef struct sr_conv_args {
    int32_t batch, height, width, cout;
    int32_t n_panels;        /* 1, or 2 when the 1x1 downsample conv accumulates into the same tile     */
    sr_conv_panel panel[2];
    const float* shift;      /* [cout] fp32 added to the accumulator (folded BN shift); NULL = 0        */
    const void* residual;    /* bf16 NHWC [batch,height,width,cout] added before the activation; or NULL */
    const void* residual_lo; /* error-compensated mode: low plane of the residual                        */
    float slope;             /* LeakyReLU negative slope (0.1 in the reference)                          */
    int32_t epilogue;        /* SR_EPI_*                                                                 */
    void* out;               /* see SR_EPI_*                                                             */
    void* out_lo;            /* error-compensated mode, bf16 outputs: low plane (same shape as out)      */
    double* stats;           /* SR_EPI_RAW_STATS: per-channel sums; NULL = raw fp32 output only           */
    /* The two fields below serve the tensor-core classifier head (csrc/head_tc.cu), which runs its two GEMMs
     * (logits = X W^T, dW = dZ^T X; resnet_language.py:187 and its autograd) on this kernel as 1x1 convolutions. */
    int32_t weights_per_image;      /* image n uses weight rows [n*cout, (n+1)*cout): a batch of independent GEMMs  */
                                    /*   (split-K partial products); needs height*width >= 1024                     */
    const int32_t* skip_if_nonzero; /* optional DEVICE flag read at kernel start: non-zero = the launch does nothing */
    int32_t max_cout_per_cta;       /* 0 = 256; smaller N tiles (>= 32) trade MMA width for more tiles and a second   */
                                    /*   TMEM accumulator stage (epilogue overlaps the next tile's main loop)          */
} sr_conv_args;

int32_t sr_conv(const sr_conv_args* a, void* stream);

/* Host-only (no device work, pointers in `a` are not dereferenced): the launch plan sr_conv would use for these shapes -
 * out16 = { TW, TH, TN, row_stacked, tiles_w, tiles_h, n_cta, n_splits, tmem_stages, tap_reuse(panel 0),
 *           row_bytes(panel 0), ring_stages, stage_bytes, dynamic_smem_bytes, channel_blocks(panel 0), last_ksteps(panel 0) }.
 * There is no reference counterpart (cuDNN chooses its own algorithms behind nn.Conv2d, resnet_language.py:268-301);
 * it exists so that tests can pin the tile / pipeline choice per layer shape. */
int32_t sr_conv_plan(const sr_conv_args* a, int32_t* out16);

/* ------------------------------------------------------------------------------------------------
 * Backbone: train-mode BatchNorm (epoch 1 of every session, language_eval.py:211,252,257)
 * ---------------------------------------------------------------------------------------------- */

/* From fp64 sums over `count` = B*H*W values per channel: batch mean, invstd = 1/sqrt(biased var + eps);
 * running_mean/var updated in place with `momentum` (unbiased variance), nn.BatchNorm2d train semantics. */
int32_t sr_bn_finalize(const double* stats, int64_t count, float eps, float momentum, float* running_mean,
                       float* running_var, float* mean, float* invstd, int32_t channels, void* stream);

typedef struct sr_bn_apply_args {
    int32_t batch, height, width, channels;
    const float* raw;       /* fp32 NHWC conv output (SR_EPI_RAW_STATS)                                  */
    const float *mean, *invstd, *gamma, *beta;
    const float* res_raw;   /* optional: fp32 NHWC raw output of the 1x1 downsample conv ...             */
    const float *res_mean, *res_invstd, *res_gamma, *res_beta; /* ... with its own batch-norm            */
    const void* res_act;    /* optional: bf16 NHWC identity residual                                     */
    const void* res_act_lo; /* optional: its low plane (error-compensated mode)                          */
    int32_t lrelu;          /* apply LeakyReLU(slope) after the (optional) residual add                  */
    float slope;
    int32_t pool;           /* 0 none, 2 = MaxPool2d(2), -1 = global average (-> fp32 [B,C] in `out`)    */
    const uint8_t* keep;    /* optional NCHW uint8 [B,C,Ho,Wo] keep-mask applied AFTER pooling           */
    float keep_scale;       /* multiplier for kept elements (1/(1-p) for dropout, numel/kept for DropBlock) */
    void* out;              /* bf16 NHWC [B,Ho,Wo,C], or fp32 [B,C] when pool == -1                      */
    void* out_lo;           /* optional: low plane of the bf16 output (error-compensated mode)           */
    const float* keep_scale_dev; /* optional DEVICE scalar used instead of keep_scale (sr_dropblock_keep's output) */
} sr_bn_apply_args;

int32_t sr_bn_apply(const sr_bn_apply_args* a, void* stream);

/* One BasicBlock of the train-mode pass up to its last BatchNorm (resnet_language.py:273-286), sequenced by the library:
 *   conv1 -> batch statistics -> BN + LeakyReLU -> conv2 -> ... -> conv3 [, 1x1 downsample conv of the block input] with
 *   every BatchNorm's running statistics updated in place - 9 (11) launches behind one call.
 * The caller finishes the block with ONE sr_bn_apply (bn3 [+ downsample BN | identity residual], LeakyReLU, pooling,
 * dropout / DropBlock keep-mask) on raw[2] / raw[3] and the mean / invstd rows this call leaves in `mean_invstd`; that
 * split lets the host draw the block's keep-mask while these launches run.  All buffers are caller-owned. */
typedef struct sr_train_block_args {
    int32_t batch, height, width, cin_pad, cout;
    int32_t downsample;          /* 1: the block has the 1x1 conv + BN residual branch                              */
    const void *x, *x_lo;        /* block input NHWC bf16 [B,H,W,cin_pad] (+ low plane in the error-compensated tier) */
    const void* w[4];            /* packed raw weights (sr_pack_weight, no BN scale): conv1, conv2, conv3, downsample */
    const void* w_lo[4];         /* their low planes, or NULL                                                        */
    const float* gamma[2];       /* BN affine of bn1, bn2 (bn3 / downsample BN are applied by the caller)            */
    const float* beta[2];
    float* running_mean[4];      /* running statistics of bn1, bn2, bn3, downsample BN: EMA-updated in place         */
    float* running_var[4];
    float eps, momentum, slope;
    double* stats;               /* [(3|4)][2*cout] fp64, zeroed by the caller                                       */
    float* mean_invstd;          /* out [(3|4)][2][cout]: batch mean, then 1/sqrt(biased var + eps), per conv         */
    float* raw[4];               /* fp32 NHWC [B,H,W,cout] conv outputs; raw[0] / raw[1] are scratch and may alias    */
    void *h1, *h1_lo, *h2, *h2_lo; /* scratch activations bf16 NHWC [B,H,W,cout] (low planes in the x3 tier)         */
} sr_train_block_args;

int32_t sr_train_block(const sr_train_block_args* a, void* stream);

/* The whole eval-mode backbone (ResNet.forward up to `feat`, resnet_language.py:170-181, on folded-BN weights) behind ONE
 * call: 3 convolutions per BasicBlock (the 1x1 downsample conv rides as a second K panel of conv3), LeakyReLU, residual,
 * MaxPool2d(2), and the final AdaptiveAvgPool2d(1) fused into the last epilogue (a pooled last block - resnet12 - gets
 * sr_global_avg).  18 launches for resnet18.  `blocks` is a HOST array. */
typedef struct sr_eval_block {
    int32_t cout;
    int32_t pool;            /* MaxPool2d stride of the block: 1 or 2                                      */
    int32_t downsample;      /* 1: wd is the 1x1 downsample conv of the residual branch (BN folded in)      */
    int32_t reserved;
    const void *w1, *w2, *w3, *wd;             /* packed weights with the BN scale folded in (sr_pack_weight) */
    const void *w1_lo, *w2_lo, *w3_lo, *wd_lo; /* low planes (error-compensated tier) or NULL                 */
    const float *s1, *s2, *s3;                 /* folded shifts; s3 already holds bn3 + downsample-BN shifts   */
} sr_eval_block;

typedef struct sr_backbone_eval_args {
    int32_t n_blocks;
    const sr_eval_block* blocks;
    int32_t batch, height, width, cin_pad;     /* of the packed input                                         */
    const void *x, *x_lo;                      /* NHWC bf16 [batch,height,width,cin_pad] (+ low plane)         */
    float slope;
    void* workspace;                           /* sr_backbone_eval_workspace_bytes(...) bytes, 256-byte aligned */
    int64_t workspace_bytes;
    float* features;                           /* out: fp32 [batch, cout of the last block]                    */
} sr_backbone_eval_args;

int64_t sr_backbone_eval_workspace_bytes(const sr_backbone_eval_args* a);
int32_t sr_backbone_eval(const sr_backbone_eval_args* a, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Subspace regulariser: orthonormal factor of span(base weights)
 * Replaces torch.qr(base_weight^T, some=True) in LangPuller.get_projected_weight
 * (resnet_language.py:92-97).  Factors the Gram matrix of the smaller side.
 * ---------------------------------------------------------------------------------------------- */
/* base [n_base, dim] row-major fp32.  Writes an orthonormal row basis qt [rank_rows, dim] of span(rows of base)
 * where rank_rows = min(n_base, dim); info[0..3]: info[0] = rank_rows, info[1] = 1 when the span is all of R^dim
 * (projection == identity, n_base >= dim), info[2] = index of the first non-positive pivot + 1, else 0.
 * workspace: sr_subspace_factor_workspace_bytes(n_base, dim) bytes. */
int64_t sr_subspace_factor_workspace_bytes(int32_t n_base, int32_t dim);
int32_t sr_subspace_factor(const float* base, int32_t n_base, int32_t dim, float* qt, int32_t* info, void* workspace,
                           int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused classifier-head fine-tuning (language_eval.py:242-318)
 * One "epoch" = CE(support) + CE(memory) + lmbd_b*||W[:nb]-W0|| + lmbd_n*||W[nb:nb+np]-Wres|| +
 * gamma*||pull - W[new]||^2, its gradient w.r.t. W, and the SGD-momentum / Adam update of W
 * (eval/util.py:92-102), all on cached 640-d features.
 * ---------------------------------------------------------------------------------------------- */
#define SR_PULL_NONE 0    /* no label_pull term                                                        */
#define SR_PULL_FIXED 1   /* pullers constant (semantic / linear-mapping modes, resnet_language.py:75-87) */
#define SR_PULL_PROJECT 2 /* pullers = projection of W[new] on span(base), differentiable (:92-97)      */
#define SR_OPT_SGD 0
#define SR_OPT_ADAM 1
#define SR_TRACE_COLS 8

typedef struct sr_head_args {
    /* data */
    const float* feat;        /* feature cache, row-major [*, dim]                                       */
    int32_t dim;
    int32_t n_support;        /* rows [support_row0, support_row0+n_support) are the support batch        */
    int32_t support_row0;
    int32_t n_memory;         /* rows [memory_row0, ...) are the replay batch (0 = none)                  */
    int32_t memory_row0;
    const int64_t* labels_support; /* [n_support] class ids                                              */
    const int64_t* labels_memory;  /* [n_memory]                                                         */
    /* parameters / state (updated in place) */
    float* weight;            /* [n_classes, dim] classifier.weight                                       */
    int32_t n_classes;
    float* opt_state;         /* SGD: momentum buffer [n_classes*dim]; Adam: exp_avg then exp_avg_sq      */
    /* regularisers */
    const float* base_weight; /* W0 [n_base, dim] or NULL (no lmbd_reg_transform_w)                       */
    int32_t n_base;
    const float* reserve_weight; /* [n_prev_novel, dim] or NULL                                          */
    int32_t n_prev_novel;
    int32_t n_new;            /* rows [n_classes-n_new, n_classes) are this session's classes             */
    int32_t pull_mode;        /* SR_PULL_*                                                                */
    const float* pull;        /* FIXED: pullers [n_new, dim]; PROJECT: qt [q_rows, dim] from sr_subspace_factor */
    int32_t q_rows;
    float lmbd_base, lmbd_novel, gamma;
    /* optimiser */
    int32_t optimizer;        /* SR_OPT_*                                                                 */
    float lr, momentum, weight_decay, beta1, beta2, adam_eps;
    int32_t step0;            /* optimiser steps already taken on this state (0 for a fresh session)      */
    /* stopping rule (language_eval.py:298-318) */
    int32_t max_epochs;       /* run at most this many epochs in this call                                */
    int32_t epoch0;           /* epochs already run in this session (for min/max_novel_epochs)            */
    int32_t stable;           /* opt.stable: use the |dloss| < eps rule                                   */
    int32_t stable_epochs;
    int32_t stable_count0;    /* carry-over of the stable-epoch counter                                   */
    int32_t min_novel_epochs, max_novel_epochs;
    double convergence_epsilon;
    double target_train_loss;
    float prev_loss;          /* train_loss carried in (15 at session start)                              */
    /* outputs */
    float* loss_trace;        /* [max_epochs][SR_TRACE_COLS]: total, ce_support, ce_memory, reg_base,     */
                              /*   reg_novel, pull, support top-1 hits, support top-5 hits (pre-update W) */
    int32_t* status;          /* [8]: epochs run in this call, stopped flag, stable counter, error flag,   */
                              /*   epochs run in total (epoch0 + [0]), last loss (float bits), 0, 0        */
    const int32_t* resume_status; /* optional: the `status` block of the PREVIOUS sr_head_run of this session, */
                              /*   still on the device.  epoch0 / step0 / stable_count0 / prev_loss are then */
                              /*   taken from it instead of the host values (and nothing runs if it says the */
                              /*   stopping rule already fired), so consecutive calls chain without a host   */
                              /*   read-back in between.  Must differ from `status`.                         */
    float* logits_support;    /* optional [n_support, n_classes]: logits of the LAST epoch (pre-update W)  */
    void* workspace;
    int64_t workspace_bytes;
    int32_t cta_budget;       /* 0 = the fastest shape for a run that has the GPU to itself (82-98 CTAs at paper  */
                              /*   sizes).  > 0: the paper-size kernel uses at most this many CTAs (fatter row /  */
                              /*   column CTAs, ~10-20 % slower per epoch) so that the cooperative launches of      */
                              /*   several runs sharing the GPU are resident together (3 runs: 148 / 3 = 49).       */
    int32_t reserved0;
} sr_head_args;

int64_t sr_head_workspace_bytes(const sr_head_args* a);
/* Host only (no device needed): out4 = {kernel sr_head_run picks: 0 fp32 SIMT tiles, 1 paper-size persistent kernel, 2 its
 * cluster variant, 3 tensor-core head; rows per row CTA; feature columns per column CTA; CTAs of the cooperative launch}. */
int32_t sr_head_plan(const sr_head_args* a, int32_t* out4);
/* Persistent kernel: loops epochs on the device until the stopping rule fires or max_epochs is reached. */
int32_t sr_head_run(const sr_head_args* a, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Query / base scoring on cached features (validate / eval_base / accuracy:
 * language_eval.py:18-69, eval/util.py:26-40)
 * ---------------------------------------------------------------------------------------------- */
typedef struct sr_eval_args {
    const float* feat;      /* [n, dim]                                                                  */
    const float* weight;    /* [n_classes, dim]                                                          */
    const int64_t* labels;  /* [n]                                                                       */
    int32_t n, dim, n_classes;
    float* logits;          /* required [n, n_classes]                                                   */
    int32_t* pred;          /* [n] argmax (lowest index wins ties, as torch.argmax)                      */
    int32_t* counts;        /* [2]: += top-1 hits, += top-5 hits                                         */
    float* loss_sum;        /* [1]: += sum over rows of (logsumexp - z_y)                                */
    int64_t* confusion;     /* optional [conf_dim, conf_dim] += 1 at (label, pred)                       */
    int32_t conf_dim;
    void* workspace;        /* optional: sr_eval_workspace_bytes(n, dim, n_classes) bytes let large problems run  */
    int64_t workspace_bytes;/*   their logits GEMM on tcgen05 (error-compensated bf16x3); NULL = fp32 SIMT tiles */
} sr_eval_args;

/* Bytes of workspace with which sr_eval_logits takes the tensor-core path for this shape; 0 = the shape stays on the
 * SIMT kernel (small problems are launch-bound anyway). */
int64_t sr_eval_workspace_bytes(int32_t n, int32_t dim, int32_t n_classes);
int32_t sr_eval_logits(const sr_eval_args* a, void* stream);
/* Same scoring on logits the caller already has (feat / weight / dim ignored): eval/util.py accuracy :26-40. */
int32_t sr_score_logits(const sr_eval_args* a, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Per-op surface (one launch per reference library call; used by the autograd wrappers behind the
 * drop-in modules so the UNMODIFIED reference loop runs on the B200 modules)
 * ---------------------------------------------------------------------------------------------- */
/* LangPuller.forward, semantic mode (resnet_language.py:75-82):
 * pullers = softmax(novel_embeds @ base_embeds^T / temperature, dim=1) @ base_weight. */
int32_t sr_semantic_pullers(const float* novel_embeds, const float* base_embeds, const float* base_weight,
                            int32_t n_novel, int32_t n_base, int32_t embed_dim, int32_t dim, float temperature,
                            int32_t mask_diagonal, float* pullers, void* stream);
/* nn.Linear forward / weight+bias backward (classifier :140,187; LinearMap :12-18):
 * y[n,m] = x[n,k] w[m,k]^T + bias;  dw[m,k] = dy^T x, dbias[m] = column sums of dy (dbias may be NULL). */
int32_t sr_linear_fwd(const float* x, const float* w, const float* bias, int32_t n, int32_t k, int32_t m, float* y,
                      void* stream);
int32_t sr_linear_bwd(const float* dy, const float* x, int32_t n, int32_t k, int32_t m, float* dw, float* dbias,
                      void* stream);
/* learn_mapping.py:50-62 pieces: nn.MSELoss (mean) value + gradient, and the momentum-free SGD step with weight decay.
 * loss[0] += mean((y-target)^2) (caller zeroes it); dy = 2 (y - target) / n;  param -= lr * (grad + weight_decay * param). */
int32_t sr_mse_grad(const float* y, const float* target, int64_t n, float* dy, float* loss, void* stream);
int32_t sr_sgd_update(float* param, const float* grad, int64_t n, float lr, float weight_decay, void* stream);
/* learn_mapping.py:41-67 in ONE launch: `epochs` full-batch SGD steps (lr, weight_decay, no momentum) on nn.MSELoss(mean)
 * of y = x weight^T + bias against target.  x [n,e], target [n,d], weight [d,e] and bias [d] updated in place,
 * loss_trace [epochs] (optional) = the loss BEFORE each step.  Output dimensions are independent, so every CTA fits its own
 * rows of `weight` out of shared memory without any exchange.  sr_fit_linear_map_workspace_bytes returns 0 when n x e does
 * not fit in shared memory (use the per-op entry points above then). */
int64_t sr_fit_linear_map_workspace_bytes(int32_t n, int32_t e, int32_t d, int32_t epochs);
int32_t sr_fit_linear_map(const float* x, const float* target, float* weight, float* bias, int32_t n, int32_t e, int32_t d,
                          int32_t epochs, float lr, float weight_decay, float* loss_trace, void* workspace,
                          int64_t workspace_bytes, void* stream);
/* out[0] = sum((a-b)^2)  (torch.norm(a-b)**2, :90,232,239). */
int32_t sr_sqdist(const float* a, const float* b, int64_t n, float* out, void* stream);
/* out = (a-b) * scale * gout[0] * (sq ? 1/sqrt(sq[0]), 0 when sq[0]==0 : 1): backward of s*||a-b||^2 (sq NULL,
 * scale = 2s) and of s*||a-b|| (sq = the forward's squared norm, scale = s; torch's zero-subgradient at 0). */
int32_t sr_diff_scale(const float* a, const float* b, int64_t n, float scale, const float* gout, const float* sq,
                      float* out, void* stream);
/* out[n,dim] = (x qt^T) qt: projection of rows on span(base) (get_projected_weight :92-97); self-adjoint. */
int32_t sr_project_rows(const float* x, const float* qt, int32_t n, int32_t q_rows, int32_t dim, float* out,
                        void* stream);

/* ------------------------------------------------------------------------------------------------
 * Host helper (no device work): replay of PyTorch's CPU mt19937 Bernoulli stream for the dropout / DropBlock
 * keep-masks of the train-mode epoch (resnet_language.py:292-299, 311-325), bit-identical to
 * tensor.bernoulli_(p) (kind 0: double p, two 32-bit draws per element) and torch.bernoulli(float p tensor)
 * (kind 1: one draw per element); kind 2 skips n 32-bit draws without output.  `state_blob` = bytes of
 * torch.get_rng_state(), advanced in place.
 * out: HOST uint8[n] (1 = drawn one).  Returns the number of ones, or -1 on a malformed state.
 * ---------------------------------------------------------------------------------------------- */
int64_t sr_host_bernoulli(void* state_blob, int64_t blob_bytes, int32_t kind, double p, int64_t n, uint8_t* out);

/* Host helper: DropBlock._compute_block_mask (resnet_language.py:327-352) on the drawn seeds.
 * seeds: HOST uint8[planes, hs, ws] (1 = block seed); keep: HOST uint8[planes, hs+bs-1, ws+bs-1], 0 wherever a
 * bs x bs block anchored at a seed covers the pixel, 1 elsewhere (= 1 - padded mask).  Returns the number of ones in
 * `keep` (count_ones of :321), or -1 on bad arguments. */
int64_t sr_host_dropblock(const uint8_t* seeds, int64_t planes, int32_t hs, int32_t ws, int32_t bs, uint8_t* keep);

/* ------------------------------------------------------------------------------------------------
 * Jump-ahead for PyTorch's CPU generator (mt19937): what lets the keep-masks be drawn ON THE DEVICE from the host
 * generator's state (sr_device_bernoulli below) while the host generator is moved past them without drawing anything.
 * table: n_polys polynomials of 624 uint32 (19937 coefficient bits each), entry w-1 = t^(w * SR_MT_JUMP_WORDS) mod the
 * generator's characteristic polynomial; computed on the host (cached inside the library, ~4 ms per entry the first time).
 * ---------------------------------------------------------------------------------------------- */
#define SR_MT_JUMP_WORDS (1 << 18)
int64_t sr_mt_jump_table_bytes(int32_t n_polys);
int32_t sr_mt_jump_table(uint32_t* table_host, int32_t n_polys);
/* state_blob (bytes of torch.get_rng_state(), HOST) is advanced by n_words 32-bit draws; the result is byte-identical to
 * torch's own state after that many draws.  Needs n_polys >= n_words / SR_MT_JUMP_WORDS table entries (HOST pointer). */
int32_t sr_host_mt_advance(void* state_blob, int64_t blob_bytes, int64_t n_words, const uint32_t* table_host,
                           int32_t n_polys);

/* Device replay of the generator's Bernoulli stream: the keep-masks of one train-mode forward (resnet_language.py:292-299,
 * 311-325) drawn on the GPU, bit-identical to what torch's CPU generator would produce from `state_blob`, in region
 * order.  kind 0: tensor.bernoulli_(double p), two words per element; kind 1: torch.bernoulli(float32 p tensor), one word;
 * kind 2: n words skipped.  out: DEVICE uint8[n] (1 = drawn one), NCHW order = draw order.  Every region but the last must
 * use an even number of words.  state_blob (HOST) is advanced in place past all regions, like sr_host_bernoulli does.
 * table_dev / table_host: the same sr_mt_jump_table on the device and on the host, n_polys >= total words / SR_MT_JUMP_WORDS. */
typedef struct sr_mask_region {
    int32_t kind;
    int32_t reserved;
    double p;
    int64_t n;
    uint8_t* out;
} sr_mask_region;
int64_t sr_device_bernoulli_workspace_bytes(int64_t total_words);
int32_t sr_device_bernoulli(void* state_blob, int64_t blob_bytes, const sr_mask_region* regions, int32_t n_regions,
                            const uint32_t* table_dev, const uint32_t* table_host, int32_t n_polys, void* workspace,
                            int64_t workspace_bytes, void* stream);
/* DropBlock._compute_block_mask (resnet_language.py:327-352) + the numel / kept scale (:321-323) on the device.
 * seeds: DEVICE uint8 [planes, hs, ws]; keep: DEVICE uint8 [planes, hs+bs-1, ws+bs-1]; scale_out: DEVICE, 16 bytes, 16-byte
 * aligned: [0] = fp32(numel) / fp32(kept) as the reference computes it (the rest is scratch). */
int32_t sr_dropblock_keep(const uint8_t* seeds, int64_t planes, int32_t hs, int32_t ws, int32_t bs, uint8_t* keep,
                          float* scale_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SRB200_H_ */
