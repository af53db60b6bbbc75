/*
 * adnb200.h - C ABI of the B200-native (sm_100a) ADN-SSD token mixer, Haar WTConv2d and
 * threshold-count kernels.  Plain pointers and sizes only; no torch / CUDA types in signatures
 * (`stream` is a cudaStream_t passed as void*).
 *
 * The reference (kanyu369/ADNM-UNet) is pure Python and has no FFI; each entry point below
 * replaces the body of one reference Python method, and the reference-side binding a maintainer
 * would add is the ctypes stub shown in INTEGRATION.md.
 *
 *   adnssd_forward / adnssd_backward   <- models/ADNssd.py:302-462  Mamba2.forward (+ its autograd)
 *   wtconv_forward / wtconv_backward   <- models/WTConv2d.py:100-153 WTConv2d.forward (+ its autograd)
 *   adn_threshold_counts               <- datasets/Shanghai_metrics.py:45-47,105-114 float2int + _cal_frame
 *   adn_sumsq_f32 / adn_adamw_flat     <- train.py:140-145 clip_grad_norm_ + AdamW.step + zero_grad (train_untils.py:35-42)
 *
 * Conventions:
 *  - every pointer is a DEVICE pointer on the current CUDA device unless stated otherwise;
 *  - the caller owns every buffer; the library allocates nothing, keeps no pointer after return;
 *  - work is enqueued on `stream` with no host synchronisation (CUDA-graph capturable);
 *  - parameters and parameter gradients are float32 in the reference's native state_dict layout;
 *    gradient buffers are OVERWRITTEN, never accumulated into;
 *  - return 0 on success, non-zero ADN_ERR_* otherwise, message via adn_last_error() (thread-local).
 *    Unsupported shapes are errors: there is no CPU fallback and no alternative backend.
 */
#ifndef ADNB200_H_
#define ADNB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ADNB200_ABI_VERSION 1

enum { ADN_OK = 0, ADN_ERR_SHAPE = 1, ADN_ERR_DTYPE = 2, ADN_ERR_NULL = 3, ADN_ERR_CUDA = 4, ADN_ERR_ARCH = 5 };

/* Activation dtype at the seam (u, out, dout, du, x, y, ...).  ADN_F32 is the "check mode": every
 * intermediate and every contraction in fp32 (tolerance 1e-4 vs the reference).  ADN_BF16: bf16 I/O and
 * saved intermediates, bf16 tensor-core operands, fp32 accumulation / statistics / parameter gradients. */
enum { ADN_F32 = 0, ADN_BF16 = 1 };

/* ------------------------------------------------------------------ ADN-SSD mixer ---------- */

typedef struct AdnShape {
  int32_t B;      /* batch                                                                        */
  int32_t H, W;   /* token grid, L = H*W (the reference only ever passes H == W, ADNMUNet.py:120) */
  int32_t D;      /* d_model                                                                      */
  int32_t Di;     /* d_inner = expand * d_model          (models/ADNssd.py:84), Di % 4 == 0       */
  int32_t P;      /* headdim, Di % P == 0                (:85,:90-91)                             */
  int32_t G;      /* ngroups; only 2 is supported (the branch taken at :278)                      */
  int32_t N;      /* d_state                             (:86), (G*N) % 4 == 0                    */
  int32_t dtype;  /* ADN_F32 | ADN_BF16                                                           */
  int32_t flags;  /* reserved, must be 0                                                          */
} AdnShape;

/* The 21 state_dict tensors of one Mamba2 mixer (models/ADNssd.py:100-248), float32, native layout.
 * scale / shift / alpha2 are declared by the reference but never read (:227-228,:246); they may be NULL. */
typedef struct AdnWeights {
  const float* dt_bias;        /* (nh)                 nh = Di / P                     */
  const float* A_log;          /* (nh)                                                 */
  const float* D;              /* (nh)                                                 */
  const float* scale;          /* ()   unused                                          */
  const float* shift;          /* ()   unused                                          */
  const float* alpha1;         /* ()                                                   */
  const float* alpha2;         /* ()   unused                                          */
  const float* in_proj_w;      /* (2Di + 2GN + nh, D)  rows ordered [z | x | B | C | dt] */
  const float* conv_13_x1_w;   /* (Di/4, 1, 1, 3)                                      */
  const float* conv_31_x1_w;   /* (Di/4, 1, 3, 1)                                      */
  const float* conv_13_x2_w;   /* (Di/4, 1, 1, 3)                                      */
  const float* conv_31_x2_w;   /* (Di/4, 1, 3, 1)                                      */
  const float* conv_13_bc1_w;  /* (GN/2, 1, 1, 3)                                      */
  const float* conv_31_bc1_w;  /* (GN/2, 1, 3, 1)                                      */
  const float* conv_13_bc2_w;  /* (GN/2, 1, 1, 3)                                      */
  const float* conv_31_bc2_w;  /* (GN/2, 1, 3, 1)                                      */
  const float* conv2d_w;       /* ((Di + 2GN)/2, 1, 3, 3)                              */
  const float* norm_w;         /* (Di)                                                 */
  const float* norm_b;         /* (Di)                                                 */
  const float* conv2d_z_w;     /* (Di, 1, 3, 3)                                        */
  const float* out_proj_w;     /* (D, 2Di)                                             */
} AdnWeights;

/* Same field order and shapes; every non-NULL pointer receives that parameter's gradient (overwritten).
 * scale / shift / alpha2 are ignored (the reference leaves their .grad None). */
typedef struct AdnWeightGrads {
  float* dt_bias; float* A_log; float* D; float* scale; float* shift; float* alpha1; float* alpha2;
  float* in_proj_w;
  float* conv_13_x1_w; float* conv_31_x1_w; float* conv_13_x2_w; float* conv_31_x2_w;
  float* conv_13_bc1_w; float* conv_31_bc1_w; float* conv_13_bc2_w; float* conv_31_bc2_w;
  float* conv2d_w; float* norm_w; float* norm_b; float* conv2d_z_w; float* out_proj_w;
} AdnWeightGrads;

/* Sizes (bytes) of the caller-allocated `saved` (forward -> backward) and `workspace` (scratch, may be
 * shared between calls on one stream; must hold max(fwd, bwd)) buffers.  256-byte alignment required. */
int adnssd_workspace_bytes(const AdnShape* s, size_t* saved_bytes, size_t* fwd_workspace_bytes,
                           size_t* bwd_workspace_bytes);

/* out(B,L,D) = Mamba2.forward(u(B,L,D), H, W).  `saved` may be NULL for inference (no backward). */
int adnssd_forward(const AdnShape* s, const AdnWeights* w, const void* u, void* out, void* saved,
                   void* workspace, void* stream);

/* du(B,L,D) and all parameter gradients from dout(B,L,D); `saved` is what adnssd_forward wrote. */
int adnssd_backward(const AdnShape* s, const AdnWeights* w, const void* u, const void* saved,
                    const void* dout, void* du, const AdnWeightGrads* g, void* workspace, void* stream);

/* ------------------------------------------------------------------ WTConv2d --------------- */

#define ADN_WT_MAX_LEVELS 8

typedef struct WtShape {
  int32_t B, C, H, W;   /* NCHW input == output shape (stride 1, in_channels == out_channels, WTConv2d.py:67) */
  int32_t k;            /* odd depthwise kernel size (5 or 3 in ADNM-UNet)                                   */
  int32_t levels;       /* wt_levels, 1..ADN_WT_MAX_LEVELS                                                    */
  int32_t has_bias;     /* base_conv.bias present                                                             */
  int32_t dtype;        /* ADN_F32 | ADN_BF16 (activations); parameters float32                               */
} WtShape;

/* state_dict of one WTConv2d minus the frozen Haar filters wt_filter / iwt_filter (fixed db1 butterflies,
 * models/WTConv2d.py:9-29; the host module keeps them as non-trainable Parameters for checkpoint parity). */
typedef struct WtWeights {
  const float* base_conv_w;                       /* (C,1,k,k)  */
  const float* base_conv_b;                       /* (C) | NULL */
  const float* base_scale_w;                      /* (1,C,1,1)  */
  const float* wavelet_conv_w[ADN_WT_MAX_LEVELS]; /* (4C,1,k,k) */
  const float* wavelet_scale_w[ADN_WT_MAX_LEVELS];/* (1,4C,1,1) */
} WtWeights;

typedef struct WtWeightGrads {
  float* base_conv_w; float* base_conv_b; float* base_scale_w;
  float* wavelet_conv_w[ADN_WT_MAX_LEVELS]; float* wavelet_scale_w[ADN_WT_MAX_LEVELS];
} WtWeightGrads;

int wtconv_workspace_bytes(const WtShape* s, size_t* saved_bytes, size_t* fwd_workspace_bytes,
                           size_t* bwd_workspace_bytes);
int wtconv_forward(const WtShape* s, const WtWeights* w, const void* x, void* y, void* saved,
                   void* workspace, void* stream);
int wtconv_backward(const WtShape* s, const WtWeights* w, const void* x, const void* saved, const void* dy,
                    void* dx, const WtWeightGrads* g, void* workspace, void* stream);

/* ------------------------------------------------------------------ threshold counts ------- */

/* table[t*4 + {0,1,2,3}] = TP, FN, FP, TN over n float32 elements with
 * q(v) = (uint16)(clamp(v,0,1) * value_scale), event = q >= thresholds[t]   (Shanghai_metrics.py:45-47,105-114).
 * `thresholds` is a HOST pointer (n_thresholds <= 8); `table` is a DEVICE int64 buffer, overwritten. */
int adn_threshold_counts(const float* obs, const float* sim, int64_t n, const int32_t* thresholds,
                         int32_t n_thresholds, float value_scale, int64_t* table, void* stream);

/* One batch of SimplifiedEvaluator.evaluate (datasets/Shanghai_metrics.py:49-103) on device, replacing the per-batch
 * .cpu().numpy() + Python loops over batch x frame x threshold of validate.py:100-118:
 *   true_batch, pred_batch: [batch][seq_len][frame_elems] float32 (declared argument order of evaluate, :49);
 *   table[t*4 + {TP,FN,FP,TN}] += the counts of every frame (same quantisation as adn_threshold_counts), int64, ACCUMULATED
 *     (zero it when the evaluator is reset);
 *   mse_t[t] += sum_b mean((clip(pred)*scale - clip(true)*scale)^2) of frame (b, t)  (:116-121), fp64, ACCUMULATED: `done`
 *     (:276) computes RMSE = mean_t sqrt(mse_t[t] / total_samples).
 * `thresholds` is a HOST pointer; batch * seq_len <= 65535 per call. */
int adn_eval_batch(const float* true_batch, const float* pred_batch, int64_t batch, int32_t seq_len, int64_t frame_elems,
                   const int32_t* thresholds, int32_t n_thresholds, float value_scale, int64_t* table, double* mse_t,
                   void* stream);

/* ------------------------------------------------------------------ training-step tail ----- */

/* Flat-buffer tail of one data-parallel training step, replacing train.py:140-145 of the reference
 * (clip_grad_norm_ -> AdamW.step -> zero_grad; AdamW hyper-parameters of train_untils.py:35-42) for the batch-sharded
 * trainer: all live parameters / gradients / moments are contiguous float32 buffers of n elements.
 *
 * adn_sumsq_f32: out[0] = sum(x[i]^2), deterministic (fixed reduction tree, no atomics) so that every rank derives the
 *   same clip factor from the same all-reduced gradients.  partial_ws: adn_sumsq_workspace_floats() floats. */
int adn_sumsq_workspace_floats(void);
int adn_sumsq_f32(const float* x, int64_t n, float* partial_ws, float* out, void* stream);
/* One pass: g_eff = g * grad_scale * min(1, max_norm / (sqrt(*sumsq) * grad_scale + 1e-6))  (max_norm <= 0: no clip;
 *   grad_scale = 1 / world when the buffer holds the SUM over ranks), decoupled weight decay, Adam moments with bias
 *   correction for step number `step` (>= 1), parameter update, and g <- 0.  If norm_out != NULL it receives the
 *   unclipped global norm sqrt(*sumsq) * grad_scale (device float; what train.py:141 reads back with .item()). */
int adn_adamw_flat(float* p, float* g, float* m, float* v, int64_t n, const float* sumsq, float* norm_out, float lr,
                   float beta1, float beta2, float eps, float weight_decay, int32_t step, float grad_scale,
                   float max_norm, void* stream);

/* ------------------------------------------------------------------ Block pre-norm -------- */

/* Standalone RMSNorm of the reference README (README.md:22-30, the `norm_layer` of create_block, models/ADNMUNet.py:278) fused
 * with the Block's scalar affine (models/ADNMUNet.py:149,155):
 *     y = scale * (x * rsqrt(mean(x^2) + eps) * weight) + shift          per token over D
 * x, y, dy, dx: [tokens][D] contiguous, dtype ADN_F32 or ADN_BF16; weight / scale / shift / rstd / gradients fp32.
 * scale / shift may be NULL (1 and 0: the bare RMSNorm module).  rstd ([tokens], optional in forward) is what backward needs
 * besides x.  dweight / dscale / dshift are OVERWRITTEN (dscale / dshift may be NULL).  D % 4 == 0, D <= 8192.
 * dres (optional, [tokens][D]) is added to dx: the gradient reaching x through the Block's residual path (ADNMUNet.py:152),
 * so the sum of the two branches is formed in fp32 and rounded once. */
int adn_rmsnorm_forward(const void* x, const float* weight, const float* scale, const float* shift, void* y, float* rstd,
                        int64_t tokens, int32_t D, float eps, int32_t dtype, void* stream);
int adn_rmsnorm_backward(const void* x, const float* weight, const float* scale, const float* rstd, const void* dy,
                         const void* dres, void* dx, float* dweight, float* dscale, float* dshift, int64_t tokens, int32_t D,
                         int32_t dtype, void* stream);

/* Residual mix of the Block (models/ADNMUNet.py:152,158,161): out = (beta1 * x + beta2 * y) * gamma[c]; gamma NULL = no
 * per-channel scale.  x, y, out, dout, dx, dy: [tokens][D] contiguous (dtype); beta1 / beta2: device scalars; gamma (D) fp32.
 * Backward OVERWRITES dx, dy, dbeta1, dbeta2 (scalars, accumulated in fp64 inside) and dgamma.  ws: 64 bytes of scratch. */
int adn_residual_forward(const void* x, const void* y, const float* beta1, const float* beta2, const float* gamma, void* out,
                         int64_t tokens, int32_t D, int32_t dtype, void* stream);
int adn_residual_backward(const void* x, const void* y, const void* dout, const float* beta1, const float* beta2,
                          const float* gamma, void* dx, void* dy, float* dbeta1, float* dbeta2, float* dgamma, void* ws,
                          int64_t tokens, int32_t D, int32_t dtype, void* stream);

/* FeedForward of the Block (models/model_untils.py:172-197) on TOKEN-MAJOR activations x, y: [B][H*W][D] - the layout the
 * mixer uses, so that the NCHW round trip of models/ADNMUNet.py:158 disappears:
 *   y = project_out( gelu(a[:, :C4/2]) * sigmoid(a[:, C4/2:]) ),  a = dwconv3x3(project_in(x)),  C4 = 4 * D in ADNM-UNet.
 * Weights are the module's state_dict tensors, fp32, native layout (1x1 conv weights are [out][in]). */
typedef struct AdnFfnShape {
  int32_t B, H, W;   /* token grid                                        */
  int32_t D;         /* dim, D % 4 == 0                                   */
  int32_t C4;        /* project_in output channels (2 * hidden), C4 % 8 == 0 */
  int32_t dtype;     /* ADN_F32 | ADN_BF16                                */
} AdnFfnShape;
typedef struct AdnFfnWeights {   /* also used for the gradients (same shapes, OVERWRITTEN) */
  void* w_in;    /* project_in.conv.weight  (C4, D, 1, 1)   */
  void* b_in;    /* project_in.conv.bias    (C4)            */
  void* w_dw;    /* dwconv.conv.weight      (C4, 1, 3, 3)   */
  void* b_dw;    /* dwconv.conv.bias        (C4)            */
  void* w_out;   /* project_out.conv.weight (D, C4/2, 1, 1) */
  void* b_out;   /* project_out.conv.bias   (D)             */
} AdnFfnWeights;
int adn_ffn_workspace_bytes(const AdnFfnShape* s, size_t* saved_bytes, size_t* fwd_workspace_bytes, size_t* bwd_workspace_bytes);
/* saved may be NULL for inference. */
int adn_ffn_forward(const AdnFfnShape* s, const AdnFfnWeights* w, const void* x, void* y, void* saved, void* workspace, void* stream);
int adn_ffn_backward(const AdnFfnShape* s, const AdnFfnWeights* w, const void* x, const void* saved, const void* dy, void* dx,
                     const AdnFfnWeights* grads, void* workspace, void* stream);

/* Linear over tokens (the Block's out_proj, models/ADNMUNet.py:108-110,162-163): y[tokens][N] = x[tokens][K] W[N][K]^T + bias.
 * bf16 activations with K % 8 == 0 and N % 8 == 0 run on the tcgen05 GEMM; dw / dbias are OVERWRITTEN (dbias may be NULL). */
int adn_linear_workspace_bytes(int64_t tokens, int32_t K, int32_t N, int32_t dtype, size_t* workspace_bytes);
int adn_linear_forward(const void* x, const float* w, const float* bias, void* y, int64_t tokens, int32_t K, int32_t N,
                       int32_t dtype, void* workspace, void* stream);
int adn_linear_backward(const void* x, const float* w, const void* dy, void* dx, float* dw, float* dbias, int64_t tokens,
                        int32_t K, int32_t N, int32_t dtype, void* workspace, void* stream);

/* ------------------------------------------------------------------ fused attention --------- */

/* Core of StandardAttention (models/ADNssd.py:41-46; the Attention bridges, models/ADNMUNet.py:172-238), SURVEY.md 8(f)3:
 *   out[b][i][h*dh + d] = sum_j softmax_j(scale * q_i . k_j) v_j[d]      per sample b and head h
 * qkv: [B][L][3 * heads * dh], the packed to_qkv output (q | k | v, each in '(h d)' order); out / dout: [B][L][heads * dh];
 * dqkv: like qkv (OVERWRITTEN); lse: [B][heads][L] fp32, written by forward (may be NULL for inference), read by backward.
 * dh in {4, 8, 16} (ADNM-UNet: 4).  The L x L score matrix is never materialised. */
int adn_sdpa_forward(const void* qkv, void* out, float* lse, int32_t B, int32_t L, int32_t heads, int32_t dh, float scale,
                     int32_t dtype, void* stream);
int adn_sdpa_backward(const void* qkv, const void* out, const float* lse, const void* dout, void* dqkv, int32_t B, int32_t L,
                      int32_t heads, int32_t dh, float scale, int32_t dtype, void* stream);

/* ------------------------------------------------------------------ conv stages (WTLayer / PatchEmbed / OutProj) ---- */

/* the full-resolution stages around the native WTConv2d - WTLayer (models/model_untils.py:358-426),
 * PatchEmbed (:226-314), OutProj (:799-892).  All activations at this seam are token-major (B, L, C) = channels-last,
 * except where a tensor feeds / leaves wtconv_forward (NCHW planes). */

/* Dense 3x3 convolution, stride 1, zero padding 1 (Conv2dLayer.conv, models/model_untils.py:71-93) on channels-last
 * activations: x (B, H*W, Cin) -> y (B, H*W, Cout).  w: nn.Conv2d weight (Cout, Cin, 3, 3) fp32; bias (Cout) or NULL;
 * gamma (Cin) or NULL: per-input-channel scale applied to x before the conv (the layer scale of WTLayer :420-421 /
 * OutProj :882-883), folded into the weights.  bf16 with Cin % 8 == 0, Cout % 8 == 0 (<= 256) and a grid that tiles into
 * 64- / 128-token boxes runs as an implicit GEMM on tcgen05 (adn_conv3x3_path() == 1); everything else on CUDA cores. */
typedef struct AdnConvShape {
  int32_t B, H, W, Cin, Cout;
  int32_t dtype; /* ADN_F32 | ADN_BF16 */
} AdnConvShape;
int adn_conv3x3_path(const AdnConvShape* s);
int adn_conv3x3_workspace_bytes(const AdnConvShape* s, size_t* workspace_bytes);
int adn_conv3x3_forward(const AdnConvShape* s, const void* x, const float* w, const float* bias, const float* gamma, void* y,
                        void* workspace, void* stream);
/* dx may be NULL (PatchEmbed: the input is data); dw is OVERWRITTEN; dbias / dgamma may be NULL. */
int adn_conv3x3_backward(const AdnConvShape* s, const void* x, const float* w, const float* gamma, const void* dy, void* dx,
                         float* dw, float* dbias, float* dgamma, void* workspace, void* stream);

/* (B, L, C1) [+ (B, L, C2)] token-major -> (B, C1 + C2, H, W) planes: out[b][c][p] = g1 x[b][p][c] | g2 res[b][p][c - C1]
 * (WTLayer :404-414: cat(gama1 x, gama2 residual) followed by the NCHW permute; g1 / g2 device scalars or NULL = 1). */
int adn_nchw_pack_forward(const void* x, const void* res, const float* g1, const float* g2, void* out, int32_t B, int64_t HW,
                          int32_t C1, int32_t C2, int32_t dtype, void* stream);
/* workspace: 64 bytes.  dx / dres / dg1 / dg2 may be NULL. */
int adn_nchw_pack_backward(const void* x, const void* res, const float* g1, const float* g2, const void* dout, void* dx, void* dres,
                           float* dg1, float* dg2, void* workspace, int32_t B, int64_t HW, int32_t C1, int32_t C2, int32_t dtype,
                           void* stream);

/* nn.InstanceNorm2d statistics (eps, biased variance, no affine): stats[plane] = (mean, rstd) fp32, planes = B * C. */
int adn_plane_stats(const void* y, float* stats, int64_t planes, int64_t HW, float eps, int32_t dtype, void* stream);

/* out[b][p][c] = gamma[c] * (alpha * act(u) + beta * xs[b][c][p]),  u = scale * (y - mean) * rstd + shift (stats NULL: u = y).
 * y, xs: (B, C, HW) planes; out: (B, HW, C) token-major; act 0 none, 1 GELU; scale / shift / gamma may be NULL.
 *   WTLayer    :416       alpha * wtconv(x) + beta * shortcut  with WTConvLayer's scale * norm(x) + shift (:112-113)
 *   PatchEmbed :303-307   alpha1 * GELU(wtconv(x)) + beta1 * x;  (alpha2 * IN(wtconv(s)) + beta2 * s) * gamma
 *   OutProj    :880-883   (alpha * GELU(IN(wtconv(x))) + beta * shortcut) * gamma */
int adn_plane_mix_forward(const void* y, const void* xs, const float* stats, const float* scale, const float* shift,
                          const float* alpha, const float* beta, const float* gamma, void* out, int32_t B, int32_t C, int64_t HW,
                          int32_t act, int32_t dtype, void* stream);
int adn_plane_mix_workspace_bytes(int32_t B, int32_t C, size_t* workspace_bytes);
/* dout: (B, HW, C); dy / dxs: (B, C, HW), either may be NULL; dscal[4] = (dscale, dshift, dalpha, dbeta); dgamma (C) or NULL. */
int adn_plane_mix_backward(const void* y, const void* xs, const float* stats, const float* scale, const float* shift,
                           const float* alpha, const float* beta, const float* gamma, const void* dout, void* dy, void* dxs,
                           float* dscal, float* dgamma, void* workspace, int32_t B, int32_t C, int64_t HW, int32_t act,
                           int32_t dtype, void* stream);

/* Activation after a conv / Linear: kind 1 = GELU (erf form), 2 = Swish x * sigmoid(beta * x) (models/model_untils.py:162-169,
 * beta a device scalar or NULL = 1).  Backward workspace: 64 bytes; dbeta may be NULL. */
int adn_act_forward(const void* x, void* y, int64_t n, int32_t kind, const float* beta, int32_t dtype, void* stream);
int adn_act_backward(const void* x, const void* dy, void* dx, int64_t n, int32_t kind, const float* beta, float* dbeta,
                     void* workspace, int32_t dtype, void* stream);

/* Grouped convolution with 4 channels per group (the `groups = dim / 4` convs of the EncoderToDecoder bridges,
 * models/model_untils.py:621-675; cuDNN runs them as one launch per group): x, y (B, H*W, C) channels-last, w (C, 4, kh, kw) fp32 in
 * nn.Conv2d's layout, bias (C) or NULL, kh / kw in {1, 3}, stride 1, zero padding kh/2, kw/2.  Backward: dx may be NULL, dw and
 * dbias (may be NULL) are OVERWRITTEN. */
int adn_gconv4_forward(const void* x, const float* w, const float* bias, void* y, int32_t B, int32_t H, int32_t W, int32_t C,
                       int32_t kh, int32_t kw, int32_t dtype, void* stream);
int adn_gconv4_backward(const void* x, const float* w, const void* dy, void* dx, float* dw, float* dbias, int32_t B, int32_t H,
                        int32_t W, int32_t C, int32_t kh, int32_t kw, int32_t dtype, void* stream);

/* ------------------------------------------------------------------ misc ------------------- */

const char* adn_last_error(void);
int adn_abi_version(void);
/* 1 if the current device is compute capability 10.x (the only target this library is built for). */
int adn_device_supported(void);


/* ------------------------------------------------------------------ diagnostics ------------ */

/* Number of CUDA kernels this library has launched in this process (all entry points, all threads). */
unsigned long long adn_launch_count(void);
/* Per-launch device timing: while enabled, every launch is bracketed by CUDA events on its stream (not
 * thread-safe; for bench.py / profiling only).  adn_prof_enable(1) clears the log; after a stream sync,
 * adn_prof_get(i, &name, &ms) returns the i-th launch's kernel name (static string) and duration. */
int adn_prof_enable(int on);
int adn_prof_count(void);
int adn_prof_get(int i, const char** name, float* ms);
/* In-kernel phase timers of the tcgen05 / conv kernels (SM-clock cycles of thread 0 of every CTA, summed):
 * adn_phase_enable(1) zeroes and arms them, adn_phase_read copies the 8 x 8 table [kernel][phase] to HOST memory
 * (synchronises the device).  Kernel ids: 0 k_bwd1, 1 k_bwd2, 2 k_bwd4, 3 k_conv_bwd_tile. */
int adn_phase_enable(int on);
int adn_phase_read(unsigned long long* out64);
/* Per-CTA start / end times (GPU globaltimer, ns) of the row kernels of the last armed step, HOST buffer of 3 x 160 x 4
 * values [kernel: 0 k_fconv, 1 k_bconv_du, 2 k_bconv_wg][CTA][start, end, after prologue, MMA loop end]; only filled by -DADN_PHASE_TIMING builds. */
int adn_cta_times_read(unsigned long long* out1920);
/* Hardware self-test of the tcgen05 building blocks (one 128 x N x K bf16 GEMM through shared-memory descriptors
 * and TMEM).  mode 0: A[128][K], B[N][K] (K-major operands); mode 1: A[K][128], B[K][N] (MN-major operands).
 * C is float[128][N]; *status (device int) is set to 1 if the MMA completion barrier timed out. */
int adn_selftest_umma(int mode, int N, int K, const void* A, const void* B, float* C, int* status, void* stream);
/* Same, with the operands staged in a row-padded buffer ([chunk][pitch rows][8]) and addressed through descriptors whose
 * start address is shifted by whole rows (the addressing the fused conv-as-GEMM kernels rely on).
 * mode 0: A[pitch][K], B[pitch][K]:  C = A[shiftA:shiftA+128] . B[shiftB:shiftB+N]^T
 * mode 1: A[pitch][128], B[pitch][N]: C[m][n] = sum_{k<K} A[shiftA+k][m] * B[shiftB+k][n] */
int adn_selftest_umma_shift(int mode, int N, int K, int pitch, int shiftA, int shiftB, const void* A, const void* B,
                            float* C, int* status, void* stream);

/* Diagnostic kernel-family switches ("rowconv", "row_wide", "bwd_ws", "wide", "rows_per_cta", "du_dbg"): same meaning as the
 * ADN_<NAME> environment variables, which are read once at first use.  Flip only between whole forward + backward passes. */
int adn_set_option(const char* name, int value);
/* Kernel family that serves a shape: 0 generic CUDA-core kernels (the fp32 check mode; bf16 shapes whose rows are not
 * whole 16-byte pieces), 1 tile kernels and 2 row kernels (d_model 32, fused tcgen05), 3 wide path (d_model >= 64 or
 * d_state 128: tcgen05 GEMMs + bf16 bandwidth kernels); -1 for an invalid shape. */
int adnssd_kernel_family(const AdnShape* s);
/* Standalone entry to the general tcgen05 GEMM of the wide path (bf16 operands, fp32 accumulation in TMEM):
 *   C[b] = alpha * (A0[b] . B0[b] [+ A1[b] . B1[b]])      M x N, K0 (+ K1) deep, `batches` problems, optional split-K.
 * a_mn / b_mn: 0 = operand stored row-major [M or N][K], 1 = stored [K][M or N]; row pitches ld*, batch strides *_bs (elements).
 * c_mode 0: bf16 store, 1: fp32 store, 2: fp32 atomic add into a pre-zeroed C (required for splitk > 1).
 * parity_mask 1 keeps only (m & 1) == (n & 1).  *status (device int, caller-zeroed) is set to 1 on a pipeline time-out. */
int adn_selftest_gemm(int M, int N, int K0, int K1, int a_mn, int b_mn, const void* A0, long long lda0, long long a_bs0,
                      const void* B0, long long ldb0, long long b_bs0, const void* A1, long long lda1, long long a_bs1,
                      const void* B1, long long ldb1, long long b_bs1, void* C, long long ldc, long long c_bs, int c_mode,
                      int batches, int splitk, const float* alpha, int parity_mask, int* status, void* stream);

/* tcgen05.mma issue-rate probe: `ctas` CTAs each issue `iters` 128 x N x 16 bf16 MMAs (mode 0 K-major, 1 MN-major operands,
 * row pitch `pitch`, start shifted by `shift` rows); cycles[cta] (DEVICE int64) = SM clocks from first issue to completion. */
int adn_bench_umma(int mode, int N, int pitch, int shift, int iters, int ctas, long long* cycles, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ADNB200_H_ */
