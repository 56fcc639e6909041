/*
 * wnb200.h -- C-ABI of the B200-native wavenet-speech hot path (libwnb200.so).
 *
 * The reference (paultsw/wavenet-speech) has no FFI layer: its hot path is a set of torch.nn
 * modules (modules/conv_ops.py, block.py, wavenet.py, raw_ctcnet.py, classifier.py,
 * layernorm.py, linear_conv_ops.py) that call ATen/cuDNN.  This header declares the entry points
 * that stand where those library calls stood; each one cites the reference call site it replaces.
 * The Python drop-in modules (wavenet_speech_b200/modules/*) are the only callers.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless it says "host"; the caller owns every buffer,
 *     including workspaces; nothing is allocated, freed or synchronised in here;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*);
 *   - return 0 on success, non-zero on error; wnb200_last_error() gives the message (thread local);
 *   - dtype: WNB200_F32 or WNB200_BF16 = storage type of activations AND weights; accumulation is
 *     always fp32; gradients of parameters are always written as fp32;
 *   - "NCL" = (batch, channels, time) with time contiguous (the reference's layout);
 *     "NLC" = (batch, time, channels) with channels contiguous (the tensor-core path's layout);
 *   - every argument STRUCT starts with `uint32_t struct_size`: the caller sets it to sizeof(the struct) as it
 *     was compiled; a mismatch (binding written against another version of this header) is refused with an
 *     error instead of reading garbage as pointers.  Zero-initialise the struct, then fill it in;
 *   - sm_100a only.  There is no CPU fallback and no other-arch fallback.
 */
#ifndef WNB200_H_
#define WNB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WNB200_F32 0
#define WNB200_BF16 1

/* Two-byte activation / weight format of the tensor-core path.
 *   WNB200_ACT_BF16  : bf16 operands; the residual stream is one bf16 tensor per layer (fastest; also the training
 *                      format).  Rounding of the stream is amplified by every following block: about 2e-2 of the
 *                      logits at 10 blocks and 5e-2..9e-2 at 16-20 with the reference's initialisation.
 *   WNB200_ACT_F16X2 : fp16 operands (same tensor-core rate, 3 more mantissa bits; bf16-rounded weights are exact in
 *                      fp16), the residual stream is carried as an fp16 (hi, lo) PAIR of tensors (hi = fp16(v),
 *                      lo = fp16(v - hi)); dilated taps read hi, the residual projection contracts both halves; the
 *                      gate is evaluated with ex2/rcp instead of tanh.approx.  Holds 2e-2 at 20 blocks (measured
 *                      6e-3..1e-2).  Range: |stream| <= 65504 (saturating). */
#define WNB200_ACT_BF16 0
#define WNB200_ACT_F16X2 1

#define WNB200_EPI_NONE 0  /* y = acc + bias                                              */
#define WNB200_EPI_LEAKY 1 /* y = LeakyReLU_0.01(acc + bias)   (wavenet.py:67-71)         */
#define WNB200_EPI_GATE 2  /* y = tanh(acc_t + b_t) * sigmoid(acc_s + b_s) (block.py:185) */

#define WNB200_EPI_MU 3    /* y = g1 * tanh(g2 * h + g3 * u), g_i = sigmoid(acc_i + b_i), u = tanh(acc_4 + b_4)
                              (MultiplicativeUnit, block.py:213-220); wnb200_taps_fwd_ex only                       */

#define WNB200_PRE_NONE 0   /* source used as stored                                                               */
#define WNB200_PRE_LEAKY 1  /* LeakyReLU(0.01) applied as the source is loaded (wavenet.py:67-71)                  */
#define WNB200_PRE_LNRELU 2 /* ReLU(LayerNorm(source)) applied as it is loaded (ByteNet blocks, block.py:103-110,
                               150-160); needs the source's wnb200_ln_t; wnb200_taps_fwd_ex / _wgrad_ex only        */

#define WNB200_MAX_SRC 4

/* One (input tensor, weight slab, time offset) term of a tap-sum contraction. */
typedef struct {
  const void* x;        /* NCL activations [B, C, T_src] (strided)                        */
  const void* w;        /* weight slab [rows, C] row-major, same dtype as x               */
  int64_t batch_stride; /* elements between batch items of x                              */
  int64_t chan_stride;  /* elements between channels of x (time stride is 1)              */
  int32_t C;            /* channels contracted                                            */
  int32_t T_src;        /* valid time extent of x; reads outside [0, T_src) give 0        */
  int32_t t_off;        /* output frame t reads x[.., t + t_off]                          */
  int32_t pre_act;      /* WNB200_PRE_*                                                   */
} wnb200_src_t;

/* LayerNorm (layernorm.py:25-28) of a source, for WNB200_PRE_LNRELU: per-frame statistics written by wnb200_ln_stats
 * and the module's gamma / beta. */
typedef struct {
  const float* stats; /* [B, T_src, 2] = (mean over channels, 1 / (unbiased std + eps)) */
  const float* gamma; /* [C] */
  const float* beta;  /* [C] */
} wnb200_ln_t;

/* Arguments of wnb200_taps_fwd_ex: wnb200_taps_fwd's plus the ByteNet fusions. */
typedef struct {
  uint32_t struct_size; /* = sizeof(wnb200_taps_t) */
  int32_t dtype, B, T_out, M, nsrc;
  int32_t epilogue;     /* WNB200_EPI_* */
  int32_t accumulate;
  const wnb200_src_t* srcs; /* host, nsrc entries */
  const wnb200_ln_t* ln;    /* host, nsrc entries (read for sources with WNB200_PRE_LNRELU), or NULL */
  const float* bias;
  void* out;
  void* th;
  void* sg;
  const void* residual; /* [B, M, T_out], added to the result after the epilogue (`seq + stack(seq)`, block.py:119,166);
                           NULL: none; not with EPI_GATE */
  const void* mu_h;     /* EPI_MU: h [B, M, T_out].  Weight rows / bias are packed per 32 channels as
                           rows [0, 64): channel r / 4 of the group, unit r % 4 in the order (gate1, gate2, gate3, update),
                           rows [64, 128): channel 16 + (r - 64) / 4 likewise: 4 * ceil(M / 32) * 32 rows */
  int32_t mu_h_ln;      /* EPI_MU: 1 = mu_h is source 0's raw tensor and its LayerNorm + ReLU is applied on the fly */
  int32_t reserved0;
} wnb200_taps_t;

const char* wnb200_last_error(void);
int wnb200_version(void);
/* 0 if the current device is sm_100 (B200); error otherwise. */
int wnb200_check_device(void);

/* ------------------------------------------------------------------------------------------
 * Generic tap-sum contraction (any channel count, kernel width, dilation; fp32 FFMA math):
 *     out[b, m, t] (+)= epi( bias[m] + sum_s sum_c w_s[m, c] * pre(x_s[b, c, t + t_off_s]) )
 * Replaces nn.Conv1d + slice in CausalConv1d/NonCausalConv1d.forward (conv_ops.py:39-44, 74-79),
 * the 1x1 convs and the nn.Linear residual projection with its two transposes (block.py:73-79,
 * conv_ops.py:91-101), the skip bottlenecks (wavenet.py:100), the output stacks (wavenet.py:103)
 * and LinearConv1d.linear (linear_conv_ops.py:39-68).
 * EPI_GATE: weight rows/bias are packed per 64 output channels as [64 tanh rows; 64 sigmoid rows]
 * (2*ceil(M/64)*64 rows in total); `th`/`sg` (optional, may be NULL) receive tanh(.) and
 * sigmoid(.) for the backward pass.  `accumulate`=1 adds into `out` (running skip sum).
 * ------------------------------------------------------------------------------------------ */
int wnb200_taps_fwd(int dtype, int B, int T_out, int M, int nsrc, const wnb200_src_t* srcs /*host*/,
                    const float* bias, int epilogue, int accumulate, void* out, void* th, void* sg,
                    void* stream);

int wnb200_taps_fwd_ex(const wnb200_taps_t* args /*host*/, void* stream);

/* dW[m, c] += sum_{b,t} dout[b, m, t] * pre(x[b, c, t + t_off]) ; dout is contiguous [B, M, T_out].
 * Weight gradient of every conv / linear above (autograd of conv_ops.py:43, block.py:73-78). */
int wnb200_taps_wgrad(int dtype, int B, int T_out, int M, const wnb200_src_t* src /*host, w unused*/,
                      const void* dout, float* dw, void* stream);

/* The same with a wnb200_ln_t for a source with WNB200_PRE_LNRELU (ln may be NULL otherwise). */
int wnb200_taps_wgrad_ex(int dtype, int B, int T_out, int M, const wnb200_src_t* src /*host, w unused*/,
                         const wnb200_ln_t* ln /*host*/, const void* dout, float* dw, void* stream);

/* out[c] += sum_{b,t} a[b,c,t] * (b_or_null ? b[b,c,t] : 1)   (bias / gamma / beta gradients) */
int wnb200_channel_reduce(int dtype, int B, int C, int T, const void* a, const void* b_or_null,
                          float* out, void* stream);

/* Gate backward (block.py:185): d_ab[b, 0:C, t] = dact*sg*(1-th^2); d_ab[b, C:2C, t] = dact*th*sg*(1-sg). */
int wnb200_gate_bwd(int dtype, int B, int C, int T, const void* dact, const void* th, const void* sg,
                    void* d_ab, void* stream);

/* Stand-alone gate (block.py:184-185): out = tanh(a) * sigmoid(b); th / sg (optional) keep the two factors. */
int wnb200_gate_fwd(int dtype, int64_t n, const void* a, const void* b, void* out, void* th, void* sg, void* stream);

/* dx = dy * (ref > 0 ? 1 : 0.01)  (LeakyReLU backward; ref = input or output of the LeakyReLU). */
int wnb200_leaky_bwd(int dtype, int64_t n, const void* dy, const void* ref, void* dx, void* stream);

/* Channel (dim=1) softmax / log-softmax of an NCL tensor, without the reshape_in/reshape_out copies
 * (wavenet.py:108-109, raw_ctcnet.py:152-153, classifier.py:119-120). */
int wnb200_softmax_fwd(int dtype, int B, int C, int T, const void* x, void* y, int log_mode, void* stream);
int wnb200_softmax_bwd(int dtype, int B, int C, int T, const void* y, const void* dy, void* dx,
                       int log_mode, void* stream);

/* nn.AvgPool1d(kernel=pool) (stride=pool, no padding, floor) (classifier.py:53,102). */
int wnb200_avgpool_fwd(int dtype, int B, int C, int T, int pool, const void* x, void* y, void* stream);
int wnb200_avgpool_bwd(int dtype, int B, int C, int T, int pool, const void* dy, void* dx, void* stream);

/* LayerNorm over channels with unbiased std and eps added to the std (layernorm.py:25-28).
 * stats: fp32 [B, T, 2] = (mean, 1/(std+eps)). */
int wnb200_layernorm_fwd(int dtype, int B, int C, int T, const void* x, const float* gamma,
                         const float* beta, float eps, void* y, float* stats, void* stream);
int wnb200_layernorm_bwd(int dtype, int B, int C, int T, const void* x, const float* gamma,
                         const float* stats, float eps, const void* dy, void* dx, void* stream);

/* dgamma[c] += sum_{b,t} dy*(x-mean)*r ; dbeta[c] += sum_{b,t} dy  (stats from wnb200_layernorm_fwd). */
int wnb200_layernorm_bwd_params(int dtype, int B, int C, int T, const void* x, const float* stats, const void* dy,
                                float* dgamma, float* dbeta, void* stream);

/* Fused log-softmax + NLL over channels: replaces the T-iteration Python loop of
 * nn.CrossEntropyLoss in legacy_code/train.py:36-39.  logits NCL [B,C,T], target int64 [B,T].
 * loss_bt[b,t] = logsumexp_c(logits) - logits[b,target,t];  lse[b,t] kept for backward.
 * bwd: dlogits[b,c,t] = (exp(logits - lse) - [c == target]) * (*gscale) . */
int wnb200_xent_fwd(int dtype, int B, int C, int T, const void* logits, const int64_t* target,
                    float* loss_bt, float* lse, void* stream);
int wnb200_xent_bwd(int dtype, int B, int C, int T, const void* logits, const int64_t* target,
                    const float* lse, const float* gscale /*device scalar*/, void* dlogits, void* stream);

/* out[0] = sum(x[0:n]) deterministically (two-stage); scratch >= 1024 floats. */
int wnb200_sum_f32(int64_t n, const float* x, float* out, float* scratch, void* stream);

/* RawCTCNet position mixing (raw_ctcnet.py:131-135): out[b,f,t] += hardtanh(w[f]*(t + t0) + bias[f]). */
int wnb200_positions_add(int dtype, int B, int F, int T, int t0, const float* w, const float* bias,
                         void* out, void* stream);

/* Gradients of the position layer's parameters (autograd of raw_ctcnet.py:131-135; trained by pretrain_tnt.py:121-124):
 * dw[f] += sum_{b,t} g[b,f,t] (t + t0) [|w[f](t+t0)+bias[f]| < 1];  db[f] += sum_{b,t} g[b,f,t] [..]; g NCL [B,F,T]. */
int wnb200_positions_bwd(int dtype, int B, int F, int T, int t0, const float* w, const float* bias, const void* g,
                         float* dw, float* db, void* stream);

/* Per-frame argmax over channels of NCL logits -> int64 [B,T]  (sequence_decoders.py:21-23,
 * legacy_code/train.py:36). */
int wnb200_argmax_channels(int dtype, int B, int C, int T, const void* x, int64_t* out, void* stream);


/* ------------------------------------------------------------------------------------------
 * Tensor-core path (bf16 storage, fp32 accumulation in TMEM, tcgen05.mma + TMA), NLC layout.
 * One launch evaluates, per 128-frame tile, a chain of up to two channel contractions with the
 * intermediate kept in shared memory (never written to HBM):
 *   stage 1: acc1 = sum_j W1[:, j*C:(j+1)*C] x[t + t_off[j]]          (dilated taps, zero padded)
 *   epi 1  : GATE  act = tanh(acc1[0:C]+b1[0:C]) * sigmoid(acc1[C:2C]+b1[C:2C])   (block.py:66-70,185)
 *            LEAKY act = LeakyReLU(acc1+b1)    LINEAR act = acc1+b1
 *   stage 2: acc2 = W2[:, 0:C] act (+ W2[0:C, C:2C] x[t] into rows 0:C when use_x2)
 *   epi 2  : RESBLOCK res = acc2[0:C]+b2[0:C] -> y_nlc (bf16);  skips (+)= acc2[C:2C]+b2[C:2C] (fp32)
 *            HEAD     (softmax over) acc2[0:n_out]+b2 -> out_ncl                  (wavenet.py:103-109)
 * With n2 == 0 the call is a single contraction whose epilogue writes y_nlc [B,T,n1].
 * A whole ResidualBlock + skip bottleneck (block.py:54-82, wavenet.py:100) is one call:
 *   W1 = [Wtanh ; Wsigmoid] (rows) with tap-major columns, W2 = [[Wres, Wproj], [Wbn*Wskip, 0]],
 *   b2 = [bres+bproj ; Wbn*bskip+bbn]  (the skip->bottleneck product is exact: nothing non-linear sits
 *   between conv1x1_skip and the bottleneck).
 * ------------------------------------------------------------------------------------------ */
#define WNB200_TC_GATE 0
#define WNB200_TC_LEAKY 1
#define WNB200_TC_LINEAR 2
#define WNB200_TC_EPI2_RESBLOCK 1
#define WNB200_TC_EPI2_HEAD 2

typedef struct {
  uint32_t struct_size;   /* = sizeof(wnb200_chain_t)                                            */
  int32_t B, T, C;        /* C = contracted channels = width of act; 64, 128 or 256              */
  int32_t ntaps;          /* 1..3                                                               */
  int32_t t_off[3];
  int32_t epi1;           /* WNB200_TC_GATE / LEAKY / LINEAR                                     */
  int32_t n1;             /* rows of w1: 2C for GATE, else C when a stage 2 follows, else <= 2C  */
  int32_t n2;             /* rows of w2 (multiple of 16; 0 = no second stage)                    */
  int32_t use_x2;
  int32_t epi2;           /* WNB200_TC_EPI2_*                                                    */
  int32_t skips_init;     /* 1: skips = ..., 0: skips += ...                                     */
  int32_t out_f32;        /* HEAD output type: 1 fp32, 0 bf16                                    */
  int32_t n_out;          /* HEAD: valid output channels (<= n2)                                 */
  int32_t softmax;        /* HEAD: 1 = channel softmax, 0 = logits                               */
  const void* x;          /* NLC bf16 [B,T,C]                                                    */
  const void* w1;         /* bf16 [n1][ntaps*C]                                                  */
  const float* bias1;     /* [n1]                                                                */
  const void* w2;         /* bf16 [n2][C or 2C]                                                  */
  const float* bias2;     /* [n2]                                                                */
  void* y_nlc;            /* bf16 NLC: res (RESBLOCK; may be NULL = not needed) or stage-1 output */
  float* skips;           /* fp32 NLC [B,T,C] running skip sum (RESBLOCK)                        */
  void* skips_act;        /* optional bf16 NLC: LeakyReLU(skips) for the head (may be NULL)      */
  void* out_ncl;          /* HEAD output, NCL [B, n_out, T]                                      */
  void* dbg;              /* optional int64[8*16] device buffer: per-phase clock64 stamps of CTA 0 */
} wnb200_chain_t;

int wnb200_chain_fwd_tc(const wnb200_chain_t* args /*host*/, void* stream);


/* Fused ResidualBlock + skip bottleneck, pipelined variant for C = 128 / 256 (resblock_tc.cu): TMEM is
 * split in two regions so the gate / output epilogues overlap the next contraction, res is written with
 * TMA stores and the skip sum is accumulated in HBM with TMA reduce-add (the SM never reads it).
 *   w1   bf16 [2C][ntaps*C]: rows = [tanh 0:C/2 ; sigmoid 0:C/2 ; tanh C/2:C ; sigmoid C/2:C], tap-major columns
 *   b1   fp32 [2C] in the same row order
 *   w2   bf16 [2C][2C] = [[Wres, Wproj], [Wbn*Wskip, 0]];   b2 fp32 [2C] = [bres+bproj ; Wbn*bskip+bbn]
 *   res  bf16 NLC [B,T,C] or NULL (last layer: not needed);  skips fp32 NLC [B,T,C].
 * Two ways to get the skip sum (wavenet.py:97-100): `skips` accumulated in HBM by every call (TMA reduce-add; the
 * single-CTA kernel and the round-1 pipeline), or `gate_out` (deferred: see the field) -- the default of the inference
 * AND the training forward (training adds `save_sg`: the gate stack is also what backward reads). */
typedef struct {
  uint32_t struct_size;   /* = sizeof(wnb200_resblock_t) */
  int32_t B, T, C, ntaps;
  int32_t t_off[3];
  int32_t skips_init;     /* 1: skips = contribution, 0: skips += contribution */
  int32_t variant;        /* 0 = default (CTA-pair kernel, cta_group::2), 1 = single-CTA kernel */
  int32_t act_fmt;        /* WNB200_ACT_BF16 | WNB200_ACT_F16X2 (CTA-pair kernel, inference): format of x, w1, w2, res,
                             skips_act.  F16X2: bias1 is PRE-SCALED (tanh rows by -2 log2 e, sigmoid rows by -log2 e) */
  const void* x;          /* NLC [B,T,C], 2-byte elements of act_fmt (F16X2: the hi half of the stream) */
  const void* w1;
  const float* bias1;
  const void* w2;
  const float* bias2;
  void* res;
  float* skips;
  void* dbg;              /* optional int64[8*16] timeline buffer */
  void* save_act;         /* optional (training): the gate tanh(.)*sigmoid(.), NLC bf16 [B,T,C] ...                 */
  void* save_th;          /* ... optionally tanh(.) (NULL: not kept -- backward derives it as gate / sigmoid) ...      */
  void* save_sg;          /* ... and sigmoid(.) (required with save_act; with gate_out it is the only save_* allowed:
                             the stored gate is the saved activation)                                                 */
  void* skips_act;        /* optional (last layer of an inference stack, res = NULL): NLC bf16 [B,T,C] that receives
                             LeakyReLU(skips + contribution) (wavenet.py:100-103); `skips` is then read, not updated */
  const void* x_lo;       /* F16X2: lo half of the input stream, NLC fp16 [B,T,C]; NULL = the input is exactly x */
  void* res_lo;           /* F16X2: lo half of the output stream; NULL with gate_out: the output stream carries its hi half
                             only (the last blocks of a deep stack, whose rounding is hardly amplified any more) */
  void* gate_out;         /* optional: NLC [B,T,C] in act_fmt that receives the gate tanh(.)*sigmoid(.).  When
                             given, the call does NOT touch `skips` (may be NULL): the skip sum of the whole stack,
                             sum_l (Wbn_l Wskip_l) gate_l + biases, is left to ONE wnb200_dense_fwd_tc call with
                             `nlayers` over the stacked gates (K = layers x channels, accumulated in tensor memory).
                             res = NULL on the last layer skips the residual contraction altogether. */
  int32_t* sat_flag;      /* optional (F16X2 with gate_out): device int that is OR-ed with 1 when a value of the output
                             stream reached the end of the fp16 range (|v| >= 65504: stored saturated, not inf).  The
                             host falls back to WNB200_ACT_BF16 for such a model (fastpath.py). */
} wnb200_resblock_t;
int wnb200_resblock_fwd_tc(const wnb200_resblock_t* args /*host*/, void* stream);


/* Dense channel contraction on tensor cores (CTA pair), NLC bf16 in:
 *   y[b,t,:] = epi(W [x[b,t+off_0,:]; x[b,t+off_1,:]; ...] + bias),   W bf16 [N][ntaps*Cin] (tap-major columns)
 * mode 0: y = bf16 NLC [B,T,N] (N %% 64 == 0), optional LeakyReLU  -- entry conv (wavenet.py:54,93), first 1x1 of
 *         the heads (wavenet.py:68-69), RawCTCNet feature 1x1 (raw_ctcnet.py:60-61)
 * mode 1: y = NCL [B,n_out,T] (bf16 or fp32), optional channel softmax -- last 1x1 of the heads + softmax
 *         (wavenet.py:70-71,103-109).  N = n_out rounded up to 16, padded rows of W / bias are zero. */
typedef struct {
  uint32_t struct_size;   /* = sizeof(wnb200_dense_t) */
  int32_t B, T, Cin, ntaps;
  int32_t t_off[3];
  int32_t N;
  int32_t mode, leaky, n_out, softmax, out_f32;
  int32_t Cin2, ntaps2;   /* optional second source x2 (NLC bf16 [B,T,Cin2]): its taps follow x's in W's columns */
  int32_t t_off2[3];
  const void* x;
  const void* w;
  const float* bias;
  void* y;
  const void* x2;         /* NULL = single source */
  float* colsum;          /* mode 0, optional: fp32 [N], += the column sums of y over all frames (the bias gradient of
                             the contraction that consumes y as its output gradient: autograd of conv biases) */
  int32_t act_fmt;        /* WNB200_ACT_BF16 | WNB200_ACT_F16X2: format of x, x2, w and of the NLC output */
  int32_t nlayers;        /* > 0: x is a stack [nlayers][B][T][Cin] (ntaps = 1, offset 0, no x2) and W is [N][nlayers*Cin]:
                             y = epi(sum_l W_l x_l + bias) -- the skip sum of a whole residual stack in ONE contraction,
                             accumulated in tensor memory over K = layers x channels (wavenet.py:97-100) */
  void* y_lo;             /* F16X2, mode 0, optional: y leaves as an fp16 (hi, lo) pair (y, y_lo) -- the producer of a
                             residual stream (entry conv, RawCTCNet feature 1x1) */
  const void* gb_gate;    /* BF16, mode 0, optional (training): gate-backward epilogue.  The contraction's result is
                             d(gate) of a residual block (block.py:66-71 backwards); with the forward's stored gate =
                             tanh*sigmoid and ... */
  const void* gb_sg;      /* ... sigmoid (both NLC bf16 [B,T,N], 16-byte aligned), y is NLC bf16 [B,T,2N]:
                             y[..,0:N] = d * sg * (1 - th^2), y[..,N:2N] = d * th * sg * (1 - sg), th = gate / sg;
                             colsum is then fp32 [2N].  Replaces wnb200_gate_bwd_nlc_from_gate + one HBM round trip. */
  const float* pos_w;     /* mode 0, optional: after bias / LeakyReLU, y[b,t,c] += hardtanh(pos_w[c] * (pos_t0 + t) + pos_b[c])
                             -- RawCTCNet(positions=True) position mixing (raw_ctcnet.py:131-135) fused into the feature
                             layer's 1x1; fp32 [N] each */
  const float* pos_b;
  int32_t pos_t0;         /* global frame index of t = 0 (a time shard of a longer read) */
} wnb200_dense_t;
int wnb200_dense_fwd_tc(const wnb200_dense_t* args /*host*/, void* stream);


/* ------------------------------------------------------------------------------------------
 * Weight packing (pack.cu): the operand matrices above, built on the device straight from the reference's parameter
 * tensors (state_dict layout: conv weights [out, in, k], nn.Linear weight [out, in]) -- one launch per block / head.
 * w_dtype = dtype of ALL source parameters (WNB200_F32 master weights or WNB200_BF16).
 *   w1  [2C][k*C]  act_fmt words, tap-major columns; row_order 1 = [tanh 0:C/2; sig 0:C/2; tanh C/2:C; sig C/2:C]
 *                  (wnb200_resblock_fwd_tc), 0 = [tanh 0:C; sig 0:C] (wnb200_chain_fwd_tc)
 *   b1  [2C] fp32 in w1's row order; for WNB200_ACT_F16X2 pre-scaled (tanh rows x -2 log2 e, sigmoid rows x -log2 e)
 *   w2  [2C][2C] = [[Wres, Wproj], [Wbn*Wskip, 0]];   b2 [2C] = [bres + bproj ; Wbn*bskip + bbn]
 *   backward (optional, all four or none, bf16):  wdg [C][2C] = [Wres^T | (Wbn Wskip)^T],  wdg_skip [C][C],
 *   wdx [C][k*2C + C] = [Wt_0^T | Ws_0^T | ... | Wproj^T],  wdx_taps [C][k*2C] (autograd of block.py:66-79).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  uint32_t struct_size;   /* = sizeof(wnb200_pack_block_t) */
  int32_t C, k;           /* channels (in = out = bottleneck width), kernel width 1..3 */
  int32_t act_fmt;        /* WNB200_ACT_* of w1 / w2 */
  int32_t w_dtype;        /* WNB200_F32 | WNB200_BF16 */
  int32_t row_order;
  const void *wt, *bt, *ws, *bs;      /* conv_tanh / conv_sigmoid .conv1d.weight [C,C,k], .bias [C]  (block.py:35-44) */
  const void *wres, *bres;            /* conv1x1_residual (block.py:45)                                                */
  const void *wskip, *bskip;          /* conv1x1_skip (block.py:46)                                                    */
  const void *wproj, *bproj;          /* residual_proj, nn.Linear (block.py:48)                                        */
  const void *wbn, *bbn;              /* the block's skip bottleneck (wavenet.py:60-63)                                */
  void* w1; float* b1; void* w2; float* b2;
  void *wdg, *wdg_skip, *wdx, *wdx_taps;
} wnb200_pack_block_t;
int wnb200_pack_block(const wnb200_pack_block_t* args /*host*/, void* stream);

/* Output head (LeakyReLU -> 1x1 -> LeakyReLU -> 1x1; wavenet.py:67-71): pw1 [C][C], pb1 [C], pw2 [n2][C] and pb2 [n2]
 * with n2 = n_out rounded up to 16 (padded rows zero); backward (optional, bf16): w3t [C][npad] = W3^T with npad = n_out
 * rounded up to 64, w1t [C][C] = W1^T. */
typedef struct {
  uint32_t struct_size;   /* = sizeof(wnb200_pack_head_t) */
  int32_t C, n_out, act_fmt, w_dtype;
  int32_t reserved0;
  const void *w1, *b1, *w3, *b3;      /* output_stack.1 / .3 weight [C,C,1] / [n_out,C,1] and bias */
  void* pw1; float* pb1; void* pw2; float* pb2;
  void *w3t, *w1t;
} wnb200_pack_head_t;
int wnb200_pack_head(const wnb200_pack_head_t* args /*host*/, void* stream);

/* Gradients of the two factors of the folded skip -> bottleneck product, from weight space, all L <= 64 layers in one
 * launch: with M[l] = dskips (x) gate_l (fp32 [L][C][C], written by wnb200_wgrad2_tc) and csk = column sums of dskips,
 *   dwskip[l] = Wbn_l^T M[l]     dwbn[l] = M[l] Wskip_l^T + csk (x) bskip_l     dbskip[l] = Wbn_l^T csk
 * wbn / wskip / bskip: HOST arrays of L device pointers to the parameters (w_dtype). */
int wnb200_fold_grads(int w_dtype, int L, int C, const void* const* wbn /*host*/, const void* const* wskip /*host*/,
                      const void* const* bskip /*host*/, const float* M, const float* csk, float* dwskip, float* dwbn,
                      float* dbskip, void* stream);

/* Bytes of device memory one forward of a C-channel residual stack over B x T frames needs besides its input and
 * output (stream ping-pong -- twice that for the fp16 (hi, lo) format --, fp32 skip sum, the head's two activations),
 * for hosts that allocate the buffers themselves. */
size_t wnb200_workspace_bytes(int B, int T, int C, int act_fmt);

/* RawCTCNet featuriser, first layer: Conv1d(1, F, fk, padding=fk-1) + LeakyReLU (raw_ctcnet.py:57-59) on the raw
 * signal x [B, 1, T] -> y NLC [B, T+fk-1, F] in act_fmt (bf16 / fp16).  w fp32 [F, fk], bias fp32 [F]. */
int wnb200_featurize_nlc(int dtype, int B, int T, int F, int fk, const void* x, const float* w, const float* bias,
                         int act_fmt, void* y, void* stream);

/* Backward of the above: dh NLC bf16 [B, floor(T/pool), C] -> dx NCL [B, C, T] (bf16 or fp32),
 * dx[b,c,t] = dh[b, t/pool, c] / pool for t < floor(T/pool)*pool, 0 after (autograd of nn.AvgPool1d). */
int wnb200_avgpool_bwd_nlc_to_ncl(int dtype_out, int B, int C, int T, int pool, const void* dh, void* dx, void* stream);

/* WaveNet entry conv (wavenet.py:54,93) on quantised LEVELS instead of their one-hot encoding (fns.py:6-15,
 * pore_model.py:88-96): y[b,t,:] = bias + sum_j wemb[j][levels[b,t+t_off[j]]][:], taps outside [0,T) contribute nothing.
 * levels int32 [B,T] (clamped to [0,in_dim)), wemb bf16 [ntaps][in_dim][C] (= conv weight [C,in_dim,ntaps] permuted),
 * bias fp32 [C], y NLC bf16 [B,T,C]; t_off: host int32[ntaps].  Bit-identical to the dense kernel on the one-hot input.
 * act_fmt = WNB200_ACT_F16X2: wemb is fp16 and y (y_lo, optional) receive the fp16 (hi, lo) pair. */
int wnb200_entry_embed_nlc(int B, int T, int C, int in_dim, int ntaps, const int32_t* t_off /*host*/,
                           const int32_t* levels, const void* wemb, const float* bias, int act_fmt, void* y, void* y_lo,
                           void* stream);

/* AvgPool1d(pool) (classifier.py:53,102) fused with the NCL -> NLC layout change:
 * x NCL [B, C, T] -> y NLC [B, floor(T/pool), C] in act_fmt (bf16 / fp16). */
int wnb200_avgpool_ncl_to_nlc(int dtype, int B, int C, int T, int pool, const void* x, int act_fmt, void* y,
                              void* stream);


/* Weight gradient on tensor cores (CTA pair, MN-major operands):
 *   dw[m, n] += sum_{b,t} g[b, t, m0 + m] * x[b, t + off, n],   m < 256, n < N (N = 128 or 256)
 * g NLC bf16 [B,T,Cg], x NLC bf16 [B,T,N], dw fp32 [256][N] (accumulated into; zero it first).  Frames outside
 * [0,T) read as zero.  Weight gradient of every contraction of the residual stack on the bf16 training step. */
int wnb200_wgrad_tc(int B, int T, int Cg, int m0, int N, int off, const void* g_nlc, const void* x_nlc, float* dw,
                    void* stream);
/* Same with one or two X tensors sharing the G tiles (nsrc = 1 | 2, off: host int32[nsrc]):
 *   dw[m, s*N + n] += sum_{b,t} g[b, t, m0+m] * x_s[b, t+off_s, n]        dw fp32 [256, nsrc*N]
 * e.g. both taps of a dilated conv (x at two offsets) or [gate | x] for conv1x1_residual and residual_proj. */
int wnb200_wgrad2_tc(int B, int T, int Cg, int m0, int N, int nsrc, const int32_t* off /*host*/, const void* g_nlc,
                     const void* x_nlc, const void* x2_nlc, float* dw, void* stream);
/* Up to 6 such products over the same frames in ONE launch -- all weight gradients of a residual block (block.py:60-78:
 * [dtanh ; dsigmoid] x the taps of x, dres x [gate | x], dskips x gate).  The CTA pairs are divided over the jobs by MMA
 * work and all jobs sweep the frames at the same pace, so x, the gate and the gradients come from HBM once per launch and
 * from L2 for the other jobs.  jobs: host array. */
typedef struct {
  const void* g;          /* NLC bf16 [B,T,Cg] */
  int32_t Cg, m0;
  const void* x;          /* NLC bf16 [B,T,N] */
  const void* x2;         /* second source (nsrc = 2) or NULL */
  int32_t N, nsrc;
  int32_t off[2];
  float* dw;              /* fp32 [256][nsrc*N], accumulated into */
} wnb200_wgrad_job_t;
int wnb200_wgrad_jobs_tc(int B, int T, int njobs, const wnb200_wgrad_job_t* jobs /*host*/, void* stream);

/* Gate backward on NLC bf16 tensors (block.py:185): dab[r, 0:C] = dact*sg*(1-th^2), dab[r, C:2C] = dact*th*sg*(1-sg)
 * for each of `rows` = B*T frames; if dbias != NULL, dbias[0:2C] (fp32) += the column sums of dab (the two conv
 * biases' gradients). */
int wnb200_gate_bwd_nlc(int64_t rows, int C, const void* dact, const void* th, const void* sg, void* dab, float* dbias,
                        void* stream);

/* Same from what the training forward keeps (gate = tanh * sigmoid, and sigmoid): tanh = gate / sigmoid. */
int wnb200_gate_bwd_nlc_from_gate(int64_t rows, int C, const void* dact, const void* gate, const void* sg, void* dab,
                                  float* dbias, void* stream);

/* out[c] += sum over rows of x[r, c]  (x NLC bf16 [rows, C]; bias gradients on the tensor-core training path). */
int wnb200_colsum_nlc(int64_t rows, int C, const void* x, float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * CTC loss (replaces warpctc_pytorch.CTCLoss at legacy_code/train.py:42-46, run_raw_ctc.py:59-62, Loss.py:50-53;
 * third party in the reference).  Activations are pre-softmax, class 0 is the blank, `act` is addressed by element
 * strides (sb, sc, st) for (read, class, frame) so both the reference's (T,B,C) layout and the classifier's (B,C,T)
 * output are read in place.  labels: concatenated int32 labels; label_offsets: int64[B+1] prefix sums;
 * act_lengths: int32[B] or NULL (= T for every read).  workspace: wnb200_ctc_workspace_bytes() bytes, shared by
 * fwd and bwd.  fwd writes nll[b] = -log p(labels_b | act_b) (+inf if no alignment exists);
 * bwd writes grad = gscale[0] * d(sum_b nll[b]) / d act with act's strides (zero for infeasible reads). */
size_t wnb200_ctc_workspace_bytes(int B, int L, int T, int max_label_len);
int wnb200_ctc_fwd(int dtype, int B, int L, int T, int max_label_len, const void* act, int64_t sb, int64_t sc,
                   int64_t st, const int32_t* labels, const int64_t* label_offsets, const int32_t* act_lengths,
                   float* workspace, float* nll, void* stream);
int wnb200_ctc_bwd(int dtype, int B, int L, int T, int max_label_len, const int32_t* labels,
                   const int64_t* label_offsets, const int32_t* act_lengths, const float* workspace, const float* nll,
                   const float* gscale, void* grad, int64_t sb, int64_t sc, int64_t st, void* stream);

/* ------------------------------------------------------------------------------------------
 * Greedy decoding (modules/sequence_decoders.py:9-23 argmax_decode; collapse-repeats + drop-blank as done by hand in
 * ipynbs/Size 1 Pore Model Check.ipynb cell 24).  `act` is addressed by element strides (read, class, frame).
 * frame_argmax: out[b, t] = argmax_c act[b, c, t] (int64, ties -> lowest class).
 * ctc_greedy_decode: out_labels[b, 0:out_lengths[b]] = collapsed, blank-free label sequence of read b (int32 [B, T]
 * buffer, entries past the length are left untouched); act_lengths int32[B] or NULL (= T). */
int wnb200_frame_argmax(int dtype, int B, int L, int T, const void* act, int64_t sb, int64_t sc, int64_t st,
                        int64_t* out, void* stream);
int wnb200_ctc_greedy_decode(int dtype, int B, int L, int T, const void* act, int64_t sb, int64_t sc, int64_t st,
                             const int32_t* act_lengths, int blank, int32_t* out_labels, int32_t* out_lengths,
                             void* stream);

/* Backward of the RawCTCNet featuriser's first layer (raw_ctcnet.py:57-59) on the tensor-core training path, one pass:
 * df, fact: NLC bf16 [B, T+fk-1, F] (gradient w.r.t. the LeakyReLU output, and that output); seq [B, T] (seq_dtype);
 * w fp32 [F, fk].  Accumulates dw fp32 [F, fk], db fp32 [F] and (if not NULL) dseq fp32 [B, T]; zero them first. */
int wnb200_featurize_bwd_nlc(int seq_dtype, int B, int T, int F, int fk, const void* df, const void* fact,
                             const void* seq, const float* w, float* dw, float* db, float* dseq, void* stream);

/* MultiplicativeUnit gate (block.py:213-220) on NCL tensors: pre [B, 4C, T] = outputs of (gate1; gate2; gate3;
 * update) stacked on the channel axis, h [B, C, T]:  out = sig(pre1) * tanh(sig(pre2) * h + sig(pre3) * tanh(pre4)).
 * bwd: dpre [B, 4C, T] and the direct term of dh [B, C, T] from dout (the convolutions' share of dh is added by
 * their own backward). */
int wnb200_mu_gate_fwd(int dtype, int B, int C, int T, const void* pre, const void* h, void* out, void* stream);
int wnb200_mu_gate_bwd(int dtype, int B, int C, int T, const void* pre, const void* h, const void* dout, void* dpre,
                       void* dh, void* stream);

/* ------------------------------------------------------------------------------------------
 * On-device synthetic pore-model signal (utils/raw_signal_generator.py:77-118,189-203; utils/pore_model.py:58-96).
 * siggen_raw: per read, `nbases` bases ~ U{1..4} -> 5-mer ids -> samples per k-mer max(1, int(Gamma(2.461964,
 * 1/587.2858) * 800)) -> T picoamp samples mean[k] + stdv[k] * z.  Outputs the draws too: bases int32 [B, nbases],
 * reps int32 [B, nbases-4], z fp32 [B, T] (optional), n_used int32 [B] = bases covered by the T samples (the CTC
 * labels are bases[2 : 2 + n_used]; -1 if nbases was too small: retry with more), sig fp32 [B, T].
 * siggen_onehot: per read mu-law quantisation to `levels` and one-hot NCL [B, levels, T]; lev_out int64 [B, T] optional. */
int wnb200_siggen_raw(int B, int T, int nbases, uint64_t seed, const float* means, const float* stdvs, int32_t* bases,
                      int32_t* reps, int32_t* n_used, float* z, float* sig, void* stream);
int wnb200_siggen_onehot(int dtype, int B, int T, int levels, const float* sig, void* onehot, int64_t* lev_out,
                         void* stream);

/* y = bf16(LeakyReLU_0.01(x)), n a multiple of 4: turns the fp32 skip sum into the head's input
 * (first LeakyReLU of output_stack, wavenet.py:67). */
int wnb200_leaky_to_bf16(int64_t n, const float* x, void* y, void* stream);

/* NCL (fp32/bf16) -> NLC bf16, and NLC (fp32/bf16) -> NCL (fp32/bf16): layout change at the module
 * boundary only (the reference's reshape_in/reshape_out, conv_ops.py:91-101, ran once per block). */
int wnb200_ncl_to_nlc_bf16(int dtype, int B, int C, int T, const void* x, void* y, void* stream);
/* Same for x[b,c,t] at x + b*sb + c*sc + t (element strides): a time slice of a longer tensor is read in place
 * (legacy_code/train.py:30 feeds sig[:, :, 0:-1]). */
int wnb200_ncl_to_nlc_bf16_strided(int dtype, int B, int C, int T, int64_t sb, int64_t sc, const void* x, void* y,
                                   void* stream);
/* Same with the 2-byte output format chosen by act_fmt (bf16 / fp16). */
int wnb200_ncl_to_nlc_act(int dtype, int act_fmt, int B, int C, int T, int64_t sb, int64_t sc, const void* x, void* y,
                          void* stream);
int wnb200_nlc_to_ncl(int out_dtype, int src_is_f32, int B, int C, int T, const void* x, void* y, void* stream);

/* ------------------------------------------------------------------------------------------
 * ByteNet residual blocks (block.py:86-173) and the frame-at-a-time LinearConv1d (linear_conv_ops.py:39-68).
 * ln_stats:    stats[b, t] = (mean_c x[b, c, t], 1 / (unbiased std_c + eps))   (layernorm.py:25-28); x contiguous [B, C, T].
 *              A contraction reads the normalised, rectified tensor through WNB200_PRE_LNRELU without it ever being stored.
 * ln_relu_fwd: y = ReLU(gamma * (x - mean) * r + beta) stored (training keeps it for the MultiplicativeUnit's backward).
 * ln_relu_bwd: dy = gradient w.r.t. that y -> dx (may be NULL), dgamma / dbeta += (fp32 [C], may be NULL).
 * ------------------------------------------------------------------------------------------ */
int wnb200_ln_stats(int dtype, int B, int C, int T, const void* x, float eps, float* stats, void* stream);
int wnb200_ln_relu_fwd(int dtype, int B, int C, int T, const void* x, const float* stats, const float* gamma,
                       const float* beta, void* y, void* stream);
int wnb200_ln_relu_bwd(int dtype, int B, int C, int T, const void* x, const float* stats, const float* gamma,
                       const float* beta, float eps, const void* dy, void* dx, float* dgamma, float* dbeta, void* stream);

/* LinearConv1d.linear (linear_conv_ops.py:39-68): y[n, m] = bias[m] + sum_{c, j} w[m, c, j] * frame[n, c, j * dilation],
 * frame [N, Cin, rf] with rf = k + (dilation - 1) * (k - 1), w [Cout, Cin, k], y [N, Cout]; k <= 32. */
int wnb200_linear_frame(int dtype, int N, int Cin, int Cout, int k, int dilation, const void* w, const float* bias,
                        const void* frame, void* y, void* stream);
/* The same frame evaluated incrementally: the caller hands over frame number `step` (x [N, Cin]) and a history ring
 * hist [rf, N, Cin] (zero before step 0, owned by the caller, rf = (k - 1) * dilation + 1); the call computes
 * y[n, :] from x and the k - 1 earlier frames the kernel's taps address, and files x in the ring.  A decoder that
 * emits one frame per step (bytenet_decoder.py:146-186 re-evaluates its whole receptive field per step) does O(k)
 * instead of O(rf) work per layer and step. */
int wnb200_linear_step(int dtype, int N, int Cin, int Cout, int k, int dilation, int64_t step, const void* w,
                       const float* bias, const void* x, void* hist, void* y, void* stream);

/* Tensor-core form of the ByteNet blocks (bf16 inference; contractions = wnb200_dense_fwd_tc on NLC activations):
 * lnrelu_rows:          y[r, :] = ReLU(LayerNorm(x[r, :])) for NLC rows (r = frame), bf16 in / out, C % 8 == 0, C <= 1024.
 * mu_gate_rows:         out = g1 * tanh(g2 * h + g3 * tanh(u)) with the pre-activation of unit u (gate1, gate2, gate3,
 *                       update) at pre_u + r * pre_pitch + c (four tensors or four column blocks of one).
 * nlc_parts_to_ncl_add: out[b, p * Cp + c, t] = residual[b, p * Cp + c, t] + parts[p][b, t, c]: back to NCL with the block's
 *                       `seq +` (block.py:119,166); parts = host array of nparts <= 4 device pointers. */
int wnb200_lnrelu_rows(int64_t rows, int C, const void* x, const float* gamma, const float* beta, float eps, void* y,
                       void* stream);
int wnb200_mu_gate_rows(int64_t rows, int C, const void* pre0, const void* pre1, const void* pre2, const void* pre3,
                        int64_t pre_pitch, const void* h, void* out, void* stream);
int wnb200_nlc_parts_to_ncl_add(int B, int C, int T, int nparts, int Cp, const void* const* parts /*host*/,
                                const void* residual, void* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Adam over every parameter tensor in ONE launch (legacy_code/train.py:55 `opt.step()`; torch.optim.Adam semantics
 * without amsgrad: g += wd * p; m = b1 m + (1 - b1) g; v = b2 v + (1 - b2) g^2;
 * p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)), fp32 state, fp32 or bf16 parameters / gradients.
 * items (DEVICE array): one entry per tensor; chunks (DEVICE array of nchunks (item, first element) int32 pairs):
 * tensor i contributes ceil(numel_i / wnb200_adam_chunk_elems()) consecutive chunks.  step counts from 1.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  void* param;       /* fp32 or bf16 [numel] */
  const void* grad;  /* same dtype as param */
  float* exp_avg;    /* fp32 [numel] */
  float* exp_avg_sq; /* fp32 [numel] */
  int64_t numel;
  int32_t is_bf16;
  int32_t reserved0;
} wnb200_adam_item_t;
int wnb200_adam_chunk_elems(void);
int wnb200_adam_step(int nchunks, const wnb200_adam_item_t* items /*device*/, const int32_t* chunks /*device*/, float lr,
                     float beta1, float beta2, float eps, float weight_decay, int64_t step, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WNB200_H_ */
