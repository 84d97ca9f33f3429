/* oniris_b200.h -- C ABI of the B200-native Oniris denoiser hot path.
 *
 * The reference (Francesco215/autoregressive_diffusion) has no FFI: its "operator interface" for this path is
 * the Python module API of edm2/conv.py, edm2/attention/ and edm2/utils.py, whose leaves are PyTorch library
 * calls.  Each entry point below replaces one such leaf (or one fused group of leaves); the reference call
 * site it stands in for is cited on every declaration (paths relative to the reference root).
 *
 * Conventions
 *  - Every pointer is a DEVICE pointer unless stated otherwise.  The library never allocates, frees or keeps
 *    caller memory; outputs and workspaces are caller-allocated.
 *  - Activations are bf16, NHWC: [frames, H, W, C] with C contiguous ("rows x channels").  Frames are ordered
 *    (b, s, t): batch-major, then the clean (s=0) / noised (s=1) half of the DART training sequence, then time.
 *    S=1 means a plain (eval) sequence.
 *  - Conv weights travel as ONE bf16 operand matrix wg[Cout][taps][Cin] produced by ob_wnorm_fwd
 *    (taps = k*k for plain convs; 27 = 9 current-frame + 2x9 causal taps for the gated conv).
 *  - All work is enqueued on `stream` (a cudaStream_t passed as void*); no entry point synchronises.
 *  - Return value: 0 = ok, <0 = error (OB_ERR_*); ob_last_error() returns a thread-local message.
 *  - Channel counts must be multiples of 8 (16 for conv inputs); the Python layer pads where the model's are not.
 */
#ifndef ONIRIS_B200_H
#define ONIRIS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OB_OK 0
#define OB_ERR_INVALID (-1)
#define OB_ERR_UNSUPPORTED (-2)
#define OB_ERR_CUDA (-3)

int ob_version(void);
const char* ob_last_error(void);

/* ---------------------------------------------------------------------------------------------- weights
 * edm2/conv.py:14-21 NormalizedWeight.forward (+ edm2/utils.py:83-88 normalize).
 * w: fp32 [Cout][taps][Cin] -- the reference's [Cout, Cin, (kt,) kh, kw] parameter stored TAP-MAJOR (torch channels_last /
 * channels_last_3d memory format; same logical shape and state_dict, so reference checkpoints load unchanged).  This is
 * the order of the GEMM operand and of the weight-gradient partials, so both directions are straight row streams.
 * Writes the bf16 GEMM operand
 * wg[Cout][taps_total][cin_pad] at tap offset tap_off (channels >= Cin zero-filled).  training != 0 also
 * overwrites w in place with its forced-normalised value and normalises THAT for the operand, exactly as the
 * reference's in-place copy_ + second normalize does. */
int ob_wnorm_fwd(float* w, void* wg, int cout, int cin, int taps, int cin_pad, int taps_total, int tap_off, float gain,
                 float eps, int training, void* stream);

/* The same for MANY weight tensors in one launch (an optimizer step invalidates every operand of a network at once).
 * jobs: DEVICE array of n_jobs descriptors; row_start: DEVICE int32 [n_jobs + 1], row_start[j] = sum of cout of the jobs
 * before j (row_start[n_jobs] = total_rows).  Field meaning as the arguments of ob_wnorm_fwd. */
typedef struct ob_wnorm_job {
  float* w;
  void* wg;
  int32_t cin, taps, cin_pad, taps_total, tap_off, training;
  float gain;
  int32_t pad_;
} ob_wnorm_job;
int ob_wnorm_fwd_multi(const ob_wnorm_job* jobs, const int* row_start, int n_jobs, int total_rows, float eps, void* stream);

/* Backward of the above w.r.t. the (forced) weights: dwg fp32 [n_split][Cout][taps_total][cin_pad] are the
 * split-K partial sums written by ob_conv_wgrad; dw fp32 [Cout][taps][Cin] (same storage order as w) is overwritten, or
 * (accumulate & 1) added to -- gradient accumulation over micro-batches (cs_train.py:108-109) without a separate pass.
 * accumulate & 2: dwg is the running sum kept by the accumulating weight-gradient entry point below, n_split = 1: it is
 * cleared as it is consumed (the kernel writes through the const pointer), ready for the next accumulation cycle.
 * ob_wnorm_bwd_gated does the 2D (9 taps at offset 0) and 3D (18 taps at offset 9) weights of a gated conv in one launch. */
int ob_wnorm_bwd(const float* w, const float* dwg, float* dw, int cout, int cin, int taps, int cin_pad, int taps_total,
                 int tap_off, int n_split, float gain, float eps, int accumulate, void* stream);
int ob_wnorm_bwd_gated(const float* w2, float* dw2, const float* w3, float* dw3, const float* dwg, int cout, int cin,
                       int cin_pad, int n_split, float eps, int accumulate, void* stream);

/* ---------------------------------------------------------------------------------------------- convolutions
 * edm2/conv.py:36-42 (MPConv: F.conv2d k=1|3) and :59-95 (MPCausal3DGatedConv: F.conv2d + F.conv3d + mp_sum).
 *   gated == 0:  out[f] = conv_kxk(x[f], wg)                                   x: [n_seq*S*T frames, H, W, Cin]
 *   gated != 0:  out[b,s,t] = alpha[b,s,t] * conv3x3(x[b,s,t], wg[:, 0:9])
 *                           + beta[b,s,t]  * sum_{tau in {0,1}} conv3x3(ctx[b, t+tau], wg[:, 9+9*tau : 18+9*tau])
 *                ctx: bf16 [n_seq, T+2, H, W, Cin] = two pad frames (ones, or the cached activations) followed
 *                by the T clean frames; the context term is computed ONCE per (b,t) and shared by both halves.
 *                out_d (optional, may be NULL): fp16, context term minus current-frame term (saved for backward; its inner
 *                product with dy feeds the gate scalars' gradients: fp16 keeps 11 mantissa bits at bf16's two bytes).
 * out: [n_seq*S*T, H, W, Cout], bf16 or (out_f32 != 0) fp32.  alpha/beta: fp32 [n_seq*S*T].
 * bias (plain convs; may be NULL): fp32 [cout] added in the epilogue before the rounding to bf16 (nn.Conv3d(bias=True) of the VAE).
 * w_taps (plain convs; 0 = ksize*ksize): taps per row of wg -- 27 lets the 2-D form of a gated conv (just_2d, edm2/conv.py:60)
 * read the first 9 taps of the SAME operand matrix the gated form uses (one operand, one normalisation per weight).
 * split_ws (optional, may be NULL): fp32 scratch sized by ob_conv_split_ws_bytes.  When given and the layer has too
 * few output tiles for 148 SMs (the 4x4 / 8x8 levels), the channel chunks are sliced over extra CTAs and reduced there
 * (split-K); the call zeroes it itself.  For the input gradient pass query with cin/cout SWAPPED (it is the transposed
 * problem). */
int64_t ob_conv_split_ws_bytes(int n_seq, int S, int T, int H, int W, int cin, int cout, int ksize, int gated);
int ob_conv_fwd(const void* x, const void* ctx, const void* wg, const float* alpha, const float* beta, void* out,
                void* out_d, void* split_ws, int n_seq, int S, int T, int H, int W, int cin, int cout, int ksize, int gated,
                int out_f32, int w_taps, const float* bias, void* stream);

/* ob_conv_fwd with a fused post-op behind the output stage (the "E1" / "E4" epilogues of edm2/networks_edm2.py:75-77 and :86,93;
 * the reference runs them as separate eager kernels).  A SECOND bf16 output out2 (same shape as out) is computed from the fp32
 * result y before it is rounded:
 *   OB_POST_SCALE_SILU: out2 = mp_silu(y * cscale[frame][channel])   cscale fp32, row stride cscale_ld (0 = cout)
 *   OB_POST_MP_SUM:     out2 = clip(mp_sum(res, y, t))               res bf16 like out; clip <= 0: none
 * out (the raw y) may be NULL when nothing needs it (evaluation); in training it is still written (the backward pass reads
 * it).  Works for plain and gated convs, with and without split-K. */
#define OB_POST_SCALE_SILU 1
#define OB_POST_MP_SUM 2
int ob_conv_fwd_fused(const void* x, const void* ctx, const void* wg, const float* alpha, const float* beta, void* out,
                      void* out_d, void* split_ws, int n_seq, int S, int T, int H, int W, int cin, int cout, int ksize, int gated,
                      int w_taps, int post, void* out2, const float* cscale, int cscale_ld, const void* res, float t, float clip,
                      void* stream);

/* Input gradient of ob_conv_fwd.  gy: bf16 [n_seq*S*T, H, W, Cout] = dL/dout (unscaled);
 * gated: gb: bf16 [n_seq*T, H, W, Cout] = sum_s beta_s*dL/dout_s (from ob_gate_bwd) and
 *        dx[b,s,t] = alpha[b,s,t] * convT3x3(gy[b,s,t]) + beta[b,s,t] * causal_convT(gb[b, t+1..t+2]),
 *        where the caller passes beta = 1 on clean rows and 0 on noised rows (the context is built from clean rows).
 * plain: dx = convT(gy); alpha/beta ignored (may be NULL).
 * dx: bf16 [n_seq*S*T, H, W, Cin].  Reads the SAME wg as the forward pass (as an MN-major operand). */
int ob_conv_dgrad(const void* gy, const void* gb, const void* wg, const float* alpha, const float* beta, void* dx,
                  void* split_ws, int n_seq, int S, int T, int H, int W, int cin, int cout, int ksize, int gated,
                  int w_taps, void* stream);

/* Weight gradient of ob_conv_fwd into dwg fp32 [n_split][Cout][taps][Cin] (taps = k*k or 27).  n_split must come
 * from ob_conv_wgrad_splits for the same shape. */
int ob_conv_wgrad_splits(int n_seq, int S, int T, int H, int W, int cin, int cout, int ksize, int gated);
int ob_conv_wgrad(const void* gya, const void* x, const void* gb, const void* ctx, float* dwg, int n_seq, int S, int T,
                  int H, int W, int cin, int cout, int ksize, int gated, int n_split, void* stream);
/* Same, ADDED into dw_sum fp32 [Cout][taps][Cin] (every K slice uses red.global.add on the one buffer): the running sum of
 * the raw weight gradient over the micro-batches of a gradient-accumulation cycle (cs_train.py:108-109).  The weight-norm
 * backward is linear in it for fixed weights, so ob_wnorm_bwd runs once per cycle on the sum (n_split = 1) instead of once
 * per micro-batch; the caller zeroes dw_sum when a cycle starts. */
int ob_conv_wgrad_acc(const void* gya, const void* x, const void* gb, const void* ctx, float* dw_sum, int n_seq, int S, int T,
                      int H, int W, int cin, int cout, int ksize, int gated, int n_split, void* stream);

/* Backward pre-pass of the gate (mp_sum with a per-frame tensor t, edm2/conv.py:95 -> edm2/utils.py:122-123):
 * from dy, the saved y (bf16) and d (fp16) it emits gya = alpha*dy, gb = sum_s beta*dy and the per-frame inner products
 * s_y[f] += <dy,y>, s_d[f] += <dy,d> (fp32 [n_seq*S*T], must be zeroed by the caller) that the five Gating
 * scalars' gradients are built from. */
int ob_gate_bwd(const void* dy, const void* y, const void* d, const float* alpha, const float* beta, void* gya, void* gb,
                float* s_y, float* s_d, int n_seq, int S, int T, int64_t frame_elems, void* stream);

/* Fused variants used by the training path (same maths as the three calls they replace, fewer launches):
 * ob_conv_prologue = ob_ctx_build + ob_gate_fwd + zeroing of `scratch` (fp32 [2*frames + 1], may be NULL in eval);
 *   pad_batch_stride: elements between the two cached frames of consecutive sequences in `pad` (0 = dense 2*frame_elems),
 *   so the last two frames of the previous call's context tensor can be passed in place (decode path, no copy);
 *   n_ctx_dev (optional, may be NULL): device int32 added to n_ctx -- the context-frame count of a graph-replayed decode
 *   step lives on the device so that one captured graph serves every generated frame;
 * ob_gate_bwd_fused = ob_gate_bwd with s_y = scratch, s_d = scratch + frames, and the last CTA to finish (ticket
 * counter at scratch[2*frames]) doing the work of ob_gate_bwd_params. */
int ob_conv_prologue(const void* x, const void* pad, void* ctx, int b, int S, int T, int64_t frame_elems, int cin,
                     int cin_pad, const float* offset, const float* mult, const float* max_gating, const float* min_gating,
                     const float* c_noise, float* alpha, float* beta, float* scratch, int n_ctx, int64_t pad_batch_stride,
                     const int* n_ctx_dev, void* stream);
int ob_gate_bwd_fused(const void* dy, const void* y, const void* d, const float* alpha, const float* beta, void* gya,
                      void* gb, float* scratch, int n_seq, int S, int T, int64_t frame_elems, const float* offset,
                      const float* mult, const float* max_gating, const float* min_gating, const float* c_noise,
                      float* g_offset, float* g_mult, float* g_max, float* g_min, int n_ctx, void* stream);

/* edm2/conv.py:113-127 Gating.forward + the mp_sum weights of edm2/utils.py:122-123, one launch:
 * alpha[f] = (1-g)/sqrt((1-g)^2+g^2), beta[f] = g/sqrt(..); position of frame f = (f % T) % half + n_ctx where T is
 * the frames per batch row of c_noise [frames] and half = T/2 in training (clean+noised halves share positions). */
int ob_gate_fwd(const float* offset, const float* mult, const float* max_gating, const float* min_gating,
                const float* c_noise, float* alpha, float* beta, int frames, int T, int half, int n_ctx, void* stream);
/* Gradients of the six gate scalars from ob_gate_bwd's inner products; ADDED into g_offset[2], g_mult[2], g_max[1],
 * g_min[1] (the parameters' .grad buffers). */
int ob_gate_bwd_params(const float* offset, const float* mult, const float* max_gating, const float* min_gating,
                       const float* c_noise, const float* alpha, const float* beta, const float* s_y, const float* s_d,
                       float* g_offset, float* g_mult, float* g_max, float* g_min, int frames, int T, int half, int n_ctx,
                       void* stream);
/* edm2/conv.py:68-69,78-84: ctx[b] = [pad frames (cache['activations'] or ones) | clean frames of x], bf16
 * [B, T+2, H, W, cin_pad]; x: [B*S*T, H, W, cin_pad]; pad: [B, 2, H, W, cin_pad] or NULL (ones on channels < cin). */
int ob_ctx_build(const void* x, const void* pad, void* ctx, int b, int S, int T, int64_t frame_elems, int cin, int cin_pad,
                 void* stream);

/* ---------------------------------------------------------------------------------------------- elementwise
 * edm2/networks_edm2.py:70 normalize(x, dim=1) + edm2/utils.py:112-113 mp_silu, one pass.
 * mode 0: xn = x/(eps+rms_C(x)), act = mp_silu(xn).   mode 1: act = mp_silu(x) (xn unused, may be NULL). */
int ob_pixnorm_silu_fwd(const void* x, void* xn, void* act, int64_t rows, int c, float eps, int mode, void* stream);
int ob_pixnorm_silu_bwd(const void* x, const void* g_xn, const void* g_act, void* dx, int64_t rows, int c, float eps,
                        int mode, void* stream);

/* edm2/networks_edm2.py:75-77: out = mp_silu(y * cscale[frame, channel]); cscale fp32, row `frame` starts at
 * cscale + frame*ld (ld = 0 means ld = C; a column slice of the all-blocks embedding GEMM is passed in place).
 * dc: fp32 [frames][C] dense. */
int ob_scale_silu_fwd(const void* y, const float* cscale, void* out, int64_t rows, int c, int rows_per_frame, int ld,
                      void* stream);
int ob_scale_silu_bwd(const void* y, const float* cscale, const void* g, void* dy, float* dc, int frames, int c,
                      int rows_per_frame, int ld, void* stream);

/* edm2/utils.py:118-123 mp_sum with float t, fused with the clip_ of edm2/networks_edm2.py:93 (clip <= 0: none). */
int ob_mp_sum_fwd(const void* a, const void* b, void* out, int64_t n, float t, float clip, void* stream);
int ob_mp_sum_bwd(const void* g, const void* out, void* da, void* db, int64_t n, float t, float clip, void* stream);

/* mp_cat (edm2/utils.py:128-134; the decoder skip connections, edm2/networks_edm2.py:244) over bf16 NHWC rows:
 *   out[row] = [ a[row] * wa | b[row] * wb ],  wa = C/sqrt(ca)*(1-t), wb = C/sqrt(cb)*t, C = sqrt((ca+cb)/((1-t)^2+t^2)).
 * bwd: da = g[:, :ca]*wa, db = g[:, ca:]*wb.  ca, cb multiples of 8. */
int ob_mp_cat_fwd(const void* a, const void* b, void* out, int64_t rows, int ca, int cb, float t, void* stream);
int ob_mp_cat_bwd(const void* g, void* da, void* db, int64_t rows, int ca, int cb, float t, void* stream);

/* 2x resampling of bf16 NHWC frames with the UNet's [1,1] filter (edm2/utils.py:94-107).  h, w: the LARGE side.
 * pool != 0: out[f,y,x,:] = scale * sum of in's 2x2 block (down: scale 0.25; gradient of up: scale 1);
 * pool == 0: out[f,2y+i,2x+j,:] = scale * in[f,y,x,:]     (up: scale 1; gradient of down: scale 0.25). */
int ob_resample2x(const void* in, void* out, int64_t frames, int h, int w, int c, int pool, float scale, void* stream);

/* VAE ResBlock activation (edm2/vae/vae.py:77-83, :86-87), one pass each way over bf16 [B, rows_per_batch, C] rows:
 *   y = x / sqrt(mean_c(x^2) + eps);   film != NULL: y = y*(1 + film[b][c]) + film[b][C + c]  (fp32 [B][2C], the decoder's
 *   t_cond scale | shift);   out = silu(y).
 * bwd: g = dL/dout -> dx (the input is re-normalised, nothing but x was saved) and, with film, dfilm fp32 [B][2C]
 * (overwritten).  c % 8 == 0, c <= 1024; c_mean <= c is the number of REAL channels the mean runs over (rows padded
 * with zero channels up to a multiple of 8). */
int ob_vae_norm_silu_fwd(const void* x, const float* film, void* out, int b, int64_t rows_per_batch, int c, int c_mean, float eps,
                         void* stream);
int ob_vae_norm_silu_bwd(const void* x, const float* film, const void* g, void* dx, float* dfilm, int b, int64_t rows_per_batch,
                         int c, int c_mean, float eps, void* stream);

/* Data movement of the VAE's grouped causal conv (edm2/vae/vae.py:40-53), bf16 NHWC rows, c % 8 == 0:
 * ob_time_window, backward == 0: xs[b, t', px, j*c + ch] = x[b, t'*g + j - (kt-g), px, ch] for j < kt -- the kt input frames output
 *   group t' reads, side by side on the channel axis (x: [b, t, hw, c]; xs: [b, t/g, hw, kt*c]); frames before the start come
 *   from pad [b, kt-g, hw, c] (the conv cache) or, pad == NULL, from the first kt-g frames of x (:43-44);
 * backward != 0: the transpose, src = dxs -> dst = dx (padding frames receive nothing: they are detached copies).
 * ob_ungroup: 'b (c g) t h w -> b c (t g) h w' for channels ordered (g, cc): out[f*g + r, px, ch] = in[f, px, r*cc + ch]
 *   (inverse != 0: the other direction, used by the backward pass). */
int ob_time_window(const void* src, const void* pad, void* dst, int b, int t, int64_t hw, int c, int g, int kt, int backward, void* stream);
int ob_ungroup(const void* in, void* out, int64_t frames, int64_t hw, int g, int cc, int inverse, void* stream);

/* Bias gradient of a conv with bias (autograd of nn.Conv3d(bias=True), edm2/vae/vae.py:26-31): out[ch] += sum over rows of
 * g[row, ch]; g: bf16 [rows, c] (c % 8 == 0), out: fp32 [c], zeroed by the caller (partial sums are added with
 * atomics). */
int ob_colsum(const void* g, float* out, int64_t rows, int c, void* stream);

/* Programmatic dependent launch: when on (default; ONIRIS_PDL=0 in the environment forces it off), every kernel of the
 * library is launched so that its prologue overlaps the tail of the previous kernel on the stream.  mode 0: off, 1: every
 * kernel, 2: only the light (elementwise) kernels.  Returns the previous mode.  The training backward pass switches it off
 * while its second stream is active (see csrc/launch.cuh). */
int ob_set_pdl(int mode);

/* Optimizer step of the training loop (cs_train.py:121-125: torch.optim.AdamW.step, zero_grad, and the
 * PowerFunctionEMA.update copies of the weights, edm2/phema.py:104-109) over one flat fp32 range of n elements
 * (n % 4 == 0, 16-byte aligned buffers):
 *   m = b1*m + (1-b1)*g;  v = b2*v + (1-b2)*g*g;  p = p*(1 - lr*wd) - lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
 *   ema_k += (1 - beta_k) * (p - ema_k)   (either pointer may be NULL);   g = 0.
 * EMA coefficient: ema_ratio > 0 selects the power-function profile of edm2/phema.py:68-70,
 *   beta_k = (1 - ema_ratio/t)^ema_a_k   with ema_a_k = std_to_exp(std_k) + 1  and  ema_ratio/t = t_delta/t_next
 * evaluated on the device from the step count (so a CUDA-graph replay follows the schedule); ema_ratio <= 0 uses the
 * constant beta_k = ema_a_k (TraditionalEMA without ramp-up).
 * g is multiplied by grad_scale first (1/world_size after a SUM all-reduce: the mean costs no extra pass) and, when
 * max_grad_norm > 0, by min(1, max_grad_norm / (grad_scale*sqrt(opt_state[2]) + 1e-6)) -- clip_grad_norm_ of
 * gym_train.py:105 with the squared norm accumulated by ob_sumsq.
 * opt_state: device fp32 {t, lr, grad_sumsq} with t the 1-based step count of THIS update.  The call may cover any
 * 16-byte aligned sub-range, so a bucketed all-reduce can be pipelined with the update of the buckets already reduced. */
int ob_adamw_ema(float* p, float* g, float* m, float* v, float* ema1, float* ema2, int64_t n, const float* opt_state,
                 float beta1, float beta2, float eps, float weight_decay, float ema_a1, float ema_a2, float ema_ratio,
                 float grad_scale, float max_grad_norm, void* stream);
/* out[0] += sum_i g[i]^2 (n % 4 == 0, g 16-byte aligned); the caller zeroes out[0]. */
int ob_sumsq(const float* g, int64_t n, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------- attention
 * edm2/attention/attention_modules.py:59-77: compiled_flex_attention(q,k,v, make_train_mask / make_infer_mask) and
 * F.scaled_dot_product_attention.  q: bf16 [B, Lq, heads, 64], k,v: bf16 [B, Lk, heads, 64] (NHWC rows); q,k RMS-normalised and
 * rotary-embedded by the caller (logits bounded by 8, which the kernel relies on).  mask: OB_ATTN_FULL (one new
 * frame vs the whole cache; per-frame attention), OB_ATTN_CAUSAL (frame-causal prefill, InferenceMask
 * attention_masking.py:56-62) or OB_ATTN_DART (TrainingMask attention_masking.py:8-24 over 2*n_frames frames).
 * hw = tokens per frame.  o: bf16 [B, Lq, heads, 64]; lse: fp32 [B, heads, Lq] (log-sum-exp of the scaled logits), may be NULL. */
/* Block lists of edm2/attention/attention_masking.py:27-53 (training != 0: make_train_mask) / :64-90 (make_infer_mask),
 * bit-exact with the reference's BlockMask.from_kv_blocks inputs, for ONE (batch, head) slice, written to HOST memory:
 * kv_num_blocks int32 [n_rows], kv_indices int32 [n_rows][n_rows].  n_rows = 2n' (training) or n' with n' = n_frames, or
 * n_frames*image_size/128 when image_size < 128 (the 128-token regrouping); *n_rows = 0 where the reference
 * builds no block list (returns None / takes its dense path).  Call with NULL arrays first to query n_rows / block_size.
 * The kernels do not consume these lists (they iterate the same pattern implicitly); they exist for callers and tests. */
int ob_build_block_lists(int training, int n_frames, int image_size, int32_t* kv_num_blocks, int32_t* kv_indices, int* n_rows,
                         int* block_size);

#define OB_ATTN_FULL 0
#define OB_ATTN_CAUSAL 1
#define OB_ATTN_DART 2
/* OB_ATTN_DART_LISTED: TrainingMask AND the 128-token blocks make_train_mask lists when hw < 128 (attention_masking.py:32-53)
 * -- what the reference's compiled FlexAttention actually evaluates on the GPU for small frames (a noised query in block i
 * of its half sees only clean blocks < i and its own block).  Identical to OB_ATTN_DART when hw >= 128. */
#define OB_ATTN_DART_LISTED 3
int ob_attn_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int b, int heads, int lq, int lk, int hw,
                int n_frames, int mask, float scale, void* stream);

/* q/k/v preparation, edm2/attention/attention_modules.py:48-49 (split of the (head, c, {q,k,v}) channel order +
 * normalize(dim=-1) of q, k and v) fused with edm2/attention/RoPe.py:43-74 (rotary + xPos on q and k).
 * qkv: bf16 [rows, heads*192] (the 1x1 conv output, NHWC).  q, k, v: bf16 [rows, heads*64].  k_raw (optional): the
 * normalised but un-rotated keys (what the reference keeps in its KV cache).  cos_t/sin_t/scl_t: fp32 [P][64] tables
 * (already fp16-rounded like RoPe.py:24,28); pos_q/pos_k: int32 table row per FRAME (row / hw), <0 or NULL = no rotary. */
int ob_qkv_prep_fwd(const void* qkv, void* q, void* k, void* v, void* k_raw, const float* cos_t, const float* sin_t,
                    const float* scl_t, const int* pos_q, const int* pos_k, int64_t rows, int heads, int hw, float eps,
                    void* stream);
int ob_qkv_prep_bwd(const void* qkv, const void* dq, const void* dk, const void* dv, void* dqkv, const float* cos_t,
                    const float* sin_t, const float* scl_t, const int* pos_q, const int* pos_k, int64_t rows, int heads,
                    int hw, float eps, void* stream);
/* Rotary (key flavour: divided by the xPos scale) over cached un-rotated keys x -> y, both bf16 [rows, heads*64]. */
int ob_rope_k(const void* x, void* y, const float* cos_t, const float* sin_t, const float* scl_t, const int* pos,
              int64_t rows, int heads, int hw, void* stream);

/* ---------------------------------------------------------------------------------------------- paged KV-cache decode
 * The sampler's cached evaluation (edm2/sampler.py:53-75 -> attention_modules.py:51-57,69-70): the reference clones and
 * concatenates the WHOLE (k, v) cache and re-rotates every cached key on each of the 2*num_steps-1 evaluations per frame.
 * Here keys/values live in frame-sized pages of a pool  k_pages / v_pages: bf16 [n_pages, hw, heads, 64]  addressed through
 * page_table: int32 [B, max_pages] (page of frame t of sequence b) with lengths: int32 [B] committed frames per sequence,
 * both in DEVICE memory -- the launch parameters do not change as a sequence grows, so one CUDA graph serves every step.
 *
 * ob_kv_append: q/k/v preparation of ONE new frame per sequence (the split + RMS-norm + rotary of ob_qkv_prep_fwd) with the
 *   frame's position taken from lengths[b]; q: bf16 [B, hw, heads, 64]; the rotated key and the value are written in place
 *   into page page_table[b][lengths[b]].  Keys are stored rotated with the tables' fixed xPos centre (it cancels in q.k).
 *   The length is NOT advanced: a caller commits the frame by incrementing lengths (update_cache=True) or lets the next
 *   evaluation overwrite the slot.  cos_t/sin_t/scl_t: fp32 [n_pos][64].
 * ob_dart_attn_decode: o[b] = softmax(q[b] K[b]^T * scale) V[b] over the (lengths[b] + extra_frames) * hw keys of sequence
 *   b, unmasked (attention_modules.py:69-70), gathered tile by tile through the page table, with the key tiles sliced over
 *   n_split CTAs per (sequence, head, query tile) (split-KV).  o_part: fp32 [n_split, B, hw, heads, 64], l_part: fp32
 *   [n_split, B*heads, hw] scratch (may be NULL when n_split == 1); n_split from ob_dart_attn_decode_splits.
 *   hw must divide 128 or be a multiple of 128 (and of 8). */
int ob_kv_append(const void* qkv, void* q, void* k_pages, void* v_pages, const int* page_table, const int* lengths,
                 const float* cos_t, const float* sin_t, const float* scl_t, int b, int heads, int hw, int max_pages, int n_pos,
                 float eps, void* stream);
int ob_dart_attn_decode_splits(int b, int heads, int hw, int max_pages);
int ob_dart_attn_decode(const void* q, const void* k_pages, const void* v_pages, const int* page_table, const int* lengths,
                        void* o, float* o_part, float* l_part, int b, int heads, int hw, int max_pages, int n_pages,
                        int n_split, int extra_frames, float scale, void* stream);

/* Backward of ob_attn_fwd (autograd of the same reference calls).  o, lse: the forward's outputs; dout: bf16
 * [B, Lq, heads, 64]; ws: fp32 workspace of 2 * B * heads * Lp floats, Lp = lq rounded up to a multiple of 128, 256-byte
 * aligned (receives -rowsum(dout*o)*scale and -lse*log2(e), zero-padded per row).  Writes dq, dk, dv (bf16, shaped like
 * q, k, v).
 * Every output element is produced by exactly one CTA (no atomics, deterministic). */
int ob_attn_bwd(const void* q, const void* k, const void* v, const void* o, const void* dout, const float* lse,
                float* ws, void* dq, void* dk, void* dv, int b, int heads, int lq, int lk, int hw, int n_frames,
                int mask, float scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ONIRIS_B200_H */
