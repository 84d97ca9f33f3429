// extern "C" surface of liboniris_b200.so (declared in include/oniris_b200.h).
#include "../../include/oniris_b200.h"
#include "launch.cuh"
#include <cstdlib>

#include <vector>

#include "hbm_host.h"
#include "tapconv.cuh"
#include "tapconv_host.h"
#include "wgrad.cuh"

using namespace ob;

namespace ob {
int attn_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int B, int heads, int Lq, int Lk, int hw,
             int n_frames, int mask, float scale, cudaStream_t st);
int attn_bwd(const void* q, const void* k, const void* v, const void* o, const void* dout, const float* lse, float* dsum,
             void* dq, void* dk, void* dv, int B, int heads, int Lq, int Lk, int hw, int n_frames, int mask, float scale,
             cudaStream_t st);
int attn_decode_splits(int B, int heads, int hw, int max_pages);
int attn_decode(const void* q, const void* k_pages, const void* v_pages, const int* page_table, const int* lengths, void* o,
                float* o_part, float* l_part, int B, int heads, int hw, int max_pages, int n_pages, int n_split,
                int extra_frames, float scale, cudaStream_t st);
}

// A column of vertical taps at horizontal shift dx.  `flip` mirrors the kernel (input-gradient passes): the tap applied
// at shift (dy, dx) is then the forward tap (-dy, -dx).  tap_base: first tap of this 3x3 slice in the weight matrix.
static TapCol tap_col3(int src, int dt, int dx, int n_a, int acc, int seq_mul, int tap_base, bool flip) {
  TapCol c{};
  c.src = (int8_t)src; c.dt = (int8_t)dt; c.dx = (int8_t)dx; c.n_a = (int8_t)n_a; c.acc = (int8_t)acc;
  c.seq_mul = (int8_t)seq_mul; c.n_taps = 3;
  for (int d = 0; d < 3; ++d) {
    const int dy = d - 1;
    const int ky = flip ? 1 - dy : dy + 1, kx = flip ? 1 - dx : dx + 1;
    c.wtap[d] = tap_base + ky * 3 + kx;
  }
  return c;
}
static TapCol tap_col1(int src, int n_a, int acc, int seq_mul) {
  TapCol c{};
  c.src = (int8_t)src; c.n_a = (int8_t)n_a; c.acc = (int8_t)acc; c.seq_mul = (int8_t)seq_mul; c.n_taps = 1;
  c.wtap[0] = 0;
  return c;
}
static void set_src(TapConvLaunch& L, int s, const void* p, int seq, int T, int H, int W, int C) {
  L.a[s] = p; L.a_seq[s] = seq; L.a_T[s] = T;
  L.a_stride_w[s] = C; L.a_stride_h[s] = (long)W * C; L.a_stride_t[s] = (long)H * W * C;
  L.a_stride_seq[s] = (long)T * H * W * C;
}
static int check_shape(const char* who, int n_seq, int S, int T, int H, int W, int ksize, int gated) {
  if (n_seq < 0 || T < 0 || H <= 0 || W <= 0 || (S != 1 && S != 2) || (ksize != 1 && ksize != 3) || (gated && ksize != 3)) {
    set_error("%s: bad shape n_seq=%d S=%d T=%d H=%d W=%d ksize=%d gated=%d", who, n_seq, S, T, H, W, ksize, gated);
    return OB_ERR_INVALID;
  }
  return OB_OK;
}

extern "C" {

int ob_version(void) { return 100; }
const char* ob_last_error(void) { return last_error(); }

int ob_wnorm_fwd(float* w, void* wg, int cout, int cin, int taps, int cin_pad, int taps_total, int tap_off, float gain,
                 float eps, int training, void* stream) {
  return wnorm_fwd(w, wg, cout, cin, taps, cin_pad, taps_total, tap_off, gain, eps, training, (cudaStream_t)stream);
}
int ob_wnorm_fwd_multi(const ob_wnorm_job* jobs, const int* row_start, int n_jobs, int total_rows, float eps, void* stream) {
  return wnorm_fwd_multi(jobs, row_start, n_jobs, total_rows, eps, (cudaStream_t)stream);
}
int ob_wnorm_bwd(const float* w, const float* dwg, float* dw, int cout, int cin, int taps, int cin_pad, int taps_total,
                 int tap_off, int n_split, float gain, float eps, int accumulate, void* stream) {
  return wnorm_bwd(w, dwg, dw, cout, cin, taps, cin_pad, taps_total, tap_off, n_split, gain, eps, accumulate,
                   (cudaStream_t)stream);
}

int ob_wnorm_bwd_gated(const float* w2, float* dw2, const float* w3, float* dw3, const float* dwg, int cout, int cin,
                       int cin_pad, int n_split, float eps, int accumulate, void* stream) {
  return wnorm_bwd2(w2, dw2, 9, 0, 1.f, w3, dw3, 18, 9, 1.f, dwg, cout, cin, cin_pad, 27, n_split, eps, accumulate,
                    (cudaStream_t)stream);
}

int64_t ob_conv_split_ws_bytes(int n_seq, int S, int T, int H, int W, int cin, int cout, int ksize, int gated) {
  int ks; long wsb;
  if (gated) tapconv_plan(n_seq, S, 1, 27, T, H, W, cin, cout, &ks, &wsb);
  else tapconv_plan(1, 1, 0, ksize * ksize, n_seq * S * T, H, W, cin, cout, &ks, &wsb);
  return (int64_t)wsb;
}

struct PostArgs { int post; void* out2; const float* cscale; int cscale_ld; const void* res; float t, clip; };
static int conv_fwd_impl(const void* x, const void* ctx, const void* wg, const float* alpha, const float* beta, void* out,
                         void* out_d, void* split_ws, int n_seq, int S, int T, int H, int W, int cin, int cout, int ksize, int gated,
                         int out_f32, int w_taps, const float* bias, const PostArgs& pa, void* stream) {
  if (int r = check_shape("ob_conv_fwd", n_seq, S, T, H, W, ksize, gated)) return r;
  TapConvLaunch L;
  std::vector<TapCol> cols;
  if (!gated) {
    // frames are independent: fold (seq, S, T) into one long frame axis so tiles never straddle less than they must
    const int frames = n_seq * S * T;
    set_src(L, 0, x, 1, frames, H, W, cin);
    if (ksize == 1) cols.push_back(tap_col1(0, 1, 0, 1));
    else for (int dx = -1; dx <= 1; ++dx) cols.push_back(tap_col3(0, 0, dx, 1, 0, 1, 0, false));
    L.n_seq = 1; L.n_out = 1; L.T = frames; L.w_taps = w_taps > 0 ? w_taps : ksize * ksize; L.epi = EPI_PLAIN; L.halo = ksize == 3;
    if (L.w_taps < ksize * ksize) { set_error("ob_conv_fwd: w_taps %d < %d", w_taps, ksize * ksize); return OB_ERR_INVALID; }
  } else {
    set_src(L, 0, x, n_seq * S, T, H, W, cin);
    set_src(L, 1, ctx, n_seq, T + 2, H, W, cin);
    // context taps first: they accumulate into the shared accumulator, which the persistent kernel rotates between two
    // TMEM ranges, so this half of the main loop overlaps the previous tile's epilogue
    for (int tau = 0; tau < 2; ++tau)
      for (int dx = -1; dx <= 1; ++dx) cols.push_back(tap_col3(1, tau, dx, 1, S, 1, 9 + tau * 9, false));
    for (int dx = -1; dx <= 1; ++dx) cols.push_back(tap_col3(0, 0, dx, S, 0, S, 0, false));
    L.n_seq = n_seq; L.n_out = S; L.T = T; L.w_taps = 27; L.epi = EPI_GATED; L.halo = 1;
    L.alpha = alpha; L.beta = beta; L.out_d = out_d;
  }
  L.wg = wg; L.cols = cols.data(); L.n_cols = (int)cols.size();
  L.H = H; L.W = W; L.Cin = cin; L.Cout = cout; L.out_f32 = out_f32; L.out = out; L.split_ws = (float*)split_ws;
  if (bias != nullptr && gated) { set_error("ob_conv_fwd: bias is supported for plain convs only"); return OB_ERR_INVALID; }
  L.bias = bias;
  if (pa.post != 0) {
    const float nrm = 1.f / sqrtf((1.f - pa.t) * (1.f - pa.t) + pa.t * pa.t);
    L.post = pa.post; L.out2 = pa.out2; L.cscale = pa.cscale; L.cscale_ld = pa.cscale_ld ? pa.cscale_ld : cout; L.res = pa.res;
    L.post_wa = (1.f - pa.t) * nrm; L.post_wb = pa.t * nrm; L.post_clip = pa.clip;
  }
  return tapconv_launch(L, (cudaStream_t)stream);
}
}  // extern "C" (re-opened below)

extern "C" {
int ob_conv_fwd(const void* x, const void* ctx, const void* wg, const float* alpha, const float* beta, void* out,
                void* out_d, void* split_ws, int n_seq, int S, int T, int H, int W, int cin, int cout, int ksize, int gated,
                int out_f32, int w_taps, const float* bias, void* stream) {
  return conv_fwd_impl(x, ctx, wg, alpha, beta, out, out_d, split_ws, n_seq, S, T, H, W, cin, cout, ksize, gated, out_f32, w_taps,
                       bias, PostArgs{0, nullptr, nullptr, 0, nullptr, 0.f, 0.f}, stream);
}
int ob_conv_fwd_fused(const void* x, const void* ctx, const void* wg, const float* alpha, const float* beta, void* out,
                      void* out_d, void* split_ws, int n_seq, int S, int T, int H, int W, int cin, int cout, int ksize, int gated,
                      int w_taps, int post, void* out2, const float* cscale, int cscale_ld, const void* res, float t, float clip,
                      void* stream) {
  if (post != 1 && post != 2) { set_error("ob_conv_fwd_fused: post must be OB_POST_SCALE_SILU or OB_POST_MP_SUM"); return OB_ERR_INVALID; }
  return conv_fwd_impl(x, ctx, wg, alpha, beta, out, out_d, split_ws, n_seq, S, T, H, W, cin, cout, ksize, gated, 0, w_taps, nullptr,
                       PostArgs{post, out2, cscale, cscale_ld, res, t, clip}, stream);
}


int ob_conv_dgrad(const void* gy, const void* gb, const void* wg, const float* alpha, const float* beta, void* dx,
                  void* split_ws, int n_seq, int S, int T, int H, int W, int cin, int cout, int ksize, int gated,
                  int w_taps, void* stream) {
  if (int r = check_shape("ob_conv_dgrad", n_seq, S, T, H, W, ksize, gated)) return r;
  // transposed problem: GEMM K = cout (channels of the incoming gradient), GEMM N = cin; taps are mirrored
  TapConvLaunch L;
  std::vector<TapCol> cols;
  if (!gated) {
    const int frames = n_seq * S * T;
    set_src(L, 0, gy, 1, frames, H, W, cout);
    if (ksize == 1) cols.push_back(tap_col1(0, 1, 0, 1));
    else for (int sx = -1; sx <= 1; ++sx) cols.push_back(tap_col3(0, 0, sx, 1, 0, 1, 0, true));
    L.n_seq = 1; L.n_out = 1; L.T = frames; L.w_taps = w_taps > 0 ? w_taps : ksize * ksize; L.epi = EPI_PLAIN; L.halo = ksize == 3;
    if (L.w_taps < ksize * ksize) { set_error("ob_conv_dgrad: w_taps %d < %d", w_taps, ksize * ksize); return OB_ERR_INVALID; }
  } else {
    set_src(L, 0, gy, n_seq * S, T, H, W, cout);
    set_src(L, 1, gb, n_seq, T, H, W, cout);
    for (int sx = -1; sx <= 1; ++sx) cols.push_back(tap_col3(0, 0, sx, S, 0, S, 0, true));
    // context frame t' fed output frames t'+2-tau through tap tau; that term goes to its own accumulator and the
    // epilogue adds it with weight beta (1 on clean rows, 0 on noised rows), so dy itself is never pre-scaled.
    for (int tau = 0; tau < 2; ++tau)
      for (int sx = -1; sx <= 1; ++sx) cols.push_back(tap_col3(1, 2 - tau, sx, 1, S, 1, 9 + tau * 9, true));
    L.n_seq = n_seq; L.n_out = S; L.T = T; L.w_taps = 27; L.epi = EPI_GATED; L.halo = 1;
    L.alpha = alpha; L.beta = beta;
  }
  L.b_mn_major = 1;
  L.wg = wg; L.cols = cols.data(); L.n_cols = (int)cols.size();
  L.H = H; L.W = W; L.Cin = cout; L.Cout = cin; L.out_f32 = 0; L.out = dx; L.split_ws = (float*)split_ws;
  return tapconv_launch(L, (cudaStream_t)stream);
}

int ob_conv_wgrad_splits(int n_seq, int S, int T, int H, int W, int cin, int cout, int ksize, int gated) {
  return wgrad_suggest_split(gated ? 27 : ksize * ksize, n_seq * S * T, H, W, cin, cout);
}

static int conv_wgrad_impl(const void* gya, const void* x, const void* gb, const void* ctx, float* dwg, int n_seq, int S, int T,
                           int H, int W, int cin, int cout, int ksize, int gated, int n_split, int accumulate, void* stream) {
  if (int r = check_shape("ob_conv_wgrad", n_seq, S, T, H, W, ksize, gated)) return r;
  WgradLaunch L;
  L.accumulate = accumulate;
  std::vector<WgradItem> items;
  auto mk = [](int pair, int dt, int dy, int dx, int wtap) {
    WgradItem t{}; t.pair = (int8_t)pair; t.dt = (int8_t)dt; t.dy = (int8_t)dy; t.dx = (int8_t)dx; t.wtap = wtap; return t;
  };
  if (!gated) {
    const int frames = n_seq * S * T;
    L.g[0] = gya; L.a[0] = x; L.g_seq[0] = 1; L.g_T[0] = frames; L.a_T[0] = frames;
    for (int ky = 0; ky < ksize; ++ky)
      for (int kx = 0; kx < ksize; ++kx) items.push_back(mk(0, 0, ky - ksize / 2, kx - ksize / 2, ky * ksize + kx));
    L.w_taps = ksize * ksize;
  } else {
    L.g[0] = gya; L.a[0] = x; L.g_seq[0] = n_seq * S; L.g_T[0] = T; L.a_T[0] = T;
    L.g[1] = gb; L.a[1] = ctx; L.g_seq[1] = n_seq; L.g_T[1] = T; L.a_T[1] = T + 2;
    for (int k = 0; k < 9; ++k) items.push_back(mk(0, 0, k / 3 - 1, k % 3 - 1, k));
    for (int tau = 0; tau < 2; ++tau)
      for (int k = 0; k < 9; ++k) items.push_back(mk(1, tau, k / 3 - 1, k % 3 - 1, 9 + tau * 9 + k));
    L.w_taps = 27;
  }
  L.items = items.data(); L.n_items = (int)items.size();
  L.H = H; L.W = W; L.Cin = cin; L.Cout = cout; L.n_split = n_split; L.out = dwg;
  return wgrad_launch(L, (cudaStream_t)stream);
}
int ob_conv_wgrad(const void* gya, const void* x, const void* gb, const void* ctx, float* dwg, int n_seq, int S, int T,
                  int H, int W, int cin, int cout, int ksize, int gated, int n_split, void* stream) {
  return conv_wgrad_impl(gya, x, gb, ctx, dwg, n_seq, S, T, H, W, cin, cout, ksize, gated, n_split, 0, stream);
}
int ob_conv_wgrad_acc(const void* gya, const void* x, const void* gb, const void* ctx, float* dw_sum, int n_seq, int S, int T,
                      int H, int W, int cin, int cout, int ksize, int gated, int n_split, void* stream) {
  return conv_wgrad_impl(gya, x, gb, ctx, dw_sum, n_seq, S, T, H, W, cin, cout, ksize, gated, n_split, 1, stream);
}

int ob_gate_bwd(const void* dy, const void* y, const void* d, const float* alpha, const float* beta, void* gya, void* gb,
                float* s_y, float* s_d, int n_seq, int S, int T, int64_t frame_elems, void* stream) {
  return gate_bwd(dy, y, d, alpha, beta, gya, gb, s_y, s_d, n_seq, S, T, (long)frame_elems, nullptr, nullptr, nullptr,
                  nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, (cudaStream_t)stream);
}
int ob_gate_bwd_fused(const void* dy, const void* y, const void* d, const float* alpha, const float* beta, void* gya,
                      void* gb, float* scratch, int n_seq, int S, int T, int64_t frame_elems, const float* offset,
                      const float* mult, const float* max_gating, const float* min_gating, const float* c_noise,
                      float* g_offset, float* g_mult, float* g_max, float* g_min, int n_ctx, void* stream) {
  const int f = n_seq * S * T;
  return gate_bwd(dy, y, d, alpha, beta, gya, gb, scratch, scratch + f, n_seq, S, T, (long)frame_elems, offset, mult,
                  max_gating, min_gating, c_noise, g_offset, g_mult, g_max, g_min, reinterpret_cast<unsigned*>(scratch + 2 * f),
                  n_ctx, (cudaStream_t)stream);
}
int ob_conv_prologue(const void* x, const void* pad, void* ctx, int b, int S, int T, int64_t frame_elems, int cin,
                     int cin_pad, const float* offset, const float* mult, const float* max_gating, const float* min_gating,
                     const float* c_noise, float* alpha, float* beta, float* scratch, int n_ctx, int64_t pad_batch_stride,
                     const int* n_ctx_dev, void* stream) {
  return conv_prologue(x, pad, ctx, b, S, T, (long)frame_elems, cin, cin_pad, offset, mult, max_gating, min_gating, c_noise,
                       alpha, beta, scratch, scratch ? 2 * b * S * T + 1 : 0, n_ctx, (long)pad_batch_stride, n_ctx_dev,
                       (cudaStream_t)stream);
}
int ob_gate_fwd(const float* offset, const float* mult, const float* max_gating, const float* min_gating,
                const float* c_noise, float* alpha, float* beta, int frames, int T, int half, int n_ctx, void* stream) {
  return gate_fwd(offset, mult, max_gating, min_gating, c_noise, alpha, beta, frames, T, half, n_ctx, (cudaStream_t)stream);
}
int ob_gate_bwd_params(const float* offset, const float* mult, const float* max_gating, const float* min_gating,
                       const float* c_noise, const float* alpha, const float* beta, const float* s_y, const float* s_d,
                       float* g_offset, float* g_mult, float* g_max, float* g_min, int frames, int T, int half, int n_ctx,
                       void* stream) {
  return gate_bwd_params(offset, mult, max_gating, min_gating, c_noise, alpha, beta, s_y, s_d, g_offset, g_mult, g_max,
                         g_min, frames, T, half, n_ctx, (cudaStream_t)stream);
}
int ob_ctx_build(const void* x, const void* pad, void* ctx, int b, int S, int T, int64_t frame_elems, int cin, int cin_pad,
                 void* stream) {
  return ctx_build(x, pad, ctx, b, S, T, (long)frame_elems, cin, cin_pad, (cudaStream_t)stream);
}
int ob_pixnorm_silu_fwd(const void* x, void* xn, void* act, int64_t rows, int c, float eps, int mode, void* stream) {
  return pixnorm_silu_fwd(x, xn, act, (long)rows, c, eps, mode, (cudaStream_t)stream);
}
int ob_pixnorm_silu_bwd(const void* x, const void* g_xn, const void* g_act, void* dx, int64_t rows, int c, float eps,
                        int mode, void* stream) {
  return pixnorm_silu_bwd(x, g_xn, g_act, dx, (long)rows, c, eps, mode, (cudaStream_t)stream);
}
int ob_scale_silu_fwd(const void* y, const float* cscale, void* out, int64_t rows, int c, int rows_per_frame, int ld, void* stream) {
  return scale_silu_fwd(y, cscale, out, (long)rows, c, rows_per_frame, ld, (cudaStream_t)stream);
}
int ob_scale_silu_bwd(const void* y, const float* cscale, const void* g, void* dy, float* dc, int frames, int c,
                      int rows_per_frame, int ld, void* stream) {
  return scale_silu_bwd(y, cscale, g, dy, dc, frames, c, rows_per_frame, ld, (cudaStream_t)stream);
}
int ob_mp_sum_fwd(const void* a, const void* b, void* out, int64_t n, float t, float clip, void* stream) {
  return mp_sum_fwd(a, b, out, (long)n, t, clip, (cudaStream_t)stream);
}
int ob_mp_sum_bwd(const void* g, const void* out, void* da, void* db, int64_t n, float t, float clip, void* stream) {
  return mp_sum_bwd(g, out, da, db, (long)n, t, clip, (cudaStream_t)stream);
}
int ob_mp_cat_fwd(const void* a, const void* b, void* out, int64_t rows, int ca, int cb, float t, void* stream) {
  return mp_cat(const_cast<void*>(a), const_cast<void*>(b), out, (long)rows, ca, cb, t, 0, (cudaStream_t)stream);
}
int ob_mp_cat_bwd(const void* g, void* da, void* db, int64_t rows, int ca, int cb, float t, void* stream) {
  return mp_cat(da, db, const_cast<void*>(g), (long)rows, ca, cb, t, 1, (cudaStream_t)stream);
}
int ob_resample2x(const void* in, void* out, int64_t frames, int h, int w, int c, int pool, float scale, void* stream) {
  return resample2x(in, out, (long)frames, h, w, c, pool, scale, (cudaStream_t)stream);
}
int ob_vae_norm_silu_fwd(const void* x, const float* film, void* out, int b, int64_t rows_per_batch, int c, int c_mean, float eps,
                         void* stream) {
  return vae_norm_silu_fwd(x, film, out, b, (long)rows_per_batch, c, c_mean, eps, (cudaStream_t)stream);
}
int ob_vae_norm_silu_bwd(const void* x, const float* film, const void* g, void* dx, float* dfilm, int b, int64_t rows_per_batch,
                         int c, int c_mean, float eps, void* stream) {
  return vae_norm_silu_bwd(x, film, g, dx, dfilm, b, (long)rows_per_batch, c, c_mean, eps, (cudaStream_t)stream);
}
int ob_time_window(const void* src, const void* pad, void* dst, int b, int t, int64_t hw, int c, int g, int kt, int backward, void* stream) {
  return time_window(src, pad, dst, b, t, (long)hw, c, g, kt, backward, (cudaStream_t)stream);
}
int ob_ungroup(const void* in, void* out, int64_t frames, int64_t hw, int g, int cc, int inverse, void* stream) {
  return ungroup(in, out, (long)frames, (long)hw, g, cc, inverse, (cudaStream_t)stream);
}
int ob_colsum(const void* g, float* out, int64_t rows, int c, void* stream) {
  return colsum(g, out, (long)rows, c, (cudaStream_t)stream);
}
int ob_set_pdl(int enabled) {
  static const bool forced_off = [] { const char* e = getenv("ONIRIS_PDL"); return e != nullptr && e[0] == '0'; }();
  const int prev = pdl_mode();
  if (!forced_off) pdl_flag().store(enabled < 0 ? 0 : enabled > 2 ? 2 : enabled, std::memory_order_relaxed);
  return prev;
}
int ob_adamw_ema(float* p, float* g, float* m, float* v, float* ema1, float* ema2, int64_t n, const float* opt_state,
                 float beta1, float beta2, float eps, float weight_decay, float ema_a1, float ema_a2, float ema_ratio,
                 float grad_scale, float max_grad_norm, void* stream) {
  return adamw_ema(p, g, m, v, ema1, ema2, (long)n, opt_state, beta1, beta2, eps, weight_decay, ema_a1, ema_a2, ema_ratio,
                   grad_scale, max_grad_norm, (cudaStream_t)stream);
}
int ob_sumsq(const float* g, int64_t n, float* out, void* stream) { return sumsq(g, (long)n, out, (cudaStream_t)stream); }

int ob_build_block_lists(int training, int n_frames, int image_size, int32_t* kv_num_blocks, int32_t* kv_indices, int* n_rows,
                         int* block_size) {
  // attention_masking.py:27-53 (make_train_mask) / :64-90 (make_infer_mask): HOST arrays, one (batch, head) slice -- the
  // reference broadcasts it over batch and heads.  Frames below 128 tokens are regrouped into 128-token blocks (F3).
  if (n_frames <= 0 || image_size <= 0 || n_rows == nullptr || block_size == nullptr) { set_error("ob_build_block_lists: bad arguments"); return OB_ERR_INVALID; }
  long n = n_frames;
  int bs = image_size;
  if (!training && static_cast<long>(n_frames) * image_size < 128) { *n_rows = 0; *block_size = 0; return OB_OK; }   // reference: score_mod path
  if (image_size < 128) {
    if ((n * image_size) % 128 != 0) { *n_rows = 0; *block_size = 0; return OB_OK; }                  // reference returns None / dense mask
    n = n * image_size / 128;
    bs = 128;
  }
  const long rows = training ? 2 * n : n;
  *n_rows = static_cast<int>(rows);
  *block_size = bs;
  if (kv_num_blocks == nullptr || kv_indices == nullptr) return OB_OK;                                  // size query
  for (long r = 0; r < rows; ++r) {
    const long i = r % n;
    kv_num_blocks[r] = static_cast<int32_t>(i + 1);
    int32_t* row = kv_indices + r * rows;
    for (long c = 0; c < rows; ++c) row[c] = 0;
    if (!training || r < n) { for (long c = 0; c <= i; ++c) row[c] = static_cast<int32_t>(c); }
    else { for (long c = 0; c < i; ++c) row[c] = static_cast<int32_t>(c); row[i] = static_cast<int32_t>(n + i); }
  }
  return OB_OK;
}

int ob_attn_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int b, int heads, int lq, int lk, int hw,
                int n_frames, int mask, float scale, void* stream) {
  return attn_fwd(q, k, v, o, lse, b, heads, lq, lk, hw, n_frames, mask, scale, (cudaStream_t)stream);
}

int ob_kv_append(const void* qkv, void* q, void* k_pages, void* v_pages, const int* page_table, const int* lengths,
                 const float* cos_t, const float* sin_t, const float* scl_t, int b, int heads, int hw, int max_pages, int n_pos,
                 float eps, void* stream) {
  return kv_append(qkv, q, k_pages, v_pages, page_table, lengths, cos_t, sin_t, scl_t, b, heads, hw, max_pages, n_pos, eps,
                   (cudaStream_t)stream);
}
int ob_dart_attn_decode_splits(int b, int heads, int hw, int max_pages) { return attn_decode_splits(b, heads, hw, max_pages); }
int ob_dart_attn_decode(const void* q, const void* k_pages, const void* v_pages, const int* page_table, const int* lengths,
                        void* o, float* o_part, float* l_part, int b, int heads, int hw, int max_pages, int n_pages,
                        int n_split, int extra_frames, float scale, void* stream) {
  return attn_decode(q, k_pages, v_pages, page_table, lengths, o, o_part, l_part, b, heads, hw, max_pages, n_pages, n_split,
                     extra_frames, scale, (cudaStream_t)stream);
}

int ob_attn_bwd(const void* q, const void* k, const void* v, const void* o, const void* dout, const float* lse,
                float* dsum, void* dq, void* dk, void* dv, int b, int heads, int lq, int lk, int hw, int n_frames,
                int mask, float scale, void* stream) {
  return attn_bwd(q, k, v, o, dout, lse, dsum, dq, dk, dv, b, heads, lq, lk, hw, n_frames, mask, scale,
                  (cudaStream_t)stream);
}

int ob_qkv_prep_fwd(const void* qkv, void* q, void* k, void* v, void* k_raw, const float* cos_t, const float* sin_t,
                    const float* scl_t, const int* pos_q, const int* pos_k, int64_t rows, int heads, int hw, float eps,
                    void* stream) {
  return qkv_prep_fwd(qkv, q, k, v, k_raw, cos_t, sin_t, scl_t, pos_q, pos_k, (long)rows, heads, hw, eps, (cudaStream_t)stream);
}
int ob_qkv_prep_bwd(const void* qkv, const void* dq, const void* dk, const void* dv, void* dqkv, const float* cos_t,
                    const float* sin_t, const float* scl_t, const int* pos_q, const int* pos_k, int64_t rows, int heads,
                    int hw, float eps, void* stream) {
  return qkv_prep_bwd(qkv, dq, dk, dv, dqkv, cos_t, sin_t, scl_t, pos_q, pos_k, (long)rows, heads, hw, eps,
                      (cudaStream_t)stream);
}
int ob_rope_k(const void* x, void* y, const float* cos_t, const float* sin_t, const float* scl_t, const int* pos,
              int64_t rows, int heads, int hw, void* stream) {
  return rope_k(x, y, cos_t, sin_t, scl_t, pos, (long)rows, heads, hw, (cudaStream_t)stream);
}

}  // extern "C"
