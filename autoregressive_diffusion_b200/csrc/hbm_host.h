// Host launchers of the bandwidth-bound kernels (hbm_kernels.cu).
#pragma once
#include <cuda_runtime.h>

namespace ob {
int wnorm_fwd(float* w, void* wg, int Co, int Ci, int taps, int Ci_pad, int taps_total, int tap_off, float gain, float eps,
              int training, cudaStream_t st);
int wnorm_fwd_multi(const void* jobs, const int* row_start, int n_jobs, int total_rows, float eps, cudaStream_t st);
int wnorm_bwd2(const float* w0, float* dw0, int taps0, int tap_off0, float gain0, const float* w1, float* dw1, int taps1,
               int tap_off1, float gain1, const float* dwg, int Co, int Ci, int Ci_pad, int taps_total, int n_split, float eps,
               int accumulate, cudaStream_t st);
int wnorm_bwd(const float* w, const float* dwg, float* dw, int Co, int Ci, int taps, int Ci_pad, int taps_total,
              int tap_off, int n_split, float gain, float eps, int accumulate, cudaStream_t st);
int gate_bwd(const void* dy, const void* y, const void* d, const float* alpha, const float* beta, void* gya, void* gb,
             float* s_y, float* s_d, int n_seq, int S, int T, long frame_elems, const float* offset, const float* mult,
             const float* max_g, const float* min_g, const float* c_noise, float* g_offset, float* g_mult, float* g_max,
             float* g_min, unsigned* counter, int n_ctx, cudaStream_t st);
int conv_prologue(const void* x, const void* pad, void* ctx, int B, int S, int T, long frame_elems, int cin, int cin_pad,
                  const float* offset, const float* mult, const float* max_g, const float* min_g, const float* c_noise,
                  float* alpha, float* beta, float* scratch, int scratch_n, int n_ctx, long pad_bstride, const int* n_ctx_dev,
                  cudaStream_t st);
int gate_fwd(const float* offset, const float* mult, const float* max_g, const float* min_g, const float* c_noise,
             float* alpha, float* beta, int frames, int T, int half, int n_ctx, cudaStream_t st);
int gate_bwd_params(const float* offset, const float* mult, const float* max_g, const float* min_g, const float* c_noise,
                    const float* alpha, const float* beta, const float* s_y, const float* s_d, float* g_offset,
                    float* g_mult, float* g_max, float* g_min, int frames, int T, int half, int n_ctx, cudaStream_t st);
int ctx_build(const void* x, const void* pad, void* ctx, int B, int S, int T, long frame_elems, int cin, int cin_pad,
              cudaStream_t st);
int pixnorm_silu_fwd(const void* x, void* xn, void* act, long rows, int C, float eps, int mode, cudaStream_t st);
int pixnorm_silu_bwd(const void* x, const void* g_xn, const void* g_act, void* dx, long rows, int C, float eps, int mode,
                     cudaStream_t st);
int scale_silu_fwd(const void* y, const float* cscale, void* out, long rows, int C, int rows_per_frame, int ld, cudaStream_t st);
int scale_silu_bwd(const void* y, const float* cscale, const void* g, void* dy, float* dc, int frames, int C,
                   int rows_per_frame, int ld, cudaStream_t st);
int mp_sum_fwd(const void* a, const void* b, void* out, long n, float t, float clip, cudaStream_t st);
int mp_sum_bwd(const void* g, const void* out, void* da, void* db, long n, float t, float clip, cudaStream_t st);
int mp_cat(void* a, void* b, void* cat, long rows, int ca, int cb, float t, int backward, cudaStream_t st);
int vae_norm_silu_fwd(const void* x, const float* film, void* out, int B, long rows_per_batch, int C, int c_mean, float eps,
                      cudaStream_t st);
int vae_norm_silu_bwd(const void* x, const float* film, const void* g, void* dx, float* dfilm, int B, long rows_per_batch, int C,
                      int c_mean, float eps, cudaStream_t st);
int time_window(const void* src, const void* pad, void* dst, int B, int T, long hw, int C, int g, int kt, int backward, cudaStream_t st);
int ungroup(const void* in, void* out, long frames, long hw, int g, int Cc, int inverse, cudaStream_t st);
int colsum(const void* g, float* out, long rows, int C, cudaStream_t st);
int resample2x(const void* in, void* out, long frames, int h, int w, int c, int pool, float scale, cudaStream_t st);
int adamw_ema(float* p, float* g, float* m, float* v, float* e1, float* e2, long n, const float* opt_state, float beta1,
              float beta2, float eps, float wd, float ema_a1, float ema_a2, float ema_ratio, float grad_scale, float max_norm,
              cudaStream_t st);
int sumsq(const float* g, long n, float* out, cudaStream_t st);
int qkv_prep_fwd(const void* qkv, void* q, void* k, void* v, void* k_raw, const float* cosT, const float* sinT,
                 const float* sclT, const int* pos_q, const int* pos_k, long rows, int heads, int hw, float eps,
                 cudaStream_t st);
int qkv_prep_bwd(const void* qkv, const void* dq, const void* dk, const void* dv, void* dqkv, const float* cosT,
                 const float* sinT, const float* sclT, const int* pos_q, const int* pos_k, long rows, int heads, int hw,
                 float eps, cudaStream_t st);
int kv_append(const void* qkv, void* q, void* k_pages, void* v_pages, const int* page_table, const int* lengths,
              const float* cosT, const float* sinT, const float* sclT, int B, int heads, int hw, int max_pages, int n_pos,
              float eps, cudaStream_t st);
int rope_k(const void* x, void* y, const float* cosT, const float* sinT, const float* sclT, const int* pos, long rows,
           int heads, int hw, cudaStream_t st);
}  // namespace ob
