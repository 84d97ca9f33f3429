// Internal host-side interface shared by the kernel translation units and the C-ABI layer.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>

#include "../../include/oniris_b200.h"  // OB_OK / OB_ERR_* status codes

namespace ob {

const char* last_error();
void set_error(const char* fmt, ...);

struct TapCol;  // defined in tapconv.cuh

int encode_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                     const uint32_t* box);

// One tap-GEMM problem. Activations are bf16 NHWC tensors [seq, T, H, W, Cin] with explicit element
// strides (channel stride 1); weights are bf16 [Cout, w_taps, Cin] contiguous.
struct TapConvLaunch {
  const void* a[2] = {nullptr, nullptr};
  int a_seq[2] = {0, 0}, a_T[2] = {0, 0};
  long a_stride_w[2] = {0, 0}, a_stride_h[2] = {0, 0}, a_stride_t[2] = {0, 0}, a_stride_seq[2] = {0, 0};
  const void* wg = nullptr;
  int w_taps = 0;
  const void* cols = nullptr;  // TapCol[n_cols]
  int n_cols = 0;
  int halo = 1;  // 1 for 3x3 kernels (taps dy = -1,0,+1 share one activation tile), 0 for 1x1
  int n_seq = 0, n_out = 1, T = 0, H = 0, W = 0, Cin = 0, Cout = 0;
  int epi = 0, out_f32 = 0;
  const float* alpha = nullptr;
  const float* beta = nullptr;
  const float* bias = nullptr;   // optional fp32 [Cout] (plain epilogue)
  int post = 0;                  // POST_* (tapconv.cuh): fused second output
  void* out2 = nullptr;
  const float* cscale = nullptr;
  int cscale_ld = 0;
  const void* res = nullptr;
  float post_wa = 0.f, post_wb = 0.f, post_clip = 0.f;
  void* out = nullptr;
  void* out_d = nullptr;
  int force_bn = 0;  // test hook: pin the N tile
  int b_mn_major = 0;  // weights given as [Cin][w_taps][Cout] (input-gradient passes)
  float* split_ws = nullptr;   // optional fp32 workspace enabling split-K (size from tapconv_plan)
  long long* trace = nullptr;  // probe only: per-CTA globaltimer stamps [ctas][8]
  int no_persist = 0;          // probe only: one CTA per tile (the pre-persistent schedule)
  int use_pair = 1;            // allow CTA-pair (cta_group::2) execution where the shape qualifies
};

// Split-K plan for a problem: number of channel-chunk slices and the fp32 workspace they need (0 = no split).
void tapconv_plan(int n_seq, int n_out, int gated, int taps, int T, int H, int W, int Cin, int Cout, int* ksplit, long* ws_bytes);

int tapconv_launch(const TapConvLaunch& L, cudaStream_t stream);

}  // namespace ob

namespace ob {

// One weight-gradient problem: up to two (G, A) tensor pairs sharing H, W, Cin, Cout.
//   G pair p: bf16 [g_seq[p], g_T[p], H, W, Cout] contiguous; A pair p: bf16 [g_seq[p], a_T[p], H, W, Cin] contiguous.
struct WgradLaunch {
  const void* g[2] = {nullptr, nullptr};
  const void* a[2] = {nullptr, nullptr};
  int g_seq[2] = {0, 0}, g_T[2] = {0, 0}, a_T[2] = {0, 0};
  const void* items = nullptr;  // WgradItem[n_items]
  int n_items = 0;
  int H = 0, W = 0, Cin = 0, Cout = 0, w_taps = 0;
  int n_split = 1;
  int accumulate = 0;    // 1: the K slices add into out[0] (a running sum the caller zeroed) instead of storing n_split slices
  int force_mode = 0;    // probe only: 1 = one tap per CTA and no pairs, 2 = CTA pairs, 3 = halo groups, where the shape allows
  float* out = nullptr;  // [n_split, Cout, w_taps, Cin]
};

int wgrad_suggest_split(int n_items, int max_rows, int H, int W, int Cin, int Cout);
int wgrad_launch(const WgradLaunch& L, cudaStream_t stream);

}  // namespace ob
