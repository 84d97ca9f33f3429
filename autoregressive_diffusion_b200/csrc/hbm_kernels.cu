// Bandwidth-bound kernels of the Oniris hot path: magnitude-preserving weight normalisation
// (forward + backward), the gate backward pre-pass, pixel-norm / mp_silu / mp_sum epilogues.
// All activations are bf16 NHWC ([rows, C], channel-contiguous); all arithmetic is fp32.
// 128-bit global accesses, warp-shuffle reductions, one pass over HBM per tensor.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "launch.cuh"
#include <cuda_runtime.h>
#include <cstdint>

#include "tapconv_host.h"

namespace ob {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum; every thread gets the result. red must hold 32 floats.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
  if (warp == 0) {
    t = warp_sum(t);
    if (lane == 0) red[0] = t;
  }
  __syncthreads();
  return red[0];
}

__device__ __forceinline__ float bf2f(__nv_bfloat16 v) { return __bfloat162float(v); }

struct alignas(16) bf16x8 {
  __nv_bfloat162 v[4];
};
__device__ __forceinline__ void unpack8(const bf16x8& p, float* f) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(p.v[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ bf16x8 pack8(const float* f) {
  bf16x8 p;
#pragma unroll
  for (int i = 0; i < 4; ++i) p.v[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return p;
}

// ============================================================================ weight norm
// Reference: edm2/conv.py:14-21 (NormalizedWeight.forward) + edm2/utils.py:83-88 (normalize).
// Parameters are STORED tap-major: the reference's logical [Co, Ci, (kt,) kh, kw] tensor in channels_last memory
// format, i.e. physically w[Co][taps][Ci] -- the order of the bf16 GEMM operand wg[Co][taps_total][Ci_pad] and of the
// wgrad partials dwg[n_split][Co][taps_total][Ci_pad].  Forward and backward are therefore straight row streams (no
// transposes, no index divisions): one CTA per output channel, 128-bit accesses.
// training: w <- w/(eps+rms(w)) in place (forced weight norm), operand = normalize(that) * gain/sqrt(fan_in).
__device__ __forceinline__ void wnorm_fwd_row(float* __restrict__ w, __nv_bfloat16* __restrict__ wg, int co, int Ci, int taps,
                                              int Ci_pad, int taps_total, int tap_off, float gain, float eps, int training,
                                              float* red) {
  const int K = Ci * taps;
  float* wr = w + static_cast<long>(co) * K;
  __nv_bfloat16* dst = wg + (static_cast<long>(co) * taps_total + tap_off) * Ci_pad;
  const bool vec = (Ci_pad == Ci) && ((K & 7) == 0) && ((reinterpret_cast<uintptr_t>(wr) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
  float ss = 0.f;
  if (vec) {
    const float4* w4 = reinterpret_cast<const float4*>(wr);
    for (int i = threadIdx.x; i < K / 4; i += blockDim.x) {
      const float4 v = w4[i];
      ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
  } else {
    for (int i = threadIdx.x; i < K; i += blockDim.x) ss += wr[i] * wr[i];
  }
  ss = block_sum(ss, red);
  const float rms = sqrtf(ss / K);
  const float s1 = 1.f / (eps + rms);          // first normalisation
  float scale = s1;
  if (training) scale = s1 / (eps + rms * s1);  // rms * s1 = rms of the forced weights
  scale *= gain * rsqrtf(static_cast<float>(K));
  if (vec) {   // second pass hits L1/L2
    for (int i = threadIdx.x; i < K / 8; i += blockDim.x) {
      const float4 a = reinterpret_cast<const float4*>(wr)[2 * i], c = reinterpret_cast<const float4*>(wr)[2 * i + 1];
      const float o[8] = {a.x * scale, a.y * scale, a.z * scale, a.w * scale, c.x * scale, c.y * scale, c.z * scale, c.w * scale};
      reinterpret_cast<bf16x8*>(dst)[i] = pack8(o);
      if (training) {
        reinterpret_cast<float4*>(wr)[2 * i] = make_float4(a.x * s1, a.y * s1, a.z * s1, a.w * s1);
        reinterpret_cast<float4*>(wr)[2 * i + 1] = make_float4(c.x * s1, c.y * s1, c.z * s1, c.w * s1);
      }
    }
  } else {
    for (int j = threadIdx.x; j < taps * Ci_pad; j += blockDim.x) {
      const int tap = j / Ci_pad, ci = j - tap * Ci_pad;
      dst[j] = __float2bfloat16((ci < Ci) ? wr[tap * Ci + ci] * scale : 0.f);
    }
    if (training) {
      __syncthreads();  // every read of the old values is done before the in-place overwrite
      for (int i = threadIdx.x; i < K; i += blockDim.x) wr[i] *= s1;
    }
  }
}

__global__ void __launch_bounds__(256) wnorm_fwd_kernel(float* __restrict__ w, __nv_bfloat16* __restrict__ wg, int Ci,
                                                        int taps, int Ci_pad, int taps_total, int tap_off, float gain,
                                                        float eps, int training) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[32];
  wnorm_fwd_row(w, wg, blockIdx.x, Ci, taps, Ci_pad, taps_total, tap_off, gain, eps, training, red);
}

// Every weight tensor of a network in ONE launch (the optimizer step invalidates all of them at once; 210 separate
// launches of ~13 us are latency-bound at 0.22 of the HBM roofline).  jobs: device array, one entry per weight tensor;
// row_start[j] = first global row (output channel) of job j, row_start[n_jobs] = total rows; one CTA per row.
__global__ void __launch_bounds__(256) wnorm_fwd_multi_kernel(const ob_wnorm_job* __restrict__ jobs,
                                                              const int* __restrict__ row_start, int n_jobs, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[32];
  const int row = blockIdx.x;
  int lo = 0, hi = n_jobs - 1;                 // last job whose first row is <= row
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (row_start[mid] <= row) lo = mid; else hi = mid - 1;
  }
  const ob_wnorm_job j = jobs[lo];
  wnorm_fwd_row(j.w, static_cast<__nv_bfloat16*>(j.wg), row - row_start[lo], j.cin, j.taps, j.cin_pad, j.taps_total, j.tap_off,
                j.gain, eps, j.training, red);
}

// Backward of operand = normalize(w) * gain/sqrt(K) w.r.t. the (forced) weights w:
//   d = eps + rms(w);  dw_i = c/d * (g_i - w_i * <g,w> / (K * rms * d)),  g = sum of the wgrad split partials.
// One CTA per (output channel, parameter): blockIdx.y selects one of up to two parameters that share the partials
// tensor (the 2D and the 3D weight of a gated conv: taps 9 at offset 0, taps 18 at offset 9) so a layer is ONE launch.
// Pass 1 streams g and w once and prefetches the old gradient row into L2; pass 2 re-reads g and w (L2 hits; short rows
// stay in registers) and writes dw (+)= ...  The kernel runs on the weight-gradient stream BESIDE the input-gradient
// chain, so its footprint is kept small on purpose (<= 64 registers, no shared-memory row): a variant caching whole
// 9216-float rows in registers was 20 % faster alone but took every register of the SM and slowed the two-stream
// backward pass (10.75 -> 10.94 ms).
constexpr int WNB_CACHE = 2;
struct WnormBwdPart {
  const float* w;
  float* dw;
  int taps, tap_off;
  float gain;
};
struct WnormBwdArgs {
  WnormBwdPart part[2];
  const float* dwg;
  int Co, Ci, Ci_pad, taps_total, n_split, accumulate;
  float eps;
};

__global__ void __launch_bounds__(256, 4) wnorm_bwd_kernel(const WnormBwdArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red2[2][8];
  const WnormBwdPart pt = a.part[blockIdx.y];
  const int co = blockIdx.x;
  const bool accumulate = (a.accumulate & 1) != 0;
  const bool zero_src = (a.accumulate & 2) != 0;     // the partials are a running sum (ob_conv_wgrad_acc): consume and clear it
  const int Ci = a.Ci, taps = pt.taps;
  const int K = Ci * taps;
  const float* wr = pt.w + static_cast<long>(co) * K;
  float* dwr = pt.dw + static_cast<long>(co) * K;
  const long split_stride = static_cast<long>(a.Co) * a.taps_total * a.Ci_pad;
  float* gsrc = const_cast<float*>(a.dwg) + (static_cast<long>(co) * a.taps_total + pt.tap_off) * a.Ci_pad;
  const bool vec = (a.Ci_pad == Ci) && ((K & 3) == 0) && ((split_stride & 3) == 0) &&
                   (((reinterpret_cast<uintptr_t>(wr) | reinterpret_cast<uintptr_t>(dwr) | reinterpret_cast<uintptr_t>(gsrc)) & 15) == 0);
  const int nv = K / 4;
  const bool cached = vec && nv <= WNB_CACHE * 256;
  float4 gv[WNB_CACHE], wv[WNB_CACHE];
  float ss = 0.f, dot = 0.f;
  auto load_g = [&](int i4) {
    float4 g = *reinterpret_cast<const float4*>(gsrc + i4 * 4);
    for (int sp = 1; sp < a.n_split; ++sp) {
      const float4 t = *reinterpret_cast<const float4*>(gsrc + sp * split_stride + i4 * 4);
      g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
    }
    return g;
  };
  if (cached) {
#pragma unroll
    for (int v = 0; v < WNB_CACHE; ++v) {
      const int i4 = threadIdx.x + v * 256;
      if (i4 < nv) {
        wv[v] = *reinterpret_cast<const float4*>(wr + i4 * 4);
        gv[v] = load_g(i4);
        if (accumulate) asm volatile("prefetch.global.L2 [%0];" ::"l"(dwr + i4 * 4));
      }
    }
#pragma unroll
    for (int v = 0; v < WNB_CACHE; ++v) {
      if (threadIdx.x + v * 256 < nv) {
        ss += wv[v].x * wv[v].x + wv[v].y * wv[v].y + wv[v].z * wv[v].z + wv[v].w * wv[v].w;
        dot += gv[v].x * wv[v].x + gv[v].y * wv[v].y + gv[v].z * wv[v].z + gv[v].w * wv[v].w;
      }
    }
  } else if (vec) {
    for (int i4 = threadIdx.x; i4 < nv; i4 += 256) {
      const float4 w4 = *reinterpret_cast<const float4*>(wr + i4 * 4);
      const float4 g = load_g(i4);
      ss += w4.x * w4.x + w4.y * w4.y + w4.z * w4.z + w4.w * w4.w;
      dot += g.x * w4.x + g.y * w4.y + g.z * w4.z + g.w * w4.w;
    }
  } else {
    for (int i = threadIdx.x; i < K; i += 256) {
      const int tap = i / Ci, ci = i - tap * Ci;
      float g = 0.f;
      for (int sp = 0; sp < a.n_split; ++sp) g += gsrc[sp * split_stride + tap * a.Ci_pad + ci];
      ss += wr[i] * wr[i];
      dot += g * wr[i];
    }
  }
  {   // both sums in one block reduction
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    ss = warp_sum(ss);
    dot = warp_sum(dot);
    if (lane == 0) { red2[0][warp] = ss; red2[1][warp] = dot; }
    __syncthreads();
    ss = 0.f; dot = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) { ss += red2[0][w8]; dot += red2[1][w8]; }
  }
  const float rms = sqrtf(ss / K);
  const float d = a.eps + rms;
  const float c = pt.gain * rsqrtf(static_cast<float>(K)) / d;
  const float proj = (rms > 0.f) ? dot / (K * rms * d) : 0.f;
  auto emit = [&](int i4, const float4& g, const float4& w4) {
    float4 o = make_float4(c * (g.x - w4.x * proj), c * (g.y - w4.y * proj), c * (g.z - w4.z * proj), c * (g.w - w4.w * proj));
    float4* dst = reinterpret_cast<float4*>(dwr + i4 * 4);
    if (accumulate) {
      const float4 old = *dst;
      o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
    }
    *dst = o;
    if (zero_src) *reinterpret_cast<float4*>(gsrc + i4 * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
  };
  if (cached) {
#pragma unroll
    for (int v = 0; v < WNB_CACHE; ++v) {
      const int i4 = threadIdx.x + v * 256;
      if (i4 < nv) emit(i4, gv[v], wv[v]);
    }
  } else if (vec) {
    for (int i4 = threadIdx.x; i4 < nv; i4 += 256) emit(i4, load_g(i4), *reinterpret_cast<const float4*>(wr + i4 * 4));
  } else {
    for (int i = threadIdx.x; i < K; i += 256) {
      const int tap = i / Ci, ci = i - tap * Ci;
      float g = 0.f;
      for (int sp = 0; sp < a.n_split; ++sp) g += gsrc[sp * split_stride + tap * a.Ci_pad + ci];
      const float v = c * (g - wr[i] * proj);
      dwr[i] = accumulate ? dwr[i] + v : v;
      if (zero_src) gsrc[tap * a.Ci_pad + ci] = 0.f;
    }
  }
}

// ============================================================================ gate backward pre-pass
// Reference: the mp_sum(last_frame_conv, context, gating) of edm2/conv.py:95 differentiated.
//   y = alpha*a + beta*b  (alpha,beta per frame), saved: y (bf16) and d = b - a (fp16: 11 mantissa bits at two bytes).
// Produces in ONE pass over dy:  gya = alpha*dy (all frames), gb[b,t] = sum_s beta_s*dy_s (context rows),
// and per-frame <dy,y>, <dy,d> (what the 5 gate scalars' gradients need).  S = 1 (eval-style) or 2 (clean+noised).
// Optional tail of gate_bwd_kernel: the gate parameters, their gradient buffers and a zeroed ticket counter.
__device__ __forceinline__ void unpack8_f16(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 p2 = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
    f[2 * i] = p2.x; f[2 * i + 1] = p2.y;
  }
}

struct GateGradArgs {
  const float *offset, *mult, *max_g, *min_g, *c_noise;
  float *g_offset, *g_mult, *g_max, *g_min;
  unsigned* counter;   // nullptr: skip (gradients are then produced by gate_bwd_params_kernel)
  int n_ctx;
};

__global__ void __launch_bounds__(256, 4) gate_bwd_kernel(const __nv_bfloat16* __restrict__ dy,
                                                       const __nv_bfloat16* __restrict__ y,
                                                       const __half* __restrict__ d,
                                                       const float* __restrict__ alpha, const float* __restrict__ beta,
                                                       __nv_bfloat16* __restrict__ gya, __nv_bfloat16* __restrict__ gb,
                                                       float* __restrict__ s_y, float* __restrict__ s_d, int n_seq, int S,
                                                       int T, long frame_elems, GateGradArgs gg) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[32];
  const int bt = blockIdx.y;  // (b, t)
  const int b = bt / T, t = bt - b * T;
  const long chunk0 = static_cast<long>(blockIdx.x) * blockDim.x * 8 + threadIdx.x * 8;
  const long stride = static_cast<long>(gridDim.x) * blockDim.x * 8;
  float sy[2] = {0.f, 0.f}, sd[2] = {0.f, 0.f};
  const long f0 = (static_cast<long>(b) * S) * T + t, f1 = f0 + T;   // clean / noised frame of this (b, t)
  const float al0 = alpha[f0], be0 = beta[f0];
  const float al1 = S > 1 ? alpha[f1] : 0.f, be1 = S > 1 ? beta[f1] : 0.f;
  for (long e = chunk0; e < frame_elems; e += stride) {
    // all loads of both halves first (8 independent 128-bit requests per thread), then the arithmetic and the stores
    const long o0 = f0 * frame_elems + e, o1 = f1 * frame_elems + e;
    const bf16x8 g0v = *reinterpret_cast<const bf16x8*>(dy + o0), y0v = *reinterpret_cast<const bf16x8*>(y + o0);
    const uint4 d0v = *reinterpret_cast<const uint4*>(d + o0);     // 8 fp16 values
    bf16x8 g1v = g0v, y1v = y0v;
    uint4 d1v = d0v;
    if (S > 1) {
      g1v = *reinterpret_cast<const bf16x8*>(dy + o1); y1v = *reinterpret_cast<const bf16x8*>(y + o1);
      d1v = *reinterpret_cast<const uint4*>(d + o1);
    }
    float g0[8], y0[8], g1[8], y1[8], oa[8], ob_[8], acc_b[8];
    unpack8(g0v, g0); unpack8(y0v, y0);
    float dv0[8];
    unpack8_f16(d0v, dv0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      oa[j] = al0 * g0[j];
      acc_b[j] = be0 * g0[j];
      sy[0] += g0[j] * y0[j];
      sd[0] += g0[j] * dv0[j];
    }
    *reinterpret_cast<bf16x8*>(gya + o0) = pack8(oa);
    if (S > 1) {
      unpack8(g1v, g1); unpack8(y1v, y1);
      float dv1[8];
      unpack8_f16(d1v, dv1);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        ob_[j] = al1 * g1[j];
        acc_b[j] += be1 * g1[j];
        sy[1] += g1[j] * y1[j];
        sd[1] += g1[j] * dv1[j];
      }
      *reinterpret_cast<bf16x8*>(gya + o1) = pack8(ob_);
    }
    *reinterpret_cast<bf16x8*>(gb + static_cast<long>(bt) * frame_elems + e) = pack8(acc_b);
  }
  // one combined block reduction of the four inner products (each CTA carries little data: keep its tail short)
  {
    __shared__ float red4[4][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float v[4] = {warp_sum(sy[0]), warp_sum(sd[0]), warp_sum(sy[1]), warp_sum(sd[1])};
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < 4; ++k) red4[k][warp] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < 2 * S) {
      const int k = threadIdx.x;            // k = 2*s + {0: <dy,y>, 1: <dy,d>}
      float tsum = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) tsum += red4[k][w];
      const long f = (static_cast<long>(b) * S + (k >> 1)) * T + t;
      atomicAdd((k & 1) ? &s_d[f] : &s_y[f], tsum);
    }
  }
  if (gg.counter == nullptr) return;
  // ---- last CTA to finish turns the completed inner products into the six gate-scalar gradients
  __shared__ unsigned last;
  __threadfence();
  if (threadIdx.x == 0) last = (atomicAdd(gg.counter, 1u) == gridDim.x * gridDim.y - 1) ? 1u : 0u;
  __syncthreads();
  if (!last) return;
  __threadfence();
  const int frames = n_seq * S * T;
  const int Trow = S * T;
  const float lo = 1.f / (1.f + expf(-gg.min_g[0])), hi = 1.f / (1.f + expf(-gg.max_g[0]));
  float a_state = 0.f, a_m0 = 0.f, a_m1 = 0.f, a_lo = 0.f, a_hi = 0.f;
  for (int f = threadIdx.x; f < frames; f += blockDim.x) {
    const float pos = log1pf(static_cast<float>((f % Trow) % T + gg.n_ctx));
    const float cn = gg.c_noise[f];
    const float state = cn * gg.mult[0] + gg.offset[0] + pos * gg.mult[1] + gg.offset[1];
    const float sg = 1.f / (1.f + expf(-state));
    const float g = lo + (1.f - lo) * hi * sg;
    const float al = alpha[f], be = beta[f];
    const float sy = __ldcg(&s_y[f]), sd = __ldcg(&s_d[f]);
    const float da = (sy - be * sd) / (al + be);
    const float db = da + sd;
    const float D = (1.f - g) * (1.f - g) + g * g;
    const float dg = (-g * da + (1.f - g) * db) / (D * sqrtf(D));
    const float dstate = dg * (1.f - lo) * hi * sg * (1.f - sg);
    a_state += dstate;
    a_m0 += dstate * cn;
    a_m1 += dstate * pos;
    a_lo += dg * (1.f - hi * sg);
    a_hi += dg * (1.f - lo) * sg;
  }
  a_state = block_sum(a_state, red);
  a_m0 = block_sum(a_m0, red);
  a_m1 = block_sum(a_m1, red);
  a_lo = block_sum(a_lo, red);
  a_hi = block_sum(a_hi, red);
  if (threadIdx.x == 0) {
    gg.g_offset[0] += a_state; gg.g_offset[1] += a_state;
    gg.g_mult[0] += a_m0; gg.g_mult[1] += a_m1;
    gg.g_min[0] += a_lo * lo * (1.f - lo);
    gg.g_max[0] += a_hi * hi * (1.f - hi);
  }
}

// ============================================================================ gate scalars (forward / backward)
// Reference: edm2/conv.py:113-127 (Gating.forward) followed by the mp_sum weights of edm2/utils.py:122-123:
//   g = lo + (1-lo)*hi*sigmoid(mult0*c_noise + off0 + mult1*log1p(pos) + off1),  lo = sigmoid(min_gating), hi = sigmoid(max_gating)
//   alpha = (1-g)/sqrt((1-g)^2+g^2),  beta = g/sqrt(...)       pos = (frame index within its sequence) % half + n_ctx
// One launch replaces ~15 eager ops per conv layer (and ~25 in backward).
__global__ void gate_fwd_kernel(const float* __restrict__ offset, const float* __restrict__ mult,
                                const float* __restrict__ max_g, const float* __restrict__ min_g,
                                const float* __restrict__ c_noise, float* __restrict__ alpha, float* __restrict__ beta,
                                int frames, int T, int half, int n_ctx) {
  pdl_launch_dependents();
  pdl_wait();
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= frames) return;
  const float lo = 1.f / (1.f + expf(-min_g[0])), hi = 1.f / (1.f + expf(-max_g[0]));
  const float pos = log1pf(static_cast<float>((f % T) % half + n_ctx));
  const float state = c_noise[f] * mult[0] + offset[0] + pos * mult[1] + offset[1];
  const float g = lo + (1.f - lo) * hi / (1.f + expf(-state));
  const float inv = rsqrtf((1.f - g) * (1.f - g) + g * g);
  alpha[f] = (1.f - g) * inv;
  beta[f] = g * inv;
}

// From the per-frame inner products s_y = <dy,y>, s_d = <dy,d> (y = alpha*a + beta*b, d = b - a):
//   <dy,a> = (s_y - beta*s_d)/(alpha+beta), <dy,b> = <dy,a> + s_d;  dg = (-g*<dy,a> + (1-g)*<dy,b>) / D^{3/2}
// then the chain rule through the gate; the six scalar gradients are ADDED to g_* (single CTA, no atomics).
__global__ void __launch_bounds__(128) gate_bwd_params_kernel(const float* __restrict__ offset, const float* __restrict__ mult,
                                                              const float* __restrict__ max_g, const float* __restrict__ min_g,
                                                              const float* __restrict__ c_noise, const float* __restrict__ alpha,
                                                              const float* __restrict__ beta, const float* __restrict__ s_y,
                                                              const float* __restrict__ s_d, float* __restrict__ g_offset,
                                                              float* __restrict__ g_mult, float* __restrict__ g_max,
                                                              float* __restrict__ g_min, int frames, int T, int half, int n_ctx) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[32];
  const float lo = 1.f / (1.f + expf(-min_g[0])), hi = 1.f / (1.f + expf(-max_g[0]));
  float a_state = 0.f, a_m0 = 0.f, a_m1 = 0.f, a_lo = 0.f, a_hi = 0.f;
  for (int f = threadIdx.x; f < frames; f += blockDim.x) {
    const float pos = log1pf(static_cast<float>((f % T) % half + n_ctx));
    const float cn = c_noise[f];
    const float state = cn * mult[0] + offset[0] + pos * mult[1] + offset[1];
    const float sg = 1.f / (1.f + expf(-state));
    const float g = lo + (1.f - lo) * hi * sg;
    const float al = alpha[f], be = beta[f];
    const float da = (s_y[f] - be * s_d[f]) / (al + be);
    const float db = da + s_d[f];
    const float D = (1.f - g) * (1.f - g) + g * g;
    const float dg = (-g * da + (1.f - g) * db) / (D * sqrtf(D));
    const float dstate = dg * (1.f - lo) * hi * sg * (1.f - sg);
    a_state += dstate;
    a_m0 += dstate * cn;
    a_m1 += dstate * pos;
    a_lo += dg * (1.f - hi * sg);
    a_hi += dg * (1.f - lo) * sg;
  }
  a_state = block_sum(a_state, red);
  a_m0 = block_sum(a_m0, red);
  a_m1 = block_sum(a_m1, red);
  a_lo = block_sum(a_lo, red);
  a_hi = block_sum(a_hi, red);
  if (threadIdx.x == 0) {
    g_offset[0] += a_state; g_offset[1] += a_state;
    g_mult[0] += a_m0; g_mult[1] += a_m1;
    g_min[0] += a_lo * lo * (1.f - lo);
    g_max[0] += a_hi * hi * (1.f - hi);
  }
}

// ============================================================================ causal context assembly
// Reference: edm2/conv.py:68-69,78-84 (causal_pad / cache['activations'] + torch.cat with the clean frames).
// ctx[b, 0:2] = pad[b] (or ones on the first `cin` channels when pad == nullptr), ctx[b, 2+t] = x[(b*S + 0)*T + t].
__global__ void __launch_bounds__(256) ctx_build_kernel(const __nv_bfloat16* __restrict__ x,
                                                        const __nv_bfloat16* __restrict__ pad,
                                                        __nv_bfloat16* __restrict__ ctx, int S, int T, long frame_elems,
                                                        int cin, int cin_pad, long total_vec) {
  pdl_launch_dependents();
  pdl_wait();
  const long v = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (v >= total_vec) return;
  const long e = v * 8;
  const long per_b = static_cast<long>(T + 2) * frame_elems;
  const long b = e / per_b;
  const long r = e - b * per_b;
  const long t = r / frame_elems;        // 0..T+1
  const long off = r - t * frame_elems;
  bf16x8 out;
  if (t >= 2) {
    out = *reinterpret_cast<const bf16x8*>(x + ((b * S) * T + (t - 2)) * frame_elems + off);
  } else if (pad != nullptr) {
    out = *reinterpret_cast<const bf16x8*>(pad + (b * 2 + t) * frame_elems + off);
  } else {
    float f[8];
    const int c0 = static_cast<int>(off % cin_pad);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = (c0 + j < cin) ? 1.f : 0.f;
    out = pack8(f);
  }
  *reinterpret_cast<bf16x8*>(ctx + e) = out;
}

// Forward prologue of the gated conv in ONE launch: blocks [0, gridDim.x-1) assemble the causal context (as
// ctx_build_kernel), the last block evaluates the gate (as gate_fwd_kernel) and zeroes the scratch that the backward
// pre-pass accumulates its per-frame inner products into.
__global__ void __launch_bounds__(256) conv_prologue_kernel(const __nv_bfloat16* __restrict__ x,
                                                            const __nv_bfloat16* __restrict__ pad,
                                                            __nv_bfloat16* __restrict__ ctx, int S, int T, long frame_elems,
                                                            int cin, int cin_pad, long total_vec,
                                                            const float* __restrict__ offset, const float* __restrict__ mult,
                                                            const float* __restrict__ max_g, const float* __restrict__ min_g,
                                                            const float* __restrict__ c_noise, float* __restrict__ alpha,
                                                            float* __restrict__ beta, float* __restrict__ scratch,
                                                            int scratch_n, int frames, int n_ctx, long pad_bstride,
                                                            const int* __restrict__ n_ctx_dev) {
  pdl_launch_dependents();
  pdl_wait();
  if (blockIdx.x == gridDim.x - 1) {
    if (n_ctx_dev != nullptr) n_ctx += n_ctx_dev[0];   // context frames counted on the device (graph-replayed decode)
    const float lo = 1.f / (1.f + expf(-min_g[0])), hi = 1.f / (1.f + expf(-max_g[0]));
    const int Trow = S * T;
    for (int f = threadIdx.x; f < frames; f += blockDim.x) {
      const float pos = log1pf(static_cast<float>((f % Trow) % T + n_ctx));
      const float state = c_noise[f] * mult[0] + offset[0] + pos * mult[1] + offset[1];
      const float g = lo + (1.f - lo) * hi / (1.f + expf(-state));
      const float inv = rsqrtf((1.f - g) * (1.f - g) + g * g);
      alpha[f] = (1.f - g) * inv;
      beta[f] = g * inv;
    }
    if (scratch != nullptr)
      for (int i = threadIdx.x; i < scratch_n; i += blockDim.x) scratch[i] = 0.f;
    return;
  }
  const long v = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (v >= total_vec) return;
  const long e = v * 8;
  const long per_b = static_cast<long>(T + 2) * frame_elems;
  const long b = e / per_b;
  const long r = e - b * per_b;
  const long t = r / frame_elems;
  const long off = r - t * frame_elems;
  bf16x8 out;
  if (t >= 2) {
    out = *reinterpret_cast<const bf16x8*>(x + ((b * S) * T + (t - 2)) * frame_elems + off);
  } else if (pad != nullptr) {
    out = *reinterpret_cast<const bf16x8*>(pad + b * pad_bstride + t * frame_elems + off);
  } else {
    float f[8];
    const int c0 = static_cast<int>(off % cin_pad);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = (c0 + j < cin) ? 1.f : 0.f;
    out = pack8(f);
  }
  *reinterpret_cast<bf16x8*>(ctx + e) = out;
}

// ============================================================================ pixel norm (+ mp_silu)
// Reference: edm2/networks_edm2.py:70 normalize(x, dim=1) followed by mp_silu (edm2/utils.py:112-113).
// One warp per pixel row of C channels. Writes xn (the residual stream) and, optionally, act = mp_silu(xn).
// mode 0: xn = normalize(x), act = mp_silu(xn).  mode 1 (decoder blocks): no norm, act = mp_silu(x) only.
template <int MODE>
__global__ void __launch_bounds__(256) pixnorm_silu_fwd_kernel(const __nv_bfloat16* __restrict__ x,
                                                               __nv_bfloat16* __restrict__ xn,
                                                               __nv_bfloat16* __restrict__ act, long rows, int C,
                                                               float eps) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long row = (static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  const __nv_bfloat16* xr = x + row * C;
  float inv = 1.f;
  if (MODE == 0) {
    float ss = 0.f;
    for (int c = lane * 8; c < C; c += 256) {
      float f[8];
      unpack8(*reinterpret_cast<const bf16x8*>(xr + c), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) ss += f[j] * f[j];
    }
    ss = warp_sum(ss);
    inv = 1.f / (eps + sqrtf(ss / C));
  }
  for (int c = lane * 8; c < C; c += 256) {
    float f[8], a[8];
    unpack8(*reinterpret_cast<const bf16x8*>(xr + c), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      f[j] *= inv;
      a[j] = f[j] / (1.f + __expf(-f[j])) * (1.f / 0.596f);
    }
    if (MODE == 0) *reinterpret_cast<bf16x8*>(xn + row * C + c) = pack8(f);
    *reinterpret_cast<bf16x8*>(act + row * C + c) = pack8(a);
  }
}

__device__ __forceinline__ float mp_silu_grad(float v) {
  const float sg = 1.f / (1.f + __expf(-v));
  return sg * (1.f + v * (1.f - sg)) * (1.f / 0.596f);
}

// Backward: dx = d(normalize)(g_xn + g_act * silu'(xn)) ; mode 1: dx = g_x + g_act * silu'(x).
//   y = x/d, d = eps + rms:  dx = g/d - y * <g,y> / (C*rms)   (per pixel row)
template <int MODE>
__global__ void __launch_bounds__(256) pixnorm_silu_bwd_kernel(const __nv_bfloat16* __restrict__ x,
                                                               const __nv_bfloat16* __restrict__ g_xn,
                                                               const __nv_bfloat16* __restrict__ g_act,
                                                               __nv_bfloat16* __restrict__ dx, long rows, int C,
                                                               float eps) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long row = (static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  const __nv_bfloat16* xr = x + row * C;
  float inv = 1.f, rms = 1.f, dot = 0.f;
  if (MODE == 0) {
    float ss = 0.f;
    for (int c = lane * 8; c < C; c += 256) {
      float f[8];
      unpack8(*reinterpret_cast<const bf16x8*>(xr + c), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) ss += f[j] * f[j];
    }
    ss = warp_sum(ss);
    rms = sqrtf(ss / C);
    inv = 1.f / (eps + rms);
    for (int c = lane * 8; c < C; c += 256) {
      float f[8], g1[8], g2[8];
      unpack8(*reinterpret_cast<const bf16x8*>(xr + c), f);
      unpack8(*reinterpret_cast<const bf16x8*>(g_act + row * C + c), g2);
      if (g_xn) unpack8(*reinterpret_cast<const bf16x8*>(g_xn + row * C + c), g1);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float yv = f[j] * inv;
        const float g = (g_xn ? g1[j] : 0.f) + g2[j] * mp_silu_grad(yv);
        dot += g * yv;
      }
    }
    dot = warp_sum(dot);
  }
  const float proj = (MODE == 0 && rms > 0.f) ? dot / (C * rms) : 0.f;
  for (int c = lane * 8; c < C; c += 256) {
    float f[8], g1[8], g2[8], o[8];
    unpack8(*reinterpret_cast<const bf16x8*>(xr + c), f);
    unpack8(*reinterpret_cast<const bf16x8*>(g_act + row * C + c), g2);
    if (g_xn) unpack8(*reinterpret_cast<const bf16x8*>(g_xn + row * C + c), g1);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float yv = f[j] * inv;
      const float g = (g_xn ? g1[j] : 0.f) + g2[j] * mp_silu_grad(yv);
      o[j] = (MODE == 0) ? (g * inv - yv * proj) : g;
    }
    *reinterpret_cast<bf16x8*>(dx + row * C + c) = pack8(o);
  }
}

// ============================================================================ emb-scale + mp_silu
// Reference: edm2/networks_edm2.py:75-77:  y = mp_silu(y * c[frame, channel])   (c = emb_linear(emb)+1).
__global__ void __launch_bounds__(256) scale_silu_fwd_kernel(const __nv_bfloat16* __restrict__ y,
                                                             const float* __restrict__ cscale,
                                                             __nv_bfloat16* __restrict__ out, long rows, int C, int ld,
                                                             int rows_per_frame) {
  pdl_launch_dependents();
  pdl_wait();
  const long vec = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int vec_per_row = C >> 3;
  if (vec >= rows * vec_per_row) return;
  const long row = vec / vec_per_row;
  const int c = static_cast<int>(vec - row * vec_per_row) << 3;
  const long frame = row / rows_per_frame;
  float f[8], o[8];
  unpack8(*reinterpret_cast<const bf16x8*>(y + row * C + c), f);
  const float4 s0 = *reinterpret_cast<const float4*>(cscale + frame * ld + c);
  const float4 s1 = *reinterpret_cast<const float4*>(cscale + frame * ld + c + 4);
  const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float z = f[j] * sc[j];
    o[j] = z / (1.f + __expf(-z)) * (1.f / 0.596f);
  }
  *reinterpret_cast<bf16x8*>(out + row * C + c) = pack8(o);
}

// Backward: dy = g * silu'(y*c) * c ;  dc[frame, ch] = sum over the frame's pixels of g * silu'(y*c) * y.
// grid = (channel groups, pixel chunks, frames): enough CTAs to saturate HBM; each CTA reduces its pixels in registers
// and shared memory, then adds one partial per channel into dc (zeroed by the launcher).
__global__ void __launch_bounds__(256) scale_silu_bwd_kernel(const __nv_bfloat16* __restrict__ y,
                                                             const float* __restrict__ cscale,
                                                             const __nv_bfloat16* __restrict__ g,
                                                             __nv_bfloat16* __restrict__ dy, float* __restrict__ dc, int C, int ld,
                                                             int rows_per_frame, int rows_per_chunk) {
  pdl_launch_dependents();
  pdl_wait();
  const int frame = blockIdx.z;
  const int cv_per_blk = min(C >> 3, 32);
  const int pl = blockDim.x / cv_per_blk;
  const int cv = threadIdx.x % cv_per_blk, p0 = threadIdx.x / cv_per_blk;
  const int c = (blockIdx.x * cv_per_blk + cv) << 3;
  const int r_lo = blockIdx.y * rows_per_chunk;
  const int r_hi = min(r_lo + rows_per_chunk, rows_per_frame);
  __shared__ float part[256][9];
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (c < C) {
    const float4 s0 = *reinterpret_cast<const float4*>(cscale + static_cast<long>(frame) * ld + c);
    const float4 s1 = *reinterpret_cast<const float4*>(cscale + static_cast<long>(frame) * ld + c + 4);
    const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
    for (int p = r_lo + p0; p < r_hi; p += pl) {
      const long off = (static_cast<long>(frame) * rows_per_frame + p) * C + c;
      float yv[8], gv[8], o[8];
      unpack8(*reinterpret_cast<const bf16x8*>(y + off), yv);
      unpack8(*reinterpret_cast<const bf16x8*>(g + off), gv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float t = gv[j] * mp_silu_grad(yv[j] * sc[j]);
        o[j] = t * sc[j];
        acc[j] += t * yv[j];
      }
      *reinterpret_cast<bf16x8*>(dy + off) = pack8(o);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) part[threadIdx.x][j] = acc[j];
  __syncthreads();
  if (p0 == 0 && c < C) {
    for (int q = 1; q < pl; ++q)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += part[q * cv_per_blk + cv][j];
    float* o = dc + static_cast<long>(frame) * C + c;
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(o + j, acc[j]);
  }
}

// ============================================================================ mp_sum (+ clip)
// Reference: edm2/utils.py:118-123 with float t, then edm2/networks_edm2.py:93 clip_(+-clip).
//   out = clamp((a*(1-t) + b*t) / sqrt((1-t)^2+t^2), +-clip)     (clip <= 0: no clamp)
__global__ void __launch_bounds__(256) mp_sum_fwd_kernel(const __nv_bfloat16* __restrict__ a,
                                                         const __nv_bfloat16* __restrict__ b,
                                                         __nv_bfloat16* __restrict__ out, long n8, float wa, float wb,
                                                         float clip) {
  pdl_launch_dependents();
  pdl_wait();
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  float fa[8], fb[8], o[8];
  unpack8(reinterpret_cast<const bf16x8*>(a)[i], fa);
  unpack8(reinterpret_cast<const bf16x8*>(b)[i], fb);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float v = fa[j] * wa + fb[j] * wb;
    if (clip > 0.f) v = fminf(fmaxf(v, -clip), clip);
    o[j] = v;
  }
  reinterpret_cast<bf16x8*>(out)[i] = pack8(o);
}
// Backward: gradients pass where the (saved) output is strictly inside the clip range (torch clamp semantics: <=).
__global__ void __launch_bounds__(256) mp_sum_bwd_kernel(const __nv_bfloat16* __restrict__ g,
                                                         const __nv_bfloat16* __restrict__ out,
                                                         __nv_bfloat16* __restrict__ da, __nv_bfloat16* __restrict__ db,
                                                         long n8, float wa, float wb, float clip) {
  pdl_launch_dependents();
  pdl_wait();
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  float fg[8], fo[8], oa[8], ob_[8];
  unpack8(reinterpret_cast<const bf16x8*>(g)[i], fg);
  if (clip > 0.f) unpack8(reinterpret_cast<const bf16x8*>(out)[i], fo);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float v = fg[j];
    if (clip > 0.f && (fo[j] >= clip || fo[j] <= -clip)) v = 0.f;
    oa[j] = v * wa;
    ob_[j] = v * wb;
  }
  reinterpret_cast<bf16x8*>(da)[i] = pack8(oa);
  reinterpret_cast<bf16x8*>(db)[i] = pack8(ob_);
}

// ============================================================================ host launchers
static int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s launch: %s", what, cudaGetErrorString(e));
    return OB_ERR_CUDA;
  }
  return OB_OK;
}

int wnorm_fwd(float* w, void* wg, int Co, int Ci, int taps, int Ci_pad, int taps_total, int tap_off, float gain, float eps,
              int training, cudaStream_t st) {
  if (Co <= 0) return OB_OK;
  launch(wnorm_fwd_kernel, Co, 256, 0, st, 1, w, static_cast<__nv_bfloat16*>(wg), Ci, taps, Ci_pad, taps_total, tap_off, gain,
                                       eps, training);
  return check_launch("wnorm_fwd");
}

int wnorm_fwd_multi(const void* jobs, const int* row_start, int n_jobs, int total_rows, float eps, cudaStream_t st) {
  if (n_jobs <= 0 || total_rows <= 0) return OB_OK;
  launch(wnorm_fwd_multi_kernel, total_rows, 256, 0, st, 1, static_cast<const ob_wnorm_job*>(jobs), row_start, n_jobs, eps);
  return check_launch("wnorm_fwd_multi");
}

int wnorm_bwd2(const float* w0, float* dw0, int taps0, int tap_off0, float gain0, const float* w1, float* dw1, int taps1,
               int tap_off1, float gain1, const float* dwg, int Co, int Ci, int Ci_pad, int taps_total, int n_split, float eps,
               int accumulate, cudaStream_t st) {
  if (Co <= 0) return OB_OK;
  WnormBwdArgs a;
  a.part[0] = WnormBwdPart{w0, dw0, taps0, tap_off0, gain0};
  a.part[1] = WnormBwdPart{w1, dw1, taps1, tap_off1, gain1};
  a.dwg = dwg; a.Co = Co; a.Ci = Ci; a.Ci_pad = Ci_pad; a.taps_total = taps_total; a.n_split = n_split;
  a.accumulate = accumulate; a.eps = eps;
  launch(wnorm_bwd_kernel, dim3(Co, w1 != nullptr ? 2 : 1), 256, 0, st, 1, a);
  return check_launch("wnorm_bwd");
}

int wnorm_bwd(const float* w, const float* dwg, float* dw, int Co, int Ci, int taps, int Ci_pad, int taps_total,
              int tap_off, int n_split, float gain, float eps, int accumulate, cudaStream_t st) {
  return wnorm_bwd2(w, dw, taps, tap_off, gain, nullptr, nullptr, 0, 0, 0.f, dwg, Co, Ci, Ci_pad, taps_total, n_split, eps,
                    accumulate, st);
}

int gate_bwd(const void* dy, const void* y, const void* d, const float* alpha, const float* beta, void* gya, void* gb,
             float* s_y, float* s_d, int n_seq, int S, int T, long frame_elems, const float* offset, const float* mult,
             const float* max_g, const float* min_g, const float* c_noise, float* g_offset, float* g_mult, float* g_max,
             float* g_min, unsigned* counter, int n_ctx, cudaStream_t st) {
  if (frame_elems % 8 != 0 || S < 1 || S > 2) {
    set_error("gate_bwd: frame size %ld must be a multiple of 8 and S in {1,2}", frame_elems);
    return OB_ERR_INVALID;
  }
  if (n_seq * T <= 0) return OB_OK;
  // enough CTAs to fill the GPU, but several 16-byte vectors per thread so a CTA is not all reduction tail
  int bx = static_cast<int>((frame_elems / 8 + 255) / 256);
  while (bx > 1 && static_cast<long>(bx) * n_seq * T > 4 * 148) bx = (bx + 1) / 2;
  dim3 grid(bx, n_seq * T);
  launch(gate_bwd_kernel, grid, 256, 0, st, 1, static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(y),
                                        static_cast<const __half*>(d), alpha, beta,
                                        static_cast<__nv_bfloat16*>(gya), static_cast<__nv_bfloat16*>(gb), s_y, s_d, n_seq,
                                        S, T, frame_elems,
                                        GateGradArgs{offset, mult, max_g, min_g, c_noise, g_offset, g_mult, g_max, g_min,
                                                     (offset != nullptr) ? counter : nullptr, n_ctx});
  return check_launch("gate_bwd");
}

int gate_fwd(const float* offset, const float* mult, const float* max_g, const float* min_g, const float* c_noise,
             float* alpha, float* beta, int frames, int T, int half, int n_ctx, cudaStream_t st) {
  if (frames <= 0) return OB_OK;
  launch(gate_fwd_kernel, (frames + 127) / 128, 128, 0, st, 1, offset, mult, max_g, min_g, c_noise, alpha, beta, frames, T, half, n_ctx);
  return check_launch("gate_fwd");
}

int gate_bwd_params(const float* offset, const float* mult, const float* max_g, const float* min_g, const float* c_noise,
                    const float* alpha, const float* beta, const float* s_y, const float* s_d, float* g_offset,
                    float* g_mult, float* g_max, float* g_min, int frames, int T, int half, int n_ctx, cudaStream_t st) {
  if (frames <= 0) return OB_OK;
  launch(gate_bwd_params_kernel, 1, 128, 0, st, 1, offset, mult, max_g, min_g, c_noise, alpha, beta, s_y, s_d, g_offset, g_mult,
                                            g_max, g_min, frames, T, half, n_ctx);
  return check_launch("gate_bwd_params");
}

int ctx_build(const void* x, const void* pad, void* ctx, int B, int S, int T, long frame_elems, int cin, int cin_pad,
              cudaStream_t st) {
  if (frame_elems % 8 != 0 || cin_pad % 8 != 0) { set_error("ctx_build: frame size must be a multiple of 8"); return OB_ERR_INVALID; }
  const long total_vec = static_cast<long>(B) * (T + 2) * frame_elems / 8;
  if (total_vec <= 0) return OB_OK;
  launch(ctx_build_kernel, (total_vec + 255) / 256, 256, 0, st, 1, static_cast<const __nv_bfloat16*>(x),
                                                            static_cast<const __nv_bfloat16*>(pad),
                                                            static_cast<__nv_bfloat16*>(ctx), S, T, frame_elems, cin, cin_pad,
                                                            total_vec);
  return check_launch("ctx_build");
}

int conv_prologue(const void* x, const void* pad, void* ctx, int B, int S, int T, long frame_elems, int cin, int cin_pad,
                  const float* offset, const float* mult, const float* max_g, const float* min_g, const float* c_noise,
                  float* alpha, float* beta, float* scratch, int scratch_n, int n_ctx, long pad_bstride, const int* n_ctx_dev,
                  cudaStream_t st) {
  if (frame_elems % 8 != 0 || cin_pad % 8 != 0) { set_error("conv_prologue: frame size must be a multiple of 8"); return OB_ERR_INVALID; }
  if (pad_bstride == 0) pad_bstride = 2 * frame_elems;
  if (pad_bstride % 8 != 0 || pad_bstride < 2 * frame_elems) { set_error("conv_prologue: bad pad batch stride %ld", pad_bstride); return OB_ERR_INVALID; }
  const long total_vec = static_cast<long>(B) * (T + 2) * frame_elems / 8;
  if (total_vec <= 0) return OB_OK;
  const unsigned blocks = static_cast<unsigned>((total_vec + 255) / 256) + 1;
  launch(conv_prologue_kernel, blocks, 256, 0, st, 1, static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(pad),
                                               static_cast<__nv_bfloat16*>(ctx), S, T, frame_elems, cin, cin_pad, total_vec,
                                               offset, mult, max_g, min_g, c_noise, alpha, beta, scratch, scratch_n,
                                               B * S * T, n_ctx, pad_bstride, n_ctx_dev);
  return check_launch("conv_prologue");
}

int pixnorm_silu_fwd(const void* x, void* xn, void* act, long rows, int C, float eps, int mode, cudaStream_t st) {
  if (C % 8 != 0) { set_error("pixnorm_silu: C=%d must be a multiple of 8", C); return OB_ERR_INVALID; }
  if (rows <= 0) return OB_OK;
  const long blocks = (rows * 32 + 255) / 256;
  if (mode == 0)
    launch(pixnorm_silu_fwd_kernel<0>, blocks, 256, 0, st, 1, static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(xn),
                                                       static_cast<__nv_bfloat16*>(act), rows, C, eps);
  else
    launch(pixnorm_silu_fwd_kernel<1>, blocks, 256, 0, st, 1, static_cast<const __nv_bfloat16*>(x), nullptr,
                                                       static_cast<__nv_bfloat16*>(act), rows, C, eps);
  return check_launch("pixnorm_silu_fwd");
}

int pixnorm_silu_bwd(const void* x, const void* g_xn, const void* g_act, void* dx, long rows, int C, float eps, int mode,
                     cudaStream_t st) {
  if (C % 8 != 0) { set_error("pixnorm_silu: C=%d must be a multiple of 8", C); return OB_ERR_INVALID; }
  if (rows <= 0) return OB_OK;
  const long blocks = (rows * 32 + 255) / 256;
  if (mode == 0)
    launch(pixnorm_silu_bwd_kernel<0>, blocks, 256, 0, st, 1, static_cast<const __nv_bfloat16*>(x),
                                                       static_cast<const __nv_bfloat16*>(g_xn),
                                                       static_cast<const __nv_bfloat16*>(g_act),
                                                       static_cast<__nv_bfloat16*>(dx), rows, C, eps);
  else
    launch(pixnorm_silu_bwd_kernel<1>, blocks, 256, 0, st, 1, static_cast<const __nv_bfloat16*>(x),
                                                       static_cast<const __nv_bfloat16*>(g_xn),
                                                       static_cast<const __nv_bfloat16*>(g_act),
                                                       static_cast<__nv_bfloat16*>(dx), rows, C, eps);
  return check_launch("pixnorm_silu_bwd");
}

int scale_silu_fwd(const void* y, const float* cscale, void* out, long rows, int C, int rows_per_frame, int ld, cudaStream_t st) {
  if (ld == 0) ld = C;
  if (ld % 4 != 0 || ld < C || reinterpret_cast<uintptr_t>(cscale) % 16 != 0) { set_error("scale_silu: scale rows must be 16-byte aligned (ld=%d)", ld); return OB_ERR_INVALID; }
  if (C % 8 != 0) { set_error("scale_silu: C=%d must be a multiple of 8", C); return OB_ERR_INVALID; }
  if (rows <= 0) return OB_OK;
  const long n = rows * (C / 8);
  launch(scale_silu_fwd_kernel, (n + 255) / 256, 256, 0, st, 1, static_cast<const __nv_bfloat16*>(y), cscale,
                                                         static_cast<__nv_bfloat16*>(out), rows, C, ld, rows_per_frame);
  return check_launch("scale_silu_fwd");
}

int scale_silu_bwd(const void* y, const float* cscale, const void* g, void* dy, float* dc, int frames, int C,
                   int rows_per_frame, int ld, cudaStream_t st) {
  if (ld == 0) ld = C;
  if (ld % 4 != 0 || ld < C || reinterpret_cast<uintptr_t>(cscale) % 16 != 0) { set_error("scale_silu: scale rows must be 16-byte aligned (ld=%d)", ld); return OB_ERR_INVALID; }
  if (C % 8 != 0) { set_error("scale_silu: C=%d must be a multiple of 8", C); return OB_ERR_INVALID; }
  if (frames <= 0) return OB_OK;
  const int cv = C / 8;
  const int cv_per_blk = cv < 32 ? cv : 32;
  if (256 % cv_per_blk != 0) { set_error("scale_silu_bwd: C/8=%d must divide 256 or be >=32", cv); return OB_ERR_UNSUPPORTED; }
  const int pl = 256 / cv_per_blk;                       // pixel lanes per CTA
  const int cgroups = (cv + cv_per_blk - 1) / cv_per_blk;
  int chunks = (4 * 148 + cgroups * frames - 1) / (cgroups * frames);   // aim for ~4 CTAs per SM
  const int max_chunks = (rows_per_frame + pl - 1) / pl;
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  const int rows_per_chunk = (rows_per_frame + chunks - 1) / chunks;
  cudaMemsetAsync(dc, 0, static_cast<size_t>(frames) * C * sizeof(float), st);
  dim3 grid(cgroups, chunks, frames);
  launch(scale_silu_bwd_kernel, grid, 256, 0, st, 1, static_cast<const __nv_bfloat16*>(y), cscale,
                                              static_cast<const __nv_bfloat16*>(g), static_cast<__nv_bfloat16*>(dy), dc, C, ld,
                                              rows_per_frame, rows_per_chunk);
  return check_launch("scale_silu_bwd");
}

int mp_sum_fwd(const void* a, const void* b, void* out, long n, float t, float clip, cudaStream_t st) {
  if (n % 8 != 0) { set_error("mp_sum: element count %ld must be a multiple of 8", n); return OB_ERR_INVALID; }
  if (n <= 0) return OB_OK;
  const float nrm = 1.f / sqrtf((1.f - t) * (1.f - t) + t * t);
  launch(mp_sum_fwd_kernel, (n / 8 + 255) / 256, 256, 0, st, 1, static_cast<const __nv_bfloat16*>(a),
                                                         static_cast<const __nv_bfloat16*>(b),
                                                         static_cast<__nv_bfloat16*>(out), n / 8, (1.f - t) * nrm, t * nrm,
                                                         clip);
  return check_launch("mp_sum_fwd");
}

int mp_sum_bwd(const void* g, const void* out, void* da, void* db, long n, float t, float clip, cudaStream_t st) {
  if (n % 8 != 0) { set_error("mp_sum: element count %ld must be a multiple of 8", n); return OB_ERR_INVALID; }
  if (n <= 0) return OB_OK;
  const float nrm = 1.f / sqrtf((1.f - t) * (1.f - t) + t * t);
  launch(mp_sum_bwd_kernel, (n / 8 + 255) / 256, 256, 0, st, 1, static_cast<const __nv_bfloat16*>(g),
                                                         static_cast<const __nv_bfloat16*>(out),
                                                         static_cast<__nv_bfloat16*>(da), static_cast<__nv_bfloat16*>(db),
                                                         n / 8, (1.f - t) * nrm, t * nrm, clip);
  return check_launch("mp_sum_bwd");
}

// ============================================================================ mp_cat
// Reference: edm2/utils.py:128-134 (decoder skip connections, edm2/networks_edm2.py:244): channel concatenation with
// magnitude-preserving weights, out[row] = [a[row]*wa | b[row]*wb] over NHWC rows.  One pass instead of two scaled
// copies and a cat; the backward splits and scales the gradient the same way.  DIR 0: forward, 1: backward.
template <int DIR>
__global__ void __launch_bounds__(256) mp_cat_kernel(__nv_bfloat16* __restrict__ a, __nv_bfloat16* __restrict__ b,
                                                     __nv_bfloat16* __restrict__ cat, long rows, int ca8, int cb8, float wa,
                                                     float wb) {
  pdl_launch_dependents();
  pdl_wait();
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int c8 = ca8 + cb8;
  if (i >= rows * c8) return;
  const long row = i / c8;
  const int v = static_cast<int>(i - row * c8);
  const bool first = v < ca8;
  bf16x8* side = first ? reinterpret_cast<bf16x8*>(a) + row * ca8 + v : reinterpret_cast<bf16x8*>(b) + row * cb8 + (v - ca8);
  bf16x8* joint = reinterpret_cast<bf16x8*>(cat) + i;
  const float w = first ? wa : wb;
  float f[8];
  unpack8(DIR == 0 ? *side : *joint, f);
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] *= w;
  if (DIR == 0) *joint = pack8(f);
  else *side = pack8(f);
}

int mp_cat(void* a, void* b, void* cat, long rows, int ca, int cb, float t, int backward, cudaStream_t st) {
  if (ca % 8 != 0 || cb % 8 != 0) { set_error("mp_cat: channel counts %d, %d must be multiples of 8", ca, cb); return OB_ERR_INVALID; }
  if (rows <= 0) return OB_OK;
  const float c = sqrtf(static_cast<float>(ca + cb) / ((1.f - t) * (1.f - t) + t * t));
  const float wa = c / sqrtf(static_cast<float>(ca)) * (1.f - t), wb = c / sqrtf(static_cast<float>(cb)) * t;
  const long n = rows * ((ca + cb) / 8);
  const unsigned blocks = static_cast<unsigned>((n + 255) / 256);
  if (backward)
    launch(mp_cat_kernel<1>, blocks, 256, 0, st, 1, static_cast<__nv_bfloat16*>(a), static_cast<__nv_bfloat16*>(b),
                                             static_cast<__nv_bfloat16*>(cat), rows, ca / 8, cb / 8, wa, wb);
  else
    launch(mp_cat_kernel<0>, blocks, 256, 0, st, 1, static_cast<__nv_bfloat16*>(a), static_cast<__nv_bfloat16*>(b),
                                             static_cast<__nv_bfloat16*>(cat), rows, ca / 8, cb / 8, wa, wb);
  return check_launch("mp_cat");
}

// ============================================================================ VAE ResBlock activation
// Reference: edm2/vae/vae.py:77-83 and :86-87 --  y = x / sqrt(mean_c(x^2) + eps);  [y = y*(1+scale_b) + shift_b  (the
// decoder's FiLM from t_cond)];  out = silu(y).  The reference runs it as ~7 eager kernels forward and ~15 backward over
// the largest activations of the VAE (up to 16x256x256x32); here one pass each way, fp32 arithmetic, bf16 in / out, the
// input recomputed (not saved) in the backward.  One warp per pixel row of C channels (C % 8 == 0, C <= 1024); a CTA's
// rows all belong to one batch element (blockIdx.y) so the FiLM gradients reduce in registers -> shared memory -> one
// atomicAdd per (CTA, channel).  film: fp32 [B][2C] = (scale | shift), or nullptr.
constexpr int VNS_MAXV = 4;   // 8-channel vectors per lane
__device__ __forceinline__ float silu_f(float y) { return y / (1.f + __expf(-y)); }
// LPR lanes share one row (LPR = min(32, C/8) rounded up to a power of two): a warp works on 32/LPR rows at once, so
// narrow rows (the VAE's 32-channel level is 64 bytes per pixel) still fill every lane and every 128-byte line.
template <int LPR>
__device__ __forceinline__ float row_sum(float v) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int LPR>
__global__ void __launch_bounds__(256) vae_norm_silu_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ film,
                                                                __nv_bfloat16* __restrict__ out, long rows_per_batch, int C,
                                                                int c_mean, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int RPW = 32 / LPR;                   // rows per warp and iteration
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane % LPR, rsel = lane / LPR;
  const int b = blockIdx.y;
  const float* fs = film ? film + static_cast<long>(b) * 2 * C : nullptr;
  for (long r0 = (static_cast<long>(blockIdx.x) * 8 + warp) * RPW; r0 < rows_per_batch; r0 += static_cast<long>(gridDim.x) * 8 * RPW) {
    const long r = r0 + rsel;
    const bool live = r < rows_per_batch;
    const long off = (static_cast<long>(b) * rows_per_batch + (live ? r : 0)) * C;
    float v[VNS_MAXV][8];
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < VNS_MAXV; ++k) {
      const int c = (sub + k * LPR) * 8;
      if (c < C && live) {
        unpack8(*reinterpret_cast<const bf16x8*>(x + off + c), v[k]);
#pragma unroll
        for (int j = 0; j < 8; ++j) ss += v[k][j] * v[k][j];
      }
    }
    const float inv = rsqrtf(row_sum<LPR>(ss) / c_mean + eps);
#pragma unroll
    for (int k = 0; k < VNS_MAXV; ++k) {
      const int c = (sub + k * LPR) * 8;
      if (c < C && live) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float y = v[k][j] * inv;
          if (fs) y = y * (1.f + fs[c + j]) + fs[C + c + j];
          o[j] = silu_f(y);
        }
        *reinterpret_cast<bf16x8*>(out + off + c) = pack8(o);
      }
    }
  }
}

template <int LPR>
__global__ void __launch_bounds__(256, 2) vae_norm_silu_bwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ film,
                                                                const __nv_bfloat16* __restrict__ g, __nv_bfloat16* __restrict__ dx,
                                                                float* __restrict__ dfilm, long rows_per_batch, int C, int c_mean,
                                                                float eps) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float sh[];          // FILM: [2C] per-CTA partial sums of (dscale | dshift)
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane % LPR, rsel = lane / LPR;
  const int b = blockIdx.y;
  const float* fs = film ? film + static_cast<long>(b) * 2 * C : nullptr;
  float a_scale[VNS_MAXV][8], a_shift[VNS_MAXV][8];
  if (fs) {
#pragma unroll
    for (int k = 0; k < VNS_MAXV; ++k)
#pragma unroll
      for (int j = 0; j < 8; ++j) { a_scale[k][j] = 0.f; a_shift[k][j] = 0.f; }
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sh[i] = 0.f;
    __syncthreads();
  }
  for (long r0 = (static_cast<long>(blockIdx.x) * 8 + warp) * RPW; r0 < rows_per_batch; r0 += static_cast<long>(gridDim.x) * 8 * RPW) {
    const long r = r0 + rsel;
    const bool live = r < rows_per_batch;
    const long off = (static_cast<long>(b) * rows_per_batch + (live ? r : 0)) * C;
    float n[VNS_MAXV][8], gn[VNS_MAXV][8];
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < VNS_MAXV; ++k) {
      const int c = (sub + k * LPR) * 8;
      if (c < C && live) {
        unpack8(*reinterpret_cast<const bf16x8*>(x + off + c), n[k]);
        unpack8(*reinterpret_cast<const bf16x8*>(g + off + c), gn[k]);
#pragma unroll
        for (int j = 0; j < 8; ++j) ss += n[k][j] * n[k][j];
      }
    }
    const float inv = rsqrtf(row_sum<LPR>(ss) / c_mean + eps);
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < VNS_MAXV; ++k) {
      const int c = (sub + k * LPR) * 8;
      if (c < C && live) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float nn = n[k][j] * inv;
          float y = nn, sc = 1.f;
          if (fs) { sc = 1.f + fs[c + j]; y = nn * sc + fs[C + c + j]; }
          const float sg = 1.f / (1.f + __expf(-y));
          const float gy = gn[k][j] * sg * (1.f + y * (1.f - sg));     // through silu
          if (fs) { a_shift[k][j] += gy; a_scale[k][j] += gy * nn; }
          n[k][j] = nn;
          gn[k][j] = gy * sc;                                           // gradient w.r.t. the normalised value
          dot += gn[k][j] * nn;
        }
      }
    }
    dot = row_sum<LPR>(dot) / c_mean;
#pragma unroll
    for (int k = 0; k < VNS_MAXV; ++k) {
      const int c = (sub + k * LPR) * 8;
      if (c < C && live) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = inv * (gn[k][j] - n[k][j] * dot);
        *reinterpret_cast<bf16x8*>(dx + off + c) = pack8(o);
      }
    }
  }
  if (fs) {
#pragma unroll
    for (int k = 0; k < VNS_MAXV; ++k) {
      const int c = (sub + k * LPR) * 8;
      if (c < C) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { atomicAdd(&sh[c + j], a_scale[k][j]); atomicAdd(&sh[C + c + j], a_shift[k][j]); }
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(&dfilm[static_cast<long>(b) * 2 * C + i], sh[i]);
  }
}

static int vns_lpr(int C) {
  const int v = C / 8;
  int l = 1;
  while (l < v && l < 32) l <<= 1;
  return l;
}
static int vns_grid(long rows_per_batch, int B, int lpr) {
  const long rpb = 8L * (32 / lpr);                 // rows per CTA and iteration
  long bx = (rows_per_batch + rpb - 1) / rpb;
  const long cap = (8L * 148 + B - 1) / B;       // ~8 CTAs per SM in total
  if (bx > cap) bx = cap;
  return static_cast<int>(bx < 1 ? 1 : bx);
}

int vae_norm_silu_fwd(const void* x, const float* film, void* out, int B, long rows_per_batch, int C, int c_mean, float eps,
                      cudaStream_t st) {
  if (c_mean <= 0 || c_mean > C) c_mean = C;
  if (C % 8 != 0 || C > 256 * VNS_MAXV) { set_error("vae_norm_silu: C=%d must be a multiple of 8 and <= %d", C, 256 * VNS_MAXV); return OB_ERR_UNSUPPORTED; }
  if (B <= 0 || rows_per_batch <= 0) return OB_OK;
  const int lpr = vns_lpr(C);
  const dim3 grid(vns_grid(rows_per_batch, B, lpr), B);
#define VNS_FWD(L) launch(vae_norm_silu_fwd_kernel<L>, grid, 256, 0, st, 1, static_cast<const __nv_bfloat16*>(x), film, \
                          static_cast<__nv_bfloat16*>(out), rows_per_batch, C, c_mean, eps)
  switch (lpr) {
    case 1: VNS_FWD(1); break; case 2: VNS_FWD(2); break; case 4: VNS_FWD(4); break;
    case 8: VNS_FWD(8); break; case 16: VNS_FWD(16); break; default: VNS_FWD(32); break;
  }
#undef VNS_FWD
  return check_launch("vae_norm_silu_fwd");
}

int vae_norm_silu_bwd(const void* x, const float* film, const void* g, void* dx, float* dfilm, int B, long rows_per_batch, int C,
                      int c_mean, float eps, cudaStream_t st) {
  if (c_mean <= 0 || c_mean > C) c_mean = C;
  if (C % 8 != 0 || C > 256 * VNS_MAXV) { set_error("vae_norm_silu: C=%d must be a multiple of 8 and <= %d", C, 256 * VNS_MAXV); return OB_ERR_UNSUPPORTED; }
  if (film != nullptr && dfilm == nullptr) { set_error("vae_norm_silu_bwd: dfilm is required with film"); return OB_ERR_INVALID; }
  if (B <= 0 || rows_per_batch <= 0) return OB_OK;
  if (film != nullptr) cudaMemsetAsync(dfilm, 0, static_cast<size_t>(B) * 2 * C * sizeof(float), st);
  const int lpr = vns_lpr(C);
  const dim3 grid(vns_grid(rows_per_batch, B, lpr), B);
  const size_t shb = film ? 2 * C * sizeof(float) : 0;
#define VNS_BWD(L) launch(vae_norm_silu_bwd_kernel<L>, grid, 256, shb, st, 1, static_cast<const __nv_bfloat16*>(x), film, \
                          static_cast<const __nv_bfloat16*>(g), static_cast<__nv_bfloat16*>(dx), dfilm, rows_per_batch, C, c_mean, eps)
  switch (lpr) {
    case 1: VNS_BWD(1); break; case 2: VNS_BWD(2); break; case 4: VNS_BWD(4); break;
    case 8: VNS_BWD(8); break; case 16: VNS_BWD(16); break; default: VNS_BWD(32); break;
  }
#undef VNS_BWD
  return check_launch("vae_norm_silu_bwd");
}

// ============================================================================ VAE grouped causal conv: data movement
// Reference: edm2/vae/vae.py:40-53.  The conv3d with kernel (kt, 3, 3) and temporal stride g reads, for output group t', the
// input frames t'g - p .. t'g - p + kt - 1 (p = kt - g frames of left padding: the cache, or a copy of the first p frames).
// time_window_gather lays those kt frames side by side on the channel axis (one pass; the reference does pad + cat + the
// strided conv), time_window_scatter is its transpose (every input frame collects its <= kt/g windows), and ungroup moves
// the (g, Cc)-ordered output channels of a group into g consecutive frames ('b (c g) t h w -> b c (t g) h w').
// All tensors bf16 NHWC rows; one thread per 8 channels.
__global__ void __launch_bounds__(256) time_window_gather_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ pad,
                                                                 __nv_bfloat16* __restrict__ xs, long total8, int T, long hw, int C8,
                                                                 int g, int kt) {
  pdl_launch_dependents();
  pdl_wait();
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  const int p = kt - g, Tg = T / g;
  long r = i;
  const int c8 = static_cast<int>(r % C8); r /= C8;
  const int j = static_cast<int>(r % kt); r /= kt;
  const long px = r % hw; r /= hw;
  const int tg = static_cast<int>(r % Tg);
  const long b = r / Tg;
  const int t = tg * g + j - p;
  const bf16x8* src;
  if (t >= 0) src = reinterpret_cast<const bf16x8*>(x) + ((b * T + t) * hw + px) * C8 + c8;
  else if (pad != nullptr) src = reinterpret_cast<const bf16x8*>(pad) + ((b * p + (t + p)) * hw + px) * C8 + c8;
  else src = reinterpret_cast<const bf16x8*>(x) + ((b * T + (t + p)) * hw + px) * C8 + c8;     // first p frames stand in for the past
  reinterpret_cast<bf16x8*>(xs)[i] = *src;
}

__global__ void __launch_bounds__(256) time_window_scatter_kernel(const __nv_bfloat16* __restrict__ dxs, __nv_bfloat16* __restrict__ dx,
                                                                  long total8, int T, long hw, int C8, int g, int kt) {
  pdl_launch_dependents();
  pdl_wait();
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  const int p = kt - g, Tg = T / g;
  long r = i;
  const int c8 = static_cast<int>(r % C8); r /= C8;
  const long px = r % hw; r /= hw;
  const int t = static_cast<int>(r % T);
  const long b = r / T;
  float acc[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) acc[u] = 0.f;
  // windows t' with 0 <= j = t + p - t'g < kt   (the padding frames are detached copies: they receive nothing)
  for (int tg = (t + p) / g; tg >= 0 && t + p - tg * g < kt; --tg) {
    if (tg >= Tg) continue;
    const int j = t + p - tg * g;
    float f[8];
    unpack8(reinterpret_cast<const bf16x8*>(dxs)[(((b * Tg + tg) * hw + px) * kt + j) * C8 + c8], f);
#pragma unroll
    for (int u = 0; u < 8; ++u) acc[u] += f[u];
  }
  reinterpret_cast<bf16x8*>(dx)[i] = pack8(acc);
}

// inverse == 0: out[(f*g + r), px, c] = in[f, px, r*Cc + c];  inverse != 0: the other way round
__global__ void __launch_bounds__(256) ungroup_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, long total8,
                                                      long hw, int g, int Cc8, int inverse) {
  pdl_launch_dependents();
  pdl_wait();
  // one CTA row = one pixel of one group: its g*Cc8 vectors are contiguous in the grouped tensor and form g runs of Cc8
  // vectors, hw*Cc8 apart, in the un-grouped one (32-bit index math: 64-bit divisions dominated the first version)
  const unsigned per_px = static_cast<unsigned>(g) * Cc8;
  const long n_px = total8 / per_px;                       // f * hw pixels
  for (long p0 = static_cast<long>(blockIdx.x) * (blockDim.x / 32) + (threadIdx.x >> 5); p0 < n_px; p0 += static_cast<long>(gridDim.x) * (blockDim.x / 32)) {
    const long f = p0 / hw, px = p0 - f * hw;
    const long grouped0 = p0 * per_px;
    for (unsigned v = threadIdx.x & 31; v < per_px; v += 32) {
      const unsigned rr = v / Cc8, c8 = v - rr * Cc8;
      const long ung = ((f * g + rr) * hw + px) * Cc8 + c8;
      if (inverse) reinterpret_cast<bf16x8*>(out)[grouped0 + v] = reinterpret_cast<const bf16x8*>(in)[ung];
      else reinterpret_cast<bf16x8*>(out)[ung] = reinterpret_cast<const bf16x8*>(in)[grouped0 + v];
    }
  }
}

int time_window(const void* src, const void* pad, void* dst, int B, int T, long hw, int C, int g, int kt, int backward, cudaStream_t st) {
  if (C % 8 != 0 || g <= 0 || kt < g || T % g != 0) { set_error("time_window: C=%d %% 8, g=%d, kt=%d, T=%d", C, g, kt, T); return OB_ERR_INVALID; }
  if (B <= 0 || T <= 0) return OB_OK;
  const int C8 = C / 8;
  if (!backward) {
    const long total8 = static_cast<long>(B) * (T / g) * hw * kt * C8;
    launch(time_window_gather_kernel, static_cast<unsigned>((total8 + 255) / 256), 256, 0, st, 1, static_cast<const __nv_bfloat16*>(src),
           static_cast<const __nv_bfloat16*>(pad), static_cast<__nv_bfloat16*>(dst), total8, T, hw, C8, g, kt);
  } else {
    const long total8 = static_cast<long>(B) * T * hw * C8;
    launch(time_window_scatter_kernel, static_cast<unsigned>((total8 + 255) / 256), 256, 0, st, 1, static_cast<const __nv_bfloat16*>(src),
           static_cast<__nv_bfloat16*>(dst), total8, T, hw, C8, g, kt);
  }
  return check_launch("time_window");
}

int ungroup(const void* in, void* out, long frames, long hw, int g, int Cc, int inverse, cudaStream_t st) {
  if (Cc % 8 != 0 || g <= 0) { set_error("ungroup: Cc=%d must be a multiple of 8", Cc); return OB_ERR_INVALID; }
  const long total8 = frames * g * hw * (Cc / 8);
  if (total8 <= 0) return OB_OK;
  const long n_px = frames * hw;
  long blocks = (n_px + 7) / 8;
  if (blocks > 148 * 16) blocks = 148 * 16;
  launch(ungroup_kernel, static_cast<unsigned>(blocks), 256, 0, st, 1, static_cast<const __nv_bfloat16*>(in),
         static_cast<__nv_bfloat16*>(out), total8, hw, g, Cc / 8, inverse);
  return check_launch("ungroup");
}

// ============================================================================ column sums (bias gradient)
// db[c] = sum over rows of g[row, c]: the bias gradient of the VAE's convs (edm2/vae/vae.py: nn.Conv3d(bias=True) under
// autograd).  torch computes it as a cast to fp32 (one full copy) followed by a reduction: 12 ms of the 74 ms VAE step.
// One pass here: a thread owns 8 adjacent channels and strides over the rows, the block reduces through shared memory and
// adds its partial sums into out (zeroed by the caller).  C % 8 == 0; rows wider than 2048 channels are split over grid.y.
__global__ void __launch_bounds__(256) colsum_kernel(const __nv_bfloat16* __restrict__ g, float* __restrict__ out, long rows,
                                                     int cv /* 16-byte vectors per row */, int cvb /* of them per block */) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[256 * 8];
  const int vl = threadIdx.x % cvb;               // which 8 channels, within this block's slice of the row
  const int v = blockIdx.y * cvb + vl;
  const int lanes = blockDim.x / cvb;             // rows handled concurrently by the block
  const int r = threadIdx.x / cvb;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (r < lanes && v < cv) {
    for (long row = static_cast<long>(blockIdx.x) * lanes + r; row < rows; row += static_cast<long>(gridDim.x) * lanes) {
      const bf16x8 p = *reinterpret_cast<const bf16x8*>(g + (row * cv + v) * 8);
      float f[8];
      unpack8(p, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[threadIdx.x * 8 + j] = acc[j];
  __syncthreads();
  // thread t < cvb * 8 sums channel (t / 8 -> vector, t % 8) over the block's row lanes
  for (int t = threadIdx.x; t < cvb * 8; t += blockDim.x) {
    const int vv = t >> 3, j = t & 7;
    if (blockIdx.y * cvb + vv >= cv) continue;
    float s = 0.f;
    for (int rr = 0; rr < lanes; ++rr) s += red[(rr * cvb + vv) * 8 + j];
    atomicAdd(&out[(blockIdx.y * cvb + vv) * 8 + j], s);
  }
}

int colsum(const void* g, float* out, long rows, int C, cudaStream_t st) {
  if (C % 8 != 0 || C <= 0) { set_error("colsum: C=%d must be a positive multiple of 8", C); return OB_ERR_INVALID; }
  if (rows <= 0) return OB_OK;
  const int cv = C / 8, cvb = cv < 256 ? cv : 256, lanes = 256 / cvb, ny = (cv + cvb - 1) / cvb;
  long blocks = (rows + lanes - 1) / lanes;
  blocks = (blocks + 15) / 16;                    // >= 16 rows per thread: the atomics stay a small fraction of the work
  const long cap = (148 * 8 + ny - 1) / ny;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  launch(colsum_kernel, dim3(static_cast<unsigned>(blocks), ny), 256, 0, st, 1, static_cast<const __nv_bfloat16*>(g), out, rows, cv, cvb);
  return check_launch("colsum");
}

// ============================================================================ 2x resampling
// Reference: edm2/utils.py:94-107 with the [1,1] filter the UNet uses: 'down' = 2x2 mean, 'up' = nearest-neighbour 2x.
// Each is the other's transpose, so two kernels cover both directions of both modes:
//   pool2x2:   out[f, y, x, :] = scale * sum of the 2x2 block of in        (down fwd: 0.25;  up bwd: 1)
//   expand2x2: out[f, 2y+i, 2x+j, :] = scale * in[f, y, x, :]              (up fwd: 1;       down bwd: 0.25)
__global__ void __launch_bounds__(256) pool2x2_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                                      long n_out8, int ho, int wo, int c8, float scale) {
  pdl_launch_dependents();
  pdl_wait();
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_out8) return;
  const int v = static_cast<int>(i % c8);
  long r = i / c8;
  const int x = static_cast<int>(r % wo); r /= wo;
  const int y = static_cast<int>(r % ho);
  const long f = r / ho;
  const bf16x8* src = reinterpret_cast<const bf16x8*>(in) + ((f * (2 * ho) + 2 * y) * (2 * wo) + 2 * x) * c8 + v;
  float acc[8], t[8];
  unpack8(src[0], acc);
  unpack8(src[c8], t);
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] += t[j];
  unpack8(src[static_cast<long>(2 * wo) * c8], t);
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] += t[j];
  unpack8(src[static_cast<long>(2 * wo) * c8 + c8], t);
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = (acc[j] + t[j]) * scale;
  reinterpret_cast<bf16x8*>(out)[i] = pack8(acc);
}
__global__ void __launch_bounds__(256) expand2x2_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                                        long n_out8, int ho, int wo, int c8, float scale) {
  pdl_launch_dependents();
  pdl_wait();
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_out8) return;
  const int v = static_cast<int>(i % c8);
  long r = i / c8;
  const int x = static_cast<int>(r % wo); r /= wo;
  const int y = static_cast<int>(r % ho);
  const long f = r / ho;
  float t[8];
  unpack8(reinterpret_cast<const bf16x8*>(in)[((f * (ho / 2) + y / 2) * (wo / 2) + x / 2) * c8 + v], t);
#pragma unroll
  for (int j = 0; j < 8; ++j) t[j] *= scale;
  reinterpret_cast<bf16x8*>(out)[i] = pack8(t);
}

// h, w: spatial size of the LARGE side (even).  pool != 0: in is [f,h,w,c] -> out [f,h/2,w/2,c]; else the reverse.
int resample2x(const void* in, void* out, long frames, int h, int w, int c, int pool, float scale, cudaStream_t st) {
  if (c % 8 != 0 || h % 2 != 0 || w % 2 != 0) { set_error("resample2x: need c %% 8 == 0 and even h, w (got %d, %d, %d)", c, h, w); return OB_ERR_INVALID; }
  const int ho = pool ? h / 2 : h, wo = pool ? w / 2 : w;
  const long n = frames * ho * wo * (c / 8);
  if (n <= 0) return OB_OK;
  const unsigned blocks = static_cast<unsigned>((n + 255) / 256);
  if (pool) launch(pool2x2_kernel, blocks, 256, 0, st, 1, static_cast<const __nv_bfloat16*>(in), static_cast<__nv_bfloat16*>(out), n, ho, wo, c / 8, scale);
  else launch(expand2x2_kernel, blocks, 256, 0, st, 1, static_cast<const __nv_bfloat16*>(in), static_cast<__nv_bfloat16*>(out), n, ho, wo, c / 8, scale);
  return check_launch("resample2x");
}

// ----------------------------------------------------------------------------- optimizer
// AdamW (decoupled weight decay, bias-corrected; the update torch.optim.AdamW applies in cs_train.py:121-124) fused with
// the two power-function-free EMA copies of the weights (cs_train.py:125) and the gradient reset, over ONE flat fp32
// range: 6 streams read, 6 written, once per optimizer step.  step_lr points at {step (already incremented), lr}.
// EMA coefficient of one tracked copy.  ratio > 0: power-function EMA (edm2/phema.py:68-70, beta = (1 - t_delta/t_next)^(exp+1))
// with a = exp + 1 and t_delta/t_next = ratio / t (t = optimizer step count); ratio <= 0: constant beta = a.
__device__ __forceinline__ float ema_beta(float a, float ratio, float step) {
  if (ratio <= 0.f) return a;
  const float r = fminf(ratio / step, 1.f);
  return r >= 1.f ? 0.f : __expf(a * log1pf(-r));
}

__global__ void __launch_bounds__(256) adamw_ema_kernel(float4* __restrict__ p, float4* __restrict__ g, float4* __restrict__ m,
                                                        float4* __restrict__ v, float4* __restrict__ e1, float4* __restrict__ e2,
                                                        long n4, const float* __restrict__ opt_state, float beta1, float beta2,
                                                        float eps, float wd, float ema_a1, float ema_a2, float ema_ratio,
                                                        float grad_scale, float max_norm) {
  pdl_launch_dependents();
  pdl_wait();
  const float step = opt_state[0], lr = opt_state[1];
  if (max_norm > 0.f) {   // torch.nn.utils.clip_grad_norm_ (gym_train.py:105): coefficient from the squared norm in opt_state[2]
    const float total = sqrtf(opt_state[2]) * grad_scale;
    grad_scale *= fminf(1.f, max_norm / (total + 1e-6f));
  }
  const float bc1 = 1.f - powf(beta1, step);
  const float bc2_sqrt = sqrtf(1.f - powf(beta2, step));
  const float step_size = lr / bc1;
  const float decay = 1.f - lr * wd;
  const float w1 = 1.f - ema_beta(ema_a1, ema_ratio, step), w2 = 1.f - ema_beta(ema_a2, ema_ratio, step);
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += static_cast<long>(gridDim.x) * blockDim.x) {
    float4 pv = p[i], mv = m[i], vv = v[i];
    float4 gv = g[i];
    gv.x *= grad_scale; gv.y *= grad_scale; gv.z *= grad_scale; gv.w *= grad_scale;   // 1/world after a summing all-reduce
    float* pp = &pv.x; float* mm = &mv.x; float* vp = &vv.x; const float* gg = &gv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      mm[k] = beta1 * mm[k] + (1.f - beta1) * gg[k];
      vp[k] = beta2 * vp[k] + (1.f - beta2) * gg[k] * gg[k];
      const float denom = sqrtf(vp[k]) / bc2_sqrt + eps;
      pp[k] = pp[k] * decay - step_size * (mm[k] / denom);
    }
    p[i] = pv; m[i] = mv; v[i] = vv;
    g[i] = zero;
    if (e1 != nullptr) {
      float4 e = e1[i]; float* ee = &e.x;
#pragma unroll
      for (int k = 0; k < 4; ++k) ee[k] += w1 * (pp[k] - ee[k]);
      e1[i] = e;
    }
    if (e2 != nullptr) {
      float4 e = e2[i]; float* ee = &e.x;
#pragma unroll
      for (int k = 0; k < 4; ++k) ee[k] += w2 * (pp[k] - ee[k]);
      e2[i] = e;
    }
  }
}

int adamw_ema(float* p, float* g, float* m, float* v, float* e1, float* e2, long n, const float* opt_state, float beta1,
              float beta2, float eps, float wd, float ema_a1, float ema_a2, float ema_ratio, float grad_scale, float max_norm,
              cudaStream_t st) {
  if (n % 4 != 0) { set_error("adamw_ema: element count %ld must be a multiple of 4", n); return OB_ERR_INVALID; }
  for (const void* q : {(const void*)p, (const void*)g, (const void*)m, (const void*)v, (const void*)e1, (const void*)e2})
    if (reinterpret_cast<uintptr_t>(q) % 16 != 0) { set_error("adamw_ema: buffers must be 16-byte aligned"); return OB_ERR_INVALID; }
  if (n <= 0) return OB_OK;
  const long n4 = n / 4;
  long blocks = (n4 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  launch(adamw_ema_kernel, static_cast<unsigned>(blocks), 256, 0, st, 1, reinterpret_cast<float4*>(p), reinterpret_cast<float4*>(g),
                                                                 reinterpret_cast<float4*>(m), reinterpret_cast<float4*>(v),
                                                                 reinterpret_cast<float4*>(e1), reinterpret_cast<float4*>(e2), n4,
                                                                 opt_state, beta1, beta2, eps, wd, ema_a1, ema_a2, ema_ratio,
                                                                 grad_scale, max_norm);
  return check_launch("adamw_ema");
}

// out[0] += sum of g[i]^2 (fp32): the squared gradient norm clip_grad_norm_ needs (gym_train.py:105).
__global__ void __launch_bounds__(256) sumsq_kernel(const float4* __restrict__ g, long n4, float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  float acc = 0.f;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const float4 v = g[i];
    acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += part[k];
    atomicAdd(out, t);
  }
}

int sumsq(const float* g, long n, float* out, cudaStream_t st) {
  if (n % 4 != 0 || reinterpret_cast<uintptr_t>(g) % 16 != 0) { set_error("sumsq: n %% 4 and 16-byte alignment required"); return OB_ERR_INVALID; }
  if (n <= 0) return OB_OK;
  const long n4 = n / 4;
  long blocks = (n4 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  launch(sumsq_kernel, static_cast<unsigned>(blocks), 256, 0, st, 1, reinterpret_cast<const float4*>(g), n4, out);
  return check_launch("sumsq");
}

}  // namespace ob
