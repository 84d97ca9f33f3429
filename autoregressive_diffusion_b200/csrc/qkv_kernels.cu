// q/k/v preparation for the attention kernels (reference: edm2/attention/attention_modules.py:48-49,59 and
// edm2/attention/RoPe.py:43-74): split the 1x1-conv output whose channels are ordered (head, c, {q,k,v}),
// RMS-normalise every 64-vector (q, k AND v), and apply the per-frame rotary + xPos tables to q and k.
// One warp per (token, head): the 192 interleaved channels of a head are one contiguous 384-byte run.
#include <cuda_bf16.h>
#include "launch.cuh"
#include <cuda_runtime.h>
#include <cstdint>

#include "hbm_host.h"
#include "tapconv_host.h"

namespace ob {

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float2 ld_bf2(const __nv_bfloat16* p) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
}
__device__ __forceinline__ void st_bf2(__nv_bfloat16* p, float a, float b) {
  *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
}

// rotate-half partner of channel pair (2l, 2l+1) lives in lane l^16; sign is - for the first half, + for the second
__device__ __forceinline__ float2 rot_half(float2 v, int lane) {
  float2 o;
  o.x = __shfl_xor_sync(0xffffffffu, v.x, 16);
  o.y = __shfl_xor_sync(0xffffffffu, v.y, 16);
  const float sgn = (lane < 16) ? -1.f : 1.f;
  o.x *= sgn; o.y *= sgn;
  return o;
}

// tables: fp32 [P][64] each (cos, sin, xPos scale) already carrying the reference's fp16 rounding.
// pos_q / pos_k: int32 per FRAME (token / hw); a negative entry means "no rotary" (FrameAttention / just_2d).
__global__ void __launch_bounds__(256) qkv_prep_fwd_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                           __nv_bfloat16* __restrict__ q, __nv_bfloat16* __restrict__ k,
                                                           __nv_bfloat16* __restrict__ v, __nv_bfloat16* __restrict__ k_raw,
                                                           const float* __restrict__ cosT, const float* __restrict__ sinT,
                                                           const float* __restrict__ sclT, const int* __restrict__ pos_q,
                                                           const int* __restrict__ pos_k, long rows, int heads, int hw,
                                                           float eps) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long wid = (static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (wid >= rows * heads) return;
  const long row = wid / heads;
  const int m = static_cast<int>(wid - row * heads);
  const __nv_bfloat16* src = qkv + (row * heads + m) * 192 + lane * 6;
  const float2 e0 = ld_bf2(src), e1 = ld_bf2(src + 2), e2 = ld_bf2(src + 4);
  // channel c0 = 2*lane: (q,k,v) = (e0.x, e0.y, e1.x); channel c1 = 2*lane+1: (e1.y, e2.x, e2.y)
  float2 qv = make_float2(e0.x, e1.y), kv = make_float2(e0.y, e2.x), vv = make_float2(e1.x, e2.y);
  const float iq = 1.f / (eps + sqrtf(wsum(qv.x * qv.x + qv.y * qv.y) * (1.f / 64.f)));
  const float ik = 1.f / (eps + sqrtf(wsum(kv.x * kv.x + kv.y * kv.y) * (1.f / 64.f)));
  const float iv = 1.f / (eps + sqrtf(wsum(vv.x * vv.x + vv.y * vv.y) * (1.f / 64.f)));
  qv.x *= iq; qv.y *= iq; kv.x *= ik; kv.y *= ik; vv.x *= iv; vv.y *= iv;
  const long dst = (row * heads + m) * 64 + lane * 2;
  if (k_raw) st_bf2(k_raw + dst, kv.x, kv.y);
  st_bf2(v + dst, vv.x, vv.y);
  const long frame = row / hw;
  const int pq = pos_q ? pos_q[frame] : -1, pk = pos_k ? pos_k[frame] : -1;
  const float2 qr = rot_half(qv, lane), kr = rot_half(kv, lane);
  if (pq >= 0) {
    const float2 c = *reinterpret_cast<const float2*>(cosT + pq * 64 + lane * 2);
    const float2 s = *reinterpret_cast<const float2*>(sinT + pq * 64 + lane * 2);
    const float2 f = *reinterpret_cast<const float2*>(sclT + pq * 64 + lane * 2);
    qv.x = (qv.x * c.x + qr.x * s.x) * f.x;
    qv.y = (qv.y * c.y + qr.y * s.y) * f.y;
  }
  if (pk >= 0) {
    const float2 c = *reinterpret_cast<const float2*>(cosT + pk * 64 + lane * 2);
    const float2 s = *reinterpret_cast<const float2*>(sinT + pk * 64 + lane * 2);
    const float2 f = *reinterpret_cast<const float2*>(sclT + pk * 64 + lane * 2);
    kv.x = (kv.x * c.x + kr.x * s.x) / f.x;
    kv.y = (kv.y * c.y + kr.y * s.y) / f.y;
  }
  st_bf2(q + dst, qv.x, qv.y);
  st_bf2(k + dst, kv.x, kv.y);
}

// Decode-step flavour of the above (one new frame per sequence): the frame's position is the sequence's cached length,
// read from DEVICE memory, and the rotated key + the value are written straight into the frame-sized page the page table
// names for that position (in place; nothing is concatenated or copied -- attention_modules.py:51-57 clones and cats the
// whole cache every evaluation).  Keys are stored ALREADY rotated with a fixed xPos centre (the tables' centre): the
// centre cancels in q.k, so cached pages never have to be re-rotated when the sequence grows.
__global__ void __launch_bounds__(256) kv_append_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ q,
                                                        __nv_bfloat16* __restrict__ k_pages, __nv_bfloat16* __restrict__ v_pages,
                                                        const int* __restrict__ page_table, const int* __restrict__ lengths,
                                                        const float* __restrict__ cosT, const float* __restrict__ sinT,
                                                        const float* __restrict__ sclT, long rows, int heads, int hw,
                                                        int max_pages, int n_pos, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long wid = (static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (wid >= rows * heads) return;
  const long row = wid / heads;
  const int m = static_cast<int>(wid - row * heads);
  const __nv_bfloat16* src = qkv + (row * heads + m) * 192 + lane * 6;
  const float2 e0 = ld_bf2(src), e1 = ld_bf2(src + 2), e2 = ld_bf2(src + 4);
  float2 qv = make_float2(e0.x, e1.y), kv = make_float2(e0.y, e2.x), vv = make_float2(e1.x, e2.y);
  const float iq = 1.f / (eps + sqrtf(wsum(qv.x * qv.x + qv.y * qv.y) * (1.f / 64.f)));
  const float ik = 1.f / (eps + sqrtf(wsum(kv.x * kv.x + kv.y * kv.y) * (1.f / 64.f)));
  const float iv = 1.f / (eps + sqrtf(wsum(vv.x * vv.x + vv.y * vv.y) * (1.f / 64.f)));
  qv.x *= iq; qv.y *= iq; kv.x *= ik; kv.y *= ik; vv.x *= iv; vv.y *= iv;
  const int b = static_cast<int>(row / hw), within = static_cast<int>(row - static_cast<long>(b) * hw);
  int pos = lengths[b];
  if (pos >= max_pages) pos = max_pages - 1;          // the host grows the pool before this can happen; never write out of bounds
  const int tp = pos < n_pos ? pos : n_pos - 1;
  const long page = page_table[static_cast<long>(b) * max_pages + pos];
  const float2 qr = rot_half(qv, lane), kr = rot_half(kv, lane);
  const float2 c = *reinterpret_cast<const float2*>(cosT + tp * 64 + lane * 2);
  const float2 s = *reinterpret_cast<const float2*>(sinT + tp * 64 + lane * 2);
  const float2 f = *reinterpret_cast<const float2*>(sclT + tp * 64 + lane * 2);
  st_bf2(q + (row * heads + m) * 64 + lane * 2, (qv.x * c.x + qr.x * s.x) * f.x, (qv.y * c.y + qr.y * s.y) * f.y);
  const long dst = ((page * hw + within) * heads + m) * 64 + lane * 2;
  st_bf2(k_pages + dst, (kv.x * c.x + kr.x * s.x) / f.x, (kv.y * c.y + kr.y * s.y) / f.y);
  st_bf2(v_pages + dst, vv.x, vv.y);
}

// Backward: undo the rotary on the incoming gradients, then the RMS-norm backward, then re-interleave.
//   y = rope(n)*f  =>  dn = f*(dy*cos - rot(dy)*sin);   n = x/d, d = eps+rms  =>  dx = dn/d - n*<dn,n>/(64*rms)
__global__ void __launch_bounds__(256) qkv_prep_bwd_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                           const __nv_bfloat16* __restrict__ dq,
                                                           const __nv_bfloat16* __restrict__ dk,
                                                           const __nv_bfloat16* __restrict__ dv,
                                                           __nv_bfloat16* __restrict__ dqkv, const float* __restrict__ cosT,
                                                           const float* __restrict__ sinT, const float* __restrict__ sclT,
                                                           const int* __restrict__ pos_q, const int* __restrict__ pos_k,
                                                           long rows, int heads, int hw, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long wid = (static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (wid >= rows * heads) return;
  const long row = wid / heads;
  const int m = static_cast<int>(wid - row * heads);
  const long so = (row * heads + m) * 192 + lane * 6;
  const float2 e0 = ld_bf2(qkv + so), e1 = ld_bf2(qkv + so + 2), e2 = ld_bf2(qkv + so + 4);
  float2 x[3] = {make_float2(e0.x, e1.y), make_float2(e0.y, e2.x), make_float2(e1.x, e2.y)};
  const long go = (row * heads + m) * 64 + lane * 2;
  float2 g[3] = {ld_bf2(dq + go), ld_bf2(dk + go), ld_bf2(dv + go)};
  const long frame = row / hw;
  const int pos[2] = {pos_q ? pos_q[frame] : -1, pos_k ? pos_k[frame] : -1};
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const float2 gr = rot_half(g[s], lane);
    if (pos[s] >= 0) {
      const float2 c = *reinterpret_cast<const float2*>(cosT + pos[s] * 64 + lane * 2);
      const float2 sn = *reinterpret_cast<const float2*>(sinT + pos[s] * 64 + lane * 2);
      float2 f = *reinterpret_cast<const float2*>(sclT + pos[s] * 64 + lane * 2);
      if (s == 1) { f.x = 1.f / f.x; f.y = 1.f / f.y; }
      g[s].x = f.x * (g[s].x * c.x - gr.x * sn.x);
      g[s].y = f.y * (g[s].y * c.y - gr.y * sn.y);
    }
  }
  float2 dx[3];
#pragma unroll
  for (int s = 0; s < 3; ++s) {
    const float rms = sqrtf(wsum(x[s].x * x[s].x + x[s].y * x[s].y) * (1.f / 64.f));
    const float inv = 1.f / (eps + rms);
    const float nx = x[s].x * inv, ny = x[s].y * inv;
    const float dot = wsum(g[s].x * nx + g[s].y * ny);
    const float proj = rms > 0.f ? dot / (64.f * rms) : 0.f;
    dx[s].x = g[s].x * inv - nx * proj;
    dx[s].y = g[s].y * inv - ny * proj;
  }
  st_bf2(dqkv + so, dx[0].x, dx[1].x);
  st_bf2(dqkv + so + 2, dx[2].x, dx[0].y);
  st_bf2(dqkv + so + 4, dx[1].y, dx[2].y);
}

// Rotary on cached (un-roped, normalised) keys: x [rows, heads*64] -> y, k-style (divide by the xPos scale).
__global__ void __launch_bounds__(256) rope_k_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                     const float* __restrict__ cosT, const float* __restrict__ sinT,
                                                     const float* __restrict__ sclT, const int* __restrict__ pos,
                                                     long rows, int heads, int hw) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long wid = (static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (wid >= rows * heads) return;
  const long row = wid / heads;
  const long off = wid * 64 + lane * 2;
  float2 v = ld_bf2(x + off);
  const float2 r = rot_half(v, lane);
  const int pk = pos[row / hw];
  const float2 c = *reinterpret_cast<const float2*>(cosT + pk * 64 + lane * 2);
  const float2 s = *reinterpret_cast<const float2*>(sinT + pk * 64 + lane * 2);
  const float2 f = *reinterpret_cast<const float2*>(sclT + pk * 64 + lane * 2);
  st_bf2(y + off, (v.x * c.x + r.x * s.x) / f.x, (v.y * c.y + r.y * s.y) / f.y);
}

static int chk(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("%s launch: %s", what, cudaGetErrorString(e)); return OB_ERR_CUDA; }
  return OB_OK;
}

int qkv_prep_fwd(const void* qkv, void* q, void* k, void* v, void* k_raw, const float* cosT, const float* sinT,
                 const float* sclT, const int* pos_q, const int* pos_k, long rows, int heads, int hw, float eps,
                 cudaStream_t st) {
  if (rows <= 0 || heads <= 0) return OB_OK;
  const long warps = rows * heads;
  launch(qkv_prep_fwd_kernel, (warps * 32 + 255) / 256, 256, 0, st, 1, 
      static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(q), static_cast<__nv_bfloat16*>(k),
      static_cast<__nv_bfloat16*>(v), static_cast<__nv_bfloat16*>(k_raw), cosT, sinT, sclT, pos_q, pos_k, rows, heads, hw, eps);
  return chk("qkv_prep_fwd");
}
int qkv_prep_bwd(const void* qkv, const void* dq, const void* dk, const void* dv, void* dqkv, const float* cosT,
                 const float* sinT, const float* sclT, const int* pos_q, const int* pos_k, long rows, int heads, int hw,
                 float eps, cudaStream_t st) {
  if (rows <= 0 || heads <= 0) return OB_OK;
  const long warps = rows * heads;
  launch(qkv_prep_bwd_kernel, (warps * 32 + 255) / 256, 256, 0, st, 1, 
      static_cast<const __nv_bfloat16*>(qkv), static_cast<const __nv_bfloat16*>(dq), static_cast<const __nv_bfloat16*>(dk),
      static_cast<const __nv_bfloat16*>(dv), static_cast<__nv_bfloat16*>(dqkv), cosT, sinT, sclT, pos_q, pos_k, rows, heads,
      hw, eps);
  return chk("qkv_prep_bwd");
}
int kv_append(const void* qkv, void* q, void* k_pages, void* v_pages, const int* page_table, const int* lengths,
              const float* cosT, const float* sinT, const float* sclT, int B, int heads, int hw, int max_pages, int n_pos,
              float eps, cudaStream_t st) {
  if (B <= 0 || heads <= 0) return OB_OK;
  if (hw <= 0 || max_pages <= 0 || n_pos <= 0) { set_error("kv_append: bad sizes hw=%d max_pages=%d n_pos=%d", hw, max_pages, n_pos); return OB_ERR_INVALID; }
  const long rows = static_cast<long>(B) * hw, warps = rows * heads;
  launch(kv_append_kernel, (warps * 32 + 255) / 256, 256, 0, st, 1, static_cast<const __nv_bfloat16*>(qkv),
         static_cast<__nv_bfloat16*>(q), static_cast<__nv_bfloat16*>(k_pages), static_cast<__nv_bfloat16*>(v_pages), page_table,
         lengths, cosT, sinT, sclT, rows, heads, hw, max_pages, n_pos, eps);
  return chk("kv_append");
}
int rope_k(const void* x, void* y, const float* cosT, const float* sinT, const float* sclT, const int* pos, long rows,
           int heads, int hw, cudaStream_t st) {
  if (rows <= 0 || heads <= 0) return OB_OK;
  const long warps = rows * heads;
  launch(rope_k_kernel, (warps * 32 + 255) / 256, 256, 0, st, 1, static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y),
                                                          cosT, sinT, sclT, pos, rows, heads, hw);
  return chk("rope_k");
}

}  // namespace ob
