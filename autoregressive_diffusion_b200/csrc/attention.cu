// Host side of the attention kernels.
#include "attention.cuh"
#include "launch.cuh"
#include "attention_bwd.cuh"

#include <cstring>

#include "tapconv_host.h"

namespace ob {

// [B, L, heads, 64] tensor seen as (channel, token, head, batch); a tile is box_rows tokens of one head.
static int make_qkv_map(CUtensorMap* m, const void* ptr, int B, int heads, int L, int box_rows = 128) {
  uint64_t dims[4] = {64, (uint64_t)L, (uint64_t)heads, (uint64_t)B};
  uint64_t str[4] = {1, (uint64_t)heads * 64, 64, (uint64_t)L * heads * 64};
  uint32_t box[4] = {64, (uint32_t)box_rows, 1, 1};
  return encode_tmap_bf16(m, ptr, 4, dims, str, box);
}

int attn_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int B, int heads, int Lq, int Lk, int hw,
             int n_frames, int mask, float scale, cudaStream_t st) {
  const int BH = B * heads;
  if (BH <= 0 || Lq <= 0 || Lk <= 0) return OB_OK;
  if (mask == ATTN_DART_LISTED && (hw >= 128 || (static_cast<long>(n_frames) * hw) % 128 != 0)) mask = ATTN_DART;  // no regrouping: lists == mask_mod
  if (hw <= 0 || mask < ATTN_FULL || mask > ATTN_DART_LISTED || (mask >= ATTN_DART && (n_frames <= 0 || Lq != Lk || Lq != 2 * n_frames * hw)) ||
      (mask == ATTN_CAUSAL && Lq != Lk)) {
    set_error("attn_fwd: inconsistent arguments (Lq=%d Lk=%d hw=%d n_frames=%d mask=%d)", Lq, Lk, hw, n_frames, mask);
    return OB_ERR_INVALID;
  }
  AttnParams p;
  memset(&p, 0, sizeof(p));
  if (int r = make_qkv_map(&p.mapQ, q, B, heads, Lq)) return r;
  if (int r = make_qkv_map(&p.mapK, k, B, heads, Lk)) return r;
  if (int r = make_qkv_map(&p.mapV, v, B, heads, Lk)) return r;
  p.BH = BH; p.heads = heads; p.Lq = Lq; p.Lk = Lk; p.hw = hw; p.n_frames = n_frames; p.mask = mask; p.scale = scale;
  p.o = static_cast<__nv_bfloat16*>(o); p.lse = lse;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATTN_SMEM_BYTES);
    if (e != cudaSuccess) { set_error("attn_fwd smem attr: %s", cudaGetErrorString(e)); return OB_ERR_CUDA; }
    attr = true;
  }
  dim3 grid((Lq + ATTN_BM - 1) / ATTN_BM, BH);
  launch(attn_fwd_kernel, grid, ATTN_THREADS, ATTN_SMEM_BYTES, st, 1, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("attn_fwd launch: %s", cudaGetErrorString(e)); return OB_ERR_CUDA; }
  return OB_OK;
}

int attn_bwd(const void* q, const void* k, const void* v, const void* o, const void* dout, const float* lse, float* dsum,
             void* dq, void* dk, void* dv, int B, int heads, int Lq, int Lk, int hw, int n_frames, int mask, float scale,
             cudaStream_t st) {
  const int BH = B * heads;
  if (BH <= 0 || Lq <= 0 || Lk <= 0) return OB_OK;
  if (mask == ATTN_DART_LISTED && (hw >= 128 || (static_cast<long>(n_frames) * hw) % 128 != 0)) mask = ATTN_DART;  // no regrouping: lists == mask_mod
  if (hw <= 0 || mask < ATTN_FULL || mask > ATTN_DART_LISTED || (mask >= ATTN_DART && (n_frames <= 0 || Lq != Lk || Lq != 2 * n_frames * hw)) ||
      (mask == ATTN_CAUSAL && Lq != Lk)) {
    set_error("attn_bwd: inconsistent arguments (Lq=%d Lk=%d hw=%d n_frames=%d mask=%d)", Lq, Lk, hw, n_frames, mask);
    return OB_ERR_INVALID;
  }
  AttnBwdParams p;
  memset(&p, 0, sizeof(p));
  if (int r = make_qkv_map(&p.mapQ128, q, B, heads, Lq, 128)) return r;
  if (int r = make_qkv_map(&p.mapdO128, dout, B, heads, Lq, 128)) return r;
  if (int r = make_qkv_map(&p.mapK64, k, B, heads, Lk, 64)) return r;
  if (int r = make_qkv_map(&p.mapV64, v, B, heads, Lk, 64)) return r;
  if (int r = make_qkv_map(&p.mapK128, k, B, heads, Lk, 128)) return r;
  if (int r = make_qkv_map(&p.mapV128, v, B, heads, Lk, 128)) return r;
  if (int r = make_qkv_map(&p.mapQ64, q, B, heads, Lq, 64)) return r;
  if (int r = make_qkv_map(&p.mapdO64, dout, B, heads, Lq, 64)) return r;
  p.BH = BH; p.heads = heads; p.Lq = Lq; p.Lk = Lk; p.hw = hw; p.n_frames = n_frames; p.mask = mask; p.scale = scale;
  p.lse = lse; p.dsum = dsum;
  p.dq = static_cast<__nv_bfloat16*>(dq); p.dk = static_cast<__nv_bfloat16*>(dk); p.dv = static_cast<__nv_bfloat16*>(dv);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ABW_DQ_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ABW_DKV_SMEM);
    if (e != cudaSuccess) { set_error("attn_bwd smem attr: %s", cudaGetErrorString(e)); return OB_ERR_CUDA; }
    attr = true;
  }
  const long rows = static_cast<long>(BH) * Lq;
  launch(attn_bwd_prep_kernel, (rows * 8 + 255) / 256, 256, 0, st, 1, static_cast<const __nv_bfloat16*>(o),
                                                               static_cast<const __nv_bfloat16*>(dout), dsum, rows, Lq, heads);
  launch(attn_bwd_dq_kernel, dim3((Lq + ABW_BM - 1) / ABW_BM, BH), ABW_THREADS, ABW_DQ_SMEM, st, 1, p);
  launch(attn_bwd_dkv_kernel, dim3((Lk + ABW_BM - 1) / ABW_BM, BH), ABW_THREADS, ABW_DKV_SMEM, st, 1, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("attn_bwd launch: %s", cudaGetErrorString(e)); return OB_ERR_CUDA; }
  return OB_OK;
}

}  // namespace ob
