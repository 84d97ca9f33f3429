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
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATTN_SMEM_BYTES);
    if (e != cudaSuccess) { set_error("attn_fwd smem attr: %s", cudaGetErrorString(e)); return OB_ERR_CUDA; }
    attr = true;
  }
  dim3 grid(BH, (Lq + ATTN_BM - 1) / ATTN_BM);   // query tiles on grid.y, heaviest first (attn_fwd_kernel)
  launch(attn_fwd_kernel<false>, grid, ATTN_THREADS, ATTN_SMEM_BYTES, st, 1, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("attn_fwd launch: %s", cudaGetErrorString(e)); return OB_ERR_CUDA; }
  return OB_OK;
}

// Split factor for the decode kernel: enough CTAs for the 148 SMs, at most one per key tile of the pool's capacity.
int attn_decode_splits(int B, int heads, int hw, int max_pages) {
  const int q_tiles = (hw + ATTN_BM - 1) / ATTN_BM;
  const long cap_tiles = (static_cast<long>(max_pages) * hw + ATTN_BN - 1) / ATTN_BN;
  long s = 148 / (static_cast<long>(B) * heads * q_tiles);
  if (s > cap_tiles) s = cap_tiles;
  if (s > 32) s = 32;
  if (s < 1) s = 1;
  return static_cast<int>(s);
}

int attn_decode(const void* q, const void* k_pages, const void* v_pages, const int* page_table, const int* lengths, void* o,
                float* o_part, float* l_part, int B, int heads, int hw, int max_pages, int n_pages, int n_split,
                int extra_frames, float scale, cudaStream_t st) {
  if (B <= 0 || heads <= 0) return OB_OK;
  if (hw <= 0 || hw % 8 != 0 || (hw > 128 && hw % 128 != 0) || (hw < 128 && 128 % hw != 0) || n_split < 1 ||
      (n_split > 1 && (o_part == nullptr || l_part == nullptr)) || max_pages <= 0 || n_pages <= 0) {
    set_error("attn_decode: unsupported arguments (hw=%d must divide or be a multiple of 128, n_split=%d, max_pages=%d, n_pages=%d)",
              hw, n_split, max_pages, n_pages);
    return OB_ERR_UNSUPPORTED;
  }
  AttnParams p;
  memset(&p, 0, sizeof(p));
  if (int r = make_qkv_map(&p.mapQ, q, B, heads, hw)) return r;
  const int box_rows = hw < 128 ? hw : 128;
  if (int r = make_qkv_map(&p.mapK, k_pages, n_pages, heads, hw, box_rows)) return r;
  if (int r = make_qkv_map(&p.mapV, v_pages, n_pages, heads, hw, box_rows)) return r;
  p.BH = B * heads; p.heads = heads; p.Lq = hw; p.Lk = 0; p.hw = hw; p.n_frames = 0; p.mask = ATTN_FULL; p.scale = scale;
  p.o = static_cast<__nv_bfloat16*>(o); p.lse = nullptr;
  p.page_table = page_table; p.lengths = lengths; p.max_pages = max_pages; p.n_pages = n_pages; p.box_rows = box_rows;
  p.extra_frames = extra_frames; p.n_split = n_split; p.o_part = o_part; p.l_part = l_part;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATTN_SMEM_BYTES);
    if (e != cudaSuccess) { set_error("attn_decode smem attr: %s", cudaGetErrorString(e)); return OB_ERR_CUDA; }
    attr = true;
  }
  dim3 grid((hw + ATTN_BM - 1) / ATTN_BM, B * heads, n_split);
  launch(attn_fwd_kernel<true>, grid, ATTN_THREADS, ATTN_SMEM_BYTES, st, 1, p);
  if (n_split > 1) {
    const long total = static_cast<long>(B) * hw * heads * (ATTN_D / 4);
    launch(attn_decode_combine_kernel, static_cast<unsigned>((total + 255) / 256), 256, 0, st, 1,
           static_cast<const float*>(o_part), static_cast<const float*>(l_part), static_cast<__nv_bfloat16*>(o), n_split, B, hw, heads);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("attn_decode launch: %s", cudaGetErrorString(e)); return OB_ERR_CUDA; }
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
  if (int r = make_qkv_map(&p.mapK128, k, B, heads, Lk, 128)) return r;
  if (int r = make_qkv_map(&p.mapV128, v, B, heads, Lk, 128)) return r;
  p.BH = BH; p.heads = heads; p.Lq = Lq; p.Lk = Lk; p.hw = hw; p.n_frames = n_frames; p.mask = mask; p.scale = scale;
  p.lse = lse; p.ws = dsum;
  p.dq = static_cast<__nv_bfloat16*>(dq); p.dk = static_cast<__nv_bfloat16*>(dk); p.dv = static_cast<__nv_bfloat16*>(dv);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ABW_DQ_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ABW_DKV_SMEM);
    if (e != cudaSuccess) { set_error("attn_bwd smem attr: %s", cudaGetErrorString(e)); return OB_ERR_CUDA; }
    attr = true;
  }
  const int Lp = (Lq + ABW_BN - 1) / ABW_BN * ABW_BN;
  const long rows = static_cast<long>(BH) * Lp;          // padded positions are zero-filled
  launch(attn_bwd_prep_kernel, (rows * 8 + 255) / 256, 256, 0, st, 1, static_cast<const __nv_bfloat16*>(o),
         static_cast<const __nv_bfloat16*>(dout), lse, dsum, rows, Lq, Lp, heads, static_cast<long>(BH), scale);
  launch(attn_bwd_dq_kernel, dim3(BH, (Lq + ABW_BM - 1) / ABW_BM), ABW_THREADS, ABW_DQ_SMEM, st, 1, p);
  launch(attn_bwd_dkv_kernel, dim3(BH, (Lk + ABW_BM - 1) / ABW_BM), ABW_THREADS, ABW_DKV_SMEM, st, 1, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("attn_bwd launch: %s", cudaGetErrorString(e)); return OB_ERR_CUDA; }
  return OB_OK;
}

}  // namespace ob
