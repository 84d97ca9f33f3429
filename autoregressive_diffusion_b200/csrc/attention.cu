// Host side of the attention kernels.
#include "attention.cuh"

#include <cstring>

#include "tapconv_host.h"

namespace ob {

static int make_qkv_map(CUtensorMap* m, const void* ptr, int BH, int L) {
  uint64_t dims[3] = {64, (uint64_t)L, (uint64_t)BH};
  uint64_t str[3] = {1, 64, (uint64_t)L * 64};
  uint32_t box[3] = {64, 128, 1};
  return encode_tmap_bf16(m, ptr, 3, dims, str, box);
}

int attn_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int BH, int Lq, int Lk, int hw, int n_frames,
             int mask, float scale, cudaStream_t st) {
  if (BH <= 0 || Lq <= 0 || Lk <= 0) return OB_OK;
  if (hw <= 0 || mask < ATTN_FULL || mask > ATTN_DART || (mask == ATTN_DART && (n_frames <= 0 || Lq != Lk || Lq != 2 * n_frames * hw)) ||
      (mask == ATTN_CAUSAL && Lq != Lk)) {
    set_error("attn_fwd: inconsistent arguments (Lq=%d Lk=%d hw=%d n_frames=%d mask=%d)", Lq, Lk, hw, n_frames, mask);
    return OB_ERR_INVALID;
  }
  AttnParams p;
  memset(&p, 0, sizeof(p));
  if (int r = make_qkv_map(&p.mapQ, q, BH, Lq)) return r;
  if (int r = make_qkv_map(&p.mapK, k, BH, Lk)) return r;
  if (int r = make_qkv_map(&p.mapV, v, BH, Lk)) return r;
  p.BH = BH; p.Lq = Lq; p.Lk = Lk; p.hw = hw; p.n_frames = n_frames; p.mask = mask; p.scale = scale;
  p.o = static_cast<__nv_bfloat16*>(o); p.lse = lse;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATTN_SMEM_BYTES);
    if (e != cudaSuccess) { set_error("attn_fwd smem attr: %s", cudaGetErrorString(e)); return OB_ERR_CUDA; }
    attr = true;
  }
  dim3 grid((Lq + ATTN_BM - 1) / ATTN_BM, BH);
  attn_fwd_kernel<<<grid, ATTN_THREADS, ATTN_SMEM_BYTES, st>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("attn_fwd launch: %s", cudaGetErrorString(e)); return OB_ERR_CUDA; }
  return OB_OK;
}

}  // namespace ob
