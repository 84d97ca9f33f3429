// Host side of the weight-gradient kernel.
#include "wgrad.cuh"

#include <cstring>

#include "tapconv_host.h"

namespace ob {

static int pow2_ceil_(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}
static int chunk_for(int c) { return (c % 64 == 0) ? 64 : (c % 32 == 0) ? 32 : 16; }

static void pixel_box(int H, int W, int* bw, int* bh, int* bt) {
  *bw = pow2_ceil_(W) > WGRAD_KT ? WGRAD_KT : pow2_ceil_(W);
  *bh = pow2_ceil_(H);
  if (*bh > WGRAD_KT / *bw) *bh = WGRAD_KT / *bw;
  *bt = WGRAD_KT / (*bw * *bh);
}
static int pick_bn(int Cin, int chunk) {
  int bn = pow2_ceil_(Cin);
  if (bn > 256) bn = 256;
  if (bn < chunk) bn = chunk;
  if (bn < 16) bn = 16;
  return bn;
}

int wgrad_suggest_split(int n_items, int max_frames, int H, int W, int Cin, int Cout) {
  int bw, bh, bt;
  pixel_box(H, W, &bw, &bh, &bt);
  const int chunk = chunk_for(Cin) < chunk_for(Cout) ? chunk_for(Cin) : chunk_for(Cout);
  const int bn = pick_bn(Cin, chunk);
  const long ctas = static_cast<long>((Cout + 127) / 128) * ((Cin + bn - 1) / bn) * n_items;
  const long k_tiles = static_cast<long>((max_frames + bt - 1) / bt) * ((H + bh - 1) / bh) * ((W + bw - 1) / bw);
  long split = (2 * 148 + ctas - 1) / ctas;
  if (split > k_tiles / 4) split = k_tiles / 4;  // keep >= 4 K tiles per CTA
  if (split < 1) split = 1;
  if (split > 64) split = 64;
  return static_cast<int>(split);
}

template <int CHUNK, int BN>
static int launch_inst(const WgradParams& p, dim3 grid, cudaStream_t stream) {
  using Cfg = WgradCfg<CHUNK, BN>;
  if constexpr (BN < CHUNK) {
    set_error("wgrad: tile N %d below chunk %d", BN, CHUNK);
    return OB_ERR_UNSUPPORTED;
  } else {
    static bool attr_set = false;
    if (!attr_set) {
      cudaError_t e =
          cudaFuncSetAttribute(wgrad_kernel<CHUNK, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
      if (e != cudaSuccess) {
        set_error("cudaFuncSetAttribute(wgrad<%d,%d>): %s", CHUNK, BN, cudaGetErrorString(e));
        return OB_ERR_CUDA;
      }
      attr_set = true;
    }
    wgrad_kernel<CHUNK, BN><<<grid, WGRAD_THREADS, Cfg::SMEM_BYTES, stream>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
      set_error("wgrad<%d,%d> launch: %s", CHUNK, BN, cudaGetErrorString(e));
      return OB_ERR_CUDA;
    }
    return OB_OK;
  }
}

template <int CHUNK>
static int launch_bn(int bn, const WgradParams& p, dim3 grid, cudaStream_t s) {
  switch (bn) {
    case 16: return launch_inst<CHUNK, 16>(p, grid, s);
    case 32: return launch_inst<CHUNK, 32>(p, grid, s);
    case 64: return launch_inst<CHUNK, 64>(p, grid, s);
    case 128: return launch_inst<CHUNK, 128>(p, grid, s);
    case 256: return launch_inst<CHUNK, 256>(p, grid, s);
  }
  set_error("wgrad: unsupported tile N %d", bn);
  return OB_ERR_UNSUPPORTED;
}

int wgrad_launch(const WgradLaunch& L, cudaStream_t stream) {
  if (L.Cin % 8 != 0 || L.Cout % 8 != 0 || L.Cin % 4 != 0) {
    set_error("wgrad: Cin (%d) and Cout (%d) must be multiples of 8", L.Cin, L.Cout);
    return OB_ERR_INVALID;
  }
  if (L.n_items < 1 || L.n_items > WGRAD_MAX_ITEMS || L.n_split < 1) {
    set_error("wgrad: bad item count %d / split %d", L.n_items, L.n_split);
    return OB_ERR_INVALID;
  }
  WgradParams p;
  memset(&p, 0, sizeof(p));
  pixel_box(L.H, L.W, &p.bw, &p.bh, &p.bt);
  p.tiles_w = (L.W + p.bw - 1) / p.bw;
  p.tiles_h = (L.H + p.bh - 1) / p.bh;
  const int chunk = chunk_for(L.Cin) < chunk_for(L.Cout) ? chunk_for(L.Cin) : chunk_for(L.Cout);
  const int bn = pick_bn(L.Cin, chunk);
  for (int s = 0; s < 2; ++s) {
    if (L.g[s] == nullptr) continue;
    p.n_seq[s] = L.g_seq[s];
    p.T[s] = L.g_T[s];
    p.tiles_t[s] = (L.g_T[s] + p.bt - 1) / p.bt;
    const long hw = static_cast<long>(L.H) * L.W;
    {
      uint64_t dims[5] = {(uint64_t)L.Cout, (uint64_t)L.W, (uint64_t)L.H, (uint64_t)L.g_T[s], (uint64_t)L.g_seq[s]};
      uint64_t str[5] = {1, (uint64_t)L.Cout, (uint64_t)L.W * L.Cout, (uint64_t)hw * L.Cout, (uint64_t)L.g_T[s] * hw * L.Cout};
      uint32_t box[5] = {(uint32_t)chunk, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bt, 1};
      int r = encode_tmap_bf16(&p.mapG[s], L.g[s], 5, dims, str, box);
      if (r != OB_OK) return r;
    }
    {
      uint64_t dims[5] = {(uint64_t)L.Cin, (uint64_t)L.W, (uint64_t)L.H, (uint64_t)L.a_T[s], (uint64_t)L.g_seq[s]};
      uint64_t str[5] = {1, (uint64_t)L.Cin, (uint64_t)L.W * L.Cin, (uint64_t)hw * L.Cin, (uint64_t)L.a_T[s] * hw * L.Cin};
      uint32_t box[5] = {(uint32_t)chunk, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bt, 1};
      int r = encode_tmap_bf16(&p.mapA[s], L.a[s], 5, dims, str, box);
      if (r != OB_OK) return r;
    }
  }
  if (L.g[1] == nullptr) { p.mapG[1] = p.mapG[0]; p.mapA[1] = p.mapA[0]; }
  for (int i = 0; i < L.n_items; ++i) {
    p.items[i] = static_cast<const WgradItem*>(L.items)[i];
    if (L.g[p.items[i].pair] == nullptr || p.items[i].wtap >= L.w_taps) {
      set_error("wgrad: item %d is inconsistent", i);
      return OB_ERR_INVALID;
    }
  }
  p.n_items = L.n_items;
  p.H = L.H; p.W = L.W; p.Cin = L.Cin; p.Cout = L.Cout; p.w_taps = L.w_taps;
  p.ci_tiles = (L.Cin + bn - 1) / bn;
  p.n_split = L.n_split;
  p.out = L.out;
  dim3 grid(((L.Cout + 127) / 128) * p.ci_tiles, L.n_items, L.n_split);
  switch (chunk) {
    case 64: return launch_bn<64>(bn, p, grid, stream);
    case 32: return launch_bn<32>(bn, p, grid, stream);
    default: return launch_bn<16>(bn, p, grid, stream);
  }
}

}  // namespace ob
