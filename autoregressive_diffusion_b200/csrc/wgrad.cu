// Host side of the weight-gradient kernel.
#include "wgrad.cuh"
#include "launch.cuh"

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

// Tile plan shared by the split heuristic and the launch: per-tap input-channel span, taps per CTA, pairing.
struct WgradPlan { int chunk, bn, nt, pair, co_tiles, ci_tiles, halo; };
static WgradPlan make_plan(int n_items, int Cin, int Cout) {
  WgradPlan q;
  q.chunk = chunk_for(Cin) < chunk_for(Cout) ? chunk_for(Cin) : chunk_for(Cout);
  q.bn = pick_bn(Cin, q.chunk);
  // narrow layers (Cin <= 128): two taps share one gradient tile and form an N=256 MMA
  q.nt = (q.chunk == 64 && q.bn == 128 && n_items >= 2) ? 2 : 1;
  q.co_tiles = (Cout + 127) / 128;
  q.ci_tiles = (Cin + q.bn - 1) / q.bn;
  // CTA pairs (cta_group::2, two output-channel tiles sharing the activation boxes) are implemented and validated but
  // OFF by default: measured 5-10 % slower than independent CTAs on every CS shape (512->512 16x16: 150 vs 141 us).
  // CTAs that differ only in their output-channel tile request the same activation boxes at the same time and L2
  // already merges those reads, so the pair saves no traffic and only couples two SMs' pipelines.
  q.pair = 0;
  // wgrad_halo_kernel (the vertical taps (dy, dy+1) of a column share one CTA and one activation load: 150 instead of
  // 87 FLOP per L2->SM byte) is implemented and validated by the probe but OFF.  It cuts the operand traffic as planned,
  // but its 2-tap groups on the full-length K loop of the current-frame taps are 4x the work of a 1-tap context group,
  // the 144 CTAs of a 512->512 layer fit one wave, and the heaviest CTA sets the time: 171 us against 126 us for the
  // flat kernel.  Balancing needs per-group K splits, whose extra fp32 partials cost as much as the halo saves.
  q.halo = 0;
  return q;
}
static int group_count(int n_items, int nt, int halo = 0) {
  if (halo) return n_items / 3 * 2;      // per (dt, dx) column: one two-tap group + one single tap
  if (nt == 1) return n_items;
  return n_items == 27 ? 14 : (n_items + 1) / 2;   // 27 = gated conv: 9 current-frame taps + 18 context taps (two tensor pairs)
}

int wgrad_suggest_split(int n_items, int max_frames, int H, int W, int Cin, int Cout) {
  int bw, bh, bt;
  pixel_box(H, W, &bw, &bh, &bt);
  const WgradPlan q = make_plan(n_items, Cin, Cout);
  const long ctas = static_cast<long>(q.co_tiles) * q.ci_tiles * group_count(n_items, q.nt, q.halo);
  const long k_tiles = static_cast<long>((max_frames + bt - 1) / bt) * ((H + bh - 1) / bh) * ((W + bw - 1) / bw);
  // ONE wave of CTAs: every extra split is another fp32 copy of dW written here and read back by the weight-norm
  // backward.  Measured sweeps (us): 512->512 16x16 {1: 125, 2: 141, 4: 161}; 256->256 16x16 {2: 65, 3: 55, 6: 58};
  // 128->128 32x32 {6: 79, 11: 58, 22: 66}; 256->128 32x32 {3: 144, 6: 89, 11: 97}
  long split = (2 * 148 + ctas) / (2 * ctas);      // nearest to one wave
  if (split > k_tiles / 4) split = k_tiles / 4;  // keep >= 4 K tiles per CTA
  if (split < 1) split = 1;
  if (split > 64) split = 64;
  return static_cast<int>(split);
}

template <int CHUNK, int BN, int NT, bool PAIR>
static int launch_inst(const WgradParams& p, dim3 grid, cudaStream_t stream) {
  using Cfg = WgradCfg<CHUNK, BN, NT, PAIR>;
  if constexpr (BN < CHUNK || (NT == 2 && (BN != 128 || CHUNK != 64)) || (PAIR && (CHUNK != 64 || BN * NT < 128))) {
    set_error("wgrad: unsupported tile <%d,%d,%d,%d>", CHUNK, BN, NT, (int)PAIR);
    return OB_ERR_UNSUPPORTED;
  } else {
    static bool attr_set = false;
    auto kern = wgrad_kernel<CHUNK, BN, NT, PAIR>;
    if (!attr_set) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
      if (e != cudaSuccess) {
        set_error("cudaFuncSetAttribute(wgrad<%d,%d,%d,%d>): %s", CHUNK, BN, NT, (int)PAIR, cudaGetErrorString(e));
        return OB_ERR_CUDA;
      }
      attr_set = true;
    }
    cudaError_t e = launch(kern, grid, dim3(WGRAD_THREADS), Cfg::SMEM_BYTES, stream, PAIR ? 2 : 1, p);
    if (e != cudaSuccess) {
      set_error("wgrad<%d,%d,%d,%d> launch: %s", CHUNK, BN, NT, (int)PAIR, cudaGetErrorString(e));
      return OB_ERR_CUDA;
    }
    return OB_OK;
  }
}

template <int CHUNK>
static int launch_bn(const WgradPlan& q, const WgradParams& p, dim3 grid, cudaStream_t s) {
  if (q.nt == 2) return q.pair ? launch_inst<CHUNK, 128, 2, true>(p, grid, s) : launch_inst<CHUNK, 128, 2, false>(p, grid, s);
  if (q.pair) {
    switch (q.bn) {
      case 128: return launch_inst<CHUNK, 128, 1, true>(p, grid, s);
      case 256: return launch_inst<CHUNK, 256, 1, true>(p, grid, s);
    }
  } else {
    switch (q.bn) {
      case 16: return launch_inst<CHUNK, 16, 1, false>(p, grid, s);
      case 32: return launch_inst<CHUNK, 32, 1, false>(p, grid, s);
      case 64: return launch_inst<CHUNK, 64, 1, false>(p, grid, s);
      case 128: return launch_inst<CHUNK, 128, 1, false>(p, grid, s);
      case 256: return launch_inst<CHUNK, 256, 1, false>(p, grid, s);
    }
  }
  set_error("wgrad: unsupported tile N %d", q.bn);
  return OB_ERR_UNSUPPORTED;
}

// Host side of wgrad_halo_kernel: columns of three vertical taps become a two-tap group (halo boxes) and a single tap.
static int wgrad_halo_launch(const WgradLaunch& L, const WgradPlan& q, cudaStream_t stream) {
  WgradHaloParams p;
  memset(&p, 0, sizeof(p));
  // pixel tile of 64 = bh x bt x bw, at most 16 wide (fewer halo rows per row), rows ordered (row, frame, column)
  p.bw = pow2_ceil_(L.W) > 16 ? 16 : pow2_ceil_(L.W);
  p.bh = pow2_ceil_(L.H);
  if (p.bh > WGRAD_KT / p.bw) p.bh = WGRAD_KT / p.bw;
  p.bt = WGRAD_KT / (p.bw * p.bh);
  if ((p.bt * p.bw) % 8 != 0) return OB_ERR_UNSUPPORTED;     // a vertical shift must be whole 8-row groups
  p.tiles_w = (L.W + p.bw - 1) / p.bw;
  p.tiles_h = (L.H + p.bh - 1) / p.bh;
  const long hw = static_cast<long>(L.H) * L.W;
  for (int s = 0; s < 2; ++s) {
    if (L.g[s] == nullptr) continue;
    p.n_seq[s] = L.g_seq[s];
    p.tiles_t[s] = (L.g_T[s] + p.bt - 1) / p.bt;
    {
      uint64_t dims[5] = {(uint64_t)L.Cout, (uint64_t)L.W, (uint64_t)L.g_T[s], (uint64_t)L.H, (uint64_t)L.g_seq[s]};
      uint64_t str[5] = {1, (uint64_t)L.Cout, (uint64_t)hw * L.Cout, (uint64_t)L.W * L.Cout, (uint64_t)L.g_T[s] * hw * L.Cout};
      uint32_t box[5] = {64, (uint32_t)p.bw, (uint32_t)p.bt, (uint32_t)p.bh, 1};
      int r = encode_tmap_bf16(&p.mapG[s], L.g[s], 5, dims, str, box);
      if (r != OB_OK) return r;
    }
    uint64_t dims[5] = {(uint64_t)L.Cin, (uint64_t)L.W, (uint64_t)L.a_T[s], (uint64_t)L.H, (uint64_t)L.g_seq[s]};
    uint64_t str[5] = {1, (uint64_t)L.Cin, (uint64_t)hw * L.Cin, (uint64_t)L.W * L.Cin, (uint64_t)L.a_T[s] * hw * L.Cin};
    uint32_t box[5] = {64, (uint32_t)p.bw, (uint32_t)p.bt, (uint32_t)p.bh, 1};
    int r = encode_tmap_bf16(&p.mapA[s], L.a[s], 5, dims, str, box);
    if (r != OB_OK) return r;
    box[3] = (uint32_t)p.bh + 1;
    r = encode_tmap_bf16(&p.mapAh[s], L.a[s], 5, dims, str, box);
    if (r != OB_OK) return r;
  }
  if (L.g[1] == nullptr) { p.mapG[1] = p.mapG[0]; p.mapA[1] = p.mapA[0]; p.mapAh[1] = p.mapAh[0]; }
  // groups: items arrive as columns of three vertical taps dy = -1, 0, +1 per (pair, dt, dx)
  const WgradItem* items = static_cast<const WgradItem*>(L.items);
  int ng = 0;
  for (int i = 0; i < L.n_items; ++i) {
    const WgradItem& a = items[i];
    if (L.g[a.pair] == nullptr || a.wtap >= L.w_taps) { set_error("wgrad: item %d is inconsistent", i); return OB_ERR_INVALID; }
    if (a.dy == 0) continue;                       // the centre tap rides with dy = -1
    if (ng >= WGRAD_MAX_ITEMS) { set_error("wgrad: too many tap groups"); return OB_ERR_INVALID; }
    WgradGroup& g = p.groups[ng++];
    g.pair = a.pair; g.dt[0] = a.dt; g.dy[0] = a.dy; g.dx[0] = a.dx; g.wtap[0] = a.wtap; g.wtap[1] = -1; g.pad_ = 1;
    if (a.dy == -1) {
      const WgradItem* c = nullptr;
      for (int j = 0; j < L.n_items; ++j)
        if (items[j].pair == a.pair && items[j].dt == a.dt && items[j].dx == a.dx && items[j].dy == 0) c = &items[j];
      if (c == nullptr) { set_error("wgrad: tap column without a centre tap"); return OB_ERR_INVALID; }
      g.wtap[1] = c->wtap; g.pad_ = 2;
    }
  }
  p.n_groups = ng;
  p.Cin = L.Cin; p.Cout = L.Cout; p.w_taps = L.w_taps;
  p.ci_tiles = q.ci_tiles; p.co_tiles = q.co_tiles;
  p.n_split = L.n_split;
  p.accumulate = L.accumulate;
  p.out = L.out;
  p.shift_bytes = p.bt * p.bw * 128;
  p.a_box_bytes = (p.bh + 1) * p.bt * p.bw * 128;
  p.stage_bytes = 2 * WGRAD_KT * 128 + (WGH_BN / 64) * p.a_box_bytes;
  p.stages = (200 * 1024) / p.stage_bytes;
  if (p.stages > WGH_MAX_STAGES) p.stages = WGH_MAX_STAGES;
  if (p.stages < 2) return OB_ERR_UNSUPPORTED;
  const size_t smem = static_cast<size_t>(p.stages) * p.stage_bytes + 1024 + 256;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(wgrad_halo): %s", cudaGetErrorString(e)); return OB_ERR_CUDA; }
    attr_set = true;
  }
  cudaError_t e = launch(wgrad_halo_kernel, dim3(q.co_tiles * q.ci_tiles, ng, L.n_split), dim3(WGRAD_THREADS), smem, stream, 1, p);
  if (e != cudaSuccess) { set_error("wgrad_halo launch: %s", cudaGetErrorString(e)); return OB_ERR_CUDA; }
  return OB_OK;
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
  WgradPlan q = make_plan(L.n_items, L.Cin, L.Cout);
  if (L.force_mode == 1) { q.nt = 1; q.pair = 0; q.halo = 0; }
  if (L.force_mode == 3) q.halo = (q.chunk == 64 && q.bn == 256 && q.nt == 1 && (L.n_items == 9 || L.n_items == 27)) ? 1 : 0;
  if (q.halo) {
    const int r = wgrad_halo_launch(L, q, stream);
    if (r != OB_ERR_UNSUPPORTED) return r;         // shapes the halo kernel does not take fall through to the flat one
  }
  if (L.force_mode == 2) q.pair = (q.chunk == 64 && q.bn * q.nt >= 128 && q.co_tiles % 2 == 0) ? 1 : 0;
  const int chunk = q.chunk;
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
  // group taps: NT consecutive items of the same tensor pair share one gradient tile
  const WgradItem* items = static_cast<const WgradItem*>(L.items);
  int ng = 0;
  for (int i = 0; i < L.n_items;) {
    if (L.g[items[i].pair] == nullptr || items[i].wtap >= L.w_taps || ng >= WGRAD_MAX_ITEMS) {
      set_error("wgrad: item %d is inconsistent", i);
      return OB_ERR_INVALID;
    }
    WgradGroup& g = p.groups[ng++];
    g.pair = items[i].pair;
    int n = 1;
    if (q.nt == 2 && i + 1 < L.n_items && items[i + 1].pair == items[i].pair && items[i + 1].wtap < L.w_taps) n = 2;
    for (int j = 0; j < 2; ++j) {
      const WgradItem& it = items[i + (j < n ? j : 0)];   // a missing second tap re-reads the first (its columns are dropped)
      g.dt[j] = it.dt; g.dy[j] = it.dy; g.dx[j] = it.dx;
      g.wtap[j] = j < n ? it.wtap : -1;
    }
    i += n;
  }
  p.n_groups = ng;
  p.H = L.H; p.W = L.W; p.Cin = L.Cin; p.Cout = L.Cout; p.w_taps = L.w_taps;
  p.ci_tiles = q.ci_tiles;
  p.co_tiles = q.co_tiles;
  p.n_split = L.n_split;
  p.accumulate = L.accumulate;
  p.out = L.out;
  dim3 grid(q.co_tiles * q.ci_tiles, ng, L.n_split);
  switch (chunk) {
    case 64: return launch_bn<64>(q, p, grid, stream);
    case 32: return launch_bn<32>(q, p, grid, stream);
    default: return launch_bn<16>(q, p, grid, stream);
  }
}

}  // namespace ob
