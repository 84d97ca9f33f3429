// Host side of the tap-GEMM kernel: tensor-map construction, tile-shape choice, launch.
#include "tapconv_host.h"
#include "launch.cuh"

#include <cudaTypedefs.h>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <mutex>

#include "tapconv.cuh"

namespace ob {

static thread_local char g_err[512] = "";
const char* last_error() { return g_err; }
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// cuTensorMapEncodeTiled is a driver entry point; fetch it through the runtime so the library
// has no link-time dependency on libcuda.
static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  return fn;
}

static CUtensorMapSwizzle swizzle_for_bytes(int inner_bytes) {
  switch (inner_bytes) {
    case 128: return CU_TENSOR_MAP_SWIZZLE_128B;
    case 64: return CU_TENSOR_MAP_SWIZZLE_64B;
    case 32: return CU_TENSOR_MAP_SWIZZLE_32B;
    default: return CU_TENSOR_MAP_SWIZZLE_NONE;
  }
}

int encode_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                     const uint32_t* box) {
  // cuTensorMapEncodeTiled is a driver call: it needs the primary context to be current on THIS host thread
  // (autograd's backward threads may reach us before making any runtime call of their own).
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    cudaFree(nullptr);
    ctx_bound = true;
  }
  auto enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return OB_ERR_CUDA;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) gstr[i - 1] = strides_elems[i] * 2;  // bytes
  }
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gdim, gstr, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_bytes(static_cast<int>(box[0]) * 2),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: CUresult %d (rank %d, dims %llu %llu %llu, box %u %u %u)", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)(rank > 2 ? dims[2] : 0),
              box[0], box[1], rank > 2 ? box[2] : 0);
    return OB_ERR_CUDA;
  }
  return OB_OK;
}

static int pow2_ceil(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

template <int CHUNK, int BN, bool BMN>
static int launch_pair(const TapConvParams& p, dim3 grid, cudaStream_t stream) {
  using Cfg = TapConvCfg<CHUNK, BN>;
  if constexpr (BN < 128 || CHUNK != 64) {
    set_error("tapconv: pair mode needs tile N >= 128 and 64-channel chunks");
    return OB_ERR_UNSUPPORTED;
  } else {
    constexpr int MAX_DYN = 226 * 1024;
    static bool attr_set = false;
    auto kern = tapconv_kernel<CHUNK, BN, BMN, true>;
    if (!attr_set) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_DYN);
      if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(tapconv pair): %s", cudaGetErrorString(e)); return OB_ERR_CUDA; }
      attr_set = true;
    }
    const size_t smem = 1024 + p.a_slots * p.a_slot_bytes + p.b_slots * (Cfg::B_BYTES_AL / 2) + 256;
    cudaError_t e = launch(kern, grid, dim3(TAPCONV_THREADS), smem, stream, 2, p);
    if (e != cudaSuccess) { set_error("tapconv pair<%d,%d> launch: %s", CHUNK, BN, cudaGetErrorString(e)); return OB_ERR_CUDA; }
    return OB_OK;
  }
}

template <int CHUNK, int BN, bool BMN>
static int launch_inst(const TapConvParams& p, dim3 grid, cudaStream_t stream) {
  using Cfg = TapConvCfg<CHUNK, BN>;
  if constexpr (BMN && (BN % CHUNK != 0)) {
    set_error("tapconv: MN-major weights need tile N %d to be a multiple of %d", BN, CHUNK);
    return OB_ERR_UNSUPPORTED;
  } else {
    if (p.m_tiles_pad > 0) return launch_pair<CHUNK, BN, BMN>(p, grid, stream);
    constexpr int MAX_DYN = 226 * 1024;
    static bool attr_set = false;  // benign race: setting twice is harmless
    if (!attr_set) {
      cudaError_t e = cudaFuncSetAttribute(tapconv_kernel<CHUNK, BN, BMN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_DYN);
      if (e != cudaSuccess) {
        set_error("cudaFuncSetAttribute(tapconv<%d,%d>, %d B): %s", CHUNK, BN, MAX_DYN, cudaGetErrorString(e));
        return OB_ERR_CUDA;
      }
      attr_set = true;
    }
    const int smem = 1024 + p.a_slots * p.a_slot_bytes + p.b_slots * Cfg::B_BYTES_AL + p.stage_pad + 256;
    if (smem > MAX_DYN) { set_error("tapconv<%d,%d>: %d bytes of shared memory exceed the limit", CHUNK, BN, smem); return OB_ERR_UNSUPPORTED; }
    // cluster split-K: the ksplit CTAs of a tile (consecutive blockIdx.y) form one cluster
    launch_xy(tapconv_kernel<CHUNK, BN, BMN, false>, grid, dim3(TAPCONV_THREADS), smem, stream, 1, p.csplit ? p.ksplit : 1, p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
      set_error("tapconv<%d,%d> launch: %s", CHUNK, BN, cudaGetErrorString(e));
      return OB_ERR_CUDA;
    }
    return OB_OK;
  }
}

template <int CHUNK, bool BMN>
static int launch_bn(int bn, const TapConvParams& p, dim3 grid, cudaStream_t s) {
  switch (bn) {
    case 16: return launch_inst<CHUNK, 16, BMN>(p, grid, s);
    case 32: return launch_inst<CHUNK, 32, BMN>(p, grid, s);
    case 64: return launch_inst<CHUNK, 64, BMN>(p, grid, s);
    case 128: return launch_inst<CHUNK, 128, BMN>(p, grid, s);
    case 256: return launch_inst<CHUNK, 256, BMN>(p, grid, s);
  }
  set_error("unsupported tile N %d", bn);
  return OB_ERR_UNSUPPORTED;
}

static void tile_shape(int H, int W, int* bw, int* bh, int* bt) {
  // at most 16 pixels wide: a squarer box carries fewer halo rows per output row ((bh+2)/bh: 1.25 for 8 rows against
  // 1.5 for 4), and the halo is pure extra L2->SM traffic on the operand that bounds the wide layers
  *bw = pow2_ceil(W) > 16 ? 16 : pow2_ceil(W);
  *bh = pow2_ceil(H);
  if (*bh > 128 / *bw) *bh = 128 / *bw;
  if (*bh > 16) *bh = 16;
  *bt = 128 / (*bw * *bh);
}

// Tile-N and split-K policy.  Wide N keeps the MMA off the shared-memory read limit (N=64 re-reads the activation
// tile twice as often per FLOP), so when a layer has too few output tiles for 148 SMs the first remedy is to slice
// its K loop (channel chunks) over extra CTAs, and only then to narrow N.
static void pick_tiling(int n_acc, int Cout, int Cin, int chunk, int taps, int m_tiles, int force_bn, int b_mn_major, bool can_split,
                        bool cluster, int* bn_out, int* ks_out) {
  const int bn_max = (n_acc == 3) ? 128 : 256;
  int bn = pow2_ceil(Cout) < 16 ? 16 : pow2_ceil(Cout);
  if (bn > bn_max) bn = bn_max;
  const int n_chunks = (Cin + chunk - 1) / chunk;
  auto split_for = [&](int b) {
    const int tiles = m_tiles * ((Cout + b - 1) / b);
    int ks = 1;
    // a split launch costs a memset and a finishing kernel: only worth it when the K loop is long (3x3 / 3x3x3 taps)
    // cluster split-K stages n_acc * b * 128 fp32 partials in shared memory beside the rings: only narrow tiles qualify
    const bool fits = !cluster || static_cast<long>(n_acc) * b * 128 * 4 <= 100 * 1024;
    if (can_split && fits && tiles <= 74 && n_chunks >= 4 && n_chunks * taps >= 48 && Cout % 4 == 0) {
      ks = 148 / tiles;
      // at least two chunks per slice -- except on the 4x4 level (<= 16 wide tiles): there one chunk per slice with N=128
      // CTA pairs beats two chunks with N=64 (512->512 4x4: 31.8 / 33.8 us against 36.1 / 36.3 us, fwd / dgrad;
      // profiles/r02_small_layer_sweep.txt) because the pair halves the weight bytes each SM pulls through L2
      const int min_chunks = (tiles <= 16 && b >= 128) ? 1 : 2;
      if (ks > n_chunks / min_chunks) ks = n_chunks / min_chunks;
      if (ks > 8) ks = 8;
      if (ks < 1) ks = 1;
    }
    return ks;
  };
  static const int env_bn = [] { const char* e = getenv("ONIRIS_SMALL_BN"); return e ? atoi(e) : 0; }();     // experiment knobs for the
  static const int env_ks = [] { const char* e = getenv("ONIRIS_SMALL_KS"); return e ? atoi(e) : 0; }();     // 4x4 / 8x8 levels
  if (env_bn > 0 && m_tiles <= 16 && taps >= 9 && Cout >= env_bn && n_chunks >= 4) {
    *bn_out = (n_acc == 3 && env_bn > 128) ? 128 : env_bn;
    int ks = env_ks > 0 ? env_ks : split_for(*bn_out);
    if (ks > n_chunks) ks = n_chunks;
    *ks_out = can_split ? ks : 1;
    return;
  }
  if (force_bn > 0) bn = force_bn;
  else {
    while (bn > 64) {
      const int tiles = m_tiles * ((Cout + bn - 1) / bn);
      // 96: the 768-channel 8x8 input gradient (16 x 6 tiles) stays on N=128 CTA pairs in one wave instead of 192 N=64
      // tiles in two (140 -> ~75 us, the time of the 1024-channel layer with the same K)
      if (tiles * split_for(bn) >= 96) break;
      bn >>= 1;
    }
  }
  if (b_mn_major && bn < chunk) bn = chunk;
  *bn_out = bn;
  *ks_out = split_for(bn);
}

// Split-K reduction strategy.  Default: red.global.add into an fp32 workspace + tapconv_finish_kernel.
// ONIRIS_CSPLIT=1 selects the experimental cluster variant (the slices of a tile form a thread-block cluster and reduce
// their partial accumulators through distributed shared memory inside the one launch).  Measured on the CS layers
// (profiles/r02_conv_breakdown_csplit.txt): NOT faster -- 512->512 4x4 41-43 us against 36-39 us -- because these layers are
// bound by L2->SM operand traffic (129 MB per launch at 4x4), not by the reduction; and its staging area forces narrow
// tiles, which costs the 8x8 level its CTA pairs (81 us against 58-62 us).  The MN-major (input-gradient) instantiation
// with N=128 still returns garbage in columns 64..127 under it; the variant stays off and is not covered by the tests.
static bool cluster_split_enabled() {
  static const bool on = [] { const char* e = getenv("ONIRIS_CSPLIT"); return e != nullptr && e[0] == '1'; }();
  return on;
}
static int pow2_floor(int v) {
  int p = 1;
  while (2 * p <= v) p <<= 1;
  return p;
}

// Small-spatial layers (the 4x4 / 8x8 levels) have too few output tiles to occupy 148 SMs while their K loop is
// hundreds of steps long: slice the channel chunks over blockIdx.y (split-K).
void tapconv_plan(int n_seq, int n_out, int gated, int taps, int T, int H, int W, int Cin, int Cout, int* ksplit, long* ws_bytes) {
  int bw, bh, bt;
  tile_shape(H, W, &bw, &bh, &bt);
  const int m_tiles = n_seq * ((T + bt - 1) / bt) * ((H + bh - 1) / bh) * ((W + bw - 1) / bw);
  const int n_acc = n_out + (gated ? 1 : 0);
  const int chunk = (Cin % 64 == 0) ? 64 : (Cin % 32 == 0) ? 32 : 16;
  int bn, ks;
  pick_tiling(n_acc, Cout, Cin, chunk, taps, m_tiles, 0, 0, true, cluster_split_enabled(), &bn, &ks);
  *ksplit = ks;
  *ws_bytes = (ks > 1 && !cluster_split_enabled()) ? static_cast<long>(n_acc) * n_seq * T * H * W * Cout * 4 : 0;
}

int tapconv_launch(const TapConvLaunch& L, cudaStream_t stream) {
  if (L.Cin % 8 != 0 || L.Cout % 8 != 0) {
    set_error("tapconv: Cin (%d) and Cout (%d) must be multiples of 8", L.Cin, L.Cout);
    return OB_ERR_INVALID;
  }
  if (L.n_cols < 1 || L.n_cols > TAPCONV_MAX_COLS) {
    set_error("tapconv: bad column count %d", L.n_cols);
    return OB_ERR_INVALID;
  }
  if (L.n_seq <= 0 || L.T <= 0) return OB_OK;  // empty problem
  const int n_acc = L.n_out + (L.epi == EPI_GATED ? 1 : 0);
  if (n_acc > 3 || L.n_out < 1) {
    set_error("tapconv: unsupported accumulator count %d", n_acc);
    return OB_ERR_INVALID;
  }
  const int chunk = (L.Cin % 64 == 0) ? 64 : (L.Cin % 32 == 0) ? 32 : 16;

  TapConvParams p;
  memset(&p, 0, sizeof(p));
  // ---- pixel tile: 128 rows ordered (hh, tt, ww); bt*bw >= 8 keeps a vertical shift a whole number of swizzle atoms
  tile_shape(L.H, L.W, &p.bw, &p.bh, &p.bt);
  p.halo = L.halo ? 1 : 0;
  p.tiles_w = (L.W + p.bw - 1) / p.bw;
  p.tiles_h = (L.H + p.bh - 1) / p.bh;
  p.tiles_t = (L.T + p.bt - 1) / p.bt;
  const int m_tiles = L.n_seq * p.tiles_t * p.tiles_h * p.tiles_w;

  // ---- tile N: as wide as TMEM allows, narrowed while the grid cannot fill the 148 SMs
  int bn, ks_plan;
  int taps = 0;
  for (int i = 0; i < L.n_cols; ++i) taps += static_cast<const TapCol*>(L.cols)[i].n_taps;
  const bool want_cs = cluster_split_enabled() && L.split_ws == nullptr && L.trace == nullptr && !L.no_persist && L.post == 0;
  pick_tiling(n_acc, L.Cout, L.Cin, chunk, taps, m_tiles, L.force_bn, L.b_mn_major, L.split_ws != nullptr || want_cs, want_cs, &bn, &ks_plan);
  // cluster split-K: power-of-two slices (rows of the tile are dealt out evenly), at most the portable cluster size, one
  // tile per cluster
  bool csplit = false;
  if (want_cs && ks_plan > 1) {
    int ks = pow2_floor(ks_plan > 8 ? 8 : ks_plan);
    const int tiles_all = m_tiles * ((L.Cout + bn - 1) / bn);
    while (ks > 1 && tiles_all * ks > 148) ks >>= 1;
    const long stage = static_cast<long>(n_acc) * bn * 128 * 4;
    if (ks > 1 && bn / ks >= 8 && stage <= 100 * 1024) { csplit = true; ks_plan = ks; }
    else ks_plan = 1;
  } else if (L.split_ws == nullptr) {
    ks_plan = 1;
  }
  if (n_acc * bn > 512) {
    set_error("tapconv: %d accumulators x N=%d exceed TMEM", n_acc, bn);
    return OB_ERR_INVALID;
  }
  p.tiles_n = (L.Cout + bn - 1) / bn;

  // CTA pairs (cta_group::2) for the wide-N, 64-channel-chunk layers with at least one full pair of pixel tiles:
  // each CTA of a pair stages only half of every weight tile
  static const bool no_pair_env = [] { const char* e = getenv("ONIRIS_NO_PAIR"); return e != nullptr && e[0] == '1'; }();   // A/B + test hook
  const bool pair = (L.use_pair != 0) && !no_pair_env && bn >= 128 && chunk == 64 && m_tiles >= 2 && !csplit;
  p.m_tiles_pad = pair ? (m_tiles + 1) / 2 * 2 : 0;

  // ---- shared-memory rings
  const int rows_a = (p.bh + 2 * p.halo) * p.bt * p.bw;
  p.a_tile_bytes = rows_a * chunk * 2;
  p.a_slot_bytes = (L.n_out * p.a_tile_bytes + 1023) / 1024 * 1024;
  const int b_al = ((bn * chunk * 2 + 1023) / 1024 * 1024) / (pair ? 2 : 1);
  // cluster split-K keeps a staging area for the partial accumulators beside the rings (their K loops are short)
  const int budget = 208 * 1024 - (csplit ? n_acc * bn * 128 * 4 : 0);
  p.a_slots = pair ? TAPCONV_MAX_A_SLOTS : 3;
  while (p.a_slots > 2 && (budget - p.a_slots * p.a_slot_bytes) / b_al < 4) --p.a_slots;   // keep >= 4 weight tiles in flight
  p.b_slots = (budget - p.a_slots * p.a_slot_bytes) / b_al;
  if (p.b_slots > 8) p.b_slots = 8;
  if (p.b_slots < 2) {
    set_error("tapconv: shared memory cannot hold the rings (A slot %d B, B tile %d B)", p.a_slot_bytes, b_al);
    return OB_ERR_UNSUPPORTED;
  }

  {
    // short K loops (1x1 kernels) never fill the rings: shrink them to the loop length so that two CTAs fit on an SM
    // and one's epilogue overlaps the other's loads
    const int n_chunks = (L.Cin + chunk - 1) / chunk;
    const int per_cta = (n_chunks + ks_plan - 1) / ks_plan;
    if (p.a_slots > L.n_cols * per_cta) p.a_slots = L.n_cols * per_cta;
    if (p.b_slots > taps * per_cta) p.b_slots = taps * per_cta;
  }

  // ---- tensor maps: activations as (C, W, T, H, SEQ) so that the tile's rows come out ordered (hh, tt, ww)
  for (int s = 0; s < 2; ++s) {
    if (L.a[s] == nullptr) continue;
    uint64_t dims[5] = {(uint64_t)L.Cin, (uint64_t)L.W, (uint64_t)L.a_T[s], (uint64_t)L.H, (uint64_t)L.a_seq[s]};
    uint64_t str[5] = {1, (uint64_t)L.a_stride_w[s], (uint64_t)L.a_stride_t[s], (uint64_t)L.a_stride_h[s],
                       (uint64_t)L.a_stride_seq[s]};
    uint32_t box[5] = {(uint32_t)chunk, (uint32_t)p.bw, (uint32_t)p.bt, (uint32_t)(p.bh + 2 * p.halo), 1};
    int r = encode_tmap_bf16(&p.mapA[s], L.a[s], 5, dims, str, box);
    if (r != OB_OK) return r;
  }
  if (L.a[1] == nullptr) p.mapA[1] = p.mapA[0];
  if (!L.b_mn_major) {  // weights [Cout][w_taps*Cin]
    uint64_t dims[2] = {(uint64_t)L.w_taps * L.Cin, (uint64_t)L.Cout};
    uint64_t str[2] = {1, (uint64_t)L.w_taps * L.Cin};
    uint32_t box[2] = {(uint32_t)chunk, (uint32_t)(pair ? bn / 2 : bn)};
    int r = encode_tmap_bf16(&p.mapB, L.wg, 2, dims, str, box);
    if (r != OB_OK) return r;
  } else {  // weights [Cin][w_taps*Cout] (the forward matrix of the transposed problem)
    uint64_t dims[2] = {(uint64_t)L.w_taps * L.Cout, (uint64_t)L.Cin};
    uint64_t str[2] = {1, (uint64_t)L.w_taps * L.Cout};
    uint32_t box[2] = {(uint32_t)chunk, (uint32_t)chunk};
    int r = encode_tmap_bf16(&p.mapB, L.wg, 2, dims, str, box);
    if (r != OB_OK) return r;
  }
  for (int i = 0; i < L.n_cols; ++i) {
    p.cols[i] = static_cast<const TapCol*>(L.cols)[i];
    const TapCol& c = p.cols[i];
    bool ok = L.a[c.src] != nullptr && c.acc + c.n_a <= n_acc && c.n_a <= L.n_out && (c.n_taps == 1 || (c.n_taps == 3 && p.halo));
    for (int d = 0; ok && d < c.n_taps; ++d) ok = c.wtap[d] >= 0 && c.wtap[d] < L.w_taps;
    if (!ok) {
      set_error("tapconv: tap column %d is inconsistent", i);
      return OB_ERR_INVALID;
    }
  }
  p.n_cols = L.n_cols;
  p.n_seq = L.n_seq; p.T = L.T; p.H = L.H; p.W = L.W;
  p.Cin = L.Cin; p.Cout = L.Cout;
  p.n_out = L.n_out; p.epi = L.epi; p.out_f32 = L.out_f32;
  p.alpha = L.alpha; p.beta = L.beta; p.bias = L.bias; p.out = L.out; p.out_d = L.out_d;
  p.post = L.post; p.out2 = L.out2; p.cscale = L.cscale; p.cscale_ld = L.cscale_ld; p.res = L.res;
  p.post_wa = L.post_wa; p.post_wb = L.post_wb; p.post_clip = L.post_clip;
  if (L.post != POST_NONE && (L.out2 == nullptr || L.Cout % 8 != 0 || (L.post == POST_SCALE_SILU && (L.cscale == nullptr || L.cscale_ld % 4 != 0)) ||
                              (L.post == POST_MP_SUM && L.res == nullptr))) {
    set_error("tapconv: inconsistent post-op arguments");
    return OB_ERR_INVALID;
  }
  if (L.out == nullptr && L.post == POST_NONE) { set_error("tapconv: no output"); return OB_ERR_INVALID; }

  p.trace = L.trace;
  p.wide_store = (L.Cout % 16 == 0) && (reinterpret_cast<uintptr_t>(L.out) % 32 == 0) &&
                 (reinterpret_cast<uintptr_t>(L.out_d) % 32 == 0) && bn >= 32;
  p.ksplit = 1;
  p.split_ws = nullptr;
  p.csplit = 0;
  p.stage_pad = 0;
  if (csplit) {
    p.ksplit = ks_plan;
    p.csplit = 1;
    p.stage_pad = n_acc * bn * 128 * 4;
  } else if (L.split_ws != nullptr) {
    const int ks = ks_plan;
    const long wsb = static_cast<long>(n_acc) * L.n_seq * L.T * L.H * L.W * L.Cout * 4;
    if (ks > 1) {
      p.ksplit = ks;
      p.split_ws = L.split_ws;
      cudaError_t e = cudaMemsetAsync(L.split_ws, 0, static_cast<size_t>(wsb), stream);
      if (e != cudaSuccess) { set_error("tapconv: workspace memset: %s", cudaGetErrorString(e)); return OB_ERR_CUDA; }
    }
  }
  // accumulator placement for the persistent tile loop (see tapconv.cuh)
  if (2 * n_acc * bn <= 512) p.tmem_mode = TMEM_DOUBLE;
  else if (n_acc == 3 && 4 * bn <= 512) p.tmem_mode = TMEM_ROTATE;
  else p.tmem_mode = TMEM_SINGLE;
  // persistent grid: at most one CTA per SM (one cluster per SM pair); each walks tiles with stride gridDim.x
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm <= 0) n_sm = 148;
  }
  int tiles_x = pair ? p.m_tiles_pad * p.tiles_n : m_tiles * p.tiles_n;
  int max_x = n_sm / p.ksplit;
  if (max_x < 1) max_x = 1;
  if (pair) max_x &= ~1;
  if (max_x < (pair ? 2 : 1)) max_x = pair ? 2 : 1;
  if (L.no_persist) max_x = tiles_x;
  const dim3 grid(tiles_x < max_x ? tiles_x : max_x, p.ksplit);
  int rc;
  if (L.b_mn_major) {
    switch (chunk) {
      case 64: rc = launch_bn<64, true>(bn, p, grid, stream); break;
      case 32: rc = launch_bn<32, true>(bn, p, grid, stream); break;
      default: rc = launch_bn<16, true>(bn, p, grid, stream); break;
    }
  } else {
    switch (chunk) {
      case 64: rc = launch_bn<64, false>(bn, p, grid, stream); break;
      case 32: rc = launch_bn<32, false>(bn, p, grid, stream); break;
      default: rc = launch_bn<16, false>(bn, p, grid, stream); break;
    }
  }
  if (rc != OB_OK || p.ksplit == 1 || p.csplit) return rc;
  {
    const long hw = static_cast<long>(L.H) * L.W;
    const long total = static_cast<long>(L.n_seq) * L.n_out * L.T * hw * (L.Cout / 4);
    launch(tapconv_finish_kernel, static_cast<unsigned>((total + 255) / 256), 256, 0, stream, 1, 
        p.split_ws, p.alpha, p.beta, p.out, static_cast<__half*>(p.out_d), L.n_seq, L.n_out, L.T, hw, L.Cout, p.epi, p.out_f32, p.bias,
        p.post, static_cast<__nv_bfloat16*>(p.out2), p.cscale, p.cscale_ld, static_cast<const __nv_bfloat16*>(p.res), p.post_wa, p.post_wb, p.post_clip);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("tapconv_finish launch: %s", cudaGetErrorString(e)); return OB_ERR_CUDA; }
  }
  return OB_OK;
}

}  // namespace ob
