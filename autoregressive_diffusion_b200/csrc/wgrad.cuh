// Weight-gradient tap-GEMM ("MN-major x MN-major"): for every tap item
//
//   dW[split, co, wtap, ci] = sum over the CTA's share of pixel rows p:  G[p, co] * A[p + shift, ci]
//
// G is the (already gate-scaled) output gradient, A the saved layer input, both bf16 NHWC, so the
// contraction index (pixels) is the slow axis of both operands. TMA drops [64 pixels][CHUNK channels]
// boxes into shared memory; those are MN-major UMMA operands (LBO = one box, SBO = 8 rows).
// grid = (co tiles * ci tiles, items, k-splits); split partials are summed by the weight-norm backward
// kernel, which has to read dW anyway.  Reference op being differentiated: edm2/conv.py:41,86.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include "ptx.cuh"
#include "tapconv.cuh"

namespace ob {

constexpr int WGRAD_KT = 64;  // pixel rows per pipeline stage
constexpr int WGRAD_THREADS = 192;
constexpr int WGRAD_MAX_ITEMS = 32;

struct WgradItem {
  int8_t pair;  // which (G, A) tensor pair
  int8_t dt, dy, dx;
  int32_t wtap;
};

// One CTA's work: NT taps that share the gradient tile (same tensor pair).  wtap[j] < 0 marks a missing tap.
struct WgradGroup {
  int8_t pair;
  int8_t dt[2], dy[2], dx[2];
  int8_t pad_;
  int32_t wtap[2];
};

struct WgradParams {
  CUtensorMap mapG[2];
  CUtensorMap mapA[2];
  WgradGroup groups[WGRAD_MAX_ITEMS];
  int n_groups;
  int n_seq[2], T[2];  // pixel-row space of each pair (rows of G)
  int H, W;
  int Cin, Cout, w_taps;
  int bw, bh, bt;
  int tiles_w, tiles_h;
  int tiles_t[2];
  int ci_tiles, co_tiles;
  int n_split;
  int accumulate;   // 1: every K slice adds (red.global.add) into slice 0 of out instead of storing its own slice
  float* out;  // [n_split, Cout, w_taps, Cin] fp32 (accumulate: [1, Cout, w_taps, Cin], holding the running sum)
};

// BN = input-channel span of one tap; the MMA's N is BN*NT.  PAIR: two CTAs (adjacent output-channel tiles) run one
// M=256 MMA stream, each staging its own gradient tile but only half of the activation boxes.
template <int CHUNK, int BN, int NT, bool PAIR>
struct WgradCfg {
  static constexpr int N = BN * NT;
  static constexpr int ROW_BYTES = CHUNK * 2;
  static constexpr int BOX_BYTES = WGRAD_KT * ROW_BYTES;
  static constexpr int G_BOXES = 128 / CHUNK;
  static constexpr int A_BOXES = N / CHUNK;                         // of the whole MMA
  static constexpr int A_LOCAL = PAIR ? A_BOXES / 2 : A_BOXES;      // staged by this CTA
  static constexpr int BOXES_PER_TAP = BN / CHUNK;
  static constexpr int G_BYTES = G_BOXES * BOX_BYTES;
  static constexpr int A_BYTES = A_LOCAL * BOX_BYTES;
  static constexpr int STAGE_BYTES = G_BYTES + A_BYTES;
  static constexpr int TX_BYTES = PAIR ? 2 * STAGE_BYTES : STAGE_BYTES;   // credited to the leader's barrier
  static constexpr int STAGES_RAW = (200 * 1024) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
  static constexpr int CW = N >= 32 ? 32 : 16;
};

template <int CHUNK, int BN, int NT, bool PAIR>
__global__ void __launch_bounds__(WGRAD_THREADS, 1) wgrad_kernel(const __grid_constant__ WgradParams p) {
  using Cfg = WgradCfg<CHUNK, BN, NT, PAIR>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int N = Cfg::N;
  constexpr uint32_t SWZ = SwizzleFor<CHUNK>::mode;
  constexpr uint32_t SBO = 8 * Cfg::ROW_BYTES;
  static_assert(!PAIR || Cfg::A_BOXES % 2 == 0, "pair mode splits the activation boxes in two");

  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;

  const WgradGroup grp = p.groups[blockIdx.y];
  const int pr = grp.pair;
  const int co_tile = blockIdx.x % p.co_tiles;   // output-channel tile fastest: the CTAs of a pair share the ci tile
  const int ci_tile = blockIdx.x / p.co_tiles;
  const int co0 = co_tile * 128, ci0 = ci_tile * BN;
  const int k_tiles = p.n_seq[pr] * p.tiles_t[pr] * p.tiles_h * p.tiles_w;
  const int k_begin = static_cast<int>(static_cast<long>(k_tiles) * blockIdx.z / p.n_split);
  const int k_end = static_cast<int>(static_cast<long>(k_tiles) * (blockIdx.z + 1) / p.n_split);

  constexpr uint32_t tmem_cols = N < 32 ? 32 : N;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
    tma_prefetch_desc(&p.mapG[pr]);
    tma_prefetch_desc(&p.mapA[pr]);
  }
  if (warp == 1) {
    if constexpr (PAIR) { tmem_alloc_pair(tmem_slot, tmem_cols); tmem_relinquish_pair(); }
    else { tmem_alloc(tmem_slot, tmem_cols); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();   // setup above overlapped the previous kernel's tail

  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int kt = k_begin; kt < k_end; ++kt) {
      int r = kt;
      const int tw_i = r % p.tiles_w; r /= p.tiles_w;
      const int th_i = r % p.tiles_h; r /= p.tiles_h;
      const int tt_i = r % p.tiles_t[pr];
      const int seq = r / p.tiles_t[pr];
      const int w0 = tw_i * p.bw, h0 = th_i * p.bh, t0 = tt_i * p.bt;
      mbar_wait(empty_bar(stage), phase ^ 1);
      if (elect_one()) {
        const uint32_t sG = smem_base + stage * Cfg::STAGE_BYTES;
        const uint32_t sA = sG + Cfg::G_BYTES;
        if (!PAIR || leader) mbar_arrive_expect_tx(full_bar(stage), Cfg::TX_BYTES);
#pragma unroll
        for (int j = 0; j < Cfg::G_BOXES; ++j) {
          if constexpr (PAIR) tma_load_5d_pair(sG + j * Cfg::BOX_BYTES, &p.mapG[pr], full_bar(stage), co0 + j * CHUNK, w0, h0, t0, seq);
          else tma_load_5d(sG + j * Cfg::BOX_BYTES, &p.mapG[pr], full_bar(stage), co0 + j * CHUNK, w0, h0, t0, seq);
        }
#pragma unroll
        for (int jl = 0; jl < Cfg::A_LOCAL; ++jl) {
          const int j = jl + static_cast<int>(rank) * Cfg::A_LOCAL;   // box index within the whole N extent
          const int tap = j / Cfg::BOXES_PER_TAP, cb = j % Cfg::BOXES_PER_TAP;
          const int ca = ci0 + cb * CHUNK;
          if constexpr (PAIR)
            tma_load_5d_pair(sA + jl * Cfg::BOX_BYTES, &p.mapA[pr], full_bar(stage), ca, w0 + grp.dx[tap], h0 + grp.dy[tap],
                             t0 + grp.dt[tap], seq);
          else
            tma_load_5d(sA + jl * Cfg::BOX_BYTES, &p.mapA[pr], full_bar(stage), ca, w0 + grp.dx[tap], h0 + grp.dy[tap],
                        t0 + grp.dt[tap], seq);
        }
      }
      __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(PAIR ? 256 : 128, N, 1, 1);
    const uint64_t desc0 = make_smem_desc(0, Cfg::BOX_BYTES, SBO, SWZ);
    int stage = 0;
    uint32_t phase = 0;
    if (leader) {
      for (int kt = k_begin; kt < k_end; ++kt) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sG = smem_base + stage * Cfg::STAGE_BYTES;
          const uint64_t adesc = desc0 + (sG >> 4), bdesc = desc0 + ((sG + Cfg::G_BYTES) >> 4);
          if constexpr (PAIR) {
            umma_bf16_ss_pair(tmem_base, adesc, bdesc, idesc, kt > k_begin ? 1u : 0u);
#pragma unroll
            for (int k = 1; k < WGRAD_KT / 16; ++k)
              umma_bf16_ss_pair(tmem_base, adesc + k * ((2 * SBO) >> 4), bdesc + k * ((2 * SBO) >> 4), idesc, 1u);
            umma_commit_pair(empty_bar(stage));
          } else {
            umma_bf16_ss(tmem_base, adesc, bdesc, idesc, kt > k_begin ? 1u : 0u);
#pragma unroll
            for (int k = 1; k < WGRAD_KT / 16; ++k)
              umma_bf16_ss(tmem_base, adesc + k * ((2 * SBO) >> 4), bdesc + k * ((2 * SBO) >> 4), idesc, 1u);
            umma_commit(empty_bar(stage));
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) { if constexpr (PAIR) umma_commit_pair(tmem_full_bar); else umma_commit(tmem_full_bar); }
      __syncwarp();
    }
  } else {
    constexpr int CW = Cfg::CW;
    const int q = warp & 3;
    const int co = co0 + q * 32 + lane;
    mbar_wait_sleep(tmem_full_bar, 0);
    tc_fence_after();
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const bool have = k_end > k_begin;
    for (int c = 0; c < N / CW; ++c) {
      float v[CW];
      if constexpr (CW == 32) tmem_ld32(lane_base + c * CW, v);
      else tmem_ld16(lane_base + c * CW, v);
      tmem_ld_wait();
      const int tap = (c * CW) / BN;              // CW divides BN
      const int wtap = grp.wtap[tap];
      if (co < p.Cout && wtap >= 0) {
        float* dst_row = p.out + ((static_cast<long>(p.accumulate ? 0 : blockIdx.z) * p.Cout + co) * p.w_taps + wtap) * p.Cin;
        const int col0 = ci0 + (c * CW) % BN;
        if (p.accumulate) {           // running sum over K slices AND over the micro-steps of an accumulation cycle
          if (have) {
#pragma unroll
            for (int j = 0; j < CW; j += 4)
              if (col0 + j + 4 <= p.Cin)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst_row + col0 + j), "f"(v[j]), "f"(v[j + 1]),
                             "f"(v[j + 2]), "f"(v[j + 3])
                             : "memory");
          }
        } else {
#pragma unroll
        for (int j = 0; j < CW; j += 4)
          if (col0 + j + 4 <= p.Cin)
            *reinterpret_cast<float4*>(dst_row + col0 + j) =
                have ? make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc_pair(tmem_base, tmem_cols); else tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ============================================================================ halo variant
// Wide layers (256-channel input tiles, 64-channel chunks) sit on the L2->SM operand bandwidth at 87 FLOP per byte
// (ncu: 1.8 GB moved for 155 GFLOP, 13.7 TB/s).  Here a CTA takes the TWO vertical taps (dy0, dy0+1) of one (dt, dx)
// column: the pixel tile is ordered (row, frame, column) and the activation boxes carry ONE extra image row, so the two
// taps are the same shared-memory boxes read through UMMA descriptors that differ by a whole number of 8-row groups,
// and the gradient tile is shared as well: 56 KB per 8.4 MFLOP (150 FLOP/B) instead of 48 KB per 4.2 MFLOP.  The third
// tap of a column (dy = +1) is a single-tap group of the same kernel (plain boxes, one accumulator).
struct WgradHaloParams {
  CUtensorMap mapG[2];    // (C, W, T, H, SEQ), box (64, bw, bt, bh, 1)
  CUtensorMap mapA[2];    // same box
  CUtensorMap mapAh[2];   // box (64, bw, bt, bh + 1, 1)
  WgradGroup groups[WGRAD_MAX_ITEMS];   // dt/dy/dx[0] = first tap; pad_ = number of taps (1 or 2); wtap[j]
  int n_groups;
  int n_seq[2];
  int tiles_t[2];
  int tiles_w, tiles_h;
  int bw, bh, bt;
  int Cin, Cout, w_taps;
  int ci_tiles, co_tiles;
  int n_split;
  int accumulate;     // as WgradParams::accumulate
  int a_box_bytes;    // (bh + 1) * bt * bw * 128: one 64-channel halo box (also the box stride of single-tap groups)
  int shift_bytes;    // bt * bw * 128: where the second tap's first pixel row sits inside a halo box
  int stages, stage_bytes;
  float* out;         // [n_split, Cout, w_taps, Cin] fp32
};

constexpr int WGH_BN = 256;
constexpr int WGH_MAX_STAGES = 4;

static __global__ void __launch_bounds__(WGRAD_THREADS, 1) wgrad_halo_kernel(const __grid_constant__ WgradHaloParams p) {
  constexpr int ROW_BYTES = 128, BOX_BYTES = WGRAD_KT * ROW_BYTES, G_BYTES = 2 * BOX_BYTES, A_BOXES = WGH_BN / 64;
  constexpr uint32_t SBO = 8 * ROW_BYTES;
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int stages = p.stages;
  const uint32_t bar_base = smem_base + stages * p.stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (WGH_MAX_STAGES + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * WGH_MAX_STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * WGH_MAX_STAGES + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const WgradGroup grp = p.groups[blockIdx.y];
  const int pr = grp.pair;
  const int n_taps = grp.pad_;
  const int co_tile = blockIdx.x % p.co_tiles;
  const int ci_tile = blockIdx.x / p.co_tiles;
  const int co0 = co_tile * 128, ci0 = ci_tile * WGH_BN;
  const int k_tiles = p.n_seq[pr] * p.tiles_t[pr] * p.tiles_h * p.tiles_w;
  const int k_begin = static_cast<int>(static_cast<long>(k_tiles) * blockIdx.z / p.n_split);
  const int k_end = static_cast<int>(static_cast<long>(k_tiles) * (blockIdx.z + 1) / p.n_split);
  constexpr uint32_t tmem_cols = 2 * WGH_BN;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
    tma_prefetch_desc(&p.mapG[pr]);
    tma_prefetch_desc(n_taps == 2 ? &p.mapAh[pr] : &p.mapA[pr]);
  }
  if (warp == 1) { tmem_alloc(tmem_slot, tmem_cols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();

  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    const void* mapA = n_taps == 2 ? static_cast<const void*>(&p.mapAh[pr]) : static_cast<const void*>(&p.mapA[pr]);
    const uint32_t tx = G_BYTES + A_BOXES * (n_taps == 2 ? p.a_box_bytes : BOX_BYTES);
    for (int kt = k_begin; kt < k_end; ++kt) {
      int r = kt;
      const int tw_i = r % p.tiles_w; r /= p.tiles_w;
      const int th_i = r % p.tiles_h; r /= p.tiles_h;
      const int tt_i = r % p.tiles_t[pr];
      const int seq = r / p.tiles_t[pr];
      const int w0 = tw_i * p.bw, h0 = th_i * p.bh, t0 = tt_i * p.bt;
      mbar_wait(empty_bar(stage), phase ^ 1);
      if (elect_one()) {
        const uint32_t sG = smem_base + stage * p.stage_bytes;
        const uint32_t sA = sG + G_BYTES;
        mbar_arrive_expect_tx(full_bar(stage), tx);
#pragma unroll
        for (int j = 0; j < 2; ++j) tma_load_5d(sG + j * BOX_BYTES, &p.mapG[pr], full_bar(stage), co0 + j * 64, w0, t0, h0, seq);
#pragma unroll
        for (int j = 0; j < A_BOXES; ++j)
          tma_load_5d(sA + j * p.a_box_bytes, mapA, full_bar(stage), ci0 + j * 64, w0 + grp.dx[0], t0 + grp.dt[0],
                      h0 + grp.dy[0], seq);
      }
      __syncwarp();
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(128, WGH_BN, 1, 1);
    const uint64_t adesc0 = make_smem_desc(0, BOX_BYTES, SBO, SWZ_128B);
    const uint64_t bdesc0 = make_smem_desc(0, p.a_box_bytes, SBO, SWZ_128B);
    int stage = 0;
    uint32_t phase = 0;
    for (int kt = k_begin; kt < k_end; ++kt) {
      mbar_wait(full_bar(stage), phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sG = smem_base + stage * p.stage_bytes;
        const uint64_t adesc = adesc0 + (sG >> 4);
        const uint32_t acc = kt > k_begin ? 1u : 0u;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          if (j < n_taps) {
            const uint64_t bdesc = bdesc0 + ((sG + G_BYTES + j * p.shift_bytes) >> 4);
            umma_bf16_ss(tmem_base + j * WGH_BN, adesc, bdesc, idesc, acc);
#pragma unroll
            for (int k = 1; k < WGRAD_KT / 16; ++k)
              umma_bf16_ss(tmem_base + j * WGH_BN, adesc + k * ((2 * SBO) >> 4), bdesc + k * ((2 * SBO) >> 4), idesc, 1u);
          }
        }
        umma_commit(empty_bar(stage));
      }
      __syncwarp();
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
    if (elect_one()) umma_commit(tmem_full_bar);
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int co = co0 + q * 32 + lane;
    mbar_wait_sleep(tmem_full_bar, 0);
    tc_fence_after();
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const bool have = k_end > k_begin;
    for (int j = 0; j < n_taps; ++j) {
      float* dst_row = p.out + ((static_cast<long>(p.accumulate ? 0 : blockIdx.z) * p.Cout + co) * p.w_taps + grp.wtap[j]) * p.Cin;
      for (int c = 0; c < WGH_BN / 32; ++c) {
        float v[32];
        tmem_ld32(lane_base + j * WGH_BN + c * 32, v);
        tmem_ld_wait();
        if (co < p.Cout) {
          const int col0 = ci0 + c * 32;
          if (p.accumulate) {
            if (have) {
#pragma unroll
              for (int u = 0; u < 32; u += 4)
                if (col0 + u + 4 <= p.Cin)
                  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst_row + col0 + u), "f"(v[u]), "f"(v[u + 1]),
                               "f"(v[u + 2]), "f"(v[u + 3])
                               : "memory");
            }
          } else {
#pragma unroll
          for (int u = 0; u < 32; u += 4)
            if (col0 + u + 4 <= p.Cin)
              *reinterpret_cast<float4*>(dst_row + col0 + u) =
                  have ? make_float4(v[u], v[u + 1], v[u + 2], v[u + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

}  // namespace ob
