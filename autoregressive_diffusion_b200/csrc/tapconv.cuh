// Tap-GEMM: the one tensor-core kernel behind every "activations x weights" contraction on the Oniris
// denoiser hot path (reference: edm2/conv.py:36-42 MPConv, :59-95 MPCausal3DGatedConv; their input-gradient
// passes reuse it with the same weight matrix read as an MN-major operand).
//
//   acc_j[m, n] = sum over taps (dt,dy,dx) routed to accumulator j, over channels c:
//                     A[m shifted by (dt,dy,dx), c] * Wg[n, tap, c]
//
// * m walks a 128-row tile of output pixels ordered (row hh, frame tt, column ww) -- bh x bt x bw = 128.  With
//   that order a vertical shift dy moves an operand by a whole number of 8-row swizzle atoms, so ONE shared-
//   memory tile holding bh+2 image rows serves the three taps dy = -1, 0, +1 through three UMMA descriptors that
//   differ only in their start address.  The activation tile is therefore fetched once per (dt, dx, channel
//   chunk) instead of once per tap: 3x less L2->SM traffic on the operand that dominates it.
// * Horizontal / temporal shifts and all zero padding are TMA box coordinates + out-of-bounds zero fill.
//   No im2col buffer exists.  Channel chunks of 64/32/16 use the 128/64/32-byte hardware swizzle.
// * Two independent smem rings: big activation tiles (A ring) and weight tiles (B ring, one per tap).
// * tcgen05.mma (M=128 per CTA, N=BN, K=16; cta_group::2 with M=256 across a CTA pair for the wide layers, see the
//   kernel's comment) accumulates in TMEM; up to 3 accumulators per tile (clean rows, noised rows, shared
//   causal-context term) so the context term of the DART training sequence is computed once for both halves
//   (edm2/conv.py:90-91 duplicates it) and each current-frame weight tile feeds two MMAs.
// * Persistent: a CTA (pair) walks several tiles; TMEM is double-buffered / rotated so loads and MMAs of the next tile
//   overlap the epilogue of the previous one.  Small layers slice K over blockIdx.y instead (split-K + finish kernel).
// * Warp roles: warp0 = activation TMA, warp1 = TMEM owner + MMA issuer, warp2 = weight TMA, warps 3..6 = epilogue
//   (TMEM -> registers -> gated combine -> global).  Producer/MMA loops are warp-uniform with one elected lane
//   issuing, which keeps descriptors in uniform registers (4 UTCHMMA back to back instead of an ELECT loop each).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include "ptx.cuh"
#include "launch.cuh"

namespace ob {

constexpr int TAPCONV_MAX_COLS = 12;
#ifndef TAPCONV_EPI_WARPS_N
#define TAPCONV_EPI_WARPS_N 8
#endif
// warp0: activation TMA, warp1: MMA, warp2: weight TMA, then the epilogue warps: one (4 warps) or two (8 warps) per TMEM lane
// quarter; with two, each takes half of the tile's columns -- the epilogue is a fixed, fully exposed cost of every
// single-wave launch (5-9 us of a 30-55 us layer: profiles/r02_tapconv_phase_trace.txt)
constexpr int TAPCONV_EPI_WARPS = TAPCONV_EPI_WARPS_N;
constexpr int TAPCONV_THREADS = 96 + 32 * TAPCONV_EPI_WARPS;
constexpr int TAPCONV_MAX_A_SLOTS = 4;

enum : int { EPI_PLAIN = 0, EPI_GATED = 1 };
// Post-ops fused behind the output stage (edm2/networks_edm2.py:75-77 and :86,93): a SECOND bf16 output
//   POST_SCALE_SILU: out2 = mp_silu(y * cscale[frame, channel])            (conv_res0 -> embedding scale -> mp_silu)
//   POST_MP_SUM:     out2 = clip(res * wa + y * wb)                         (conv_res1 -> mp_sum with the residual -> clip)
// computed from the fp32 result y before it is rounded.  `out` (the raw y) may then be NULL (evaluation: nothing needs it).
enum : int { POST_NONE = 0, POST_SCALE_SILU = 1, POST_MP_SUM = 2 };

// One "column" of taps: fixed (source, dt, dx); its n_taps entries are the vertical taps dy = -1,0,+1 (or the single
// centre tap of a 1x1 kernel).
struct TapCol {
  int8_t src;      // which activation tensor map (0 or 1)
  int8_t dt;       // frame shift
  int8_t dx;       // horizontal shift
  int8_t n_a;      // number of A tiles sharing each weight tile (1, or 2 in dual mode)
  int8_t acc;      // accumulator of A tile 0 (tile i goes to acc+i)
  int8_t seq_mul;  // source sequence coordinate = seq*seq_mul + i
  int8_t n_taps;   // 1 or 3
  int8_t pad_;
  int32_t wtap[3]; // weight column block (units of Cin) for dy = -1, 0, +1  (n_taps == 1: wtap[0])
};

struct TapConvParams {
  CUtensorMap mapA[2];
  CUtensorMap mapB;
  TapCol cols[TAPCONV_MAX_COLS];
  int n_cols;
  int n_seq, T, H, W;
  int Cin, Cout;
  int bw, bh, bt;
  int halo;            // 1: A tiles carry one extra image row above and below (3x3 kernels)
  int a_tile_bytes;    // (bh + 2*halo) * bt * bw rows * CHUNK*2 bytes
  int a_slot_bytes;    // n_out tiles, rounded up to 1024
  int a_slots;
  int b_slots;
  int tiles_w, tiles_h, tiles_t, tiles_n;
  int tmem_mode;       // TMEM_SINGLE / TMEM_DOUBLE / TMEM_ROTATE (accumulator placement per tile parity)
  int m_tiles_pad;     // PAIR mode: pixel tiles rounded up to a multiple of 2 (blockIdx.x = m + m_tiles_pad*n)
  int n_out;    // output row sets per tile (2 in dual mode)
  int epi;      // EPI_*
  int out_f32;  // 0: bf16 out, 1: fp32 out
  int wide_store;  // 1: the epilogue may use 256-bit stores (Cout % 16 == 0, out / out_d 32-byte aligned)
  const float* alpha;  // [n_seq*n_out*T] per output frame (EPI_GATED)
  const float* beta;
  const float* bias;   // optional fp32 [Cout] added to every output row (EPI_PLAIN; the VAE's nn.Conv3d layers)
  int post;            // POST_*
  void* out2;          // bf16, same shape as out
  const float* cscale; // POST_SCALE_SILU: fp32, row `frame` at cscale + frame*cscale_ld
  int cscale_ld;
  const void* res;     // POST_MP_SUM: bf16 residual, same shape as out
  float post_wa, post_wb, post_clip;
  void* out;    // [n_seq*n_out*T, H, W, Cout]
  void* out_d;  // optional (EPI_GATED): shared - own accumulator, fp16, same shape as out
  int ksplit;   // >1: blockIdx.y owns a slice of the channel chunks (split-K)
  int csplit;   // 1: the ksplit CTAs of a tile form a cluster (1, ksplit, 1) and reduce their partial accumulators through
                //    distributed shared memory inside this launch; 0: they add into split_ws and a finish kernel follows
  int stage_pad;  // csplit: bytes of the accumulator staging area (n_acc*BN*128*4) placed after the operand rings
  long long* trace;   // optional [grid.x*grid.y][8] globaltimer stamps (probe builds only; nullptr in production)
  float* split_ws;  // [n_acc][n_seq*T*H*W][Cout] fp32 (own sets first, then the shared accumulator), zeroed by tapconv_launch
};

template <int CHUNK>
struct SwizzleFor;
template <>
struct SwizzleFor<64> { static constexpr uint32_t mode = SWZ_128B; };
template <>
struct SwizzleFor<32> { static constexpr uint32_t mode = SWZ_64B; };
template <>
struct SwizzleFor<16> { static constexpr uint32_t mode = SWZ_32B; };

template <int CHUNK, int BN>
struct TapConvCfg {
  static constexpr int ROW_BYTES = CHUNK * 2;
  static constexpr int B_BYTES = BN * ROW_BYTES;
  static constexpr int B_BYTES_AL = (B_BYTES + 1023) / 1024 * 1024;
  static constexpr int MAX_SMEM = 200 * 1024;
  static constexpr int CW = BN >= 32 ? 32 : 16;  // epilogue column chunk
  static constexpr int MAX_B_SLOTS = 8;
};

// BMN=false: weights are [N rows][K contiguous] (forward convs).  BMN=true: weights are [K rows][N contiguous],
// i.e. the SAME forward weight matrix read as an MN-major B operand, which is what the input-gradient pass
// needs -- no transposed weight copy is ever materialised.
// PAIR=true: launched as clusters of two CTAs that own two adjacent pixel tiles of the same channel tile and execute
// ONE tcgen05.mma.cta_group::2 stream (M=256): each CTA stages its own activation tiles but only HALF of every weight
// tile, which halves the weight traffic into shared memory and the shared-memory operand reads per FLOP -- the limit a
// single-CTA M=128 x N=128 MMA runs into.
//
// The kernel is PERSISTENT: a CTA (pair) walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...  The producers and the MMA
// issuer run ahead into the next tile while the epilogue warps drain the previous one; without this every CTA of a
// wave reaches its epilogue at the same moment and the whole GPU alternates between a tensor-bound and a
// store-bound phase (measured: 8.4 us of a 27 us tile).  TMEM accumulator placement per tile parity (p.tmem_mode):
//   TMEM_DOUBLE  2*n_acc*BN <= 512: the whole accumulator set alternates between two column ranges.
//   TMEM_ROTATE  three accumulators of 128 columns: the two own accumulators stay in place and the shared (context)
//                accumulator alternates between columns 256 and 384; the host orders the context taps first, so that
//                half of the next tile's main loop overlaps the previous tile's epilogue.
//   TMEM_SINGLE  no spare columns: the next tile's MMAs wait for the epilogue (loads still run ahead).
enum : int { TMEM_SINGLE = 0, TMEM_DOUBLE = 1, TMEM_ROTATE = 2 };

__device__ __forceinline__ long long global_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define TAPCONV_STAMP(slot) \
  do { if (p.trace != nullptr && lane == 0) p.trace[(static_cast<long>(blockIdx.y) * gridDim.x + blockIdx.x) * 8 + (slot)] = global_ns(); } while (0)
// inside a region that only one elected thread executes
#define TAPCONV_STAMP1(slot) \
  do { if (p.trace != nullptr) p.trace[(static_cast<long>(blockIdx.y) * gridDim.x + blockIdx.x) * 8 + (slot)] = global_ns(); } while (0)

template <int CHUNK, int BN, bool BMN, bool PAIR>
__global__ void __launch_bounds__(TAPCONV_THREADS, 1) tapconv_kernel(const __grid_constant__ TapConvParams p) {
  using Cfg = TapConvCfg<CHUNK, BN>;
  static_assert(!BMN || BN % CHUNK == 0, "MN-major weights need BN to be a multiple of CHUNK");
  static_assert(!PAIR || (BN >= 32 && (!BMN || (BN / 2) % CHUNK == 0)), "pair mode needs a splittable weight tile");
  constexpr int B_STRIDE = PAIR ? Cfg::B_BYTES_AL / 2 : Cfg::B_BYTES_AL;   // weight bytes this CTA stages per tap
  constexpr uint32_t SWZ = SwizzleFor<CHUNK>::mode;
  constexpr uint32_t SBO = 8 * Cfg::ROW_BYTES;
  constexpr int MAXB = Cfg::MAX_B_SLOTS;

  pdl_launch_dependents();   // the next kernel may start its prologue; it still waits for this grid to complete
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA0 = smem_base;
  const uint32_t sB0 = sA0 + p.a_slots * p.a_slot_bytes;
  const uint32_t sStage = sB0 + p.b_slots * B_STRIDE;      // csplit: staging area of the partial accumulators
  const uint32_t bar_base = sStage + p.stage_pad;
  // barriers: a_full[MAXA], a_empty[MAXA], b_full[MAXB], b_empty[MAXB], tmem_full[2], tmem_empty[2], TMEM address slot
  constexpr int MAXA = TAPCONV_MAX_A_SLOTS;
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (MAXA + s); };
  auto b_full = [&](int s) { return bar_base + 8u * (2 * MAXA + s); };
  auto b_empty = [&](int s) { return bar_base + 8u * (2 * MAXA + MAXB + s); };
  auto tmem_full = [&](int s) { return bar_base + 8u * (2 * MAXA + 2 * MAXB + s); };
  auto tmem_empty = [&](int s) { return bar_base + 8u * (2 * MAXA + 2 * MAXB + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * MAXA + 2 * MAXB + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;

  // ---- tile walk.  PAIR: tiles (2k, 2k+1) -- the two CTAs of a cluster -- are adjacent pixel tiles of one channel tile
  const int total_tiles = PAIR ? p.m_tiles_pad * p.tiles_n : p.tiles_w * p.tiles_h * p.tiles_t * p.n_seq * p.tiles_n;
  struct Tile { int w0, h0, t0, seq, n0; };
  auto decode = [&](int tile) {
    int n_tile;
    if constexpr (PAIR) {
      // pairs walk the channel tiles fastest: concurrently running clusters then share activation tiles (same pixel
      // pair, different channel tile) as well as weight tiles, and L2 merges the coincident reads
      const int pt = tile >> 1;
      n_tile = pt % p.tiles_n;
      tile = 2 * (pt / p.tiles_n) + (tile & 1);
    } else { n_tile = tile % p.tiles_n; tile /= p.tiles_n; }
    Tile c;
    c.w0 = (tile % p.tiles_w) * p.bw; tile /= p.tiles_w;
    c.h0 = (tile % p.tiles_h) * p.bh; tile /= p.tiles_h;
    c.t0 = (tile % p.tiles_t) * p.bt;
    c.seq = tile / p.tiles_t;             // >= n_seq on the padding tile of an odd pair count: loads zero-fill, stores masked
    c.n0 = n_tile * BN;
    return c;
  };

  const int n_acc = p.n_out + (p.epi == EPI_GATED ? 1 : 0);
  const int mode = p.tmem_mode;
  const int cols_needed = mode == TMEM_DOUBLE ? 2 * n_acc * BN : mode == TMEM_ROTATE ? 4 * BN : n_acc * BN;
  uint32_t tmem_cols = 32;
  while (tmem_cols < static_cast<uint32_t>(cols_needed)) tmem_cols <<= 1;
  // TMEM column of accumulator `acc` for the tile of parity `par`
  auto acc_col = [&](int acc, int par) -> uint32_t {
    if (mode == TMEM_DOUBLE) return static_cast<uint32_t>((par * n_acc + acc) * BN);
    if (mode == TMEM_ROTATE && acc == p.n_out) return static_cast<uint32_t>((2 + par) * BN);
    return static_cast<uint32_t>(acc * BN);
  };
  const int n_chunks_all = (p.Cin + CHUNK - 1) / CHUNK;  // a ragged last chunk is TMA zero-filled
  // split-K: this CTA owns channel chunks [ck_lo, ck_hi)
  const int ck_lo = static_cast<int>(static_cast<long>(n_chunks_all) * blockIdx.y / p.ksplit);
  const int ck_hi = static_cast<int>(static_cast<long>(n_chunks_all) * (blockIdx.y + 1) / p.ksplit);
  const uint32_t dy_stride = static_cast<uint32_t>(p.bt * p.bw) * Cfg::ROW_BYTES;  // one image row of the tile

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.a_slots; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); }
    for (int s = 0; s < p.b_slots; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tmem_full(s), 1); mbar_init(tmem_empty(s), (PAIR ? 2 : 1) * TAPCONV_EPI_WARPS); }
    fence_barrier_init();
    tma_prefetch_desc(&p.mapA[0]);
    tma_prefetch_desc(&p.mapA[1]);
    tma_prefetch_desc(&p.mapB);
  }
  if (warp == 1) {
    if constexpr (PAIR) { tmem_alloc_pair(tmem_slot, tmem_cols); tmem_relinquish_pair(); }
    else { tmem_alloc(tmem_slot, tmem_cols); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();   // the peer's barriers exist before anything is signalled across the pair
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  if (warp == 1) TAPCONV_STAMP(0);   // setup done
  pdl_wait();   // everything above overlapped the previous kernel's tail; global memory is touched only from here on

  if (warp == 0) {
    // ===================== activation (A) TMA producer: ONE elected thread runs the whole loop ============
    if (elect_one()) {
    int as = 0;
    uint32_t aph = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const Tile tc = decode(tile);
      for (int ic = 0; ic < p.n_cols; ++ic) {
        const TapCol col = p.cols[ic];
        const void* mapA = &p.mapA[col.src];
        for (int ck = ck_lo; ck < ck_hi; ++ck) {
          mbar_wait(a_empty(as), aph ^ 1);
          if constexpr (PAIR) {
            if (leader) mbar_arrive_expect_tx(a_full(as), 2 * col.n_a * p.a_tile_bytes);   // both CTAs' tiles
            for (int i = 0; i < col.n_a; ++i)
              tma_load_5d_pair(sA0 + as * p.a_slot_bytes + i * p.a_tile_bytes, mapA, a_full(as), ck * CHUNK,
                               tc.w0 + col.dx, tc.t0 + col.dt, tc.h0 - p.halo, tc.seq * col.seq_mul + i);
          } else {
            mbar_arrive_expect_tx(a_full(as), col.n_a * p.a_tile_bytes);
            for (int i = 0; i < col.n_a; ++i)
              tma_load_5d(sA0 + as * p.a_slot_bytes + i * p.a_tile_bytes, mapA, a_full(as), ck * CHUNK, tc.w0 + col.dx,
                          tc.t0 + col.dt, tc.h0 - p.halo, tc.seq * col.seq_mul + i);
          }
          if (++as == p.a_slots) { as = 0; aph ^= 1; }
        }
      }
    }
    }
    __syncwarp();
  } else if (warp == 2) {
    // ===================== weight (B) TMA producer (one elected thread): runs ahead independently of the A ring ======
    if (elect_one()) {
    int bs = 0;
    uint32_t bph = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int n0 = decode(tile).n0;
      for (int ic = 0; ic < p.n_cols; ++ic) {
        const TapCol col = p.cols[ic];
        for (int ck = ck_lo; ck < ck_hi; ++ck) {
          for (int d = 0; d < col.n_taps; ++d) {
            mbar_wait(b_empty(bs), bph ^ 1);
            const uint32_t sB = sB0 + bs * B_STRIDE;
            if constexpr (PAIR) {
              // this CTA stages output channels [n0 + rank*BN/2, +BN/2) of the tile; the MMA reads both halves
              if (leader) mbar_arrive_expect_tx(b_full(bs), Cfg::B_BYTES);
              if constexpr (!BMN) {
                tma_load_2d_pair(sB, &p.mapB, b_full(bs), col.wtap[d] * p.Cin + ck * CHUNK, n0 + rank * (BN / 2));
              } else {
#pragma unroll
                for (int j = 0; j < BN / 2 / CHUNK; ++j)
                  tma_load_2d_pair(sB + j * (CHUNK * Cfg::ROW_BYTES), &p.mapB, b_full(bs),
                                   col.wtap[d] * p.Cout + n0 + rank * (BN / 2) + j * CHUNK, ck * CHUNK);
              }
            } else {
              mbar_arrive_expect_tx(b_full(bs), Cfg::B_BYTES);
              if constexpr (!BMN) {
                tma_load_2d(sB, &p.mapB, b_full(bs), col.wtap[d] * p.Cin + ck * CHUNK, n0);
              } else {
#pragma unroll
                for (int j = 0; j < BN / CHUNK; ++j)
                  tma_load_2d(sB + j * (CHUNK * Cfg::ROW_BYTES), &p.mapB, b_full(bs), col.wtap[d] * p.Cout + n0 + j * CHUNK,
                              ck * CHUNK);
              }
            }
            if (++bs == p.b_slots) { bs = 0; bph ^= 1; }
          }
        }
      }
    }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // One elected thread runs the whole loop (no per-tap elect / reconvergence: measured on the attention kernels, the
    // issuer's dependent instruction count per step is what limits how far it runs ahead).  Descriptors differ only in
    // their 14-bit start-address field, so they are formed by adding to a base descriptor.
    constexpr uint32_t idesc = make_idesc_bf16(PAIR ? 256 : 128, BN, 0, BMN ? 1 : 0);
    const uint64_t adesc0 = make_smem_desc(0, 16, SBO, SWZ);
    const uint64_t bdesc0 = BMN ? make_smem_desc(0, CHUNK * Cfg::ROW_BYTES, SBO, SWZ) : make_smem_desc(0, 16, SBO, SWZ);
    constexpr uint32_t B_KSTEP = BMN ? (2 * SBO) >> 4 : 32 >> 4;   // descriptor units of 16 bytes
    int as = 0, bs = 0;
    uint32_t aph = 0, bph = 0;
    // how many tiles back an accumulator's previous user lies: its epilogue must have drained it first
    const int own_depth = mode == TMEM_DOUBLE ? 2 : 1;
    const int shr_depth = mode == TMEM_SINGLE ? 1 : 2;
    int drained = -1;   // epilogues of iterations <= drained are known complete
    if (leader && elect_one()) {   // in pair mode the peer's MMA warp only owns its half of the TMEM allocation
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int par = it & 1;
      uint32_t started = 0;  // bit j: accumulator j already holds a partial sum
      const uint32_t own_base = tmem_base + acc_col(0, par), shr_base = tmem_base + acc_col(p.n_out, par);
      for (int ic = 0; ic < p.n_cols; ++ic) {
        const TapCol col = p.cols[ic];
        const bool is_shr = col.acc >= p.n_out;
        const uint32_t d_tmem0 = is_shr ? shr_base : own_base + static_cast<uint32_t>(col.acc * BN);
        const int need = it - (is_shr ? shr_depth : own_depth);
        while (drained < need) {
          ++drained;
          if constexpr (PAIR) mbar_wait_cluster(tmem_empty(drained & 1), (drained >> 1) & 1);
          else mbar_wait(tmem_empty(drained & 1), (drained >> 1) & 1);
          tc_fence_after();
        }
        for (int ck = ck_lo; ck < ck_hi; ++ck) {
          mbar_wait(a_full(as), aph);
          if (it == 0 && ic == 0 && ck == ck_lo) TAPCONV_STAMP1(1);   // first activation tile landed
          const uint32_t sA = sA0 + as * p.a_slot_bytes;
          for (int d = 0; d < col.n_taps; ++d) {
            mbar_wait(b_full(bs), bph);
            tc_fence_after();
            const uint64_t bdesc = bdesc0 + ((sB0 + bs * B_STRIDE) >> 4);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              if (i < col.n_a) {
                const int acc = col.acc + i;
                const uint32_t d_tmem = d_tmem0 + static_cast<uint32_t>(i * BN);
                const uint64_t adesc = adesc0 + ((sA + i * p.a_tile_bytes + d * dy_stride) >> 4);
                if constexpr (PAIR) {
                  umma_bf16_ss_pair(d_tmem, adesc, bdesc, idesc, (started >> acc) & 1u);
#pragma unroll
                  for (int k = 1; k < CHUNK / 16; ++k) umma_bf16_ss_pair(d_tmem, adesc + 2 * k, bdesc + B_KSTEP * k, idesc, 1u);
                } else {
                  umma_bf16_ss(d_tmem, adesc, bdesc, idesc, (started >> acc) & 1u);
#pragma unroll
                  for (int k = 1; k < CHUNK / 16; ++k) umma_bf16_ss(d_tmem, adesc + 2 * k, bdesc + B_KSTEP * k, idesc, 1u);
                }
              }
            }
            if constexpr (PAIR) umma_commit_pair(b_empty(bs)); else umma_commit(b_empty(bs));  // frees the weight slot(s)
            started |= ((1u << col.n_a) - 1u) << col.acc;
            if (++bs == p.b_slots) { bs = 0; bph ^= 1; }
          }
          if constexpr (PAIR) umma_commit_pair(a_empty(as)); else umma_commit(a_empty(as));
          if (++as == p.a_slots) { as = 0; aph ^= 1; }
        }
      }
      if constexpr (PAIR) umma_commit_pair(tmem_full(par)); else umma_commit(tmem_full(par));
    }
    TAPCONV_STAMP1(2);   // all MMAs issued
    }
    __syncwarp();
  } else {
    // ===================== epilogue (warps 3..) =====================
    constexpr int CW = Cfg::CW;
    constexpr int NCH = BN / CW;                                       // column chunks of the tile
    constexpr int CH_PER = (NCH + TAPCONV_EPI_WARPS / 4 - 1) / (TAPCONV_EPI_WARPS / 4);
    const int c_lo = ((warp - 3) >> 2) * CH_PER;                       // this warp's chunks [c_lo, c_hi)
    const int c_hi = c_lo + CH_PER < NCH ? c_lo + CH_PER : NCH;
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const int m = q * 32 + lane;
    const int hh = m / (p.bt * p.bw);        // tile rows are ordered (hh, tt, ww)
    const int rem = m - hh * (p.bt * p.bw);
    const int tt = rem / p.bw;
    const int ww = rem - tt * p.bw;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const long rows_per_set = static_cast<long>(p.n_seq) * p.T * p.H * p.W;   // pixel rows of ONE output set
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int par = it & 1;
      const Tile tc = decode(tile);
      const int seq = tc.seq, n0 = tc.n0;
      const int t = tc.t0 + tt, h = tc.h0 + hh, w = tc.w0 + ww;
      const bool row_ok = (t < p.T) && (h < p.H) && (w < p.W) && (seq < p.n_seq);

      mbar_wait_sleep(tmem_full(par), (it >> 1) & 1);
      tc_fence_after();
      if (warp == 3 && it == 0) TAPCONV_STAMP(3);   // first tile's accumulators complete

      if (p.csplit) {
        // cluster split-K: park the raw partial accumulators, column-major ([acc][column][row], rows adjacent: conflict-free
        // writes here and coalesced distributed-shared-memory reads in the reduction below), in this CTA's shared memory
        for (int a = 0; a < n_acc; ++a)
          for (int c = c_lo; c < c_hi; ++c) {
            float v[CW];
            if constexpr (CW == 32) tmem_ld32(lane_base + acc_col(a, par) + c * CW, v);
            else tmem_ld16(lane_base + acc_col(a, par) + c * CW, v);
            tmem_ld_wait();
            const uint32_t dst = sStage + static_cast<uint32_t>(((a * BN + c * CW) * 128 + m) * 4);
#pragma unroll
            for (int j = 0; j < CW; ++j) asm volatile("st.shared.f32 [%0], %1;" ::"r"(dst + j * 512), "f"(v[j]) : "memory");
          }
      } else if (p.ksplit > 1) {
        // partial sums: add the raw accumulators into the fp32 workspace; tapconv_finish_kernel applies the epilogue
        const long pix = ((static_cast<long>(seq) * p.T + t) * p.H + h) * p.W + w;
        for (int a = 0; a < n_acc; ++a) {
          // own accumulators: row = (set a, pixel); shared accumulator: stored after the n_out own sets
          float* dst_row = p.split_ws + (static_cast<long>(a) * rows_per_set + pix) * p.Cout;
          for (int c = c_lo; c < c_hi; ++c) {
            float v[CW];
            if constexpr (CW == 32) tmem_ld32(lane_base + acc_col(a, par) + c * CW, v);
            else tmem_ld16(lane_base + acc_col(a, par) + c * CW, v);
            tmem_ld_wait();
            const int col0 = n0 + c * CW;
            if (row_ok && ck_hi > ck_lo) {
#pragma unroll
              for (int j = 0; j < CW; j += 4)
                if (col0 + j + 4 <= p.Cout)
                  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst_row + col0 + j), "f"(v[j]), "f"(v[j + 1]),
                               "f"(v[j + 2]), "f"(v[j + 3])
                               : "memory");
            }
          }
        }
      } else
      for (int o = 0; o < p.n_out; ++o) {
        const long frame = static_cast<long>(seq * p.n_out + o) * p.T + t;
        float al = 1.f, be = 0.f;
        if (p.epi == EPI_GATED && row_ok) { al = p.alpha[frame]; be = p.beta[frame]; }
        const long row_off = ((frame * p.H + h) * p.W + w) * static_cast<long>(p.Cout);
        for (int c = c_lo; c < c_hi; ++c) {
          float own[CW], shr[CW];
          if constexpr (CW == 32) tmem_ld32(lane_base + acc_col(o, par) + c * CW, own);
          else tmem_ld16(lane_base + acc_col(o, par) + c * CW, own);
          if (p.epi == EPI_GATED) {
            if constexpr (CW == 32) tmem_ld32(lane_base + acc_col(p.n_out, par) + c * CW, shr);
            else tmem_ld16(lane_base + acc_col(p.n_out, par) + c * CW, shr);
          }
          tmem_ld_wait();
          if (o == p.n_out - 1 && c == c_hi - 1) {
            // last TMEM read of this tile: hand the accumulators back to the MMA issuer before the stores go out
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if constexpr (PAIR) mbar_arrive_remote(tmem_empty(par), 0);
              else mbar_arrive(tmem_empty(par));
            }
          }
          float y[CW];
          if (p.epi == EPI_GATED) {
#pragma unroll
            for (int j = 0; j < CW; ++j) y[j] = al * own[j] + be * shr[j];
          } else {
#pragma unroll
            for (int j = 0; j < CW; ++j) y[j] = own[j];
          }
          const int col0 = n0 + c * CW;
          if (p.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < CW; ++j) y[j] += (col0 + j < p.Cout) ? p.bias[col0 + j] : 0.f;
          }
          if (row_ok && p.post != POST_NONE) {
            // fused post-op on the fp32 result: second output
            float z[CW];
            if (p.post == POST_SCALE_SILU) {
              const float* cs = p.cscale + frame * p.cscale_ld + col0;
#pragma unroll
              for (int j = 0; j < CW; ++j) {
                const float v = y[j] * ((col0 + j < p.Cout) ? cs[j] : 0.f);
                z[j] = v / (1.f + __expf(-v)) * (1.f / 0.596f);
              }
            } else {
              const __nv_bfloat16* rr = static_cast<const __nv_bfloat16*>(p.res) + row_off + col0;
#pragma unroll
              for (int j = 0; j < CW; j += 8) {
                if (col0 + j + 8 <= p.Cout) {
                  const uint4 raw = *reinterpret_cast<const uint4*>(rr + j);
                  const uint32_t w4[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
                  for (int u = 0; u < 4; ++u) {
                    const float2 f2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w4[u]));
                    z[j + 2 * u] = f2.x; z[j + 2 * u + 1] = f2.y;
                  }
                } else {
#pragma unroll
                  for (int u = 0; u < 8; ++u) z[j + u] = 0.f;
                }
              }
#pragma unroll
              for (int j = 0; j < CW; ++j) {
                float v = z[j] * p.post_wa + y[j] * p.post_wb;
                if (p.post_clip > 0.f) v = fminf(fmaxf(v, -p.post_clip), p.post_clip);
                z[j] = v;
              }
            }
            __nv_bfloat16* dst2 = static_cast<__nv_bfloat16*>(p.out2) + row_off + col0;
#pragma unroll
            for (int j = 0; j < CW; j += 8)
              if (col0 + j + 8 <= p.Cout)
                *reinterpret_cast<uint4*>(dst2 + j) = make_uint4(pack_bf16x2(z[j], z[j + 1]), pack_bf16x2(z[j + 2], z[j + 3]),
                                                                 pack_bf16x2(z[j + 4], z[j + 5]), pack_bf16x2(z[j + 6], z[j + 7]));
          }
          if (row_ok && p.out != nullptr) {
            if (p.wide_store) {
              // 256-bit stores: one full 32-byte sector per thread and instruction (Cout % 16 == 0, 32-byte aligned bases)
              if (p.out_f32) {
                float* dst = static_cast<float*>(p.out) + row_off + col0;
#pragma unroll
                for (int j = 0; j < CW; j += 8)
                  if (col0 + j + 8 <= p.Cout) st_global_f32x8(dst + j, y + j);
              } else {
                __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(p.out) + row_off + col0;
#pragma unroll
                for (int j = 0; j < CW; j += 16)
                  if (col0 + j + 16 <= p.Cout) {
                    uint32_t pk[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) pk[u] = pack_bf16x2(y[j + 2 * u], y[j + 2 * u + 1]);
                    st_global_b32x8(dst + j, pk);
                  }
              }
              if (p.epi == EPI_GATED && p.out_d != nullptr) {
                __half* dd = static_cast<__half*>(p.out_d) + row_off + col0;
#pragma unroll
                for (int j = 0; j < CW; j += 16)
                  if (col0 + j + 16 <= p.Cout) {
                    uint32_t pk[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) pk[u] = pack_f16x2(shr[j + 2 * u] - own[j + 2 * u], shr[j + 2 * u + 1] - own[j + 2 * u + 1]);
                    st_global_b32x8(dd + j, pk);
                  }
              }
            } else {
            if (p.out_f32) {
              float* dst = static_cast<float*>(p.out) + row_off + col0;
#pragma unroll
              for (int j = 0; j < CW; j += 4)
                if (col0 + j + 4 <= p.Cout) *reinterpret_cast<float4*>(dst + j) = make_float4(y[j], y[j + 1], y[j + 2], y[j + 3]);
            } else {
              __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(p.out) + row_off + col0;
#pragma unroll
              for (int j = 0; j < CW; j += 8)
                if (col0 + j + 8 <= p.Cout)
                  *reinterpret_cast<uint4*>(dst + j) = make_uint4(pack_bf16x2(y[j], y[j + 1]), pack_bf16x2(y[j + 2], y[j + 3]),
                                                                  pack_bf16x2(y[j + 4], y[j + 5]), pack_bf16x2(y[j + 6], y[j + 7]));
            }
            if (p.epi == EPI_GATED && p.out_d != nullptr) {
              // fp16, not bf16: <dy, d> feeds the gate scalars' gradients, where bf16 rounding of d shows up at ~2 %; fp16
              // carries three more mantissa bits at the same two bytes (|d| is O(1): far inside fp16's range)
              __half* dd = static_cast<__half*>(p.out_d) + row_off + col0;
#pragma unroll
              for (int j = 0; j < CW; j += 8)
                if (col0 + j + 8 <= p.Cout)
                  *reinterpret_cast<uint4*>(dd + j) =
                      make_uint4(pack_f16x2(shr[j] - own[j], shr[j + 1] - own[j + 1]), pack_f16x2(shr[j + 2] - own[j + 2], shr[j + 3] - own[j + 3]),
                                 pack_f16x2(shr[j + 4] - own[j + 4], shr[j + 5] - own[j + 5]), pack_f16x2(shr[j + 6] - own[j + 6], shr[j + 7] - own[j + 7]));
            }
            }
          }
        }
      }
      if (p.ksplit > 1 || c_lo >= c_hi) {     // (a warp without columns -- one-chunk tiles -- only hands the accumulators back)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (PAIR) mbar_arrive_remote(tmem_empty(par), 0);
          else mbar_arrive(tmem_empty(par));
        }
      }
    }
    tc_fence_before();
    if (warp == 3) TAPCONV_STAMP(4);   // last epilogue's stores issued
  }

  __syncthreads();
  if constexpr (!PAIR) {
    if (p.csplit) {
      // ---- cluster split-K reduction: CTA r of the cluster finishes rows [r*128/ks, (r+1)*128/ks) of the tile.  Each of
      // its 128 epilogue threads owns one row and BN/ks columns, sums the ks partial values of every accumulator straight
      // out of the peers' shared memory and applies the ordinary output stage (gate combine, bf16 store, fp16 difference).
      cluster_sync_all();                       // every CTA's partials are staged and visible cluster-wide
      if (warp >= 3 && warp < 7) {
        const int ks = p.ksplit, rows_per = 128 / ks, cols_per = BN / ks;
        const int t128 = threadIdx.x - 96;
        const int m = static_cast<int>(blockIdx.y) * rows_per + t128 % rows_per;
        const int col_lo = (t128 / rows_per) * cols_per;
        const Tile tc = decode(blockIdx.x);
        const int hh = m / (p.bt * p.bw), rem = m - hh * (p.bt * p.bw), tt = rem / p.bw, ww = rem - tt * p.bw;
        const int t = tc.t0 + tt, h = tc.h0 + hh, w = tc.w0 + ww;
        const bool row_ok = (t < p.T) && (h < p.H) && (w < p.W) && (tc.seq < p.n_seq);
        uint32_t peer[8];
#pragma unroll
        for (int s = 0; s < 8; ++s) peer[s] = s < ks ? mapa_cluster(sStage, static_cast<uint32_t>(s)) : 0u;
        for (int c0 = col_lo; c0 < col_lo + cols_per; c0 += 8) {
          float shr[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) shr[j] = 0.f;
          if (p.epi == EPI_GATED) {
            for (int s = 0; s < ks; ++s) {
              const uint32_t src = peer[s] + static_cast<uint32_t>(((p.n_out * BN + c0) * 128 + m) * 4);
#pragma unroll
              for (int j = 0; j < 8; ++j) shr[j] += ld_cluster_f32(src + j * 512);
            }
          }
          for (int o = 0; o < p.n_out; ++o) {
            float own[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) own[j] = 0.f;
            for (int s = 0; s < ks; ++s) {
              const uint32_t src = peer[s] + static_cast<uint32_t>(((o * BN + c0) * 128 + m) * 4);
#pragma unroll
              for (int j = 0; j < 8; ++j) own[j] += ld_cluster_f32(src + j * 512);
            }
            const int col = tc.n0 + c0;
            if (!row_ok || col + 8 > p.Cout) continue;
            const long frame = static_cast<long>(tc.seq * p.n_out + o) * p.T + t;
            float al = 1.f, be = 0.f;
            if (p.epi == EPI_GATED) { al = p.alpha[frame]; be = p.beta[frame]; }
            const long off = ((frame * p.H + h) * p.W + w) * static_cast<long>(p.Cout) + col;
            float y[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) y[j] = al * own[j] + be * shr[j] + (p.bias != nullptr ? p.bias[col + j] : 0.f);
            if (p.out_f32) {
              float* dst = static_cast<float*>(p.out) + off;
              *reinterpret_cast<float4*>(dst) = make_float4(y[0], y[1], y[2], y[3]);
              *reinterpret_cast<float4*>(dst + 4) = make_float4(y[4], y[5], y[6], y[7]);
            } else {
              *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + off) =
                  make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7]));
            }
            if (p.epi == EPI_GATED && p.out_d != nullptr) {
              *reinterpret_cast<uint4*>(static_cast<__half*>(p.out_d) + off) =
                  make_uint4(pack_f16x2(shr[0] - own[0], shr[1] - own[1]), pack_f16x2(shr[2] - own[2], shr[3] - own[3]),
                             pack_f16x2(shr[4] - own[4], shr[5] - own[5]), pack_f16x2(shr[6] - own[6], shr[7] - own[7]));
            }
          }
        }
      }
      cluster_sync_all();                       // no CTA leaves while a peer may still read its staging area
    }
  }
  if constexpr (PAIR) cluster_sync_all();   // the peer has consumed every multicast arrival and drained its accumulators
  if (warp == 1) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc_pair(tmem_base, tmem_cols); else tmem_dealloc(tmem_base, tmem_cols);
  }
}

// Epilogue of a split-K launch: reads the reduced accumulators from the workspace and applies the same output stage
// as the fused epilogue (gate combine, bf16 / fp32 store, optional fp32 difference tensor).
static __global__ void __launch_bounds__(256) tapconv_finish_kernel(const float* __restrict__ ws, const float* __restrict__ alpha,
                                                             const float* __restrict__ beta, void* __restrict__ out,
                                                             __half* __restrict__ out_d, int n_seq, int n_out, int T,
                                                             long hw, int Cout, int epi, int out_f32,
                                                             const float* __restrict__ bias, int post, __nv_bfloat16* __restrict__ out2,
                                                             const float* __restrict__ cscale, int cscale_ld,
                                                             const __nv_bfloat16* __restrict__ res, float post_wa, float post_wb,
                                                             float post_clip) {
  pdl_launch_dependents();
  pdl_wait();
  const long rows_per_set = static_cast<long>(n_seq) * T * hw;
  const long vec = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;   // one float4 of one output row
  const int v_per_row = Cout >> 2;
  const long total = rows_per_set * n_out * v_per_row;
  if (vec >= total) return;
  const long orow = vec / v_per_row;                 // output row: ((seq*n_out + o)*T + t)*hw + px
  const int c = static_cast<int>(vec - orow * v_per_row) << 2;
  const long frame = orow / hw;
  const long px = orow - frame * hw;
  const long so = frame / T;                          // seq*n_out + o
  const int t = static_cast<int>(frame - so * T);
  const long sq = so / n_out;
  const int o = static_cast<int>(so - sq * n_out);
  const long pix = (sq * T + t) * hw + px;
  const float4 own = *reinterpret_cast<const float4*>(ws + (static_cast<long>(o) * rows_per_set + pix) * Cout + c);
  float4 y = own;
  if (epi == EPI_GATED) {
    const float4 shr = *reinterpret_cast<const float4*>(ws + (static_cast<long>(n_out) * rows_per_set + pix) * Cout + c);
    const float al = alpha[frame], be = beta[frame];
    y = make_float4(al * own.x + be * shr.x, al * own.y + be * shr.y, al * own.z + be * shr.z, al * own.w + be * shr.w);
    if (out_d) *reinterpret_cast<uint2*>(out_d + orow * Cout + c) = make_uint2(pack_f16x2(shr.x - own.x, shr.y - own.y), pack_f16x2(shr.z - own.z, shr.w - own.w));
  }
  if (bias != nullptr) { y.x += bias[c]; y.y += bias[c + 1]; y.z += bias[c + 2]; y.w += bias[c + 3]; }
  if (post != POST_NONE) {
    float4 z;
    if (post == POST_SCALE_SILU) {
      const float4 cs = *reinterpret_cast<const float4*>(cscale + frame * cscale_ld + c);
      const float v[4] = {y.x * cs.x, y.y * cs.y, y.z * cs.z, y.w * cs.w};
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = v[j] / (1.f + __expf(-v[j])) * (1.f / 0.596f);
      z = make_float4(o[0], o[1], o[2], o[3]);
    } else {
      const uint2 raw = *reinterpret_cast<const uint2*>(res + orow * Cout + c);
      const float2 r0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
      const float2 r1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
      float o[4] = {r0.x * post_wa + y.x * post_wb, r0.y * post_wa + y.y * post_wb, r1.x * post_wa + y.z * post_wb, r1.y * post_wa + y.w * post_wb};
      if (post_clip > 0.f) {
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = fminf(fmaxf(o[j], -post_clip), post_clip);
      }
      z = make_float4(o[0], o[1], o[2], o[3]);
    }
    *reinterpret_cast<uint2*>(out2 + orow * Cout + c) = make_uint2(pack_bf16x2(z.x, z.y), pack_bf16x2(z.z, z.w));
  }
  if (out == nullptr) return;
  if (out_f32) *reinterpret_cast<float4*>(static_cast<float*>(out) + orow * Cout + c) = y;
  else *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(out) + orow * Cout + c) = make_uint2(pack_bf16x2(y.x, y.y), pack_bf16x2(y.z, y.w));
}

}  // namespace ob
