// Tap-GEMM: the one tensor-core kernel behind every "K-major x K-major" contraction on the
// Oniris denoiser hot path (reference: edm2/conv.py:36-42 MPConv, :59-95 MPCausal3DGatedConv;
// their input-gradient passes reuse it with flipped/transposed weights).
//
//   acc_j[m, n] = sum over items i with acc(i)=j, over channels c:  A_i[m + shift_i, c] * Wg[n, wtap_i, c]
//
// * m walks a 128-row tile of output pixels laid out (frames bt) x (rows bh) x (cols bw) of an NHWC
//   activation tensor; an item's shift (dt,dy,dx) is applied through the TMA box coordinates, so spatial
//   zero padding and "frame past the end" are TMA out-of-bounds zero fill. No im2col buffer exists.
// * A tiles land in shared memory as [128 rows][CHUNK channels] with the hardware swizzle that matches
//   CHUNK*2 bytes; weights as [BN rows][CHUNK]. Both are K-major UMMA operands.
// * tcgen05.mma (cta_group::1, M=128, N=BN, K=16) accumulates in TMEM; up to 3 accumulators per tile
//   (clean rows, noised rows, shared causal-context term) so the context term of the DART training
//   sequence is computed once for both halves (edm2/conv.py:90-91 duplicates it instead).
// * Warp roles: warp0 = TMA producer, warp1 = TMEM owner + MMA issuer, warps 2..5 = epilogue
//   (TMEM -> registers -> gated combine -> global).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include "ptx.cuh"

namespace ob {

constexpr int TAPCONV_MAX_ITEMS = 32;
constexpr int TAPCONV_THREADS = 192;

enum : int { EPI_PLAIN = 0, EPI_GATED = 1 };

struct TapItem {
  int8_t src;      // which activation tensor map (0 or 1)
  int8_t dt;       // frame shift
  int8_t dy, dx;   // spatial shift
  int8_t n_a;      // number of A tiles sharing this weight tile (1, or 2 in dual mode)
  int8_t acc;      // accumulator of A tile 0 (tile i goes to acc+i)
  int8_t seq_mul;  // source sequence coordinate = seq*seq_mul + i
  int8_t pad_;
  int32_t wtap;    // column block of the weight matrix (in units of Cin)
};

struct TapConvParams {
  CUtensorMap mapA[2];
  CUtensorMap mapB;
  TapItem items[TAPCONV_MAX_ITEMS];
  int n_items;
  int n_seq, T, H, W;
  int Cin, Cout;
  int bw, bh, bt;
  int tiles_w, tiles_h, tiles_t, tiles_n;
  int n_out;    // output row sets per tile (2 in dual mode)
  int epi;      // EPI_*
  int out_f32;  // 0: bf16 out, 1: fp32 out
  const float* alpha;  // [n_seq*n_out*T] per output frame (EPI_GATED)
  const float* beta;
  void* out;    // [n_seq*n_out*T, H, W, Cout]
  void* out_d;  // optional (EPI_GATED): shared - own accumulator, fp32, same shape as out
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
  static constexpr int A_BYTES = 128 * ROW_BYTES;
  static constexpr int B_BYTES = BN * ROW_BYTES;
  static constexpr int B_BYTES_AL = (B_BYTES + 1023) / 1024 * 1024;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + B_BYTES_AL;  // room for two A tiles (dual mode)
  static constexpr int MAX_SMEM = 200 * 1024;
  static constexpr int STAGES_RAW = MAX_SMEM / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int CW = BN >= 32 ? 32 : 16;  // epilogue column chunk
};

// BMN=false: weights are [N rows][K contiguous] (forward convs).  BMN=true: weights are [K rows][N contiguous],
// i.e. the SAME forward weight matrix read as an MN-major B operand, which is what the input-gradient pass
// needs -- no transposed weight copy is ever materialised.
template <int CHUNK, int BN, bool BMN>
__global__ void __launch_bounds__(TAPCONV_THREADS, 1) tapconv_kernel(const __grid_constant__ TapConvParams p) {
  using Cfg = TapConvCfg<CHUNK, BN>;
  static_assert(!BMN || BN % CHUNK == 0, "MN-major weights need BN to be a multiple of CHUNK");
  constexpr int STAGES = Cfg::STAGES;
  constexpr uint32_t SWZ = SwizzleFor<CHUNK>::mode;
  constexpr uint32_t SBO = 8 * Cfg::ROW_BYTES;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES;
  // barriers: full[STAGES], empty[STAGES], tmem_full, then the TMEM base address slot
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- tile coordinates
  int tile = blockIdx.x;
  const int n_tile = tile % p.tiles_n;
  tile /= p.tiles_n;
  const int tw_i = tile % p.tiles_w;
  tile /= p.tiles_w;
  const int th_i = tile % p.tiles_h;
  tile /= p.tiles_h;
  const int tt_i = tile % p.tiles_t;
  const int seq = tile / p.tiles_t;
  const int w0 = tw_i * p.bw, h0 = th_i * p.bh, t0 = tt_i * p.bt;
  const int n0 = n_tile * BN;

  const int n_acc = p.n_out + (p.epi == EPI_GATED ? 1 : 0);
  uint32_t tmem_cols = 32;
  while (tmem_cols < static_cast<uint32_t>(n_acc * BN)) tmem_cols <<= 1;
  const int n_chunks = (p.Cin + CHUNK - 1) / CHUNK;  // a ragged last chunk is TMA zero-filled

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
    tma_prefetch_desc(&p.mapA[0]);
    tma_prefetch_desc(&p.mapA[1]);
    tma_prefetch_desc(&p.mapB);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < p.n_items; ++it) {
        const TapItem item = p.items[it];
        const void* mapA = &p.mapA[item.src];
        for (int ck = 0; ck < n_chunks; ++ck) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sA = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t sB = sA + 2 * Cfg::A_BYTES;
          mbar_arrive_expect_tx(full_bar(stage), item.n_a * Cfg::A_BYTES + Cfg::B_BYTES);
          for (int i = 0; i < item.n_a; ++i) {
            tma_load_5d(sA + i * Cfg::A_BYTES, mapA, full_bar(stage), ck * CHUNK, w0 + item.dx, h0 + item.dy,
                        t0 + item.dt, seq * item.seq_mul + i);
          }
          if constexpr (!BMN) {
            tma_load_2d(sB, &p.mapB, full_bar(stage), item.wtap * p.Cin + ck * CHUNK, n0);
          } else {
#pragma unroll
            for (int j = 0; j < BN / CHUNK; ++j)
              tma_load_2d(sB + j * (CHUNK * Cfg::ROW_BYTES), &p.mapB, full_bar(stage), item.wtap * p.Cout + n0 + j * CHUNK,
                          ck * CHUNK);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, BN, 0, BMN ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t started = 0;  // bit j: accumulator j already holds a partial sum
      for (int it = 0; it < p.n_items; ++it) {
        const TapItem item = p.items[it];
        for (int ck = 0; ck < n_chunks; ++ck) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sA = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t sB = sA + 2 * Cfg::A_BYTES;
          for (int i = 0; i < item.n_a; ++i) {
            const int acc = item.acc + i;
            const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
#pragma unroll
            for (int k = 0; k < CHUNK / 16; ++k) {
              const uint64_t adesc = make_smem_desc(sA + i * Cfg::A_BYTES + k * 32, 16, SBO, SWZ);
              const uint64_t bdesc = BMN ? make_smem_desc(sB + k * 2 * SBO, CHUNK * Cfg::ROW_BYTES, SBO, SWZ)
                                         : make_smem_desc(sB + k * 32, 16, SBO, SWZ);
              umma_bf16_ss(d_tmem, adesc, bdesc, idesc, (k > 0) || ((started >> acc) & 1u));
            }
            started |= 1u << acc;
          }
          umma_commit(empty_bar(stage));  // frees the smem slot once these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
      umma_commit(tmem_full_bar);  // all accumulators final
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    constexpr int CW = Cfg::CW;
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const int m = q * 32 + lane;
    const int tt = m / (p.bh * p.bw);
    const int rem = m - tt * (p.bh * p.bw);
    const int hh = rem / p.bw;
    const int ww = rem - hh * p.bw;
    const int t = t0 + tt, h = h0 + hh, w = w0 + ww;
    const bool row_ok = (t < p.T) && (h < p.H) && (w < p.W);

    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();

    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    for (int o = 0; o < p.n_out; ++o) {
      const long frame = static_cast<long>(seq * p.n_out + o) * p.T + t;
      float al = 1.f, be = 0.f;
      if (p.epi == EPI_GATED && row_ok) { al = p.alpha[frame]; be = p.beta[frame]; }
      const long row_off = ((frame * p.H + h) * p.W + w) * static_cast<long>(p.Cout);
      for (int c = 0; c < BN / CW; ++c) {
        float own[CW], shr[CW];
        if constexpr (CW == 32) tmem_ld32(lane_base + o * BN + c * CW, own);
        else tmem_ld16(lane_base + o * BN + c * CW, own);
        if (p.epi == EPI_GATED) {
          if constexpr (CW == 32) tmem_ld32(lane_base + p.n_out * BN + c * CW, shr);
          else tmem_ld16(lane_base + p.n_out * BN + c * CW, shr);
        }
        tmem_ld_wait();
        float y[CW];
        if (p.epi == EPI_GATED) {
#pragma unroll
          for (int j = 0; j < CW; ++j) y[j] = al * own[j] + be * shr[j];
        } else {
#pragma unroll
          for (int j = 0; j < CW; ++j) y[j] = own[j];
        }
        const int col0 = n0 + c * CW;
        if (row_ok) {
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
            // fp32 on purpose: <dy, d> feeds the gate scalars' gradients and bf16 rounding of d shows up there at ~2%
            float* dd = static_cast<float*>(p.out_d) + row_off + col0;
#pragma unroll
            for (int j = 0; j < CW; j += 4)
              if (col0 + j + 4 <= p.Cout)
                *reinterpret_cast<float4*>(dd + j) =
                    make_float4(shr[j] - own[j], shr[j + 1] - own[j + 1], shr[j + 2] - own[j + 2], shr[j + 3] - own[j + 3]);
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
