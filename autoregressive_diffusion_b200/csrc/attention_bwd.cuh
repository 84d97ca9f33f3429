// Backward of the frame-masked attention (what autograd of compiled FlexAttention / SDPA computes in the
// reference, edm2/attention/attention_modules.py:66,70,75), as two tcgen05 kernels that each recompute the
// probabilities from the saved log-sum-exp:
//
//   dq kernel : one CTA per 128-row query tile, loops over visible 64-row key tiles
//               S = Q K^T, dP = dO V^T  ->  dS = P*(dP - D)*scale  ->  dQ += dS K
//   dkv kernel: one CTA per 128-row key tile, loops over visible 64-row query tiles, transposed scores
//               S^T = K Q^T, dP^T = V dO^T  ->  P^T, dS^T  ->  dV += P^T dO,  dK += dS^T Q
//
// No atomics and no fp32 dQ scratch: every output element is owned by exactly one CTA.
// D[i] = sum_d dO[i,d]*O[i,d] comes from attn_bwd_prep_kernel.
#pragma once
#include "attention.cuh"
#include "launch.cuh"

namespace ob {

constexpr int ABW_BM = 128;  // rows owned by the CTA (queries for dq, keys for dkv)
constexpr int ABW_BN = 64;   // rows streamed per step

struct AttnBwdParams {
  CUtensorMap mapQ128, mapdO128, mapK64, mapV64;   // dq kernel
  CUtensorMap mapK128, mapV128, mapQ64, mapdO64;   // dkv kernel
  int BH, heads, Lq, Lk, hw, n_frames, mask;
  float scale;
  const float* lse;  // [BH, Lq]
  const float* dsum; // [BH, Lq]  D = rowsum(dO*O)
  __nv_bfloat16 *dq, *dk, *dv;
};

// Up to three disjoint, ascending ranges of streamed tiles.
struct TileRanges {
  int s[3], e[3], n;
  __device__ __forceinline__ void init() { n = 0; }
  __device__ __forceinline__ void add_tokens(int a, int b, int tile, int max_tiles) {  // token interval [a, b)
    if (b <= a) return;
    int ts = a / tile, te = min((b + tile - 1) / tile, max_tiles);
    if (te <= ts) return;
    if (n > 0 && ts <= e[n - 1]) { e[n - 1] = max(e[n - 1], te); return; }
    s[n] = ts; e[n] = te; ++n;
  }
  __device__ __forceinline__ int count() const {
    int c = 0;
    for (int i = 0; i < n; ++i) c += e[i] - s[i];
    return c;
  }
  __device__ __forceinline__ int tile(int j) const {
    for (int i = 0; i < n; ++i) {
      const int len = e[i] - s[i];
      if (j < len) return s[i] + j;
      j -= len;
    }
    return 0;
  }
};

// keys visible from query tokens [q0, q1)
__device__ __forceinline__ TileRanges visible_keys(const AttnBwdParams& p, int q0, int q1) {
  TileRanges r; r.init();
  const int tiles = (p.Lk + ABW_BN - 1) / ABW_BN;
  const int hw = p.hw, n = p.n_frames;
  const int qf_lo = q0 / hw, qf_hi = (q1 - 1) / hw;
  if (p.mask == ATTN_FULL) r.add_tokens(0, p.Lk, ABW_BN, tiles);
  else if (p.mask == ATTN_CAUSAL) r.add_tokens(0, (qf_hi + 1) * hw, ABW_BN, tiles);
  else {
    int end1;
    if (qf_hi < n) end1 = (qf_hi + 1) * hw;
    else { end1 = (qf_hi - n) * hw; if (qf_lo < n) end1 = max(end1, n * hw); }
    r.add_tokens(0, end1, ABW_BN, tiles);
    if (qf_hi >= n) r.add_tokens(max(qf_lo, n) * hw, (qf_hi + 1) * hw, ABW_BN, tiles);
  }
  return r;
}

// queries that see key tokens [k0, k1)
__device__ __forceinline__ TileRanges visible_queries(const AttnBwdParams& p, int k0, int k1) {
  TileRanges r; r.init();
  const int tiles = (p.Lq + ABW_BN - 1) / ABW_BN;
  const int hw = p.hw, n = p.n_frames;
  const int kf_lo = k0 / hw, kf_hi = (k1 - 1) / hw;
  if (p.mask == ATTN_FULL) r.add_tokens(0, p.Lq, ABW_BN, tiles);
  else if (p.mask == ATTN_CAUSAL) r.add_tokens(kf_lo * hw, p.Lq, ABW_BN, tiles);
  else {
    if (kf_lo < n) {
      // clean keys: clean queries of frames >= kf_lo (and, if the tile reaches into the noised half, those frames too)
      const int b1 = (kf_hi >= n) ? (kf_hi + 1) * hw : n * hw;
      r.add_tokens(kf_lo * hw, b1, ABW_BN, tiles);
      r.add_tokens((n + kf_lo + 1) * hw, 2 * n * hw, ABW_BN, tiles);  // noised queries strictly later than kf_lo
    } else {
      r.add_tokens(kf_lo * hw, (kf_hi + 1) * hw, ABW_BN, tiles);       // noised keys: only their own frames
    }
  }
  return r;
}

// D = rowsum(dO * O): one 8-lane group per (token, head) row of the [B, L, heads, 64] tensors; dsum is [B, heads, L].
__global__ void __launch_bounds__(256) attn_bwd_prep_kernel(const __nv_bfloat16* __restrict__ o,
                                                            const __nv_bfloat16* __restrict__ dout,
                                                            float* __restrict__ dsum, long rows, int L, int heads) {
  pdl_launch_dependents();
  pdl_wait();
  const long gid = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long row = gid >> 3;
  const int sub = static_cast<int>(gid & 7);
  float acc = 0.f;
  if (row < rows) {
    const uint4 a = *reinterpret_cast<const uint4*>(o + row * 64 + sub * 8);
    const uint4 b = *reinterpret_cast<const uint4*>(dout + row * 64 + sub * 8);
    const uint32_t av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
      acc += __uint_as_float(av[i] << 16) * __uint_as_float(bv[i] << 16) +
             __uint_as_float(av[i] & 0xffff0000u) * __uint_as_float(bv[i] & 0xffff0000u);
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  if (row < rows && sub == 0) {
    const long tok = row / heads;
    const int hh = static_cast<int>(row - tok * heads);
    const long bb = tok / L;
    dsum[(bb * heads + hh) * L + (tok - bb * L)] = acc;
  }
}

constexpr int ABW_STAGES = 3;
constexpr int ABW_T128 = 128 * 128;  // [128 rows][64 bf16]
constexpr int ABW_T64 = 64 * 128;    // [64 rows][64 bf16]
constexpr int ABW_SM_WARPS = 8;      // two softmax warps per TMEM lane quarter, 32 streamed columns each
constexpr int ABW_THREADS = 64 + 32 * ABW_SM_WARPS;
constexpr int ABW_DQ_SMEM = 1024 + 2 * ABW_T128 + ABW_STAGES * 2 * ABW_T64 + 2 * ABW_T128 + 256;
constexpr int ABW_DKV_SMEM = 1024 + 2 * ABW_T128 + ABW_STAGES * 2 * ABW_T64 + 4 * ABW_T128 + 2 * 2 * 64 * 4 + 256;

// store 32 packed bf16 columns [c*32, c*32+32) of row r into a K-major, 128B-swizzled [128][64] tile
__device__ __forceinline__ void store_row_chunk(uint32_t tile_base, int r, int c, const uint32_t (&packed)[16]) {
  const uint32_t row_base = tile_base + r * 128;
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
    const uint32_t chunk = static_cast<uint32_t>(c * 4 + ch) ^ static_cast<uint32_t>(r & 7);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_base + (chunk << 4)), "r"(packed[ch * 4]),
                 "r"(packed[ch * 4 + 1]), "r"(packed[ch * 4 + 2]), "r"(packed[ch * 4 + 3])
                 : "memory");
  }
}

// dS = P * (dP - D) * scale for 32 (query row, key) pairs of one row; P = exp2(s*c1 - lse2).
template <bool MASKED>
__device__ __forceinline__ void ds_chunk_row(const float (&s)[32], const float (&dp)[32], uint32_t (&packed)[16], float c1,
                                             float lse2, float dsum, float scale, int mask, int n_frames, int qf, int ik0,
                                             int Lk, int hw, int iq) {
  int kf = 0, rem = 0;
  if (MASKED) { kf = ik0 / hw; rem = ik0 - kf * hw; }
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    float ds[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const float arg = s[i + u] * c1 - lse2;
      float pr = ATTN_EXP_POLY(i + u) ? poly_exp2(arg) : fast_exp2(arg);
      if (MASKED) {
        bool ok = (ik0 + i + u < Lk) && frame_visible(mask, n_frames, qf, kf);
        if (mask == ATTN_DART_LISTED) ok = ok && block_listed(n_frames * hw, iq, ik0 + i + u);
        if (++rem == hw) { rem = 0; ++kf; }
        pr = ok ? pr : 0.f;
      }
      ds[u] = pr * (dp[i + u] - dsum) * scale;
    }
    packed[i >> 1] = pack_bf16x2(ds[0], ds[1]);
  }
}

// Transposed flavour (one key row, 32 query columns with their own lse / D read from shared memory).
template <bool MASKED>
__device__ __forceinline__ void pds_chunk_col(const float (&s)[32], const float (&dp)[32], uint32_t (&pk_p)[16],
                                              uint32_t (&pk_ds)[16], float c1, float scale, uint32_t stat_lse,
                                              uint32_t stat_d, int mask, int n_frames, int kf, int iq0, int Lq, int hw, int ik) {
  int qf = 0, rem = 0;
  if (MASKED) { qf = iq0 / hw; rem = iq0 - qf * hw; }
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    float l4[4], d4[4];
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(l4[0]), "=f"(l4[1]), "=f"(l4[2]), "=f"(l4[3]) : "r"(stat_lse + i * 4));
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(d4[0]), "=f"(d4[1]), "=f"(d4[2]), "=f"(d4[3]) : "r"(stat_d + i * 4));
    float pv[4], ds[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float arg = s[i + u] * c1 - l4[u];
      float pr = ATTN_EXP_POLY(i + u) ? poly_exp2(arg) : fast_exp2(arg);
      if (MASKED) {
        bool ok = (iq0 + i + u < Lq) && frame_visible(mask, n_frames, qf, kf);
        if (mask == ATTN_DART_LISTED) ok = ok && block_listed(n_frames * hw, iq0 + i + u, ik);
        if (++rem == hw) { rem = 0; ++qf; }
        pr = ok ? pr : 0.f;
      }
      pv[u] = pr;
      ds[u] = pr * (dp[i + u] - d4[u]) * scale;
    }
    pk_p[i >> 1] = pack_bf16x2(pv[0], pv[1]);
    pk_p[(i >> 1) + 1] = pack_bf16x2(pv[2], pv[3]);
    pk_ds[i >> 1] = pack_bf16x2(ds[0], ds[1]);
    pk_ds[(i >> 1) + 1] = pack_bf16x2(ds[2], ds[3]);
  }
}

// ------------------------------------------------------------------------------------------------ dQ
__global__ void __launch_bounds__(ABW_THREADS, 1) attn_bwd_dq_kernel(const __grid_constant__ AttnBwdParams p) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = base, sdO = sQ + ABW_T128;
  const uint32_t sKV = sdO + ABW_T128;                       // stage s: K at +s*16K, V at +s*16K+8K
  const uint32_t sdS = sKV + ABW_STAGES * 2 * ABW_T64;       // 2 buffers of [128][64]
  const uint32_t bar = sdS + 2 * ABW_T128;
  const uint32_t q_full = bar;
  auto kv_full = [&](int s) { return bar + 8u * (1 + s); };
  auto kv_empty = [&](int s) { return bar + 8u * (4 + s); };
  auto sdp_full = [&](int b) { return bar + 8u * (7 + b); };
  auto sdp_empty = [&](int b) { return bar + 8u * (9 + b); };
  auto ds_full = [&](int b) { return bar + 8u * (11 + b); };
  auto ds_empty = [&](int b) { return bar + 8u * (13 + b); };
  const uint32_t acc_full = bar + 8u * 15;
  const uint32_t tmem_slot = bar + 8u * 16;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int bb = bh / p.heads, hh = bh - bb * p.heads;
  const int q0 = blockIdx.x * ABW_BM;
  const TileRanges kr = visible_keys(p, q0, min(q0 + ABW_BM, p.Lq));
  const int n_kv = kr.count();

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < ABW_STAGES; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(sdp_full(b), 1); mbar_init(sdp_empty(b), ABW_SM_WARPS); mbar_init(ds_full(b), ABW_SM_WARPS); mbar_init(ds_empty(b), 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot));
  // S buffers: cols [0,64) [64,128); dP buffers: [128,192) [192,256); dQ: [256,320)

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, 2 * ABW_T128);
      tma_load_4d(sQ, &p.mapQ128, q_full, 0, q0, hh, bb);
      tma_load_4d(sdO, &p.mapdO128, q_full, 0, q0, hh, bb);
    }
    __syncwarp();
    for (int j = 0; j < n_kv; ++j) {
      const int st = j % ABW_STAGES;
      mbar_wait(kv_empty(st), ((j / ABW_STAGES) & 1) ^ 1);
      if (elect_one()) {
        const uint32_t sK = sKV + st * 2 * ABW_T64, sV = sK + ABW_T64;
        mbar_arrive_expect_tx(kv_full(st), 2 * ABW_T64);
        const int k0 = kr.tile(j) * ABW_BN;
        tma_load_4d(sK, &p.mapK64, kv_full(st), 0, k0, hh, bb);
        tma_load_4d(sV, &p.mapV64, kv_full(st), 0, k0, hh, bb);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    if (n_kv > 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, ABW_BN, 0, 0);
      constexpr uint32_t idesc_q = make_idesc_bf16(128, ATTN_D, 0, 1);
      const uint64_t kdesc0 = make_smem_desc(0, 16, 1024, SWZ_128B);
      const uint64_t mdesc0 = make_smem_desc(0, ABW_T64, 1024, SWZ_128B);
      auto issue_sdp = [&](int j) {
        const int st = j % ABW_STAGES, b = j & 1;
        mbar_wait(kv_full(st), (j / ABW_STAGES) & 1);
        if (j >= 2) mbar_wait(sdp_empty(b), ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sK = sKV + st * 2 * ABW_T64, sV = sK + ABW_T64;
          const uint64_t qd = kdesc0 + (sQ >> 4), dod = kdesc0 + (sdO >> 4), kd = kdesc0 + (sK >> 4), vd = kdesc0 + (sV >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem + b * 64, qd + 2 * k, kd + 2 * k, idesc_s, k > 0);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem + 128 + b * 64, dod + 2 * k, vd + 2 * k, idesc_s, k > 0);
          umma_commit(sdp_full(b));
        }
        __syncwarp();
      };
      mbar_wait(q_full, 0);
      issue_sdp(0);
      for (int j = 0; j < n_kv; ++j) {
        if (j + 1 < n_kv) issue_sdp(j + 1);
        const int st = j % ABW_STAGES, b = j & 1;
        mbar_wait(ds_full(b), (j >> 1) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t dsd = kdesc0 + ((sdS + b * ABW_T128) >> 4);
          const uint64_t kd = mdesc0 + ((sKV + st * 2 * ABW_T64) >> 4);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)   // dQ += dS[128 x 64 keys] * K[64 keys x 64]
            umma_bf16_ss(tmem + 256, dsd + 2 * kk, kd + kk * (2048 >> 4), idesc_q, (j > 0) || (kk > 0));
          umma_commit(kv_empty(st));
          umma_commit(ds_empty(b));
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(acc_full);
      __syncwarp();
    }
  } else {
    const int qw = warp & 3;
    const int half = (warp - 2) >> 2;      // which 32 of the 64 streamed key columns this warp owns
    const int r = qw * 32 + lane;
    const int iq = q0 + r;
    const int qf = iq / p.hw;
    const uint32_t lane_off = static_cast<uint32_t>(qw * 32) << 16;
    const float LOG2E = 1.4426950408889634f;
    const float c1 = p.scale * LOG2E;
    float lse2 = 0.f, dsum = 0.f;
    if (iq < p.Lq) {
      lse2 = p.lse[static_cast<long>(bh) * p.Lq + iq] * LOG2E;
      dsum = p.dsum[static_cast<long>(bh) * p.Lq + iq];
    }
    for (int j = 0; j < n_kv; ++j) {
      const int b = j & 1;
      const int k0 = kr.tile(j) * ABW_BN;
      mbar_wait(sdp_full(b), (j >> 1) & 1);
      tc_fence_after();
      if (j >= 2) mbar_wait(ds_empty(b), ((j >> 1) & 1) ^ 1);
      float s[32], dp[32];
      tmem_ld32(tmem + lane_off + b * 64 + half * 32, s);
      tmem_ld32(tmem + lane_off + 128 + b * 64 + half * 32, dp);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(sdp_empty(b));
      uint32_t packed[16];
      const int ik0 = k0 + half * 32;
      bool all_vis = (iq < p.Lq) && (ik0 + 32 <= p.Lk);
      {
        const int kf_a = ik0 / p.hw, kf_b = (ik0 + 31) / p.hw;
        if (p.mask == ATTN_CAUSAL) all_vis = all_vis && (kf_b <= qf);
        else if (p.mask == ATTN_DART)
          all_vis = all_vis && ((qf < p.n_frames) ? (kf_b <= qf) : ((kf_b < qf - p.n_frames) || (kf_a == qf && kf_b == qf)));
        else if (p.mask == ATTN_DART_LISTED) all_vis = all_vis && (qf < p.n_frames) && (kf_b <= qf);
      }
      if (all_vis) ds_chunk_row<false>(s, dp, packed, c1, lse2, dsum, p.scale, p.mask, p.n_frames, qf, ik0, p.Lk, p.hw, iq);
      else if (iq < p.Lq) ds_chunk_row<true>(s, dp, packed, c1, lse2, dsum, p.scale, p.mask, p.n_frames, qf, ik0, p.Lk, p.hw, iq);
      else {
#pragma unroll
        for (int i = 0; i < 16; ++i) packed[i] = 0u;
      }
      store_row_chunk(sdS + b * ABW_T128, r, half, packed);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(ds_full(b));
    }
    if (n_kv > 0) { mbar_wait(acc_full, 0); tc_fence_after(); }
    __nv_bfloat16* drow = p.dq + ((static_cast<long>(bb) * p.Lq + iq) * p.heads + hh) * ATTN_D + half * 32;
    {
      float o[32];
      if (n_kv > 0) { tmem_ld32(tmem + lane_off + 256 + half * 32, o); tmem_ld_wait(); }
      else {
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = 0.f;
      }
      if (iq < p.Lq) {
#pragma unroll
        for (int i = 0; i < 32; i += 8)
          *reinterpret_cast<uint4*>(drow + i) = make_uint4(pack_bf16x2(o[i], o[i + 1]), pack_bf16x2(o[i + 2], o[i + 3]),
                                                           pack_bf16x2(o[i + 4], o[i + 5]), pack_bf16x2(o[i + 6], o[i + 7]));
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// ------------------------------------------------------------------------------------------------ dK, dV
__global__ void __launch_bounds__(ABW_THREADS, 1) attn_bwd_dkv_kernel(const __grid_constant__ AttnBwdParams p) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sK = base, sV = sK + ABW_T128;
  const uint32_t sQdO = sV + ABW_T128;                       // stage s: Q at +s*16K, dO at +s*16K+8K
  const uint32_t sP = sQdO + ABW_STAGES * 2 * ABW_T64;       // P^T buffers 0,1 then dS^T buffers 0,1, each [128][64]
  const uint32_t sStat = sP + 4 * ABW_T128;                  // [2 buffers][lse(64) | D(64)] floats
  const uint32_t bar = sStat + 2 * 2 * 64 * 4;
  const uint32_t kv_full = bar;
  auto q_full = [&](int s) { return bar + 8u * (1 + s); };
  auto q_empty = [&](int s) { return bar + 8u * (4 + s); };
  auto sdp_full = [&](int b) { return bar + 8u * (7 + b); };
  auto sdp_empty = [&](int b) { return bar + 8u * (9 + b); };
  auto pds_full = [&](int b) { return bar + 8u * (11 + b); };
  auto pds_empty = [&](int b) { return bar + 8u * (13 + b); };
  const uint32_t acc_full = bar + 8u * 15;
  const uint32_t tmem_slot = bar + 8u * 16;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int bb = bh / p.heads, hh = bh - bb * p.heads;
  const int k0 = blockIdx.x * ABW_BM;
  const TileRanges qr = visible_queries(p, k0, min(k0 + ABW_BM, p.Lk));
  const int n_q = qr.count();

  if (threadIdx.x == 0) {
    mbar_init(kv_full, 1);
    for (int s = 0; s < ABW_STAGES; ++s) { mbar_init(q_full(s), 1); mbar_init(q_empty(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(sdp_full(b), 1); mbar_init(sdp_empty(b), ABW_SM_WARPS); mbar_init(pds_full(b), ABW_SM_WARPS); mbar_init(pds_empty(b), 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot));
  // S^T buffers: cols [0,64) [64,128); dP^T: [128,192) [192,256); dK: [256,320); dV: [320,384)

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(kv_full, 2 * ABW_T128);
      tma_load_4d(sK, &p.mapK128, kv_full, 0, k0, hh, bb);
      tma_load_4d(sV, &p.mapV128, kv_full, 0, k0, hh, bb);
    }
    __syncwarp();
    for (int j = 0; j < n_q; ++j) {
      const int st = j % ABW_STAGES;
      mbar_wait(q_empty(st), ((j / ABW_STAGES) & 1) ^ 1);
      if (elect_one()) {
        const uint32_t sQ = sQdO + st * 2 * ABW_T64, sdO = sQ + ABW_T64;
        mbar_arrive_expect_tx(q_full(st), 2 * ABW_T64);
        const int q0 = qr.tile(j) * ABW_BN;
        tma_load_4d(sQ, &p.mapQ64, q_full(st), 0, q0, hh, bb);
        tma_load_4d(sdO, &p.mapdO64, q_full(st), 0, q0, hh, bb);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    if (n_q > 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, ABW_BN, 0, 0);
      constexpr uint32_t idesc_a = make_idesc_bf16(128, ATTN_D, 0, 1);
      const uint64_t kdesc0 = make_smem_desc(0, 16, 1024, SWZ_128B);
      const uint64_t mdesc0 = make_smem_desc(0, ABW_T64, 1024, SWZ_128B);
      auto issue_sdp = [&](int j) {
        const int st = j % ABW_STAGES, b = j & 1;
        mbar_wait(q_full(st), (j / ABW_STAGES) & 1);
        if (j >= 2) mbar_wait(sdp_empty(b), ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sQ = sQdO + st * 2 * ABW_T64, sdO = sQ + ABW_T64;
          const uint64_t kd = kdesc0 + (sK >> 4), vd = kdesc0 + (sV >> 4), qd = kdesc0 + (sQ >> 4), dod = kdesc0 + (sdO >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem + b * 64, kd + 2 * k, qd + 2 * k, idesc_s, k > 0);          // S^T = K Q^T
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem + 128 + b * 64, vd + 2 * k, dod + 2 * k, idesc_s, k > 0);   // dP^T = V dO^T
          umma_commit(sdp_full(b));
        }
        __syncwarp();
      };
      mbar_wait(kv_full, 0);
      issue_sdp(0);
      for (int j = 0; j < n_q; ++j) {
        if (j + 1 < n_q) issue_sdp(j + 1);
        const int st = j % ABW_STAGES, b = j & 1;
        mbar_wait(pds_full(b), (j >> 1) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sQ = sQdO + st * 2 * ABW_T64, sdO = sQ + ABW_T64;
          const uint64_t pd = kdesc0 + ((sP + b * ABW_T128) >> 4), dsd = kdesc0 + ((sP + (2 + b) * ABW_T128) >> 4);
          const uint64_t dod = mdesc0 + (sdO >> 4), qd = mdesc0 + (sQ >> 4);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)   // dV += P^T[128 keys x 64 q] * dO[64 q x 64]
            umma_bf16_ss(tmem + 320, pd + 2 * kk, dod + kk * (2048 >> 4), idesc_a, (j > 0) || (kk > 0));
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)   // dK += dS^T * Q
            umma_bf16_ss(tmem + 256, dsd + 2 * kk, qd + kk * (2048 >> 4), idesc_a, (j > 0) || (kk > 0));
          umma_commit(q_empty(st));
          umma_commit(pds_empty(b));
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(acc_full);
      __syncwarp();
    }
  } else {
    const int qw = warp & 3;
    const int half = (warp - 2) >> 2;     // which 32 of the 64 streamed query columns this warp owns
    const int r = qw * 32 + lane;         // key row of the tile
    const int tid = threadIdx.x - 64;     // 0..255 among the softmax threads
    const int ik = k0 + r;
    const int kf = ik / p.hw;
    const uint32_t lane_off = static_cast<uint32_t>(qw * 32) << 16;
    const float LOG2E = 1.4426950408889634f;
    const float c1 = p.scale * LOG2E;
    for (int j = 0; j < n_q; ++j) {
      const int b = j & 1;
      const int q0 = qr.tile(j) * ABW_BN;
      {  // stage this step's 64 (lse, D) pairs; the named barrier also orders reuse of the buffer (see header note)
        if (tid < 128) {
          const int i = tid & 63;
          const long gi = static_cast<long>(bh) * p.Lq + q0 + i;
          float v = 0.f;
          if (q0 + i < p.Lq) v = (tid < 64) ? p.lse[gi] * LOG2E : p.dsum[gi];
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(sStat + (b * 128 + tid) * 4), "f"(v) : "memory");
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      mbar_wait(sdp_full(b), (j >> 1) & 1);
      tc_fence_after();
      if (j >= 2) mbar_wait(pds_empty(b), ((j >> 1) & 1) ^ 1);
      float s[32], dp[32];
      tmem_ld32(tmem + lane_off + b * 64 + half * 32, s);
      tmem_ld32(tmem + lane_off + 128 + b * 64 + half * 32, dp);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(sdp_empty(b));
      uint32_t pk_p[16], pk_ds[16];
      const int iq0 = q0 + half * 32;
      bool all_vis = (ik < p.Lk) && (iq0 + 32 <= p.Lq);
      {
        const int qf_a = iq0 / p.hw, qf_b = (iq0 + 31) / p.hw, n = p.n_frames;
        if (p.mask == ATTN_CAUSAL) all_vis = all_vis && (kf <= qf_a);
        else if (p.mask == ATTN_DART) {
          if (kf < n) all_vis = all_vis && ((qf_b < n && kf <= qf_a) || (qf_a >= n && kf < qf_a - n));
          else all_vis = all_vis && (qf_a == kf && qf_b == kf);
        } else if (p.mask == ATTN_DART_LISTED) {
          all_vis = all_vis && (kf < n) && (qf_b < n) && (kf <= qf_a);   // clean x clean only; the rest is tested per element
        }
      }
      const uint32_t st_l = sStat + (b * 128 + half * 32) * 4, st_d = sStat + (b * 128 + 64 + half * 32) * 4;
      if (all_vis) pds_chunk_col<false>(s, dp, pk_p, pk_ds, c1, p.scale, st_l, st_d, p.mask, p.n_frames, kf, iq0, p.Lq, p.hw, ik);
      else if (ik < p.Lk) pds_chunk_col<true>(s, dp, pk_p, pk_ds, c1, p.scale, st_l, st_d, p.mask, p.n_frames, kf, iq0, p.Lq, p.hw, ik);
      else {
#pragma unroll
        for (int i = 0; i < 16; ++i) { pk_p[i] = 0u; pk_ds[i] = 0u; }
      }
      store_row_chunk(sP + b * ABW_T128, r, half, pk_p);
      store_row_chunk(sP + (2 + b) * ABW_T128, r, half, pk_ds);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_full(b));
    }
    if (n_q > 0) { mbar_wait(acc_full, 0); tc_fence_after(); }
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      __nv_bfloat16* drow = (which == 0 ? p.dk : p.dv) + ((static_cast<long>(bb) * p.Lk + ik) * p.heads + hh) * ATTN_D + half * 32;
      float o[32];
      if (n_q > 0) { tmem_ld32(tmem + lane_off + 256 + which * 64 + half * 32, o); tmem_ld_wait(); }
      else {
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = 0.f;
      }
      if (ik < p.Lk) {
#pragma unroll
        for (int i = 0; i < 32; i += 8)
          *reinterpret_cast<uint4*>(drow + i) = make_uint4(pack_bf16x2(o[i], o[i + 1]), pack_bf16x2(o[i + 2], o[i + 3]),
                                                           pack_bf16x2(o[i + 4], o[i + 5]), pack_bf16x2(o[i + 6], o[i + 7]));
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

}  // namespace ob
