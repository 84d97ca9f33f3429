// Backward of the frame-masked attention (what autograd of compiled FlexAttention / SDPA computes in the
// reference, edm2/attention/attention_modules.py:66,70,75), as two tcgen05 kernels that each recompute the
// probabilities from the saved log-sum-exp:
//
//   dq kernel : one CTA per 128-row query tile, loops over visible 128-row key tiles
//               S = Q K^T, dP = dO V^T  ->  dS = P*(dP - D)*scale  ->  dQ += dS K
//   dkv kernel: one CTA per 128-row key tile, loops over visible 128-row query tiles, transposed scores
//               S^T = K Q^T, dP^T = V dO^T  ->  P^T, dS^T  ->  dV += P^T dO,  dK += dS^T Q
//
// No atomics and no fp32 dQ scratch: every output element is owned by exactly one CTA.
// D[i] = sum_d dO[i,d]*O[i,d] comes from attn_bwd_prep_kernel.
#pragma once
#include "attention.cuh"
#include "launch.cuh"

namespace ob {

constexpr int ABW_BM = 128;  // rows owned by the CTA (queries for dq, keys for dkv)
constexpr int ABW_BN = 128;  // rows streamed per step (64 before: every step pays ~60 fixed instructions per softmax warp and three
                             // barrier round trips, which 128-wide steps halve per element)

struct AttnBwdParams {
  CUtensorMap mapQ128, mapdO128, mapK128, mapV128;   // 4-D maps (64, L, heads, B) with a 128-row box: own and streamed tiles
  int BH, heads, Lq, Lk, hw, n_frames, mask;
  float scale;
  const float* lse;  // [BH, Lq]
  const float* ws;   // [2][BH][Lp] pre-scaled row statistics written by attn_bwd_prep_kernel (Lp = Lq rounded up to 64)
  __nv_bfloat16 *dq, *dk, *dv;
};

// Up to three disjoint, ascending ranges of streamed tiles.  Scalar members and if-chains only: dynamically indexed
// arrays put the struct in local memory, and tile(j) is evaluated by every softmax warp on every step (ncu: 5 % of the
// dK/dV kernel's stall samples sat on its local loads).
struct TileRanges {
  int s0, e0, s1, e1, s2, e2, n;
  __device__ __forceinline__ void init() { n = 0; s0 = e0 = s1 = e1 = s2 = e2 = 0; }
  __device__ __forceinline__ void add_tokens(int a, int b, int tile, int max_tiles) {  // token interval [a, b)
    if (b <= a) return;
    const int ts = a / tile, te = min((b + tile - 1) / tile, max_tiles);
    if (te <= ts) return;
    if (n == 0) { s0 = ts; e0 = te; n = 1; }
    else if (n == 1) {
      if (ts <= e0) e0 = max(e0, te);
      else { s1 = ts; e1 = te; n = 2; }
    } else if (n == 2) {
      if (ts <= e1) e1 = max(e1, te);
      else { s2 = ts; e2 = te; n = 3; }
    } else if (ts <= e2) e2 = max(e2, te);
  }
  __device__ __forceinline__ int count() const { return (e0 - s0) + (e1 - s1) + (e2 - s2); }
  __device__ __forceinline__ int tile(int j) const {
    const int l0 = e0 - s0, l1 = e1 - s1;
    return j < l0 ? s0 + j : (j < l0 + l1 ? s1 + (j - l0) : s2 + (j - l0 - l1));
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

// Per-row statistics of the backward pass, pre-scaled so the hot loops use them as FMA addends:
//   ws[0][bh][i] = -D[i] * scale   with D = rowsum(dO * O)          (dS = P * (dP*scale - D*scale))
//   ws[1][bh][i] = -lse[i] * log2(e)                                 (P = exp2(S*scale*log2e - lse*log2e))
// Rows are padded to Lp = ceil(L / 128) * 128 floats (padding zero-filled) so the dK/dV kernel can fetch the 128 values of a
// streamed query tile with one aligned bulk copy.  One 8-lane group per (token, head) row of the [B, L, heads, 64] tensors.
__global__ void __launch_bounds__(256) attn_bwd_prep_kernel(const __nv_bfloat16* __restrict__ o,
                                                            const __nv_bfloat16* __restrict__ dout,
                                                            const float* __restrict__ lse, float* __restrict__ ws, long rows,
                                                            int L, int Lp, int heads, long BH, float scale) {
  pdl_launch_dependents();
  pdl_wait();
  const long gid = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long row = gid >> 3;                 // over [B, Lp, heads]: padded positions write zeros
  const int sub = static_cast<int>(gid & 7);
  const bool valid = row < rows;
  const long tok_p = (valid ? row : 0) / heads;
  const int hh = static_cast<int>((valid ? row : 0) - tok_p * heads);
  const long bb = tok_p / Lp;
  const int i = static_cast<int>(tok_p - bb * Lp);
  float acc = 0.f;
  if (valid && i < L) {
    const long src = ((bb * L + i) * heads + hh) * 64 + sub * 8;
    const uint4 a = *reinterpret_cast<const uint4*>(o + src);
    const uint4 b = *reinterpret_cast<const uint4*>(dout + src);
    const uint32_t av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int t = 0; t < 4; ++t)
      acc += __uint_as_float(av[t] << 16) * __uint_as_float(bv[t] << 16) +
             __uint_as_float(av[t] & 0xffff0000u) * __uint_as_float(bv[t] & 0xffff0000u);
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  if (valid && sub == 0) {
    const long bh = bb * heads + hh;
    const long dst = bh * Lp + i;
    ws[dst] = -acc * scale;
    ws[BH * Lp + dst] = i < L ? -lse[bh * L + i] * 1.4426950408889634f : 0.f;
  }
}

constexpr int ABW_STAGES = 4;
constexpr int ABW_T128 = 128 * 128;  // [128 rows][64 bf16]
constexpr int ABW_SM_WARPS = 16;     // four softmax warps per TMEM lane quarter, 32 streamed columns each (as two chunks of 16)
constexpr int ABW_COLS = ABW_BN / (ABW_SM_WARPS / 4);
constexpr int ABW_THREADS = 96 + 32 * ABW_SM_WARPS;   // TMA, two MMA issuers, softmax warps
constexpr int ABW_DS_BUFS = 2;       // dS buffers of the dQ kernel in tensor memory
constexpr int ABW_STAT_BYTES = 2 * ABW_BN * 4;        // per stage of the dK/dV kernel: -lse*log2e | -D*scale of 128 queries
constexpr int ABW_DQ_SMEM = 1024 + 2 * ABW_T128 + ABW_STAGES * 2 * ABW_T128 + 512;
constexpr int ABW_DKV_SMEM = 1024 + 2 * ABW_T128 + ABW_STAGES * (2 * ABW_T128 + ABW_STAT_BYTES) + 512;

// Visibility of the streamed index (keys for a query row, queries for a key row) as two windows [a1, b1) u [a2, b2):
// every mask of this file has that shape per row, so tile and element tests are integer compares.
struct Win2 {
  int a1, b1, a2, b2;
  __device__ __forceinline__ bool whole(int c0, int len) const {
    return (c0 >= a1 && c0 + len <= b1) || (c0 >= a2 && c0 + len <= b2);
  }
};
// keys seen by query iq
__device__ __forceinline__ Win2 key_windows(int mask, int n_frames, int hw, int iq, int Lq, int Lk) {
  Win2 w;
  if (iq >= Lq) { w.a1 = w.b1 = w.a2 = w.b2 = 0; return w; }
  const RowWindows r = row_windows(mask, n_frames, hw, iq, Lk);
  w.a1 = 0; w.b1 = r.w1_end; w.a2 = r.w2_lo; w.b2 = max(r.w2_hi, r.w2_lo);
  return w;
}
// queries that see key ik (the transpose of row_windows, including the ATTN_DART_LISTED block rule)
__device__ __forceinline__ Win2 query_windows(int mask, int n_frames, int hw, int ik, int Lq, int Lk) {
  Win2 w;
  w.a1 = w.b1 = w.a2 = w.b2 = 0;
  if (ik >= Lk) return w;
  const int kf = ik / hw;
  if (mask == ATTN_FULL) { w.b1 = Lq; return w; }
  if (mask == ATTN_CAUSAL) { w.a1 = kf * hw; w.b1 = Lq; return w; }
  const int n_hw = n_frames * hw;
  if (kf < n_frames) {
    w.a1 = kf * hw; w.b1 = n_hw;                       // clean queries of frames >= kf
    w.a2 = (n_frames + kf + 1) * hw; w.b2 = 2 * n_hw;  // noised queries of strictly later frames
    if (mask == ATTN_DART_LISTED) w.a2 = max(w.a2, n_hw + (((ik >> 7) + 1) << 7));
  } else {
    w.a1 = kf * hw; w.b1 = w.a1 + hw;                  // noised keys: their own frame's queries
    if (mask == ATTN_DART_LISTED) { const int blk = (ik >> 7) << 7; w.a1 = max(w.a1, blk); w.b1 = min(w.b1, blk + 128); }
  }
  w.b1 = max(min(w.b1, Lq), w.a1);
  w.b2 = max(min(w.b2, Lq), w.a2);
  return w;
}

// which pairs of a 16-element chunk take the FMA-pipe exponential in the backward kernels (tuned separately from the forward)
#ifndef ABW_PAIR_POLY
#define ABW_PAIR_POLY(pi) (((pi) % 3) == 2)
#endif

// One 16-element chunk of a row: P = exp2(s*c1 + nl), dS = P * (dp*scale + nd).  nl / nd are pairs (per column) so the
// same code serves the dQ kernel (row statistics, broadcast) and the dK/dV kernel (column statistics).  MASKED: elements
// outside the row's windows (given relative to the chunk's first column) are zeroed.
template <bool MASKED, bool WANT_P, typename NL, typename ND>
__device__ __forceinline__ void pds_chunk16(const float (&s)[16], const float (&dp)[16], uint32_t (&pk_p)[8],
                                            uint32_t (&pk_ds)[8], float c1, float scale, NL nl, ND nd, int rel_a1,
                                            unsigned len1, int rel_a2, unsigned len2) {
  const uint64_t C1 = pack2(c1, c1), SC = pack2(scale, scale);
#pragma unroll
  for (int pi = 0; pi < 8; ++pi) {
    const uint64_t arg = fma2(pack2(s[2 * pi], s[2 * pi + 1]), C1, nl(pi));
    float e0, e1;
    if (ABW_PAIR_POLY(pi)) poly_exp2_pair(arg, e0, e1);
    else {
      float a0, a1;
      unpack2(arg, a0, a1);
      e0 = fast_exp2(a0);
      e1 = fast_exp2(a1);
    }
    if (MASKED) {
      const int i0 = 2 * pi, i1 = 2 * pi + 1;
      e0 = (static_cast<unsigned>(i0 - rel_a1) < len1 || static_cast<unsigned>(i0 - rel_a2) < len2) ? e0 : 0.f;
      e1 = (static_cast<unsigned>(i1 - rel_a1) < len1 || static_cast<unsigned>(i1 - rel_a2) < len2) ? e1 : 0.f;
    }
    const uint64_t t = fma2(pack2(dp[2 * pi], dp[2 * pi + 1]), SC, nd(pi));
    float d0, d1;
    unpack2(mul2(pack2(e0, e1), t), d0, d1);
    if (WANT_P) pk_p[pi] = pack_bf16x2(e0, e1);
    pk_ds[pi] = pack_bf16x2(d0, d1);
  }
}

// heaviest query tiles first (same mapping as attn_fwd_kernel)
__device__ __forceinline__ int abw_heavy_first(int mask, int n_frames, int hw, int nt, int i) {
  if ((mask == ATTN_DART || mask == ATTN_DART_LISTED) && !(nt & 1) && (n_frames * hw) % ABW_BM == 0) {
    const int th = nt >> 1;
    return (i & 1) * th + th - 1 - (i >> 1);
  }
  return mask == ATTN_FULL ? i : nt - 1 - i;
}

// ------------------------------------------------------------------------------------------------ dQ
// Warp roles: 0 = TMA producer, 1 = S / dP MMA issuer, 2 = dQ MMA issuer (each ONE elected thread running its whole loop),
// 3..18 = softmax warps.  dS goes from the softmax warps to the dQ MMA through tensor memory (TS mode: the A operand is
// read from TMEM), not shared memory -- with dS staged in shared memory the kernel was bound by shared-memory bandwidth.
// TMEM columns: S [0,128); dP [128,256) (single-buffered: the next step's MMAs start as soon as the softmax warps have
// drained them, early in their step); dQ [256,320); dS (bf16 pairs) 2 x 64 from 320.
__global__ void __launch_bounds__(ABW_THREADS, 1) attn_bwd_dq_kernel(const __grid_constant__ AttnBwdParams p) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = base, sdO = sQ + ABW_T128;
  const uint32_t sKV = sdO + ABW_T128;                       // stage s: K at +s*32K, V at +s*32K+16K
  const uint32_t bar = sKV + ABW_STAGES * 2 * ABW_T128;
  const uint32_t q_full = bar;
  auto kv_full = [&](int s) { return bar + 8u * (1 + s); };
  auto kv_empty = [&](int s) { return bar + 8u * (1 + ABW_STAGES + s); };
  const uint32_t sdp_full = bar + 8u * (1 + 2 * ABW_STAGES), sdp_empty = bar + 8u * (2 + 2 * ABW_STAGES);
  auto ds_full = [&](int b) { return bar + 8u * (3 + 2 * ABW_STAGES + b); };
  auto ds_empty = [&](int b) { return bar + 8u * (3 + ABW_DS_BUFS + 2 * ABW_STAGES + b); };
  const uint32_t acc_full = bar + 8u * (3 + 2 * ABW_DS_BUFS + 2 * ABW_STAGES);
  const uint32_t tmem_slot = bar + 8u * (4 + 2 * ABW_DS_BUFS + 2 * ABW_STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.x;
  const int bb = bh / p.heads, hh = bh - bb * p.heads;
  const int q0 = abw_heavy_first(p.mask, p.n_frames, p.hw, gridDim.y, blockIdx.y) * ABW_BM;
  const TileRanges kr = visible_keys(p, q0, min(q0 + ABW_BM, p.Lq));
  const int n_kv = kr.count();

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < ABW_STAGES; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
    mbar_init(sdp_full, 1); mbar_init(sdp_empty, ABW_SM_WARPS);
    for (int b = 0; b < ABW_DS_BUFS; ++b) { mbar_init(ds_full(b), ABW_SM_WARPS); mbar_init(ds_empty(b), 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot));
  const uint32_t tS = tmem, tdP = tmem + 128, tdQ = tmem + 256, tdS = tmem + 320;

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, 2 * ABW_T128);
      tma_load_4d(sQ, &p.mapQ128, q_full, 0, q0, hh, bb);
      tma_load_4d(sdO, &p.mapdO128, q_full, 0, q0, hh, bb);
      int st = 0;
      uint32_t ph = 1;
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(kv_empty(st), ph);
        const uint32_t sK = sKV + st * 2 * ABW_T128, sV = sK + ABW_T128;
        mbar_arrive_expect_tx(kv_full(st), 2 * ABW_T128);
        const int k0 = kr.tile(j) * ABW_BN;
        tma_load_4d(sK, &p.mapK128, kv_full(st), 0, k0, hh, bb);
        tma_load_4d(sV, &p.mapV128, kv_full(st), 0, k0, hh, bb);
        if (++st == ABW_STAGES) { st = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---- S = Q K^T and dP = dO V^T
    if (n_kv > 0 && elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, ABW_BN, 0, 0);
      const uint64_t kdesc0 = make_smem_desc(0, 16, 1024, SWZ_128B);
      const uint64_t qd = kdesc0 + (sQ >> 4), dod = kdesc0 + (sdO >> 4);
      int st = 0;
      uint32_t ph = 0;
      mbar_wait(q_full, 0);
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(kv_full(st), ph);
        if (j >= 1) mbar_wait(sdp_empty, (j & 1) ^ 1);      // the softmax warps have drained S(j-1), dP(j-1)
        tc_fence_after();
        const uint64_t kd = kdesc0 + ((sKV + st * 2 * ABW_T128) >> 4), vd = kd + (ABW_T128 >> 4);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ss(tS, qd + 2 * k, kd + 2 * k, idesc_s, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ss(tdP, dod + 2 * k, vd + 2 * k, idesc_s, k > 0);
        umma_commit(sdp_full);
        if (++st == ABW_STAGES) { st = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 2) {
    // ---- dQ += dS[128 x 128 keys] * K[128 keys x 64]: dS from tensor memory, K (MN-major) from shared memory
    if (n_kv > 0 && elect_one()) {
      constexpr uint32_t idesc_q = make_idesc_bf16(128, ATTN_D, 0, 1);
      const uint64_t mdesc0 = make_smem_desc(0, ABW_T128, 1024, SWZ_128B);
      int st = 0, db = 0;
      uint32_t ph = 0, dph = 0;
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(kv_full(st), ph);
        mbar_wait(ds_full(db), dph);
        tc_fence_after();
        const uint64_t kd = mdesc0 + ((sKV + st * 2 * ABW_T128) >> 4);
#pragma unroll
        for (int kk = 0; kk < ABW_BN / 16; ++kk)     // 16 keys of dS = 8 TMEM columns
          umma_bf16_ts(tdQ, tdS + db * 64 + kk * 8, kd + kk * (2048 >> 4), idesc_q, (j > 0) || (kk > 0));
        umma_commit(kv_empty(st));       // S(j), dP(j) retired before the softmax produced dS(j): K and V are free
        umma_commit(ds_empty(db));
        if (++st == ABW_STAGES) { st = 0; ph ^= 1; }
        if (++db == ABW_DS_BUFS) { db = 0; dph ^= 1; }
      }
      umma_commit(acc_full);
    }
    __syncwarp();
  } else {
    const int qw = warp & 3;
    const int part = (warp - 3) >> 2;      // which 32 of the 128 streamed key columns this warp owns
    const int r = qw * 32 + lane;
    const int iq = q0 + r;
    const uint32_t lane_off = static_cast<uint32_t>(qw * 32) << 16;
    const float c1 = p.scale * 1.4426950408889634f;
    const int Lp = (p.Lq + ABW_BN - 1) / ABW_BN * ABW_BN;
    float nlse2 = 0.f, ndsum = 0.f;
    if (iq < p.Lq) {
      ndsum = p.ws[static_cast<long>(bh) * Lp + iq];
      nlse2 = p.ws[(static_cast<long>(p.BH) + bh) * Lp + iq];
    }
    const uint64_t NL = pack2(nlse2, nlse2), ND = pack2(ndsum, ndsum);
    const Win2 win = key_windows(p.mask, p.n_frames, p.hw, iq, p.Lq, p.Lk);
    const uint32_t t_s = tS + lane_off + part * ABW_COLS, t_dp = tdP + lane_off + part * ABW_COLS;
    const uint32_t t_ds = tdS + lane_off + part * (ABW_COLS / 2);
    auto nl = [&](int) { return NL; };
    auto nd = [&](int) { return ND; };
    int db = 0;
    uint32_t dph = 0;
    for (int j = 0; j < n_kv; ++j) {
      const int ik0 = kr.tile(j) * ABW_BN + part * ABW_COLS;
      mbar_wait(sdp_full, j & 1);
      tc_fence_after();
      uint32_t pk_ds[16];
#pragma unroll
      for (int h = 0; h < 2; ++h) {          // two chunks of 16 columns: 32 live accumulator registers at a time
        float s[16], dp[16];
        tmem_ld16(t_s + h * 16, s);
        tmem_ld16(t_dp + h * 16, dp);
        tmem_ld_wait();
        if (h == 1) {                         // S and dP of this step are in registers: the next step's MMAs may overwrite them
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(sdp_empty);
        }
        uint32_t unused[8];
        uint32_t (&out)[8] = *reinterpret_cast<uint32_t (*)[8]>(&pk_ds[h * 8]);
        const int c0 = ik0 + h * 16;
        if (win.whole(c0, 16)) pds_chunk16<false, false>(s, dp, unused, out, c1, p.scale, nl, nd, 0, 0u, 0, 0u);
        else
          pds_chunk16<true, false>(s, dp, unused, out, c1, p.scale, nl, nd, win.a1 - c0, static_cast<unsigned>(win.b1 - win.a1),
                                   win.a2 - c0, static_cast<unsigned>(win.b2 - win.a2));
      }
      mbar_wait(ds_empty(db), dph ^ 1);    // first pass falls through
      tmem_st16(t_ds + db * 64, pk_ds);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(ds_full(db));
      if (++db == ABW_DS_BUFS) { db = 0; dph ^= 1; }
    }
    if (n_kv > 0) { mbar_wait(acc_full, 0); tc_fence_after(); }
    __nv_bfloat16* drow = p.dq + ((static_cast<long>(bb) * p.Lq + iq) * p.heads + hh) * ATTN_D + part * 16;
    {
      float o[16];
      if (n_kv > 0) { tmem_ld16(tdQ + lane_off + part * 16, o); tmem_ld_wait(); }
      else {
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] = 0.f;
      }
      if (iq < p.Lq) {
#pragma unroll
        for (int i = 0; i < 16; i += 8)
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
// Same warp roles; rows are keys, streamed columns are queries.  P^T and dS^T go to the dV / dK MMAs through tensor memory.
// The 128 streamed queries' statistics (-lse*log2e, -D*scale) ride along with their Q / dO tiles as bulk copies per stage.
// TMEM columns (all 512, everything single-buffered): S^T [0,128); dP^T [128,256); dK [256,320); dV [320,384); P^T (bf16
// pairs) [384,448); dS^T [448,512).
__global__ void __launch_bounds__(ABW_THREADS, 1) attn_bwd_dkv_kernel(const __grid_constant__ AttnBwdParams p) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sK = base, sV = sK + ABW_T128;
  const uint32_t sQdO = sV + ABW_T128;                       // stage s: Q at +s*32K, dO at +s*32K+16K
  const uint32_t sStat = sQdO + ABW_STAGES * 2 * ABW_T128;   // stage s: [-lse*log2e (128) | -D*scale (128)] floats
  const uint32_t bar = sStat + ABW_STAGES * ABW_STAT_BYTES;
  const uint32_t kv_full = bar;
  auto q_full = [&](int s) { return bar + 8u * (1 + s); };
  auto q_empty = [&](int s) { return bar + 8u * (1 + ABW_STAGES + s); };
  const uint32_t sdp_full = bar + 8u * (1 + 2 * ABW_STAGES), sdp_empty = bar + 8u * (2 + 2 * ABW_STAGES);
  const uint32_t pds_full = bar + 8u * (3 + 2 * ABW_STAGES), pds_empty = bar + 8u * (4 + 2 * ABW_STAGES);
  const uint32_t acc_full = bar + 8u * (5 + 2 * ABW_STAGES);
  const uint32_t tmem_slot = bar + 8u * (6 + 2 * ABW_STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.x;
  const int bb = bh / p.heads, hh = bh - bb * p.heads;
  const int k0 = blockIdx.y * ABW_BM;       // ascending key tiles already run heaviest first (clean keys, early frames)
  const TileRanges qr = visible_queries(p, k0, min(k0 + ABW_BM, p.Lk));
  const int n_q = qr.count();
  const int Lp = (p.Lq + ABW_BN - 1) / ABW_BN * ABW_BN;

  if (threadIdx.x == 0) {
    mbar_init(kv_full, 1);
    for (int s = 0; s < ABW_STAGES; ++s) { mbar_init(q_full(s), 1); mbar_init(q_empty(s), 1); }
    mbar_init(sdp_full, 1); mbar_init(sdp_empty, ABW_SM_WARPS);
    mbar_init(pds_full, ABW_SM_WARPS); mbar_init(pds_empty, 1);
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot));
  const uint32_t tS = tmem, tdP = tmem + 128, tdK = tmem + 256, tdV = tmem + 320, tP = tmem + 384, tdS = tmem + 448;

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(kv_full, 2 * ABW_T128);
      tma_load_4d(sK, &p.mapK128, kv_full, 0, k0, hh, bb);
      tma_load_4d(sV, &p.mapV128, kv_full, 0, k0, hh, bb);
      const float* nd_row = p.ws + static_cast<long>(bh) * Lp;
      const float* nl_row = p.ws + (static_cast<long>(p.BH) + bh) * Lp;
      int st = 0;
      uint32_t ph = 1;
      for (int j = 0; j < n_q; ++j) {
        mbar_wait(q_empty(st), ph);
        const uint32_t sQ = sQdO + st * 2 * ABW_T128, sdO = sQ + ABW_T128, sS = sStat + st * ABW_STAT_BYTES;
        mbar_arrive_expect_tx(q_full(st), 2 * ABW_T128 + ABW_STAT_BYTES);
        const int q0 = qr.tile(j) * ABW_BN;
        tma_load_4d(sQ, &p.mapQ128, q_full(st), 0, q0, hh, bb);
        tma_load_4d(sdO, &p.mapdO128, q_full(st), 0, q0, hh, bb);
        bulk_load_1d(sS, nl_row + q0, ABW_BN * 4, q_full(st));
        bulk_load_1d(sS + ABW_BN * 4, nd_row + q0, ABW_BN * 4, q_full(st));
        if (++st == ABW_STAGES) { st = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---- S^T = K Q^T and dP^T = V dO^T
    if (n_q > 0 && elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, ABW_BN, 0, 0);
      const uint64_t kdesc0 = make_smem_desc(0, 16, 1024, SWZ_128B);
      const uint64_t kd = kdesc0 + (sK >> 4), vd = kdesc0 + (sV >> 4);
      int st = 0;
      uint32_t ph = 0;
      mbar_wait(kv_full, 0);
      for (int j = 0; j < n_q; ++j) {
        mbar_wait(q_full(st), ph);
        if (j >= 1) mbar_wait(sdp_empty, (j & 1) ^ 1);
        tc_fence_after();
        const uint64_t qd = kdesc0 + ((sQdO + st * 2 * ABW_T128) >> 4), dod = qd + (ABW_T128 >> 4);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ss(tS, kd + 2 * k, qd + 2 * k, idesc_s, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ss(tdP, vd + 2 * k, dod + 2 * k, idesc_s, k > 0);
        umma_commit(sdp_full);
        if (++st == ABW_STAGES) { st = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 2) {
    // ---- dV += P^T[128 keys x 128 q] * dO[128 q x 64];  dK += dS^T * Q   (A operands from tensor memory)
    if (n_q > 0 && elect_one()) {
      constexpr uint32_t idesc_a = make_idesc_bf16(128, ATTN_D, 0, 1);
      const uint64_t mdesc0 = make_smem_desc(0, ABW_T128, 1024, SWZ_128B);
      int st = 0;
      uint32_t ph = 0;
      for (int j = 0; j < n_q; ++j) {
        mbar_wait(q_full(st), ph);
        mbar_wait(pds_full, j & 1);
        tc_fence_after();
        const uint64_t qd = mdesc0 + ((sQdO + st * 2 * ABW_T128) >> 4), dod = qd + (ABW_T128 >> 4);
#pragma unroll
        for (int kk = 0; kk < ABW_BN / 16; ++kk)
          umma_bf16_ts(tdV, tP + kk * 8, dod + kk * (2048 >> 4), idesc_a, (j > 0) || (kk > 0));
#pragma unroll
        for (int kk = 0; kk < ABW_BN / 16; ++kk)
          umma_bf16_ts(tdK, tdS + kk * 8, qd + kk * (2048 >> 4), idesc_a, (j > 0) || (kk > 0));
        umma_commit(q_empty(st));
        umma_commit(pds_empty);
        if (++st == ABW_STAGES) { st = 0; ph ^= 1; }
      }
      umma_commit(acc_full);
    }
    __syncwarp();
  } else {
    const int qw = warp & 3;
    const int part = (warp - 3) >> 2;     // which 32 of the 128 streamed query columns this warp owns
    const int r = qw * 32 + lane;         // key row of the tile
    const int ik = k0 + r;
    const uint32_t lane_off = static_cast<uint32_t>(qw * 32) << 16;
    const float c1 = p.scale * 1.4426950408889634f;
    const Win2 win = query_windows(p.mask, p.n_frames, p.hw, ik, p.Lq, p.Lk);
    const uint32_t t_s = tS + lane_off + part * ABW_COLS, t_dp = tdP + lane_off + part * ABW_COLS;
    const uint32_t t_p = tP + lane_off + part * (ABW_COLS / 2), t_ds = tdS + lane_off + part * (ABW_COLS / 2);
    int st = 0;
    uint32_t ph = 0;
    for (int j = 0; j < n_q; ++j) {
      const int iq0 = qr.tile(j) * ABW_BN + part * ABW_COLS;
      mbar_wait(q_full(st), ph);           // this stage's statistics are visible (the S^T MMA needed the same barrier)
      mbar_wait(sdp_full, j & 1);
      tc_fence_after();
      uint32_t pk_p[16], pk_ds[16];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float s[16], dp[16];
        tmem_ld16(t_s + h * 16, s);
        tmem_ld16(t_dp + h * 16, dp);
        tmem_ld_wait();
        if (h == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(sdp_empty);
        }
        // this chunk's column statistics: 16 + 16 floats, the same for every lane (broadcast reads)
        uint64_t nlv[8], ndv[8];
        {
          const uint32_t a_l = sStat + st * ABW_STAT_BYTES + (part * ABW_COLS + h * 16) * 4, a_d = a_l + ABW_BN * 4;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(nlv[2 * i]), "=l"(nlv[2 * i + 1]) : "r"(a_l + i * 16));
            asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(ndv[2 * i]), "=l"(ndv[2 * i + 1]) : "r"(a_d + i * 16));
          }
        }
        auto nl = [&](int pi) { return nlv[pi]; };
        auto nd = [&](int pi) { return ndv[pi]; };
        uint32_t (&op)[8] = *reinterpret_cast<uint32_t (*)[8]>(&pk_p[h * 8]);
        uint32_t (&od)[8] = *reinterpret_cast<uint32_t (*)[8]>(&pk_ds[h * 8]);
        const int c0 = iq0 + h * 16;
        if (win.whole(c0, 16)) pds_chunk16<false, true>(s, dp, op, od, c1, p.scale, nl, nd, 0, 0u, 0, 0u);
        else
          pds_chunk16<true, true>(s, dp, op, od, c1, p.scale, nl, nd, win.a1 - c0, static_cast<unsigned>(win.b1 - win.a1),
                                  win.a2 - c0, static_cast<unsigned>(win.b2 - win.a2));
      }
      if (j >= 1) mbar_wait(pds_empty, (j & 1) ^ 1);     // the dV / dK MMAs of the previous step have consumed P^T, dS^T
      tmem_st16(t_p, pk_p);
      tmem_st16(t_ds, pk_ds);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_full);
      if (++st == ABW_STAGES) { st = 0; ph ^= 1; }
    }
    if (n_q > 0) { mbar_wait(acc_full, 0); tc_fence_after(); }
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      __nv_bfloat16* drow = (which == 0 ? p.dk : p.dv) + ((static_cast<long>(bb) * p.Lk + ik) * p.heads + hh) * ATTN_D + part * 16;
      float o[16];
      if (n_q > 0) { tmem_ld16((which == 0 ? tdK : tdV) + lane_off + part * 16, o); tmem_ld_wait(); }
      else {
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] = 0.f;
      }
      if (ik < p.Lk) {
#pragma unroll
        for (int i = 0; i < 16; i += 8)
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
