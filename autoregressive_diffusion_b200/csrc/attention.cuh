// Frame-structured attention on tcgen05/TMEM (reference: edm2/attention/attention_modules.py:59-77 --
// compiled FlexAttention with make_train_mask / make_infer_mask, and F.scaled_dot_product_attention).
//
// q, k, v: bf16 [B, L, heads, 64] (token-major rows of heads*64 channels -- the NHWC activation layout), q/k already RMS-normalised and rotary-embedded, so every logit
// q.k/8 lies in [-8, 8] (|q|,|k| <= 8 and the xPos factor of an allowed pair is <= 1).  That bound replaces
// the running max of online softmax: p = exp(s - 8) can neither overflow nor vanish, the O accumulator in
// TMEM is never rescaled, and the row statistic saved for backward is lse = 8 + log(sum p).
//
// Masks are functions of FRAME indices (token / hw), evaluated in registers; only KV tiles that contain an
// allowed frame are visited, so the DART training mask costs n(n+1) frame pairs instead of (2n)^2.
#pragma once
#include <cuda.h>
#include "launch.cuh"
#include <cuda_runtime.h>
#include <cstdint>
#include "ptx.cuh"

namespace ob {

enum : int { ATTN_FULL = 0, ATTN_CAUSAL = 1, ATTN_DART = 2, ATTN_DART_LISTED = 3 };

constexpr int ATTN_BM = 128;  // query rows per CTA
constexpr int ATTN_BN = 128;  // key rows per step
constexpr int ATTN_D = 64;    // head dim
constexpr float ATTN_SMAX = 8.0f;

struct AttnParams {
  CUtensorMap mapQ, mapK, mapV;  // 4D: (64, L, heads, B), box (64, 128, 1, 1)
  int BH, heads, Lq, Lk;
  int hw;         // tokens per frame
  int n_frames;   // DART: frames per half (clean / noised)
  int mask;
  float scale;    // 1/sqrt(64)
  __nv_bfloat16* o;  // [B, Lq, heads, 64]
  float* lse;        // [BH, Lq]
  // ---- paged decode (attn_fwd_kernel<true>): keys / values live in frame-sized pages of a pool
  //      [n_pages, hw, heads, 64]; mapK / mapV are 4D maps (64, hw, heads, n_pages) with box (64, box_rows, 1, 1)
  const int* page_table;   // [B, max_pages] page of frame t of sequence b
  const int* lengths;      // [B] committed frames per sequence (device side: one CUDA graph serves every decode step)
  int max_pages, n_pages, box_rows;
  int extra_frames;        // frames visible beyond lengths[b] (1: the frame being generated sits in slot lengths[b])
  int n_split;             // split-KV factor (gridDim.z); > 1: partial sums go to o_part / l_part
  float* o_part;           // [n_split, B, Lq, heads, 64] un-normalised partial outputs
  float* l_part;           // [n_split, BH, Lq] partial row sums
};

// Is key frame kf visible from query frame qf?
__device__ __forceinline__ bool frame_visible(int mask, int n, int qf, int kf) {
  if (mask == ATTN_FULL) return true;
  if (mask == ATTN_CAUSAL) return kf <= qf;
  // DART (attention_masking.py:15-24): clean->clean causal, noised->strictly earlier clean, noised->itself
  if (qf < n) return kf <= qf;
  return (kf < qf - n) || (kf == qf);
}

// ATTN_DART_LISTED: what COMPILED FlexAttention computes for make_train_mask when a frame has fewer than 128 tokens
// (attention_masking.py:32-53; measured on the B200, profiles/r02_ref_gpu_baseline.json): the frames are regrouped into
// 128-token blocks, the FRAME-level block list is reused for them, and the kernel only visits listed blocks, so the
// effective mask is mask_mod AND listed(block(q), block(k)).  A noised query in block i of its half is listed the clean
// blocks < i and its own block; clean queries lose nothing (frame-causal implies block-causal).  n_hw = tokens per half.
__device__ __forceinline__ bool block_listed(int n_hw, int iq, int ik) {
  if (iq < n_hw) return true;
  return ik < n_hw ? (ik >> 7) < ((iq - n_hw) >> 7) : (ik >> 7) == (iq >> 7);
}

// KV tiles a query tile must visit: [0, n1) and [s2, e2) (tile indices, second range may be empty).
struct KvRange {
  int n1, s2, e2;
  __device__ __forceinline__ int count() const { return n1 + (e2 > s2 ? e2 - s2 : 0); }
  __device__ __forceinline__ int tile(int j) const { return j < n1 ? j : s2 + (j - n1); }
};

__device__ __forceinline__ KvRange kv_range(const AttnParams& p, int q0, int Lk) {
  KvRange r;
  const int kv_tiles = (Lk + ATTN_BN - 1) / ATTN_BN;
  const int q_last = min(q0 + ATTN_BM, p.Lq) - 1;
  r.s2 = r.e2 = 0;
  if (p.mask == ATTN_FULL) {
    r.n1 = kv_tiles;
  } else if (p.mask == ATTN_CAUSAL) {
    const int qf_hi = q_last / p.hw;
    r.n1 = min(kv_tiles, ((qf_hi + 1) * p.hw + ATTN_BN - 1) / ATTN_BN);
  } else {
    const int n = p.n_frames;
    const int qf_lo = q0 / p.hw, qf_hi = q_last / p.hw;
    int end1;  // clean keys [0, end1)
    if (qf_hi < n) end1 = (qf_hi + 1) * p.hw;                 // all rows clean
    else {
      end1 = (qf_hi - n) * p.hw;                               // noised rows: clean frames < qf-n
      if (qf_lo < n) end1 = max(end1, n * p.hw);               // tile straddles the halves: clean rows see up to frame n-1
    }
    r.n1 = min(kv_tiles, (end1 + ATTN_BN - 1) / ATTN_BN);
    if (qf_hi >= n) {                                          // noised rows also see their own frame
      const int lo = max(qf_lo, n) * p.hw, hi = (qf_hi + 1) * p.hw;
      r.s2 = max(r.n1, lo / ATTN_BN);
      r.e2 = min(kv_tiles, (hi + ATTN_BN - 1) / ATTN_BN);
      if (r.e2 < r.s2) r.e2 = r.s2;
    }
  }
  return r;
}

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// exp2 on the FMA pipe: round-to-nearest split x = i + f with the 1.5*2^23 trick, a degree-3 minimax polynomial for 2^f on
// [-0.5, 0.5] (max relative error 7.5e-5, far below the bf16 rounding of P), and 2^i added into the exponent field.
// ncu (profiles/r02_ncu_attention_summary.txt + source page): 31 % of the forward kernel's stall samples sit on MUFU.EX2
// -- the softmax warps are bound by the 16/clk/SM transcendental unit, not by the tensor pipe (24 % active) -- so every
// third exponential is computed here instead (9 FMA-pipe operations at 128/clk/SM), which balances the two pipes.
// Valid for x in [-120, 120]: the callers' arguments lie in [-45, 0] (|logit| <= 8, fixed maximum / saved lse).
__device__ __forceinline__ float poly_exp2(float x) {
  const float t = x + 12582912.f;                  // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const float f = x - (t - 12582912.f);            // in [-0.5, 0.5]
  float p = fmaf(0.0551706217f, f, 0.242608815f);
  p = fmaf(p, f, 0.693260968f);
  p = fmaf(p, f, 0.999928236f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
// which elements of an unrolled chunk take the polynomial
#define ATTN_EXP_POLY(idx) (((idx) % 3) == 2)

#ifndef ATTN_KV_STAGES_N
#define ATTN_KV_STAGES_N 6
#endif
constexpr int ATTN_KV_STAGES = ATTN_KV_STAGES_N;
constexpr int ATTN_TILE_BYTES = 128 * 128;  // [128 rows][64 bf16]
constexpr int ATTN_SMEM_BYTES = 1024 + ATTN_TILE_BYTES /*Q*/ + ATTN_KV_STAGES * 2 * ATTN_TILE_BYTES /*K,V*/ +
                                2048 /*row sums*/ + 256;
// four softmax warps per TMEM lane quarter (= per scheduler): with two the warps were latency-bound (MUFU + tcgen05.ld
// round trips; 526 -> 628 TFLOP/s at 131k tokens going to four)
#ifndef ATTN_DIAG
#define ATTN_DIAG 0
#endif
#ifndef ATTN_HEAVY_FIRST
#define ATTN_HEAVY_FIRST 1
#endif
#ifndef ATTN_P_BUFS_N
#define ATTN_P_BUFS_N 3
#endif
constexpr int ATTN_P_BUFS = ATTN_P_BUFS_N;
constexpr int ATTN_SOFTMAX_WARPS = 16;
constexpr int ATTN_PARTS = ATTN_SOFTMAX_WARPS / 4;
constexpr int ATTN_THREADS = 96 + 32 * ATTN_SOFTMAX_WARPS;   // TMA, S-MMA and PV-MMA warps + softmax warps

// ---- packed fp32 pairs: one FFMA2 / FADD2 issue slot does two lanes' worth of work per thread (sm_100 fma.rn.f32x2)
__device__ __forceinline__ uint64_t pack2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// poly_exp2 on a pair
__device__ __forceinline__ void poly_exp2_pair(uint64_t x, float& e0, float& e1) {
  const uint64_t magic = pack2(12582912.f, 12582912.f);
  const uint64_t t = add2(x, magic);
  const uint64_t f = sub2(x, sub2(t, magic));
  uint64_t q = fma2(pack2(0.0551706217f, 0.0551706217f), f, pack2(0.242608815f, 0.242608815f));
  q = fma2(q, f, pack2(0.693260968f, 0.693260968f));
  q = fma2(q, f, pack2(0.999928236f, 0.999928236f));
  float q0, q1, t0, t1;
  unpack2(q, q0, q1);
  unpack2(t, t0, t1);
  e0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
  e1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
}
// which PAIRS of an unrolled 32-element chunk take the polynomial
// (3 of 16: once the kernel stopped being MUFU-bound -- P in tensor memory, packed math -- fewer polynomial pairs won:
//  0 / 2 / 3 / 5 of 16 measured 873 / 918 / 923 / 882 TFLOP/s at 65 536 tokens)
#ifndef ATTN_PAIR_POLY
#define ATTN_PAIR_POLY(pi) (((pi) % 5) == 4)
#endif

// The keys a query row sees, as two windows of token indices: [0, w1_end) and [w2_lo, w2_hi).  Every mask of this file
// has that shape per row -- causal / clean DART rows: one window up to the end of the row's frame; noised DART rows: the
// clean frames strictly before their own, plus their own frame; ATTN_DART_LISTED clips both to the listed 128-token
// blocks (block_listed) -- so tile and element tests are integer compares, no per-tile division.
struct RowWindows {
  int w1_end, w2_lo, w2_hi;
  __device__ __forceinline__ bool whole_tile(int k0) const {
    return (k0 + ATTN_BN <= w1_end) || (k0 >= w2_lo && k0 + ATTN_BN <= w2_hi);
  }
};
__device__ __forceinline__ RowWindows row_windows(int mask, int n_frames, int hw, int iq, int Lk) {
  RowWindows w;
  w.w2_lo = w.w2_hi = 0;
  const int qf = iq / hw;
  if (mask == ATTN_FULL) w.w1_end = Lk;
  else if (mask == ATTN_CAUSAL || qf < n_frames) w.w1_end = (qf + 1) * hw;
  else {
    w.w1_end = (qf - n_frames) * hw;
    w.w2_lo = qf * hw;
    w.w2_hi = w.w2_lo + hw;
    if (mask == ATTN_DART_LISTED) {
      const int n_hw = n_frames * hw, blk = (iq >> 7) << 7;
      w.w1_end = min(w.w1_end, ((iq - n_hw) >> 7) << 7);
      w.w2_lo = max(w.w2_lo, blk);
      w.w2_hi = min(w.w2_hi, blk + 128);
    }
  }
  w.w1_end = min(w.w1_end, Lk);
  w.w2_hi = min(w.w2_hi, Lk);
  return w;
}

// p = exp2(s*c1 - c2) for 32 logits, rounded to bf16 pairs; returns the (fp32) sum.  MASKED: keys are tested one by
// one against the row's windows, given relative to the chunk's first key (only for tiles that straddle a mask edge).
template <bool MASKED>
__device__ __forceinline__ float softmax_chunk(const float (&s)[32], uint32_t (&packed)[16], float c1, float c2, int rel_end,
                                               int rel_lo, int rel_hi) {
  const uint64_t C1 = pack2(c1, c1), C2 = pack2(-c2, -c2);
  uint64_t acc = pack2(0.f, 0.f);
  const unsigned w2_len = static_cast<unsigned>(rel_hi - rel_lo);
#pragma unroll
  for (int pi = 0; pi < 16; ++pi) {
    const uint64_t arg = fma2(pack2(s[2 * pi], s[2 * pi + 1]), C1, C2);
    float e0, e1;
#if ATTN_DIAG == 1
    if (true) unpack2(arg, e0, e1);
    else
#endif
    if (ATTN_PAIR_POLY(pi)) poly_exp2_pair(arg, e0, e1);
    else {
      float a0, a1;
      unpack2(arg, a0, a1);
      e0 = fast_exp2(a0);
      e1 = fast_exp2(a1);
    }
    if (MASKED) {
      const int i0 = 2 * pi, i1 = 2 * pi + 1;
      e0 = (i0 < rel_end || static_cast<unsigned>(i0 - rel_lo) < w2_len) ? e0 : 0.f;
      e1 = (i1 < rel_end || static_cast<unsigned>(i1 - rel_lo) < w2_len) ? e1 : 0.f;
    }
    acc = add2(acc, pack2(e0, e1));
    packed[pi] = pack_bf16x2(e0, e1);
  }
  float x, y;
  unpack2(acc, x, y);
  return x + y;
}

// PAGED = false: q, k, v are contiguous [B, L, heads, 64] tensors (training, prefill, per-frame attention).
// PAGED = true : the decode path of the sampler (attention_modules.py:69-70 after the cat with the cache, :51-57): one
//   frame of queries per sequence against the frames cached in pages + the frame being generated, unmasked.  The key
//   length comes from device memory, every 128-key tile is gathered page by page through the page table, and
//   blockIdx.z selects a slice of the key tiles (split-KV).  The softmax uses the fixed maximum ATTN_SMAX, so the
//   slices' un-normalised outputs and row sums simply ADD (attn_decode_combine_kernel) -- no running-max bookkeeping.
template <bool PAGED>
__global__ void __launch_bounds__(ATTN_THREADS, 1) attn_fwd_kernel(const __grid_constant__ AttnParams p) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = base;
  const uint32_t sKV = sQ + ATTN_TILE_BYTES;                       // stage s: K at +s*32K, V at +s*32K+16K
  const uint32_t sL = sKV + ATTN_KV_STAGES * 2 * ATTN_TILE_BYTES;  // [ATTN_PARTS][128 rows] partial row sums
  const uint32_t bar = sL + 2048;
  const uint32_t q_full = bar;
  auto kv_full = [&](int s) { return bar + 8u * (1 + s); };
  auto kv_empty = [&](int s) { return bar + 8u * (1 + ATTN_KV_STAGES + s); };
  auto s_full = [&](int b) { return bar + 8u * (1 + 2 * ATTN_KV_STAGES + b); };
  auto s_empty = [&](int b) { return bar + 8u * (3 + 2 * ATTN_KV_STAGES + b); };
  auto p_full = [&](int b) { return bar + 8u * (5 + 2 * ATTN_KV_STAGES + b); };
  auto p_empty = [&](int b) { return bar + 8u * (8 + 2 * ATTN_KV_STAGES + b); };
  const uint32_t o_full = bar + 8u * (11 + 2 * ATTN_KV_STAGES);
  const uint32_t tmem_slot = bar + 8u * (12 + 2 * ATTN_KV_STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Non-paged launches put (batch, head) on grid.x and the query tiles on grid.y in order of DECREASING work: under the
  // causal / DART masks a tile's key range grows with its position in its half, and CTAs are dispatched in grid order, so
  // the long tiles start first and the short ones fill the tail (ncu: SMs idle 9 % of the launch in ascending order).
  int bh, qt;
  if constexpr (PAGED) { bh = blockIdx.y; qt = blockIdx.x; }
  else {
    bh = blockIdx.x;
    const int nt = gridDim.y, i = blockIdx.y;
#if ATTN_HEAVY_FIRST
    if ((p.mask == ATTN_DART || p.mask == ATTN_DART_LISTED) && !(nt & 1) && (p.n_frames * p.hw) % ATTN_BM == 0) {
      const int th = nt >> 1;
      qt = (i & 1) * th + th - 1 - (i >> 1);           // clean and noised halves interleaved, each descending
    } else qt = p.mask == ATTN_FULL ? i : nt - 1 - i;
#else
    qt = i;
#endif
  }
  const int bb = bh / p.heads, hh = bh - bb * p.heads;
  const int q0 = qt * ATTN_BM;
  int Lk = p.Lk;
  KvRange kr;
  if constexpr (PAGED) {
    Lk = (p.lengths[bb] + p.extra_frames) * p.hw;
    const int kv_tiles = (Lk + ATTN_BN - 1) / ATTN_BN;
    const int per = (kv_tiles + p.n_split - 1) / p.n_split;
    const int lo = min(kv_tiles, static_cast<int>(blockIdx.z) * per), hi = min(kv_tiles, lo + per);
    kr.n1 = 0; kr.s2 = lo; kr.e2 = hi;        // this slice's tiles: [lo, hi)
  } else {
    kr = kv_range(p, q0, Lk);
  }
  const int n_kv = kr.count();

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < ATTN_KV_STAGES; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(s_full(b), 1); mbar_init(s_empty(b), ATTN_SOFTMAX_WARPS);
    }
    for (int b = 0; b < ATTN_P_BUFS; ++b) { mbar_init(p_full(b), ATTN_SOFTMAX_WARPS); mbar_init(p_empty(b), 1); }
    mbar_init(o_full, 1);
    fence_barrier_init();
    tma_prefetch_desc(&p.mapQ); tma_prefetch_desc(&p.mapK); tma_prefetch_desc(&p.mapV);
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot));
  // S buffers at cols [0,128) and [128,256); O at [256,320); three P buffers (bf16 pairs, the A operand of the PV MMA)
  // at [320,384), [384,448), [448,512).  P never touches shared memory: with P staged there the kernel was bound by shared-memory
  // bandwidth (Q+K operand reads 32 KB, P+V reads 48 KB, P stores 32 KB, TMA fills 32 KB per tile at 128 B/clk ~ 1150 clk
  // against 512 clk of tensor time; removing the exponentials altogether only reached 896 TFLOP/s).
  const uint32_t tS0 = tmem, tO = tmem + 256, tP0 = tmem + 320;

  if (warp == 0) {
    // ---- TMA producer: ONE elected thread runs the whole loop (no per-tile elect / reconvergence)
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, ATTN_TILE_BYTES);
      tma_load_4d(sQ, &p.mapQ, q_full, 0, q0, hh, bb);
      int st = 0;
      uint32_t ph = 1;                                   // kv_empty parity to wait for (first pass falls through)
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(kv_empty(st), ph);
        const uint32_t sK = sKV + st * 2 * ATTN_TILE_BYTES, sV = sK + ATTN_TILE_BYTES;
        mbar_arrive_expect_tx(kv_full(st), 2 * ATTN_TILE_BYTES);
        const int k0 = kr.tile(j) * ATTN_BN;
        if constexpr (PAGED) {
          // gather the tile page by page; frames past the sequence's length read page n_pages, which is out of bounds for
          // the tensor map: TMA zero-fills it (and still delivers the bytes the barrier expects)
          const int* table = p.page_table + static_cast<long>(bb) * p.max_pages;
          for (int r0 = 0; r0 < ATTN_BN; r0 += p.box_rows) {
            const int tok = k0 + r0, fr = tok / p.hw, within = tok - fr * p.hw;
            const int page = (tok < Lk && fr < p.max_pages) ? table[fr] : p.n_pages;
            tma_load_4d(sK + r0 * 128, &p.mapK, kv_full(st), 0, within, hh, page);
            tma_load_4d(sV + r0 * 128, &p.mapV, kv_full(st), 0, within, hh, page);
          }
        } else {
          tma_load_4d(sK, &p.mapK, kv_full(st), 0, k0, hh, bb);
          tma_load_4d(sV, &p.mapV, kv_full(st), 0, k0, hh, bb);
        }
        if (++st == ATTN_KV_STAGES) { st = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---- S = Q K^T issuer (one elected thread).  S runs up to two tiles ahead of the softmax: S(j) only needs its K tile
    //      and the S buffer (j & 1) drained by the softmax of tile j-2.  A separate warp issues the PV MMAs: with one
    //      warp doing both, that warp was busy 70 % of the time (ncu source page: ~130 dependent instructions per tile)
    //      and its latency sat in the S -> softmax -> PV round trip.
    if (n_kv > 0 && elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, ATTN_BN, 0, 0);
      const uint64_t kdesc0 = make_smem_desc(0, 16, 1024, SWZ_128B);               // K-major operands (Q, K)
      const uint64_t qd = kdesc0 + (sQ >> 4);
      uint64_t kd = kdesc0 + (sKV >> 4);
      int st = 0;
      uint32_t ph = 0;
      mbar_wait(q_full, 0);
      for (int j = 0; j < n_kv; ++j) {
        const int b = j & 1;
        mbar_wait(kv_full(st), ph);
        if (j >= 2) mbar_wait(s_empty(b), ((j >> 1) & 1) ^ 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < ATTN_D / 16; ++k) umma_bf16_ss(tS0 + b * ATTN_BN, qd + 2 * k, kd + 2 * k, idesc_s, k > 0);
        umma_commit(s_full(b));
        kd += (2 * ATTN_TILE_BYTES) >> 4;
        if (++st == ATTN_KV_STAGES) { st = 0; ph ^= 1; kd = kdesc0 + (sKV >> 4); }
      }
    }
    __syncwarp();
  } else if (warp == 2) {
    // ---- O += P V issuer (one elected thread): P from tensor memory (TS mode), V from shared memory
    if (n_kv > 0 && elect_one()) {
      constexpr uint32_t idesc_o = make_idesc_bf16(128, ATTN_D, 0, 1);
      const uint64_t vdesc0 = make_smem_desc(0, ATTN_TILE_BYTES, 1024, SWZ_128B);  // MN-major operand (V)
      uint64_t vd = vdesc0 + ((sKV + ATTN_TILE_BYTES) >> 4);
      int st = 0, pb = 0;
      uint32_t ph = 0, pph = 0;
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(kv_full(st), ph);                      // V(j) landed (already true: S(j) needed the same barrier)
        mbar_wait(p_full(pb), pph);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < ATTN_BN / 16; ++kk)       // 16 keys of P = 8 TMEM columns
          umma_bf16_ts(tO, tP0 + pb * 64 + kk * 8, vd + kk * (2048 >> 4), idesc_o, (j > 0) || (kk > 0));
        umma_commit(kv_empty(st));                       // K(j) and V(j) are both consumed once this PV retires
        umma_commit(p_empty(pb));
        vd += (2 * ATTN_TILE_BYTES) >> 4;
        if (++st == ATTN_KV_STAGES) { st = 0; ph ^= 1; vd = vdesc0 + ((sKV + ATTN_TILE_BYTES) >> 4); }
        if (++pb == ATTN_P_BUFS) { pb = 0; pph ^= 1; }
      }
      umma_commit(o_full);
    }
    __syncwarp();
  } else {
    // ---- softmax: warps (q, q+4, q+8, q+12) share TMEM lane quarter q, each owns 32 key columns of the tile
    //      (measured against two groups of eight warps ping-ponging over alternate tiles: 718 vs 672 TFLOP/s at 131k tokens)
    const int sw = warp - 3;
    const int qw = warp & 3;               // the TMEM lane quarter a warp may access is fixed by its index
    const int part = sw >> 2;
    const int r = qw * 32 + lane;          // row of the tile == TMEM lane
    const int iq = q0 + r;
    const RowWindows win = row_windows(p.mask, p.n_frames, p.hw, iq, Lk);
    const uint32_t lane_off = static_cast<uint32_t>(qw * 32) << 16;
    const float c1 = p.scale * 1.4426950408889634f, c2 = ATTN_SMAX * 1.4426950408889634f;
    const uint32_t t_p = tP0 + lane_off + part * 16;   // this warp's 32 keys of P: 16 columns of bf16 pairs
    const uint32_t t_col = tS0 + lane_off + part * 32;
    float l = 0.f;
    int pb = 0;
    uint32_t pph = 0;
    for (int j = 0; j < n_kv; ++j) {
      const int b = j & 1;
      const int k0 = kr.tile(j) * ATTN_BN;
      const bool all_vis = win.whole_tile(k0);     // does this row see EVERY key of the tile?
      mbar_wait(s_full(b), (j >> 1) & 1);
      tc_fence_after();
      float s0[32];
#if ATTN_DIAG == 2
#pragma unroll
      for (int i = 0; i < 32; ++i) s0[i] = __int_as_float(j + i);
#else
      tmem_ld32(t_col + b * ATTN_BN, s0);
      tmem_ld_wait();
#endif
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_empty(b));          // S buffer may be overwritten by the MMA of tile j+2
      uint32_t pk0[16];
      if (all_vis) l += softmax_chunk<false>(s0, pk0, c1, c2, 0, 0, 0);
      else {
        const int ik0 = k0 + part * 32;
        l += softmax_chunk<true>(s0, pk0, c1, c2, win.w1_end - ik0, win.w2_lo - ik0, win.w2_hi - ik0);
      }
      mbar_wait(p_empty(pb), pph ^ 1);                 // the buffer's previous P consumed by its PV MMA (first pass falls through)
      tmem_st16(t_p + pb * 64, pk0);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full(pb));
      if (++pb == ATTN_P_BUFS) { pb = 0; pph ^= 1; }
    }
    // ---- combine the parts' row sums, then O / l; parts 0 and 1 each write 32 of the 64 output channels
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(sL + (part * 128 + r) * 4), "f"(l) : "memory");
    asm volatile("bar.sync 1, %0;" ::"n"(32 * ATTN_SOFTMAX_WARPS) : "memory");
    l = 0.f;
#pragma unroll
    for (int pp = 0; pp < ATTN_PARTS; ++pp) {            // same order in every part: identical l in all of them
      float lp;
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(lp) : "r"(sL + (pp * 128 + r) * 4));
      l += lp;
    }
    if (part < 2) {
    const int half = part;
    if (n_kv > 0) {
      mbar_wait(o_full, 0);
      tc_fence_after();
    }
    const float inv_l = l > 0.f ? 1.f / l : 0.f;
    __nv_bfloat16* orow = p.o + ((static_cast<long>(bb) * p.Lq + iq) * p.heads + hh) * ATTN_D + half * 32;
    float o[32];
    if (n_kv > 0) { tmem_ld32(tO + lane_off + half * 32, o); tmem_ld_wait(); }
    else {
#pragma unroll
      for (int i = 0; i < 32; ++i) o[i] = 0.f;
    }
    if (PAGED && p.n_split > 1) {
      // split-KV: un-normalised partial output and row sum of this key slice (an empty slice contributes zeros)
      if (iq < p.Lq) {
        float* prow = p.o_part + (((static_cast<long>(blockIdx.z) * (p.BH / p.heads) + bb) * p.Lq + iq) * p.heads + hh) * ATTN_D + half * 32;
#pragma unroll
        for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(prow + i) = make_float4(o[i], o[i + 1], o[i + 2], o[i + 3]);
        if (half == 0) p.l_part[(static_cast<long>(blockIdx.z) * p.BH + bh) * p.Lq + iq] = l;
      }
    } else
    if (iq < p.Lq) {
#pragma unroll
      for (int i = 0; i < 32; i += 8)
        *reinterpret_cast<uint4*>(orow + i) =
            make_uint4(pack_bf16x2(o[i] * inv_l, o[i + 1] * inv_l), pack_bf16x2(o[i + 2] * inv_l, o[i + 3] * inv_l),
                       pack_bf16x2(o[i + 4] * inv_l, o[i + 5] * inv_l), pack_bf16x2(o[i + 6] * inv_l, o[i + 7] * inv_l));
      if (half == 0 && p.lse != nullptr) p.lse[static_cast<long>(bh) * p.Lq + iq] = ATTN_SMAX + __logf(fmaxf(l, 1e-37f));
    }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// Sum of the split-KV slices: o = (sum_s o_part[s]) / (sum_s l_part[s]); one thread per 4 output channels.
__global__ void __launch_bounds__(256) attn_decode_combine_kernel(const float* __restrict__ o_part, const float* __restrict__ l_part,
                                                                  __nv_bfloat16* __restrict__ o, int n_split, int B, int Lq,
                                                                  int heads) {
  pdl_launch_dependents();
  pdl_wait();
  const long vec = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long total = static_cast<long>(B) * Lq * heads * (ATTN_D / 4);
  if (vec >= total) return;
  const long row = vec / (ATTN_D / 4);                 // (b, iq, head)
  const int hh = static_cast<int>(row % heads);
  const long bq = row / heads;
  const int iq = static_cast<int>(bq % Lq), bb = static_cast<int>(bq / Lq);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  float l = 0.f;
  for (int s = 0; s < n_split; ++s) {
    const float4 v = *reinterpret_cast<const float4*>(o_part + (static_cast<long>(s) * B * Lq * heads * (ATTN_D / 4) + vec) * 4);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    l += l_part[(static_cast<long>(s) * B * heads + bb * heads + hh) * Lq + iq];
  }
  const float inv = l > 0.f ? 1.f / l : 0.f;
  *reinterpret_cast<uint2*>(o + vec * 4) = make_uint2(pack_bf16x2(acc.x * inv, acc.y * inv), pack_bf16x2(acc.z * inv, acc.w * inv));
}

}  // namespace ob
