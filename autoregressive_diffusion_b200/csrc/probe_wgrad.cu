// GPU probe for the weight-gradient kernel (MN-major UMMA operands) against a CPU reference.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "wgrad.cuh"
#include "tapconv_host.h"

using namespace ob;
static uint32_t rng_state = 777;
static float frand() { rng_state = rng_state * 1664525u + 1013904223u; return ((rng_state >> 8) & 0xFFFF) / 32768.0f - 1.0f; }
static float bf16r(float v) { return __bfloat162float(__float2bfloat16(v)); }

struct Problem {
  const char* name;
  int g_seq[2], g_T[2], a_T[2];
  int H, W, Cin, Cout, w_taps, n_split;
  int flat;   // WgradLaunch::force_mode
  std::vector<WgradItem> items;
};
static WgradItem mk(int pair, int dt, int dy, int dx, int wtap) { WgradItem t{}; t.pair = pair; t.dt = dt; t.dy = dy; t.dx = dx; t.wtap = wtap; return t; }

static bool run(const Problem& P, bool check, int reps) {
  std::vector<float> G[2], A[2];
  __nv_bfloat16 *dG[2] = {nullptr, nullptr}, *dA[2] = {nullptr, nullptr};
  for (int s = 0; s < 2; ++s) {
    if (!P.g_seq[s]) continue;
    size_t ng = (size_t)P.g_seq[s] * P.g_T[s] * P.H * P.W * P.Cout, na = (size_t)P.g_seq[s] * P.a_T[s] * P.H * P.W * P.Cin;
    G[s].resize(ng); A[s].resize(na);
    std::vector<__nv_bfloat16> hg(ng), ha(na);
    for (size_t i = 0; i < ng; ++i) { G[s][i] = bf16r(frand() * 0.25f); hg[i] = __float2bfloat16(G[s][i]); }
    for (size_t i = 0; i < na; ++i) { A[s][i] = bf16r(frand()); ha[i] = __float2bfloat16(A[s][i]); }
    cudaMalloc(&dG[s], ng * 2); cudaMalloc(&dA[s], na * 2);
    cudaMemcpy(dG[s], hg.data(), ng * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dA[s], ha.data(), na * 2, cudaMemcpyHostToDevice);
  }
  size_t nw = (size_t)P.Cout * P.w_taps * P.Cin;
  float* dOut; cudaMalloc(&dOut, nw * 4 * P.n_split); cudaMemset(dOut, 0xFF, nw * 4 * P.n_split);
  WgradLaunch L;
  for (int s = 0; s < 2; ++s) { L.g[s] = dG[s]; L.a[s] = dA[s]; L.g_seq[s] = P.g_seq[s]; L.g_T[s] = P.g_T[s]; L.a_T[s] = P.a_T[s]; }
  L.items = P.items.data(); L.n_items = (int)P.items.size(); L.H = P.H; L.W = P.W; L.Cin = P.Cin; L.Cout = P.Cout;
  L.w_taps = P.w_taps; L.n_split = P.n_split; L.out = dOut; L.force_mode = P.flat;
  int rc = wgrad_launch(L, 0);
  cudaError_t e = cudaDeviceSynchronize();
  if (rc != OB_OK || e != cudaSuccess) { printf("[%s] LAUNCH FAIL rc=%d (%s) cuda=%s\n", P.name, rc, last_error(), cudaGetErrorString(e)); return false; }
  bool ok = true;
  if (check) {
    std::vector<float> got(nw * P.n_split);
    cudaMemcpy(got.data(), dOut, nw * 4 * P.n_split, cudaMemcpyDeviceToHost);
    double max_err = 0, max_ref = 0;
    for (const WgradItem& it : P.items) {
      int s = it.pair;
      std::vector<double> ref((size_t)P.Cout * P.Cin, 0.0);
      for (int sq = 0; sq < P.g_seq[s]; ++sq) for (int t = 0; t < P.g_T[s]; ++t) for (int h = 0; h < P.H; ++h) for (int w = 0; w < P.W; ++w) {
        int tt = t + it.dt, hh = h + it.dy, ww = w + it.dx;
        if (tt < 0 || tt >= P.a_T[s] || hh < 0 || hh >= P.H || ww < 0 || ww >= P.W) continue;
        const float* gp = &G[s][((((size_t)sq * P.g_T[s] + t) * P.H + h) * P.W + w) * P.Cout];
        const float* ap = &A[s][((((size_t)sq * P.a_T[s] + tt) * P.H + hh) * P.W + ww) * P.Cin];
        for (int co = 0; co < P.Cout; ++co) { double g = gp[co]; double* r = &ref[(size_t)co * P.Cin]; for (int ci = 0; ci < P.Cin; ++ci) r[ci] += g * ap[ci]; }
      }
      for (int co = 0; co < P.Cout; ++co) for (int ci = 0; ci < P.Cin; ++ci) {
        double v = 0;
        for (int sp = 0; sp < P.n_split; ++sp) v += got[(size_t)sp * nw + ((size_t)co * P.w_taps + it.wtap) * P.Cin + ci];
        double r = ref[(size_t)co * P.Cin + ci];
        max_ref = fmax(max_ref, fabs(r)); max_err = fmax(max_err, fabs(r - v));
      }
    }
    ok = max_err / (max_ref + 1e-30) < 1e-4;
    printf("[%s] %s  max_err=%.3e max_ref=%.3e rel=%.3e\n", P.name, ok ? "PASS" : "FAIL", max_err, max_ref, max_err / (max_ref + 1e-30));
  }
  if (reps > 0) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) wgrad_launch(L, 0);
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) wgrad_launch(L, 0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    double flops = 0;
    for (const WgradItem& it : P.items) flops += 2.0 * P.g_seq[it.pair] * P.g_T[it.pair] * P.H * P.W * (double)P.Cin * P.Cout;
    printf("[%s] split %d time %.1f us  %.1f TFLOP/s\n", P.name, P.n_split, ms * 1e3, flops / ms / 1e9);
  }
  for (int s = 0; s < 2; ++s) { cudaFree(dG[s]); cudaFree(dA[s]); }
  cudaFree(dOut);
  return ok;
}

static Problem gated(const char* name, int B, int n, int H, int W, int Cin, int Cout, int split) {
  Problem P{}; P.name = name; P.g_seq[0] = 2 * B; P.g_T[0] = n; P.a_T[0] = n; P.g_seq[1] = B; P.g_T[1] = n; P.a_T[1] = n + 2;
  P.H = H; P.W = W; P.Cin = Cin; P.Cout = Cout; P.w_taps = 27;
  P.n_split = split > 0 ? split : wgrad_suggest_split(27, 2 * B * n, H, W, Cin, Cout);
  for (int ky = 0; ky < 3; ++ky) for (int kx = 0; kx < 3; ++kx) P.items.push_back(mk(0, 0, ky - 1, kx - 1, ky * 3 + kx));
  for (int tau = 0; tau < 2; ++tau) for (int ky = 0; ky < 3; ++ky) for (int kx = 0; kx < 3; ++kx) P.items.push_back(mk(1, tau, ky - 1, kx - 1, 9 + tau * 9 + ky * 3 + kx));
  return P;
}
static Problem plain(const char* name, int F, int H, int W, int Cin, int Cout, int k, int split) {
  Problem P{}; P.name = name; P.g_seq[0] = 1; P.g_T[0] = F; P.a_T[0] = F; P.H = H; P.W = W; P.Cin = Cin; P.Cout = Cout; P.w_taps = k * k;
  P.n_split = split > 0 ? split : wgrad_suggest_split(k * k, F, H, W, Cin, Cout);
  for (int ky = 0; ky < k; ++ky) for (int kx = 0; kx < k; ++kx) P.items.push_back(mk(0, 0, ky - k / 2, kx - k / 2, ky * k + kx));
  return P;
}

int main(int argc, char** argv) {
  bool perf = argc > 1 && atoi(argv[1]) > 0;
  int fails = 0;
  fails += !run(plain("wgrad 1x1 c64 n128 16x16 F4", 4, 16, 16, 64, 128, 1, 1), true, 0);
  fails += !run(plain("wgrad 1x1 c128 n64 8x8 F6 split3", 6, 8, 8, 128, 64, 1, 3), true, 0);
  fails += !run(plain("wgrad 3x3 c64 n64 8x8 F5", 5, 8, 8, 64, 64, 3, 2), true, 0);
  fails += !run(plain("wgrad 3x3 c256 n128 4x4 F9", 9, 4, 4, 256, 128, 3, 1), true, 0);
  fails += !run(plain("wgrad 3x3 c32 n64 16x16 (chunk32)", 3, 16, 16, 32, 64, 3, 2), true, 0);
  fails += !run(plain("wgrad 3x3 c16 n32 32x32 (chunk16)", 2, 32, 32, 16, 32, 3, 4), true, 0);
  fails += !run(plain("wgrad 3x3 c16 n8 64x64", 1, 64, 64, 16, 8, 3, 4), true, 0);
  fails += !run(plain("wgrad 3x3 c96 n40 10x12 F3 ragged", 3, 10, 12, 96, 40, 3, 1), true, 0);
  fails += !run(plain("wgrad linear c256 n128 T37", 37, 1, 1, 256, 128, 1, 1), true, 0);
  fails += !run(gated("wgrad gated c128 n128 8x8 B2 n4", 2, 4, 8, 8, 128, 128, 2), true, 0);
  fails += !run(gated("wgrad gated c64 n192 4x4 B1 n8", 1, 8, 4, 4, 64, 192, 0), true, 0);
  // CTA pairs (Cout a multiple of 256) and two-tap groups (Cin = 128), alone and combined
  { Problem q = gated("wgrad gated c256 n256 8x8 B1 n4 (pair)", 1, 4, 8, 8, 256, 256, 2); q.flat = 2; fails += !run(q, true, 0); }
  { Problem q = gated("wgrad gated c128 n256 8x8 B1 n4 (pair + two-tap)", 1, 4, 8, 8, 128, 256, 2); q.flat = 2; fails += !run(q, true, 0); }
  fails += !run(gated("wgrad gated c128 n256 8x8 B1 n4 (two-tap)", 1, 4, 8, 8, 128, 256, 2), true, 0);
  fails += !run(plain("wgrad 3x3 c128 n128 8x8 F6 (two-tap, odd count)", 6, 8, 8, 128, 128, 3, 2), true, 0);
  { Problem q = plain("wgrad 1x1 c512 n512 4x4 F8 (pair)", 8, 4, 4, 512, 512, 1, 1); q.flat = 2; fails += !run(q, true, 0); }
  { Problem q = plain("wgrad 3x3 c384 n256 8x8 F3 (pair, ragged ci tile)", 3, 8, 8, 384, 256, 3, 1); q.flat = 2; fails += !run(q, true, 0); }
  { Problem q = gated("wgrad gated c256 n256 8x8 B1 n4 (halo)", 1, 4, 8, 8, 256, 256, 2); q.flat = 3; fails += !run(q, true, 0); }
  { Problem q = plain("wgrad 3x3 c256 n128 4x4 F9 (halo)", 9, 4, 4, 256, 128, 3, 1); q.flat = 3; fails += !run(q, true, 0); }
  { Problem q = plain("wgrad 3x3 c384 n256 16x16 F3 (halo, ragged ci tile)", 3, 16, 16, 384, 256, 3, 1); q.flat = 3; fails += !run(q, true, 0); }
  printf("== wgrad correctness: %d failing ==\n", fails);
  if (perf) {
    run(gated("CS 512->512 16x16 B2 n16", 2, 16, 16, 16, 512, 512, 0), false, 20);
    run(gated("CS 256->256 16x16 B2 n16", 2, 16, 16, 16, 256, 256, 0), false, 20);
    run(gated("CS 128->128 32x32 B2 n16", 2, 16, 32, 32, 128, 128, 0), false, 20);
    run(gated("CS 512->512 8x8 B2 n16", 2, 16, 8, 8, 512, 512, 0), false, 20);
    run(gated("CS 512->512 4x4 B2 n16", 2, 16, 4, 4, 512, 512, 0), false, 20);
    run(gated("CS 1024->512 8x8 B2 n16", 2, 16, 8, 8, 1024, 512, 0), false, 20);
    { Problem q = gated("CS 512->512 16x16 FLAT", 2, 16, 16, 16, 512, 512, 0); q.flat = 1; run(q, false, 20); }
    { Problem q = gated("CS 256->256 16x16 FLAT", 2, 16, 16, 16, 256, 256, 0); q.flat = 1; run(q, false, 20); }
    { Problem q = gated("CS 128->128 32x32 FLAT", 2, 16, 32, 32, 128, 128, 0); q.flat = 1; run(q, false, 20); }
    { Problem q = gated("CS 512->512 8x8 FLAT", 2, 16, 8, 8, 512, 512, 0); q.flat = 1; run(q, false, 20); }
    { Problem q = gated("CS 512->512 4x4 FLAT", 2, 16, 4, 4, 512, 512, 0); q.flat = 1; run(q, false, 20); }
    { Problem q = gated("CS 512->512 16x16 PAIR", 2, 16, 16, 16, 512, 512, 0); q.flat = 2; run(q, false, 20); }
    { Problem q = gated("CS 512->512 16x16 HALO", 2, 16, 16, 16, 512, 512, 0); q.flat = 3; run(q, false, 20); }
    { Problem q = gated("CS 1024->512 8x8 HALO", 2, 16, 8, 8, 1024, 512, 0); q.flat = 3; run(q, false, 20); }
    run(gated("CS 256->128 32x32 B2 n16", 2, 16, 32, 32, 256, 128, 0), false, 20);
    for (int sp : {2, 3, 4, 6, 8}) run(gated("CS 256->256 16x16 splitN", 2, 16, 16, 16, 256, 256, sp), false, 20);
    for (int sp : {6, 11, 16, 22}) run(gated("CS 128->128 32x32 splitN", 2, 16, 32, 32, 128, 128, sp), false, 20);
    for (int sp : {3, 6, 11}) run(gated("CS 256->128 32x32 splitN", 2, 16, 32, 32, 256, 128, sp), false, 20);
    for (int sp : {2, 3, 6}) run(gated("CS 256->256 8x8 splitN", 2, 16, 8, 8, 256, 256, sp), false, 20);
    run(gated("CS 512->512 16x16 split1", 2, 16, 16, 16, 512, 512, 1), false, 20);
    run(gated("CS 512->512 16x16 split4", 2, 16, 16, 16, 512, 512, 4), false, 20);
  }
  return fails;
}
