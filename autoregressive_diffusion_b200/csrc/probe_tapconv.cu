// Standalone GPU probe for the tap-GEMM kernel: checks each layout/shape family against a CPU
// reference computed from the same bf16-rounded inputs, then times the Counter-Strike-sized layers.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <climits>
#include "tapconv.cuh"
#include "tapconv_host.h"

using namespace ob;

static uint32_t rng_state = 12345;
static float frand() {
  rng_state = rng_state * 1664525u + 1013904223u;
  return ((rng_state >> 8) & 0xFFFF) / 32768.0f - 1.0f;
}
static float bf16r(float v) { return __bfloat162float(__float2bfloat16(v)); }

struct Problem {
  const char* name;
  int n_seq, n_out, T, H, W, Cin, Cout, epi, out_f32;
  int seqA[2], TA[2];
  int w_taps;
  std::vector<TapCol> cols;
  int force_bn;
  int bmn;
  int halo;
  int split;   // give the launch a split-K workspace
  int nopair;  // 1: forbid CTA-pair execution
  int nopersist;  // 1: one CTA per tile
};

// column of vertical taps at horizontal shift dx; flip mirrors the kernel (input-gradient form)
static TapCol mk3(int src, int dt, int dx, int n_a, int acc, int seq_mul, int tap_base, bool flip) {
  TapCol c{};
  c.src = src; c.dt = dt; c.dx = dx; c.n_a = n_a; c.acc = acc; c.seq_mul = seq_mul; c.n_taps = 3;
  for (int d = 0; d < 3; ++d) {
    int dy = d - 1;
    int ky = flip ? 1 - dy : dy + 1, kx = flip ? 1 - dx : dx + 1;
    c.wtap[d] = tap_base + ky * 3 + kx;
  }
  return c;
}
static TapCol mk1(int src, int n_a, int acc, int seq_mul) {
  TapCol c{};
  c.src = src; c.n_a = n_a; c.acc = acc; c.seq_mul = seq_mul; c.n_taps = 1; c.wtap[0] = 0;
  return c;
}

static bool run(const Problem& P, bool check, int reps) {
  const int n_acc = P.n_out + (P.epi == EPI_GATED);
  std::vector<std::vector<float>> A(2);
  std::vector<__nv_bfloat16*> dA(2, nullptr);
  for (int s = 0; s < 2; ++s) {
    if (P.seqA[s] == 0) continue;
    size_t n = (size_t)P.seqA[s] * P.TA[s] * P.H * P.W * P.Cin;
    A[s].resize(n);
    std::vector<__nv_bfloat16> h(n);
    for (size_t i = 0; i < n; ++i) { A[s][i] = bf16r(frand()); h[i] = __float2bfloat16(A[s][i]); }
    cudaMalloc(&dA[s], n * 2);
    cudaMemcpy(dA[s], h.data(), n * 2, cudaMemcpyHostToDevice);
  }
  size_t nw = (size_t)P.Cout * P.w_taps * P.Cin;
  std::vector<float> Wt(nw);
  std::vector<__nv_bfloat16> hw(nw);
  // Wt is indexed [n][tap][c]; in MN-major mode the device copy is stored [c][tap][n]
  for (size_t i = 0; i < nw; ++i) Wt[i] = bf16r(frand() * 0.1f);
  for (int n = 0; n < P.Cout; ++n) for (int tp = 0; tp < P.w_taps; ++tp) for (int c = 0; c < P.Cin; ++c) {
    size_t src = ((size_t)n * P.w_taps + tp) * P.Cin + c;
    size_t dst = P.bmn ? ((size_t)c * P.w_taps + tp) * P.Cout + n : src;
    hw[dst] = __float2bfloat16(Wt[src]);
  }
  __nv_bfloat16* dW; cudaMalloc(&dW, nw * 2); cudaMemcpy(dW, hw.data(), nw * 2, cudaMemcpyHostToDevice);
  const int frames = P.n_seq * P.n_out * P.T;
  std::vector<float> al(frames), be(frames);
  for (int i = 0; i < frames; ++i) { al[i] = 0.5f + 0.5f * frand(); be[i] = 0.3f * frand(); }
  float *dal, *dbe; cudaMalloc(&dal, frames * 4); cudaMalloc(&dbe, frames * 4);
  cudaMemcpy(dal, al.data(), frames * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dbe, be.data(), frames * 4, cudaMemcpyHostToDevice);
  size_t nout = (size_t)frames * P.H * P.W * P.Cout;
  void* dOut; cudaMalloc(&dOut, nout * 4); cudaMemset(dOut, 0xFF, nout * 4);
  float* dD; cudaMalloc(&dD, nout * 4); cudaMemset(dD, 0xFF, nout * 4);

  TapConvLaunch L;
  for (int s = 0; s < 2; ++s) {
    L.a[s] = dA[s]; L.a_seq[s] = P.seqA[s]; L.a_T[s] = P.TA[s];
    L.a_stride_w[s] = P.Cin; L.a_stride_h[s] = (long)P.W * P.Cin; L.a_stride_t[s] = (long)P.H * P.W * P.Cin;
    L.a_stride_seq[s] = (long)P.TA[s] * P.H * P.W * P.Cin;
  }
  L.wg = dW; L.w_taps = P.w_taps; L.cols = P.cols.data(); L.n_cols = (int)P.cols.size(); L.halo = P.halo;
  L.n_seq = P.n_seq; L.n_out = P.n_out; L.T = P.T; L.H = P.H; L.W = P.W; L.Cin = P.Cin; L.Cout = P.Cout;
  L.epi = P.epi; L.out_f32 = P.out_f32; L.alpha = dal; L.beta = dbe; L.out = dOut;
  L.out_d = (P.epi == EPI_GATED) ? dD : nullptr; L.force_bn = P.force_bn; L.b_mn_major = P.bmn;
  L.use_pair = P.nopair ? 0 : 1;
  L.no_persist = P.nopersist;
  float* dWs = nullptr;
  if (P.split) { cudaMalloc(&dWs, (size_t)n_acc * nout / P.n_out * 4); L.split_ws = dWs; }

  int rc = tapconv_launch(L, 0);
  cudaError_t e = cudaDeviceSynchronize();
  if (rc != OB_OK || e != cudaSuccess) {
    printf("[%s] LAUNCH FAIL rc=%d (%s) cuda=%s\n", P.name, rc, last_error(), cudaGetErrorString(e));
    return false;
  }
  bool ok = true;
  if (check) {
    std::vector<float> got(nout);
    if (P.out_f32) cudaMemcpy(got.data(), dOut, nout * 4, cudaMemcpyDeviceToHost);
    else {
      std::vector<__nv_bfloat16> hb(nout);
      cudaMemcpy(hb.data(), dOut, nout * 2, cudaMemcpyDeviceToHost);
      for (size_t i = 0; i < nout; ++i) got[i] = __bfloat162float(hb[i]);
    }
    std::vector<float> hd(nout);
    cudaMemcpy(hd.data(), dD, nout * 4, cudaMemcpyDeviceToHost);
    double max_err = 0, max_ref = 0, max_err_d = 0;
    std::vector<double> acc(n_acc);
    for (int seq = 0; seq < P.n_seq; ++seq)
      for (int t = 0; t < P.T; ++t)
        for (int h = 0; h < P.H; ++h)
          for (int w = 0; w < P.W; ++w)
            for (int n = 0; n < P.Cout; ++n) {
              for (auto& a : acc) a = 0;
              for (const TapCol& it : P.cols)
                for (int d = 0; d < it.n_taps; ++d)
                  for (int i = 0; i < it.n_a; ++i) {
                    int dy = it.n_taps == 3 ? d - 1 : 0;
                    int sq = seq * it.seq_mul + i, tt = t + it.dt, hh = h + dy, ww = w + it.dx;
                    if (tt < 0 || tt >= P.TA[it.src] || hh < 0 || hh >= P.H || ww < 0 || ww >= P.W) continue;
                    const float* ap = &A[it.src][((((size_t)sq * P.TA[it.src] + tt) * P.H + hh) * P.W + ww) * P.Cin];
                    const float* wp = &Wt[((size_t)n * P.w_taps + it.wtap[d]) * P.Cin];
                    double s = 0;
                    for (int c = 0; c < P.Cin; ++c) s += (double)ap[c] * wp[c];
                    acc[it.acc + i] += s;
                  }
              for (int o = 0; o < P.n_out; ++o) {
                size_t frame = (size_t)(seq * P.n_out + o) * P.T + t;
                size_t idx = ((frame * P.H + h) * P.W + w) * P.Cout + n;
                double ref = acc[o], refd = 0;
                if (P.epi == EPI_GATED) { ref = al[frame] * acc[o] + be[frame] * acc[P.n_out]; refd = acc[P.n_out] - acc[o]; }
                max_ref = fmax(max_ref, fabs(ref));
                max_err = fmax(max_err, fabs(ref - got[idx]));
                if (P.epi == EPI_GATED) max_err_d = fmax(max_err_d, fabs(refd - hd[idx]));
              }
            }
    double rel = max_err / (max_ref + 1e-30);
    ok = rel < (P.out_f32 ? 2e-4 : 1e-2) && (max_err_d / (max_ref + 1e-30) < 2e-2);
    printf("[%s] %s  max_err=%.3e max_ref=%.3e rel=%.3e  d_err=%.3e\n", P.name, ok ? "PASS" : "FAIL", max_err, max_ref,
           rel, max_err_d);
  }
  if (reps > 0) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) tapconv_launch(L, 0);
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) tapconv_launch(L, 0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    double flops = 0;
    for (const TapCol& it : P.cols) flops += 2.0 * it.n_taps * it.n_a * P.n_seq * P.T * P.H * P.W * (double)P.Cin * P.Cout;
    printf("[%s] time %.1f us  %.1f TFLOP/s  (%.3f GFLOP)\n", P.name, ms * 1e3, flops / ms / 1e9, flops / 1e9);
    if (getenv("TAPCONV_TRACE")) {
      const int max_ctas = 1 << 16;
      long long* dT; cudaMalloc(&dT, (size_t)max_ctas * 8 * 8); cudaMemset(dT, 0, (size_t)max_ctas * 8 * 8);
      L.trace = dT;
      tapconv_launch(L, 0); cudaDeviceSynchronize();
      L.trace = nullptr;
      std::vector<long long> hT((size_t)max_ctas * 8);
      cudaMemcpy(hT.data(), dT, hT.size() * 8, cudaMemcpyDeviceToHost);
      long long tmin = LLONG_MAX, tmax = 0; int n = 0;
      for (int i = 0; i < max_ctas; ++i) if (hT[i * 8]) { tmin = std::min(tmin, hT[i * 8]); ++n; }
      double s01 = 0, s12 = 0, s23 = 0, s34 = 0; int n2 = 0, n34 = 0;
      std::vector<double> starts;
      for (int i = 0; i < max_ctas; ++i) {
        const long long* t = &hT[i * 8];
        if (!t[0]) continue;
        starts.push_back((t[0] - tmin) * 1e-3);
        if (t[2]) { s01 += t[1] - t[0]; s12 += t[2] - t[1]; s23 += t[3] - t[2]; ++n2; }
        if (t[3] && t[4]) { s34 += t[4] - t[3]; ++n34; tmax = std::max(tmax, t[4]); }
      }
      std::sort(starts.begin(), starts.end());
      printf("  trace: %d CTAs (%d leaders)  setup->first tile %.2f us  mainloop issue %.2f us  drain %.2f us  epilogue %.2f us  span %.1f us\n",
             n, n2, s01 / n2 * 1e-3, s12 / n2 * 1e-3, s23 / n2 * 1e-3, s34 / n34 * 1e-3, (tmax - tmin) * 1e-3);
      printf("  CTA start times (us): p0 %.1f p25 %.1f p50 %.1f p60 %.1f p75 %.1f p90 %.1f p100 %.1f\n", starts[0], starts[n / 4], starts[n / 2],
             starts[n * 6 / 10], starts[n * 3 / 4], starts[n * 9 / 10], starts[n - 1]);
      cudaFree(dT);
    }
  }
  for (int s = 0; s < 2; ++s) cudaFree(dA[s]);
  cudaFree(dW); cudaFree(dal); cudaFree(dbe); cudaFree(dOut); cudaFree(dD);
  return ok;
}

static Problem gated(const char* name, int B, int S, int n, int H, int W, int Cin, int Cout, int force_bn = 0, int split = 0) {
  Problem P{};
  P.split = split;
  P.name = name; P.n_seq = B; P.n_out = S; P.T = n; P.H = H; P.W = W; P.Cin = Cin; P.Cout = Cout;
  P.epi = EPI_GATED; P.out_f32 = 0; P.seqA[0] = B * S; P.TA[0] = n; P.seqA[1] = B; P.TA[1] = n + 2; P.w_taps = 27;
  P.force_bn = force_bn; P.halo = 1;
  const bool cur_first = getenv("TAPCONV_CUR_FIRST") != nullptr;
  if (cur_first) for (int dx = -1; dx <= 1; ++dx) P.cols.push_back(mk3(0, 0, dx, S, 0, S, 0, false));
  for (int tau = 0; tau < 2; ++tau)   // context taps first (see capi.cu)
    for (int dx = -1; dx <= 1; ++dx) P.cols.push_back(mk3(1, tau, dx, 1, S, 1, 9 + tau * 9, false));
  if (!cur_first) for (int dx = -1; dx <= 1; ++dx) P.cols.push_back(mk3(0, 0, dx, S, 0, S, 0, false));
  return P;
}
static Problem plain(const char* name, int F, int H, int W, int Cin, int Cout, int k, int f32, int force_bn = 0) {
  Problem P{};
  P.name = name; P.n_seq = 1; P.n_out = 1; P.T = F; P.H = H; P.W = W; P.Cin = Cin; P.Cout = Cout;
  P.epi = EPI_PLAIN; P.out_f32 = f32; P.seqA[0] = 1; P.TA[0] = F; P.seqA[1] = 0; P.TA[1] = 0; P.w_taps = k * k;
  P.force_bn = force_bn; P.halo = (k == 3);
  if (k == 1) P.cols.push_back(mk1(0, 1, 0, 1));
  else for (int dx = -1; dx <= 1; ++dx) P.cols.push_back(mk3(0, 0, dx, 1, 0, 1, 0, false));
  return P;
}
// input-gradient shape of the gated conv: dual rows from src0, causal terms (dt=+1,+2) from src1 into acc 0
static Problem dgrad(const char* name, int B, int n, int H, int W, int Cin, int Cout, int bmn = 0) {
  Problem P{};
  P.bmn = bmn;
  P.name = name; P.n_seq = B; P.n_out = 2; P.T = n; P.H = H; P.W = W; P.Cin = Cin; P.Cout = Cout;
  P.epi = EPI_PLAIN; P.out_f32 = 0; P.seqA[0] = B * 2; P.TA[0] = n; P.seqA[1] = B; P.TA[1] = n; P.w_taps = 27;
  P.halo = 1;
  for (int dx = -1; dx <= 1; ++dx) P.cols.push_back(mk3(0, 0, dx, 2, 0, 2, 0, true));
  for (int tau = 0; tau < 2; ++tau)
    for (int dx = -1; dx <= 1; ++dx) P.cols.push_back(mk3(1, 2 - tau, dx, 1, 0, 1, 9 + tau * 9, true));
  return P;
}

int main(int argc, char** argv) {
  bool perf = argc > 1 && atoi(argv[1]) > 0;
  int fails = 0;
  if (argc > 1 && atoi(argv[1]) == 2) {   // a short list for ncu captures
    run(gated("CS 128->128 32x32 B2 n16", 2, 2, 16, 32, 32, 128, 128), false, 2);
    { Problem q = gated("CS 128->128 32x32 B2 n16 NOPERSIST", 2, 2, 16, 32, 32, 128, 128); q.nopersist = 1; run(q, false, 2); }
    run(gated("CS 512->512 16x16 B2 n16", 2, 2, 16, 16, 16, 512, 512), false, 2);
    run(plain("1x1 256->128 32x32 F64", 64, 32, 32, 256, 128, 1, 0), false, 2);
    run(gated("CS 512->512 8x8 split auto", 2, 2, 16, 8, 8, 512, 512, 0, 1), false, 2);
    run(gated("CS 512->512 4x4 split auto", 2, 2, 16, 4, 4, 512, 512, 0, 1), false, 2);
    run(gated("CS 512->512 4x4 nosplit", 2, 2, 16, 4, 4, 512, 512, 0, 0), false, 2);
    run(dgrad("dgrad BMN 128->128 32x32 B2 n16", 2, 16, 32, 32, 128, 128, 1), false, 2);
    return 0;
  }
  fails += !run(plain("gemm1x1 c64 n128 16x16 f32", 4, 16, 16, 64, 128, 1, 1), true, 0);
  fails += !run(plain("gemm1x1 c128 n64 16x16", 4, 16, 16, 128, 64, 1, 0), true, 0);
  fails += !run(plain("conv3x3 c64 n64 8x8 T6", 6, 8, 8, 64, 64, 3, 0), true, 0);
  fails += !run(plain("conv3x3 c32 n64 16x16 (chunk32)", 3, 16, 16, 32, 64, 3, 0), true, 0);
  fails += !run(plain("conv3x3 c16 n32 32x32 (chunk16)", 2, 32, 32, 16, 32, 3, 0), true, 0);
  fails += !run(plain("conv3x3 c16 n8 64x64 (N pad)", 1, 64, 64, 16, 8, 3, 1), true, 0);
  fails += !run(plain("conv3x3 c64 n256 4x4 T20 bn256", 20, 4, 4, 64, 256, 3, 0, 256), true, 0);
  fails += !run(plain("conv3x3 c96 n40 10x12 T3 (ragged)", 3, 10, 12, 96, 40, 3, 0), true, 0);
  fails += !run(plain("linear c256 n128 1x1 T37", 37, 1, 1, 256, 128, 1, 1), true, 0);
  fails += !run(gated("gated dual c128 n128 8x8 B2 n4", 2, 2, 4, 8, 8, 128, 128), true, 0);
  fails += !run(gated("gated dual c64 n64 4x4 B2 n8", 2, 2, 8, 4, 4, 64, 64), true, 0);
  fails += !run(gated("gated dual c64 n128 16x16 B1 n3 bn128", 1, 2, 3, 16, 16, 64, 128, 128), true, 0);
  fails += !run(gated("gated eval c64 n256 8x8 B2 T5 bn256", 2, 1, 5, 8, 8, 64, 256, 256), true, 0);
  fails += !run(gated("gated eval decode c128 n128 4x4 B3 T1", 3, 1, 1, 4, 4, 128, 128), true, 0);
  fails += !run(gated("gated dual split c256 n128 4x4 B1 n8", 1, 2, 8, 4, 4, 256, 128, 0, 1), true, 0);
  fails += !run(gated("gated eval split c256 n64 4x4 B2 T4", 2, 1, 4, 4, 4, 256, 64, 0, 1), true, 0);
  // more tiles than SMs: several rounds of the persistent loop in every TMEM mode (ROTATE, DOUBLE, SINGLE), pair and single
  fails += !run(gated("gated dual c64 n128 32x32 B2 n12 (rotate, 3 rounds)", 2, 2, 12, 32, 32, 64, 128), true, 0);
  { Problem q = gated("gated dual c64 n128 32x32 B2 n12 nopair", 2, 2, 12, 32, 32, 64, 128); q.nopair = 1; fails += !run(q, true, 0); }
  fails += !run(gated("gated dual c64 n64 32x32 B2 n10 (double)", 2, 2, 10, 32, 32, 64, 64), true, 0);
  fails += !run(gated("gated eval c64 n256 32x32 B2 T10 (single)", 2, 1, 10, 32, 32, 64, 256, 256), true, 0);
  fails += !run(plain("conv3x3 c64 n128 32x32 F40 (double, pair)", 40, 32, 32, 64, 128, 3, 0), true, 0);
  fails += !run(plain("gemm1x1 c64 n256 32x32 F30 f32", 30, 32, 32, 64, 256, 1, 1), true, 0);
  fails += !run(dgrad("dgrad dual c128 n64 8x8 B2 n4", 2, 4, 8, 8, 128, 64), true, 0);
  fails += !run(dgrad("dgrad dual BMN c128 n64 8x8 B2 n4", 2, 4, 8, 8, 128, 64, 1), true, 0);
  fails += !run(dgrad("dgrad dual BMN c64 n256 4x4 B2 n8", 2, 8, 4, 4, 64, 256, 1), true, 0);
  fails += !run(dgrad("dgrad dual BMN c32 n96 8x8 B1 n3", 1, 3, 8, 8, 32, 96, 1), true, 0);
  fails += !run(dgrad("dgrad dual BMN c16 n32 16x16 B1 n2", 1, 2, 16, 16, 16, 32, 1), true, 0);
  fails += !run(dgrad("dgrad dual BMN c8(16) n64 8x8 B1 n2", 1, 2, 8, 8, 16, 64, 1), true, 0);
  printf("== correctness: %d failing ==\n", fails);
  if (perf) {
    run(gated("CS 512->512 16x16 B2 n16", 2, 2, 16, 16, 16, 512, 512), false, 20);
    run(gated("CS 256->256 16x16 B2 n16", 2, 2, 16, 16, 16, 256, 256), false, 20);
    run(gated("CS 128->128 32x32 B2 n16", 2, 2, 16, 32, 32, 128, 128), false, 20);
    run(gated("CS 512->512 8x8 B2 n16", 2, 2, 16, 8, 8, 512, 512), false, 20);
    run(gated("CS 512->512 4x4 B2 n16", 2, 2, 16, 4, 4, 512, 512), false, 20);
    run(gated("CS 1024->512 8x8 B2 n16", 2, 2, 16, 8, 8, 1024, 512), false, 20);
    run(gated("CS 512->512 16x16 bn64", 2, 2, 16, 16, 16, 512, 512, 64), false, 20);
    run(gated("CS 512->512 8x8 bn128", 2, 2, 16, 8, 8, 512, 512, 128), false, 20);
    run(gated("CS 512->512 8x8 bn64", 2, 2, 16, 8, 8, 512, 512, 64), false, 20);
    run(gated("CS 512->512 8x8 bn32", 2, 2, 16, 8, 8, 512, 512, 32), false, 20);
    { Problem q = gated("CS 512->512 16x16 NOPAIR", 2, 2, 16, 16, 16, 512, 512); q.nopair = 1; run(q, false, 20); }
    { Problem q = gated("CS 256->256 16x16 bn128 NOPAIR", 2, 2, 16, 16, 16, 256, 256, 128); q.nopair = 1; run(q, false, 20); }
    { Problem q = plain("conv3x3 512->512 16x16 F64 bn256 NOPAIR", 64, 16, 16, 512, 512, 3, 0, 256); q.nopair = 1; run(q, false, 20); }
    run(gated("CS 512->512 8x8 split auto", 2, 2, 16, 8, 8, 512, 512, 0, 1), false, 20);
    run(gated("CS 512->512 8x8 split bn128", 2, 2, 16, 8, 8, 512, 512, 128, 1), false, 20);
    run(gated("CS 1024->512 8x8 split auto", 2, 2, 16, 8, 8, 1024, 512, 0, 1), false, 20);
    run(gated("CS 512->512 4x4 split auto", 2, 2, 16, 4, 4, 512, 512, 0, 1), false, 20);
    run(gated("CS 512->512 4x4 split bn128", 2, 2, 16, 4, 4, 512, 512, 128, 1), false, 20);
    run(gated("CS 256->256 16x16 bn128", 2, 2, 16, 16, 16, 256, 256, 128), false, 20);
    run(gated("CS 256->256 16x16 bn64", 2, 2, 16, 16, 16, 256, 256, 64), false, 20);
    run(gated("CS 128->128 32x32 bn64", 2, 2, 16, 32, 32, 128, 128, 64), false, 20);
    run(gated("CS 512->512 4x4 bn128", 2, 2, 16, 4, 4, 512, 512, 128), false, 20);
    run(gated("CS 512->512 4x4 bn64", 2, 2, 16, 4, 4, 512, 512, 64), false, 20);
    run(gated("CS 512->512 4x4 bn32", 2, 2, 16, 4, 4, 512, 512, 32), false, 20);
    run(plain("gemm 16384x512x1536", 64, 16, 16, 512, 1536, 1, 0), false, 20);
    for (int bnf : {0, 256, 128, 64}) {
      { Problem q = plain("1x1 512->512 4x4 F64", 64, 4, 4, 512, 512, 1, 0, bnf); q.split = 1; printf("bn=%d ", bnf); run(q, false, 20); }
      { Problem q = plain("1x1 512->1536 4x4 F64", 64, 4, 4, 512, 1536, 1, 0, bnf); q.split = 1; printf("bn=%d ", bnf); run(q, false, 20); }
      { Problem q = plain("1x1 512->512 8x8 F64", 64, 8, 8, 512, 512, 1, 0, bnf); q.split = 1; printf("bn=%d ", bnf); run(q, false, 20); }
      { Problem q = plain("1x1 512->1536 8x8 F64", 64, 8, 8, 512, 1536, 1, 0, bnf); q.split = 1; printf("bn=%d ", bnf); run(q, false, 20); }
      { Problem q = plain("1x1 256->128 32x32 F64", 64, 32, 32, 256, 128, 1, 0, bnf); q.split = 1; printf("bn=%d ", bnf); run(q, false, 20); }
    }
    run(plain("gemm 16384x512x512 bn128", 64, 16, 16, 512, 512, 1, 0, 128), false, 20);
    run(plain("conv3x3 512->512 16x16 F64 bn256", 64, 16, 16, 512, 512, 3, 0, 256), false, 20);
    run(plain("conv3x3 512->512 16x16 F64 bn128", 64, 16, 16, 512, 512, 3, 0, 128), false, 20);
    run(dgrad("dgrad BMN 512->512 16x16 B2 n16", 2, 16, 16, 16, 512, 512, 1), false, 20);
    run(dgrad("dgrad BMN 128->128 32x32 B2 n16", 2, 16, 32, 32, 128, 128, 1), false, 20);
  }
  return fails;
}
