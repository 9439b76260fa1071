// Development harness (GPU box only): runs the tcgen05 implicit-GEMM conv against a naive fp32
// conv on the same bf16-rounded operands, for several shapes, and times it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tc_conv_test tc_conv_test.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../hconv_tc.cuh"
#include "../tc_host.cuh"

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);  \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)

__global__ void naive_conv(const float* in, const float* w, const float* bias, float* out, int N,
                           int H, int W, int k, int S) {
  size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t total = (size_t)N * H * W * k;
  if (idx >= total) return;
  int co = idx % k;
  size_t p = idx / k;
  int x = p % W;
  int y = (p / W) % H;
  int n = p / ((size_t)W * H);
  int pad = (S - 1) / 2;
  float acc = 0.f;
  for (int dy = 0; dy < S; ++dy) {
    int yy = y + dy - pad;
    if (yy < 0 || yy >= H) continue;
    for (int dx = 0; dx < S; ++dx) {
      int xx = x + dx - pad;
      if (xx < 0 || xx >= W) continue;
      const float* ip = in + ((size_t)(n * H + yy) * W + xx) * k;
      const float* wp = w + ((size_t)(dy * S + dx) * k) * k + co;
      for (int ci = 0; ci < k; ++ci) acc += ip[ci] * wp[(size_t)ci * k];
    }
  }
  out[idx] = acc + bias[co];
}

static float bf16r(float v) { return __bfloat162float(__float2bfloat16(v)); }

template <int S, int KSTEPS, int CO_PAD, int TILES_X, int G, int WSTAGES>
int run_case(int N, int H, int W, int k, int grid_override, int iters) {
  using Cfg = hgru::TcConvCfg<S, KSTEPS, CO_PAD, TILES_X, G, WSTAGES>;
  const int CG = 2 * KSTEPS;
  printf("case S=%d k=%d (KSTEPS=%d CO_PAD=%d TILES_X=%d G=%d) N=%d H=%d W=%d smem=%d\n", S, k,
         KSTEPS, CO_PAD, TILES_X, G, N, H, W, Cfg::kSmemBytes);
  size_t npix = (size_t)N * H * W;
  std::vector<float> in(npix * k), w((size_t)S * S * k * k), bias(CO_PAD, 0.f);
  srand(123);
  for (auto& v : in) v = bf16r((rand() / (float)RAND_MAX) * 2.f - 1.f);
  for (auto& v : w) v = bf16r(((rand() / (float)RAND_MAX) * 2.f - 1.f) * 0.05f);
  for (int c = 0; c < k; ++c) bias[c] = (rand() / (float)RAND_MAX) - 0.5f;
  // operand copies
  std::vector<__nv_bfloat16> act((size_t)N * CG * H * W * 8, __float2bfloat16(0.f));
  for (int n = 0; n < N; ++n)
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x)
        for (int c = 0; c < k; ++c)
          act[((((size_t)n * CG + c / 8) * H + y) * W + x) * 8 + (c % 8)] =
              __float2bfloat16(in[(((size_t)n * H + y) * W + x) * k + c]);
  std::vector<__nv_bfloat16> wpk((size_t)KSTEPS * S * S * 2 * CO_PAD * 8, __float2bfloat16(0.f));
  for (int q = 0; q < KSTEPS; ++q)
    for (int tap = 0; tap < S * S; ++tap)
      for (int ch = 0; ch < 2; ++ch)
        for (int co = 0; co < k; ++co)
          for (int j = 0; j < 8; ++j) {
            int ci = q * 16 + ch * 8 + j;
            if (ci < k)
              wpk[((((size_t)q * S * S + tap) * 2 + ch) * CO_PAD + co) * 8 + j] =
                  __float2bfloat16(w[((size_t)tap * k + ci) * k + co]);
          }
  float *d_in, *d_w, *d_bias, *d_ref, *d_out;
  __nv_bfloat16 *d_act, *d_wpk;
  CK(cudaMalloc(&d_in, in.size() * 4));
  CK(cudaMalloc(&d_w, w.size() * 4));
  CK(cudaMalloc(&d_bias, CO_PAD * 4));
  CK(cudaMalloc(&d_ref, npix * k * 4));
  CK(cudaMalloc(&d_out, npix * CO_PAD * 4));
  CK(cudaMalloc(&d_act, act.size() * 2));
  CK(cudaMalloc(&d_wpk, wpk.size() * 2));
  CK(cudaMemcpy(d_in, in.data(), in.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_w, w.data(), w.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_bias, bias.data(), CO_PAD * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_act, act.data(), act.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_wpk, wpk.data(), wpk.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(d_out, 0xFF, npix * CO_PAD * 4));

  CUtensorMap map;
  int rc = hgru::make_act_tensor_map(&map, d_act, N, CG, H, W, Cfg::kCols, Cfg::kRows);
  if (rc) { printf("tensor map encode failed rc=%d\n", rc); return 1; }

  hgru::TcConvArgs a;
  a.N = N; a.H = H; a.W = W; a.KP = CO_PAD; a.kreal = k; a.scale = nullptr; a.shift = nullptr; a.out_bf16 = nullptr;
  a.units_x = (W + 8 * TILES_X - 1) / (8 * TILES_X);
  a.units_y = (H + hgru::kTileRows - 1) / hgru::kTileRows;
  a.num_units = N * a.units_x * a.units_y;
  a.wpk = d_wpk; a.bias = d_bias; a.out = d_out;
  auto kern = hgru::hconv_tc_kernel<S, KSTEPS, CO_PAD, TILES_X, G, WSTAGES, hgru::EpiBias>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  int grid = grid_override > 0 ? grid_override : (a.num_units < sms ? a.num_units : sms);
  kern<<<grid, 256, Cfg::kSmemBytes>>>(map, a);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());

  size_t total = npix * k;
  naive_conv<<<(unsigned)((total + 255) / 256), 256>>>(d_in, d_w, d_bias, d_ref, N, H, W, k, S);
  CK(cudaDeviceSynchronize());
  std::vector<float> out(total), ref(total), outp(npix * CO_PAD);
  CK(cudaMemcpy(outp.data(), d_out, npix * CO_PAD * 4, cudaMemcpyDeviceToHost));
  for (size_t p = 0; p < npix; ++p)
    for (int c = 0; c < k; ++c) {
      size_t nn = p / ((size_t)H * W), pin = p % ((size_t)H * W);
      out[p * k + c] = outp[((nn * (CO_PAD / 4) + c / 4) * (size_t)H * W + pin) * 4 + (c % 4)];
    }
  CK(cudaMemcpy(ref.data(), d_ref, total * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0;
  size_t bad = 0, nan = 0, worst = 0;
  for (size_t i = 0; i < total; ++i) {
    if (!(out[i] == out[i])) { ++nan; continue; }
    double e = fabs((double)out[i] - ref[i]);
    if (e > maxerr) { maxerr = e; worst = i; }
    if (fabs(ref[i]) > maxref) maxref = fabs(ref[i]);
    if (e > 1e-3 * (1.0 + fabs(ref[i]))) ++bad;
  }
  printf("  grid=%d units=%d  maxerr=%.3e maxref=%.3e rel=%.3e bad=%zu nan=%zu (worst idx %zu: got %f ref %f)\n",
         grid, a.num_units, maxerr, maxref, maxerr / (maxref + 1e-30), bad, nan, worst, out[worst],
         ref[worst]);
  if (bad || nan) {
    // print a small map of which pixels of frame 0, channel 0 are wrong (first 32x32)
    for (int y = 0; y < (H < 32 ? H : 32); ++y) {
      for (int x = 0; x < (W < 64 ? W : 64); ++x) {
        size_t i = (((size_t)0 * H + y) * W + x) * k;
        double e = fabs((double)out[i] - ref[i]);
        putchar(!(out[i] == out[i]) ? 'N' : (e > 1e-3 * (1.0 + fabs(ref[i])) ? 'x' : '.'));
      }
      putchar('\n');
    }
  }
  if (iters > 0 && !bad && !nan) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 2; ++i) kern<<<grid, 256, Cfg::kSmemBytes>>>(map, a);
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) kern<<<grid, 256, Cfg::kSmemBytes>>>(map, a);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= iters;
    double flops = 2.0 * npix * (double)S * S * k * k;
    double flops_pad = 2.0 * npix * (double)S * S * (KSTEPS * 16) * CO_PAD;
    printf("  time %.3f ms  -> %.1f TFLOP/s algorithmic (%.1f padded)\n", ms, flops / ms * 1e-9,
           flops_pad / ms * 1e-9);
  }
  cudaFree(d_in); cudaFree(d_w); cudaFree(d_bias); cudaFree(d_ref); cudaFree(d_out);
  cudaFree(d_act); cudaFree(d_wpk);
  return (bad || nan) ? 1 : 0;
}

int main(int argc, char** argv) {
  int which = argc > 1 ? atoi(argv[1]) : 0;
  int fails = 0;
  if (which == 0 || which == 1) fails += run_case<1, 4, 64, 4, 1, 4>(2, 16, 32, 64, 0, 0);
  if (which == 0 || which == 2) fails += run_case<3, 4, 64, 4, 9, 4>(2, 32, 32, 64, 0, 0);
  if (which == 0 || which == 3) fails += run_case<15, 4, 64, 4, 5, 4>(2, 64, 64, 64, 0, 0);
  if (which == 0 || which == 4) fails += run_case<15, 4, 64, 4, 5, 4>(3, 64, 64, 64, 5, 0);   // multi-unit per CTA
  if (which == 0 || which == 5) fails += run_case<15, 2, 32, 8, 15, 4>(2, 64, 64, 25, 0, 0);  // k = 25 padded
  if (which == 0 || which == 6) fails += run_case<15, 4, 64, 4, 5, 4>(2, 40, 24, 64, 0, 0);   // ragged H, W
  if (which == 0 || which == 7) fails += run_case<15, 4, 64, 4, 5, 4>(256, 64, 64, 64, 0, 5);  // timing, BASELINE
  if (which == 0 || which == 8) fails += run_case<15, 2, 32, 8, 15, 4>(256, 64, 64, 25, 0, 5);
  printf(fails ? "FAILED %d\n" : "ALL OK\n", fails);
  return fails;
}
