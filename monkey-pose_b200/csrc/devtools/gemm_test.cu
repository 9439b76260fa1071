// Development harness (GPU box only): split-K tcgen05 GEMM vs a naive fp32 GEMM on bf16-rounded data.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../gemm_tc.cuh"
#include "../tc_host.cuh"
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(2);} } while (0)

__global__ void naive(const float* A, const float* B, float* C, int M, int N, int K) {
  int n = blockIdx.x * blockDim.x + threadIdx.x, m = blockIdx.y;
  if (n >= N || m >= M) return;
  double acc = 0.0;
  for (int k = 0; k < K; ++k) acc += (double)A[(size_t)m * K + k] * (double)B[(size_t)n * K + k];
  C[(size_t)m * N + n] = (float)acc;
}
__global__ void pack_hilo(const float* X, __nv_bfloat16* out, int rows, int K, int Kpad) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= (size_t)rows * K) return;
  int k = i % K; size_t r = i / K;
  __nv_bfloat16 hi = __float2bfloat16(X[i]);
  out[r * 2 * Kpad + k] = hi;
  out[r * 2 * Kpad + Kpad + k] = __float2bfloat16(X[i] - __bfloat162float(hi));
}
__global__ void reduce(const float* part, float* C, int splits, size_t mn) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= mn) return;
  float a = 0.f;
  for (int z = 0; z < splits; ++z) a += part[z * mn + i];
  C[i] = a;
}
int run(int M, int N, int K, int iters) {
  printf("gemm M=%d N=%d K=%d\n", M, N, K);
  std::vector<float> A((size_t)M * K), B((size_t)N * K);
  srand(7);
  for (auto& v : A) v = rand() / (float)RAND_MAX - 0.5f;
  for (auto& v : B) v = rand() / (float)RAND_MAX - 0.5f;
  __nv_bfloat16 *dA, *dB; float *dAf, *dBf, *dC, *dR, *dP;
  int total_kb = (K + 63) / 64;
  int Kpad = total_kb * 64;
  int tiles = ((N + 255) / 256) * ((M + 127) / 128);
  int splits = 148 / tiles; if (splits < 1) splits = 1; if (splits > total_kb) splits = total_kb;
  int kbps = (total_kb + splits - 1) / splits; splits = (total_kb + kbps - 1) / kbps;
  CK(cudaMalloc(&dA, (size_t)M * 2 * Kpad * 2)); CK(cudaMalloc(&dB, (size_t)N * 2 * Kpad * 2));
  CK(cudaMalloc(&dAf, A.size() * 4)); CK(cudaMalloc(&dBf, B.size() * 4));
  CK(cudaMemset(dA, 0, (size_t)M * 2 * Kpad * 2)); CK(cudaMemset(dB, 0, (size_t)N * 2 * Kpad * 2));
  CK(cudaMalloc(&dC, (size_t)M * N * 4)); CK(cudaMalloc(&dR, (size_t)M * N * 4)); CK(cudaMalloc(&dP, (size_t)splits * M * N * 4));
  CK(cudaMemcpy(dAf, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dBf, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  pack_hilo<<<(unsigned)((A.size() + 255) / 256), 256>>>(dAf, dA, M, K, Kpad);
  pack_hilo<<<(unsigned)((B.size() + 255) / 256), 256>>>(dBf, dB, N, K, Kpad);
  CUtensorMap ma, mb;
  if (hgru::make_kmajor_bf16_map(&ma, dA, M, 2 * (size_t)Kpad, 128) || hgru::make_kmajor_bf16_map(&mb, dB, N, 2 * (size_t)Kpad, 256)) { printf("map fail\n"); return 1; }
  hgru::GemmArgs g{M, N, K, Kpad, kbps, dP};
  CK(cudaFuncSetAttribute(hgru::gemm_tc_splitk_kernel<hgru::kGemmBN>, cudaFuncAttributeMaxDynamicSharedMemorySize, hgru::kGemmSmemBytes));
  dim3 grid((N + 255) / 256, (M + 127) / 128, splits);
  hgru::gemm_tc_splitk_kernel<hgru::kGemmBN><<<grid, 256, hgru::kGemmSmemBytes>>>(ma, ma, mb, g);
  CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
  reduce<<<(unsigned)(((size_t)M * N + 255) / 256), 256>>>(dP, dC, splits, (size_t)M * N);
  naive<<<dim3((N + 127) / 128, M), 128>>>(dAf, dBf, dR, M, N, K);
  CK(cudaDeviceSynchronize());
  std::vector<float> C((size_t)M * N), R((size_t)M * N);
  CK(cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(R.data(), dR, R.size() * 4, cudaMemcpyDeviceToHost));
  double me = 0, mr = 0; size_t bad = 0;
  for (size_t i = 0; i < C.size(); ++i) { double e = fabs((double)C[i] - R[i]); if (e > me) me = e; if (fabs(R[i]) > mr) mr = fabs(R[i]); if (!(e <= 2e-4 * (1 + fabs(R[i])))) ++bad; }
  printf("  splits=%d kbps=%d maxerr=%.3e maxref=%.3e bad=%zu\n", splits, kbps, me, mr, bad);
  if (!bad && iters) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) hgru::gemm_tc_splitk_kernel<hgru::kGemmBN><<<grid, 256, hgru::kGemmSmemBytes>>>(ma, ma, mb, g);
    cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= iters;
    printf("  %.3f ms  %.1f TFLOP/s  B-stream %.0f GB/s\n", ms, 2.0 * M * N * K / ms * 1e-9, (double)N * K * 4 / ms * 1e-6);
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dAf); cudaFree(dBf); cudaFree(dC); cudaFree(dR); cudaFree(dP);
  return bad ? 1 : 0;
}
int main() {
  int f = 0;
  f += run(128, 256, 64, 0);
  f += run(128, 256, 512, 0);
  f += run(2, 32, 512, 0);
  f += run(3, 69, 216, 0);
  f += run(256, 1024, 4096, 0);
  f += run(256, 1024, 102400, 5);
  f += run(256, 1024, 262144, 5);
  printf(f ? "FAILED\n" : "ALL OK\n");
  return f;
}
