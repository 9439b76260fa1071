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

__global__ void naive(const __nv_bfloat16* A, const __nv_bfloat16* B, float* C, int M, int N, int K) {
  int n = blockIdx.x * blockDim.x + threadIdx.x, m = blockIdx.y;
  if (n >= N || m >= M) return;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) acc += __bfloat162float(A[(size_t)m * K + k]) * __bfloat162float(B[(size_t)n * K + k]);
  C[(size_t)m * N + n] = acc;
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
  std::vector<__nv_bfloat16> A((size_t)M * K), B((size_t)N * K);
  srand(7);
  for (auto& v : A) v = __float2bfloat16(rand() / (float)RAND_MAX - 0.5f);
  for (auto& v : B) v = __float2bfloat16(rand() / (float)RAND_MAX - 0.5f);
  __nv_bfloat16 *dA, *dB; float *dC, *dR, *dP;
  int total_kb = (K + 63) / 64;
  int tiles = ((N + 255) / 256) * ((M + 127) / 128);
  int splits = 148 / tiles; if (splits < 1) splits = 1; if (splits > total_kb) splits = total_kb;
  int kbps = (total_kb + splits - 1) / splits; splits = (total_kb + kbps - 1) / kbps;
  CK(cudaMalloc(&dA, A.size() * 2)); CK(cudaMalloc(&dB, B.size() * 2));
  CK(cudaMalloc(&dC, (size_t)M * N * 4)); CK(cudaMalloc(&dR, (size_t)M * N * 4)); CK(cudaMalloc(&dP, (size_t)splits * M * N * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice));
  CUtensorMap ma, mb;
  if (hgru::make_kmajor_bf16_map(&ma, dA, M, K, 128) || hgru::make_kmajor_bf16_map(&mb, dB, N, K, 256)) { printf("map fail\n"); return 1; }
  hgru::GemmArgs g{M, N, K, kbps, dP};
  CK(cudaFuncSetAttribute(hgru::gemm_tc_splitk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, hgru::kGemmSmemBytes));
  dim3 grid((N + 255) / 256, (M + 127) / 128, splits);
  hgru::gemm_tc_splitk_kernel<<<grid, 256, hgru::kGemmSmemBytes>>>(ma, mb, g);
  CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
  reduce<<<(unsigned)(((size_t)M * N + 255) / 256), 256>>>(dP, dC, splits, (size_t)M * N);
  naive<<<dim3((N + 127) / 128, M), 128>>>(dA, dB, dR, M, N, K);
  CK(cudaDeviceSynchronize());
  std::vector<float> C((size_t)M * N), R((size_t)M * N);
  CK(cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(R.data(), dR, R.size() * 4, cudaMemcpyDeviceToHost));
  double me = 0, mr = 0; size_t bad = 0;
  for (size_t i = 0; i < C.size(); ++i) { double e = fabs((double)C[i] - R[i]); if (e > me) me = e; if (fabs(R[i]) > mr) mr = fabs(R[i]); if (!(e <= 1e-3 * (1 + fabs(R[i])))) ++bad; }
  printf("  splits=%d kbps=%d maxerr=%.3e maxref=%.3e bad=%zu\n", splits, kbps, me, mr, bad);
  if (!bad && iters) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) hgru::gemm_tc_splitk_kernel<<<grid, 256, hgru::kGemmSmemBytes>>>(ma, mb, g);
    cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= iters;
    printf("  %.3f ms  %.1f TFLOP/s  B-stream %.0f GB/s\n", ms, 2.0 * M * N * K / ms * 1e-9, (double)N * K * 2 / ms * 1e-6);
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dR); cudaFree(dP);
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
