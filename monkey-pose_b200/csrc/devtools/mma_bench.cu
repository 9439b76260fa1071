// Development microbenchmark (GPU box only): issue-rate of tcgen05.mma kind::f16, M=128,
// SS mode, no-swizzle K-major operands, as a function of N -- answers whether the UMMA operand
// fetch from shared memory (A: 4 KB, B: 32*N bytes per instruction) bounds small-N shapes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_bench mma_bench.cu
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#include "../sm100_ptx.cuh"

using namespace sm100;

// mode 0: distinct A per MMA (conv-like, 4 tiles), same B for 4 consecutive MMAs
// mode 1: same A and B every time (best case for any operand caching)
template <int N>
__global__ void __launch_bounds__(128, 1) mma_rate(int reps, int mode, long long* cycles_out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  // zero-fill operands (values do not matter for timing; keep them finite)
  for (int i = threadIdx.x; i < (160 * 1024) / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem_raw + (base - smem_u32(smem_raw)))[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_barrier_init();
  }
  fence_proxy_async();
  if (threadIdx.x < 32) tmem_alloc<512>(smem_u32(&tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = make_idesc(1, 128, N);
    const uint32_t a_buf = base;                 // 96 KB region for A windows
    const uint32_t b_buf = base + 96 * 1024;     // B tiles
    const uint32_t row_pitch = 46 * 16, chunk_pitch = 30 * 46 * 16;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int g = 0; g < 5; ++g) {
        const uint64_t bdesc =
            make_smem_desc(b_buf + (mode ? 0 : g * (2 * N * 16)), N * 16, 128);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const uint32_t a_addr = a_buf + (mode ? 0 : (g * 16 + t * 128));
          const uint64_t adesc = make_smem_desc(a_addr, chunk_pitch, row_pitch);
          mma_bf16_ss(tmem + (t % (512 / N)) * N, adesc, bdesc, idesc, 1);
        }
      }
    }
    tc_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    cycles_out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

template <int N>
void run(int grid, int mode) {
  long long* d;
  cudaMalloc(&d, grid * sizeof(long long));
  auto k = mma_rate<N>;
  const int smem = 161 * 1024 + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int reps = 2000;
  k<<<grid, 128, smem>>>(10, mode, d);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<<<grid, 128, smem>>>(reps, mode, d);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  if (err != cudaSuccess) { printf("N=%d error %s\n", N, cudaGetErrorString(err)); exit(1); }
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long* h = (long long*)malloc(grid * sizeof(long long));
  cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < grid; ++i) if (h[i] > mx) mx = h[i];
  double n_mma = reps * 20.0;
  double cyc = mx / n_mma;
  double ideal = 128.0 * N / 256.0;     // cycles per MMA at 4096 MAC/clk/SM
  double tflops = grid * n_mma * 2.0 * 128 * N * 16 / (ms * 1e-3) * 1e-12;
  printf("N=%3d grid=%3d mode=%d: %.1f cyc/MMA (ideal %.0f, %.0f%% of issue peak)  smem B/cyc=%.0f  %.0f TFLOP/s  (%.3f ms, %.2f GHz eff)\n",
         N, grid, mode, cyc, ideal, 100.0 * ideal / cyc, (4096.0 + 32.0 * N) / cyc, tflops, ms,
         mx / (ms * 1e-3) * 1e-9);
  cudaFree(d); free(h);
}

int main() {
  for (int mode = 0; mode < 2; ++mode)
    for (int grid : {1, 148}) {
      run<32>(grid, mode);
      run<64>(grid, mode);
      run<96>(grid, mode);
      run<128>(grid, mode);
      run<160>(grid, mode);
      run<192>(grid, mode);
      run<256>(grid, mode);
    }
  return 0;
}
