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
__device__ int g_random_fill = 0;

template <int N, int mode>
__global__ void __launch_bounds__(128, 1) mma_rate(int reps, long long* cycles_out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar, bar2, bar3;
  // zero-fill operands (values do not matter for timing; keep them finite)
  for (int i = threadIdx.x; i < (208 * 1024) / 16; i += blockDim.x) {
    uint4 v = make_uint4(0, 0, 0, 0);
    if (g_random_fill) {
      // pseudo-random bf16 values in (-2, 2): data-dependent switching power, as real activations have
      uint32_t h = (i + 1) * 2654435761u + blockIdx.x * 40503u;
      uint32_t w[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        h ^= h << 13; h ^= h >> 17; h ^= h << 5;
        const uint32_t lo = 0x3F00u | (h & 0x80FFu), hi = 0x3F00u | ((h >> 16) & 0x80FFu);
        w[j] = lo | (hi << 16);
      }
      v = make_uint4(w[0], w[1], w[2], w[3]);
    }
    reinterpret_cast<uint4*>(smem_raw + (base - smem_u32(smem_raw)))[i] = v;
  }
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    mbar_init(smem_u32(&bar2), 1);
    mbar_init(smem_u32(&bar3), 1);
    fence_barrier_init();
  }
  fence_proxy_async();
  if (threadIdx.x < 32) tmem_alloc<512>(smem_u32(&tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (mode >= 10) {
    // lean issuer as in the production kernels: a convergent warp, loop state in uniform registers, one
    // elected lane issues.  bits of (mode - 10): 1 = commit per stage, 2 = blocking wait per stage on a
    // completed barrier, 4 = non-blocking probe per stage, 8 = commit only every 2nd stage
    const int wrp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
    if (wrp == 0) {
      constexpr int V = mode - 10;
      const bool leader = elect_one();
      const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
      constexpr uint32_t idesc = make_idesc(1, 128, N);
      constexpr uint32_t row_pitch = 82 * 16, chunk_pitch = 30 * 82 * 16;
      const uint64_t adesc0 = make_smem_desc(base, chunk_pitch, row_pitch);
      const uint64_t bdesc0 = make_smem_desc(base + 158 * 1024, N * 16, 128);
      long long t0 = clock64();
      uint32_t st = 0;
      for (int r = 0; r < reps; r += 30) {
        for (int dy = 0; dy < 15; ++dy) {
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            if (V & 2) mbar_wait_warp(smem_u32(&bar3), 1);
            if (V & 4) (void)mbar_test(smem_u32(&bar3), 1);
            const uint64_t a_st = adesc0 + ((q * 2 * chunk_pitch + dy * row_pitch) >> 4);
            const uint64_t b_st = bdesc0 + ((st * 3 * (2 * N * 16)) >> 4);
            if (leader) {
#pragma unroll
              for (int g = 0; g < 3; ++g) {
                mma_bf16_ss(tm, a_st + 5 * g, b_st + ((g * 2 * N * 16) >> 4), idesc, 1);
                mma_bf16_ss(tm + N, a_st + 5 * g + 8, b_st + ((g * 2 * N * 16) >> 4), idesc, 1);
              }
              if ((V & 1) || ((V & 8) && q)) tc_commit(smem_u32(&bar2));
            }
            st = (st + 1) & 3;
          }
        }
      }
      if (leader) tc_commit(smem_u32(&bar));
      mbar_wait_warp(smem_u32(&bar), 0);
      long long t1 = clock64();
      if (leader) cycles_out[blockIdx.x] = t1 - t0;
    }
  } else
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = make_idesc(1, 128, N);
    const uint32_t a_buf = base;                 // 96 KB region for A windows
    const uint32_t b_buf = base + (mode >= 2 ? 158 * 1024 : 96 * 1024);     // B tiles
    uint32_t row_pitch = 46 * 16, chunk_pitch = 30 * 46 * 16;
    if (mode >= 2) { row_pitch = 82 * 16; chunk_pitch = 30 * 82 * 16; }   // the stacked kernel's window
    long long t0 = clock64();
    if (mode >= 2) {
      // the stacked kernel's exact operand pattern: per stage (dy, q) 3 tap groups x 2 tiles, 12 B blocks
      for (int r = 0; r < reps; ++r) {
        const int sg = r % 30, dy = sg >> 1, q = sg & 1;
        const uint32_t a_st = a_buf + q * 2 * chunk_pitch + dy * row_pitch;
        const uint32_t b_st = b_buf + (r & 3) * 3 * (2 * N * 16);
#pragma unroll
        for (int g = 0; g < 3; ++g) {
          const uint64_t bdesc = make_smem_desc(b_st + g * (2 * N * 16), N * 16, 128);
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const uint64_t adesc = make_smem_desc(a_st + (8 * t + 5 * g) * 16, chunk_pitch, row_pitch);
            mma_bf16_ss(tmem + t * N, adesc, bdesc, idesc, 1);
          }
        }
        // mode 3: one commit per stage (as the weight-ring release does); mode 4: commit every 2nd stage;
        // mode 5: commit per stage + a completed-barrier probe
        // mode 6: blocking wait on an already-complete barrier per stage; mode 7: commit + blocking wait
        if (mode == 3 || mode == 5 || mode == 7 || (mode == 4 && (r & 1))) tc_commit(smem_u32(&bar2));
        if (mode == 5) (void)mbar_test(smem_u32(&bar3), 1);
        if (mode == 6 || mode == 7) mbar_wait(smem_u32(&bar3), 1);
      }
    } else
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

template <int N, int mode>
void run(int grid) {
  long long* d;
  cudaMalloc(&d, grid * sizeof(long long));
  auto k = mma_rate<N, mode>;
  const int smem = 209 * 1024 + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int reps = 2100;
  k<<<grid, 128, smem>>>(10, d);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<<<grid, 128, smem>>>(reps, d);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  if (err != cudaSuccess) { printf("N=%d error %s\n", N, cudaGetErrorString(err)); exit(1); }
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long* h = (long long*)malloc(grid * sizeof(long long));
  cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < grid; ++i) if (h[i] > mx) mx = h[i];
  double n_mma = reps * (mode >= 2 ? 6.0 : 20.0);
  double cyc = mx / n_mma;
  double ideal = 128.0 * N / 256.0;     // cycles per MMA at 4096 MAC/clk/SM
  double tflops = grid * n_mma * 2.0 * 128 * N * 16 / (ms * 1e-3) * 1e-12;
  printf("N=%3d grid=%3d mode=%d: %.1f cyc/MMA (ideal %.0f, %.0f%% of issue peak)  smem B/cyc=%.0f  %.0f TFLOP/s  (%.3f ms, %.2f GHz eff)\n",
         N, grid, mode, cyc, ideal, 100.0 * ideal / cyc, (4096.0 + 32.0 * N) / cyc, tflops, ms,
         mx / (ms * 1e-3) * 1e-9);
  cudaFree(d); free(h);
}

int main() {
  for (int rnd = 0; rnd < 2; ++rnd) {
    cudaMemcpyToSymbol(g_random_fill, &rnd, sizeof(int));
    printf("operand fill: %s\n", rnd ? "pseudo-random bf16" : "zeros");
    run<128, 2>(1); run<128, 2>(148); run<128, 10>(148); run<128, 13>(148);
    run<128, 0>(148); run<256, 0>(148); run<64, 0>(148);
  }
  return 0;
}
