// 1x1 gate convolutions of the hGRU step on tcgen05 (reference: hgru_module.py:696-711, 729-740),
// as a light non-persistent kernel: the op is HBM-bound (one bf16 operand read, one fp32 or bf16
// write per pixel, a k x k contraction in between), so instead of one persistent CTA per SM it runs
// many small CTAs per SM (10-24 KB of shared memory, 32-64 TMEM columns each) and lets occupancy
// hide the load -> MMA -> store latency chain.
//
// CTA = 128 consecutive pixels of one frame (one UMMA M tile).  The chunked operand layout
// [n][cg][pix][8] makes every channel chunk of the tile a contiguous 2 KB run, so CG 1-D bulk
// copies land the tile in shared memory directly in the canonical K-major UMMA layout.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "hconv_tc.cuh"
#include "sm100_ptx.cuh"

namespace hgru {

template <int KP>
struct GateCfg {
  static constexpr int KSTEPS = KP / 16, CG = KP / 8;
  static constexpr int A_BYTES = CG * 128 * 16;
  static constexpr int W_BYTES = KSTEPS * 2 * KP * 16;
  static constexpr int SMEM_BYTES = A_BYTES + W_BYTES + 64 + 128;
  static constexpr uint32_t TMEM_COLS = KP < 32 ? 32 : KP;
};

template <int KP, class Epi>
__global__ void __launch_bounds__(128)
gate_tc_kernel(const __nv_bfloat16* __restrict__ act /*[N][CG][HW][8]*/, const TcConvArgs a) {
  using namespace sm100;
  using Cfg = GateCfg<KP>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
  const uint32_t a_buf = base, w_buf = base + Cfg::A_BYTES;
  const uint32_t bar_ld = w_buf + Cfg::W_BYTES, bar_mma = bar_ld + 8, tmem_slot = bar_ld + 16;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);   // provably warp-uniform
  const int lane = threadIdx.x & 31;
  const int HW = a.H * a.W;
  const int tiles_per_frame = (HW + 127) / 128;
  const int n = blockIdx.x / tiles_per_frame;
  const int p0 = (blockIdx.x - n * tiles_per_frame) * 128;
  const int valid = min(128, HW - p0);

  if (threadIdx.x == 0) {
    mbar_init(bar_ld, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
    mbar_arrive_expect_tx(bar_ld, static_cast<uint32_t>(Cfg::CG * valid * 16 + Cfg::W_BYTES));
    bulk_load(w_buf, a.wpk, Cfg::W_BYTES, bar_ld);
#pragma unroll
    for (int cg = 0; cg < Cfg::CG; ++cg)
      bulk_load(a_buf + cg * 2048, act + ((static_cast<size_t>(n) * Cfg::CG + cg) * HW + p0) * 8,
                static_cast<uint32_t>(valid * 16), bar_ld);
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot_ptr, 0);

  if (warp == 0) {
    const bool leader = elect_one();
    mbar_wait(bar_ld, 0);
    tc_fence_after();
    constexpr uint32_t idesc = make_idesc(1, 128, KP);
    const uint64_t adesc = make_smem_desc(a_buf, 2048, 128);        // LBO: chunk plane, SBO: 8 pixels
    const uint64_t bdesc = make_smem_desc(w_buf, KP * 16, 128);
#pragma unroll
    for (int q = 0; q < Cfg::KSTEPS; ++q)
      if (leader) mma_bf16_ss(tmem_base, adesc + static_cast<uint64_t>((q * 2 * 2048) >> 4),
                              bdesc + static_cast<uint64_t>((q * 2 * KP * 16) >> 4), idesc, q != 0);
    if (leader) tc_commit(bar_mma);
    __syncwarp();
  }
  mbar_wait(bar_mma, 0);
  tc_fence_after();
  float acc[KP];
#pragma unroll
  for (int c0 = 0; c0 < KP; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[c0 + j] = __uint_as_float(v[j]);
  }
  const int m = warp * 32 + lane;
  if (m < valid) {
    const int pin = p0 + m;
    Epi::template apply<KP>(a, n, pin / a.W, pin % a.W, acc);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

}  // namespace hgru
