// fp32 SIMT kernels for the hGRU-pose forward: the exact (<= 1e-4) path and the glue around the
// tensor-core convolutions.  All activations are NHWC fp32 (the reference's layout), weights HWIO.
// Reference call sites are cited per kernel (paths relative to /root/reference).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace hgru {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }
// tanh with ~1e-7 relative error (tanhf from libdevice; NOT tanh.approx, the 1e-4 path needs it)
__device__ __forceinline__ float tanhf_(float x) { return tanhf(x); }

// Writes one bf16 value into the "chunked NHWC" operand copy [n][cg][y][x][8] (see hconv_tc.cuh).
__device__ __forceinline__ size_t chunked_index(size_t pix_in_frame, int n, int c, int HW, int CG) {
  return ((static_cast<size_t>(n) * CG + (c >> 3)) * HW + pix_in_frame) * 8 + (c & 7);
}

// ------------------------------------------------------------------------------------------------
// Generic direct convolution, stride 1, SAME zero padding, fp32 (tf.nn.conv2d: hgru_module.py:531-548,
// hgru_pose.py:146), with the epilogue  y = act(conv + bias) * scale + shift  (bias_add + relu of
// hgru_pose.py:147-148 and the inference-mode batch_normalization that follows, :52-80).
// Block = 256 threads = 32 pixel-groups (8 px along x) x 8 channel-groups (8 co): a 16x16 pixel tile
// times 64 output channels; input patch and one filter row are staged in shared memory.
// ------------------------------------------------------------------------------------------------
template <int S>
struct ConvSimtCfg {
  static constexpr int PH = 16 + S - 1;
  static constexpr int PWraw = 16 + S - 1;
  static constexpr int PW = ((8 + ((8 + S - 1 + 3) / 4) * 4) + 3) / 4 * 4;   // room for float4 window loads
  static constexpr int NWIN = ((8 + S - 1 + 3) / 4) * 4;                       // window floats per thread
  static constexpr int kPatchFloats = 8 * PH * PW;
  static constexpr int kWFloats = S * 8 * 64;
  static constexpr int kSmemBytes = (kPatchFloats + kWFloats) * 4;
};

template <int S>
__global__ void __launch_bounds__(256)
conv_simt_kernel(const float* __restrict__ in, const float* __restrict__ w,
                 const float* __restrict__ bias, const float* __restrict__ scale,
                 const float* __restrict__ shift, float* __restrict__ out, int N, int H, int W,
                 int Ci, int Co, int relu) {
  using Cfg = ConvSimtCfg<S>;
  extern __shared__ float smem_f[];
  float* patch = smem_f;                       // [8 ci][PH][PW]
  float* wsm = smem_f + Cfg::kPatchFloats;     // [S dx][8 ci][64 co]
  constexpr int PAD = (S - 1) / 2;
  const int tiles_x = (W + 15) / 16;
  const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
  const int co0 = blockIdx.y * 64;
  const int n = blockIdx.z;
  const int tid = threadIdx.x;
  const int cg = tid & 7;            // channel group -> co0 + cg*8 .. +8
  const int pg = tid >> 3;           // pixel group 0..31
  const int prow = pg >> 1;          // 0..15
  const int pxg = (pg & 1) * 8;      // 0 or 8
  const int y_base = ty * 16 - PAD, x_base = tx * 16 - PAD;

  float acc[8][8];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[j][c] = 0.f;

  for (int ci0 = 0; ci0 < Ci; ci0 += 8) {
    __syncthreads();
    // stage the input patch (zero fill outside the image / beyond Ci)
    for (int e = tid; e < Cfg::PH * Cfg::PWraw * 8; e += 256) {
      const int ci = e & 7;
      const int p = e >> 3;
      const int px = p % Cfg::PWraw, py = p / Cfg::PWraw;
      const int y = y_base + py, x = x_base + px;
      float v = 0.f;
      if (y >= 0 && y < H && x >= 0 && x < W && ci0 + ci < Ci)
        v = in[((static_cast<size_t>(n) * H + y) * W + x) * Ci + ci0 + ci];
      patch[(ci * Cfg::PH + py) * Cfg::PW + px] = v;
    }
    for (int dy = 0; dy < S; ++dy) {
      __syncthreads();
      for (int e = tid; e < S * 8 * 64; e += 256) {
        const int co = e & 63;
        const int ci = (e >> 6) & 7;
        const int dx = e >> 9;
        float v = 0.f;
        if (ci0 + ci < Ci && co0 + co < Co)
          v = w[((static_cast<size_t>(dy) * S + dx) * Ci + ci0 + ci) * Co + co0 + co];
        wsm[e] = v;
      }
      __syncthreads();
#pragma unroll 1
      for (int ci = 0; ci < 8; ++ci) {
        float win[Cfg::NWIN];
        const float4* src =
            reinterpret_cast<const float4*>(patch + (ci * Cfg::PH + prow + dy) * Cfg::PW + pxg);
#pragma unroll
        for (int j = 0; j < Cfg::NWIN / 4; ++j) {
          const float4 v = src[j];
          win[4 * j] = v.x; win[4 * j + 1] = v.y; win[4 * j + 2] = v.z; win[4 * j + 3] = v.w;
        }
#pragma unroll
        for (int dx = 0; dx < S; ++dx) {
          const float4* wp = reinterpret_cast<const float4*>(wsm + (dx * 8 + ci) * 64 + cg * 8);
          const float4 w0 = wp[0], w1 = wp[1];
          const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
          for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[j][c] = fmaf(win[j + dx], wv[c], acc[j][c]);
        }
      }
    }
  }
  const int y = ty * 16 + prow;
  if (y >= H) return;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int x = tx * 16 + pxg + j;
    if (x >= W) continue;
    float* dst = out + ((static_cast<size_t>(n) * H + y) * W + x) * Co;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int co = co0 + cg * 8 + c;
      if (co < Co) {
        float v = acc[j][c] + (bias ? bias[co] : 0.f);
        if (relu) v = fmaxf(v, 0.f);
        if (scale) v = v * scale[co] + shift[co];
        dst[co] = v;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Stem head: conv 3x3 (1 -> C) + bias + ReLU (hgru_pose.py:50,146-148), 2x2/2 max-pool (:51,134-137)
// and the inference batch-norm affine (:52-60), fused: [N,2H,2W,1] -> [N,H,W,C].  Bandwidth-bound
// stencil; one thread per (pooled pixel, channel); optionally also emits the bf16 chunked copy.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
stem_conv1_pool_bn_kernel(const float* __restrict__ depth, const float* __restrict__ w /*[3][3][1][C]*/,
                          const float* __restrict__ bias, const float* __restrict__ scale,
                          const float* __restrict__ shift, float* __restrict__ out,
                          __nv_bfloat16* __restrict__ out_bf16, int N, int H, int W, int C, int CG) {
  const size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  const size_t total = static_cast<size_t>(N) * H * W * C;
  if (idx >= total) return;
  const int c = idx % C;
  const size_t p = idx / C;
  const int x = p % W;
  const int y = (p / W) % H;
  const int n = p / (static_cast<size_t>(W) * H);
  const int IH = 2 * H, IW = 2 * W;
  float wv[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) wv[t] = w[t * C + c];
  // 4x4 input neighbourhood of the 2x2 pooling window
  float v[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int yy = 2 * y - 1 + r, xx = 2 * x - 1 + q;
      v[r][q] = (yy >= 0 && yy < IH && xx >= 0 && xx < IW)
                    ? depth[(static_cast<size_t>(n) * IH + yy) * IW + xx] : 0.f;
    }
  const float b = bias[c];
  float m = -INFINITY;
#pragma unroll
  for (int py = 0; py < 2; ++py)
#pragma unroll
    for (int px = 0; px < 2; ++px) {
      float a = 0.f;
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) a = fmaf(v[py + dy][px + dx], wv[dy * 3 + dx], a);
      m = fmaxf(m, fmaxf(a + b, 0.f));
    }
  const float r = m * scale[c] + shift[c];
  out[idx] = r;
  if (out_bf16) out_bf16[chunked_index(static_cast<size_t>(y) * W + x, n, c, H * W, CG)] = __float2bfloat16(r);
}

// ------------------------------------------------------------------------------------------------
// 1x1 gate convolution + sigmoid (hgru_module.py:696-707 and :729-740):
//   G = sigmoid(in *1x1 wg + bg);  optional gated copy  out_mul = in . G  (:709-711).
// Block = 64 pixels x k channels; wg [k][k] in shared memory; one thread = one pixel x 8 outputs.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gate1x1_kernel(const float* __restrict__ in, const float* __restrict__ wg /*[ci][co]*/,
               const float* __restrict__ bg, float* __restrict__ out_g, float* __restrict__ out_mul,
               __nv_bfloat16* __restrict__ out_mul_bf16, size_t npix, int k, int HW, int CG) {
  extern __shared__ float smem_f[];
  float* wsm = smem_f;                 // [k][k]
  float* xin = smem_f + k * k;         // [64][k+1]
  const int tid = threadIdx.x;
  const size_t p0 = blockIdx.x * static_cast<size_t>(64);
  for (int e = tid; e < k * k; e += 256) wsm[e] = wg[e];
  for (int e = tid; e < 64 * k; e += 256) {
    const int pp = e / k, c = e - pp * k;
    xin[pp * (k + 1) + c] = (p0 + pp < npix) ? in[(p0 + pp) * k + c] : 0.f;
  }
  __syncthreads();
  const int pp = tid & 63;             // pixel within block
  const int og = tid >> 6;             // 0..3: output channels og, og+4, ...
  const size_t p = p0 + pp;
  if (p >= npix) return;
  const int n = p / HW;
  const size_t pin = p - static_cast<size_t>(n) * HW;
  for (int co = og; co < k; co += 4) {
    float a = 0.f;
    for (int ci = 0; ci < k; ++ci) a = fmaf(xin[pp * (k + 1) + ci], wsm[ci * k + co], a);
    const float g = sigmoidf_(a + bg[co]);
    if (out_g) out_g[p * k + co] = g;
    const float mv = xin[pp * (k + 1) + co] * g;
    if (out_mul) out_mul[p * k + co] = mv;
    if (out_mul_bf16) out_mul_bf16[chunked_index(pin, n, co, HW, CG)] = __float2bfloat16(mv);
  }
}

// ------------------------------------------------------------------------------------------------
// input_integration (hgru_module.py:795-804):  H1 = tanh(X - (beta*H2 + nu) * C1),  C1 already
// includes lateral_bias (:657).  Optionally emits the bf16 chunked operand copy of H1.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
h1_kernel(const float* __restrict__ X, const float* __restrict__ H2, const float* __restrict__ C1,
          const float* __restrict__ beta, const float* __restrict__ nu, float* __restrict__ H1,
          __nv_bfloat16* __restrict__ H1_bf16, size_t total, int k, int HW, int CG) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int c = i % k;
  const float h1 = tanhf_(X[i] - (beta[c] * H2[i] + nu[c]) * C1[i]);
  H1[i] = h1;
  if (H1_bf16) {
    const size_t p = i / k;
    const int n = p / HW;
    H1_bf16[chunked_index(p - static_cast<size_t>(n) * HW, n, c, HW, CG)] = __float2bfloat16(h1);
  }
}

// ------------------------------------------------------------------------------------------------
// output_integration + adaptation (hgru_module.py:806-823, 847-849):
//   e = gamma*C2; Ht = tanh(kappa*(H1+e) + omega*(H1*e)); H2 = (G2*H2 + (1-G2)*Ht) * rho_t
// In place on H2; optional trace copies; optional bf16 chunked copy of the new H2.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
h2_kernel(const float* __restrict__ H1, const float* __restrict__ C2, const float* __restrict__ G2,
          const float* __restrict__ gamma, const float* __restrict__ kappa,
          const float* __restrict__ omega, const float* __restrict__ rho, int t,
          float* __restrict__ H2, __nv_bfloat16* __restrict__ H2_bf16, size_t total, int k, int HW,
          int CG) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int c = i % k;
  const float h1 = H1[i];
  const float e = gamma[c] * C2[i];
  const float ht = tanhf_(kappa[c] * (h1 + e) + omega[c] * (h1 * e));
  const float g = G2[i];
  const float h2 = (g * H2[i] + (1.f - g) * ht) * rho[t];
  H2[i] = h2;
  if (H2_bf16) {
    const size_t p = i / k;
    const int n = p / HW;
    H2_bf16[chunked_index(p - static_cast<size_t>(n) * HW, n, c, HW, CG)] = __float2bfloat16(h2);
  }
}

// fp32 NHWC -> bf16 chunked operand copy, with optional per-channel affine
__global__ void __launch_bounds__(256)
to_chunked_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, size_t total,
                       int k, int HW, int CG) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int c = i % k;
  const size_t p = i / k;
  const int n = p / HW;
  out[chunked_index(p - static_cast<size_t>(n) * HW, n, c, HW, CG)] = __float2bfloat16(in[i]);
}

// ------------------------------------------------------------------------------------------------
// Readout FC1 (hgru_pose.py:91,156-163): part[z][m][j] = sum_{kk in slice z} (a[m][kk]*sc[kk%k]+sh[kk%k]) * Wt[kk][j]
// The per-channel affine on A is the inference batch-norm of the hGRU output (:82-90).
// Split-K SGEMM: block tile 64(m) x 64(j), 256 threads, 4x4 per thread, K step 16.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
fc1_splitk_kernel(const float* __restrict__ A, const float* __restrict__ Wt,
                  const float* __restrict__ sc, const float* __restrict__ sh,
                  float* __restrict__ part, int M, int K, int Nout, int kch, int kslice) {
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  const int tid = threadIdx.x;
  const int j0 = blockIdx.x * 64, m0 = blockIdx.y * 64, z = blockIdx.z;
  const int kb = z * kslice, ke = min(K, kb + kslice);
  const int tm = (tid >> 4) * 4, tj = (tid & 15) * 4;
  float acc[4][4] = {};
  for (int k0 = kb; k0 < ke; k0 += 16) {
    // A tile: 64 rows x 16 k  (row-major A[m][K]); each thread loads 4 elements
    {
      const int r = tid >> 2, kk = (tid & 3) * 4;
      const int m = m0 + r;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int kidx = k0 + kk + q;
        float v = 0.f;
        if (m < M && kidx < ke) {
          const int c = kidx % kch;
          v = A[static_cast<size_t>(m) * K + kidx] * sc[c] + sh[c];
        }
        As[kk + q][r] = v;
      }
    }
    {
      const int kk = tid >> 4, jj = (tid & 15) * 4;
      const int kidx = k0 + kk;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int j = j0 + jj + q;
        Bs[kk][jj + q] = (kidx < ke && j < Nout) ? Wt[static_cast<size_t>(kidx) * Nout + j] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) { a[q] = As[kk][tm + q]; b[q] = Bs[kk][tj + q]; }
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[q][r] = fmaf(a[q], b[r], acc[q][r]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int m = m0 + tm + q;
    if (m >= M) continue;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int j = j0 + tj + r;
      if (j < Nout) part[(static_cast<size_t>(z) * M + m) * Nout + j] = acc[q][r];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Readout tail: fc1 = sum_z part + b1 (hgru_pose.py:91); relu (:92); batch-norm affine (:95-103,
// resolution R-D5); fc_out (:104, R-D6): out[m][o] = sum_j r[j] * W2[j][o] + b2[o].
// One block per frame.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
fc_tail_kernel(const float* __restrict__ part, int nsplit, const float* __restrict__ b1,
               const float* __restrict__ sc, const float* __restrict__ sh,
               const float* __restrict__ W2, const float* __restrict__ b2, float* __restrict__ fc1_out,
               float* __restrict__ out, int M, int Hid, int Nout) {
  extern __shared__ float r[];      // [Hid]
  const int m = blockIdx.x;
  for (int j = threadIdx.x; j < Hid; j += blockDim.x) {
    float a = 0.f;
    for (int z = 0; z < nsplit; ++z) a += part[(static_cast<size_t>(z) * M + m) * Hid + j];
    a += b1[j];
    if (fc1_out) fc1_out[static_cast<size_t>(m) * Hid + j] = a;
    r[j] = fmaxf(a, 0.f) * sc[j] + sh[j];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int o = warp; o < Nout; o += nw) {
    float a = 0.f;
    for (int j = lane; j < Hid; j += 32) a = fmaf(r[j], W2[static_cast<size_t>(j) * Nout + o], a);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) a += __shfl_xor_sync(0xffffffffu, a, s);
    if (lane == 0) out[static_cast<size_t>(m) * Nout + o] = a + b2[o];
  }
}

// batch-norm inference fold: scale = gamma / sqrt(var + eps), shift = beta - mean * scale
__global__ void bn_fold_kernel(const float* gamma, const float* beta, const float* mean,
                               const float* var, float eps, float* scale, float* shift, int c) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c) return;
  const float s = gamma[i] / sqrtf(var[i] + eps);
  scale[i] = s;
  shift[i] = beta[i] - mean[i] * s;
}

// Weight packing for the tcgen05 conv: HWIO fp32 [S][S][k][k] -> bf16 [KSTEPS][taps][2][CO_PAD][8 ci]
__global__ void pack_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wpk,
                                    int taps, int k, int ksteps, int co_pad) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  const size_t total = static_cast<size_t>(ksteps) * taps * 2 * co_pad * 8;
  if (i >= total) return;
  const int j = i & 7;
  size_t r = i >> 3;
  const int co = r % co_pad; r /= co_pad;
  const int ch = r & 1; r >>= 1;
  const int tap = r % taps;
  const int q = r / taps;
  const int ci = q * 16 + ch * 8 + j;
  float v = 0.f;
  if (ci < k && co < k) v = w[(static_cast<size_t>(tap) * k + ci) * k + co];
  wpk[i] = __float2bfloat16(v);
}

}  // namespace hgru
