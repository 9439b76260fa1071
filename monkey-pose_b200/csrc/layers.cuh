// Stand-alone layer kernels behind the reference model's layer METHODS (hgru_pose.py:134-163: max_pool,
// conv_layer, fc_layer) and the train-mode batch normalisation (tf.layers.batch_normalization(training=True),
// hgru_pose.py:52-103).  model.build() does not go through these -- it runs the fused tensor-core pipeline; these
// serve callers that compose the layers themselves, exactly (fp32, any channel counts), without the fusion.
// All tensors dense fp32, activations NHWC, conv weights HWIO.
#pragma once
#include <cuda_runtime.h>

namespace hgru {

// relu?(conv2d(x, w, stride 1, SAME) + b): one thread per (pixel, output channel); consecutive threads take
// consecutive output channels, so weight reads are coalesced and the input value is a warp broadcast.
__global__ void __launch_bounds__(256)
conv2d_direct_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                     float* __restrict__ out, int N, int H, int W, int Ci, int Co, int S, int relu) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  const size_t total = static_cast<size_t>(N) * H * W * Co;
  if (i >= total) return;
  const int co = static_cast<int>(i % Co);
  const size_t pix = i / Co;
  const int xx = static_cast<int>(pix % W);
  const int yy = static_cast<int>((pix / W) % H);
  const int n = static_cast<int>(pix / (static_cast<size_t>(W) * H));
  const int pad = (S - 1) / 2;          // TF SAME, stride 1, odd S
  float acc = 0.f;
  for (int dy = 0; dy < S; ++dy) {
    const int y = yy + dy - pad;
    if (y < 0 || y >= H) continue;
    for (int dx = 0; dx < S; ++dx) {
      const int xq = xx + dx - pad;
      if (xq < 0 || xq >= W) continue;
      const float* xp = x + ((static_cast<size_t>(n) * H + y) * W + xq) * Ci;
      const float* wp = w + (static_cast<size_t>(dy) * S + dx) * Ci * Co + co;
      for (int ci = 0; ci < Ci; ++ci) acc = fmaf(__ldg(xp + ci), __ldg(wp + static_cast<size_t>(ci) * Co), acc);
    }
  }
  acc += __ldg(b + co);
  out[i] = relu ? fmaxf(acc, 0.f) : acc;
}

// tf.nn.max_pool(ksize 2x2, stride 2, SAME): output ceil(H/2) x ceil(W/2); windows are clipped at the border
__global__ void __launch_bounds__(256)
max_pool2x2_kernel(const float* __restrict__ x, float* __restrict__ out, int N, int H, int W, int C) {
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  const size_t total = static_cast<size_t>(N) * Ho * Wo * C;
  if (i >= total) return;
  const int c = static_cast<int>(i % C);
  const size_t pix = i / C;
  const int xo = static_cast<int>(pix % Wo);
  const int yo = static_cast<int>((pix / Wo) % Ho);
  const int n = static_cast<int>(pix / (static_cast<size_t>(Wo) * Ho));
  float m = -INFINITY;
  for (int dy = 0; dy < 2; ++dy)
    for (int dx = 0; dx < 2; ++dx) {
      const int y = 2 * yo + dy, xq = 2 * xo + dx;
      if (y < H && xq < W) m = fmaxf(m, __ldg(x + ((static_cast<size_t>(n) * H + y) * W + xq) * C + c));
    }
  out[i] = m;
}

// tf.nn.max_pool / tf.nn.avg_pool with ksize = stride = k, SAME (the model's max_pool_4 and avg_pool helpers,
// hgru_pose.py:124-132): output ceil(H/k) x ceil(W/k); TF pads pad_total = out*k - H, pad_total / 2 of it before;
// padding never wins a max and is left out of an average's count.
__global__ void __launch_bounds__(256)
pool_same_kernel(const float* __restrict__ x, float* __restrict__ out, int N, int H, int W, int C, int k, int avg) {
  const int Ho = (H + k - 1) / k, Wo = (W + k - 1) / k;
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  const size_t total = static_cast<size_t>(N) * Ho * Wo * C;
  if (i >= total) return;
  const int c = static_cast<int>(i % C);
  const size_t pix = i / C;
  const int xo = static_cast<int>(pix % Wo);
  const int yo = static_cast<int>((pix / Wo) % Ho);
  const int n = static_cast<int>(pix / (static_cast<size_t>(Wo) * Ho));
  const int py = (Ho * k - H) / 2, px = (Wo * k - W) / 2;
  float m = -INFINITY, sum = 0.f;
  int cnt = 0;
  for (int dy = 0; dy < k; ++dy)
    for (int dx = 0; dx < k; ++dx) {
      const int y = yo * k + dy - py, xq = xo * k + dx - px;
      if (y < 0 || y >= H || xq < 0 || xq >= W) continue;
      const float v = __ldg(x + ((static_cast<size_t>(n) * H + y) * W + xq) * C + c);
      m = fmaxf(m, v); sum += v; ++cnt;
    }
  out[i] = avg ? sum / static_cast<float>(cnt) : m;
}

// The model's `batchnorm` helper (hgru_pose.py:120-122): tf.nn.moments over axis 0 (the batch, separately for every
// position and channel) and tf.nn.batch_normalization without scale / offset: (x - mean) * rsqrt(var + eps).
// x [N][inner]; one thread per inner index.
__global__ void __launch_bounds__(256)
batchnorm_moments0_kernel(const float* __restrict__ x, int N, size_t inner, float eps, float* __restrict__ out) {
  const size_t j = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (j >= inner) return;
  double s = 0.0;
  for (int n = 0; n < N; ++n) s += static_cast<double>(x[static_cast<size_t>(n) * inner + j]);
  const float mean = static_cast<float>(s / N);
  double q = 0.0;
  for (int n = 0; n < N; ++n) {
    const double d = static_cast<double>(x[static_cast<size_t>(n) * inner + j]) - static_cast<double>(mean);
    q += d * d;
  }
  const float inv = rsqrtf(static_cast<float>(q / N) + eps);
  for (int n = 0; n < N; ++n) {
    const size_t i = static_cast<size_t>(n) * inner + j;
    out[i] = x[i] * inv + (-mean * inv);
  }
}

// x [M][K] @ w [K][F] + b [F]: block = 64 outputs j of one row m... one thread per (m, j), K walked in chunks whose
// fp32 partial sums are added in double (K is 262 144 for fc_1).
__global__ void __launch_bounds__(256)
fc_direct_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                 float* __restrict__ out, int M, int K, int F) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<size_t>(M) * F) return;
  const int j = static_cast<int>(i % F);
  const int m = static_cast<int>(i / F);
  const float* xr = x + static_cast<size_t>(m) * K;
  double acc = 0.0;
  for (int k0 = 0; k0 < K; k0 += 256) {
    const int k1 = k0 + 256 < K ? k0 + 256 : K;
    float part = 0.f;
    for (int k = k0; k < k1; ++k) part = fmaf(__ldg(xr + k), __ldg(w + static_cast<size_t>(k) * F + j), part);
    acc += static_cast<double>(part);
  }
  out[i] = static_cast<float>(acc + static_cast<double>(__ldg(b + j)));
}

// ---- train-mode batch normalisation ------------------------------------------------------------------------------
// Element-wise prologue shared by the two passes: optional relu (hgru_pose.py:92) and optional inverted dropout
// (tf.nn.dropout(x, keep_prob), :93-94: kept elements are scaled by 1 / keep_prob).  TensorFlow's random stream
// cannot be reproduced, so the keep mask is a counter-based hash of (seed, element index): splitmix64, keep when the
// top 24 bits / 2^24 < keep_prob.  keep >= 1 disables dropout.
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ float bn_pre_op(float v, size_t idx, int relu, float keep, unsigned long long seed) {
  if (relu) v = fmaxf(v, 0.f);
  if (keep < 1.f) {
    const unsigned long long h = splitmix64(seed ^ (static_cast<unsigned long long>(idx) * 0xD1B54A32D192ED03ull));
    const float u = static_cast<float>(h >> 40) * (1.0f / 16777216.0f);
    v = (u < keep) ? v / keep : 0.f;
  }
  return v;
}

// Batch statistics over every axis but the last (tf.layers.batch_normalization(training=True, fused=True): mean and
// BIASED variance of the batch).  x is [rows][C] (NHWC flattened).  Pass 1: per-channel sum and sum of squares in
// double, blocks = (channel group of 32) x (row chunk), combined with double atomics into sums[c], sums[C + c]
// (zeroed by the caller).  Double keeps the one-pass variance exact to ~1e-13 at these sizes.
__global__ void __launch_bounds__(256)
bn_batch_sums_kernel(const float* __restrict__ x, size_t rows, int C, double* __restrict__ sums, int relu,
                     float keep, unsigned long long seed) {
  __shared__ double sh[2][8][32];
  const int lane = threadIdx.x & 31, rlane = threadIdx.x >> 5;          // 32 channels x 8 row lanes
  const int c = blockIdx.x * 32 + lane;
  const size_t per = (rows + gridDim.y - 1) / gridDim.y;
  const size_t r0 = blockIdx.y * per, r1 = r0 + per < rows ? r0 + per : rows;
  double s = 0.0, q = 0.0;
  if (c < C)
    for (size_t r = r0 + rlane; r < r1; r += 8) {
      const double v = static_cast<double>(bn_pre_op(__ldg(x + r * C + c), r * C + c, relu, keep, seed));
      s += v;
      q += v * v;
    }
  sh[0][rlane][lane] = s;
  sh[1][rlane][lane] = q;
  __syncthreads();
  if (rlane == 0 && c < C) {
    for (int k = 1; k < 8; ++k) { s += sh[0][k][lane]; q += sh[1][k][lane]; }
    atomicAdd(sums + c, s);
    atomicAdd(sums + C + c, q);
  }
}

// Pass 2: y = (x - mean) / sqrt(var + eps) * gamma + beta with mean = S / rows, var = Q / rows - mean^2; also the
// moving-statistics update the reference's UPDATE_OPS perform (train_cnn_networks_hgru.py:123-126): moving = moving *
// momentum + batch * (1 - momentum), with the UNBIASED batch variance as TF's fused kernel reports it -- written to
// new_mean / new_var when non-null (threads i < C do it).  batch_mean / batch_var (biased) are exported when non-null.
__global__ void __launch_bounds__(256)
bn_apply_batch_kernel(const float* __restrict__ x, float* __restrict__ y, size_t rows, int C,
                      const double* __restrict__ sums, const float* __restrict__ gamma,
                      const float* __restrict__ beta, float eps, const float* __restrict__ moving_mean,
                      const float* __restrict__ moving_var, float momentum, float* __restrict__ new_mean,
                      float* __restrict__ new_var, int relu, float keep, unsigned long long seed) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  const double inv_rows = 1.0 / static_cast<double>(rows);
  if (i < static_cast<size_t>(C) && new_mean && new_var) {
    const int c = static_cast<int>(i);
    const double mean = sums[c] * inv_rows;
    double var = sums[C + c] * inv_rows - mean * mean;
    var = var > 0.0 ? var : 0.0;
    const double unbiased = rows > 1 ? var * static_cast<double>(rows) / static_cast<double>(rows - 1) : var;
    new_mean[c] = static_cast<float>(static_cast<double>(moving_mean[c]) * momentum + mean * (1.0 - momentum));
    new_var[c] = static_cast<float>(static_cast<double>(moving_var[c]) * momentum + unbiased * (1.0 - momentum));
  }
  if (i >= rows * C) return;
  const int c = static_cast<int>(i % C);
  const double mean = sums[c] * inv_rows;
  double var = sums[C + c] * inv_rows - mean * mean;
  var = var > 0.0 ? var : 0.0;
  const double v = static_cast<double>(bn_pre_op(x[i], i, relu, keep, seed));
  y[i] = static_cast<float>((v - mean) / sqrt(var + static_cast<double>(eps)) * static_cast<double>(__ldg(gamma + c)) +
                            static_cast<double>(__ldg(beta + c)));
}


// tf.layers.batch_normalization(training=False): y = (x - moving_mean) / sqrt(moving_var + eps) * gamma + beta
__global__ void __launch_bounds__(256)
bn_inference_kernel(const float* __restrict__ x, float* __restrict__ y, size_t total, int C,
                    const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ mean,
                    const float* __restrict__ var, float eps, int relu) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int c = static_cast<int>(i % C);
  float v = x[i];
  if (relu) v = fmaxf(v, 0.f);
  const float sc = __ldg(gamma + c) / sqrtf(__ldg(var + c) + eps);
  y[i] = (v - __ldg(mean + c)) * sc + __ldg(beta + c);
}

// ---- the circuit's per-timestep building blocks (hgru_module.py:692-823) as stand-alone exact-fp32 ops -------------
// What a caller composing `full()` by hand gets (ContextualCircuit.circuit_input / circuit_output / input_integration /
// output_integration / full); `build()` never runs them -- it runs the fused tensor-core pipeline.  Tensors are
// [rows][k] channels-last (rows = n*h*w); the 15x15 convolution between them is conv2d_direct_kernel.

// gate = sigmoid(x *1x1 w + b) (:696-707, 729-740); gated = x . gate when asked for (:709-711).  One thread per
// (row, output channel).
__global__ void __launch_bounds__(256)
circuit_gate_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                    float* __restrict__ gate, float* __restrict__ gated, size_t rows, int k) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= rows * k) return;
  const int co = static_cast<int>(i % k);
  const float* xr = x + (i / k) * k;
  float acc = 0.f;
  for (int ci = 0; ci < k; ++ci) acc = fmaf(__ldg(xr + ci), __ldg(w + static_cast<size_t>(ci) * k + co), acc);
  const float g = 1.f / (1.f + expf(-(acc + __ldg(b + co))));
  gate[i] = g;
  if (gated) gated[i] = xr[co] * g;
}

// I = tanh(xi * X - (beta . O + nu) . P) (:795-804, gru_gates: no mixing with the old I)
__global__ void __launch_bounds__(256)
circuit_input_integration_kernel(const float* __restrict__ X, const float* __restrict__ O, const float* __restrict__ P,
                                 const float* __restrict__ beta, const float* __restrict__ nu, float xi, size_t total,
                                 int k, float* __restrict__ I) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int c = static_cast<int>(i % k);
  I[i] = tanhf(xi * X[i] - (__ldg(beta + c) * O[i] + __ldg(nu + c)) * P[i]);
}

// O' = G . O + (1 - G) . tanh(kappa . (zeta I + gamma . P) + omega . (zeta I . gamma . P)) (:806-823), times *rho when a
// pointer is given (the `O * rho[i0]` of full(), :847-849)
__global__ void __launch_bounds__(256)
circuit_output_integration_kernel(const float* __restrict__ I, const float* __restrict__ P, const float* __restrict__ O,
                                  const float* __restrict__ G, const float* __restrict__ gamma,
                                  const float* __restrict__ kappa, const float* __restrict__ omega, float zeta,
                                  const float* __restrict__ rho, size_t total, int k, float* __restrict__ out) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int c = static_cast<int>(i % k);
  const float zi = zeta * I[i];
  const float e = __ldg(gamma + c) * P[i];
  const float s = tanhf(__ldg(kappa + c) * (zi + e) + __ldg(omega + c) * (zi * e));
  const float g = G[i];
  const float o = g * O[i] + (1.f - g) * s;
  out[i] = rho ? o * __ldg(rho) : o;
}

}  // namespace hgru
