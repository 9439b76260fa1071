// The step right after the pose network: de-normalise the [N,3J] output by
// cube[2]/2 (train_cnn_networks_hgru.py:293-296), add the crop's centre of mass in camera space and
// project back to the image (tfMonkeyDetector.getAbsoluteCoordinates, tf_monkeydetector.py:387-391 with
// uvdtoxyz :138-160 and xyztouvd :116-136), and the mean / max per-joint error (pose_evaluation.py:10-23).
// The reference runs these per frame in numpy on the host; float32 results here are bit-identical
// (explicitly rounded IEEE operations, no FMA contraction).
#pragma once
#include <cuda_runtime.h>

#include "np_reduce.cuh"

namespace hgru {

// one thread per joint
__global__ void __launch_bounds__(256)
pose_postprocess_kernel(const float* __restrict__ out_put, const double* __restrict__ com_uvd, int N, int J,
                        double fx, double fy, double ux, double uy, float scale, float* __restrict__ xyz,
                        float* __restrict__ uvd) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * J) return;
  const int n = i / J;
  const double cu = com_uvd[3 * n], cv = com_uvd[3 * n + 1], cd = com_uvd[3 * n + 2];
  // uvdtoxyz: float64 arithmetic, stored as float32
  const float cx = static_cast<float>((ux - cu) * cd / (-fx));
  const float cy = static_cast<float>((cv - uy) * cd / (-fy));
  const float cz = static_cast<float>(-cd);
  const float x = __fadd_rn(__fmul_rn(out_put[3 * i], scale), cx);
  const float y = __fadd_rn(__fmul_rn(out_put[3 * i + 1], scale), cy);
  const float z = __fadd_rn(__fmul_rn(out_put[3 * i + 2], scale), cz);
  xyz[3 * i] = x; xyz[3 * i + 1] = y; xyz[3 * i + 2] = z;
  const float fxf = static_cast<float>(fx), fyf = static_cast<float>(fy);
  const float uxf = static_cast<float>(ux), uyf = static_cast<float>(uy);
  float u, v, d;
  if (z == 0.f) { u = uxf; v = uyf; d = 0.f; }
  else {
    u = __fsub_rn(uxf, __fmul_rn(__fdiv_rn(x, z), fxf));
    v = __fadd_rn(__fmul_rn(__fdiv_rn(y, z), fyf), uyf);
    d = -z;
  }
  uvd[3 * i] = u; uvd[3 * i + 1] = v; uvd[3 * i + 2] = d;
}

// per-frame sum and max of the per-joint Euclidean errors (NaNs skipped, as numpy.nanmean / nanmax);
// one block per frame, the tiny [N] reduction is finished by the caller-side second kernel
__global__ void __launch_bounds__(128)
joint_error_frame_kernel(const float* __restrict__ labels, const float* __restrict__ results, int J,
                         double* __restrict__ frame_mean, float* __restrict__ frame_max) {
  __shared__ double ssum[128];
  __shared__ float smax[128];
  __shared__ int scnt[128];
  const int n = blockIdx.x;
  double s = 0.0; float mx = -INFINITY; int cnt = 0;
  for (int j = threadIdx.x; j < J; j += blockDim.x) {
    const float* a = labels + (static_cast<size_t>(n) * J + j) * 3;
    const float* b = results + (static_cast<size_t>(n) * J + j) * 3;
    float acc = 0.f;                           // float32 square, sum over the 3 coordinates, sqrt (numpy order)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float dlt = __fsub_rn(a[c], b[c]);
      acc = __fadd_rn(acc, __fmul_rn(dlt, dlt));
    }
    const float e = __fsqrt_rn(acc);
    if (e == e) { s += e; ++cnt; mx = fmaxf(mx, e); }
  }
  ssum[threadIdx.x] = s; smax[threadIdx.x] = mx; scnt[threadIdx.x] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0; float m = -INFINITY; int c = 0;
    for (int k = 0; k < blockDim.x; ++k) { t += ssum[k]; m = fmaxf(m, smax[k]); c += scnt[k]; }
    frame_mean[n] = c ? t / c : nan("");
    frame_max[n] = m;
  }
}
__global__ void joint_error_final_kernel(const double* __restrict__ frame_mean, const float* __restrict__ frame_max,
                                         int N, double* __restrict__ result /* [2]: mean, max */) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double t = 0.0; float m = -INFINITY; int c = 0;
  for (int n = 0; n < N; ++n) {
    if (frame_mean[n] == frame_mean[n]) { t += frame_mean[n]; ++c; }
    m = fmaxf(m, frame_max[n]);
  }
  result[0] = c ? t / c : nan("");
  result[1] = m;
}

// ---- the whole metric set of pose_evaluation.py:10-88 in numpy's own float32 evaluation order ------------------
// Per-joint Euclidean error err[n][j] = sqrt(((l - r)^2).sum(axis=2)): float32 subtract, square, the three squares
// added left to right (a contiguous run shorter than 8), sqrt -- the expression every metric of the file starts from.
__global__ void __launch_bounds__(256)
joint_error_matrix_kernel(const float* __restrict__ labels, const float* __restrict__ results, size_t NJ,
                          float* __restrict__ err) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= NJ) return;
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float d = __fsub_rn(labels[3 * i + c], results[3 * i + c]);
    acc = __fadd_rn(acc, __fmul_rn(d, d));
  }
  err[i] = __fsqrt_rn(acc);
}

// Threads [0, N): frame_mean[n] = nanmean(err[n, :]) (getMeanErrors_N, the inner mean of getMeanError_np / _train,
// getNumFramesWithinMeanDist) and frame_max[n] = nanmax(err[n, :]) (getNumFramesWithinMaxDist).  Threads [N, N + J):
// joint_mean[j] = nanmean(err[:, j]) (getJointMeanError for every joint at once).  skip_nan = 0 gives the TensorFlow
// variants' semantics (reduce_mean / reduce_max: a NaN poisons the result).
__global__ void __launch_bounds__(128)
joint_error_reduce_kernel(const float* __restrict__ err, int N, int J, int skip_nan, float* __restrict__ frame_mean,
                          float* __restrict__ frame_max, float* __restrict__ joint_mean) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < N) {
    const float* e = err + static_cast<size_t>(t) * J;
    frame_mean[t] = np_nanmean([&](long long i) { return e[i]; }, J, skip_nan != 0);
    float m = nanf("");
    for (int j = 0; j < J; ++j) {
      const float v = e[j];
      if (v != v) { if (!skip_nan) { m = v; break; } continue; }
      m = (m != m) ? v : fmaxf(m, v);
    }
    frame_max[t] = m;
  } else if (t < N + J) {
    const int j = t - N;
    joint_mean[j] = np_nanmean([&](long long i) { return err[static_cast<size_t>(i) * J + j]; }, N, skip_nan != 0);
  }
}

// summary[0] = nanmean(frame_mean) (getMeanError_np / getMeanError_train), summary[1] = nanmax over everything
// (getMaxError_np / getMaxError)
__global__ void joint_error_summary_kernel(const float* __restrict__ frame_mean, const float* __restrict__ frame_max,
                                           int N, int skip_nan, float* __restrict__ summary) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  summary[0] = np_nanmean([&](long long i) { return frame_mean[i]; }, N, skip_nan != 0);
  float m = nanf("");
  for (int n = 0; n < N; ++n) {
    const float v = frame_max[n];
    if (v != v) { if (!skip_nan) { m = v; break; } continue; }
    m = (m != m) ? v : fmaxf(m, v);
  }
  summary[1] = m;
}

// number of frames whose statistic is <= dist (a NaN compares false, as in numpy)
__global__ void __launch_bounds__(256)
count_within_kernel(const float* __restrict__ stat, int N, float dist, int* __restrict__ count) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const bool in = n < N && stat[n] <= dist;
  const unsigned b = __ballot_sync(0xffffffffu, in);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(count, __popc(b));
}

// getMean_np / getMeanError (pose_evaluation.py:26-28, 38-44): sqrt(((a - b)^2).sum(axis=1)) averaged over axis 0.
// a, b [N][M][C] (C = 1 for rank-2 inputs).  A non-contiguous axis is summed element by element in numpy, so both
// stages are plain left-to-right float32 sums.  rows [N][C] = the square roots; out [C] = their (nan)mean.
__global__ void __launch_bounds__(256)
axis1_error_rows_kernel(const float* __restrict__ a, const float* __restrict__ b, int N, int M, int C,
                        float* __restrict__ rows) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<size_t>(N) * C) return;
  const size_t n = i / C;
  const int c = static_cast<int>(i % C);
  float acc = 0.f;
  if (C == 1) {
    // rank-2 inputs: axis 1 IS the contiguous axis -> pairwise order
    const float* pa = a + n * M;
    const float* pb = b + n * M;
    acc = np_pairwise_sum([&](long long m) { const float d = __fsub_rn(pa[m], pb[m]); return __fmul_rn(d, d); }, M);
  } else {
    for (int m = 0; m < M; ++m) {
      const size_t k = (n * M + m) * C + c;
      const float d = __fsub_rn(a[k], b[k]);
      const float sq = __fmul_rn(d, d);
      acc = (m == 0) ? sq : __fadd_rn(acc, sq);
    }
  }
  rows[i] = __fsqrt_rn(acc);
}
__global__ void axis0_mean_kernel(const float* __restrict__ rows, int N, int C, int skip_nan, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (C == 1) {      // a contiguous [N] vector: pairwise
    out[0] = np_nanmean([&](long long i) { return rows[i]; }, N, skip_nan != 0);
    return;
  }
  float tot = 0.f;
  long long cnt = 0;
  for (int n = 0; n < N; ++n) {
    float v = rows[static_cast<size_t>(n) * C + c];
    if (skip_nan && v != v) v = 0.f; else ++cnt;
    tot = (n == 0) ? v : __fadd_rn(tot, v);
  }
  out[c] = static_cast<float>(__ddiv_rn(static_cast<double>(tot), static_cast<double>(cnt)));
}

}  // namespace hgru
