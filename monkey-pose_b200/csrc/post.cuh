// The step right after the pose network: de-normalise the [N,3J] output by
// cube[2]/2 (train_cnn_networks_hgru.py:293-296), add the crop's centre of mass in camera space and
// project back to the image (tfMonkeyDetector.getAbsoluteCoordinates, tf_monkeydetector.py:387-391 with
// uvdtoxyz :138-160 and xyztouvd :116-136), and the mean / max per-joint error (pose_evaluation.py:10-23).
// The reference runs these per frame in numpy on the host; float32 results here are bit-identical
// (explicitly rounded IEEE operations, no FMA contraction).
#pragma once
#include <cuda_runtime.h>

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

}  // namespace hgru
