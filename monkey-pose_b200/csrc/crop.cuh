// Crop stage feeding the pose network: the reference's
// tfMonkeyDetector.cropArea3D (tf_monkeydetector.py:292-365) = comToBounds (:193-206, host side) ->
// getCrop (slice, zero pad, z clamp; :208-247) -> cv2.resize INTER_NEAREST (:249-263) -> paste into a
// max-depth background -> / image_max_depth (train_cnn_networks_hgru.py:61-74), one Python/cv2 call per
// frame on the host.  Here: one thread per output pixel gathers straight from the full depth frame,
// bit-exact with the reference (the nearest-neighbour index is computed in double exactly like OpenCV).
#pragma once
#include <cuda_runtime.h>

namespace hgru {

// per-frame integers: xstart, ystart, wb, hb (source window), sz_w, sz_h (resized crop), xs, ys (paste offset)
__global__ void __launch_bounds__(256)
crop_area3d_kernel(const float* __restrict__ frames, float frame_scale, const int* __restrict__ ip,
                   const float* __restrict__ zp, float background, double out_div, float* __restrict__ out,
                   int N, int H, int W, int dh, int dw) {
  const size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<size_t>(N) * dh * dw) return;
  const int ox = idx % dw;
  const int oy = (idx / dw) % dh;
  const int n = idx / (static_cast<size_t>(dw) * dh);
  const int* q = ip + n * 8;
  const int xstart = q[0], ystart = q[1], wb = q[2], hb = q[3], sz_w = q[4], sz_h = q[5], xs = q[6], ys = q[7];
  float v = background;
  const int ry = oy - ys, rx = ox - xs;
  if (ry >= 0 && ry < sz_h && rx >= 0 && rx < sz_w) {
    // OpenCV resizeNN: src = min(floor(dst * (1 / (dsize / ssize))), ssize - 1), in double
    const double ify = 1.0 / (static_cast<double>(sz_h) / static_cast<double>(hb));
    const double ifx = 1.0 / (static_cast<double>(sz_w) / static_cast<double>(wb));
    const int sy = min(static_cast<int>(floor(ry * ify)), hb - 1);
    const int sx = min(static_cast<int>(floor(rx * ifx)), wb - 1);
    const int Y = ystart + sy, X = xstart + sx;
    v = 0.f;                                                   // zero padding outside the frame
    if (Y >= 0 && Y < H && X >= 0 && X < W)
      v = __fmul_rn(__ldg(frames + (static_cast<size_t>(n) * H + Y) * W + X), frame_scale);
    const float zs = zp[2 * n], ze = zp[2 * n + 1];
    if (v != 0.f) {
      if (v < zs) v = zs;                                      // in front of the cube: clamp to its face
      else if (v > ze) v = 0.f;                                // behind it: back face, set to 0
    }
  }
  out[idx] = static_cast<float>(static_cast<double>(v) / out_div);
}

}  // namespace hgru
