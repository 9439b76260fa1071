// Crop stage feeding the pose network: the reference's
// tfMonkeyDetector.cropArea3D (tf_monkeydetector.py:292-365) = comToBounds (:193-206, host side) ->
// getCrop (slice, zero pad, z clamp; :208-247) -> cv2.resize INTER_NEAREST (:249-263) -> paste into a
// max-depth background -> / image_max_depth (train_cnn_networks_hgru.py:61-74), one Python/cv2 call per
// frame on the host.  Here: one thread per output pixel gathers straight from the full depth frame,
// bit-exact with the reference (the nearest-neighbour index is computed in double exactly like OpenCV).
#pragma once
#include <cuda_runtime.h>

#include "np_reduce.cuh"

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


// Raw depth frames (16-bit millimetres, as a depth camera delivers them and as the reference's real-data loop loads
// them) -> the float32 frames in [0, 1] the networks and the crop stage take: values outside [near, far] are replaced by
// `fill` and the result is divided by max_depth (eval_model_on_real_data, train_cnn_networks_hgru.py:381-386 + the
// `/ config.image_max_depth` of :359 / :392).  Half the PCIe bytes of uploading float32 frames.  The quotient is formed
// in double and rounded once to float32, as numpy's `im / 10000.` followed by the float32 feed does.
__global__ void __launch_bounds__(256)
depth_preprocess_kernel(const unsigned short* __restrict__ raw, size_t n, unsigned int near_mm, unsigned int far_mm,
                        double fill, double max_depth, float* __restrict__ out) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const unsigned int v = raw[i];
  const double d = (v < near_mm || v > far_mm) ? fill : static_cast<double>(v);
  out[i] = static_cast<float>(__ddiv_rn(d, max_depth));
}

// Window arithmetic of the crop on the device (tfMonkeyDetector.comToBounds + the resize / paste integers and the
// 3x3 transform of cropArea3D, tf_monkeydetector.py:193-206, 309-362), so that attention output -> crop needs no
// host round trip.  One thread per frame; every operation is an explicitly rounded IEEE double operation in the
// order the reference's numpy expressions evaluate them (no FMA contraction), so the integers are the host's.
//   tr [N][3] float32: the attention CNN's output (u / height, v / width, d / max depth as the reference reads it);
//   coms = double(tr) * tr_scale (train_cnn_networks_hgru.py:66-68).  com_in (double [N][3]) overrides tr when given.
// A window that misses the frame gets resized size 0 (an all-background patch) and invalid[n] = 1.
__global__ void __launch_bounds__(128)
crop_windows_kernel(const float* __restrict__ tr, const double* __restrict__ com_in, double s0, double s1, double s2,
                    int N, int H, int W, int dw, int dh, double fx, double fy, double cx, double cy, double cz,
                    double* __restrict__ coms, int* __restrict__ ip, float* __restrict__ zp, double* __restrict__ Ms,
                    int* __restrict__ invalid) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  double c0, c1, c2;
  if (com_in) {
    c0 = com_in[3 * n]; c1 = com_in[3 * n + 1]; c2 = com_in[3 * n + 2];
  } else {
    c0 = __dmul_rn(static_cast<double>(tr[3 * n]), s0);
    c1 = __dmul_rn(static_cast<double>(tr[3 * n + 1]), s1);
    c2 = __dmul_rn(static_cast<double>(tr[3 * n + 2]), s2);
  }
  coms[3 * n] = c0; coms[3 * n + 1] = c1; coms[3 * n + 2] = c2;
  const double hx = cx / 2., hy = cy / 2., hz = cz / 2.;
  const double zstart = __dsub_rn(c2, hz), zend = __dadd_rn(c2, hz);
  // floor((c * z / f -+ size / 2) / z * f)
  const double ax = __ddiv_rn(__dmul_rn(c0, c2), fx), ay = __ddiv_rn(__dmul_rn(c1, c2), fy);
  const double fxs = floor(__dmul_rn(__ddiv_rn(__dsub_rn(ax, hx), c2), fx));
  const double fxe = floor(__dmul_rn(__ddiv_rn(__dadd_rn(ax, hx), c2), fx));
  const double fys = floor(__dmul_rn(__ddiv_rn(__dsub_rn(ay, hy), c2), fy));
  const double fye = floor(__dmul_rn(__ddiv_rn(__dadd_rn(ay, hy), c2), fy));
  const bool finite = isfinite(c0) && isfinite(c1) && isfinite(c2) && isfinite(fxs) && isfinite(fxe) &&
                      isfinite(fys) && isfinite(fye) && fabs(fxs) < 1e9 && fabs(fxe) < 1e9 && fabs(fys) < 1e9 &&
                      fabs(fye) < 1e9;
  long long xstart = finite ? static_cast<long long>(fxs) : 0, xend = finite ? static_cast<long long>(fxe) : 1;
  long long ystart = finite ? static_cast<long long>(fys) : 0, yend = finite ? static_cast<long long>(fye) : 1;
  const bool bad = !finite || xend <= 0 || yend <= 0 || xstart >= W || ystart >= H || xend <= xstart || yend <= ystart;
  if (bad) { xstart = 0; ystart = 0; xend = 1; yend = 1; }
  const long long wb = xend - xstart, hb = yend - ystart;
  long long szx, szy;
  if (wb > hb) { szx = dw; szy = hb * dw / wb; } else { szx = wb * dh / hb; szy = dh; }      // (positive: // == /)
  if (bad) { szx = 0; szy = 0; }
  const double sc = (hb > wb) ? __ddiv_rn(static_cast<double>(szy), static_cast<double>(hb))
                              : __ddiv_rn(static_cast<double>(szx), static_cast<double>(wb));
  const long long xs = static_cast<long long>(floor(__dsub_rn(dw / 2., szx / 2.)));
  const long long ys = static_cast<long long>(floor(__dsub_rn(dh / 2., szy / 2.)));
  int* q = ip + 8 * n;
  q[0] = static_cast<int>(xstart); q[1] = static_cast<int>(ystart); q[2] = static_cast<int>(wb); q[3] = static_cast<int>(hb);
  q[4] = static_cast<int>(szx); q[5] = static_cast<int>(szy); q[6] = static_cast<int>(xs); q[7] = static_cast<int>(ys);
  zp[2 * n] = static_cast<float>(zstart);
  zp[2 * n + 1] = static_cast<float>(zend);
  // M = off @ scale @ trans written out (the products only ever add exact zeros)
  double* M = Ms + 9 * n;
  M[0] = sc; M[1] = 0.; M[2] = __dadd_rn(__dmul_rn(sc, -static_cast<double>(xstart)), static_cast<double>(xs));
  M[3] = 0.; M[4] = sc; M[5] = __dadd_rn(__dmul_rn(sc, -static_cast<double>(ystart)), static_cast<double>(ys));
  M[6] = 0.; M[7] = 0.; M[8] = 1.;
  invalid[n] = bad ? 1 : 0;
}

// tfMonkeyDetector.calculateCoM (tf_monkeydetector.py:73-90) for a batch: depth outside [minDepth, maxDepth] zeroed,
// centre of mass of the mask (scipy.ndimage.center_of_mass of dc > 0: integer sums, exact), mean depth =
// dc.sum() / count_nonzero(dc).  dc.sum() is a float32 sum whose value depends on numpy's pairwise order, so the
// blocks of numpy's tree are summed one per thread (np_block_sum) and combined level by level by heap index
// (np_reduce.cuh): the result is the host's, bit for bit.  One CTA per frame.
//   ip == nullptr: the image is the whole frame (cropArea3D with com=None, :307-308).
//   ip / zp given (crop_windows_kernel's output): the image is getCrop's window (:208-244) -- the frame's pixels where
//   the window overlaps it, zeros elsewhere, z-clamped -- and the result gets cropArea3D's `docom` treatment
//   (:318-326): if the CoM is all zero take the window's centre pixel as depth (300 if that is zero too), then add
//   (xstart, ystart).
// heap [N][heap_cap] floats of workspace; a frame whose image needs more (window larger than the caller's bound) gets
// a NaN centre of mass and overflow[n] = 1.
struct ComImage {
  const float* frame;
  float scale, zs, ze, lo, hi;
  int H, W, xstart, ystart, wb;
  bool clamp;
  // element i of the (virtual) image after getCrop's z clamp and calculateCoM's range test
  __device__ __forceinline__ float crop_value(unsigned i, unsigned* row, unsigned* col) const {
    const unsigned r = i / static_cast<unsigned>(wb), c = i - r * static_cast<unsigned>(wb);
    *row = r; *col = c;
    const int Y = ystart + static_cast<int>(r), X = xstart + static_cast<int>(c);
    float v = 0.f;
    if (Y >= 0 && Y < H && X >= 0 && X < W) v = __fmul_rn(frame[static_cast<size_t>(Y) * W + X], scale);
    if (clamp && v != 0.f) {
      if (v < zs) v = zs;
      else if (v > ze) v = 0.f;
    }
    return v;
  }
  __device__ __forceinline__ float com_value(float v) const {
    if (v < lo) v = 0.f;
    if (v > hi) v = 0.f;
    return v;
  }
};

__global__ void __launch_bounds__(1024)
calculate_com_kernel(const float* __restrict__ frames, int H, int W, float frame_scale, float min_depth,
                     float max_depth, const int* __restrict__ ip, const float* __restrict__ zp, float* __restrict__ heap,
                     unsigned heap_cap, double* __restrict__ coms, int* __restrict__ overflow) {
  __shared__ unsigned long long s_stat[4];     // #(dc > 0), #(dc != 0), sum of columns, sum of rows over the mask
  const int n = blockIdx.x;
  const int tid = threadIdx.x, T = blockDim.x;
  ComImage im;
  im.frame = frames + static_cast<size_t>(n) * H * W;
  im.scale = frame_scale; im.lo = min_depth; im.hi = max_depth; im.H = H; im.W = W;
  int hb;
  if (ip) {
    const int* q = ip + 8 * n;
    im.xstart = q[0]; im.ystart = q[1]; im.wb = q[2]; hb = q[3];
    im.zs = zp[2 * n]; im.ze = zp[2 * n + 1]; im.clamp = true;
  } else {
    im.xstart = 0; im.ystart = 0; im.wb = W; hb = H; im.zs = 0.f; im.ze = 0.f; im.clamp = false;
  }
  const long long npx = static_cast<long long>(im.wb) * hb;
  const int depth = np_pairwise_depth(npx);
  if (npx < 1 || npx > 0x7fffffffLL || (2ull << depth) > heap_cap) {
    if (tid == 0) {
      coms[3 * n] = coms[3 * n + 1] = coms[3 * n + 2] = nan("");
      overflow[n] = 1;
    }
    return;
  }
  if (tid < 4) s_stat[tid] = 0ull;
  __syncthreads();
  float* vals = heap + static_cast<size_t>(n) * heap_cap;
  unsigned long long pos = 0, nz = 0, sx = 0, sy = 0;
  np_tree_blocks([&](long long i) {
    unsigned r, c;
    const float v = im.com_value(im.crop_value(static_cast<unsigned>(i), &r, &c));
    if (v > 0.f) { ++pos; sx += c; sy += r; }
    if (v != 0.f) ++nz;
    return v;
  }, npx, tid, T, vals);
  // the integer statistics do not depend on the order of addition
  for (int o = 16; o > 0; o >>= 1) {
    pos += __shfl_xor_sync(0xffffffffu, pos, o); nz += __shfl_xor_sync(0xffffffffu, nz, o);
    sx += __shfl_xor_sync(0xffffffffu, sx, o);   sy += __shfl_xor_sync(0xffffffffu, sy, o);
  }
  if ((tid & 31) == 0) {
    atomicAdd(&s_stat[0], pos); atomicAdd(&s_stat[1], nz); atomicAdd(&s_stat[2], sx); atomicAdd(&s_stat[3], sy);
  }
  __syncthreads();
  for (int level = depth - 1; level >= 0; --level) {
    np_tree_level(npx, level, tid, T, vals);
    __syncthreads();
  }
  if (tid != 0) return;
  overflow[n] = 0;
  const double cnt = static_cast<double>(s_stat[0]), num = static_cast<double>(s_stat[1]);
  double c0 = 0., c1 = 0., c2 = 0.;
  if (s_stat[1] != 0ull) {
    // cc = sums / count of the mask; com = (cc[1] * num, cc[0] * num, dc.sum()) / num, every step a rounded double op
    c0 = __ddiv_rn(__dmul_rn(__ddiv_rn(static_cast<double>(s_stat[2]), cnt), num), num);
    c1 = __ddiv_rn(__dmul_rn(__ddiv_rn(static_cast<double>(s_stat[3]), cnt), num), num);
    c2 = __ddiv_rn(static_cast<double>(vals[1]), num);
  }
  if (ip) {
    if (fabs(c0) <= 1e-8 && fabs(c1) <= 1e-8 && fabs(c2) <= 1e-8) {          // numpy.allclose(com, 0.)
      unsigned r, c;
      c2 = static_cast<double>(im.crop_value(static_cast<unsigned>(hb / 2) * im.wb + im.wb / 2, &r, &c));
      if (fabs(c2) <= 1e-8) c2 = 300.;                                        // numpy.isclose(com[2], 0)
    }
    c0 = __dadd_rn(c0, static_cast<double>(im.xstart));
    c1 = __dadd_rn(c1, static_cast<double>(im.ystart));
  }
  coms[3 * n] = c0; coms[3 * n + 1] = c1; coms[3 * n + 2] = c2;
}

}  // namespace hgru
