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
// blocks of numpy's tree (np_reduce.cuh) are summed exactly as numpy sums them -- eight interleaved accumulators, here
// eight lanes -- and combined level by level by heap index: the result is the host's, bit for bit.
//   ip == nullptr: the image is the whole frame (cropArea3D with com=None, :307-308).
//   ip / zp given (crop_windows_kernel's output): the image is getCrop's window (:208-244) -- the frame's pixels where
//   the window overlaps it, zeros elsewhere, z-clamped -- and the result gets cropArea3D's `docom` treatment
//   (:318-326): if the CoM is all zero take the window's centre pixel as depth (300 if that is zero too), then add
//   (xstart, ystart).
// Two launches.  com_blocks_kernel is the HBM-bound one: grid (parts, N); a group of 8 lanes owns a subtree of numpy's
// tree (2^leaves_log2 consecutive blocks below a level-L node) and sums it one block at a time (lane j = numpy's
// accumulator r[j], so a warp reads four 32-byte runs per load and every byte of a frame exactly once), folding the block
// sums into the node's sum as they come; it writes that to heap[n][heap index of the node] and adds its integer
// statistics (counts, row and column sums of the mask: order-free) to stats[n].  com_finish_kernel (one CTA per frame)
// combines the top L levels in shared memory and does the few double operations of the reference in its order.
// heap [N][heap_cap] floats (heap_cap = 2^(L+1) for the largest image allowed), stats [N][4] uint64 (zeroed by the
// caller); a frame whose image needs a bigger heap (window larger than the caller's bound) gets a NaN centre of mass
// and overflow[n] = 1.
struct ComImage {
  const float* frame;
  float scale, zs, ze, lo, hi;
  int H, W, xstart, ystart, wb, hb;
  bool window;
  __device__ __forceinline__ void load(const float* frames, int n, int H_, int W_, float frame_scale, float min_depth,
                                       float max_depth, const int* ip, const float* zp) {
    frame = frames + static_cast<size_t>(n) * H_ * W_;
    scale = frame_scale; lo = min_depth; hi = max_depth; H = H_; W = W_;
    window = ip != nullptr;
    if (window) {
      const int* q = ip + 8 * n;
      xstart = q[0]; ystart = q[1]; wb = q[2]; hb = q[3];
      zs = zp[2 * n]; ze = zp[2 * n + 1];
    } else {
      xstart = 0; ystart = 0; wb = W_; hb = H_; zs = 0.f; ze = 0.f;
    }
  }
  // pixel (r, c) of the image after getCrop's zero padding and z clamp
  __device__ __forceinline__ float crop_value(int r, int c) const {
    float v;
    if (window) {
      const int Y = ystart + r, X = xstart + c;
      v = 0.f;
      if (Y >= 0 && Y < H && X >= 0 && X < W) v = __fmul_rn(frame[static_cast<size_t>(Y) * W + X], scale);
      if (v != 0.f) {
        if (v < zs) v = zs;
        else if (v > ze) v = 0.f;
      }
    } else {
      v = __fmul_rn(frame[static_cast<size_t>(r) * W + c], scale);
    }
    return v;
  }
  // calculateCoM's range test
  __device__ __forceinline__ float com_value(float v) const {
    if (v < lo) v = 0.f;
    if (v > hi) v = 0.f;
    return v;
  }
};

// One block of numpy's tree (8 <= len <= 128 pixels starting at pixel `off`) by a group of 8 lanes; lane j is numpy's
// accumulator r[j].  Straight-line, predicated code: the lane's (up to) 16 pixels and its tail pixel are requested
// first, all loads in flight together, then added in numpy's order.  WINDOW: the image is a crop window (bounds test,
// z clamp), else the frame itself (pixel index = address).  POSLO: the near plane is positive (always, in the
// reference's use), which makes "in range" and "in the mask" the same test.  Needs an image at least 8 pixels wide and
// a block of at least 64 pixels (every block of an image of more than 128): the lane's first eight pixels always exist.
template <bool WINDOW, bool POSLO>
__device__ __forceinline__ float com_block_sum8(const ComImage& im, unsigned off, int rb, int cb, int len, int lane,
                                                unsigned gmask, unsigned& bpos, unsigned& bnz, unsigned& bsx,
                                                unsigned& bsy) {
  const int j = lane & 7, wb = im.wb;
  const int full = len & ~7, tail = len & 7;
  const int e0 = static_cast<int>(off) + j;
  // (rb, cb) = row / column of the block's first pixel (kept by the caller from block to block: no division here)
  int r0 = rb, c0 = cb + j;
  if (c0 >= wb) { c0 -= wb; ++r0; }
  auto fetch = [&](int e, int r, int c, bool on) -> float {
    if (!WINDOW) return on ? __fmul_rn(im.frame[e], im.scale) : 0.f;
    const int Y = im.ystart + r, X = im.xstart + c;
    const bool inb = on && static_cast<unsigned>(Y) < static_cast<unsigned>(im.H) &&
                     static_cast<unsigned>(X) < static_cast<unsigned>(im.W);
    float v = inb ? __fmul_rn(im.frame[static_cast<size_t>(Y) * im.W + X], im.scale) : 0.f;
    v = (v != 0.f && v < im.zs) ? im.zs : ((v > im.ze) ? 0.f : v);          // getCrop: clamp to the front face, drop the back
    return v;
  };
  // tail pixel of this lane (numpy adds the len % 8 last pixels one by one at the end)
  const int te = static_cast<int>(off) + full + j;
  int tr = rb, tc = cb + full + j;
  while (tc >= wb) { tc -= wb; ++tr; }
  float tv = fetch(te, tr, tc, j < tail);
  float v[16];
  {
    int r = r0, c = c0;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      v[k] = fetch(e0 + 8 * k, r, c, k < 8 || 8 * k < full);
      c += 8;
      if (c >= wb) { c -= wb; ++r; }
    }
  }
  float acc = 0.f;
  if (POSLO && wb >= 128) {
    // lo > 0: a pixel that passes the range test is positive, hence in the mask and non-zero -- one test, one mask.
    // Unused slots hold +0, fail the test and add +0 to a sum that is never -0 (its terms are +0, positive or NaN):
    // no predicates.  Only a NaN pixel breaks the rule (numpy's comparisons are false for it, so it stays in the
    // image: non-zero, but not > 0); it also turns the sum NaN, which is the cue to redo both masks exactly.
    unsigned mpos = 0u;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float raw = v[k];
      const bool out = (raw < im.lo) || (raw > im.hi);
      const float x = out ? 0.f : raw;
      if (!out) mpos |= 1u << k;
      acc = (k == 0) ? x : __fadd_rn(acc, x);
    }
    unsigned mnz = mpos;
    if (acc != acc) {
      mpos = 0u; mnz = 0u;
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const float raw = v[k];
        const bool out = (raw < im.lo) || (raw > im.hi);
        if (!out && raw > 0.f) mpos |= 1u << k;
        if (!out && raw != 0.f) mnz |= 1u << k;
      }
    }
    const unsigned cnt = __popc(mpos);
    const unsigned sumk = __popc(mpos & 0xAAAAu) + 2u * __popc(mpos & 0xCCCCu) + 4u * __popc(mpos & 0xF0F0u) +
                          8u * __popc(mpos & 0xFF00u);
    const int kw = (wb - c0 + 7) >> 3;                           // first pixel of this lane in the next row
    const unsigned after = kw < 16 ? __popc(mpos >> kw) : 0u;
    bpos += cnt;
    bnz += __popc(mnz);
    bsy += static_cast<unsigned>(r0) * cnt + after;
    bsx += static_cast<unsigned>(c0) * cnt + 8u * sumk - static_cast<unsigned>(wb) * after;
  } else if (wb >= 128) {
    // at most one row change inside the block: the mask's row / column sums follow from two 16-bit masks
    // (bit k = pixel k of this lane is in the mask / is non-zero) instead of three additions per pixel
    unsigned mpos = 0u, mnz = 0u;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float raw = v[k];                                    // (an unused slot holds 0: it counts nothing)
      const bool out = (raw < im.lo) || (raw > im.hi);           // calculateCoM's range test (a NaN stays)
      const float x = out ? 0.f : raw;
      if (!out && raw > 0.f) mpos |= 1u << k;
      if (!out && raw != 0.f) mnz |= 1u << k;
      acc = (k == 0) ? x : ((8 * k < full) ? __fadd_rn(acc, x) : acc);
    }
    const unsigned cnt = __popc(mpos);
    const unsigned sumk = __popc(mpos & 0xAAAAu) + 2u * __popc(mpos & 0xCCCCu) + 4u * __popc(mpos & 0xF0F0u) +
                          8u * __popc(mpos & 0xFF00u);
    const int kw = (wb - c0 + 7) >> 3;                           // first pixel of this lane in the next row
    const unsigned after = kw < 16 ? __popc(mpos >> kw) : 0u;
    bpos += cnt;
    bnz += __popc(mnz);
    bsy += static_cast<unsigned>(r0) * cnt + after;
    bsx += static_cast<unsigned>(c0) * cnt + 8u * sumk - static_cast<unsigned>(wb) * after;
  } else {
    int r = r0, c = c0;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const bool on = 8 * k < full;
      const float x = im.com_value(v[k]);
      const bool p = on && x > 0.f;
      bpos += p ? 1u : 0u; bsx += p ? static_cast<unsigned>(c) : 0u; bsy += p ? static_cast<unsigned>(r) : 0u;
      bnz += (on && x != 0.f) ? 1u : 0u;
      acc = (k == 0) ? x : (on ? __fadd_rn(acc, x) : acc);
      c += 8;
      if (c >= wb) { c -= wb; ++r; }
    }
  }
  // ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7)): IEEE addition commutes, so every lane gets the same bits
  acc = __fadd_rn(acc, __shfl_xor_sync(gmask, acc, 1));
  acc = __fadd_rn(acc, __shfl_xor_sync(gmask, acc, 2));
  acc = __fadd_rn(acc, __shfl_xor_sync(gmask, acc, 4));
  tv = im.com_value(tv);
  {
    const bool p = j < tail && tv > 0.f;
    bpos += p ? 1u : 0u; bsx += p ? static_cast<unsigned>(tc) : 0u; bsy += p ? static_cast<unsigned>(tr) : 0u;
    bnz += (j < tail && tv != 0.f) ? 1u : 0u;
  }
  for (int i = 0; i < tail; ++i)                                  // one by one, in order (uniform over the group)
    acc = __fadd_rn(acc, __shfl_sync(gmask, tv, (lane & 24) + i));
  return acc;
}

// any block, pixel by pixel with a division each: images narrower than 8 pixels or of at most 128 pixels in all
// (degenerate windows); every lane of the group computes the same sum and the same counts, the caller keeps lane 0's
struct ComSlowResult { float sum; unsigned pos, nz, sx, sy; };
__device__ __noinline__ ComSlowResult com_block_sum_slow(ComImage im, unsigned off, int len) {
  const int wb = im.wb;
  ComSlowResult out = {0.f, 0u, 0u, 0u, 0u};
  out.sum = np_block_sum([&](int i) {
    const int e = static_cast<int>(off) + i, r = e / wb, c = e - r * wb;
    const float v = im.com_value(im.crop_value(r, c));
    if (v > 0.f) { ++out.pos; out.sx += static_cast<unsigned>(c); out.sy += static_cast<unsigned>(r); }
    if (v != 0.f) ++out.nz;
    return v;
  }, len);
  return out;
}

template <bool WINDOW, bool POSLO>
__global__ void __launch_bounds__(256, WINDOW ? 2 : 4)
com_blocks_kernel(const float* __restrict__ frames, int H, int W, float frame_scale, float min_depth, float max_depth,
                  const int* __restrict__ ip, const float* __restrict__ zp, float* __restrict__ heap, unsigned heap_cap,
                  int leaves_log2, unsigned long long* __restrict__ stats) {
  const int n = blockIdx.y;
  ComImage im;
  im.load(frames, n, H, W, frame_scale, min_depth, max_depth, WINDOW ? ip : nullptr, zp);
  const long long npx64 = static_cast<long long>(im.wb) * im.hb;
  if (npx64 < 1 || npx64 > 0x7fffffffLL) return;
  const unsigned npx = static_cast<unsigned>(npx64);
  const int depth = np_pairwise_depth32(npx);
  const int L = depth > leaves_log2 ? depth - leaves_log2 : 0;
  if ((2ull << L) > heap_cap) return;
  float* vals = heap + static_cast<size_t>(n) * heap_cap;
  const int lane = threadIdx.x & 31, j = lane & 7;
  const unsigned gmask = 0xffu << (lane & 24);                    // the 8 lanes of this group
  // group w of 2^L owns the blocks below the level-L node with path w: ~2^leaves_log2 consecutive blocks, folded into
  // the node's sum as they come (np_host_worker_sum in devtools/np_reduce_host.cu is this loop on the host)
  const unsigned w = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  unsigned bpos = 0, bnz = 0;                                     // per lane: < 2^31 pixels per image
  unsigned long long sx = 0, sy = 0;
  unsigned no, nl, nid;
  if (w < (1u << L) && np_pairwise_worker_node(npx, L, w, &no, &nl, &nid)) {
    int rb = static_cast<int>(no / static_cast<unsigned>(im.wb));      // row / column of the next block's first pixel
    int cb = static_cast<int>(no - static_cast<unsigned>(rb) * static_cast<unsigned>(im.wb));
    NpSubtreeSum node;
    node.init();
    for (unsigned p = no; p < no + nl;) {
      unsigned off; int len, steps;
      np_pairwise_block_in(no, nl, nid, p, &off, &len, &steps);
      unsigned bsx = 0, bsy = 0;                                  // a block holds at most 128 pixels: 32 bits are plenty
      float res;
      if (len >= 64 && im.wb >= 8) {
        res = com_block_sum8<WINDOW, POSLO>(im, off, rb, cb, len, lane, gmask, bpos, bnz, bsx, bsy);
      } else {
        const ComSlowResult sr = com_block_sum_slow(im, off, len);
        res = sr.sum;
        if (j == 0) { bpos += sr.pos; bnz += sr.nz; bsx = sr.sx; bsy = sr.sy; }
      }
      node.push(res, steps);                                      // (every lane of the group holds the same bits)
      sx += bsx; sy += bsy;
      p += static_cast<unsigned>(len);
      cb += len;
      while (cb >= im.wb) { cb -= im.wb; ++rb; }
    }
    if (j == 0) vals[nid] = node.total();
  }
  unsigned long long pos = bpos, nz = bnz;
  for (int o = 16; o > 0; o >>= 1) {
    pos += __shfl_xor_sync(0xffffffffu, pos, o); nz += __shfl_xor_sync(0xffffffffu, nz, o);
    sx += __shfl_xor_sync(0xffffffffu, sx, o);   sy += __shfl_xor_sync(0xffffffffu, sy, o);
  }
  if (lane == 0 && (pos | nz | sx | sy)) {
    unsigned long long* st = stats + 4 * static_cast<size_t>(n);
    atomicAdd(st, pos); atomicAdd(st + 1, nz); atomicAdd(st + 2, sx); atomicAdd(st + 3, sy);
  }
}

// One CTA per frame: the sums of the level-L nodes (and of blocks above that level) -> the top L levels of numpy's
// tree, level by level in shared memory (2^(L+1) floats of dynamic shared memory), then the reference's few double
// operations in its order.
__global__ void __launch_bounds__(256)
com_finish_kernel(const float* __restrict__ frames, int H, int W, float frame_scale, float min_depth, float max_depth,
                  const int* __restrict__ ip, const float* __restrict__ zp, const float* __restrict__ heap,
                  unsigned heap_cap, int leaves_log2, const unsigned long long* __restrict__ stats,
                  double* __restrict__ coms, int* __restrict__ overflow) {
  extern __shared__ float s_vals[];
  const int n = blockIdx.x;
  const int tid = threadIdx.x, T = blockDim.x;
  ComImage im;
  im.load(frames, n, H, W, frame_scale, min_depth, max_depth, ip, zp);
  const long long npx = static_cast<long long>(im.wb) * im.hb;
  const int depth = (npx >= 1 && npx <= 0x7fffffffLL) ? np_pairwise_depth32(static_cast<unsigned>(npx)) : 0;
  const int L = depth > leaves_log2 ? depth - leaves_log2 : 0;
  if (npx < 1 || npx > 0x7fffffffLL || (2ull << L) > heap_cap) {
    if (tid == 0) {
      coms[3 * n] = coms[3 * n + 1] = coms[3 * n + 2] = nan("");
      overflow[n] = 1;
    }
    return;
  }
  const float* vals = heap + static_cast<size_t>(n) * heap_cap;
  // what the workers wrote: nodes of level L, and blocks that sit above it
  for (unsigned id = 1u + tid; id < (2u << L); id += T) {
    long long off, len;
    const bool level_L = (id >> L) != 0u;
    if (np_pairwise_node(npx, id, &off, &len) && (level_L || len <= kNpBlock)) s_vals[id] = vals[id];
  }
  __syncthreads();
  for (int level = L - 1; level >= 0; --level) {
    np_tree_level(npx, level, tid, T, s_vals);
    __syncthreads();
  }
  if (tid != 0) return;
  overflow[n] = 0;
  const unsigned long long* st = stats + 4 * static_cast<size_t>(n);
  const double cnt = static_cast<double>(st[0]), num = static_cast<double>(st[1]);
  double c0 = 0., c1 = 0., c2 = 0.;
  if (st[1] != 0ull) {
    // cc = sums / count of the mask; com = (cc[1] * num, cc[0] * num, dc.sum()) / num, every step a rounded double op
    c0 = __ddiv_rn(__dmul_rn(__ddiv_rn(static_cast<double>(st[2]), cnt), num), num);
    c1 = __ddiv_rn(__dmul_rn(__ddiv_rn(static_cast<double>(st[3]), cnt), num), num);
    c2 = __ddiv_rn(static_cast<double>(s_vals[1]), num);
  }
  if (im.window) {
    if (fabs(c0) <= 1e-8 && fabs(c1) <= 1e-8 && fabs(c2) <= 1e-8) {          // numpy.allclose(com, 0.)
      c2 = static_cast<double>(im.crop_value(im.hb / 2, im.wb / 2));
      if (fabs(c2) <= 1e-8) c2 = 300.;                                        // numpy.isclose(com[2], 0)
    }
    c0 = __dadd_rn(c0, static_cast<double>(im.xstart));
    c1 = __dadd_rn(c1, static_cast<double>(im.ystart));
  }
  coms[3 * n] = c0; coms[3 * n + 1] = c1; coms[3 * n + 2] = c2;
}

}  // namespace hgru
