// numpy's float32 summation order on the device.  The reference's host-side metrics and centre-of-mass code
// (pose_evaluation.py:10-88, tf_monkeydetector.py:73-90) reduce float32 arrays with numpy.sum / nanmean, whose
// result depends on the order of the additions: over a contiguous axis numpy adds PAIRWISE -- blocks of at most
// 128 elements summed with eight interleaved accumulators, combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), a
// remainder of n % 8 elements added one by one, and longer runs split recursively at n/2 rounded down to a multiple
// of 8 -- and over any other axis it adds element by element.  Restating that order here makes the device results
// bit-identical to the host's instead of "close".  (Checked against numpy itself by tests/test_np_reduce_order.py,
// which compiles these very functions for the host -- they are __host__ __device__ for that purpose only; the
// product calls them from kernels.)
#pragma once
#include <cuda_runtime.h>

// explicitly rounded operations on the device (no FMA contraction); on the host a lone add / divide cannot contract
#ifdef __CUDA_ARCH__
#define HGRU_NP_ADD(a, b) __fadd_rn((a), (b))
#define HGRU_NP_DDIV(a, b) __ddiv_rn((a), (b))
#else
#define HGRU_NP_ADD(a, b) ((a) + (b))
#define HGRU_NP_DDIV(a, b) ((a) / (b))
#endif

namespace hgru {

constexpr int kNpBlock = 128;   // numpy's PW_BLOCKSIZE

// left child's length of a pairwise node of n > 128 elements
__host__ __device__ inline long long np_pairwise_split(long long n) {
  long long n2 = n / 2;
  return n2 - (n2 % 8);
}

// depth of the pairwise tree (0: a single block)
__host__ __device__ inline int np_pairwise_depth(long long n) {
  int d = 0;
  while (n > kNpBlock) { n -= np_pairwise_split(n); ++d; }      // the right child is never the shorter one
  return d;
}

// one block (n <= 128) by ONE thread; `at(i)` yields element i
template <typename F>
__host__ __device__ inline float np_block_sum(F at, int n) {
  if (n < 8) {
    float res = 0.f;
    for (int i = 0; i < n; ++i) res = HGRU_NP_ADD(res, at(i));
    return res;
  }
  float r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = at(j);
  int i = 8;
  for (; i < n - (n % 8); i += 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = HGRU_NP_ADD(r[j], at(i + j));
  }
  float res = HGRU_NP_ADD(HGRU_NP_ADD(HGRU_NP_ADD(r[0], r[1]), HGRU_NP_ADD(r[2], r[3])),
                        HGRU_NP_ADD(HGRU_NP_ADD(r[4], r[5]), HGRU_NP_ADD(r[6], r[7])));
  for (; i < n; ++i) res = HGRU_NP_ADD(res, at(i));
  return res;
}

// a whole run of n elements by ONE thread (explicit stack instead of recursion); for the short reductions of the
// metrics (n = joints per frame, frames per batch)
template <typename F>
__host__ __device__ inline float np_pairwise_sum(F at, long long n) {
  if (n <= kNpBlock) return np_block_sum(at, static_cast<int>(n));
  long long off[40], len[40];
  char state[40];
  float val[40];
  int sp = 0, vp = 0;
  off[0] = 0; len[0] = n; state[0] = 0;
  while (sp >= 0) {
    const long long o = off[sp], l = len[sp];
    if (l <= kNpBlock) {
      val[vp++] = np_block_sum([&](int i) { return at(o + i); }, static_cast<int>(l));
      --sp;
    } else if (state[sp] == 0) {
      state[sp] = 1;
      ++sp; off[sp] = o; len[sp] = np_pairwise_split(l); state[sp] = 0;
    } else if (state[sp] == 1) {
      state[sp] = 2;
      const long long n2 = np_pairwise_split(l);
      ++sp; off[sp] = o + n2; len[sp] = l - n2; state[sp] = 0;
    } else {
      const float b = val[--vp], a = val[--vp];
      val[vp++] = HGRU_NP_ADD(a, b);
      --sp;
    }
  }
  return val[0];
}

// numpy.nanmean over a contiguous run: NaNs replaced by 0 and left out of the count, float32 total / count
template <typename F>
__host__ __device__ inline float np_nanmean(F at, long long n, bool skip_nan) {
  long long cnt = n;
  if (skip_nan) {
    cnt = 0;
    for (long long i = 0; i < n; ++i) { const float v = at(i); cnt += (v == v) ? 1 : 0; }
  }
  const float tot = np_pairwise_sum([&](long long i) { const float v = at(i); return (skip_nan && v != v) ? 0.f : v; }, n);
  // float32 / count goes through double in numpy; double division rounded to float32 equals the float32 quotient
  return static_cast<float>(HGRU_NP_DDIV(static_cast<double>(tot), static_cast<double>(cnt)));
}

// ---- the same tree, walked in parallel --------------------------------------------------------------------------
// For long runs (a whole depth frame) the blocks of the pairwise tree are summed by different threads and combined
// level by level.  Nodes are named by their heap index (root 1, children 2i and 2i + 1), so the partial sums live in a
// flat array of 2^(depth + 1) floats and no table of the tree is needed: both functions below re-derive a node from n.

// The node reached from the root along the path written in `id` (bits below the leading one, most significant first;
// 0 = left).  False when the path runs below a block, i.e. the tree has no such node.
__host__ __device__ inline bool np_pairwise_node(long long n, unsigned id, long long* off, long long* len) {
  int top = 31;
  while (top > 0 && !((id >> top) & 1u)) --top;
  long long o = 0, l = n;
  for (int b = top - 1; b >= 0; --b) {
    if (l <= kNpBlock) return false;
    const long long n2 = np_pairwise_split(l);
    if ((id >> b) & 1u) { o += n2; l -= n2; } else { l = n2; }
  }
  *off = o; *len = l;
  return true;
}

// The block that holds element p: its heap index, offset and length.  Blocks of a run longer than 128 elements are
// 64 to 128 elements long, so probing p = 0, 64, 128, ... meets every block, and a block is met FIRST at
// p == round_up(off, 64) -- the rule the kernels use to give each block to exactly one thread.
__host__ __device__ inline unsigned np_pairwise_block_at(long long n, long long p, long long* off, int* len) {
  unsigned id = 1;
  long long o = 0, l = n;
  while (l > kNpBlock) {
    const long long n2 = np_pairwise_split(l);
    if (p - o >= n2) { o += n2; l -= n2; id = 2 * id + 1; } else { l = n2; id = 2 * id; }
  }
  *off = o; *len = static_cast<int>(l);
  return id;
}
// the same walk in 32-bit unsigned arithmetic (n < 2^31): a shift and a mask per level instead of signed 64-bit
// division -- the bandwidth-bound kernel calls it once per block
__host__ __device__ inline unsigned np_pairwise_block_at32(unsigned n, unsigned p, unsigned* off, int* len) {
  unsigned id = 1, o = 0, l = n;
  while (l > static_cast<unsigned>(kNpBlock)) {
    const unsigned n2 = (l >> 1) & ~7u;
    if (p - o >= n2) { o += n2; l -= n2; id = 2 * id + 1; } else { l = n2; id = 2 * id; }
  }
  *off = o; *len = static_cast<int>(l);
  return id;
}
__host__ __device__ inline int np_pairwise_depth32(unsigned n) {
  int d = 0;
  while (n > static_cast<unsigned>(kNpBlock)) { n -= (n >> 1) & ~7u; ++d; }
  return d;
}
// Work split of the bandwidth-bound kernel: worker w of 2^L owns the subtree below the level-L node with path w (its
// blocks are contiguous in memory, and finding the next one is a walk of a few levels, not of the whole tree).  A
// block that sits ABOVE level L (ragged tree) goes to the worker whose remaining path bits are all zero.  False: this
// worker owns nothing.
__host__ __device__ inline bool np_pairwise_worker_node(unsigned n, int L, unsigned w, unsigned* off, unsigned* len,
                                                        unsigned* id) {
  unsigned o = 0, l = n, i = 1;
  for (int b = L - 1; b >= 0; --b) {
    if (l <= static_cast<unsigned>(kNpBlock)) {
      if (w & ((2u << b) - 1u)) return false;
      break;
    }
    const unsigned n2 = (l >> 1) & ~7u;
    if ((w >> b) & 1u) { o += n2; l -= n2; i = 2 * i + 1; } else { l = n2; i = 2 * i; }
  }
  *off = o; *len = l; *id = i;
  return true;
}
// the block that holds element p inside the node (o, l, id); *steps = its depth below that node
__host__ __device__ inline unsigned np_pairwise_block_in(unsigned o, unsigned l, unsigned id, unsigned p, unsigned* off,
                                                        int* len, int* steps) {
  int d = 0;
  while (l > static_cast<unsigned>(kNpBlock)) {
    const unsigned n2 = (l >> 1) & ~7u;
    if (p - o >= n2) { o += n2; l -= n2; id = 2 * id + 1; } else { l = n2; id = 2 * id; }
    ++d;
  }
  *off = o; *len = static_cast<int>(l); *steps = d;
  return id;
}
// A worker's block sums arrive in memory order = left to right in its subtree: two neighbours of equal depth are the
// two children of one node, so they are added on the spot and the sum moves one level up.  After the last block the
// stack holds the subtree's sum -- numpy's additions, in numpy's order, without storing the blocks.
struct NpSubtreeSum {
  float v[34];                 // a tree of 2^31 elements is 25 levels deep
  int d[34];
  int n;
  __host__ __device__ void init() { n = 0; }
  __host__ __device__ void push(float x, int depth) {
    v[n] = x; d[n] = depth; ++n;
    while (n >= 2 && d[n - 1] == d[n - 2]) {
      v[n - 2] = HGRU_NP_ADD(v[n - 2], v[n - 1]);
      --d[n - 2];
      --n;
    }
  }
  __host__ __device__ float total() const { return v[0]; }      // valid once every block of the subtree is in
};
__host__ __device__ inline bool np_pairwise_first_probe(long long off, long long p) {
  return (off + 63) / 64 * 64 == p;
}

// Thread `tid` of `T`: sum the blocks this thread owns into vals[heap index].  Every element is read exactly once
// (callers may count things as a side effect of `at`).
template <typename F>
__host__ __device__ inline void np_tree_blocks(F at, long long n, int tid, int T, float* vals) {
  for (long long p = 64ll * tid; p < n; p += 64ll * T) {
    long long off; int len;
    const unsigned id = np_pairwise_block_at(n, p, &off, &len);
    if (!np_pairwise_first_probe(off, p)) continue;
    vals[id] = np_block_sum([&](int i) { return at(off + i); }, len);
  }
}
// Thread `tid` of `T`: the inner nodes of one level (root = level 0), children first -- call for level = depth - 1 ... 0
// with a barrier in between; the total ends up in vals[1].
__host__ __device__ inline void np_tree_level(long long n, int level, int tid, int T, float* vals) {
  const unsigned first = 1u << level;
  for (unsigned id = first + tid; id < 2 * first; id += T) {
    long long off, len;
    if (np_pairwise_node(n, id, &off, &len) && len > kNpBlock) vals[id] = HGRU_NP_ADD(vals[2 * id], vals[2 * id + 1]);
  }
}

}  // namespace hgru
