// Host build of csrc/np_reduce.cuh for tests/test_np_reduce_order.py: the same functions the kernels call, compiled
// for the CPU, so that numpy itself can be asked whether the addition order is its own.  Not part of the product.
//   nvcc -shared -Xcompiler -fPIC -o libnp_reduce_host.so np_reduce_host.cu
#include "../np_reduce.cuh"

#include <vector>

extern "C" {

float np_host_pairwise_sum(const float* a, long long n) {
  return hgru::np_pairwise_sum([&](long long i) { return a[i]; }, n);
}

float np_host_nanmean(const float* a, long long n, long long stride, int skip_nan) {
  return hgru::np_nanmean([&](long long i) { return a[i * stride]; }, n, skip_nan != 0);
}

// the kernels' parallel scheme run by T sequential "threads": blocks by heap index, then level by level
float np_host_tree_sum(const float* a, long long n, int T, long long* reads) {
  const int depth = hgru::np_pairwise_depth(n);
  std::vector<float> vals(static_cast<size_t>(2) << depth, -12345.f);
  long long cnt = 0;
  for (int tid = 0; tid < T; ++tid)
    hgru::np_tree_blocks([&](long long i) { ++cnt; return a[i]; }, n, tid, T, vals.data());
  for (int level = depth - 1; level >= 0; --level)
    for (int tid = 0; tid < T; ++tid) hgru::np_tree_level(n, level, tid, T, vals.data());
  if (reads) *reads = cnt;
  return vals[1];
}

int np_host_depth(long long n) { return hgru::np_pairwise_depth(n); }

// the kernel's work split: worker w of 2^L owns the subtree below the level-L node with path w; returns the number
// of blocks visited, fails (-1) if a block is visited twice, skipped, out of order, or disagrees with the 64-bit walk
long long np_host_worker_walk(long long n, int L) {
  const unsigned npx = static_cast<unsigned>(n);
  if (hgru::np_pairwise_depth32(npx) != hgru::np_pairwise_depth(n)) return -1;
  long long visited = 0, covered = 0;
  for (unsigned w = 0; w < (1u << L); ++w) {
    unsigned no, nl, nid;
    if (!hgru::np_pairwise_worker_node(npx, L, w, &no, &nl, &nid)) continue;
    if (no != covered) return -1;                    // workers own consecutive slices
    for (unsigned p = no; p < no + nl;) {
      unsigned off; int len;
      const unsigned id = hgru::np_pairwise_block_in(no, nl, nid, p, &off, &len);
      long long off64; int len64;
      if (hgru::np_pairwise_block_at(n, p, &off64, &len64) != id || off64 != off || len64 != len || off != p) return -1;
      covered += len;
      ++visited;
      p += len;
    }
  }
  return covered == n ? visited : -1;
}

}  // extern "C"
