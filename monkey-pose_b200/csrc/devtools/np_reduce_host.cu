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

// the kernel's work split: worker w of 2^L owns the subtree below the level-L node with path w, sums its blocks
// (np_block_sum) and folds them into the subtree's sum as they come (NpSubtreeSum); the top L levels are then combined
// by heap index.  Returns the total; *visited = blocks seen, -1 if a block is visited twice, skipped, out of order, or
// disagrees with the 64-bit walk.
float np_host_worker_sum(const float* a, long long n, int L, long long* visited_out) {
  const unsigned npx = static_cast<unsigned>(n);
  long long visited = 0, covered = 0;
  bool ok = hgru::np_pairwise_depth32(npx) == hgru::np_pairwise_depth(n);
  std::vector<float> vals(static_cast<size_t>(2) << L, -12345.f);
  for (unsigned w = 0; w < (1u << L); ++w) {
    unsigned no, nl, nid;
    if (!hgru::np_pairwise_worker_node(npx, L, w, &no, &nl, &nid)) continue;
    if (no != covered) ok = false;                   // workers own consecutive slices
    hgru::NpSubtreeSum st;
    st.init();
    for (unsigned p = no; p < no + nl;) {
      unsigned off; int len, steps;
      const unsigned id = hgru::np_pairwise_block_in(no, nl, nid, p, &off, &len, &steps);
      long long off64; int len64;
      if (hgru::np_pairwise_block_at(n, p, &off64, &len64) != id || off64 != off || len64 != len || off != p) ok = false;
      st.push(hgru::np_block_sum([&](int i) { return a[off + i]; }, len), steps);
      covered += len;
      ++visited;
      p += len;
    }
    if (st.n != 1) ok = false;
    vals[nid] = st.total();
  }
  for (int level = L - 1; level >= 0; --level) {
    for (unsigned id = 1u << level; id < (2u << level); ++id) {
      long long off, len;
      // inner nodes above the workers' level: both children were written (by a worker, or by this loop)
      if (hgru::np_pairwise_node(n, id, &off, &len) && len > hgru::kNpBlock) vals[id] = vals[2 * id] + vals[2 * id + 1];
    }
  }
  if (visited_out) *visited_out = (ok && covered == n) ? visited : -1;
  return vals[1];
}

}  // extern "C"
