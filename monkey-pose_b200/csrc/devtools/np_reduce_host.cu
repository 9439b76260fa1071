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

}  // extern "C"
