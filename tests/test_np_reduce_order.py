"""CPU: the addition order of csrc/np_reduce.cuh IS numpy's.  The header's functions are __host__ __device__; this
test compiles them for the host (csrc/devtools/np_reduce_host.cu, plain g++) and compares with numpy's own float32
sums bit for bit -- the one-thread pairwise sum the metric kernels use, numpy.nanmean, and the parallel scheme of the
centre-of-mass kernel (blocks by heap index, combined level by level) for any number of threads."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "monkey-pose_b200", "csrc", "devtools", "np_reduce_host.cu")
CUDA_INC = os.environ.get("CUDA_HOME", "/usr/local/cuda") + "/include"


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    if shutil.which("g++") is None or not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("needs g++ and the CUDA headers")
    so = str(tmp_path_factory.mktemp("np_reduce") / "libnp_reduce_host.so")
    subprocess.check_call(["g++", "-x", "c++", "-O2", "-std=c++17", "-ffp-contract=off", "-I", CUDA_INC, "-shared",
                           "-fPIC", "-o", so, SRC])
    lib = ctypes.CDLL(so)
    fp = ctypes.POINTER(ctypes.c_float)
    lib.np_host_pairwise_sum.restype = ctypes.c_float
    lib.np_host_pairwise_sum.argtypes = [fp, ctypes.c_longlong]
    lib.np_host_nanmean.restype = ctypes.c_float
    lib.np_host_nanmean.argtypes = [fp, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_int]
    lib.np_host_tree_sum.restype = ctypes.c_float
    lib.np_host_tree_sum.argtypes = [fp, ctypes.c_longlong, ctypes.c_int, ctypes.POINTER(ctypes.c_longlong)]
    lib.np_host_depth.restype = ctypes.c_int
    lib.np_host_depth.argtypes = [ctypes.c_longlong]
    lib.np_host_worker_sum.restype = ctypes.c_float
    lib.np_host_worker_sum.argtypes = [fp, ctypes.c_longlong, ctypes.c_int, ctypes.POINTER(ctypes.c_longlong)]
    return lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


SIZES = [1, 2, 7, 8, 9, 15, 16, 23, 36, 63, 64, 65, 127, 128, 129, 130, 136, 191, 192, 193, 255, 256, 257, 1000, 1023,
         1024, 1025, 4095, 4096, 4097, 12345, 65536, 100003, 424 * 512, 640 * 480 + 3]


def test_one_thread_pairwise_sum_is_numpy_sum(host):
    rng = np.random.default_rng(0)
    for n in SIZES:
        for scale in (1.0, 1e4):
            a = (rng.standard_normal(n) * scale).astype(np.float32)
            got = np.float32(host.np_host_pairwise_sum(_ptr(a), n))
            assert got == a.sum(), (n, scale)


def test_parallel_tree_sum_is_numpy_sum_for_any_thread_count(host):
    rng = np.random.default_rng(1)
    for n in SIZES:
        a = (rng.uniform(0, 3000, n) * (rng.uniform(size=n) < 0.7)).astype(np.float32)   # depth-like: many zeros
        want = a.sum()
        for T in (1, 3, 32, 1024):
            reads = ctypes.c_longlong(0)
            got = np.float32(host.np_host_tree_sum(_ptr(a), n, T, ctypes.byref(reads)))
            assert got == want, (n, T)
            assert reads.value == n, "every element is read exactly once"
    # a 2-D C-contiguous image sums like its flattened self (what dc.sum() does, tf_monkeydetector.py:85)
    img = (rng.uniform(0, 3000, (424, 512))).astype(np.float32)
    assert np.float32(host.np_host_tree_sum(_ptr(img), img.size, 1024, None)) == img.sum()


def test_tree_depth_bounds_the_heap(host):
    for n in SIZES + [2 ** 31 - 1]:
        d = host.np_host_depth(n)
        assert (n <= 128) == (d == 0)
        assert 64 * 2 ** d <= max(n, 64) or d == 0 or n > 64 * 2 ** (d - 1)    # blocks hold 64..128 elements
        assert n <= 128 * 2 ** d


def test_nanmean_matches_numpy(host):
    rng = np.random.default_rng(2)
    for n in (1, 5, 23, 36, 129, 256, 1000, 5000):
        a = (rng.standard_normal(n) * 50).astype(np.float32)
        assert np.float32(host.np_host_nanmean(_ptr(a), n, 1, 1)) == np.nanmean(a)
        assert np.float32(host.np_host_nanmean(_ptr(a), n, 1, 0)) == np.mean(a)
        b = a.copy()
        b[rng.uniform(size=n) < 0.2] = np.nan
        if np.isnan(b).all():
            continue
        assert np.float32(host.np_host_nanmean(_ptr(b), n, 1, 1)) == np.nanmean(b)
        assert np.isnan(host.np_host_nanmean(_ptr(b), n, 1, 0)) == np.isnan(b).any()
    # a strided column (getJointMeanError: err[:, j] gathered into a contiguous vector first)
    m = (rng.standard_normal((300, 23)) * 50).astype(np.float32)
    for j in (0, 7, 22):
        got = np.float32(host.np_host_nanmean(_ptr(m[:, j:]), 300, 23, 1))
        assert got == np.nanmean(np.ascontiguousarray(m[:, j]))
    allnan = np.full(4, np.nan, np.float32)
    assert np.isnan(host.np_host_nanmean(_ptr(allnan), 4, 1, 1))


def test_worker_split_sums_like_numpy(host):
    """The centre-of-mass kernel gives worker w of 2^L the subtree below the level-L node with path w (32-bit walk),
    folds the block sums into the subtree's sum as they come and combines the top L levels by heap index: every block
    exactly once, in memory order, same blocks as the 64-bit walk, and the total is numpy's -- also when L is deeper
    than parts of the tree (blocks above level L) or than all of it."""
    rng = np.random.default_rng(3)
    for n in SIZES + [128 * 2 ** 10, 128 * 2 ** 10 + 1, 129 * 2 ** 7 + 8, 1460 * 1461]:
        a = (rng.uniform(0, 3000, n) * (rng.uniform(size=n) < 0.7)).astype(np.float32)
        want = a.sum()
        depth = host.np_host_depth(n)
        for L in sorted({0, 1, max(depth - 4, 0), max(depth - 2, 0), depth, depth + 1, depth + 3}):
            visited = ctypes.c_longlong(0)
            got = np.float32(host.np_host_worker_sum(_ptr(a), n, L, ctypes.byref(visited)))
            assert visited.value >= 1, (n, L)
            assert got == want, (n, L)
            if n > 128:
                assert n / 128 <= visited.value <= n / 64


def test_random_sizes_and_splits_property(host):
    """Property test (hypothesis): for any length, any values and any worker level the device formulation's total is
    numpy's float32 sum."""
    hyp = pytest.importorskip("hypothesis")
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(n=st.integers(1, 40000), level=st.integers(0, 12), seed=st.integers(0, 2 ** 31 - 1),
           scale=st.sampled_from([1e-3, 1.0, 3e4]))
    def check(n, level, seed, scale):
        a = (np.random.default_rng(seed).standard_normal(n) * scale).astype(np.float32)
        visited = ctypes.c_longlong(0)
        got = np.float32(host.np_host_worker_sum(_ptr(a), n, level, ctypes.byref(visited)))
        assert visited.value >= 1
        assert got == a.sum()
        assert np.float32(host.np_host_pairwise_sum(_ptr(a), n)) == a.sum()

    check()
