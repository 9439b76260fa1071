"""GPU (B200): the stages on either side of the pose network that round 2 widened last -- the whole metric file of
the reference (pose_evaluation.py:10-88) and the centre-of-mass estimation / refinement of the crop stage
(tf_monkeydetector.py:73-90, 292-333) -- against golden vectors produced by the reference's own code.
Everything here is bit-exact: the kernels add float32 numbers in numpy's own order (csrc/np_reduce.cuh)."""
import os
import warnings

import numpy as np
import pytest
import torch

from oracle import post_oracle_np as post
from tests.golden.make_golden_metrics import nan_labels

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _dev(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("tag", ["a", "b"])
def test_metric_file_bit_exact_against_reference_golden(tag):
    from monkey_pose_b200 import pose_evaluation as pe
    z = np.load(os.path.join(GOLDEN, "metrics_ref.npz"))
    res = z[tag + "_results"]
    for name, lab in (("", z[tag + "_labels"]), ("nan_", nan_labels(z[tag + "_labels"]))):
        k = tag + "_" + name
        L, R = _dev(lab), _dev(res)
        assert np.float32(pe.getMeanError_np(L, R)) == z[k + "mean_np"]
        assert np.float32(pe.getMaxError_np(L, R)) == z[k + "max_np"]
        assert np.array_equal(pe.getMean_np(L, R), z[k + "getMean_np"])
        assert np.float32(pe.getMean_np(_dev(lab[:, 0, :]), _dev(res[:, 0, :]))) == z[k + "getMean_np_rank2"]
        for i, d in enumerate(z[tag + "_dists"]):
            assert pe.getNumFramesWithinMaxDist(L, R, d) == z[k + "within_max"][i]
            assert pe.getNumFramesWithinMeanDist(L, R, d) == z[k + "within_mean"][i]
        jm = np.array([pe.getJointMeanError(L, R, j) for j in range(lab.shape[1])], np.float32)
        assert np.array_equal(jm, z[k + "joint_mean"])
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            assert np.array_equal(pe.joint_error_matrix(L, R).cpu().numpy(), post.joint_errors(lab, res), equal_nan=True)
    # the TensorFlow variants: device tensors, the reference's formulas (TensorFlow's own reduction order is not
    # reproducible, so the contract is a float32 tolerance; against numpy's order they are in fact exact)
    L, R = _dev(z[tag + "_labels"]), _dev(res)
    t = pe.getMeanError_train(L, R)
    assert t.is_cuda and t.dim() == 0
    assert abs(float(t) - float(z[tag + "_train"])) <= 1e-6 * float(z[tag + "_train"])
    assert np.allclose(pe.getMeanError(L, R).cpu().numpy(), z[tag + "_getMeanError"], rtol=1e-6, atol=0)
    assert np.allclose(pe.getMeanErrors_N(L, R).cpu().numpy(), z[tag + "_getMeanErrors_N"], rtol=1e-6, atol=0)
    assert float(pe.getMaxError(L, R)) == float(z[tag + "_getMaxError"])
    # a NaN poisons the TensorFlow variants and is skipped by the numpy ones
    Ln = _dev(nan_labels(z[tag + "_labels"]))
    assert np.isnan(float(pe.getMeanError_train(Ln, R))) and np.isfinite(pe.getMeanError_np(Ln, R))
    with pytest.raises(AssertionError):
        pe.getMeanError_train(L, R[:, :5])
    with pytest.raises(RuntimeError):
        pe.getMeanError_np(z[tag + "_labels"], res)          # host arrays: there is no CPU fallback


def test_metrics_at_batch_size_equal_numpy():
    """Batch 4096 x 23 joints (more frames than one pairwise block holds, unaligned splits): device == numpy."""
    from monkey_pose_b200 import pose_evaluation as pe
    rng = np.random.default_rng(5)
    for n in (129, 1000, 4096):
        lab = rng.uniform(-300, 300, size=(n, 23, 3)).astype(np.float32)
        res = (lab + rng.normal(0, 9, size=lab.shape)).astype(np.float32)
        L, R = _dev(lab), _dev(res)
        assert np.float32(pe.getMeanError_np(L, R)) == np.float32(post.mean_error(lab, res))
        assert np.array_equal(pe.getMean_np(L, R), post.mean_axis1(lab, res))
        assert pe.getNumFramesWithinMeanDist(L, R, 14.0) == post.frames_within_mean_dist(lab, res, 14.0)
        assert np.float32(pe.getJointMeanError(L, R, 11)) == np.float32(post.joint_mean_error(lab, res, 11))
