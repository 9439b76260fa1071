"""Device versions of the reference's error metrics (pose_evaluation.py:10-88): same names and argument meaning,
torch CUDA tensors in.

The numpy functions of the reference (`getMeanError_np`, `getMaxError_np`, `getMean_np`, `getNumFramesWithinMaxDist`,
`getNumFramesWithinMeanDist`, `getJointMeanError`) return host values (Python floats / ints, a numpy vector for
`getMean_np` on rank-3 inputs), skip NaN joints like `numpy.nanmean` / `nanmax`, and -- for float32 inputs -- are
bit-identical to the reference's results: the kernels add in numpy's own float32 order (`csrc/np_reduce.cuh`).
The TensorFlow functions of the reference (`getMeanError_train`, `getMeanError`, `getMeanErrors_N`, `getMaxError`)
return CUDA tensors, stay on the stream without a host synchronisation, and let a NaN propagate like
`tf.reduce_mean` / `tf.reduce_max`.  Inputs are converted to float32 (the reference's graph dtype).  No CPU fallback."""
import torch

from . import _lib
from .hgru_module import _stream


def _pair(labels, results, min_dim=3):
    if not (torch.is_tensor(labels) and torch.is_tensor(results) and labels.is_cuda and results.is_cuda):
        raise RuntimeError("labels / results must be torch CUDA tensors (no CPU fallback)")
    # the reference's TF variants assert this; numpy would broadcast, which no caller relies on
    assert labels.shape == results.shape
    if labels.dim() < min_dim:
        raise ValueError("labels / results must have at least %d dimensions" % min_dim)
    return labels.to(torch.float32).contiguous(), results.to(torch.float32).contiguous()


class _Stats(object):
    """One pass of joint_error_stats_forward: err [N,J], frame_mean [N], frame_max [N], joint_mean [J], summary [2]."""

    def __init__(self, labels, results, skip_nan):
        a, b = _pair(labels, results)
        if a.dim() != 3 or a.shape[2] != 3:
            raise ValueError("labels / results must both be [N,J,3]")
        N, J = int(a.shape[0]), int(a.shape[1])
        dev = a.device
        self.N, self.J = N, J
        self.err = torch.empty((N, J), device=dev, dtype=torch.float32)
        self.frame_mean = torch.empty(N, device=dev, dtype=torch.float32)
        self.frame_max = torch.empty(N, device=dev, dtype=torch.float32)
        self.joint_mean = torch.empty(J, device=dev, dtype=torch.float32)
        self.summary = torch.empty(2, device=dev, dtype=torch.float32)
        _lib.check(_lib.load().joint_error_stats_forward(
            a.data_ptr(), b.data_ptr(), N, J, 1 if skip_nan else 0, self.err.data_ptr(), self.frame_mean.data_ptr(),
            self.frame_max.data_ptr(), self.joint_mean.data_ptr(), self.summary.data_ptr(), _stream()),
            "joint_error_stats_forward")

    def count_within(self, stat, dist):
        count = torch.empty(1, device=stat.device, dtype=torch.int32)
        _lib.check(_lib.load().joint_error_count_within_forward(stat.data_ptr(), self.N, float(dist), count.data_ptr(),
                                                                _stream()), "joint_error_count_within_forward")
        return int(count.item())


def _axis1_mean(labels, results, skip_nan):
    a, b = _pair(labels, results, min_dim=2)
    if a.dim() not in (2, 3):
        raise ValueError("labels / results must be [N,M] or [N,M,C]")
    N, M = int(a.shape[0]), int(a.shape[1])
    C = int(a.shape[2]) if a.dim() == 3 else 1
    rows = torch.empty((N, C), device=a.device, dtype=torch.float32)
    out = torch.empty(C, device=a.device, dtype=torch.float32)
    _lib.check(_lib.load().axis1_error_mean_forward(a.data_ptr(), b.data_ptr(), N, M, C, 1 if skip_nan else 0,
                                                    rows.data_ptr(), out.data_ptr(), _stream()),
               "axis1_error_mean_forward")
    return out if a.dim() == 3 else out[0]


# ---- numpy functions of the reference: host values, NaNs skipped ----------------------------------------------------
def getMeanError_np(labels, results):
    """Average error over all joints, averaged over the sequence (pose_evaluation.py:10-15)."""
    return float(_Stats(labels, results, True).summary[0].item())


def getMaxError_np(labels, results):
    """Maximum error over all joints (pose_evaluation.py:18-23)."""
    return float(_Stats(labels, results, True).summary[1].item())


def getMean_np(labels, results):
    """nanmean over axis 0 of sqrt(square(labels - results).sum(axis=1)) (pose_evaluation.py:26-28): a float for
    [N,3] inputs, a numpy vector [C] for [N,M,C] inputs."""
    out = _axis1_mean(labels, results, True)
    return float(out.item()) if out.dim() == 0 else out.cpu().numpy()


def getNumFramesWithinMaxDist(labels, results, dist):
    """Number of frames whose worst joint is within dist mm (pose_evaluation.py:63-69)."""
    s = _Stats(labels, results, True)
    return s.count_within(s.frame_max, dist)


def getNumFramesWithinMeanDist(labels, results, dist):
    """Number of frames whose mean joint error is within dist mm (pose_evaluation.py:72-78)."""
    s = _Stats(labels, results, True)
    return s.count_within(s.frame_mean, dist)


def getJointMeanError(labels, results, jointID):
    """Error of one joint, averaged over the sequence (pose_evaluation.py:81-88)."""
    s = _Stats(labels, results, True)
    if not -s.J <= int(jointID) < s.J:
        raise IndexError("index %d is out of bounds for axis 1 with size %d" % (int(jointID), s.J))
    return float(s.joint_mean[int(jointID)].item())


# ---- TensorFlow functions of the reference: device tensors, NaNs propagate ------------------------------------------
def getMeanError_train(labels, results):
    """reduce_mean over frames of reduce_mean over joints of the per-joint error (pose_evaluation.py:30-36; the
    train / validation error of train_cnn_networks_hgru.py:154-156, 171-173).  0-dim CUDA tensor."""
    return _Stats(labels, results, False).summary[0]


def getMeanError(labels, results):
    """reduce_mean over axis 0 of sqrt(reduce_sum(square(labels - results), 1)) (pose_evaluation.py:38-44)."""
    return _axis1_mean(labels, results, False)


def getMeanErrors_N(labels, results):
    """Per-frame mean joint error [N] (pose_evaluation.py:46-52)."""
    return _Stats(labels, results, False).frame_mean


def getMaxError(labels, results):
    """Maximum error over all joints (pose_evaluation.py:54-60).  0-dim CUDA tensor."""
    return _Stats(labels, results, False).summary[1]


def joint_error_matrix(labels, results):
    """The per-joint error matrix [N,J] every metric above reduces (not a reference function)."""
    return _Stats(labels, results, True).err
