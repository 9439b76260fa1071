"""Device versions of the reference's numpy error metrics (pose_evaluation.py:10-23): same names, torch
CUDA tensors [N,J,3] in millimetres in, Python floats out.  One small kernel pair (per-frame reduction +
final reduction); NaNs are skipped exactly like numpy.nanmean / nanmax."""
import torch

from . import _lib
from .hgru_module import _stream


def _errors(labels, results):
    if not (torch.is_tensor(labels) and torch.is_tensor(results) and labels.is_cuda and results.is_cuda):
        raise RuntimeError("labels / results must be torch CUDA tensors (no CPU fallback)")
    if labels.shape != results.shape or labels.dim() != 3 or labels.shape[2] != 3:
        raise ValueError("labels / results must both be [N,J,3]")
    a = labels.to(torch.float32).contiguous()
    b = results.to(torch.float32).contiguous()
    N, J = int(a.shape[0]), int(a.shape[1])
    ws_mean = torch.empty(N, device=a.device, dtype=torch.float64)
    ws_max = torch.empty(N, device=a.device, dtype=torch.float32)
    res = torch.empty(2, device=a.device, dtype=torch.float64)
    _lib.check(_lib.load().joint_error_forward(a.data_ptr(), b.data_ptr(), N, J, ws_mean.data_ptr(),
                                               ws_max.data_ptr(), res.data_ptr(), _stream()), "joint_error_forward")
    r = res.cpu()
    return float(r[0]), float(r[1])


def getMeanError_np(labels, results):
    """Average error over all joints, averaged over the sequence (pose_evaluation.py:10-15)."""
    return _errors(labels, results)[0]


def getMaxError_np(labels, results):
    """Maximum error over all joints (pose_evaluation.py:18-23)."""
    return _errors(labels, results)[1]
