"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the step right after the pose network:
de-normalisation x cube[2]/2 (train_cnn_networks_hgru.py:293-296), getAbsoluteCoordinates
(tf_monkeydetector.py:387-391; uvdtoxyz :138-160, xyztouvd :116-136) and the error metrics
getMeanError_np / getMaxError_np and the rest of the metric file (pose_evaluation.py:10-88).  Pinned against
tests/golden/post_ref.npz and metrics_ref.npz (reference source executed by tests/golden/make_golden_post.py and
make_golden_metrics.py).  Never imported by the product."""
import numpy as np


def uvdtoxyz(com_uvd, fx, fy, ux, uy):
    """One (u, v, d): float64 arithmetic stored into a float32 triple (tf_monkeydetector.py:146-150)."""
    c = np.asarray(com_uvd, np.float64)
    out = np.zeros((3,), np.float32)
    out[0] = (ux - c[0]) * c[2] / (-fx)
    out[1] = (c[1] - uy) * c[2] / (-fy)
    out[2] = -c[2]
    return out


def xyztouvd(jnts_xyz, fx, fy, ux, uy):
    """[J,3] float32 -> [J,3] float32, elementwise float32 arithmetic (tf_monkeydetector.py:126-135)."""
    j = np.asarray(jnts_xyz, np.float32)
    out = np.zeros_like(j)
    f32 = np.float32
    for i in range(j.shape[0]):
        if j[i, 2] == 0.:
            out[i, 0], out[i, 1] = ux, uy
            continue
        out[i, 0] = f32(ux) - j[i, 0] / j[i, 2] * f32(fx)
        out[i, 1] = j[i, 1] / j[i, 2] * f32(fy) + f32(uy)
        out[i, 2] = -j[i, 2]
    return out


def absolute_coordinates(out_put, coms, cam, scale=600.0):
    """out_put [N,3J] normalised -> (xyz [N,J,3], uvd [N,J,3]) float32."""
    fx, fy, ux, uy = cam
    n = out_put.shape[0]
    rel = np.reshape(np.asarray(out_put, np.float32), (n, -1, 3)) * np.float32(scale)
    xyz = np.zeros_like(rel)
    uvd = np.zeros_like(rel)
    for i in range(n):
        xyz[i] = rel[i] + uvdtoxyz(coms[i], fx, fy, ux, uy)
        uvd[i] = xyztouvd(xyz[i], fx, fy, ux, uy)
    return xyz, uvd


def mean_error(labels, results):
    """getMeanError_np (pose_evaluation.py:10-15)."""
    return np.nanmean(np.nanmean(np.sqrt(np.square(labels - results).sum(axis=2)), axis=1))


def max_error(labels, results):
    """getMaxError_np (pose_evaluation.py:18-23)."""
    return np.nanmax(np.sqrt(np.square(labels - results).sum(axis=2)))


# ---- the rest of the metric file (pose_evaluation.py:26-88); float32 in, float32 arithmetic, numpy's own order ------
def joint_errors(labels, results):
    """The expression every metric reduces: per-joint Euclidean error [N,J]."""
    return np.sqrt(np.square(labels - results).sum(axis=2))


def mean_axis1(labels, results, skip_nan=True):
    """getMean_np (:26-28; nan-skipping) / getMeanError (:38-44; TensorFlow, NaN propagates)."""
    rows = np.sqrt(np.square(labels - results).sum(axis=1))
    return np.nanmean(rows, axis=0) if skip_nan else np.mean(rows, axis=0, dtype=np.float32)


def mean_error_train(labels, results):
    """getMeanError_train (:30-36): reduce_mean over frames of reduce_mean over joints."""
    e = joint_errors(labels, results)
    return np.mean(np.mean(e, axis=1, dtype=np.float32), dtype=np.float32)


def mean_errors_n(labels, results):
    """getMeanErrors_N (:46-52): per-frame mean."""
    return np.mean(joint_errors(labels, results), axis=1, dtype=np.float32)


def max_error_tf(labels, results):
    """getMaxError (:54-60)."""
    return np.max(joint_errors(labels, results))


def frames_within_max_dist(labels, results, dist):
    """getNumFramesWithinMaxDist (:63-69)."""
    return int((np.nanmax(joint_errors(labels, results), axis=1) <= dist).sum())


def frames_within_mean_dist(labels, results, dist):
    """getNumFramesWithinMeanDist (:72-78)."""
    return int((np.nanmean(joint_errors(labels, results), axis=1) <= dist).sum())


def joint_mean_error(labels, results, joint_id):
    """getJointMeanError (:81-88)."""
    return np.nanmean(np.sqrt(np.square(labels[:, joint_id, :] - results[:, joint_id, :]).sum(axis=1)))
