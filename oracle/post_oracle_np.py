"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the step right after the pose network:
de-normalisation x cube[2]/2 (train_cnn_networks_hgru.py:293-296), getAbsoluteCoordinates
(tf_monkeydetector.py:387-391; uvdtoxyz :138-160, xyztouvd :116-136) and the error metrics
getMeanError_np / getMaxError_np (pose_evaluation.py:10-23).  Pinned against tests/golden/post_ref.npz
(reference source executed by tests/golden/make_golden_post.py).  Never imported by the product."""
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
