"""Golden vectors for the camera-model methods of `tfMonkeyDetector` that stay on the host (a handful of scalars per
frame): `xyztouvd_np` (tf_monkeydetector.py:116-136), `uvdtoxyz` (:138-160), `calcCoMRenders` (:185-191),
`comToBounds` (:193-206), `transformPoint2D` (:367-370), `getRelativeCoordinates` (:372-385),
`getAbsoluteCoordinates` (:387-391) and the TensorFlow-graph `calculateCoMfrom3DJoints` / `xyztouvd` (:66-71, :92-114;
run against a numpy stand-in for tf.reduce_mean / tf.stack), produced by EXECUTING THE REFERENCE'S OWN class.

    python tests/golden/make_golden_detector_host.py        (build container only; needs /root/reference)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden_crop import load_detector_module  # noqa: E402
from make_golden_com import CAMERA  # noqa: E402


def main():
    mod = load_detector_module()
    tf = sys.modules["tensorflow"]
    tf.reduce_mean = lambda x, axis=None: np.mean(x, axis=axis)
    tf.stack = lambda xs, axis=0: np.stack(xs, axis=axis)
    md = mod.tfMonkeyDetector(*CAMERA)
    rng = np.random.default_rng(5)
    J = 23
    jx = np.stack([rng.uniform(-400, 400, J), rng.uniform(-300, 300, J), -rng.uniform(900, 2600, J)], 1).astype(np.float32)
    jx[3, 2] = 0.0                                            # a joint on the camera plane: the z == 0 branch
    com_uvd = np.array([231.5, 187.25, 1830.0])
    M = np.array([[0.61, 0.0, -75.2], [0.0, 0.61, -40.9], [0.0, 0.0, 1.0]])
    out = {"jnts_xyz": jx, "com_uvd": com_uvd, "M": M}
    uvd = md.xyztouvd_np(jx)
    out["xyztouvd_np"] = uvd
    out["xyztouvd_np_single"] = md.xyztouvd_np(jx[0])
    out["uvdtoxyz"] = md.uvdtoxyz(uvd)
    out["uvdtoxyz_single"] = md.uvdtoxyz(com_uvd)
    out["calcCoMRenders"] = md.calcCoMRenders(jx)
    out["comToBounds"] = np.array(md.comToBounds(com_uvd, md.cube), np.float64)
    out["transformPoint2D"] = np.asarray(md.transformPoint2D(uvd[5], M), np.float64).reshape(2)
    rel_xyz, rel_uvd = md.getRelativeCoordinates(jx, uvd, com_uvd, M)
    out["rel_xyz"], out["rel_uvd"] = rel_xyz, rel_uvd
    a_xyz, a_uvd = md.getAbsoluteCoordinates(rel_xyz, com_uvd)
    out["abs_xyz"], out["abs_uvd"] = a_xyz, a_uvd
    batch = np.stack([jx, jx * np.float32(1.1) + np.float32(3.0)])
    batch[:, 3, 2] = -1500.0                                  # the TF projection has no z == 0 branch
    out["jnts_batch"] = batch
    out["calculateCoMfrom3DJoints"] = np.asarray(md.calculateCoMfrom3DJoints(batch))
    # the label half of prepare_data (train_cnn_networks_hgru.py:51-56) for two frames, restated line by line on the
    # reference's own methods (the trainer module itself imports the whole training stack)
    coms = np.array([com_uvd, [301.0, 150.5, 2410.0]])
    far = batch.copy()
    far[1] *= np.float32(3.0)                                 # some joints beyond the cube: the clip to [-1, 1] acts
    rel_labels = np.zeros((2, J * 3))
    for im in range(2):
        jnts_uvd = md.xyztouvd_np(far[im])
        rel_jnts_xyz, rel_jnts_uvd = md.getRelativeCoordinates(far[im], jnts_uvd, coms[im], M)
        rel_labels[im] = np.clip(np.asarray(np.reshape(rel_jnts_xyz, (J * 3,)), dtype='float32') / (md.cube[2] / 2.), -1, 1)
    out["label_jnts"], out["label_coms"], out["rel_labels"] = far, coms, rel_labels
    path = os.path.join(HERE, "detector_host_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: (np.asarray(v).shape, np.asarray(v).dtype) for k, v in out.items()})


if __name__ == "__main__":
    main()
