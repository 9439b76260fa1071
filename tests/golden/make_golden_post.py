"""Golden vectors for the step right after the pose network, produced by
EXECUTING THE REFERENCE'S OWN code: the x600 mm de-normalisation (train_cnn_networks_hgru.py:293-296),
`tfMonkeyDetector.getAbsoluteCoordinates` (tf_monkeydetector.py:387-391, via uvdtoxyz :138-160 and
xyztouvd :116-136) and `pose_evaluation.getMeanError_np` (pose_evaluation.py:10-15).

    python tests/golden/make_golden_post.py       (build container only)
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path.insert(0, HERE)
from make_golden_crop import load_detector_module  # noqa: E402


def load_pose_evaluation():
    src = open(os.path.join(REF, "pose_evaluation.py")).read()
    for name in ("cPickle", "argparse"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.image"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].image = sys.modules["matplotlib.image"]
    mod = types.ModuleType("pose_evaluation")
    # only the numpy metrics are needed; the module body also defines TF / plotting helpers (unused)
    src = src.replace("print ", "pass  # print ")
    exec(compile(src, os.path.join(REF, "pose_evaluation.py"), "exec"), mod.__dict__)
    return mod


def main():
    det = load_detector_module()
    pe = load_pose_evaluation()
    md = det.tfMonkeyDetector(365.456, 365.456, 256, 212, [800, 800, 1200], 200, 10000)
    rng = np.random.default_rng(31)
    n, J = 6, 23
    out_put = rng.uniform(-1.0, 1.0, size=(n, J * 3)).astype(np.float32)          # network output, normalised
    labels = (out_put + rng.normal(0, 0.02, size=out_put.shape)).astype(np.float32)
    coms = np.stack([rng.uniform(40, 470, n), rng.uniform(40, 380, n), rng.uniform(900, 3500, n)], 1)
    coms[0, 2] = 600.0
    out_put[1, 2::3] = 0.0                                                         # some joints land on z = -d
    scale = 1200 / 2.
    xyz, uvd = [], []
    for i in range(n):
        t_res = np.reshape(out_put[i], (J, 3)) * np.float32(scale)                # tf float32 graph op
        a, b = md.getAbsoluteCoordinates(t_res, coms[i])
        xyz.append(a)
        uvd.append(b)
    res_mm = np.reshape(out_put, (n, J, 3)) * np.float32(scale)
    lab_mm = np.reshape(labels, (n, J, 3)) * np.float32(scale)
    err = pe.getMeanError_np(lab_mm, res_mm)
    emax = pe.getMaxError_np(lab_mm, res_mm)
    path = os.path.join(HERE, "post_ref.npz")
    np.savez_compressed(path, out_put=out_put, labels=labels, coms=coms, xyz=np.stack(xyz), uvd=np.stack(uvd),
                        mean_error_mm=np.float64(err), max_error_mm=np.float64(emax), scale=np.float64(scale),
                        cam=np.array([365.456, 365.456, 256, 212]))
    print("wrote", path, "mean err", float(err), "max", float(emax), np.stack(xyz).dtype, np.stack(uvd).dtype)


if __name__ == "__main__":
    main()
