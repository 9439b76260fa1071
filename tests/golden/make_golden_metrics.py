"""Golden vectors for the whole metric file of the reference (pose_evaluation.py:10-88), produced by EXECUTING THE
REFERENCE'S OWN functions on float32 inputs.

    python tests/golden/make_golden_metrics.py       (build container only; needs /root/reference)

The numpy functions (`getMeanError_np`, `getMaxError_np`, `getMean_np`, `getNumFramesWithinMaxDist`,
`getNumFramesWithinMeanDist`, `getJointMeanError`) run as they are.  The TensorFlow functions (`getMeanError_train`,
`getMeanError`, `getMeanErrors_N`, `getMaxError`) run against a five-function numpy stand-in for `tf`
(reduce_mean / reduce_sum / reduce_max / sqrt / square in float32): their formulas are the reference's, the order of
TensorFlow's own reductions is not reproducible, so tests compare them with a float32 tolerance instead of bit for bit.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden_post import load_pose_evaluation  # noqa: E402


def tf_stand_in():
    tf = types.ModuleType("tensorflow")
    tf.reduce_mean = lambda x, axis=None: np.mean(np.asarray(x, np.float32), axis=axis, dtype=np.float32)
    tf.reduce_sum = lambda x, axis=None: np.sum(np.asarray(x, np.float32), axis=axis, dtype=np.float32)
    tf.reduce_max = lambda x, axis=None: np.max(np.asarray(x, np.float32), axis=axis)
    tf.sqrt = lambda x: np.sqrt(np.asarray(x, np.float32))
    tf.square = lambda x: np.square(np.asarray(x, np.float32))
    return tf


def nan_labels(lab):
    """The labels with an unlabeled joint and an unlabeled frame (tests rebuild them from the stored clean ones)."""
    nanlab = lab.copy()
    nanlab[1, 2, 0] = np.nan
    nanlab[5] = np.nan
    return nanlab


def main():
    sys.modules["tensorflow"] = tf_stand_in()
    pe = load_pose_evaluation()
    rng = np.random.default_rng(77)
    out = {}
    # two shapes: the pose network's 23 joints over a validation-size batch, and the 36 joints of main() (:94-99) over
    # a run longer than one pairwise block; millimetre-scale errors
    for tag, (n, J) in (("a", (37, 23)), ("b", (150, 36))):
        lab = rng.uniform(-300, 300, size=(n, J, 3)).astype(np.float32)
        res = (lab + rng.normal(0, 12, size=lab.shape)).astype(np.float32)
        res[3] = lab[3] + np.float32(0.25)                      # one nearly perfect frame
        nanlab = nan_labels(lab)
        dists = np.array([5.0, 15.0, 20.0, 25.0, 40.0, 80.0])
        for name, l in (("", lab), ("nan_", nanlab)):
            with np.errstate(all="ignore"):
                import warnings
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    out[tag + "_" + name + "mean_np"] = np.float32(pe.getMeanError_np(l, res))
                    out[tag + "_" + name + "max_np"] = np.float32(pe.getMaxError_np(l, res))
                    out[tag + "_" + name + "getMean_np"] = np.asarray(pe.getMean_np(l, res), np.float32)
                    out[tag + "_" + name + "getMean_np_rank2"] = np.float32(pe.getMean_np(l[:, 0, :], res[:, 0, :]))
                    out[tag + "_" + name + "within_max"] = np.array(
                        [pe.getNumFramesWithinMaxDist(l, res, d) for d in dists], np.int64)
                    out[tag + "_" + name + "within_mean"] = np.array(
                        [pe.getNumFramesWithinMeanDist(l, res, d) for d in dists], np.int64)
                    out[tag + "_" + name + "joint_mean"] = np.array(
                        [pe.getJointMeanError(l, res, j) for j in range(J)], np.float32)
        # the TensorFlow variants (no NaN handling in the reference: clean labels only)
        out[tag + "_train"] = np.float32(pe.getMeanError_train(lab, res))
        out[tag + "_getMeanError"] = np.asarray(pe.getMeanError(lab, res), np.float32)
        out[tag + "_getMeanErrors_N"] = np.asarray(pe.getMeanErrors_N(lab, res), np.float32)
        out[tag + "_getMaxError"] = np.float32(pe.getMaxError(lab, res))
        out[tag + "_labels"], out[tag + "_results"] = lab, res      # the NaN labels are rebuilt by `nan_labels`
        out[tag + "_dists"] = dists
    path = os.path.join(HERE, "metrics_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: (v.shape, v.dtype) for k, v in out.items() if "labels" not in k and "results" not in k})
    print("a mean", out["a_mean_np"], "nan mean", out["a_nan_mean_np"], "within_max", out["a_within_max"],
          out["a_within_mean"])


if __name__ == "__main__":
    main()
