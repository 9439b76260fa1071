"""CPU: the post-processing oracle against golden vectors produced by the reference's own code."""
import os

import numpy as np

from oracle import post_oracle_np as post

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "post_ref.npz")


def test_absolute_coordinates_bit_exact():
    z = np.load(GOLDEN)
    xyz, uvd = post.absolute_coordinates(z["out_put"], z["coms"], z["cam"], float(z["scale"]))
    assert np.array_equal(xyz, z["xyz"])
    assert np.array_equal(uvd, z["uvd"])


def test_error_metrics():
    z = np.load(GOLDEN)
    n = z["out_put"].shape[0]
    res = np.reshape(z["out_put"], (n, -1, 3)) * np.float32(z["scale"])
    lab = np.reshape(z["labels"], (n, -1, 3)) * np.float32(z["scale"])
    assert post.mean_error(lab, res) == z["mean_error_mm"]
    assert post.max_error(lab, res) == z["max_error_mm"]


def test_metric_file_against_reference_golden():
    """Every function of pose_evaluation.py:10-88: the oracle's restatements equal the reference's own outputs."""
    import warnings
    from tests.golden.make_golden_metrics import nan_labels
    z = np.load(os.path.join(os.path.dirname(GOLDEN), "metrics_ref.npz"))
    for tag in ("a", "b"):
        res = z[tag + "_results"]
        for name, lab in (("", z[tag + "_labels"]), ("nan_", nan_labels(z[tag + "_labels"]))):
            k = tag + "_" + name
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")          # all-NaN frame: numpy warns, as it does for the reference
                assert np.float32(post.mean_error(lab, res)) == z[k + "mean_np"]
                assert np.float32(post.max_error(lab, res)) == z[k + "max_np"]
                assert np.array_equal(post.mean_axis1(lab, res), z[k + "getMean_np"])
                assert np.float32(post.mean_axis1(lab[:, 0, :], res[:, 0, :])) == z[k + "getMean_np_rank2"]
                for i, d in enumerate(z[tag + "_dists"]):
                    assert post.frames_within_max_dist(lab, res, d) == z[k + "within_max"][i]
                    assert post.frames_within_mean_dist(lab, res, d) == z[k + "within_mean"][i]
                jm = np.array([post.joint_mean_error(lab, res, j) for j in range(lab.shape[1])], np.float32)
                assert np.array_equal(jm, z[k + "joint_mean"])
        lab = z[tag + "_labels"]
        assert post.mean_error_train(lab, res) == z[tag + "_train"]
        assert np.array_equal(post.mean_axis1(lab, res, skip_nan=False), z[tag + "_getMeanError"])
        assert np.array_equal(post.mean_errors_n(lab, res), z[tag + "_getMeanErrors_N"])
        assert post.max_error_tf(lab, res) == z[tag + "_getMaxError"]
        # the thresholds are chosen so the counts are not all 0 or all N
        assert 0 < z[tag + "_within_mean"][2] < lab.shape[0]
