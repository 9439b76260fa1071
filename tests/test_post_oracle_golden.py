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
