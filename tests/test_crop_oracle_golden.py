"""CPU: the crop-stage oracle against golden vectors produced by the reference's own cropArea3D + cv2."""
import os

import numpy as np

from oracle import crop_oracle_np as crop

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "crop_ref.npz")


def test_crop_oracle_is_bit_exact_against_reference():
    z = np.load(GOLDEN)
    patches, coms, Ms = crop.prepare_data_test(z["frames"], z["coms_norm"], z["cam"], z["cube"],
                                               image_max_depth=float(z["max_depth"]))
    assert np.array_equal(patches, z["patches64"])
    assert np.array_equal(patches.astype(np.float32), z["patches"])
    for i in range(len(Ms)):
        np.testing.assert_allclose(Ms[i], z["M"][i], rtol=0, atol=1e-12)
        np.testing.assert_allclose(coms[i], z["coms_out"][i], rtol=0, atol=0)


def test_nearest_rule_matches_opencv_on_awkward_ratios():
    # sizes whose 1/(dst/src) is not exactly src/dst in double
    for dst, src in ((128, 177), (127, 178), (128, 191), (128, 96), (99, 128), (128, 255)):
        idx = crop.nn_index(dst, src)
        assert idx[0] == 0 and idx[-1] <= src - 1 and np.all(np.diff(idx) >= 0)


def test_window_outside_frame_raises():
    z = np.load(GOLDEN)
    import pytest
    with pytest.raises(ValueError):
        crop.crop_area3d(z["frames"][0] * 1e4, np.array([-900.0, 100.0, 1500.0]), z["cube"], 365.456, 365.456, 1e4)
