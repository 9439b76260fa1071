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


def test_centre_of_mass_paths_are_bit_exact_against_reference():
    """calculateCoM, cropArea3D without a centre of mass / with the docom refinement / both, the empty-window
    fallback, and the stand-alone helpers: oracle == the reference run with the real cv2 and scipy.ndimage."""
    from tests.golden.make_golden_com import CAMERA, HOLE_COM, MAX_DEPTH, crc, frame_sets
    z = np.load(GOLDEN)
    g = np.load(os.path.join(os.path.dirname(GOLDEN), "com_ref.npz"))
    fx, fy, cube, d1, d2 = CAMERA[0], CAMERA[1], CAMERA[4], CAMERA[5], CAMERA[6]
    sets = frame_sets(z["frames"])
    for name in ("plain", "far"):
        for i in range(sets[name].shape[0]):
            dpt = sets[name][i] * MAX_DEPTH
            assert np.array_equal(crop.calculate_com(dpt, d1, d2), g[name + "_com"][i])
            for key, com, docom in (("none", None, False), ("none_docom", None, True),
                                    ("given_docom", np.array(z["coms_in"][i], np.float64), True)):
                p, M, c = crop.crop_area3d(dpt, com, cube, fx, fy, d2, docom=docom, min_depth=d1)
                k = "%s_%s_" % (name, key)
                assert crc(p) == g[k + "patch_crc"][i], (name, key, i)
                assert np.array_equal(c, g[k + "com"][i])
                np.testing.assert_allclose(M, g[k + "M"][i], rtol=0, atol=1e-12)
    p, M, c = crop.crop_area3d(sets["hole"][0] * MAX_DEPTH, HOLE_COM.copy(), cube, fx, fy, d2, docom=True, min_depth=d1)
    assert crc(p) == g["hole_patch_crc"] and np.array_equal(c, g["hole_com"]) and c[2] == 300.0
    dpt = sets["plain"][2] * MAX_DEPTH
    com = np.array(z["coms_in"][2], np.float64)
    b = crop.com_to_bounds(com, cube, fx, fy)
    assert np.array_equal(np.array(b, np.float64), g["helper_bounds"])
    cropped = crop.get_crop(dpt, *b)
    assert np.array_equal(cropped.shape, g["helper_getcrop_shape"]) and crc(cropped) == g["helper_getcrop_crc"]
    assert crc(crop.get_crop(dpt, *b, thresh_z=False)) == g["helper_getcrop_nothresh_crc"]
    assert crc(crop.resize_crop(cropped, (97, 61))) == g["helper_resize_crc"]
    assert crc(crop.apply_crop3d(dpt, com, (600, 600, 900), (96, 96), fx, fy, True, 7777.0)) == g["helper_apply_crc"]
    assert crc(crop.apply_crop3d(dpt, com, (600, 600, 900), (96, 96), fx, fy, False, 7777.0)) == \
        g["helper_apply_nothresh_crc"]
