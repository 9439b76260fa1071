"""CPU: the camera-model methods of the detector mirror that stay on the host (numpy, a handful of scalars per frame)
against golden vectors produced by the reference's own class (tests/golden/make_golden_detector_host.py)."""
import os

import numpy as np

from monkey_pose_b200 import tf_monkeydetector as tmd
from tests.golden.make_golden_com import CAMERA

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "detector_host_ref.npz")


def test_host_camera_methods_match_the_reference():
    z = np.load(GOLDEN)
    md = tmd.tfMonkeyDetector(*CAMERA)
    jx, com, M = z["jnts_xyz"], z["com_uvd"], z["M"]
    uvd = md.xyztouvd_np(jx)
    assert uvd.dtype == np.float32 and np.array_equal(uvd, z["xyztouvd_np"])
    assert np.array_equal(md.xyztouvd(jx), z["xyztouvd_np"])                 # the mirror's xyztouvd is the numpy twin
    assert np.array_equal(md.xyztouvd_np(jx[0]), z["xyztouvd_np_single"])
    assert uvd[3, 0] == CAMERA[2] and uvd[3, 1] == CAMERA[3] and uvd[3, 2] == 0      # z == 0: the principal point
    assert np.array_equal(md.uvdtoxyz(uvd), z["uvdtoxyz"])
    assert np.array_equal(md.uvdtoxyz(com), z["uvdtoxyz_single"])
    assert np.array_equal(md.calcCoMRenders(jx), z["calcCoMRenders"])
    assert np.array_equal(np.array(md.comToBounds(com, md.cube), np.float64), z["comToBounds"])
    assert np.allclose(md.transformPoint2D(uvd[5], M), z["transformPoint2D"], rtol=0, atol=1e-12)
    rel_xyz, rel_uvd = md.getRelativeCoordinates(jx, uvd, com, M)
    assert np.array_equal(rel_xyz, z["rel_xyz"]) and np.array_equal(rel_uvd, z["rel_uvd"])
    a_xyz, a_uvd = md.getAbsoluteCoordinates(rel_xyz, com)
    assert np.array_equal(a_xyz, z["abs_xyz"]) and np.array_equal(a_uvd, z["abs_uvd"])
    got = md.calculateCoMfrom3DJoints(z["jnts_batch"])
    assert np.allclose(got, z["calculateCoMfrom3DJoints"], rtol=1e-6, atol=0)


def test_largest_window_bound_covers_the_near_plane():
    md = tmd.tfMonkeyDetector(*CAMERA)
    # a centre of mass at the near plane asks for the largest window comToBounds can produce
    xs, xe, ys, ye, _, _ = md.comToBounds(np.array([256.0, 212.0, float(md.minDepth)]), md.cube)
    assert (xe - xs) * (ye - ys) <= md._max_window_pixels(424, 512)
    assert md._max_window_pixels(4000, 4000) == 4000 * 4000


def test_relative_labels_of_prepare_data_match_the_reference():
    """The label half of the training-time `prepare_data` (train_cnn_networks_hgru.py:51-56)."""
    z = np.load(GOLDEN)
    md = tmd.tfMonkeyDetector(*CAMERA)
    got = md.relative_labels(z["label_jnts"], z["label_coms"])
    assert got.dtype == np.float64 and np.array_equal(got, z["rel_labels"])
    assert got.min() == -1.0 or got.max() == 1.0                   # the clip acts on this fixture
