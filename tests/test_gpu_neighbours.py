"""GPU (B200): the stages on either side of the pose network that round 2 widened last -- the whole metric file of
the reference (pose_evaluation.py:10-88) and the centre-of-mass estimation / refinement of the crop stage
(tf_monkeydetector.py:73-90, 292-333) -- against golden vectors produced by the reference's own code.
Everything here is bit-exact: the kernels add float32 numbers in numpy's own order (csrc/np_reduce.cuh)."""
import os
import warnings

import numpy as np
import pytest
import torch

from oracle import post_oracle_np as post
from tests.golden.make_golden_metrics import nan_labels

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _dev(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("tag", ["a", "b"])
def test_metric_file_bit_exact_against_reference_golden(tag):
    from monkey_pose_b200 import pose_evaluation as pe
    z = np.load(os.path.join(GOLDEN, "metrics_ref.npz"))
    res = z[tag + "_results"]
    for name, lab in (("", z[tag + "_labels"]), ("nan_", nan_labels(z[tag + "_labels"]))):
        k = tag + "_" + name
        L, R = _dev(lab), _dev(res)
        assert np.float32(pe.getMeanError_np(L, R)) == z[k + "mean_np"]
        assert np.float32(pe.getMaxError_np(L, R)) == z[k + "max_np"]
        assert np.array_equal(pe.getMean_np(L, R), z[k + "getMean_np"])
        assert np.float32(pe.getMean_np(_dev(lab[:, 0, :]), _dev(res[:, 0, :]))) == z[k + "getMean_np_rank2"]
        for i, d in enumerate(z[tag + "_dists"]):
            assert pe.getNumFramesWithinMaxDist(L, R, d) == z[k + "within_max"][i]
            assert pe.getNumFramesWithinMeanDist(L, R, d) == z[k + "within_mean"][i]
        jm = np.array([pe.getJointMeanError(L, R, j) for j in range(lab.shape[1])], np.float32)
        assert np.array_equal(jm, z[k + "joint_mean"])
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            assert np.array_equal(pe.joint_error_matrix(L, R).cpu().numpy(), post.joint_errors(lab, res), equal_nan=True)
    # the TensorFlow variants: device tensors, the reference's formulas (TensorFlow's own reduction order is not
    # reproducible, so the contract is a float32 tolerance; against numpy's order they are in fact exact)
    L, R = _dev(z[tag + "_labels"]), _dev(res)
    t = pe.getMeanError_train(L, R)
    assert t.is_cuda and t.dim() == 0
    assert abs(float(t) - float(z[tag + "_train"])) <= 1e-6 * float(z[tag + "_train"])
    assert np.allclose(pe.getMeanError(L, R).cpu().numpy(), z[tag + "_getMeanError"], rtol=1e-6, atol=0)
    assert np.allclose(pe.getMeanErrors_N(L, R).cpu().numpy(), z[tag + "_getMeanErrors_N"], rtol=1e-6, atol=0)
    assert float(pe.getMaxError(L, R)) == float(z[tag + "_getMaxError"])
    # a NaN poisons the TensorFlow variants and is skipped by the numpy ones
    Ln = _dev(nan_labels(z[tag + "_labels"]))
    assert np.isnan(float(pe.getMeanError_train(Ln, R))) and np.isfinite(pe.getMeanError_np(Ln, R))
    with pytest.raises(AssertionError):
        pe.getMeanError_train(L, R[:, :5])
    with pytest.raises(RuntimeError):
        pe.getMeanError_np(z[tag + "_labels"], res)          # host arrays: there is no CPU fallback


def test_metrics_at_batch_size_equal_numpy():
    """Batch 4096 x 23 joints (more frames than one pairwise block holds, unaligned splits): device == numpy."""
    from monkey_pose_b200 import pose_evaluation as pe
    rng = np.random.default_rng(5)
    for n in (129, 1000, 4096):
        lab = rng.uniform(-300, 300, size=(n, 23, 3)).astype(np.float32)
        res = (lab + rng.normal(0, 9, size=lab.shape)).astype(np.float32)
        L, R = _dev(lab), _dev(res)
        assert np.float32(pe.getMeanError_np(L, R)) == np.float32(post.mean_error(lab, res))
        assert np.array_equal(pe.getMean_np(L, R), post.mean_axis1(lab, res))
        assert pe.getNumFramesWithinMeanDist(L, R, 14.0) == post.frames_within_mean_dist(lab, res, 14.0)
        assert np.float32(pe.getJointMeanError(L, R, 11)) == np.float32(post.joint_mean_error(lab, res, 11))


# ---- centre of mass of the crop stage (tf_monkeydetector.py:73-90, 292-333) ------------------------------------------
def _detector():
    from monkey_pose_b200 import tf_monkeydetector as tmd
    from tests.golden.make_golden_com import CAMERA
    return tmd.tfMonkeyDetector(*CAMERA)


def test_centre_of_mass_paths_bit_exact_against_reference_golden():
    """calculateCoM and cropArea3D without a centre of mass / with the docom refinement / both, and the empty-window
    fallback: the device results are the reference's (run with the real cv2 and scipy.ndimage), bit for bit."""
    from tests.golden.make_golden_com import HOLE_COM, MAX_DEPTH, crc, frame_sets
    z = np.load(os.path.join(GOLDEN, "crop_ref.npz"))
    g = np.load(os.path.join(GOLDEN, "com_ref.npz"))
    md = _detector()
    sets = frame_sets(z["frames"])
    for name in ("plain", "far"):
        mm = _dev(sets[name] * np.float32(MAX_DEPTH))                     # frames in mm, as cropArea3D takes them
        n = mm.shape[0]
        assert np.array_equal(md.calculateCoM_batch(mm).cpu().numpy(), g[name + "_com"])
        # ... and from [0,1] frames with the x max-depth done on the device, as prepare_data_test's callers hold them
        assert np.array_equal(md.calculateCoM_batch(_dev(sets[name]), frame_scale=MAX_DEPTH).cpu().numpy(),
                              g[name + "_com"])
        assert np.array_equal(md.calculateCoM(mm[1]), g[name + "_com"][1])
        for key, coms, docom in (("none", None, False), ("none_docom", None, True),
                                 ("given_docom", _dev(z["coms_in"].astype(np.float64)), True)):
            k = "%s_%s_" % (name, key)
            p, Ms, c = md.cropArea3D_batch_device(mm, coms=coms, docom=docom)
            assert not md.last_invalid_dev.any()
            assert np.array_equal(c.cpu().numpy(), g[k + "com"]), (name, key)
            np.testing.assert_allclose(Ms.cpu().numpy(), g[k + "M"], rtol=0, atol=1e-12)
            got = np.array([crc(p[i].cpu().numpy()) for i in range(n)], np.uint32)
            assert np.array_equal(got, g[k + "patch_crc"]), (name, key)
        # the single-frame method of the reference class
        p, M, c = md.cropArea3D(mm[3], docom=True)
        assert crc(p.cpu().numpy()) == g[name + "_none_docom_patch_crc"][3]
        assert np.array_equal(c, g[name + "_none_docom_com"][3])
    hole = _dev(sets["hole"][0] * np.float32(MAX_DEPTH))
    p, M, c = md.cropArea3D(hole, com=HOLE_COM.copy(), docom=True)
    assert np.array_equal(c, g["hole_com"]) and c[2] == 300.0
    assert crc(p.cpu().numpy()) == g["hole_patch_crc"]
    np.testing.assert_allclose(M, g["hole_M"], rtol=0, atol=1e-12)


def test_crop_building_blocks_bit_exact_against_reference_golden():
    """getCrop / resizeCrop / applyCrop3D as stand-alone calls (tf_monkeydetector.py:208-290)."""
    from tests.golden.make_golden_com import MAX_DEPTH, crc, frame_sets
    z = np.load(os.path.join(GOLDEN, "crop_ref.npz"))
    g = np.load(os.path.join(GOLDEN, "com_ref.npz"))
    md = _detector()
    dpt = _dev(frame_sets(z["frames"])["plain"][2] * np.float32(MAX_DEPTH))
    com = np.array(z["coms_in"][2], np.float64)
    b = md.comToBounds(com, md.cube)
    assert np.array_equal(np.array(b, np.float64), g["helper_bounds"])
    cropped = md.getCrop(dpt, *b)
    assert np.array_equal(tuple(cropped.shape), g["helper_getcrop_shape"])
    assert crc(cropped.cpu().numpy()) == g["helper_getcrop_crc"]
    assert crc(md.getCrop(dpt, *b, thresh_z=False).cpu().numpy()) == g["helper_getcrop_nothresh_crc"]
    assert crc(md.resizeCrop(cropped, (97, 61)).cpu().numpy()) == g["helper_resize_crc"]
    assert crc(md.applyCrop3D(dpt, com, (600, 600, 900), (96, 96), True, 7777.0).cpu().numpy()) == g["helper_apply_crc"]
    assert crc(md.applyCrop3D(dpt, com, (600, 600, 900), (96, 96), False, 7777.0).cpu().numpy()) == \
        g["helper_apply_nothresh_crc"]
    with pytest.raises(ValueError):
        md.applyCrop3D(dpt, com, (600, 600, 900), (96, 96))              # the reference's default background cannot run
    with pytest.raises(ValueError):
        md.getCrop(dpt, -900, -700, 10, 200, 0.0, 1.0)


def test_centre_of_mass_at_batch_size_vs_oracle():
    """64 Kinect-size frames + odd image sizes (pairwise trees of every shape) against the numpy oracle: centres of
    mass, refined centres, windows and patches all bit-exact; an all-empty frame gives the reference's zeros."""
    from oracle import crop_oracle_np as crop
    from tests.golden.make_golden_com import CAMERA
    md = _detector()
    fx, fy, cube, d1, d2 = CAMERA[0], CAMERA[1], CAMERA[4], CAMERA[5], CAMERA[6]
    rng = np.random.default_rng(11)
    for (n, h, w) in ((64, 424, 512), (3, 97, 131), (2, 8, 15), (2, 1, 5), (1, 480, 640)):
        base = np.round(rng.uniform(150, 11000, size=(n, h, w))).astype(np.float32)      # some below / beyond the planes
        yy, xx = np.mgrid[0:h, 0:w]
        for i in range(n):
            cy, cx, r = rng.uniform(0.2 * h, 0.8 * h), rng.uniform(0.2 * w, 0.8 * w), 0.15 * min(h, w) + 1
            blob = (yy - cy) ** 2 + (xx - cx) ** 2 <= r * r
            base[i] = np.where(blob, np.round(rng.uniform(900, 2500) + rng.uniform(-80, 80, size=(h, w))), 12000.0)
            base[i][rng.uniform(size=(h, w)) < 0.05] = 0.0
        if n > 2:
            base[2] = 0.0                                                # nothing in range: com = (0, 0, 0)
        F = _dev(base)
        got = md.calculateCoM_batch(F).cpu().numpy()
        want = np.stack([crop.calculate_com(base[i], d1, d2) for i in range(n)])
        assert np.array_equal(got, want), (n, h, w)
        assert not md.last_com_overflow_dev.any()
        if h < 64:
            continue
        p, Ms, c = md.cropArea3D_batch_device(F, docom=True)
        inv = md.last_invalid_dev.cpu().numpy()
        for i in range(n):
            if n > 2 and i == 2:
                assert inv[i] == 1                                       # com (0,0,0): no window; flagged, not fatal
                continue
            wp, wM, wc = crop.crop_area3d(base[i], None, cube, fx, fy, d2, docom=True, min_depth=d1)
            assert inv[i] == 0
            assert np.array_equal(c[i].cpu().numpy(), wc), (n, h, w, i)
            assert np.array_equal(p[i].cpu().numpy(), wp), (n, h, w, i)
            np.testing.assert_allclose(Ms[i].cpu().numpy(), wM, rtol=0, atol=1e-12)


def test_centre_of_mass_nan_pixels_and_non_positive_near_plane():
    """The kernel's fast path assumes a positive near plane and NaN-free blocks and falls back otherwise: a NaN pixel
    (stays in the image: counted as non-zero, not in the mask, the depth sum turns NaN) and a detector whose near plane
    is 0 or negative (zeros and negative depths then pass the range test) give the oracle's values."""
    from monkey_pose_b200 import tf_monkeydetector as tmd
    from oracle import crop_oracle_np as crop
    rng = np.random.default_rng(21)
    base = np.round(rng.uniform(-50, 11000, size=(4, 200, 333))).astype(np.float32)
    base[rng.uniform(size=base.shape) < 0.1] = 0.0
    base[1, 57, 100] = np.nan
    base[3, 0, 0] = np.nan
    base[3, 199, 332] = -0.0
    for d1 in (200, 0, -10):
        md = tmd.tfMonkeyDetector(365.456, 365.456, 256, 212, [800, 800, 1200], d1, 10000)
        got = md.calculateCoM_batch(_dev(base)).cpu().numpy()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            want = np.stack([crop.calculate_com(base[i], d1, 10000) for i in range(4)])
        assert np.array_equal(got, want, equal_nan=True), d1
        assert np.isnan(got[1, 2]) and np.isfinite(got[1, :2]).all() and np.isfinite(got[0]).all()


def test_frames_to_joints_without_attention_network_uses_the_detector_estimate():
    """FramesToJoints(attn=None): centres of mass from calculateCoM (+ docom) on the device; equals the stage-by-stage
    sequence bit for bit, from pinned host frames (chunked upload) and from device frames."""
    import monkey_pose_b200 as mp
    from monkey_pose_b200 import initialization as init
    from tests.golden.make_golden_com import MAX_DEPTH, frame_sets

    class Cfg(object):
        image_orig_size = [424, 512, 1]
        image_target_size = [128, 128, 1]
        image_max_depth = MAX_DEPTH

    z = np.load(os.path.join(GOLDEN, "crop_ref.npz"))
    frames = np.concatenate([frame_sets(z["frames"])["far"]] * 4)[:16]          # 16 frames: 4 upload pieces of 4
    md = _detector()
    m = mp.model()
    m.channels, m.timesteps, m.fc_hidden, m.hidden_state = 25, 2, 64, None
    m.aux["hidden_init"] = "zeros"
    m.load_params(init.pose_params(channels=25, S=15, T=2, hw=64, fc_hidden=64, out=69, seed=3))
    for docom in (False, True):
        pipe = mp.FramesToJoints(None, m, md, Cfg(), cube_z=1200.0, chunks=4, docom=docom)
        xyz, uvd = pipe(torch.as_tensor(frames).pin_memory())
        F = _dev(frames)
        patches, _, coms = md.cropArea3D_batch_device(F, frame_scale=MAX_DEPTH, out_divisor=MAX_DEPTH, docom=docom)
        want_xyz, want_uvd = md.getAbsoluteCoordinates_batch(m.build(patches[..., None], 69), coms, 600.0)
        assert torch.equal(xyz, want_xyz.cpu()) and torch.equal(uvd, want_uvd.cpu())
        xyz2, uvd2 = pipe(F)
        assert torch.equal(xyz2, xyz) and torch.equal(uvd2, uvd)
        assert torch.isfinite(xyz).all()
