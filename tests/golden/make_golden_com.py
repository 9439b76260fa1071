"""Golden vectors for the centre-of-mass paths of the crop stage, produced by EXECUTING THE REFERENCE'S OWN
`tfMonkeyDetector` (tf_monkeydetector.py) with the real OpenCV and scipy.ndimage:

  * `calculateCoM` (:73-90),
  * `cropArea3D` with no centre of mass given (:307-308), with the `docom` refinement (:316-333), and both,
  * the all-empty-window fallback of the refinement (:320-323),
  * the stand-alone helpers `getCrop` (:208-244), `resizeCrop` (:246-261), `applyCrop3D` (:263-290).

    python tests/golden/make_golden_com.py        (build container only; needs /root/reference, cv2, scipy)

The depth frames are the five of crop_ref.npz (not stored again); `frame_sets` derives the variants from them.
Image-sized outputs are stored as CRC-32s of their float32 bytes (`crc`): a bit-exact comparison needs no more, and
the fixture stays a few KB.
"""
import os
import sys
import warnings
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden_crop import load_detector_module  # noqa: E402

MAX_DEPTH = 10000.0
CAMERA = (365.456, 365.456, 256, 212, [800, 800, 1200], 200, 10000)      # train_cnn_networks_hgru.py:77


def frame_sets(frames):
    """[0,1] float32 frames -> {name: frames}: as they are (the far background is in range, so the whole-frame CoM
    sits mid-image at background depth), with the background pushed beyond maxDepth (CoM lands on the blob), and
    frame 0 with an empty rectangle (for the refinement's fallback)."""
    far = np.where(frames * np.float32(MAX_DEPTH) > np.float32(3200.0), np.float32(1.2), frames).astype(np.float32)
    hole = frames[:1].copy()
    hole[0, 100:300, 100:400] = 0.0
    return {"plain": frames, "far": far, "hole": hole}


HOLE_COM = np.array([250.0, 200.0, 2000.0])          # a window wholly inside the empty rectangle of `hole`


def crc(a):
    """CRC-32 of the array's float32 bytes (one per leading index for a stack of images) + its shape."""
    a = np.ascontiguousarray(a, np.float32)
    return np.uint32(zlib.crc32(a.tobytes()))


def main():
    warnings.simplefilter("ignore")
    mod = load_detector_module()
    md = mod.tfMonkeyDetector(*CAMERA)
    z = np.load(os.path.join(HERE, "crop_ref.npz"))
    sets = frame_sets(z["frames"])
    out = {}
    for name in ("plain", "far"):
        fr = sets[name]
        n = fr.shape[0]
        com0, res = [], {k: [] for k in ("none", "none_docom", "given_docom")}
        for i in range(n):
            dpt = fr[i] * MAX_DEPTH                      # float32 x python float -> float32, as the caller does
            assert dpt.dtype == np.float32
            com0.append(md.calculateCoM(dpt))
            res["none"].append(md.cropArea3D(dpt))
            res["none_docom"].append(md.cropArea3D(dpt, docom=True))
            res["given_docom"].append(md.cropArea3D(dpt, com=np.array(z["coms_in"][i], np.float64), docom=True))
        out[name + "_com"] = np.stack(com0).astype(np.float64)
        for k, v in res.items():
            out["%s_%s_patch_crc" % (name, k)] = np.array([crc(p) for p, _, _ in v], np.uint32)
            out["%s_%s_M" % (name, k)] = np.stack([np.asarray(M, np.float64) for _, M, _ in v])
            out["%s_%s_com" % (name, k)] = np.stack([np.asarray(c, np.float64) for _, _, c in v])
    dpt = sets["hole"][0] * MAX_DEPTH
    p, M, c = md.cropArea3D(dpt, com=HOLE_COM.copy(), docom=True)
    out["hole_patch_crc"], out["hole_M"], out["hole_com"] = crc(p), np.asarray(M, np.float64), np.asarray(c)
    # stand-alone helpers on frame 2 of `plain`
    dpt = sets["plain"][2] * MAX_DEPTH
    com = np.array(z["coms_in"][2], np.float64)
    b = md.comToBounds(com, md.cube)
    out["helper_bounds"] = np.array(b, np.float64)
    cropped = md.getCrop(dpt, *b)
    out["helper_getcrop_shape"] = np.array(cropped.shape)
    out["helper_getcrop_crc"] = crc(cropped)
    out["helper_getcrop_nothresh_crc"] = crc(md.getCrop(dpt, *b, thresh_z=False))
    out["helper_resize_crc"] = crc(md.resizeCrop(cropped, (97, 61)))
    out["helper_apply_crc"] = crc(md.applyCrop3D(dpt, com, (600, 600, 900), (96, 96), True, 7777.0))
    out["helper_apply_nothresh_crc"] = crc(md.applyCrop3D(dpt, com, (600, 600, 900), (96, 96), False, 7777.0))
    path = os.path.join(HERE, "com_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)
    print("plain com", out["plain_com"][:2], "far com", out["far_com"][:2])
    print("hole com", out["hole_com"], "none_docom com", out["far_none_docom_com"][:2])
    print({k: v for k, v in out.items() if "helper" in k})


if __name__ == "__main__":
    main()
