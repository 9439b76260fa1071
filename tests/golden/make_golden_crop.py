"""Golden vectors for the crop stage, produced by EXECUTING THE REFERENCE'S
OWN `tf_monkeydetector.tfMonkeyDetector.cropArea3D` (tf_monkeydetector.py:292-365) with the real OpenCV.

    python tests/golden/make_golden_crop.py        (build container only; needs /root/reference and cv2)

The source is read as text and exec'd with stand-ins for `tensorflow` / `matplotlib` (unused on this
path) and two mechanical Python-2 fixes applied in memory: the integer divisions that size the resized
crop (`hb * dsize[0] / wb`, `wb * dsize[1] / hb` are int/int -> floor division under Python 2) and the
removed alias `numpy.float`.  Inputs follow the caller (train_cnn_networks_hgru.py:61-74,
`prepare_data_test`): depth frames in [0,1] float32 times image_max_depth, CoM = attention output times
(image height, image width, max depth) -- the reference really does scale u by the height.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path.insert(0, HERE)
import tf_shim  # noqa: E402


def load_detector_module():
    src = open(os.path.join(REF, "tf_monkeydetector.py")).read()
    src = src.replace("hb * dsize[0] / wb", "hb * dsize[0] // wb").replace("wb * dsize[1] / hb", "wb * dsize[1] // hb")
    src = src.replace("numpy.float)", "numpy.float64)")
    tfm = types.ModuleType("tensorflow")
    for k in dir(tf_shim):
        if not k.startswith("__"):
            setattr(tfm, k, getattr(tf_shim, k))
    sys.modules["tensorflow"] = tfm
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    mod = types.ModuleType("tf_monkeydetector")
    mod.__file__ = os.path.join(REF, "tf_monkeydetector.py")
    exec(compile(src, mod.__file__, "exec"), mod.__dict__)
    return mod


def synthetic_frames(rng, n, h=424, w=512):
    """Kinect-like depth frames in [0,1] (x 10000 = mm): far background, a blob, holes (0)."""
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    out = np.zeros((n, h, w), np.float32)
    coms = np.zeros((n, 3), np.float64)
    for i in range(n):
        d = rng.uniform(3500.0, 6000.0) + 16.0 * rng.integers(-12, 13, size=(h, w))   # coarse integer mm: compressible
        cy, cx = rng.uniform(0.2 * h, 0.8 * h), rng.uniform(0.2 * w, 0.8 * w)
        if i == 0:
            cy, cx = 25.0, 30.0          # crop window sticks out of the top-left corner
        if i == 1:
            cy, cx = h - 14.0, w - 12.0  # ... and out of the bottom-right
        cz = rng.uniform(1200.0, 3000.0)
        ry, rx = rng.uniform(40, 120, 2)
        blob = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0
        d = np.where(blob, np.round(cz) + 8.0 * rng.integers(-60, 110, size=(h, w)), np.round(d))
        d[rng.uniform(size=(h, w)) < 0.03] = 0.0
        out[i] = (d / 10000.0).astype(np.float32)
        # what the attention net regresses: the caller rescales by (height, width, depth) and the
        # detector reads component 0 as x -- so a trained net emits x / height, y / width
        coms[i] = (cx / h, cy / w, cz / 10000.0)
    # a few deliberately hard cases: CoM near / beyond the image border, very near and very far depth
    coms[2, 2] *= 0.8                       # CoM in front of the surface: far part of the blob -> 0
    coms[3, 2] *= 1.35                      # CoM behind it: near part clamps to zstart, smaller window
    return out, coms


def main():
    mod = load_detector_module()
    md = mod.tfMonkeyDetector(365.456, 365.456, 256, 212, [800, 800, 1200], 200, 10000)   # train_cnn_networks_hgru.py:77
    rng = np.random.default_rng(77)
    n, h, w = 5, 424, 512
    frames, coms_norm = synthetic_frames(rng, n, h, w)
    max_depth = 10000.0
    # prepare_data_test (train_cnn_networks_hgru.py:61-74)
    patches = np.zeros((n, 128, 128, 1))
    Ms, coms_out, coms_in = [], [], []
    for im in range(n):
        com = coms_norm[im] * [h, w, max_depth]
        coms_in.append(np.array(com, np.float64))
        dpt, M, com_o = md.cropArea3D(frames[im] * max_depth, com=com)
        patches[im] = np.expand_dims(dpt, axis=2) / max_depth
        Ms.append(np.asarray(M, np.float64))
        coms_out.append(np.asarray(com_o, np.float64))
    path = os.path.join(HERE, "crop_ref.npz")
    np.savez_compressed(path, frames=frames, coms_norm=coms_norm, coms_in=np.stack(coms_in),
                        patches=patches.astype(np.float32), patches64=patches, M=np.stack(Ms),
                        coms_out=np.stack(coms_out), max_depth=np.float64(max_depth),
                        cam=np.array([365.456, 365.456, 256, 212]), cube=np.array([800., 800., 1200.]))
    print("wrote", path, patches.shape, "background fraction", float((patches == 1.0).mean()))


if __name__ == "__main__":
    main()
