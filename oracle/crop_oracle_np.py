"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's crop stage
(tf_monkeydetector.tfMonkeyDetector.cropArea3D, tf_monkeydetector.py:292-365, with comToBounds :193-206,
getCrop :208-247, resizeCrop :249-263 = cv2.resize(..., INTER_NEAREST)) as called by prepare_data_test
(train_cnn_networks_hgru.py:61-74).  OpenCV is not needed: its nearest-neighbour rule
(src = min(floor(dst * (1 / (dsize/ssize))), ssize-1), all in double) is restated here.

Pinned: bit-exact against tests/golden/crop_ref.npz, which was produced by executing the reference's own
source with the real cv2 (tests/golden/make_golden_crop.py).  Never imported by the product package.
"""
import numpy as np


def com_to_bounds(com, size, fx, fy):
    """tf_monkeydetector.py:193-206."""
    zstart = com[2] - size[2] / 2.
    zend = com[2] + size[2] / 2.
    xstart = int(np.floor((com[0] * com[2] / fx - size[0] / 2.) / com[2] * fx))
    xend = int(np.floor((com[0] * com[2] / fx + size[0] / 2.) / com[2] * fx))
    ystart = int(np.floor((com[1] * com[2] / fy - size[1] / 2.) / com[2] * fy))
    yend = int(np.floor((com[1] * com[2] / fy + size[1] / 2.) / com[2] * fy))
    return xstart, xend, ystart, yend, zstart, zend


def nn_index(dst_len, src_len):
    """cv2.resize INTER_NEAREST source index for every destination index (OpenCV resizeNN)."""
    inv_scale = float(dst_len) / float(src_len)
    ifx = 1.0 / inv_scale
    return np.minimum(np.floor(np.arange(dst_len, dtype=np.float64) * ifx).astype(np.int64), src_len - 1)


def calculate_com(dpt, min_depth, max_depth):
    """calculateCoM (tf_monkeydetector.py:73-90) without scipy: ndimage.center_of_mass of the boolean mask is
    (sum of row indices, sum of column indices) / count -- integer sums, exact in float64.  The mean depth is
    numpy's float32 `dc.sum()` (pairwise order) divided by the number of non-zero pixels."""
    dc = dpt.copy()
    dc[dc < min_depth] = 0
    dc[dc > max_depth] = 0
    mask = dc > 0
    cnt = mask.sum()
    rows, cols = np.nonzero(mask)
    with np.errstate(all="ignore"):
        cc = (np.float64(rows.sum()) / cnt, np.float64(cols.sum()) / cnt)
    num = np.count_nonzero(dc)
    com = np.array((cc[1] * num, cc[0] * num, dc.sum()), np.float64)
    if num == 0:
        return np.array((0, 0, 0), np.float64)
    return com / num


def get_crop(dpt, xstart, xend, ystart, yend, zstart, zend, thresh_z=True):
    """getCrop (tf_monkeydetector.py:208-244) for a 2-D frame: slice, zero pad to the window, clamp in z."""
    H, W = dpt.shape
    if xend <= 0 or yend <= 0 or xstart >= W or ystart >= H or xend <= xstart or yend <= ystart:
        raise ValueError("crop window does not intersect the frame (the reference's slicing is undefined there)")
    wb, hb = xend - xstart, yend - ystart
    cropped = np.zeros((hb, wb), dpt.dtype)
    y0, y1, x0, x1 = max(ystart, 0), min(yend, H), max(xstart, 0), min(xend, W)
    cropped[y0 - ystart:y1 - ystart, x0 - xstart:x1 - xstart] = dpt[y0:y1, x0:x1]
    if thresh_z:
        zs, ze = cropped.dtype.type(zstart), cropped.dtype.type(zend)
        msk1 = np.logical_and(cropped < zs, cropped != 0)
        msk2 = np.logical_and(cropped > ze, cropped != 0)
        cropped[msk1] = zs
        cropped[msk2] = 0.
    return cropped


def resize_crop(crop, sz):
    """resizeCrop (tf_monkeydetector.py:246-261) with RESIZE_CV2_NN: cv2.resize(crop, (w, h), INTER_NEAREST)."""
    return crop[nn_index(sz[1], crop.shape[0])][:, nn_index(sz[0], crop.shape[1])]


def apply_crop3d(dpt, com, size, dsize, fx, fy, thresh_z=True, background=None):
    """applyCrop3D (tf_monkeydetector.py:263-290); `background` must be given (the reference's getNDValue drops into
    a debugger)."""
    xstart, xend, ystart, yend, zstart, zend = com_to_bounds(com, size, fx, fy)
    cropped = get_crop(dpt, xstart, xend, ystart, yend, zstart, zend, thresh_z)
    wb, hb = xend - xstart, yend - ystart
    if wb > hb:
        sz = (dsize[0], hb * dsize[0] // wb)
    else:
        sz = (wb * dsize[1] // hb, dsize[1])
    rz = resize_crop(cropped, sz)
    ret = np.ones((dsize[1], dsize[0]), np.float32) * np.float32(background)
    xs = int(np.floor(dsize[0] / 2. - rz.shape[1] / 2.))
    ys = int(np.floor(dsize[1] / 2. - rz.shape[0] / 2.))
    ret[ys:ys + rz.shape[0], xs:xs + rz.shape[1]] = rz
    return ret


def crop_area3d(dpt, com, cube, fx, fy, max_depth, dsize=(128, 128), docom=False, min_depth=200):
    """cropArea3D (tf_monkeydetector.py:292-365).  dpt: [H,W] float32 in mm; com None -> calculateCoM of the frame
    (:307-308); docom -> the second refinement (:316-333).  Returns (patch float32 [dsize[1], dsize[0]] in mm,
    M float64 3x3, com)."""
    H, W = dpt.shape
    if com is None:
        com = calculate_com(dpt, min_depth, max_depth)
    xstart, xend, ystart, yend, zstart, zend = com_to_bounds(com, cube, fx, fy)
    cropped = get_crop(dpt, xstart, xend, ystart, yend, zstart, zend)
    if docom:
        com = calculate_com(cropped, min_depth, max_depth)
        if np.allclose(com, 0.):
            com[2] = cropped[cropped.shape[0] // 2, cropped.shape[1] // 2]
            if np.isclose(com[2], 0):
                com[2] = 300.
        com[0] += xstart
        com[1] += ystart
        xstart, xend, ystart, yend, zstart, zend = com_to_bounds(com, cube, fx, fy)
        cropped = get_crop(dpt, xstart, xend, ystart, yend, zstart, zend)
    wb, hb = xend - xstart, yend - ystart
    # destination size of the resized crop (Python-2 integer division, :329-332)
    if wb > hb:
        sz = (dsize[0], hb * dsize[0] // wb)
    else:
        sz = (wb * dsize[1] // hb, dsize[1])
    trans = np.eye(3)
    trans[0, 2], trans[1, 2] = -xstart, -ystart
    if hb > wb:
        scale = np.eye(3) * sz[1] / float(hb)
    else:
        scale = np.eye(3) * sz[0] / float(wb)
    scale[2, 2] = 1
    rz = cropped[nn_index(sz[1], hb)][:, nn_index(sz[0], wb)]
    ret = np.ones((dsize[1], dsize[0]), np.float32) * np.float32(max_depth)
    xs = int(np.floor(dsize[0] / 2. - rz.shape[1] / 2.))
    ys = int(np.floor(dsize[1] / 2. - rz.shape[0] / 2.))
    ret[ys:ys + rz.shape[0], xs:xs + rz.shape[1]] = rz
    off = np.eye(3)
    off[0, 2], off[1, 2] = xs, ys
    return ret, off @ scale @ trans, com


def prepare_data_test(image_np, tr_res, cam, cube, image_orig_size=(424, 512), image_max_depth=10000.0,
                      target=(128, 128)):
    """train_cnn_networks_hgru.py:61-74: frames in [0,1] -> normalised 128x128 patches, CoMs, Ms."""
    n = image_np.shape[0]
    patches = np.zeros((n, target[0], target[1], 1))
    coms, Ms = [], []
    for im in range(n):
        com = tr_res[im] * [image_orig_size[0], image_orig_size[1], image_max_depth]
        dpt, M, com = crop_area3d(image_np[im] * image_max_depth, com, cube, cam[0], cam[1], image_max_depth,
                                  dsize=(target[1], target[0]))
        patches[im] = np.expand_dims(dpt, axis=2) / image_max_depth
        coms.append(com)
        Ms.append(M)
    return patches, coms, Ms
