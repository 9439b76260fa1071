"""B200-native drop-in for the crop stage of the reference (`tf_monkeydetector.tfMonkeyDetector`,
tf_monkeydetector.py:21-391, as used by `prepare_data_test`, train_cnn_networks_hgru.py:61-74).

Same constructor and method names.  The camera / window arithmetic (a handful of scalars per frame)
stays in numpy float64 exactly as the reference writes it; the per-pixel work -- slice, zero pad,
z clamp, cv2 nearest-neighbour resize, paste, normalise -- runs in one CUDA kernel over the whole batch
(`crop_area3d_forward`), bit-exact with the reference + OpenCV; the centre-of-mass estimate the reference
falls back to when none is given, and its `docom` refinement (`calculateCoM`, :73-90, :307-333), run on the
device too (`calculate_com_forward`), bit-exact as well.  No CPU fallback.

Not mirrored: `checkImage` / `getNDValue` (:163-183) read `self.dpt`, which the reference's constructor no longer
sets (and `getNDValue` drops into a debugger) -- they cannot run in the reference either.
"""
import numpy

import torch

from . import _lib
from .hgru_module import _stream


class tfMonkeyDetector(object):
    RESIZE_BILINEAR = 0
    RESIZE_CV2_NN = 1
    RESIZE_CV2_LINEAR = 2

    def __init__(self, fx, fy, ux, uy, cube, d1, d2, importer=None):
        """tf_monkeydetector.py:30-64."""
        self.maxDepth = d2
        self.minDepth = d1
        self.fx, self.fy, self.ux, self.uy = fx, fy, ux, uy
        if len(cube) != 3:
            raise ValueError("Volume must be 3D")
        self.cube = cube
        self.resizeMethod = self.RESIZE_CV2_NN

    # ---- camera model (host, numpy; tf_monkeydetector.py:116-160, 193-206, 367-391) -------------------
    def xyztouvd(self, jnts_xyz):
        j = numpy.asarray(jnts_xyz)
        single = j.ndim == 1
        j = numpy.atleast_2d(j)
        out = numpy.zeros((j.shape[0], 3), numpy.float32)
        z = j[:, 2]
        nz = z != 0.
        out[~nz, 0], out[~nz, 1] = self.ux, self.uy
        out[nz, 0] = self.ux - j[nz, 0] / z[nz] * self.fx
        out[nz, 1] = j[nz, 1] / z[nz] * self.fy + self.uy
        out[nz, 2] = -z[nz]
        return out[0] if single else out

    def uvdtoxyz(self, jnts_uvd):
        j = numpy.asarray(jnts_uvd)
        single = j.ndim == 1
        j = numpy.atleast_2d(j)
        out = numpy.zeros((j.shape[0], 3), numpy.float32)
        out[:, 0] = (self.ux - j[:, 0]) * j[:, 2] / (-self.fx)
        out[:, 1] = (j[:, 1] - self.uy) * j[:, 2] / (-self.fy)
        out[:, 2] = -j[:, 2]
        return out[0] if single else out

    xyztouvd_np = xyztouvd          # tf_monkeydetector.py:116-136 (the reference's `xyztouvd` is its TF twin)

    def calcCoMRenders(self, jnts):
        """Centre of mass of a render's joints in 3-D (tf_monkeydetector.py:185-191)."""
        jnts = numpy.asarray(jnts)
        assert jnts.ndim == 2, 'input must be the 3D coordinates of all monkey joints'
        return numpy.sum(jnts, axis=0) / jnts.shape[0]

    def calculateCoMfrom3DJoints(self, jnts):
        """Mean joint position [N,J,3] -> [N,3], projected to the image (tf_monkeydetector.py:66-71; the TF
        projection has no z == 0 branch).  Label preparation: host arrays."""
        com_3d = numpy.mean(numpy.asarray(jnts), axis=1)
        u_s = self.ux - com_3d[:, 0] / com_3d[:, 2] * self.fx
        v_s = self.uy + com_3d[:, 1] / com_3d[:, 2] * self.fy
        return numpy.stack([u_s, v_s, -com_3d[:, 2]], axis=1)

    def relative_labels(self, jnts_xyz, coms):
        """The label half of the reference's `prepare_data` (train_cnn_networks_hgru.py:51-56) for a batch: joints
        [N,J,3] (camera space, mm) relative to each frame's centre of mass (`getRelativeCoordinates`, :372-385), as
        float32, divided by cube[2] / 2 and clipped to [-1, 1] -> numpy float64 [N, 3J] (the reference fills a
        float64 array).  A few dozen numbers per frame: host numpy, the reference's own operations."""
        j = numpy.asarray(jnts_xyz)
        coms = numpy.asarray(coms.cpu() if torch.is_tensor(coms) else coms, numpy.float64).reshape(j.shape[0], 3)
        out = numpy.zeros((j.shape[0], j.shape[1] * j.shape[2]))
        for im in range(j.shape[0]):
            rel = j[im] - self.uvdtoxyz(coms[im])
            out[im] = numpy.clip(numpy.asarray(numpy.reshape(rel, (-1,)), dtype='float32') / (self.cube[2] / 2.), -1, 1)
        return out

    # ---- centre of mass of a depth image (device; tf_monkeydetector.py:73-90) ---------------------------
    def _max_window_pixels(self, H, W):
        """The largest window comToBounds can ask for: a centre of mass at the near plane (:193-206)."""
        near = max(float(self.minDepth), 1.0)
        wb = int(numpy.ceil(self.cube[0] / near * self.fx)) + 2
        hb = int(numpy.ceil(self.cube[1] / near * self.fy)) + 2
        return min(max(wb * hb, H * W), 2 ** 31 - 1)

    def calculateCoM_batch(self, frames, frame_scale=1.0, windows=None):
        """calculateCoM for every frame of a [N,H,W] torch CUDA tensor (times frame_scale = mm) -> CUDA float64
        [N,3] (x, y, mean depth), bit for bit the reference's values for float32 frames.  `windows` = (iparams,
        zparams) CUDA tensors of `crop_windows_forward`: the centre of mass of each frame's z-clamped crop window
        with cropArea3D's `docom` fallbacks and offset (:316-326) instead."""
        if not (torch.is_tensor(frames) and frames.is_cuda and frames.dim() == 3):
            raise RuntimeError("frames must be a [N,H,W] torch CUDA tensor (no CPU fallback)")
        frames = frames.to(torch.float32).contiguous()
        N, H, W = [int(v) for v in frames.shape]
        lib = _lib.load()
        max_px = self._max_window_pixels(H, W) if windows is not None else H * W
        ws = torch.empty(int(lib.calculate_com_workspace_bytes(N, max_px)), device=frames.device, dtype=torch.uint8)
        coms = torch.empty((N, 3), device=frames.device, dtype=torch.float64)
        over = torch.empty((N,), device=frames.device, dtype=torch.int32)
        ip, zp = windows if windows is not None else (None, None)
        _lib.check(lib.calculate_com_forward(
            frames.data_ptr(), N, H, W, float(frame_scale), float(self.minDepth), float(self.maxDepth),
            ip.data_ptr() if ip is not None else None, zp.data_ptr() if zp is not None else None, max_px,
            ws.data_ptr(), coms.data_ptr(), over.data_ptr(), _stream()), "calculate_com_forward")
        self.last_com_overflow_dev = over
        return coms

    def calculateCoM(self, dpt):
        """Centre of mass (x, y, z) of one depth image [H,W] in mm (tf_monkeydetector.py:73-90): numpy float64 [3]."""
        return self.calculateCoM_batch(dpt[None]).cpu().numpy()[0]

    def comToBounds(self, com, size):
        zstart = com[2] - size[2] / 2.
        zend = com[2] + size[2] / 2.
        xstart = int(numpy.floor((com[0] * com[2] / self.fx - size[0] / 2.) / com[2] * self.fx))
        xend = int(numpy.floor((com[0] * com[2] / self.fx + size[0] / 2.) / com[2] * self.fx))
        ystart = int(numpy.floor((com[1] * com[2] / self.fy - size[1] / 2.) / com[2] * self.fy))
        yend = int(numpy.floor((com[1] * com[2] / self.fy + size[1] / 2.) / com[2] * self.fy))
        return xstart, xend, ystart, yend, zstart, zend

    def transformPoint2D(self, pt, M):
        pt2 = numpy.asarray(M, numpy.float64).reshape(3, 3) @ numpy.array([pt[0], pt[1], 1.0])
        return numpy.array([pt2[0] / pt2[2], pt2[1] / pt2[2]])

    def getRelativeCoordinates(self, jnts_xyz, jnts_uvd, com_uvd, M):
        com_xyz = self.uvdtoxyz(com_uvd)
        rel_jnts_xyz = jnts_xyz - com_xyz
        rel_jnts_uvd = numpy.zeros((jnts_uvd.shape[0], 3), numpy.float32)
        for joint in range(jnts_uvd.shape[0]):
            t = self.transformPoint2D(jnts_uvd[joint], M)
            rel_jnts_uvd[joint, 0], rel_jnts_uvd[joint, 1] = t[0], t[1]
            rel_jnts_uvd[joint, 2] = jnts_uvd[joint, 2]
        return rel_jnts_xyz, rel_jnts_uvd

    def getAbsoluteCoordinates(self, rel_jnts_xyz, com_uvd):
        com_xyz = self.uvdtoxyz(com_uvd)
        jnts_xyz = rel_jnts_xyz + com_xyz
        return jnts_xyz, self.xyztouvd(jnts_xyz)

    def getAbsoluteCoordinates_batch(self, out_put, coms, scale):
        """Batched, on the device: network output [N,3J] (normalised) x scale (cube[2]/2,
        train_cnn_networks_hgru.py:293-296) + getAbsoluteCoordinates per frame -> (xyz, uvd) [N,J,3]."""
        if not (torch.is_tensor(out_put) and out_put.is_cuda and out_put.dim() == 2 and out_put.shape[1] % 3 == 0):
            raise RuntimeError("out_put must be a [N,3J] torch CUDA tensor (no CPU fallback)")
        out_put = out_put.to(torch.float32).contiguous()
        N, J = int(out_put.shape[0]), int(out_put.shape[1]) // 3
        if torch.is_tensor(coms) and coms.is_cuda:          # centres of mass already on the device (device crop path)
            com_d = coms.to(torch.float64).reshape(N, 3).contiguous()
        else:
            com_d = torch.as_tensor(numpy.asarray(coms, numpy.float64).reshape(N, 3)).cuda()
        xyz = torch.empty((N, J, 3), device=out_put.device, dtype=torch.float32)
        uvd = torch.empty_like(xyz)
        _lib.check(_lib.load().pose_postprocess_forward(
            out_put.data_ptr(), com_d.data_ptr(), N, J, float(self.fx), float(self.fy), float(self.ux),
            float(self.uy), float(scale), xyz.data_ptr(), uvd.data_ptr(), _stream()), "pose_postprocess_forward")
        return xyz, uvd

    # ---- crop (device) ----------------------------------------------------------------------------------
    def _window(self, com, H, W, dsize):
        """Host-side integers of one frame's crop (tf_monkeydetector.py:309-362): window, resized size,
        paste offset, the 3x3 transform.  Mirrors the reference line by line (Python-2 integer division)."""
        xstart, xend, ystart, yend, zstart, zend = self.comToBounds(com, self.cube)
        if xend <= 0 or yend <= 0 or xstart >= W or ystart >= H or xend <= xstart or yend <= ystart:
            raise ValueError("crop window does not intersect the frame")
        wb, hb = xend - xstart, yend - ystart
        if wb > hb:
            sz = (dsize[0], hb * dsize[0] // wb)
        else:
            sz = (wb * dsize[1] // hb, dsize[1])
        trans = numpy.eye(3, dtype=float)
        trans[0, 2], trans[1, 2] = -xstart, -ystart
        if hb > wb:
            scale = numpy.eye(3, dtype=float) * sz[1] / float(hb)
        else:
            scale = numpy.eye(3, dtype=float) * sz[0] / float(wb)
        scale[2, 2] = 1
        xs = int(numpy.floor(dsize[0] / 2. - sz[0] / 2.))
        ys = int(numpy.floor(dsize[1] / 2. - sz[1] / 2.))
        off = numpy.eye(3, dtype=float)
        off[0, 2], off[1, 2] = xs, ys
        return (xstart, ystart, wb, hb, sz[0], sz[1], xs, ys), (zstart, zend), off @ scale @ trans

    def _windows_batch(self, coms, H, W, dsize):
        """`_window` for a whole batch in array form (same float64 operations element by element, so the
        integers are identical): at 40K+ frames/s a per-frame Python loop costs more than the network.
        Returns (ints [N,8] int32, z [N,2] float32, Ms [N,3,3] float64)."""
        c = numpy.asarray(coms, numpy.float64).reshape(-1, 3)
        sx, sy, sz = [float(v) for v in self.cube]
        zstart, zend = c[:, 2] - sz / 2., c[:, 2] + sz / 2.
        xstart = numpy.floor((c[:, 0] * c[:, 2] / self.fx - sx / 2.) / c[:, 2] * self.fx).astype(numpy.int64)
        xend = numpy.floor((c[:, 0] * c[:, 2] / self.fx + sx / 2.) / c[:, 2] * self.fx).astype(numpy.int64)
        ystart = numpy.floor((c[:, 1] * c[:, 2] / self.fy - sy / 2.) / c[:, 2] * self.fy).astype(numpy.int64)
        yend = numpy.floor((c[:, 1] * c[:, 2] / self.fy + sy / 2.) / c[:, 2] * self.fy).astype(numpy.int64)
        # A window that misses the frame (a wild centre-of-mass prediction) must not abort the other frames of the
        # batch: that frame gets an all-background patch (resized size 0: nothing is pasted) and is reported in
        # `self.last_invalid`; the single-frame `_window` / `cropArea3D` raise ValueError for it.
        bad = (xend <= 0) | (yend <= 0) | (xstart >= W) | (ystart >= H) | (xend <= xstart) | (yend <= ystart) | \
            ~numpy.isfinite(c).all(axis=1)
        self.last_invalid = numpy.nonzero(bad)[0]
        if bad.any():
            xstart, ystart = numpy.where(bad, 0, xstart), numpy.where(bad, 0, ystart)
            xend, yend = numpy.where(bad, 1, xend), numpy.where(bad, 1, yend)
        wb, hb = xend - xstart, yend - ystart
        wide = wb > hb
        szx = numpy.where(wide, dsize[0], wb * dsize[1] // hb)
        szy = numpy.where(wide, hb * dsize[0] // wb, dsize[1])
        if bad.any():
            szx, szy = numpy.where(bad, 0, szx), numpy.where(bad, 0, szy)
        sc = numpy.where(hb > wb, szy / hb.astype(numpy.float64), szx / wb.astype(numpy.float64))
        xs = numpy.floor(dsize[0] / 2. - szx / 2.).astype(numpy.int64)
        ys = numpy.floor(dsize[1] / 2. - szy / 2.).astype(numpy.int64)
        # M = off @ scale @ trans written out (the products only ever add exact zeros, so this is the same
        # floating-point result; a stacked 3x3 matmul would cost one BLAS call per frame)
        Ms = numpy.zeros((c.shape[0], 3, 3), numpy.float64)
        Ms[:, 0, 0] = Ms[:, 1, 1] = sc
        Ms[:, 0, 2] = sc * (-xstart) + xs
        Ms[:, 1, 2] = sc * (-ystart) + ys
        Ms[:, 2, 2] = 1.0
        ints = numpy.stack([xstart, ystart, wb, hb, szx, szy, xs, ys], 1).astype(numpy.int32)
        return ints, numpy.stack([zstart, zend], 1).astype(numpy.float32), Ms

    def cropArea3D_batch(self, frames, coms, dsize=(128, 128), frame_scale=1.0, out_divisor=1.0):
        """Batched cropArea3D: frames [N,H,W] torch CUDA float32 (times frame_scale = mm), coms [N,3]
        (u, v, d mm).  Returns (patches [N,dsize[1],dsize[0]] CUDA = mm / out_divisor, Ms, coms)."""
        if len(dsize) != 2:
            raise ValueError("dsize must be a 2D bounding box")
        if not (torch.is_tensor(frames) and frames.is_cuda and frames.dim() == 3):
            raise RuntimeError("frames must be a [N,H,W] torch CUDA tensor (no CPU fallback)")
        frames = frames.to(torch.float32).contiguous()
        N, H, W = [int(v) for v in frames.shape]
        coms = numpy.asarray(coms, numpy.float64).reshape(N, 3)
        ip, zp, Ms = self._windows_batch(coms, H, W, dsize)
        ip_d = torch.as_tensor(ip).cuda()
        zp_d = torch.as_tensor(zp).cuda()
        out = torch.empty((N, dsize[1], dsize[0]), device=frames.device, dtype=torch.float32)
        _lib.check(_lib.load().crop_area3d_forward(
            frames.data_ptr(), N, H, W, float(frame_scale), ip_d.data_ptr(), zp_d.data_ptr(), float(self.maxDepth),
            float(out_divisor), out.data_ptr(), int(dsize[1]), int(dsize[0]), _stream()), "crop_area3d_forward")
        return out, list(Ms), list(coms)

    def cropArea3D_batch_device(self, frames, tr=None, tr_scale=(1.0, 1.0, 1.0), coms=None, dsize=(128, 128),
                                frame_scale=1.0, out_divisor=1.0, docom=False):
        """cropArea3D_batch with the window arithmetic on the device too (`crop_windows_forward`): no host round trip
        between the attention CNN and the crop.  Centres of mass either as `coms` (CUDA float64 [N,3]; u, v, d mm), as
        attention outputs `tr` (CUDA float32 [N,3]) times `tr_scale` (train_cnn_networks_hgru.py:66-68), or -- neither
        given -- estimated from each frame by `calculateCoM` as the reference does (tf_monkeydetector.py:307-308).
        `docom` adds the reference's second refinement (:316-333): the centre of mass of the first crop window
        replaces the first estimate and the window is computed again.
        Returns (patches, Ms, coms) as CUDA tensors ([N,dh,dw] float32, [N,3,3] float64, [N,3] float64);
        `self.last_invalid_dev` (CUDA int32 [N]) flags frames whose window misses the frame (all-background patch)."""
        if len(dsize) != 2:
            raise ValueError("dsize must be a 2D bounding box")
        if not (torch.is_tensor(frames) and frames.is_cuda and frames.dim() == 3):
            raise RuntimeError("frames must be a [N,H,W] torch CUDA tensor (no CPU fallback)")
        frames = frames.to(torch.float32).contiguous()
        N, H, W = [int(v) for v in frames.shape]
        dev = frames.device
        if coms is not None:
            if not (torch.is_tensor(coms) and coms.is_cuda):
                raise RuntimeError("coms must be a CUDA tensor on the device path")
            coms = coms.to(torch.float64).reshape(N, 3).contiguous()
        elif tr is not None:
            if not (torch.is_tensor(tr) and tr.is_cuda):
                raise RuntimeError("tr must be a CUDA tensor on the device path")
            tr = tr.to(torch.float32).reshape(N, 3).contiguous()
        else:
            coms = self.calculateCoM_batch(frames, frame_scale=frame_scale)
        coms_out = torch.empty((N, 3), device=dev, dtype=torch.float64)
        ip = torch.empty((N, 8), device=dev, dtype=torch.int32)
        zp = torch.empty((N, 2), device=dev, dtype=torch.float32)
        Ms = torch.empty((N, 3, 3), device=dev, dtype=torch.float64)
        inv = torch.empty((N,), device=dev, dtype=torch.int32)
        out = torch.empty((N, dsize[1], dsize[0]), device=dev, dtype=torch.float32)
        lib = _lib.load()

        def windows(tr_, coms_):
            _lib.check(lib.crop_windows_forward(
                tr_.data_ptr() if coms_ is None else None, coms_.data_ptr() if coms_ is not None else None,
                float(tr_scale[0]), float(tr_scale[1]), float(tr_scale[2]), N, H, W, int(dsize[0]), int(dsize[1]),
                float(self.fx), float(self.fy), float(self.cube[0]), float(self.cube[1]), float(self.cube[2]),
                coms_out.data_ptr(), ip.data_ptr(), zp.data_ptr(), Ms.data_ptr(), inv.data_ptr(), _stream()),
                "crop_windows_forward")

        windows(tr, coms)
        if docom:
            # a window that misses the frame has no crop to refine on: the reference's slicing is undefined there
            # (the host path raises); its flag from the FIRST pass is kept
            first_inv = inv.clone()
            refined = self.calculateCoM_batch(frames, frame_scale=frame_scale, windows=(ip, zp))
            windows(None, refined)
            inv = torch.maximum(inv, first_inv)
        _lib.check(lib.crop_area3d_forward(
            frames.data_ptr(), N, H, W, float(frame_scale), ip.data_ptr(), zp.data_ptr(), float(self.maxDepth),
            float(out_divisor), out.data_ptr(), int(dsize[1]), int(dsize[0]), _stream()), "crop_area3d_forward")
        self.last_invalid_dev = inv
        self._last_windows_dev = (ip, zp)
        return out, Ms, coms_out

    def cropArea3D(self, dpt, com=None, dsize=(128, 128), docom=False):
        """tf_monkeydetector.py:292-365 for one frame [H,W] (mm, torch CUDA): (patch, M, com).  Without `com` the
        centre of mass is estimated from the frame (:307-308); `docom` refines it on the first crop (:316-333)."""
        if com is None or docom:
            if not (torch.is_tensor(dpt) and dpt.is_cuda and dpt.dim() == 2):
                raise RuntimeError("dpt must be a [H,W] torch CUDA tensor (no CPU fallback)")
            coms = None if com is None else torch.as_tensor(numpy.asarray(com, numpy.float64).reshape(1, 3)).cuda()
            out, Ms, coms = self.cropArea3D_batch_device(dpt[None], coms=coms, dsize=dsize, docom=docom)
            if int(self.last_invalid_dev[0].item()):
                raise ValueError("crop window does not intersect the frame")
            return out[0], Ms[0].cpu().numpy(), coms[0].cpu().numpy()
        out, Ms, coms = self.cropArea3D_batch(dpt[None], [com], dsize=dsize)
        if len(self.last_invalid):
            raise ValueError("crop window does not intersect the frame")
        return out[0], Ms[0], coms[0]

    # ---- the crop's building blocks as stand-alone calls (tf_monkeydetector.py:208-290) ------------------
    def _crop_call(self, img, ints, z, background, dsize):
        """One frame through `crop_area3d_forward` with hand-made window integers."""
        if not (torch.is_tensor(img) and img.is_cuda):
            raise RuntimeError("the image must be a torch CUDA tensor (no CPU fallback)")
        if img.dim() != 2:
            raise NotImplementedError()
        img = img.to(torch.float32).contiguous()
        H, W = int(img.shape[0]), int(img.shape[1])
        ip = torch.as_tensor(numpy.asarray([ints], numpy.int32)).cuda()
        zp = torch.as_tensor(numpy.asarray([z], numpy.float32)).cuda()
        out = torch.empty((1, dsize[1], dsize[0]), device=img.device, dtype=torch.float32)
        _lib.check(_lib.load().crop_area3d_forward(
            img.data_ptr(), 1, H, W, 1.0, ip.data_ptr(), zp.data_ptr(), float(background), 1.0, out.data_ptr(),
            int(dsize[1]), int(dsize[0]), _stream()), "crop_area3d_forward")
        return out[0]

    def getCrop(self, dpt, xstart, xend, ystart, yend, zstart, zend, thresh_z=True):
        """Crop the window out of a depth image [H,W], zero-padded where it leaves the image, clamped in z
        (tf_monkeydetector.py:208-244): [yend - ystart, xend - xstart] CUDA float32."""
        H, W = int(dpt.shape[0]), int(dpt.shape[1])
        if xend <= 0 or yend <= 0 or xstart >= W or ystart >= H or xend <= xstart or yend <= ystart:
            raise ValueError("crop window does not intersect the frame")
        wb, hb = int(xend - xstart), int(yend - ystart)
        z = (zstart, zend) if thresh_z is True else (-numpy.inf, numpy.inf)
        return self._crop_call(dpt, (xstart, ystart, wb, hb, wb, hb, 0, 0), z, 0.0, (wb, hb))

    def resizeCrop(self, crop, sz):
        """cv2.resize(crop, sz, INTER_NEAREST) (tf_monkeydetector.py:246-261; sz = (width, height))."""
        if self.resizeMethod != self.RESIZE_CV2_NN:
            raise NotImplementedError("Unknown resize method!")
        h, w = int(crop.shape[0]), int(crop.shape[1])
        return self._crop_call(crop, (0, 0, w, h, int(sz[0]), int(sz[1]), 0, 0), (-numpy.inf, numpy.inf), 0.0,
                               (int(sz[0]), int(sz[1])))

    def applyCrop3D(self, dpt, com, size, dsize, thresh_z=True, background=None):
        """Crop a `size` (mm) volume around `com` and paste it, resized, into a `background` image
        (tf_monkeydetector.py:263-290).  `background` must be given: the reference's default, getNDValue, cannot run."""
        if background is None:
            raise ValueError("background must be given (the reference's getNDValue drops into a debugger)")
        xstart, xend, ystart, yend, zstart, zend = self.comToBounds(com, size)
        H, W = int(dpt.shape[0]), int(dpt.shape[1])
        if xend <= 0 or yend <= 0 or xstart >= W or ystart >= H or xend <= xstart or yend <= ystart:
            raise ValueError("crop window does not intersect the frame")
        wb, hb = xend - xstart, yend - ystart
        if wb > hb:
            sz = (dsize[0], hb * dsize[0] // wb)
        else:
            sz = (wb * dsize[1] // hb, dsize[1])
        xs = int(numpy.floor(dsize[0] / 2. - sz[0] / 2.))
        ys = int(numpy.floor(dsize[1] / 2. - sz[1] / 2.))
        z = (zstart, zend) if thresh_z is True else (-numpy.inf, numpy.inf)
        return self._crop_call(dpt, (xstart, ystart, wb, hb, sz[0], sz[1], xs, ys), z, background, dsize)


def preprocess_real_depth(raw, near=1000, far=3000, fill=10000.0, max_depth=10000.0, out=None):
    """The pre-processing of the reference's real-data loop on the device (eval_model_on_real_data,
    train_cnn_networks_hgru.py:381-386 and the `/ image_max_depth` of :359 / :392): raw 16-bit depth frames in
    millimetres ([N,H,W] torch CUDA tensor of dtype uint16 or int16, the bits read as unsigned) -> float32 frames in
    [0, 1] with everything outside [near, far] replaced by `fill`."""
    if not (torch.is_tensor(raw) and raw.is_cuda and raw.element_size() == 2 and not raw.is_floating_point()):
        raise RuntimeError("raw must be a 16-bit integer torch CUDA tensor (no CPU fallback)")
    raw = raw.contiguous()
    if out is None:
        out = torch.empty(raw.shape, device=raw.device, dtype=torch.float32)
    _lib.check(_lib.load().depth_preprocess_forward(raw.data_ptr(), raw.numel(), int(near), int(far), float(fill),
                                                    float(max_depth), out.data_ptr(), _stream()),
               "depth_preprocess_forward")
    return out


def prepare_data_test(image_np, tr_res, md, config):
    """train_cnn_networks_hgru.py:61-74 on the device: frames in [0,1] ([N,H,W] or [N,H,W,1], torch CUDA)
    and attention outputs tr_res [N,3] -> (patches [N,128,128,1] CUDA, coms, Ms)."""
    if image_np.dim() == 4:
        image_np = image_np[..., 0]
    ts = config.image_target_size
    if torch.is_tensor(tr_res) and tr_res.is_cuda:
        # attention outputs still on the device: the whole stage stays there (coms and Ms come back as CUDA tensors)
        patches, Ms, coms = md.cropArea3D_batch_device(
            image_np, tr=tr_res, tr_scale=(config.image_orig_size[0], config.image_orig_size[1], config.image_max_depth),
            dsize=(ts[1], ts[0]), frame_scale=config.image_max_depth, out_divisor=config.image_max_depth)
        return patches[..., None], coms, Ms
    tr = numpy.asarray(tr_res, numpy.float64)
    scale = numpy.array([config.image_orig_size[0], config.image_orig_size[1], config.image_max_depth], numpy.float64)
    coms = tr * scale
    patches, Ms, coms = md.cropArea3D_batch(image_np, coms, dsize=(ts[1], ts[0]),
                                            frame_scale=config.image_max_depth,
                                            out_divisor=config.image_max_depth)
    return patches[..., None], coms, Ms


def prepare_data(image_np, image_label_shaped, tr_res, md, config, show=False):
    """train_cnn_networks_hgru.py:40-59, the training-time twin of `prepare_data_test`: the same crops (on the device)
    plus the normalised relative joint labels -> (patches [N,128,128,1] CUDA, rel_labels numpy [N, 3J])."""
    if show:
        raise NotImplementedError("plotting is not part of the port")
    patches, coms, _ = prepare_data_test(image_np, tr_res, md, config)
    return patches, md.relative_labels(image_label_shaped, coms)
