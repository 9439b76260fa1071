"""A numpy-backed stand-in for the handful of TensorFlow-1.x symbols the reference hot path touches.

Purpose: TensorFlow and Python 2 are absent from this container, so the reference's
`hgru_module.py` / `hgru_pose.py` cannot be imported as they are.  `make_golden.py` loads their
*source text* from /root/reference, applies three mechanical py2->py3 token fixes in memory, and
executes it against this shim, so the golden vectors under tests/golden/ are produced by the
reference's own Python statements (gating order, parameter wiring, loop structure) -- only the
leaf ops (conv2d, sigmoid, tanh, ...) are supplied here, with standard TF semantics, in float64.

This is test tooling.  Nothing in the product imports it, and it is not used on the GPU box.
"""
import contextlib
import types

import numpy as np


class Tensor(np.ndarray):
    """ndarray with the few tf.Tensor methods the reference calls.  In-place operators are
    re-routed to out-of-place ones: in TF `O *= g` rebinds a Python name, it never mutates."""

    def __new__(cls, arr, name=None):
        obj = np.asarray(arr, dtype=np.float64).view(cls)
        obj.name = name or "t:0"
        return obj

    def __array_finalize__(self, obj):
        self.name = getattr(obj, "name", "t:0")

    def get_shape(self):
        return _Shape(self.shape)

    def set_shape(self, shape):
        assert tuple(int(s) for s in shape) == tuple(self.shape), (shape, self.shape)

    def __imul__(self, o):
        return self * o

    def __iadd__(self, o):
        return self + o

    def __isub__(self, o):
        return self - o


class _Shape(list):
    def as_list(self):
        return list(self)


class _Registry(object):
    """Everything the reference code created, so the generator can save it next to the outputs."""

    def __init__(self, seed):
        self.rng = np.random.default_rng(seed)
        self.variables = {}
        self.scope = []
        self.drawn = []          # (shape, array) returned by initialisers, in call order
        self.conv_calls = 0
        self.bn_calls = 0        # tf.layers.batch_normalization scopes: batch_normalization, batch_normalization_1, ...
        self.bn_preset = {}      # scope -> (gamma, beta, moving_mean, moving_variance) injected by the generator
        self.init_gain = {}      # shape -> factor applied to ops.initialization draws of that shape ("stress" sets:
        #                          a wider initial distribution for the 15x15 kernel / the initial state; the
        #                          arithmetic that follows is still the reference's own statements)


REG = _Registry(0)


def reset(seed):
    global REG
    REG = _Registry(seed)
    return REG


def _t(x, name=None):
    return Tensor(x, name)


# ---- tf.* ------------------------------------------------------------------------------------
float32 = np.float32


def identity(x, name=None):
    return _t(np.array(x, copy=True), name)


def constant(v, dtype=None, name=None):
    return _t(v, name)


def ones(shape, dtype=None):
    return _t(np.ones(shape))


def zeros_like(x):
    return _t(np.zeros_like(np.asarray(x)))


def log(x):
    return _t(np.log(np.asarray(x)))


def random_uniform(shape, minval=0.0, maxval=1.0):
    a = REG.rng.uniform(minval, maxval, size=tuple(shape)).astype(np.float32)
    REG.drawn.append((tuple(shape), a))
    return _t(a)


def truncated_normal(shape, mean=0.0, stddev=1.0):
    a = REG.rng.normal(mean, stddev, size=tuple(shape))
    a = np.clip(a, mean - 2 * stddev, mean + 2 * stddev).astype(np.float32)
    REG.drawn.append((tuple(shape), a))
    return _t(a)


def minimum(a, b):
    return _t(np.minimum(a, b))


def maximum(a, b):
    return _t(np.maximum(a, b))


def gather(params, idx, axis=-1):
    return _t(np.take(np.asarray(params), int(np.asarray(idx)), axis=axis))


def transpose(x, perm):
    return _t(np.transpose(np.asarray(x), perm))


def reshape(x, shape):
    return _t(np.reshape(np.asarray(x), shape))


def matmul(a, b):
    return _t(np.asarray(a) @ np.asarray(b))


def concat(xs, axis=-1):
    return _t(np.concatenate([np.asarray(x) for x in xs], axis=axis))


def split(x, n, axis=3):
    return [_t(a) for a in np.split(np.asarray(x), n, axis=axis)]


@contextlib.contextmanager
def variable_scope(name, *a, **k):
    REG.scope.append(name)
    try:
        yield
    finally:
        REG.scope.pop()


def get_variable(name, shape=None, dtype=None, initializer=None, trainable=True):
    """Returns the initial value (there is no session; a variable IS its initial value here)."""
    full = "/".join(REG.scope + [name])
    if callable(initializer):
        val = initializer(shape)
    else:
        val = initializer
    t = _t(np.asarray(val), full + ":0")
    REG.variables[full] = np.asarray(t).astype(np.float32)
    # round-trip through float32 so saved parameters and used parameters are bit-identical
    return _t(REG.variables[full], full + ":0")


class _Graph(object):
    @contextlib.contextmanager
    def gradient_override_map(self, m):
        yield      # backward-only (hgru_module.py:521-535); forward is an ordinary conv2d


def get_default_graph():
    return _Graph()


def while_loop(cond, body, loop_vars, back_prop=True, swap_memory=False):
    vals = list(loop_vars)
    while bool(np.asarray(cond(*vals))):
        vals = list(body(*vals))
    return vals


class TensorArray(object):
    def __init__(self, dtype, size):
        self.items = [None] * size

    def write(self, i, v):
        self.items[int(np.asarray(i))] = np.array(v, copy=True)
        return self

    def stack(self):
        return _t(np.stack(self.items, axis=0))


# ---- tf.nn -----------------------------------------------------------------------------------
def _conv2d(data, weights, strides, padding="SAME"):
    """tf.nn.conv2d NHWC x HWIO, stride 1, SAME: zero-padded cross-correlation (no kernel flip)."""
    assert list(strides) == [1, 1, 1, 1] and padding == "SAME"
    x = np.asarray(data, np.float64)
    w = np.asarray(weights, np.float64)
    REG.conv_calls += 1
    n, h, wd, ci = x.shape
    fh, fw, _, co = w.shape
    pt, pl = (fh - 1) // 2, (fw - 1) // 2
    xp = np.pad(x, ((0, 0), (pt, fh - 1 - pt), (pl, fw - 1 - pl), (0, 0)))
    cols = np.lib.stride_tricks.sliding_window_view(xp, (fh, fw), axis=(1, 2))  # n,h,w,ci,fh,fw
    return _t(np.einsum("nhwcyx,yxco->nhwo", cols, w, optimize=True))


def _max_pool(x, ksize, strides, padding="SAME", name=None):
    assert list(ksize) == [1, 2, 2, 1] and list(strides) == [1, 2, 2, 1]
    a = np.asarray(x)
    n, h, w, c = a.shape
    assert h % 2 == 0 and w % 2 == 0
    return _t(a.reshape(n, h // 2, 2, w // 2, 2, c).max(axis=(2, 4)))


nn = types.SimpleNamespace(
    tanh=lambda x: _t(np.tanh(np.asarray(x))),
    sigmoid=lambda x: _t(1.0 / (1.0 + np.exp(-np.asarray(x)))),
    relu=lambda x: _t(np.maximum(np.asarray(x), 0.0)),
    selu=None, leaky_relu=None,
    conv2d=_conv2d,
    max_pool=_max_pool,
    bias_add=lambda x, b: _t(np.asarray(x) + np.asarray(b)),
    dropout=None,        # train-mode only: never reached by the inference-mode golden runs
)


# ---- tf.layers.batch_normalization (inference: moving statistics) -----------------------------------
def _batch_normalization(inputs, axis=-1, momentum=0.99, epsilon=1e-3, center=True, scale=True, training=False,
                         fused=None, name=None):
    """Inference-mode tf.layers.batch_normalization.  Variables live in scopes `batch_normalization`,
    `batch_normalization_1`, ... in call order (TF's default naming); values come from REG.bn_preset or the TF
    defaults (gamma 1, beta 0, mean 0, variance 1).  `axis` beyond the rank (the reference passes axis=3 for the
    rank-2 fc tensor: defect D5) is resolved to the last axis (R-D5)."""
    assert training in (False, None), "golden vectors are inference-mode (moving statistics)"
    x = np.asarray(inputs, np.float64)
    scope = "batch_normalization" if REG.bn_calls == 0 else "batch_normalization_%d" % REG.bn_calls
    REG.bn_calls += 1
    ax = axis if -x.ndim <= axis < x.ndim else x.ndim - 1
    c = x.shape[ax]
    g, b, m, v = REG.bn_preset.get(scope, (np.ones(c), np.zeros(c), np.zeros(c), np.ones(c)))
    full = "/".join(REG.scope + [scope])
    for nm, val in (("gamma", g), ("beta", b), ("moving_mean", m), ("moving_variance", v)):
        REG.variables[full + "/" + nm] = np.asarray(val, np.float32)
    shp = [1] * x.ndim
    shp[ax] = c
    g, b, m, v = [np.asarray(REG.variables[full + "/" + nm], np.float64).reshape(shp)
                  for nm in ("gamma", "beta", "moving_mean", "moving_variance")]
    return _t((x - m) / np.sqrt(v + epsilon) * g + b)


layers = types.SimpleNamespace(batch_normalization=_batch_normalization)


# ---- tf.image.resize_images (TF 1.x default: bilinear, align_corners=False, no half-pixel centres) ----
def _resize_images(images, size, method=0, align_corners=False):
    """ResizeBilinear as TF 1.x computes it (kernels/resize_bilinear_op.cc), in float32: scale = in / out,
    src = dst * scale, lower = floor(src), upper = min(lower + 1, in - 1), lerp = src - lower;
    top = tl + (tr - tl) * xl; bottom = bl + (br - bl) * xl; out = top + (bottom - top) * yl."""
    assert method == 0 and not align_corners
    x = np.asarray(images, np.float32)
    n, h, w, c = x.shape
    oh, ow = int(size[0]), int(size[1])
    if (oh, ow) == (h, w):
        return _t(x.astype(np.float64))

    def weights(insz, outsz):
        scale = np.float32(insz) / np.float32(outsz)
        src = np.arange(outsz, dtype=np.float32) * scale
        lo = np.floor(src).astype(np.int64)
        hi = np.minimum(lo + 1, insz - 1)
        return lo, hi, (src - lo.astype(np.float32)).astype(np.float32)

    y0, y1, yl = weights(h, oh)
    x0, x1, xl = weights(w, ow)
    xl_ = xl.reshape(1, 1, ow, 1)
    yl_ = yl.reshape(1, oh, 1, 1)
    tl, tr = x[:, y0][:, :, x0], x[:, y0][:, :, x1]
    bl, br = x[:, y1][:, :, x0], x[:, y1][:, :, x1]
    top = (tl + (tr - tl) * xl_).astype(np.float32)
    bot = (bl + (br - bl) * xl_).astype(np.float32)
    out = (top + (bot - top) * yl_).astype(np.float32)
    return _t(out.astype(np.float64))


image = types.SimpleNamespace(resize_images=_resize_images)


# ---- tf.contrib.layers.xavier_initializer* (hgru_pose.py:171,186) ----------------------------
def _xavier(uniform=True):
    def init(shape):
        shape = tuple(int(s) for s in shape)
        rf = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
        fi, fo = shape[-2] * rf, shape[-1] * rf
        if uniform:
            lim = np.sqrt(6.0 / (fi + fo))
            a = REG.rng.uniform(-lim, lim, size=shape)
        else:
            std = np.sqrt(1.3 * 2.0 / (fi + fo))
            a = np.clip(REG.rng.normal(0, std, size=shape), -2 * std, 2 * std)
        a = a.astype(np.float32)
        REG.drawn.append((shape, a))
        return a
    return init


contrib = types.SimpleNamespace(layers=types.SimpleNamespace(
    xavier_initializer=_xavier, xavier_initializer_conv2d=_xavier))


# ---- the two un-vendored helper modules (hgru_module.py:4-5) ---------------------------------
py_utils = types.SimpleNamespace(ifloor=lambda x: int(np.floor(x)), iceil=lambda x: int(np.ceil(x)))


def _xavier_initializer(shape, uniform=True, mask=None):
    """ops.initialization.xavier_initializer: returns a *tensor* (call sites hgru_module.py:278...)."""
    a = _xavier(uniform)(shape)
    g = REG.init_gain.get(tuple(int(s) for s in shape))
    if g is not None:
        a *= np.float32(g)          # in place: REG.drawn holds this same array
    return _t(a)


initialization = types.SimpleNamespace(xavier_initializer=_xavier_initializer)
