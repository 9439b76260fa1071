"""Seeded "random-init" parameters and synthetic depth crops for the hGRU-pose path (numpy, host side).

The reference draws its variables with TensorFlow initialisers at graph-build time
(hgru_module.py:262-503, hgru_pose.py:165-194).  Initialiser *distributions* do not affect parity
(the harness hands the same arrays to the oracle and to the kernels); they are restated here so a
freshly constructed module has the same statistics as the reference's.
"""
import math

import numpy as np

HGRU_PARAM_NAMES = ("p_r", "i_r", "i_b", "o_r", "o_b", "beta", "nu", "gamma", "kappa", "omega",
                    "rho", "lateral_bias")
BN_SCOPES = ("batch_normalization", "batch_normalization_1", "batch_normalization_2",
             "batch_normalization_3", "batch_normalization_4")


def _fans(shape):
    """Fan computation of tf.contrib.layers.xavier_initializer (variance_scaling, FAN_AVG)."""
    shape = tuple(int(s) for s in shape)
    if len(shape) < 1:
        return 1.0, 1.0
    if len(shape) == 1:
        return float(shape[0]), float(shape[0])
    rf = 1.0
    for s in shape[:-2]:
        rf *= s
    return shape[-2] * rf, shape[-1] * rf


def xavier_uniform(rng, shape):
    """xavier_initializer(uniform=True): U(+-sqrt(6/(fan_in+fan_out))).

    This is what `initialization.xavier_initializer(shape, uniform=self.normal_initializer)` evaluates
    to on the configured path (normal_initializer=True, hgru_module.py:32,280).
    """
    fi, fo = _fans(shape)
    lim = math.sqrt(6.0 / (fi + fo))
    return rng.uniform(-lim, lim, size=shape).astype(np.float32)


def _truncated_normal(rng, shape, std):
    out = rng.normal(0.0, std, size=shape)
    bad = np.abs(out) > 2.0 * std
    while bad.any():
        out[bad] = rng.normal(0.0, std, size=int(bad.sum()))
        bad = np.abs(out) > 2.0 * std
    return out.astype(np.float32)


def xavier_normal(rng, shape):
    """xavier_initializer(uniform=False) (hgru_pose.py:171,186): truncated normal, FAN_AVG."""
    fi, fo = _fans(shape)
    std = math.sqrt(1.3 * 2.0 / (fi + fo))
    return _truncated_normal(rng, shape, std)


def hgru_params(k, S=15, T=8, seed=42, stress=1.0):
    """Variables of scope `contextual_circuit` (hgru_module.py:262-503) on the hgru_pose aux path.

    stress > 1 scales p_r so tanh leaves its linear region (SURVEY.md section 8d "stress" set).
    """
    rng = np.random.default_rng(seed)
    vec = (1, 1, 1, k)
    p = {}
    p["p_r"] = xavier_uniform(rng, (S, S, k, k)) * np.float32(stress)      # :299-310
    p["i_r"] = xavier_uniform(rng, (1, 1, k, k))                            # :322-332
    hi = max(T - 1, 1 + 1e-3)
    p["i_b"] = (-np.log(rng.uniform(1.0, hi, size=vec))).astype(np.float32)  # :344-357 (chronos)
    p["o_r"] = xavier_uniform(rng, (1, 1, k, k))                            # :360-370
    p["o_b"] = (-p["i_b"]).astype(np.float32)                               # :382-396
    for name in ("beta", "nu", "gamma", "kappa", "omega", "lateral_bias"):  # :405-503
        p[name] = xavier_uniform(rng, vec)
    p["rho"] = np.ones((T,), np.float32)                                    # :490-493
    return p


def bn_identity(c):
    """tf.layers.batch_normalization at init: gamma=1, beta=0, moving_mean=0, moving_variance=1."""
    return {"gamma": np.ones(c, np.float32), "beta": np.zeros(c, np.float32),
            "moving_mean": np.zeros(c, np.float32), "moving_variance": np.ones(c, np.float32)}


def bn_random(rng, c):
    """A non-trivial BN state (as after training) so folding bugs cannot hide behind identity."""
    return {"gamma": rng.uniform(0.5, 1.5, c).astype(np.float32),
            "beta": rng.uniform(-0.2, 0.2, c).astype(np.float32),
            "moving_mean": rng.uniform(-0.1, 0.1, c).astype(np.float32),
            "moving_variance": rng.uniform(0.5, 1.5, c).astype(np.float32)}


def pose_params(channels=64, S=15, T=8, hw=64, fc_hidden=1024, out=69, seed=42, stress=1.0,
                random_bn=False):
    """All variables of hgru_pose.model (hgru_pose.py:47-105) keyed by the reference's names."""
    rng = np.random.default_rng(seed + 1)
    k = channels
    P = {}
    for name, cin in (("conv_1", 1), ("conv_2", k), ("conv_3", k)):
        P["%s/%s_filters" % (name, name)] = xavier_normal(rng, (3, 3, cin, k))
        P["%s/%s_biases" % (name, name)] = _truncated_normal(rng, (k,), 0.001)
    fc_in = hw * hw * k
    P["fc_1/fc_1_weights"] = xavier_normal(rng, (fc_in, fc_hidden))
    P["fc_1/fc_1_biases"] = _truncated_normal(rng, (fc_hidden,), 0.001)
    P["fc_out/fc_out_weights"] = xavier_normal(rng, (fc_hidden, out))
    P["fc_out/fc_out_biases"] = _truncated_normal(rng, (out,), 0.001)
    for scope, c in zip(BN_SCOPES, (k, k, k, k, fc_hidden)):
        bn = bn_random(rng, c) if random_bn else bn_identity(c)
        for n, v in bn.items():
            P["%s/%s" % (scope, n)] = v
    for n, v in hgru_params(k, S, T, seed, stress).items():
        P["contextual_circuit/" + n] = v
    return P


def hidden_init(shape, seed=7, limit=0.005):
    """O_0: the reference samples xavier-uniform over the activation shape (hgru_module.py:884-887),
    whose limit depends on the batch size; the harness fixes U(+-0.005)."""
    rng = np.random.default_rng(seed)
    return rng.uniform(-limit, limit, size=shape).astype(np.float32)


def synthetic_depth(n, seed=1234, size=128, uniform=False):
    """Synthetic normalised depth crops [n,size,size,1] in [0,1] mimicking
    cropArea3D(...)/image_max_depth (train_cnn_networks_hgru.py:47-50, tf_monkeydetector.py:353-359):
    background 1.0, a centred blob of Kinect-band depths, ~2% zeros inside the blob."""
    rng = np.random.default_rng(seed)
    if uniform:
        return rng.uniform(0.0, 1.0, size=(n, size, size, 1)).astype(np.float32)
    yy, xx = np.mgrid[0:size, 0:size].astype(np.float32)
    out = np.ones((n, size, size, 1), np.float32)
    for i in range(n):
        mask = np.zeros((size, size), bool)
        for _ in range(int(rng.integers(2, 5))):
            cy, cx = rng.uniform(0.35 * size, 0.65 * size, 2)
            ry, rx = rng.uniform(0.15 * size, 0.35 * size, 2)
            mask |= ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0
        c = rng.uniform(1000.0, 3000.0)
        d = (c + rng.uniform(-600.0, 600.0, size=(size, size))) / 10000.0
        d[rng.uniform(size=(size, size)) < 0.02] = 0.0
        img = np.where(mask, d, 1.0).astype(np.float32)
        out[i, :, :, 0] = img
    return out


ATTN_BN_SCOPES = BN_SCOPES[:5] + ("batch_normalization_5",)
ATTN_CONV = (("aconv_1", 3), ("aconv_2", 3), ("aconv_3", 3), ("aconv_4", 3), ("aconv_5", 5))


def attn_params(widths=(64, 128, 256, 512, 1024), fc_hidden=1024, out=3, seed=42, random_bn=False):
    """All variables of attn_model_struct (train_cnn_networks_hgru.py:440-525) keyed by the reference's names:
    xavier-normal filters / weights, truncated_normal(0, .001) biases (:566-592), batch-norm defaults."""
    rng = np.random.default_rng(seed + 11)
    P, cin = {}, 1
    for (name, fs), co in zip(ATTN_CONV, widths):
        P["%s/%s_filters" % (name, name)] = xavier_normal(rng, (fs, fs, cin, co))
        P["%s/%s_biases" % (name, name)] = _truncated_normal(rng, (co,), 0.001)
        cin = co
    P["afc_1/afc_1_weights"] = xavier_normal(rng, (16 * widths[4], fc_hidden))
    P["afc_1/afc_1_biases"] = _truncated_normal(rng, (fc_hidden,), 0.001)
    P["afc_out/afc_out_weights"] = xavier_normal(rng, (fc_hidden, out))
    P["afc_out/afc_out_biases"] = _truncated_normal(rng, (out,), 0.001)
    for scope, c in zip(ATTN_BN_SCOPES, tuple(widths) + (fc_hidden,)):
        bn = bn_random(rng, c) if random_bn else bn_identity(c)
        for n, v in bn.items():
            P["%s/%s" % (scope, n)] = v
    return P
