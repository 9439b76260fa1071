"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the hGRU-pose forward path (numpy, float64).

This file restates, in plain numpy, the arithmetic of the reference's hot path so the CUDA
kernels can be checked against it.  It is never imported by the product package
(`monkey_pose_b200`); only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs
may use it.

Pinning status: the reference ships no tests / golden vectors and cannot run
as committed (TensorFlow 1.x + Python 2, un-vendored imports).  The recurrent layer restated here
IS pinned against the reference's own `hgru_module.py` source, executed in this container under a
numpy-backed `tensorflow` shim (`tests/golden/make_golden.py` -> `tests/golden/hgru_ref_*.npz`,
checked by `tests/test_oracle_golden.py`).  The composed `hgru_pose.model.build` graph cannot be
executed even under the shim (reference defects D4-D6), so the model-level
composition is "parity unpinned" beyond its individually pinned layers (conv_layer, max_pool,
fc_layer, hgru_layer) and uses the documented resolutions R-D4/R-D5/R-D6.

All tensors are NHWC, weights HWIO, exactly as in the reference (TF conventions).
Every function cites the reference file:line it follows (paths relative to /root/reference).
"""
import numpy as np

F64 = np.float64


# ----------------------------------------------------------------------------------------------
# primitive ops (TF semantics)
# ----------------------------------------------------------------------------------------------
def conv2d_same(x, w):
    """tf.nn.conv2d(x, w, [1,1,1,1], 'SAME'): stride-1 zero-padded cross-correlation (no flip).

    x [N,H,W,Ci], w [fh,fw,Ci,Co] -> [N,H,W,Co].  Used at hgru_module.py:531-535, 544-548 and
    hgru_pose.py:146.  For odd f the SAME padding is (f-1)/2 on each side.
    """
    x = np.asarray(x, F64)
    w = np.asarray(w, F64)
    n, h, wd, ci = x.shape
    fh, fw, ci2, co = w.shape
    assert ci == ci2
    # TF SAME, stride 1: pad_total = f-1, pad_before = pad_total // 2
    pt, pl = (fh - 1) // 2, (fw - 1) // 2
    pb, pr = (fh - 1) - pt, (fw - 1) - pl
    xp = np.zeros((n, h + pt + pb, wd + pl + pr, ci), F64)
    xp[:, pt:pt + h, pl:pl + wd, :] = x
    out = np.zeros((n, h, wd, co), F64)
    for dy in range(fh):
        for dx in range(fw):
            out += xp[:, dy:dy + h, dx:dx + wd, :] @ w[dy, dx]
    return out


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def max_pool_2x2(x):
    """tf.nn.max_pool(ksize 2x2, stride 2, SAME) -- hgru_pose.py:134-137 (even H, W: no padding)."""
    n, h, w, c = x.shape
    assert h % 2 == 0 and w % 2 == 0
    return x.reshape(n, h // 2, 2, w // 2, 2, c).max(axis=(2, 4))


def pool_same(x, k, average=False):
    """tf.nn.max_pool / tf.nn.avg_pool with ksize = stride = k, SAME -- the model's max_pool_4 (k = 4) and avg_pool
    (k = 2) helpers, hgru_pose.py:124-132.  TF pads out*k - H in all, half of it (rounded down) before; padding never
    wins a max and is left out of an average's count."""
    x = np.asarray(x, F64)
    n, h, w, c = x.shape
    ho, wo = -(-h // k), -(-w // k)
    py, px = (ho * k - h) // 2, (wo * k - w) // 2
    out = np.zeros((n, ho, wo, c), F64)
    for yo in range(ho):
        for xo in range(wo):
            y0, y1 = max(yo * k - py, 0), min(yo * k - py + k, h)
            x0, x1 = max(xo * k - px, 0), min(xo * k - px + k, w)
            win = x[:, y0:y1, x0:x1, :]
            out[:, yo, xo, :] = win.mean(axis=(1, 2)) if average else win.max(axis=(1, 2))
    return out


def batchnorm_moments0(x, eps=1e-3):
    """The model's `batchnorm` helper (hgru_pose.py:120-122): tf.nn.moments(layer, [0]) and
    tf.nn.batch_normalization(layer, m, v, None, None, 1e-3)."""
    x = np.asarray(x, F64)
    m = x.mean(axis=0)
    v = ((x - m) ** 2).mean(axis=0)
    return (x - m) / np.sqrt(v + eps)


def batch_norm_inference(x, gamma, beta, mean, var, eps=1e-5):
    """tf.layers.batch_normalization(training=False) over the last axis -- hgru_pose.py:52-60 etc."""
    return (np.asarray(x, F64) - mean) / np.sqrt(np.asarray(var, F64) + eps) * gamma + beta


def batch_norm_training(x, gamma, beta, eps=1e-5):
    """training=True: batch statistics over all axes but the last (biased variance, as TF fused BN)."""
    x = np.asarray(x, F64)
    ax = tuple(range(x.ndim - 1))
    mean = x.mean(axis=ax)
    var = x.var(axis=ax)
    return (x - mean) / np.sqrt(var + eps) * gamma + beta


def dropout_keep_mask(n, keep, seed):
    """Keep mask of the training-mode dropout (tf.nn.dropout(x, keep), hgru_pose.py:93-94).  TensorFlow's random stream
    cannot be reproduced, so the C ABI documents its own counter-based mask (include/hgru_b200.h,
    layer_batch_norm_forward): element i is kept iff (splitmix64(seed ^ i * 0xD1B54A32D192ED03) >> 40) / 2^24 < keep.
    This is that definition in numpy (uint64 arithmetic wraps)."""
    with np.errstate(over="ignore"):
        i = np.arange(n, dtype=np.uint64)
        z = np.uint64(seed) ^ (i * np.uint64(0xD1B54A32D192ED03))
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    u = (z >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    return u < np.float32(keep)


def batch_norm_moving_update(x, moving_mean, moving_var, momentum=0.997):
    """What the reference's UPDATE_OPS assign in training mode (train_cnn_networks_hgru.py:123-126): moving statistics
    <- moving * momentum + batch * (1 - momentum), batch variance UNBIASED (TF's fused batch norm)."""
    x = np.asarray(x, F64)
    ax = tuple(range(x.ndim - 1))
    n = x.size // x.shape[-1]
    mean, var = x.mean(axis=ax), x.var(axis=ax) * (n / max(n - 1.0, 1.0))
    return (np.asarray(moving_mean, F64) * momentum + mean * (1 - momentum),
            np.asarray(moving_var, F64) * momentum + var * (1 - momentum))


# ----------------------------------------------------------------------------------------------
# hGRU recurrent layer  (hgru_module.py)
# ----------------------------------------------------------------------------------------------
HGRU_PARAM_NAMES = ("p_r", "i_r", "i_b", "o_r", "o_b", "beta", "nu", "gamma", "kappa", "omega",
                    "rho", "lateral_bias")


def hgru_step(X, H2, p, t):
    """One timestep = ContextualCircuit.full (hgru_module.py:825-857) on the configured path
    (hgru_pose.py:20-39 aux: tanh, gru_gates=True, multiplicative_excitation, gamma, adapation).

    Returns (H1, H2_new).
    """
    # circuit_input (hgru_module.py:692-724)
    G1 = sigmoid(conv2d_same(H2, p["i_r"]) + p["i_b"])          # :696-707
    Og = H2 * G1                                                 # :709-711 (local; caller's O unchanged)
    C1 = conv2d_same(Og, p["p_r"]) + p["lateral_bias"]           # :714-718, 657
    # input_integration (hgru_module.py:795-804), xi = 1 (:463-464), gru_gates => no mixing
    H1 = np.tanh(X - (p["beta"] * H2 + p["nu"]) * C1)
    # circuit_output (hgru_module.py:726-756)
    G2 = sigmoid(conv2d_same(H1, p["o_r"]) + p["o_b"])          # :729-740
    C2 = conv2d_same(H1, p["p_r"]) + p["lateral_bias"]           # :746-750, 657
    # output_integration (hgru_module.py:806-823), zeta = 1 (:437-438)
    e = p["gamma"] * C2
    a = p["kappa"] * (H1 + e)
    m = p["omega"] * (H1 * e)
    Ht = np.tanh(a + m)
    H2n = G2 * H2 + (1.0 - G2) * Ht
    # adapation (hgru_module.py:847-849)
    H2n = H2n * p["rho"][t]
    return H1, H2n


def hgru_forward(X, H2_init, params, timesteps, trace=False):
    """ContextualCircuit.build (hgru_module.py:872-959): `timesteps` iterations of `full`.

    X, H2_init [N,H,W,k]; params: dict with HGRU_PARAM_NAMES (p_r [S,S,k,k], i_r/o_r [1,1,k,k],
    vectors [1,1,1,k], rho [T]).  Returns H2 after the last step (and per-step H1/H2 lists).
    The initial `I` of the reference is dead when gru_gates=True.
    """
    p = {k: np.asarray(v, F64) for k, v in params.items()}
    X = np.asarray(X, F64)
    H2 = np.asarray(H2_init, F64)
    H1s, H2s = [], []
    for t in range(timesteps):
        H1, H2 = hgru_step(X, H2, p, t)
        if trace:
            H1s.append(H1)
            H2s.append(H2)
    if trace:
        return H2, H1s, H2s
    return H2


# ----------------------------------------------------------------------------------------------
# hgru_pose.model.build  (hgru_pose.py:47-105) with resolutions R-D4, R-D5, R-D6
# ----------------------------------------------------------------------------------------------
BN_SCOPES = ("batch_normalization", "batch_normalization_1", "batch_normalization_2",
             "batch_normalization_3", "batch_normalization_4")


def _bn(x, params, scope, train_mode, eps):
    g, b = params[scope + "/gamma"], params[scope + "/beta"]
    if train_mode:
        return batch_norm_training(x, g, b, eps)
    return batch_norm_inference(x, g, b, params[scope + "/moving_mean"],
                                params[scope + "/moving_variance"], eps)


def conv_layer(x, filt, bias):
    """model.conv_layer (hgru_pose.py:139-154): relu(conv2d SAME + bias)."""
    return np.maximum(conv2d_same(x, filt) + np.asarray(bias, F64), 0.0)


def fc_layer(x, w, b):
    """model.fc_layer (hgru_pose.py:156-163): reshape(x,[-1,in]) @ W + b (flatten order h,w,c)."""
    x = np.asarray(x, F64)
    return x.reshape(x.shape[0], -1) @ np.asarray(w, F64) + np.asarray(b, F64)


def pose_forward(depth, params, H2_init, timesteps=8, train_mode=False, eps=1e-5, trace=False,
                 dropout_keep=None, dropout_seed=0):
    """hgru_pose.model.build (hgru_pose.py:47-105).

    depth [N,128,128,1]; params: dict keyed by the reference's variable names
    (`conv_1/conv_1_filters`, ..., `contextual_circuit/p_r`, ..., `fc_out/fc_out_biases`,
    `batch_normalization[_i]/{gamma,beta,moving_mean,moving_variance}`).
    Dropout (hgru_pose.py:93-94) is random: excluded unless `dropout_keep` is given, in which case the documented
    counter-based mask (dropout_keep_mask) is applied to relu(fc1) before the batch norm, as the reference orders them.
    H2_init: the initial state O_0 [N,64,64,k], or the reference's hidden_init names 'identity' (O_0 = X, the hGRU's
    input conv3: hgru_module.py:876-878) / 'zeros' (:888-890).
    """
    P = params
    acts = {}
    x = np.asarray(depth, F64)
    conv1 = conv_layer(x, P["conv_1/conv_1_filters"], P["conv_1/conv_1_biases"])       # :50
    pool1 = _bn(max_pool_2x2(conv1), P, BN_SCOPES[0], train_mode, eps)                   # :51-60
    conv2 = _bn(conv_layer(pool1, P["conv_2/conv_2_filters"], P["conv_2/conv_2_biases"]),
                P, BN_SCOPES[1], train_mode, eps)                                        # :61-70
    conv3 = _bn(conv_layer(conv2, P["conv_3/conv_3_filters"], P["conv_3/conv_3_biases"]),
                P, BN_SCOPES[2], train_mode, eps)                                        # :71-80
    hp = {n: P["contextual_circuit/" + n] for n in HGRU_PARAM_NAMES}
    if isinstance(H2_init, str):
        if H2_init not in ("identity", "zeros"):
            raise RuntimeError("hidden_init")                                            # hgru_module.py:891-892
        H2_init = conv3.copy() if H2_init == "identity" else np.zeros_like(conv3)
    hgru = hgru_forward(conv3, H2_init, hp, timesteps)                                   # :81 (R-D4)
    hgru_bn = _bn(hgru, P, BN_SCOPES[3], train_mode, eps)                                # :82-90
    fc1 = fc_layer(hgru_bn, P["fc_1/fc_1_weights"], P["fc_1/fc_1_biases"])               # :91
    r = np.maximum(fc1, 0.0)                                                             # :92
    if dropout_keep is not None and dropout_keep < 1.0:                                  # :93-94
        keep = dropout_keep_mask(r.size, dropout_keep, dropout_seed).reshape(r.shape)
        r = np.where(keep, r / np.float64(np.float32(dropout_keep)), 0.0)
    relu1 = _bn(r, P, BN_SCOPES[4], train_mode, eps)                                     # :95-103 (R-D5)
    out = fc_layer(relu1, P["fc_out/fc_out_weights"], P["fc_out/fc_out_biases"])         # :104 (R-D6)
    if trace:
        acts.update(conv1=conv1, pool1=pool1, conv2=conv2, conv3=conv3, hgru=hgru,
                    hgru_bn=hgru_bn, fc1=fc1, relu1=relu1, out_put=out)
        return out, acts
    return out


# ----------------------------------------------------------------------------------------------
# metric  (pose_evaluation.py:10-15, train_cnn_networks_hgru.py:154-156)
# ----------------------------------------------------------------------------------------------
def mean_joint_error_mm(out_a, out_b, cube_z=1200.0):
    """getMeanError_np on outputs reshaped [N,23,3] and scaled by cube[2]/2 (= 600 mm)."""
    a = np.asarray(out_a, F64).reshape(out_a.shape[0], -1, 3) * (cube_z / 2.0)
    b = np.asarray(out_b, F64).reshape(out_b.shape[0], -1, 3) * (cube_z / 2.0)
    return float(np.nanmean(np.nanmean(np.sqrt(np.square(a - b).sum(axis=2)), axis=1)))


def rel_err(a, b):
    """max|a-b| / max|b| and l2 relative error (SURVEY.md section 8d parity metrics)."""
    a = np.asarray(a, F64)
    b = np.asarray(b, F64)
    d = a - b
    return float(np.abs(d).max() / max(np.abs(b).max(), 1e-30)), \
        float(np.linalg.norm(d.ravel()) / max(np.linalg.norm(b.ravel()), 1e-30))
