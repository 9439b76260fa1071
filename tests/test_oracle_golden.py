"""CPU: the oracle (numpy float64 arbiter, torch float32 full-size checker) against the golden
vectors produced by executing the reference's own source (tests/golden/make_golden.py)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import hgru_oracle_np as onp
from oracle import hgru_oracle_torch as otorch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
HGRU_FILES = sorted(glob.glob(os.path.join(GOLDEN, "hgru_ref_*.npz")))


def _params(z, scope="contextual_circuit/"):
    return {n: z["var:" + scope + n] for n in onp.HGRU_PARAM_NAMES}


def _atol(z):
    """The small sets store the reference's float64 outputs; the BASELINE-width sets (25 / 32 channels, T = 8 / 16)
    store them rounded to float32 (half an ulp of values below 4: 2.4e-7)."""
    return 1e-12 if z["O_steps"].dtype == np.float64 else 2.5e-7


def test_golden_files_present():
    assert len(HGRU_FILES) >= 6
    assert os.path.exists(os.path.join(GOLDEN, "pose_layers_ref.npz"))


@pytest.mark.parametrize("path", HGRU_FILES, ids=[os.path.basename(p) for p in HGRU_FILES])
def test_numpy_oracle_matches_reference_every_timestep(path):
    z = np.load(path)
    T = int(z["T"])
    out, H1s, H2s = onp.hgru_forward(z["X"], z["O0"], _params(z), T, trace=True)
    for t in range(T):
        np.testing.assert_allclose(H1s[t], z["I_steps"][:, t], rtol=0, atol=_atol(z))
        np.testing.assert_allclose(H2s[t], z["O_steps"][:, t], rtol=0, atol=_atol(z))
    np.testing.assert_allclose(out, z["O_final"], rtol=0, atol=_atol(z))


@pytest.mark.parametrize("path", HGRU_FILES, ids=[os.path.basename(p) for p in HGRU_FILES])
def test_initial_I_is_dead_on_configured_path(path):
    """gru_gates=True: the reference's I_0 draw never reaches the output."""
    z = np.load(path)
    a = onp.hgru_forward(z["X"], z["O0"], _params(z), int(z["T"]))
    np.testing.assert_allclose(a, z["O_final"], rtol=0, atol=_atol(z))   # oracle takes no I_0 at all


@pytest.mark.parametrize("path", HGRU_FILES, ids=[os.path.basename(p) for p in HGRU_FILES])
def test_torch_oracle_matches_reference(path):
    z = np.load(path)
    T = int(z["T"])
    out, H1s, H2s = otorch.hgru_forward(z["X"], z["O0"], _params(z), T, trace=True)
    for t in range(T):
        assert onp.rel_err(H1s[t].numpy(), z["I_steps"][:, t])[0] < 2e-6
        assert onp.rel_err(H2s[t].numpy(), z["O_steps"][:, t])[0] < 2e-6
    assert onp.rel_err(out.numpy(), z["O_final"])[0] < 2e-6


def test_weights_dict_keys_of_reference():
    """`build()` returns (O, weights, activities) with these keys (hgru_module.py:863-870,939-954)."""
    z = np.load(HGRU_FILES[0])
    assert set(z["weights_keys"].tolist()) == {
        "P_r", "I_r", "O_r", "xi_r", "beta_r", "nu_r", "zeta_r", "gamma_r", "kappa_r", "rho_r", "p_t"}


def test_pose_layers_match_reference():
    z = np.load(os.path.join(GOLDEN, "pose_layers_ref.npz"))
    c1 = onp.conv_layer(z["x"], z["var:conv_1/conv_1_filters"], z["var:conv_1/conv_1_biases"])
    np.testing.assert_allclose(c1, z["conv1"], rtol=0, atol=1e-12)
    p1 = onp.max_pool_2x2(c1)
    np.testing.assert_allclose(p1, z["pool1"], rtol=0, atol=1e-12)
    c2 = onp.conv_layer(p1, z["var:conv_2/conv_2_filters"], z["var:conv_2/conv_2_biases"])
    np.testing.assert_allclose(c2, z["conv2"], rtol=0, atol=1e-12)
    fc = onp.fc_layer(c2, z["var:fc_1/fc_1_weights"], z["var:fc_1/fc_1_biases"])
    np.testing.assert_allclose(fc, z["fc1"], rtol=0, atol=1e-12)
    hg = onp.hgru_forward(c2, z["hgru_O0"], _params(z), 8)
    np.testing.assert_allclose(hg, z["hgru_O"], rtol=0, atol=1e-12)
    assert set(z["var_dict_keys"].tolist()) == {
        "conv_1|0", "conv_1|1", "conv_2|0", "conv_2|1", "fc_1|0", "fc_1|1"}


def test_numpy_and_torch_pose_forward_agree():
    from monkey_pose_b200 import initialization as init
    P = init.pose_params(channels=8, S=5, T=3, hw=8, fc_hidden=32, out=69, seed=3, stress=4.0,
                         random_bn=True)
    depth = init.synthetic_depth(2, seed=0, size=16)
    h0 = init.hidden_init((2, 8, 8, 8), seed=5)
    a, acts = onp.pose_forward(depth, P, h0, timesteps=3, trace=True)
    b = otorch.pose_forward(depth, P, h0, timesteps=3).numpy()
    assert a.shape == (2, 69)
    assert np.abs(acts["hgru"]).max() > 1e-3
    assert onp.rel_err(b, a)[0] < 1e-5
    assert onp.mean_joint_error_mm(b, a) < 1e-2
    # the reference's named initial states (hgru_module.py:876-878, 888-890): 'identity' is O_0 = conv3
    for name, explicit in (("identity", acts["conv3"]), ("zeros", np.zeros_like(acts["conv3"]))):
        n = onp.pose_forward(depth, P, name, timesteps=3)
        assert np.array_equal(n, onp.pose_forward(depth, P, explicit, timesteps=3))
        assert onp.rel_err(otorch.pose_forward(depth, P, name, timesteps=3, dtype=torch.float64).numpy(), n)[0] < 1e-12
    with pytest.raises(RuntimeError):
        onp.pose_forward(depth, P, "bogus", timesteps=3)


def test_metric_restatement():
    """getMeanError_np (pose_evaluation.py:10-15) on x600 mm scaling (train_cnn_networks_hgru.py:154-156)."""
    a = np.zeros((2, 69))
    b = np.zeros((2, 69))
    b[:, 0:3] = [3.0 / 600, 4.0 / 600, 0.0]        # one joint off by 5 mm in each frame
    assert abs(onp.mean_joint_error_mm(a, b) - 5.0 / 23) < 1e-12


def test_training_mode_helpers_of_the_oracle():
    """Dropout keep mask (the counter-based definition documented in include/hgru_b200.h) and the moving-statistics
    update: deterministic, seed-dependent, the right keep fraction; batch-norm training = inference with the batch's
    own statistics."""
    m1 = onp.dropout_keep_mask(200000, 0.7, 1234)
    m2 = onp.dropout_keep_mask(200000, 0.7, 1234)
    m3 = onp.dropout_keep_mask(200000, 0.7, 1235)
    assert np.array_equal(m1, m2) and not np.array_equal(m1, m3)
    assert abs(m1.mean() - 0.7) < 0.005 and abs(m3.mean() - 0.7) < 0.005
    assert onp.dropout_keep_mask(1000, 1.0, 5).all()
    # element i of a longer mask equals element i of a shorter one (counter-based)
    assert np.array_equal(onp.dropout_keep_mask(1000, 0.3, 9), onp.dropout_keep_mask(5000, 0.3, 9)[:1000])
    rng = np.random.default_rng(0)
    x = rng.normal(1.5, 2.0, size=(7, 5, 3, 4))
    g, b = rng.uniform(0.5, 1.5, 4), rng.uniform(-1, 1, 4)
    mean, var = x.mean(axis=(0, 1, 2)), x.var(axis=(0, 1, 2))
    np.testing.assert_allclose(onp.batch_norm_training(x, g, b), onp.batch_norm_inference(x, g, b, mean, var), atol=1e-12)
    mm, mv = onp.batch_norm_moving_update(x, np.zeros(4), np.ones(4), momentum=0.997)
    n = 7 * 5 * 3
    np.testing.assert_allclose(mm, 0.003 * mean, atol=1e-15)
    np.testing.assert_allclose(mv, 0.997 + 0.003 * var * n / (n - 1), atol=1e-15)
    # the model-level oracle with train_mode uses them
    P = {k: v for k, v in __import__("monkey_pose_b200").initialization.pose_params(
        channels=4, S=3, T=1, hw=4, fc_hidden=8, out=6, seed=1, random_bn=True).items()}
    d = rng.uniform(0, 1, size=(3, 8, 8, 1)).astype(np.float32)
    h0 = np.zeros((3, 4, 4, 4), np.float32)
    a = onp.pose_forward(d, P, h0, timesteps=1, train_mode=True)
    bb = onp.pose_forward(d, P, h0, timesteps=1, train_mode=True, dropout_keep=0.7, dropout_seed=3)
    c = onp.pose_forward(d, P, h0, timesteps=1, train_mode=False)
    assert a.shape == (3, 6) and not np.allclose(a, bb) and not np.allclose(a, c)


def test_same_pooling_and_batch_moments_known_answers():
    """TensorFlow's SAME pooling (hgru_pose.py:124-137): pad_total = out*k - H, half of it before; padding never wins
    a max and is not counted by an average; `batchnorm` (:120-122) normalises over the batch axis."""
    x = np.arange(1, 1 + 5 * 6, dtype=np.float64).reshape(1, 5, 6, 1)
    assert np.array_equal(onp.pool_same(x, 2)[0, :, :, 0], onp.max_pool_2x2(np.pad(x, ((0, 0), (0, 1), (0, 0), (0, 0)),
                                                                                    constant_values=-1))[0, :, :, 0])
    p4 = onp.pool_same(x, 4)[0, :, :, 0]            # 5 -> 2 rows (pad 3: 1 before), 6 -> 2 columns (pad 2: 1 before)
    assert p4.shape == (2, 2)
    assert p4[0, 0] == x[0, 2, 2, 0] and p4[1, 1] == x[0, 4, 5, 0] and p4[0, 1] == x[0, 2, 5, 0]
    a2 = onp.pool_same(x, 2, average=True)[0, :, :, 0]
    assert a2[0, 0] == x[0, :2, :2, 0].mean() and a2[2, 0] == x[0, 4, :2, 0].mean()      # last row: 2 values, not 4
    b = onp.batchnorm_moments0(np.array([[1.0, 10.0], [3.0, 10.0]]), eps=0.0 + 1e-3)
    assert abs(b[0, 0] + 1 / np.sqrt(1 + 1e-3)) < 1e-12 and b[0, 1] == 0.0


def test_leaf_convolution_against_a_second_independent_implementation():
    """The oracle's conv2d_same (numpy) and the torch-CPU oracle agree with scipy.signal.correlate2d -- a third party's
    definition of a zero-padded 'same' cross-correlation -- for odd filters, several channels and non-square images."""
    signal = pytest.importorskip("scipy.signal")
    rng = np.random.default_rng(8)
    for (h, w, ci, co, f) in ((9, 7, 3, 2, 3), (16, 11, 2, 3, 15), (5, 5, 1, 1, 1), (12, 20, 4, 4, 7)):
        x = rng.standard_normal((2, h, w, ci))
        k = rng.standard_normal((f, f, ci, co))
        want = np.zeros((2, h, w, co))
        for n in range(2):
            for o in range(co):
                for i in range(ci):
                    want[n, :, :, o] += signal.correlate2d(x[n, :, :, i], k[:, :, i, o], mode="same", boundary="fill")
        got = onp.conv2d_same(x, k)
        assert np.abs(got - want).max() < 1e-11


def test_leaf_normalisation_and_pooling_against_torch_functional():
    """tf.layers.batch_normalization (inference and training forward, moving-statistics update) and SAME pooling as
    the numpy oracle restates them, against torch.nn.functional's independent implementations."""
    import torch.nn.functional as F
    rng = np.random.default_rng(9)
    x = rng.standard_normal((4, 6, 10, 5))
    gamma, beta = rng.uniform(0.5, 1.5, 5), rng.standard_normal(5)
    mean, var = rng.standard_normal(5), rng.uniform(0.5, 2.0, 5)
    xt = torch.as_tensor(x).permute(0, 3, 1, 2)
    t = lambda a: torch.as_tensor(a)                                            # noqa: E731
    want = F.batch_norm(xt, t(mean), t(var), t(gamma), t(beta), training=False, eps=1e-5).permute(0, 2, 3, 1).numpy()
    assert np.abs(onp.batch_norm_inference(x, gamma, beta, mean, var, eps=1e-5) - want).max() < 1e-12
    rm, rv = t(mean.copy()), t(var.copy())
    want = F.batch_norm(xt, rm, rv, t(gamma), t(beta), training=True, momentum=1 - 0.997, eps=1e-5)
    got = onp.batch_norm_training(x, gamma, beta, eps=1e-5)
    got = got[0] if isinstance(got, tuple) else got
    assert np.abs(got - want.permute(0, 2, 3, 1).numpy()).max() < 1e-12
    # pooling: 2x2 / stride 2 on even sizes, and the general SAME rule on odd ones (torch: explicit -inf padding)
    assert np.array_equal(onp.max_pool_2x2(x), F.max_pool2d(xt, 2, 2).permute(0, 2, 3, 1).numpy())
    y = rng.standard_normal((2, 7, 9, 3))
    yt = torch.as_tensor(y).permute(0, 3, 1, 2)
    for k in (2, 4):
        ho, wo = -(-7 // k), -(-9 // k)
        py, px = ho * k - 7, wo * k - 9
        padded = F.pad(yt, (px // 2, px - px // 2, py // 2, py - py // 2), value=float("-inf"))
        assert np.array_equal(onp.pool_same(y, k), F.max_pool2d(padded, k, k).permute(0, 2, 3, 1).numpy())
    ones = F.pad(torch.ones_like(yt), (0, 1, 0, 1))
    avg = F.avg_pool2d(F.pad(yt, (0, 1, 0, 1)), 2, 2) * 4 / (F.avg_pool2d(ones, 2, 2) * 4)
    assert np.abs(onp.pool_same(y, 2, average=True) - avg.permute(0, 2, 3, 1).numpy()).max() < 1e-12
