"""GPU (B200): the reference model's layer methods and attribute surface (hgru_pose.py:47-216), the training-mode
forward (batch statistics + dropout, what the reference's only live caller builds: train_cnn_networks_hgru.py:142),
the `adapation` switch of the circuit, checkpoint import feeding the device path, and plan-cache housekeeping."""
import os

import numpy as np
import pytest
import torch

import monkey_pose_b200 as mp
from monkey_pose_b200 import initialization as init
from oracle import hgru_oracle_np as onp
from oracle import hgru_oracle_torch as otorch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _cuda(a):
    return torch.as_tensor(np.asarray(a, np.float32)).cuda()


def test_layer_methods_match_reference_layer_golden():
    """model.conv_layer / max_pool / fc_layer against the outputs of the reference's own methods
    (tests/golden/make_golden.py, gen_pose_layers), variables injected through data_dict like the reference."""
    z = np.load(os.path.join(GOLDEN, "pose_layers_ref.npz"))
    m = mp.model()
    m.data_dict = {"conv_1": [z["var:conv_1/conv_1_filters"], z["var:conv_1/conv_1_biases"]],
                   "conv_2": [z["var:conv_2/conv_2_filters"], z["var:conv_2/conv_2_biases"]],
                   "fc_1": [z["var:fc_1/fc_1_weights"], z["var:fc_1/fc_1_biases"]]}
    c1 = m.conv_layer(_cuda(z["x"]), 1, 6, "conv_1", filter_size=3)
    p1 = m.max_pool(c1, "pool_1")
    c2 = m.conv_layer(p1, 6, 6, "conv_2", filter_size=3)
    fc = m.fc_layer(c2, 6 * 6 * 6, 10, "fc_1")
    for got, key in ((c1, "conv1"), (p1, "pool1"), (c2, "conv2"), (fc, "fc1")):
        assert onp.rel_err(got.cpu().numpy(), z[key])[0] < 2e-6, key
    assert sorted("%s|%d" % k for k in m.var_dict) == sorted(z["var_dict_keys"].tolist())
    # hgru_layer through the model's own wrapper (hgru_pose.py:107-118: T = 8, SSN = SSF = 15) on conv2's output
    m.data_dict["contextual_circuit"] = {n: z["var:contextual_circuit/" + n] for n in onp.HGRU_PARAM_NAMES}
    m.hidden_state, m.compute_mode = z["hgru_O0"], "fp32"
    h = m.hgru_layer(c2)
    assert isinstance(h, tuple) and len(h) == 3                        # reference defect D4: a tuple
    assert onp.rel_err(h[0].cpu().numpy(), z["hgru_O"])[0] < 1e-4
    with pytest.raises(NotImplementedError):
        m.conv_layer(p1, 6, 6, "conv_2", stride=[1, 2, 2, 1])
    with pytest.raises(RuntimeError):
        m.max_pool(torch.zeros(1, 4, 4, 2), "p")                      # CPU tensor: no fallback


@pytest.mark.parametrize("shape", [(2, 9, 7, 5, 11, 3), (1, 16, 16, 64, 64, 3), (1, 6, 5, 3, 4, 15)])
def test_layer_methods_vs_oracle_on_odd_shapes(shape):
    n, h, w, ci, co, S = shape
    rng = np.random.default_rng(4)
    x = rng.uniform(-1, 1, size=(n, h, w, ci)).astype(np.float32)
    m = mp.model()
    m.seed = 7
    y = m.conv_layer(_cuda(x), ci, co, "any_name", filter_size=S)
    filt, bias = m.var_dict[("any_name", 0)].cpu().numpy(), m.var_dict[("any_name", 1)].cpu().numpy()
    assert onp.rel_err(y.cpu().numpy(), onp.conv_layer(x, filt, bias))[0] < 1e-5
    # SAME pooling of odd sizes: ceil(h/2) x ceil(w/2), clipped windows
    p = m.max_pool(_cuda(x), "pool").cpu().numpy()
    xp = np.full((n, h + h % 2, w + w % 2, ci), -np.inf, np.float32)
    xp[:, :h, :w] = x
    assert np.array_equal(p, onp.max_pool_2x2(xp).astype(np.float32))
    f = m.fc_layer(_cuda(x), h * w * ci, 13, "fc_any").cpu().numpy()
    wts, b = m.var_dict[("fc_any", 0)].cpu().numpy(), m.var_dict[("fc_any", 1)].cpu().numpy()
    assert onp.rel_err(f, onp.fc_layer(x, wts, b))[0] < 1e-5


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_train_mode_forward_vs_oracle(mode):
    """build(train_mode=True): batch statistics in all five batch norms, dropout (keep 0.7, documented counter-based
    mask) after relu(fc1), moving-statistics updates -- against the numpy oracle."""
    N, ch, hw, T, S, F = 6, 16, 16, 2, 15, 48
    P = init.pose_params(channels=ch, S=S, T=T, hw=hw, fc_hidden=F, out=69, seed=3, stress=4.0, random_bn=True)
    depth = init.synthetic_depth(N, seed=2, size=2 * hw)
    h0 = init.hidden_init((N, hw, hw, ch), seed=5)
    m = mp.model()
    m.channels, m.timesteps, m.fc_hidden, m.compute_mode, m.hidden_state = ch, T, F, mode, h0
    m.dropout_seed = 99
    m.load_params(P)
    out = m.build(torch.as_tensor(depth).cuda(), 69, train_mode=True)
    ref, acts = onp.pose_forward(depth, P, h0, timesteps=T, train_mode=True, trace=True, dropout_keep=0.7,
                                 dropout_seed=99)
    tol = 1e-4 if mode == "fp32" else 1e-2
    for name in ("conv1", "pool1", "conv2", "conv3"):
        assert onp.rel_err(getattr(m, name).cpu().numpy(), acts[name])[0] < 1e-4, name
    assert onp.rel_err(m.hgru.cpu().numpy(), acts["hgru_bn"])[0] < tol
    assert onp.rel_err(m.fc1.cpu().numpy(), acts["fc1"])[0] < tol
    assert onp.rel_err(m.relu1.cpu().numpy(), acts["relu1"])[0] < 2 * tol
    assert onp.rel_err(out.cpu().numpy(), ref)[0] < 2 * tol
    assert m.out_put is out and m.fc4 is out
    # dropped units exist and are exactly the oracle's
    keep = onp.dropout_keep_mask(N * F, 0.7, 99)
    assert 0.55 < keep.mean() < 0.85
    # moving statistics as the reference's UPDATE_OPS would leave them
    mm, mv = onp.batch_norm_moving_update(onp.max_pool_2x2(acts["conv1"]), P["batch_normalization/moving_mean"],
                                          P["batch_normalization/moving_variance"])
    upd = m.updated_moving_stats["batch_normalization"]
    assert onp.rel_err(upd["moving_mean"].cpu().numpy(), mm)[0] < 1e-5
    assert onp.rel_err(upd["moving_variance"].cpu().numpy(), mv)[0] < 1e-5
    assert set(m.updated_moving_stats) == set(init.BN_SCOPES)
    # inference mode on the same model object afterwards still runs the fused pipeline
    out_inf = m.build(torch.as_tensor(depth).cuda(), 69, train_mode=False)
    ref_inf = otorch.pose_forward(depth, P, h0, timesteps=T, dtype=torch.float64).numpy()
    assert onp.rel_err(out_inf.cpu().numpy(), ref_inf)[0] < tol


def test_attention_cnn_train_mode_forward_vs_oracle():
    """attn_model_struct.build(train_mode=True) -- what the reference passes in training (train_cnn_networks_hgru.py:117)
    and, as committed, in eval_model_on_real_data (:360): batch statistics in the six batch norms, dropout after
    relu(afc_1), composed from the class's layer methods (incl. the bit-exact TF-1.x bilinear resize)."""
    from oracle import attn_oracle_np as anp
    N, H, W, widths, F = 5, 96, 80, (8, 24, 16, 40, 32), 64
    P = init.attn_params(widths, F, 3, seed=5, random_bn=True)
    rng = np.random.default_rng(3)
    frames = rng.uniform(0.06, 0.4, size=(N, H, W, 1)).astype(np.float32)
    frames[rng.uniform(size=frames.shape) < 0.05] = 0.0
    m = mp.attn_model_struct()
    m.dropout_seed = 77
    m.load_params(P)
    out = m.build(torch.as_tensor(frames).cuda(), 3, train_mode=True)
    ref, acts = anp.attn_forward(frames, P, trace=True, train_mode=True, dropout_keep=0.7, dropout_seed=77)
    assert np.array_equal(m.resize_images(torch.as_tensor(frames).cuda(), [128, 128]).cpu().numpy(), acts["resized"])
    for k in ("pool1", "pool3", "pool5"):
        assert onp.rel_err(getattr(m, k).cpu().numpy(), acts[k])[0] < 1e-4, k
    assert onp.rel_err(m.fc1.cpu().numpy(), acts["fc1"])[0] < 1e-4
    assert onp.rel_err(m.relu1.cpu().numpy(), acts["relu1"])[0] < 1e-4
    assert onp.rel_err(out.cpu().numpy(), ref)[0] < 1e-4
    assert set(m.updated_moving_stats) == set(init.ATTN_BN_SCOPES)
    # inference mode afterwards: the fused pipeline, moving statistics
    out_inf = m.build(torch.as_tensor(frames).cuda(), 3)
    ref_inf = anp.attn_forward(frames, P)
    assert onp.rel_err(out_inf.cpu().numpy(), ref_inf)[0] < 1e-4
    assert onp.rel_err(m.pool5.cpu().numpy(), anp.attn_forward(frames, P, trace=True)[1]["pool5"])[0] < 1e-4


def test_attribute_surface_after_inference_build():
    """conv1, pool1, conv2, conv3, hgru, fc1, relu1, fc4, out_put (hgru_pose.py:50-105) -- batch-normalised where
    the reference re-assigns them -- materialised on first access after the fused forward."""
    N, ch, hw, T, S, F = 2, 25, 16, 2, 15, 32
    P = init.pose_params(channels=ch, S=S, T=T, hw=hw, fc_hidden=F, out=69, seed=3, stress=4.0, random_bn=True)
    depth = init.synthetic_depth(N, seed=2, size=2 * hw)
    h0 = init.hidden_init((N, hw, hw, ch), seed=5)
    m = mp.model()
    m.channels, m.timesteps, m.fc_hidden, m.compute_mode, m.hidden_state = ch, T, F, "bf16", h0
    m.load_params(P)
    out = m.build(torch.as_tensor(depth).cuda(), 69)
    ref, acts = onp.pose_forward(depth, P, h0, timesteps=T, trace=True)
    assert m.conv1.shape == (N, 2 * hw, 2 * hw, ch)
    assert onp.rel_err(m.conv1.cpu().numpy(), acts["conv1"])[0] < 1e-5
    for name, key in (("pool1", "pool1"), ("conv2", "conv2"), ("conv3", "conv3")):
        assert onp.rel_err(getattr(m, name).cpu().numpy(), acts[key])[0] < 1e-4, name
    assert onp.rel_err(m.hgru.cpu().numpy(), acts["hgru_bn"])[0] < 1e-2           # post batch-norm, as :82 re-assigns
    assert onp.rel_err(m.activation("hgru").cpu().numpy(), acts["hgru"])[0] < 1e-2   # the circuit's own output
    assert onp.rel_err(m.fc1.cpu().numpy(), acts["fc1"])[0] < 1e-2
    assert onp.rel_err(m.relu1.cpu().numpy(), acts["relu1"])[0] < 1e-2
    assert m.fc4 is out and m.out_put is out and m["conv3"] is m.conv3 and "relu1" in m
    # a second build invalidates the cached tensors
    first = m.conv3
    m.build(torch.as_tensor(init.synthetic_depth(N, seed=3, size=2 * hw)).cuda(), 69)
    assert not torch.equal(first, m.conv3)
    # hgru_layer: the circuit's build() as the reference returns it (a tuple under return_weights=True, defect D4)
    h = m.hgru_layer(m.conv3)
    assert isinstance(h, tuple) and len(h) == 3 and h[0].shape == (N, hw, hw, ch)


def test_model_default_hidden_state_is_cached_and_named_initial_states():
    m = mp.model()
    m.channels, m.timesteps, m.fc_hidden = 16, 1, 16
    d = torch.as_tensor(init.synthetic_depth(2, seed=1, size=32)).cuda()
    a = m.build(d, 5).clone()
    h0_first = m._h0
    b = m.build(d, 5)
    assert m._h0 is h0_first and torch.equal(a, b)          # the seeded O_0 is drawn once per shape
    m2 = mp.model()
    m2.channels, m2.timesteps, m2.fc_hidden = 16, 1, 16
    m2.aux = dict(m2.aux, hidden_init="zeros")
    assert torch.isfinite(m2.build(d, 5)).all() and m2._h0 is None
    m2.aux["hidden_init"] = "bogus"
    with pytest.raises(RuntimeError):
        m2.build(d, 5)


@pytest.mark.parametrize("mode", ["fp32", "bf16", "bf16x3"])
@pytest.mark.parametrize("hidden_init", ["identity", "zeros"])
def test_model_named_initial_state_vs_oracle(mode, hidden_init):
    """aux['hidden_init'] = 'identity' (O_0 = X = conv3 of the same forward, hgru_module.py:876-878) and 'zeros'
    (:888-890) through the fused forward, device and host entry points, against the fp64 oracle."""
    N, ch, hw, T, S, F = 3, 25, 16, 3, 15, 32
    P = init.pose_params(channels=ch, S=S, T=T, hw=hw, fc_hidden=F, out=69, seed=3, stress=4.0, random_bn=True)
    depth = init.synthetic_depth(N, seed=2, size=2 * hw)
    m = mp.model()
    m.channels, m.timesteps, m.SSF, m.SSN, m.fc_hidden, m.compute_mode = ch, T, S, S, F, mode
    m.aux = dict(m.aux, hidden_init=hidden_init)
    m.load_params(P)
    out = m.build(torch.as_tensor(depth).cuda(), 69)
    ref, acts = otorch.pose_forward(depth, P, hidden_init, timesteps=T, dtype=torch.float64, trace=True)
    tol = {"fp32": 1e-4, "bf16x3": 1e-4, "bf16": 1e-2}[mode]
    e_h = onp.rel_err(m.activation("hgru").cpu().numpy(), acts["hgru"].numpy())[0]
    e_o = onp.rel_err(out.cpu().numpy(), ref.numpy())[0]
    print("hidden_init=%s %s: hgru rel %.3e, out_put rel %.3e" % (hidden_init, mode, e_h, e_o))
    assert e_h < tol and e_o < tol
    # the named state must matter (a silently ignored option would also pass a loose tolerance against itself)
    other = otorch.pose_forward(depth, P, "zeros" if hidden_init == "identity" else "identity", timesteps=T,
                                dtype=torch.float64)
    assert onp.rel_err(other.numpy(), ref.numpy())[0] > 5 * tol
    out_host = m.build(torch.as_tensor(depth).pin_memory(), 69)
    assert torch.equal(out_host, out.cpu())
    # an explicit hidden_state wins over the option, and switching back restores the named state
    m.hidden_state = init.hidden_init((N, hw, hw, ch), seed=5)
    o2 = m.build(torch.as_tensor(depth).cuda(), 69)
    assert not torch.equal(o2, out)
    m.hidden_state = None
    assert torch.equal(m.build(torch.as_tensor(depth).cuda(), 69), out)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_circuit_without_adaptation_needs_no_rho(mode):
    """adapation=False (the default of the circuit's own aux, hgru_module.py:44): no `rho` variable exists
    (:490-493), the state is not rescaled (:847-849), `weights` has no rho_r."""
    n, h, w, k, S, T = 1, 16, 16, 16, 5, 3
    rng = np.random.default_rng(5)
    X = rng.uniform(-1, 1, size=(n, h, w, k)).astype(np.float32)
    O0 = init.hidden_init((n, h, w, k), seed=3, limit=0.5)
    params = init.hgru_params(k, S, T, seed=9, stress=3.0)
    aux = dict(mp.model().aux, adapation=False)
    no_rho = {key: v for key, v in params.items() if key != "rho"}
    cc = mp.ContextualCircuit(X=_cuda(X), timesteps=T, SSN=S, SSF=S, aux=aux, params=no_rho, hidden_state=O0,
                              compute_mode=mode)
    O, weights, _ = cc.build()
    ref = otorch.hgru_forward(X, O0, dict(params, rho=np.ones(T, np.float32)), T, dtype=torch.float64)
    assert onp.rel_err(O.cpu().numpy(), ref.numpy())[0] < (1e-4 if mode == "fp32" else 1e-2)
    assert "rho_r" not in weights and not hasattr(cc, "rho")
    # with adaptation a non-unit rho is applied
    params2 = dict(params, rho=np.linspace(0.5, 1.5, T).astype(np.float32))
    O2, w2, _ = mp.ContextualCircuit(X=_cuda(X), timesteps=T, SSN=S, SSF=S, aux=dict(aux, adapation=True),
                                     params=params2, hidden_state=O0, compute_mode=mode).build()
    ref2 = otorch.hgru_forward(X, O0, params2, T, dtype=torch.float64)
    assert onp.rel_err(O2.cpu().numpy(), ref2.numpy())[0] < (1e-4 if mode == "fp32" else 1e-2)
    assert "rho_r" in w2
    with pytest.raises(KeyError):
        mp.ContextualCircuit(X=_cuda(X), timesteps=T, SSN=S, SSF=S, aux=dict(aux, adapation=True), params=no_rho,
                             hidden_state=O0, compute_mode=mode).build()


def test_plan_cache_is_bounded_and_can_be_cleared():
    from monkey_pose_b200 import hgru_module as hm
    hm.clear_plan_cache()
    free0 = torch.cuda.mem_get_info()[0]
    for i in range(hm.PLAN_CACHE_SIZE + 3):
        X = torch.rand(8, 32 + 2 * i, 32, 16, device="cuda")
        mp.ContextualCircuit(X=X, timesteps=1, SSN=5, SSF=5, aux=dict(mp.model().aux, hidden_init="zeros"),
                             compute_mode="bf16").build()
    assert len(hm._PLAN_CACHE) == hm.PLAN_CACHE_SIZE
    hm.clear_plan_cache()
    torch.cuda.synchronize()
    assert len(hm._PLAN_CACHE) == 0
    assert torch.cuda.mem_get_info()[0] >= free0 - (64 << 20)       # workspaces are back (torch's own pool aside)


def test_checkpoint_import_feeds_the_device_path_bitwise(tmp_path):
    """model.load_checkpoint(bundle) -> build equals model.load_params(flat) -> build bit for bit, for a bundle laid
    out like the reference's combined training graph: everything under `cnn/`, the pose model's batch norms at
    batch_normalization_6 .. _10 behind the attention CNN's six, Adam slots beside the variables
    (train_cnn_networks_hgru.py:96, 115-117, 188).  (The file format itself stays "parity unpinned": no file written
    by TensorFlow is available here, tests/test_tf_checkpoint.py.)"""
    from monkey_pose_b200 import tf_checkpoint as ck
    ch, hw, T, F = 25, 16, 2, 32
    P = init.pose_params(channels=ch, S=15, T=T, hw=hw, fc_hidden=F, out=69, seed=4, stress=4.0, random_bn=True)
    A = init.attn_params(widths=(8, 8, 8, 8, 8), fc_hidden=16, out=3, seed=2, random_bn=True)
    flat = {"cnn/" + k: np.asarray(v) for k, v in A.items()}
    for k, v in P.items():
        if k.startswith("batch_normalization"):
            scope, field = k.split("/")
            i = 0 if scope == "batch_normalization" else int(scope.rsplit("_", 1)[1])
            k = "batch_normalization_%d/%s" % (i + 6, field)
        flat["cnn/" + k] = np.asarray(v)
        if "filters" in k or "weights" in k:
            flat["cnn/" + k + "/Adam"] = np.zeros_like(v)
    flat["beta1_power"] = np.array(0.9, np.float32)
    prefix = str(tmp_path / "attn_cnn_model5000.ckpt")
    ck.write_checkpoint(prefix, flat)
    depth = torch.as_tensor(init.synthetic_depth(3, seed=2, size=2 * hw)).cuda()
    h0 = init.hidden_init((3, hw, hw, ch), seed=5)
    outs = []
    for how in ("checkpoint", "params"):
        m = mp.model()
        m.channels, m.timesteps, m.fc_hidden, m.hidden_state = ch, T, F, h0
        if how == "checkpoint":
            m.load_checkpoint(prefix)
        else:
            m.load_params(P)
        outs.append((m.build(depth, 69).clone(), m.activation("hgru")))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    ref = otorch.pose_forward(depth.cpu().numpy(), P, h0, timesteps=T, dtype=torch.float64).numpy()
    assert onp.rel_err(outs[0][0].cpu().numpy(), ref)[0] < 1e-2


@pytest.mark.parametrize("shape", [(2, 8, 8, 5), (3, 9, 13, 4), (1, 6, 7, 3), (2, 1, 2, 2)])
def test_declared_but_unused_helpers_of_the_model(shape):
    """max_pool_4, avg_pool and batchnorm (hgru_pose.py:120-132): the reference declares them and its build() never
    calls them; they exist as stand-alone ops with TensorFlow's SAME-pooling semantics (padding split before / after,
    left out of an average's count) and moments over the batch axis."""
    rng = np.random.default_rng(shape[1] * 31 + shape[2])
    x = rng.standard_normal(shape).astype(np.float32)
    m = mp.model()
    X = torch.as_tensor(x).cuda()
    got = m.max_pool_4(X, "p").cpu().numpy()
    assert got.shape == onp.pool_same(x, 4).shape and np.array_equal(got, onp.pool_same(x, 4).astype(np.float32))
    assert onp.rel_err(m.avg_pool(X, "a").cpu().numpy(), onp.pool_same(x, 2, average=True))[0] < 1e-6
    assert np.array_equal(m.max_pool(X, "q").cpu().numpy(), onp.pool_same(x, 2).astype(np.float32))   # the 2x2 kernel agrees
    if shape[0] > 1:
        assert onp.rel_err(m.batchnorm(X).cpu().numpy(), onp.batchnorm_moments0(x))[0] < 1e-5
    with pytest.raises(RuntimeError):
        m.avg_pool(torch.as_tensor(x), "a")
