"""GPU (B200): CUDA path against the fp64 oracle AT THE BASELINE CONFIGS (BASELINE.json configs[1]-[3]):
25 and 64 channels at T = 8, the deep variant 32 channels at T = 16, 64x64 maps, 15x15 kernels -- every timestep's
H1 and H2, random-init and "stress" weights (15x15 kernel x4-5 and a large initial state, so tanh and the gates leave
their linear region and conv error cannot hide), all three arithmetic modes; the whole pose forward at
batch 256 against the oracle on a frame subset; the host entry point at batch 256; batch shards against the
unsharded forward.

Tolerances (BASELINE.json north_star): fp32-class paths ('fp32', 'bf16x3') <= 1e-4 relative on hidden states and
outputs; bf16 tensor-core path <= 1e-2 relative and <= 0.5 mm mean per-joint deviation.  Relative = max|a-b| / max|b|
(SURVEY.md 8d parity metrics), per tensor and per timestep.  Reference loop body: hgru_module.py:825-857."""
import functools
import os
import socket

import numpy as np
import pytest
import torch

import monkey_pose_b200 as mp
from monkey_pose_b200 import initialization as init
from oracle import hgru_oracle_np as onp
from oracle import hgru_oracle_torch as otorch

pytestmark = pytest.mark.gpu

POSE_AUX = mp.model().aux
TOL = {"fp32": 1e-4, "bf16": 1e-2, "bf16x3": 1e-4}
# (channels, timesteps): BASELINE configs[1] (25 ch), the reference's own width (64 ch), the deep variant (configs[3])
LAYER_CASES = {"k25_T8": (25, 8), "k64_T8": (64, 8), "k32_T16": (32, 16)}
# (15x15 kernel scale, initial-state limit)
WEIGHT_SETS = {"random_init": (1.0, 0.005), "stress": (5.0, 0.5)}
N_FRAMES = 2


@functools.lru_cache(maxsize=None)
def _layer_case(case, wset):
    """Inputs + the fp64 oracle's per-timestep states (computed once per case, shared by the three modes)."""
    k, T = LAYER_CASES[case]
    stress, h0_lim = WEIGHT_SETS[wset]
    if T == 16 and stress > 1.0:
        stress = 4.0
    rng = np.random.default_rng(17 + k)
    X = rng.uniform(-1, 1, size=(N_FRAMES, 64, 64, k)).astype(np.float32)
    O0 = init.hidden_init((N_FRAMES, 64, 64, k), seed=3, limit=h0_lim)
    params = init.hgru_params(k, 15, T, seed=9, stress=stress)
    ref, H1s, H2s = otorch.hgru_forward(X, O0, params, T, dtype=torch.float64, trace=True)
    return X, O0, params, ref.numpy(), [h.numpy() for h in H1s], [h.numpy() for h in H2s]


@pytest.mark.parametrize("mode", ["bf16", "bf16x3", "fp32"])
@pytest.mark.parametrize("wset", sorted(WEIGHT_SETS))
@pytest.mark.parametrize("case", sorted(LAYER_CASES))
def test_hgru_every_timestep_vs_fp64_oracle_at_baseline_configs(case, wset, mode):
    k, T = LAYER_CASES[case]
    X, O0, params, ref, H1s, H2s = _layer_case(case, wset)
    cc = mp.ContextualCircuit(X=torch.as_tensor(X).cuda(), timesteps=T, SRF=1, SSN=15, SSF=15, aux=POSE_AUX,
                              params=params, hidden_state=O0, compute_mode=mode)
    O, _, _ = cc.build(trace=True)
    torch.cuda.synchronize()
    worst = 0.0
    for t in range(T):
        e1 = onp.rel_err(cc.I_steps[t].cpu().numpy(), H1s[t])[0]
        e2 = onp.rel_err(cc.O_steps[t].cpu().numpy(), H2s[t])[0]
        worst = max(worst, e1, e2)
        assert e1 < TOL[mode] and e2 < TOL[mode], (case, wset, mode, t, e1, e2)
    assert onp.rel_err(O.cpu().numpy(), ref)[0] < TOL[mode]
    if wset == "stress":      # the stress set really is off the linear region
        assert np.abs(H1s[-1]).max() > 0.5 and np.abs(ref).max() > 0.2
    print("%s %s %s: worst per-timestep rel err %.3e (budget %.0e)" % (case, wset, mode, worst, TOL[mode]))
    # the untraced forward (chained launches where the kernel supports them) gives the same final state, bit for bit
    O2, _, _ = mp.ContextualCircuit(X=torch.as_tensor(X).cuda(), timesteps=T, SRF=1, SSN=15, SSF=15, aux=POSE_AUX,
                                    params=params, hidden_state=O0, compute_mode=mode).build()
    assert torch.equal(O2, O)


GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("mode", ["bf16", "bf16x3", "fp32"])
@pytest.mark.parametrize("tag", ["S15_k25_T8", "S15_k25_T8_stress", "S15_k32_T16_stress", "S15_k64_T3_stress"])
def test_baseline_width_goldens_from_the_reference_source(tag, mode):
    """Numbers computed by the reference's own hgru_module.py statements (tests/golden/make_golden.py) at 25 channels /
    T = 8 -- the remainder-packed tap-stacked kernel --, 32 channels / T = 16 and 64 channels (the reference's own width:
    the plain tcgen05 conv kernel with epilogue-issued gates), every timestep."""
    z = np.load(os.path.join(GOLDEN, "hgru_ref_%s.npz" % tag))
    T, S = int(z["T"]), int(z["S"])
    params = {n: z["var:contextual_circuit/" + n] for n in onp.HGRU_PARAM_NAMES}
    cc = mp.ContextualCircuit(X=torch.as_tensor(z["X"]).cuda(), timesteps=T, SRF=1, SSN=S, SSF=S, aux=POSE_AUX,
                              params=params, hidden_state=z["O0"], compute_mode=mode)
    O, weights, _ = cc.build(trace=True)
    torch.cuda.synchronize()
    for t in range(T):
        e1 = onp.rel_err(cc.I_steps[t].cpu().numpy(), z["I_steps"][:, t])[0]
        e2 = onp.rel_err(cc.O_steps[t].cpu().numpy(), z["O_steps"][:, t])[0]
        assert e1 < TOL[mode] and e2 < TOL[mode], (tag, mode, t, e1, e2)
    assert onp.rel_err(O.cpu().numpy(), z["O_final"])[0] < TOL[mode]
    assert set(weights.keys()) == set(z["weights_keys"].tolist())


# ---- the whole pose forward at batch 256 (BASELINE configs[1]) ------------------------------------------------------
SUBSET = (0, 37, 100, 255)


@functools.lru_cache(maxsize=None)
def _pose_case(stress, random_bn, channels=25, T=8):
    N, hw, F = 256, 64, 1024
    P = init.pose_params(channels=channels, S=15, T=T, hw=hw, fc_hidden=F, out=69, seed=3, stress=stress,
                         random_bn=random_bn)
    depth = init.synthetic_depth(N, seed=1234, size=128)
    h0 = init.hidden_init((N, hw, hw, channels), seed=5)
    idx = list(SUBSET)
    ref, acts = otorch.pose_forward(depth[idx], P, h0[idx], timesteps=T, dtype=torch.float64, trace=True)
    return P, depth, h0, ref.numpy(), {k: v.numpy() for k, v in acts.items()}


def _model(P, h0, channels, T, mode):
    m = mp.model()
    m.channels, m.timesteps, m.compute_mode, m.hidden_state = channels, T, mode, h0
    m.load_params(P)
    return m


@pytest.mark.parametrize("stress,random_bn", [(1.0, False), (4.0, True)], ids=["random_init", "stress"])
def test_pose_forward_batch256_vs_fp64_oracle(stress, random_bn):
    """N = 256, 25 channels, T = 8, 15x15: bf16 forward of the full batch, checked on a frame subset against the fp64
    oracle (frames are independent); the fp32 and bf16x3 modes on the same frames against the fp32 budget; the host
    entry point (crops uploaded in four chunks, the stem following chunk by chunk) against the device entry point."""
    P, depth, h0, ref, acts = _pose_case(stress, random_bn)
    idx = list(SUBSET)
    m = _model(P, h0, 25, 8, "bf16")
    out = m.build(torch.as_tensor(depth).cuda(), 69)
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    assert np.isfinite(got).all()
    e = onp.rel_err(got[idx], ref)[0]
    mm = onp.mean_joint_error_mm(got[idx], ref)
    eh = onp.rel_err(m.activation("hgru").cpu().numpy()[idx], acts["hgru"])[0]
    ex = onp.rel_err(m.activation("conv3").cpu().numpy()[idx], acts["conv3"])[0]
    print("pose N=256 bf16 (stress %g): out_put rel %.3e, %.4f mm, hgru rel %.3e, conv3 rel %.3e" % (stress, e, mm, eh, ex))
    assert e < 1e-2 and mm < 0.5 and eh < 1e-2 and ex < 1e-2
    # host buffers in, host buffers out, at the size the end-to-end number is quoted on
    out_host = m.build(torch.as_tensor(depth).pin_memory(), 69)
    assert not out_host.is_cuda and np.array_equal(out_host.numpy(), got)
    # the same four frames as their own batch: bitwise the same predictions (the sharding premise)
    sub = _model(P, h0[idx], 25, 8, "bf16").build(torch.as_tensor(depth[idx]).cuda(), 69)
    assert np.array_equal(sub.cpu().numpy(), got[idx])
    for mode in ("fp32", "bf16x3"):
        o = _model(P, h0[idx], 25, 8, mode).build(torch.as_tensor(depth[idx]).cuda(), 69).cpu().numpy()
        e = onp.rel_err(o, ref)[0]
        print("pose %s (stress %g): out_put rel %.3e, %.5f mm" % (mode, stress, e, onp.mean_joint_error_mm(o, ref)))
        assert e < 1e-4 and onp.mean_joint_error_mm(o, ref) < 0.05


def test_pose_forward_reference_width_and_deep_variant_vs_fp64_oracle():
    """64 channels (the reference's own width) at T = 8 and the deep variant (32 channels, T = 16), whole model,
    stress weights, 8 frames -- more units than a launch chain needs to be exercised, few enough for the oracle."""
    for ch, T in ((64, 8), (32, 16)):
        N, hw, F = 8, 64, 256
        P = init.pose_params(channels=ch, S=15, T=T, hw=hw, fc_hidden=F, out=69, seed=3, stress=4.0, random_bn=True)
        depth = init.synthetic_depth(N, seed=77, size=128)
        h0 = init.hidden_init((N, hw, hw, ch), seed=5)
        idx = [0, 7]
        ref = otorch.pose_forward(depth[idx], P, h0[idx], timesteps=T, dtype=torch.float64).numpy()
        m = _model(P, h0, ch, T, "bf16")
        m.fc_hidden = F
        got = m.build(torch.as_tensor(depth).cuda(), 69).cpu().numpy()
        e, mm = onp.rel_err(got[idx], ref)[0], onp.mean_joint_error_mm(got[idx], ref)
        print("pose k=%d T=%d bf16: out_put rel %.3e, %.4f mm" % (ch, T, e, mm))
        assert e < 1e-2 and mm < 0.5


# ---- batch shards against the unsharded forward ---------------------------------------------------------------------
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _shard_worker(rank, world, port, n_total, channels, T, out_path):
    """One process per shard (one per GPU when the box has several; NCCL then, else both on cuda:0 over gloo)."""
    import torch.distributed as dist
    from monkey_pose_b200.sharding import PredictionGatherer, gather_predictions, shard_bounds
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    multi = torch.cuda.device_count() >= world
    torch.cuda.set_device(rank if multi else 0)
    dist.init_process_group("nccl" if multi else "gloo", rank=rank, world_size=world)
    P = init.pose_params(channels=channels, S=15, T=T, hw=64, fc_hidden=128, out=69, seed=3, stress=4.0,
                         random_bn=True)
    depth = init.synthetic_depth(n_total, seed=1234, size=128)
    h0 = init.hidden_init((n_total, 64, 64, channels), seed=5)
    lo, hi = shard_bounds(n_total, rank, world)
    m = mp.model()
    m.channels, m.timesteps, m.fc_hidden, m.compute_mode, m.hidden_state = channels, T, 128, "bf16", h0[lo:hi]
    m.load_params(P)
    out = m.build(torch.as_tensor(depth[lo:hi]).cuda(), 69)
    torch.cuda.synchronize()
    full = gather_predictions(out if multi else out.cpu(), n_total)
    if n_total % world == 0:
        # equal shards: the asynchronous double-buffered gather over three steps lands the same rows
        g = PredictionGatherer(hi - lo, 69, device="cuda" if multi else "cpu")
        for _ in range(3):
            o = m.build(torch.as_tensor(depth[lo:hi]).cuda(), 69)
            g.submit(o if multi else o.cpu())
        g.drain()
        assert torch.equal(g.result().cpu(), full.cpu())
    if rank == 0:
        np.save(out_path, full.cpu().numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total,channels,T", [(23, 25, 3), (12, 64, 2)])
def test_sharded_gather_equals_unsharded_forward_bitwise(tmp_path, n_total, channels, T):
    """World size 2, contiguous (ragged) batch shards, replicated seeded parameters, predictions all-gathered in batch
    order: bit for bit the single-process forward of the whole batch."""
    import torch.multiprocessing as tmp_
    out_path = str(tmp_path / "gathered.npy")
    tmp_.spawn(_shard_worker, args=(2, _free_port(), n_total, channels, T, out_path), nprocs=2, join=True)
    gathered = np.load(out_path)
    P = init.pose_params(channels=channels, S=15, T=T, hw=64, fc_hidden=128, out=69, seed=3, stress=4.0,
                         random_bn=True)
    depth = init.synthetic_depth(n_total, seed=1234, size=128)
    h0 = init.hidden_init((n_total, 64, 64, channels), seed=5)
    m = mp.model()
    m.channels, m.timesteps, m.fc_hidden, m.compute_mode, m.hidden_state = channels, T, 128, "bf16", h0
    m.load_params(P)
    whole = m.build(torch.as_tensor(depth).cuda(), 69).cpu().numpy()
    assert gathered.shape == (n_total, 69)
    assert np.array_equal(gathered, whole)
