"""The attention-CNN oracle (oracle/attn_oracle_np.py) against the golden vectors produced by executing the
reference's own `attn_model_struct` source (tests/golden/make_golden_attn.py), and the torch twin against it."""
import os

import numpy as np
import torch

from oracle import attn_oracle_np as anp
from oracle import attn_oracle_torch as atorch
from oracle import hgru_oracle_np as onp

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "attn_ref.npz")


def _load():
    z = np.load(GOLDEN)
    P = {k[4:]: z[k] for k in z.files if k.startswith("var:")}
    A = {k[4:]: z[k] for k in z.files if k.startswith("act:")}
    return z["frames"], P, A


def test_attn_oracle_matches_reference_golden_every_layer():
    frames, P, A = _load()
    out, acts = anp.attn_forward(frames, P, trace=True)
    assert np.array_equal(acts["resized"], A["resized"])                    # float32 bilinear: bit-exact
    for k in ("pool1", "pool2", "pool3", "pool4", "pool5"):
        assert onp.rel_err(acts[k], A[k])[0] < 2e-7, k                      # stored as float32
    for k in ("fc1", "relu1", "out_put"):
        assert onp.rel_err(acts[k], A[k])[0] < 1e-12, k
    assert out.shape == (frames.shape[0], 3)


def test_attn_torch_twin_matches_numpy_arbiter():
    frames, P, _ = _load()
    ref, racts = anp.attn_forward(frames, P, trace=True)
    out, acts = atorch.attn_forward(frames, P, dtype=torch.float64, trace=True)
    for k in ("pool1", "pool3", "pool5", "relu1", "out_put"):
        assert onp.rel_err(acts[k].numpy(), racts[k])[0] < 1e-10, k


def test_resize_bilinear_identity_and_known_values():
    x = np.arange(2 * 4 * 6, dtype=np.float32).reshape(2, 4, 6, 1)
    assert np.array_equal(anp.resize_bilinear_tf1(x, 4, 6), x)
    y = anp.resize_bilinear_tf1(x, 8, 12)                                    # 2x up: src = dst / 2
    assert y[0, 0, 0, 0] == x[0, 0, 0, 0] and y[0, 0, 1, 0] == 0.5 * (x[0, 0, 0, 0] + x[0, 0, 1, 0])
    assert y[0, 7, 11, 0] == x[0, 3, 5, 0]                                   # clamped upper neighbour
    z = anp.resize_bilinear_tf1(x, 2, 3)                                     # 2x down: picks every other pixel
    assert np.array_equal(z[..., 0], x[:, ::2, ::2, 0])
