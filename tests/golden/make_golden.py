"""Generate the golden vectors under tests/golden/ by EXECUTING THE REFERENCE'S OWN SOURCE.

Run in the build container only (needs /root/reference; the GPU box never runs this):

    python tests/golden/make_golden.py

How: `hgru_module.py` and `hgru_pose.py` are read from /root/reference as text, three mechanical
py2->py3 token substitutions are applied in memory (print statement, dict.iteritems, basestring --
reference defect D2), and the result is exec'd with `tensorflow`, `utils.py_utils` and
`ops.initialization` resolved to tests/golden/tf_shim.py (numpy, float64; defect D1).  Nothing
from the reference is written into this repository except the numbers it computes.

Outputs (small .npz files, committed):
  hgru_ref_<tag>.npz  ContextualCircuit(...).build() with the hgru_pose aux dict, its loop
                      body wrapped to record every timestep: X, the variables the reference created (by their TF
                      names), the I_0/O_0 it drew, final O, per-timestep I (H1) and O (H2).
  pose_layers_ref.npz model.conv_layer / max_pool / fc_layer / hgru_layer in->out pairs.
"""
import os
import re
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path.insert(0, HERE)
import tf_shim  # noqa: E402


def _load_reference_module(name):
    src = open(os.path.join(REF, name + ".py")).read()
    src = re.sub(r"^(\s*)print '([^']*)'\s*$", r"\1print('\2')", src, flags=re.M)
    src = src.replace(".iteritems()", ".items()").replace("basestring", "str")
    mod = types.ModuleType(name)
    mod.__file__ = os.path.join(REF, name + ".py")
    exec(compile(src, mod.__file__, "exec"), mod.__dict__)
    return mod


def _install_shims():
    tfm = types.ModuleType("tensorflow")
    for k in dir(tf_shim):
        if not k.startswith("__"):
            setattr(tfm, k, getattr(tf_shim, k))
    sys.modules["tensorflow"] = tfm
    um = types.ModuleType("utils")
    um.py_utils = tf_shim.py_utils
    sys.modules["utils"] = um
    om = types.ModuleType("ops")
    om.initialization = tf_shim.initialization
    sys.modules["ops"] = om


def pose_aux(hgru_pose_mod):
    return dict(hgru_pose_mod.model().aux)


def gen_hgru(hm, aux, tag, n, h, w, k, S, T, seed, gains=None, f32_steps=False):
    reg = tf_shim.reset(seed)
    reg.init_gain = dict(gains or {})
    X = tf_shim.Tensor(reg.rng.uniform(-1.0, 1.0, size=(n, h, w, k)).astype(np.float32))
    cc = hm.ContextualCircuit(X=X, timesteps=T, SRF=1, SSN=S, SSF=S, strides=[1, 1, 1, 1],
                              padding="SAME", aux=dict(aux))
    # Per-timestep states are captured by wrapping the reference's loop body.  (Its own
    # `store_states` option is unusable: `full(i0, O, I, store_O, store_I)` receives
    # `[i0, O, I, store_I, store_O]` positionally, hgru_module.py:825,896-902, so the two
    # TensorArrays swap roles every iteration -- reference defect D10, off on the configured path.)
    steps_I, steps_O = [], []
    ref_full = cc.full

    def recording_full(i0, O, I, a, b):
        r = ref_full(i0, O, I, a, b)
        steps_O.append(np.array(r[1], copy=True))
        steps_I.append(np.array(r[2], copy=True))
        return r

    cc.full = recording_full
    O_final, weights, acts = cc.build()
    assert len(steps_O) == T
    # the two activation-shaped draws are I_0 then O_0 (hgru_module.py:879-887)
    act_draws = [arr for shp, arr in reg.drawn if tuple(shp) == (n, h, w, k)]
    assert len(act_draws) == 2
    out = {"X": np.asarray(X).astype(np.float32), "I0": act_draws[0], "O0": act_draws[1],
           "O_final": np.asarray(O_final), "O_steps": np.stack(steps_O, 1),
           "I_steps": np.stack(steps_I, 1),
           "init_gain": np.array(sorted("%s:%g" % (shp, g) for shp, g in reg.init_gain.items())),
           "T": np.int64(T), "S": np.int64(S), "weights_keys": np.array(sorted(weights.keys()))}
    for name, val in reg.variables.items():
        out["var:" + name] = val
    assert np.array_equal(out["O_steps"][:, -1], out["O_final"])
    assert reg.conv_calls == 4 * T, reg.conv_calls
    if f32_steps:      # BASELINE-width sets: float32 storage (6e-8 relative, far below the 1e-4 gate) keeps them small
        for key in ("O_final", "O_steps", "I_steps"):
            out[key] = out[key].astype(np.float32)
    path = os.path.join(HERE, "hgru_ref_%s.npz" % tag)
    np.savez_compressed(path, **out)
    print("wrote", path, "O_final absmax", np.abs(out["O_final"]).max())


def gen_pose_layers(pm, seed):
    reg = tf_shim.reset(seed)
    m = pm.model()
    out = {}
    x = tf_shim.Tensor(reg.rng.uniform(0.0, 1.0, size=(2, 12, 12, 1)).astype(np.float32))
    c1 = m.conv_layer(x, 1, 6, "conv_1", filter_size=3)
    p1 = m.max_pool(c1, "pool_1")
    c2 = m.conv_layer(p1, 6, 6, "conv_2", filter_size=3)
    fc = m.fc_layer(c2, 6 * 6 * 6, 10, "fc_1")
    out.update(x=np.asarray(x).astype(np.float32), conv1=np.asarray(c1), pool1=np.asarray(p1),
               conv2=np.asarray(c2), fc1=np.asarray(fc))
    for name, val in reg.variables.items():
        out["var:" + name] = val
    # hgru_layer through the model's own wrapper (hgru_pose.py:107-118): T=8, SSN=SSF=15
    h = m.hgru_layer(c2)
    assert isinstance(h, tuple) and len(h) == 3      # reference defect D4: a tuple, not a tensor
    out["hgru_O"] = np.asarray(h[0])
    act = [arr for shp, arr in reg.drawn if tuple(shp) == tuple(np.asarray(c2).shape)]
    out["hgru_I0"], out["hgru_O0"] = act[0], act[1]
    for name, val in reg.variables.items():
        out["var:" + name] = val
    out["var_dict_keys"] = np.array(["%s|%d" % k for k in sorted(m.var_dict.keys())])
    path = os.path.join(HERE, "pose_layers_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


def main():
    _install_shims()
    hm = _load_reference_module("hgru_module")
    sys.modules["hgru_module"] = hm
    pm = _load_reference_module("hgru_pose")
    aux = pose_aux(pm)
    gen_hgru(hm, aux, "S15_k8", n=2, h=16, w=16, k=8, S=15, T=3, seed=11)
    gen_hgru(hm, aux, "S5_k8", n=2, h=16, w=16, k=8, S=5, T=3, seed=12)
    gen_hgru(hm, aux, "S7_k5", n=1, h=12, w=10, k=5, S=7, T=4, seed=13)
    # BASELINE widths: 25 channels / T = 8 (the remainder-packed tap-stacked kernel), random-init and a
    # "stress" set (15x15 kernel x5, initial state x10 so tanh leaves its linear region), and the deep variant
    # 32 channels / T = 16
    gen_hgru(hm, aux, "S15_k25_T8", n=1, h=32, w=32, k=25, S=15, T=8, seed=14, f32_steps=True)
    gen_hgru(hm, aux, "S15_k25_T8_stress", n=1, h=32, w=32, k=25, S=15, T=8, seed=15, f32_steps=True,
             gains={(15, 15, 25, 25): 5.0, (1, 32, 32, 25): 10.0})
    gen_hgru(hm, aux, "S15_k32_T16_stress", n=1, h=20, w=24, k=32, S=15, T=16, seed=16, f32_steps=True,
             gains={(15, 15, 32, 32): 4.0, (1, 20, 24, 32): 10.0})
    # the reference's own width (64 channels: the plain tcgen05 conv kernel with epilogue-issued gates), small map
    gen_hgru(hm, aux, "S15_k64_T3_stress", n=1, h=16, w=20, k=64, S=15, T=3, seed=17, f32_steps=True,
             gains={(15, 15, 64, 64): 4.0, (1, 16, 20, 64): 10.0})
    gen_pose_layers(pm, seed=21)


if __name__ == "__main__":
    main()
