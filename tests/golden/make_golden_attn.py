"""Golden vectors for the attention (centre-of-mass) CNN, produced by EXECUTING THE REFERENCE'S OWN CLASS.

Run in the build container only (needs /root/reference; the GPU box never runs this):

    python tests/golden/make_golden_attn.py

How: the source text of `class attn_model_struct` (train_cnn_networks_hgru.py:422-640: build, conv_layer,
max_pool, fc_layer, get_*_var) is cut out of the reference file by line (the module itself cannot be imported:
TensorFlow 1.x, Python 2, data-loader imports) and exec'd in memory with `tf` resolved to tests/golden/tf_shim.py
(numpy float64; bilinear resize in float32 as TF computes it).  Nothing from the reference is written into this
repository except the numbers it computes.

Weights are injected through the class's own `data_dict` hook with `trainable=False` (the `tf.constant` branch of
get_var, :615-633), so the widths can be small -- get_conv_var ignores its in/out channel arguments when the name is
in data_dict -- and the run takes seconds.  Batch-norm statistics are injected through the shim.
Defect resolution (same as hgru_pose.py, SURVEY R-D5): `axis=3` on the rank-2 fc tensor (:501-509) -> last axis.

Output: attn_ref.npz = input frames, every injected variable under its TF name, resized input, pool1..pool5, fc1,
relu1 (after batch-norm) and out_put.
"""
import ast
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/train_cnn_networks_hgru.py"
sys.path.insert(0, HERE)
import tf_shim  # noqa: E402

WIDTHS = (8, 16, 16, 24, 32)     # reference: 64, 128, 256, 512, 1024
FC_HIDDEN, OUT = 1024, 3         # 1024 is hard-coded in the reference's afc_out reshape (:522); num_dims = 3


def load_reference_class():
    # (the file holds Python-2 print statements elsewhere, so it is cut by line, not parsed as a whole)
    lines = open(REF).read().splitlines()
    start = [i for i, ln in enumerate(lines) if ln.startswith("class attn_model_struct")][0]
    end = [i for i, ln in enumerate(lines) if i > start and ln.startswith("class ")][0]
    text = "\n".join(lines[start:end])
    ast.parse(text)                                       # the class itself is valid Python 3
    text = text.replace('print ("attention network")', 'print("attention network")')
    tfm = types.ModuleType("tensorflow")
    for k in dir(tf_shim):
        if not k.startswith("__"):
            setattr(tfm, k, getattr(tf_shim, k))
    ns = {"tf": tfm, "np": np}
    exec(compile(text, REF, "exec"), ns)
    return ns["attn_model_struct"]


def main():
    cls = load_reference_class()
    rng = np.random.default_rng(77)
    tf_shim.reset(77)
    n, h, w = 2, 53, 64                                   # Kinect aspect ratio, reduced; resized to 128x128
    frames = rng.uniform(0.05, 0.4, size=(n, h, w, 1)).astype(np.float32)
    frames[rng.uniform(size=frames.shape) < 0.05] = 0.0
    dd, cin = {}, 1
    for i, co in enumerate(WIDTHS):
        fs = 5 if i == 4 else 3
        name = "aconv_%d" % (i + 1)
        std = np.sqrt(2.0 / (fs * fs * cin))
        dd[name] = [rng.normal(0, std, size=(fs, fs, cin, co)).astype(np.float32),
                    rng.normal(0, 0.1, size=(co,)).astype(np.float32)]
        cin = co
    flat = 4 * 4 * WIDTHS[-1]
    # (fc weights on a 1/1024 grid: the 512 x 1024 matrix then compresses to a small fixture)
    dd["afc_1"] = [(np.round(rng.normal(0, np.sqrt(2.0 / flat), size=(flat, FC_HIDDEN)) * 1024) / 1024).astype(np.float32),
                   rng.normal(0, 0.1, size=(FC_HIDDEN,)).astype(np.float32)]
    dd["afc_out"] = [rng.normal(0, np.sqrt(1.0 / FC_HIDDEN), size=(FC_HIDDEN, OUT)).astype(np.float32),
                     rng.normal(0, 0.1, size=(OUT,)).astype(np.float32)]
    for i, c in enumerate(WIDTHS + (FC_HIDDEN,)):
        scope = "batch_normalization" if i == 0 else "batch_normalization_%d" % i
        tf_shim.REG.bn_preset[scope] = (rng.uniform(0.5, 1.5, c), rng.normal(0, 0.2, c), rng.normal(0, 0.3, c),
                                        rng.uniform(0.5, 2.0, c))
    m = cls(trainable=False)
    m.data_dict = dd
    m.build(tf_shim._t(frames.astype(np.float64)), OUT, train_mode=False)
    out = {"frames": frames, "widths": np.array(WIDTHS), "fc_hidden": FC_HIDDEN}
    for name, (wt, b) in dd.items():
        kind = "weights" if name.startswith("afc") else "filters"
        out["var:%s/%s_%s" % (name, name, kind)] = wt
        out["var:%s/%s_biases" % (name, name)] = b
    for k, v in tf_shim.REG.variables.items():
        out["var:" + k] = v
    for k in ("pool1", "pool2", "pool3", "pool4", "pool5", "fc1", "relu1", "out_put"):
        a = np.asarray(getattr(m, k), np.float64)
        out["act:" + k] = a if a.ndim == 2 else a.astype(np.float32)      # feature maps: float32 keeps it small
    out["act:resized"] = np.asarray(tf_shim.image.resize_images(frames, [128, 128]), np.float32)
    path = os.path.join(HERE, "attn_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote %s (%.1f KB); out_put =\n%s" % (path, os.path.getsize(path) / 1024.0, out["act:out_put"]))


if __name__ == "__main__":
    main()
