"""TEST INFRASTRUCTURE ONLY -- torch-CPU twin of oracle/attn_oracle_np.py (same equations, library convolutions),
used where the numpy restatement is too slow: the reference's real widths (64..1024 channels) and the CPU baseline
of bench.py.  Cross-checked against the numpy arbiter (itself pinned to the reference's golden vectors) in
tests/test_attn_oracle_golden.py.  Never imported by the product."""
import numpy as np
import torch
import torch.nn.functional as F

from .attn_oracle_np import BN_SCOPES, CONV_NAMES, resize_bilinear_tf1


def attn_forward(frames, P, dtype=torch.float64, eps=1e-5, trace=False):
    t = lambda a: torch.as_tensor(np.asarray(a)).to(dtype)
    x = t(resize_bilinear_tf1(frames, 128, 128)).permute(0, 3, 1, 2)            # NCHW
    acts = {}
    for i, name in enumerate(CONV_NAMES):
        w = t(P["%s/%s_filters" % (name, name)]).permute(3, 2, 0, 1)             # HWIO -> OIHW
        x = F.relu(F.conv2d(x, w, t(P["%s/%s_biases" % (name, name)]), padding=w.shape[-1] // 2))
        x = F.max_pool2d(x, 2)
        s = BN_SCOPES[i]
        g, b, m, v = [t(P[s + "/" + f]).view(1, -1, 1, 1) for f in ("gamma", "beta", "moving_mean", "moving_variance")]
        x = (x - m) / torch.sqrt(v + eps) * g + b
        acts["pool%d" % (i + 1)] = x.permute(0, 2, 3, 1)
    flat = x.permute(0, 2, 3, 1).reshape(x.shape[0], -1)                         # tf.reshape of NHWC
    fc1 = flat @ t(P["afc_1/afc_1_weights"]) + t(P["afc_1/afc_1_biases"])
    s = BN_SCOPES[5]
    g, b, m, v = [t(P[s + "/" + f]) for f in ("gamma", "beta", "moving_mean", "moving_variance")]
    relu1 = (F.relu(fc1) - m) / torch.sqrt(v + eps) * g + b
    out = relu1 @ t(P["afc_out/afc_out_weights"]) + t(P["afc_out/afc_out_biases"])
    acts.update(fc1=fc1, relu1=relu1, out_put=out)
    return (out, acts) if trace else out
