"""TEST INFRASTRUCTURE ONLY -- torch-CPU float32 restatement of the hGRU-pose forward.

Same arithmetic as `hgru_oracle_np.py` (which is the arbiter, float64, pinned against the
reference source through tests/golden/), but on `torch.nn.functional.conv2d` so BASELINE-sized
shapes finish in seconds.  It is (i) the full-size parity checker and (ii) the CPU baseline
("port" kind: TensorFlow is absent, this multi-threaded torch-CPU forward stands in for the
TF-CPU forward -- BASELINE.md "CPU baseline plan").  Never imported by the product package.

NHWC at the API (reference layout); NCHW internally.  Line citations: see hgru_oracle_np.py.
"""
import torch
import torch.nn.functional as F

HGRU_PARAM_NAMES = ("p_r", "i_r", "i_b", "o_r", "o_b", "beta", "nu", "gamma", "kappa", "omega",
                    "rho", "lateral_bias")
BN_SCOPES = ("batch_normalization", "batch_normalization_1", "batch_normalization_2",
             "batch_normalization_3", "batch_normalization_4")


def _t(x, dtype):
    return torch.as_tensor(x).to(dtype)


def _w(w, dtype):       # HWIO -> OIHW
    return _t(w, dtype).permute(3, 2, 0, 1).contiguous()


def _vec(v, dtype):     # [1,1,1,k] -> [1,k,1,1]
    return _t(v, dtype).reshape(1, -1, 1, 1)


def _conv_same(x, w):
    fh, fw = w.shape[2], w.shape[3]
    return F.conv2d(x, w, padding=((fh - 1) // 2, (fw - 1) // 2))


def hgru_forward_nchw(X, H2, p, timesteps, trace=None):
    """hgru_module.py:825-857 per step (see hgru_oracle_np.hgru_step for the line map)."""
    for t in range(timesteps):
        G1 = torch.sigmoid(_conv_same(H2, p["i_r"]) + p["i_b"])
        C1 = _conv_same(H2 * G1, p["p_r"]) + p["lateral_bias"]
        H1 = torch.tanh(X - (p["beta"] * H2 + p["nu"]) * C1)
        G2 = torch.sigmoid(_conv_same(H1, p["o_r"]) + p["o_b"])
        C2 = _conv_same(H1, p["p_r"]) + p["lateral_bias"]
        e = p["gamma"] * C2
        Ht = torch.tanh(p["kappa"] * (H1 + e) + p["omega"] * (H1 * e))
        H2 = (G2 * H2 + (1.0 - G2) * Ht) * p["rho"][t]
        if trace is not None:
            trace.append((H1.permute(0, 2, 3, 1).contiguous(), H2.permute(0, 2, 3, 1).contiguous()))
    return H2


def _prep_hgru(params, dtype):
    p = {}
    for n in ("p_r", "i_r", "o_r"):
        p[n] = _w(params[n], dtype)
    for n in ("i_b", "o_b", "beta", "nu", "gamma", "kappa", "omega", "lateral_bias"):
        p[n] = _vec(params[n], dtype)
    p["rho"] = _t(params["rho"], dtype)
    return p


def hgru_forward(X, H2_init, params, timesteps, dtype=torch.float32, trace=False):
    """NHWC in / NHWC out wrapper."""
    with torch.no_grad():
        x = _t(X, dtype).permute(0, 3, 1, 2).contiguous()
        h = _t(H2_init, dtype).permute(0, 3, 1, 2).contiguous()
        tr = [] if trace else None
        out = hgru_forward_nchw(x, h, _prep_hgru(params, dtype), timesteps, tr)
        out = out.permute(0, 2, 3, 1).contiguous()
        if trace:
            return out, [a for a, _ in tr], [b for _, b in tr]
        return out


def _bn(x, P, scope, dtype, eps, channel_dim):
    g = _t(P[scope + "/gamma"], dtype)
    b = _t(P[scope + "/beta"], dtype)
    m = _t(P[scope + "/moving_mean"], dtype)
    v = _t(P[scope + "/moving_variance"], dtype)
    shape = [1] * x.dim()
    shape[channel_dim] = -1
    return (x - m.reshape(shape)) / torch.sqrt(v.reshape(shape) + eps) * g.reshape(shape) + b.reshape(shape)


def pose_forward(depth, P, H2_init, timesteps=8, dtype=torch.float32, eps=1e-5, trace=False):
    """hgru_pose.py:47-105, inference-mode BN, resolutions R-D4/R-D5/R-D6.  H2_init: O_0 [N,HW,HW,k], or the reference's
    hidden_init names 'identity' (O_0 = the hGRU's input, hgru_module.py:876-878) / 'zeros' (:888-890)."""
    with torch.no_grad():
        x = _t(depth, dtype).permute(0, 3, 1, 2).contiguous()
        c1 = torch.relu(_conv_same(x, _w(P["conv_1/conv_1_filters"], dtype))
                        + _t(P["conv_1/conv_1_biases"], dtype).reshape(1, -1, 1, 1))
        p1 = _bn(F.max_pool2d(c1, 2, 2), P, BN_SCOPES[0], dtype, eps, 1)
        c2 = torch.relu(_conv_same(p1, _w(P["conv_2/conv_2_filters"], dtype))
                        + _t(P["conv_2/conv_2_biases"], dtype).reshape(1, -1, 1, 1))
        c2 = _bn(c2, P, BN_SCOPES[1], dtype, eps, 1)
        c3 = torch.relu(_conv_same(c2, _w(P["conv_3/conv_3_filters"], dtype))
                        + _t(P["conv_3/conv_3_biases"], dtype).reshape(1, -1, 1, 1))
        c3 = _bn(c3, P, BN_SCOPES[2], dtype, eps, 1)
        hp = _prep_hgru({n: P["contextual_circuit/" + n] for n in HGRU_PARAM_NAMES}, dtype)
        if isinstance(H2_init, str):
            if H2_init not in ("identity", "zeros"):
                raise RuntimeError("hidden_init")                          # hgru_module.py:891-892
            h0 = c3.clone() if H2_init == "identity" else torch.zeros_like(c3)
        else:
            h0 = _t(H2_init, dtype).permute(0, 3, 1, 2).contiguous()
        hg = hgru_forward_nchw(c3, h0, hp, timesteps)
        hb = _bn(hg, P, BN_SCOPES[3], dtype, eps, 1)
        flat = hb.permute(0, 2, 3, 1).reshape(hb.shape[0], -1)            # flatten order h,w,c
        fc1 = flat @ _t(P["fc_1/fc_1_weights"], dtype) + _t(P["fc_1/fc_1_biases"], dtype)
        r1 = _bn(torch.relu(fc1), P, BN_SCOPES[4], dtype, eps, 1)
        out = r1 @ _t(P["fc_out/fc_out_weights"], dtype) + _t(P["fc_out/fc_out_biases"], dtype)
        if trace:
            return out, {"conv3": c3.permute(0, 2, 3, 1).contiguous(),
                         "hgru": hg.permute(0, 2, 3, 1).contiguous(), "fc1": fc1, "out_put": out}
        return out
