"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's attention (centre-of-mass) CNN forward,
`attn_model_struct.build` (train_cnn_networks_hgru.py:440-525; helpers :530-640).  Never imported by the product
(`monkey-pose_b200/`); only tests/, __graft_entry__.smoke() and bench.py's CPU leg may use it.

    resize to 128x128 (tf.image.resize_images: bilinear, align_corners=False)             :442
    5 x [conv SxS + bias + relu -> max-pool 2x2 -> batch-norm]   S = 3,3,3,3,5             :443-499
    fc afc_1 (flatten h,w,c -> 1024) -> relu -> batch-norm -> fc afc_out (1024 -> 3)       :501-525

PINNING: pinned.  tests/golden/attn_ref.npz holds the outputs of the reference's own class source executed under
a numpy TensorFlow stand-in (tests/golden/make_golden_attn.py); tests/test_attn_oracle_golden.py checks this
restatement against every stored activation.  Third-party arithmetic restated here because TensorFlow 1.x
(unpinned version) is absent: ResizeBilinear (kernels/resize_bilinear_op.cc: scale = in/out, src = dst*scale,
lower = floor, upper = min(lower+1, in-1), float32 lerps), conv2d / max_pool / batch_normalization / matmul.
Defect resolution: `axis=3` on the rank-2 fc tensor (:503) -> last axis (SURVEY R-D5).
"""
import numpy as np

from . import hgru_oracle_np as onp

BN_SCOPES = ("batch_normalization", "batch_normalization_1", "batch_normalization_2", "batch_normalization_3",
             "batch_normalization_4", "batch_normalization_5")
CONV_NAMES = ("aconv_1", "aconv_2", "aconv_3", "aconv_4", "aconv_5")


def resize_bilinear_tf1(x, oh, ow):
    """tf.image.resize_images(x, [oh, ow]) as TF 1.x computes it, float32 arithmetic (train_cnn_networks_hgru.py:442)."""
    x = np.asarray(x, np.float32)
    n, h, w, c = x.shape
    if (oh, ow) == (h, w):
        return x.copy()

    def weights(insz, outsz):
        scale = np.float32(insz) / np.float32(outsz)
        src = np.arange(outsz, dtype=np.float32) * scale
        lo = np.floor(src).astype(np.int64)
        return lo, np.minimum(lo + 1, insz - 1), (src - lo.astype(np.float32)).astype(np.float32)

    y0, y1, yl = weights(h, oh)
    x0, x1, xl = weights(w, ow)
    xl = xl.reshape(1, 1, ow, 1)
    yl = yl.reshape(1, oh, 1, 1)
    tl, tr = x[:, y0][:, :, x0], x[:, y0][:, :, x1]
    bl, br = x[:, y1][:, :, x0], x[:, y1][:, :, x1]
    top = (tl + (tr - tl) * xl).astype(np.float32)
    bot = (bl + (br - bl) * xl).astype(np.float32)
    return (top + (bot - top) * yl).astype(np.float32)


def _bn(x, P, scope, eps, train_mode=False):
    if train_mode:      # tf.layers.batch_normalization(training=True): statistics of the batch
        return onp.batch_norm_training(x, P[scope + "/gamma"], P[scope + "/beta"], eps)
    return onp.batch_norm_inference(x, P[scope + "/gamma"], P[scope + "/beta"], P[scope + "/moving_mean"],
                                    P[scope + "/moving_variance"], eps)


def attn_forward(frames, params, eps=1e-5, trace=False, train_mode=False, dropout_keep=None, dropout_seed=0):
    """frames [N,H,W,1] (depth / image_max_depth); params keyed by the reference's variable names
    (`aconv_1/aconv_1_filters`, ..., `afc_out/afc_out_biases`, `batch_normalization[_i]/...`).  Returns
    out_put [N,3] (float64), plus the intermediate activations when trace.  train_mode: batch statistics in the six
    batch norms (:446-513, `training=train_mode`); dropout (:504-505) only when `dropout_keep` is given, with the
    documented counter-based mask (hgru_oracle_np.dropout_keep_mask)."""
    P = params
    acts = {"resized": resize_bilinear_tf1(frames, 128, 128)}
    x = acts["resized"].astype(np.float64)
    for i, name in enumerate(CONV_NAMES):
        conv = onp.conv_layer(x, P["%s/%s_filters" % (name, name)], P["%s/%s_biases" % (name, name)])   # conv+bias+relu
        x = _bn(onp.max_pool_2x2(conv), P, BN_SCOPES[i], eps, train_mode)
        acts["pool%d" % (i + 1)] = x
    fc1 = onp.fc_layer(x, P["afc_1/afc_1_weights"], P["afc_1/afc_1_biases"])
    r = np.maximum(fc1, 0.0)
    if dropout_keep is not None and dropout_keep < 1.0:
        keep = onp.dropout_keep_mask(r.size, dropout_keep, dropout_seed).reshape(r.shape)
        r = np.where(keep, r / np.float64(np.float32(dropout_keep)), 0.0)
    relu1 = _bn(r, P, BN_SCOPES[5], eps, train_mode)
    out = onp.fc_layer(relu1, P["afc_out/afc_out_weights"], P["afc_out/afc_out_biases"])
    acts.update(fc1=fc1, relu1=relu1, out_put=out)
    return (out, acts) if trace else out
