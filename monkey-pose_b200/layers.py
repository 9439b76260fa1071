"""Stand-alone layer methods shared by the two model mirrors (`hgru_pose.model`, `attn_model_struct`): the reference
classes carry the same helpers -- `conv_layer`, `max_pool`, `fc_layer`, `get_conv_var`, `get_fc_var`, `get_var`
(hgru_pose.py:134-216, train_cnn_networks_hgru.py:530-640) -- and call `tf.layers.batch_normalization` the same way.
Exact fp32, unfused kernels of libhgru_b200.so (`layer_*_forward`); the inference-mode `build()` of either model does
not go through them (fused tensor-core pipelines), `build(train_mode=True)` does.  No CPU fallback."""
import numpy as np
import torch

from . import _lib
from . import initialization as init
from .hgru_module import _as_dev, _stream

_BN_FIELDS = ("gamma", "beta", "moving_mean", "moving_variance")


class LayerOps(object):
    _bn_scopes = ()          # tf.layers.batch_normalization scope names in call order (set by the subclass)
    dropout_seed = 1234
    seed = 42

    def get_var(self, initial_value, name, idx, var_name, in_size=None, out_size=None):
        """hgru_pose.py:196-216 / train_cnn_networks_hgru.py:615-633: value from data_dict[name][idx] when present, else the initial
        value; registered in var_dict[(name, idx)]."""
        if self.data_dict is not None and name in self.data_dict:
            value = self.data_dict[name][idx]
        else:
            value = initial_value
        src = self.__dict__.setdefault("_var_src", {})
        if src.get((name, idx)) is value and (name, idx) in self.var_dict:
            return self.var_dict[(name, idx)]          # same host array as last time: keep its device copy
        var = _as_dev(value)
        self.var_dict[(name, idx)] = var
        src[(name, idx)] = value
        return var

    # -- layer methods (hgru_pose.py:107-194) -------------------------------------------------------
    # Stand-alone, exact fp32, unfused: what a caller composing the graph by hand gets.  build() in inference mode
    # does NOT go through them (it runs the fused tensor-core pipeline); build(train_mode=True) does.
    def _rng(self, name):
        import zlib
        return np.random.default_rng([self.seed, zlib.crc32(name.encode())])

    def _default(self, key, fallback):
        """Initial value of variable `key` ('conv_2/conv_2_filters', 'batch_normalization_3/gamma', ...): inside
        build() the seeded draw of the whole graph (the same one whichever path runs), else `fallback()`."""
        d = self.__dict__.get("_defaults")
        if d is not None and key in d:
            return d[key]
        return fallback()

    def get_conv_var(self, filter_size, in_channels, out_channels, name, init_type='xavier'):
        """hgru_pose.py:165-180: xavier-normal filters (or truncated_normal(0, .001)), truncated_normal(0, .001) biases,
        unless `data_dict[name]` supplies them."""
        rng = self._rng(name)
        shape = (filter_size, filter_size, in_channels, out_channels)
        w0 = self._default("%s/%s_filters" % (name, name), lambda: (
            init.xavier_normal(rng, shape) if init_type == 'xavier' else init._truncated_normal(rng, shape, 0.001)))
        b0 = self._default("%s/%s_biases" % (name, name), lambda: init._truncated_normal(rng, (out_channels,), 0.001))
        filters = self.get_var(w0, name, 0, name + "_filters")
        biases = self.get_var(b0, name, 1, name + "_biases")
        return filters, biases

    def get_fc_var(self, in_size, out_size, name, init_type='xavier'):
        """hgru_pose.py:182-194."""
        rng = self._rng(name)
        w0 = self._default("%s/%s_weights" % (name, name), lambda: (
            init.xavier_normal(rng, (in_size, out_size)) if init_type == 'xavier' else
            init._truncated_normal(rng, (in_size, out_size), 0.001)))
        b0 = self._default("%s/%s_biases" % (name, name), lambda: init._truncated_normal(rng, (out_size,), 0.001))
        weights = self.get_var(w0, name, 0, name + "_weights")
        biases = self.get_var(b0, name, 1, name + "_biases")
        return weights, biases

    @staticmethod
    def _need_cuda(t, what):
        if not (torch.is_tensor(t) and t.is_cuda):
            raise RuntimeError("%s needs a CUDA tensor (no CPU fallback)" % what)
        return t.to(torch.float32).contiguous()

    def conv_layer(self, bottom, in_channels, out_channels, name, filter_size=3, batchnorm=None,
                   stride=[1, 1, 1, 1]):
        """hgru_pose.py:139-154: relu(conv2d(bottom, filters, SAME) + biases); variables `<name>/<name>_filters`,
        `<name>/<name>_biases`."""
        if list(stride) != [1, 1, 1, 1]:
            raise NotImplementedError("conv_layer: stride != [1,1,1,1] is not on the path hgru_pose.py configures")
        if batchnorm is not None and name in batchnorm:
            raise NotImplementedError("conv_layer(batchnorm=...) (tf.nn.moments over the batch axis, hgru_pose.py:"
                                      "120-122) is never selected by hgru_pose.build")
        x = self._need_cuda(bottom, "conv_layer")
        if x.dim() != 4 or int(x.shape[-1]) != int(in_channels):
            raise ValueError("conv_layer: bottom must be [N,H,W,%d]" % in_channels)
        filt, bias = self.get_conv_var(filter_size, in_channels, out_channels, name)
        if tuple(filt.shape) != (filter_size, filter_size, in_channels, out_channels):
            raise ValueError("%s_filters has shape %s" % (name, tuple(filt.shape)))
        N, H, W = [int(v) for v in x.shape[:3]]
        out = torch.empty((N, H, W, int(out_channels)), device=x.device, dtype=torch.float32)
        _lib.check(_lib.load().layer_conv2d_forward(x.data_ptr(), N, H, W, int(in_channels), filt.data_ptr(),
                                                    int(filter_size), int(out_channels), bias.data_ptr(), 1,
                                                    out.data_ptr(), _stream()), "layer_conv2d_forward")
        return out

    def max_pool(self, bottom, name):
        """hgru_pose.py:134-137: tf.nn.max_pool 2x2, stride 2, SAME."""
        x = self._need_cuda(bottom, "max_pool")
        N, H, W, C = [int(v) for v in x.shape]
        out = torch.empty((N, (H + 1) // 2, (W + 1) // 2, C), device=x.device, dtype=torch.float32)
        _lib.check(_lib.load().layer_max_pool2x2_forward(x.data_ptr(), N, H, W, C, out.data_ptr(), _stream()),
                   "layer_max_pool2x2_forward")
        return out

    def _pool(self, bottom, ksize, average, what):
        x = self._need_cuda(bottom, what)
        N, H, W, C = [int(v) for v in x.shape]
        out = torch.empty((N, (H + ksize - 1) // ksize, (W + ksize - 1) // ksize, C), device=x.device,
                          dtype=torch.float32)
        _lib.check(_lib.load().layer_pool_same_forward(x.data_ptr(), N, H, W, C, ksize, 1 if average else 0,
                                                       out.data_ptr(), _stream()), "layer_pool_same_forward")
        return out

    def avg_pool(self, bottom, name):
        """hgru_pose.py:124-127: tf.nn.avg_pool 2x2, stride 2, SAME (declared by the reference, unused by its build)."""
        return self._pool(bottom, 2, True, "avg_pool")

    def max_pool_4(self, bottom, name):
        """hgru_pose.py:129-132: tf.nn.max_pool 4x4, stride 4, SAME (declared by the reference, unused by its build)."""
        return self._pool(bottom, 4, False, "max_pool_4")

    def batchnorm(self, layer):
        """hgru_pose.py:120-122: moments over axis 0, normalised without scale / offset, epsilon 1e-3 (declared by the
        reference, unused by its build)."""
        x = self._need_cuda(layer, "batchnorm")
        N = int(x.shape[0])
        out = torch.empty_like(x)
        _lib.check(_lib.load().layer_batchnorm_moments0_forward(x.data_ptr(), N, x.numel() // N, 1e-3, out.data_ptr(),
                                                                _stream()), "layer_batchnorm_moments0_forward")
        return out

    def fc_layer(self, bottom, in_size, out_size, name):
        """hgru_pose.py:156-163: reshape(bottom, [-1, in_size]) @ weights + biases."""
        x = self._need_cuda(bottom, "fc_layer").reshape(-1, int(in_size))
        w, b = self.get_fc_var(int(in_size), int(out_size), name)
        if tuple(w.shape) != (int(in_size), int(out_size)):
            raise ValueError("%s_weights has shape %s" % (name, tuple(w.shape)))
        out = torch.empty((int(x.shape[0]), int(out_size)), device=x.device, dtype=torch.float32)
        _lib.check(_lib.load().layer_fc_forward(x.data_ptr(), int(x.shape[0]), int(in_size), w.data_ptr(),
                                                b.data_ptr(), int(out_size), out.data_ptr(), _stream()),
                   "layer_fc_forward")
        return out

    def batch_normalization(self, inputs, index, training, relu_first=False, dropout_keep=1.0):
        """tf.layers.batch_normalization(inputs, axis=last, momentum=.997, epsilon=1e-5, center, scale, training,
        fused=True) as hgru_pose.py:52-103 calls it; `index` 0..4 = scope batch_normalization, _1 .. _4.  In training
        mode the batch statistics normalise and `self.updated_moving_stats[scope]` receives the moving statistics the
        reference's UPDATE_OPS would assign (train_cnn_networks_hgru.py:123-126)."""
        x = self._need_cuda(inputs, "batch_normalization")
        C = int(x.shape[-1])
        scope = self._bn_scopes[index]
        bn = init.bn_identity(C)
        g, b, mu, var = [self.get_var(self._default("%s/%s" % (scope, f), lambda f=f: bn[f]), scope, i, f)
                         for i, f in enumerate(_BN_FIELDS)]
        for t in (g, b, mu, var):
            if t.numel() != C:
                raise ValueError("%s variables must have %d elements" % (scope, C))
        rows = x.numel() // C
        y = torch.empty_like(x)
        new_mu = new_var = ws = None
        if training:
            new_mu, new_var = torch.empty_like(mu), torch.empty_like(var)
            ws = torch.empty(2 * C, device=x.device, dtype=torch.float64)
        _lib.check(_lib.load().layer_batch_norm_forward(
            x.data_ptr(), rows, C, g.data_ptr(), b.data_ptr(), mu.data_ptr(), var.data_ptr(),
            float(self._BATCH_NORM_EPSILON), 1 if training else 0, 1 if relu_first else 0, float(dropout_keep),
            int(self.dropout_seed), float(self._BATCH_NORM_DECAY),
            new_mu.data_ptr() if training else None, new_var.data_ptr() if training else None,
            ws.data_ptr() if training else None, y.data_ptr(), _stream()), "layer_batch_norm_forward")
        if training:
            self.updated_moving_stats[scope] = {"moving_mean": new_mu, "moving_variance": new_var}
        return y


    def resize_images(self, images, size):
        """tf.image.resize_images(images, size) as TF 1.x computes it (bilinear, align_corners=False, float32 without
        fused multiply-adds; train_cnn_networks_hgru.py:442).  images [N,H,W] or [N,H,W,1] -> [N,size[0],size[1],1]."""
        x = self._need_cuda(images, "resize_images")
        if x.dim() == 4:
            if int(x.shape[-1]) != 1:
                raise NotImplementedError("resize_images: one channel (depth frames) only")
            x = x[..., 0].contiguous()
        N, H, W = [int(v) for v in x.shape]
        out = torch.empty((N, int(size[0]), int(size[1]), 1), device=x.device, dtype=torch.float32)
        _lib.check(_lib.load().layer_resize_bilinear_forward(x.data_ptr(), N, H, W, int(size[0]), int(size[1]),
                                                             out.data_ptr(), _stream()), "layer_resize_bilinear_forward")
        return out
