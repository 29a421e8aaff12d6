"""Stand-in for the two tensorflow_probability entry points the reference's hot path uses (see ../tensorflow)."""
import torch as _torch

import tensorflow as _tf


class _Math:
    @staticmethod
    def reduce_weighted_logsumexp(logx, w=None, axis=None, keep_dims=False, return_sign=False, name=None):
        """log|sum_i w_i exp(logx_i)| and its sign (tfp.math.reduce_weighted_logsumexp): max-shifted, like tfp."""
        logx = _tf._t(logx)
        if w is None:
            out = _tf.reduce_logsumexp(logx, axis=axis, keepdims=keep_dims)
            return (out, _torch.ones_like(out)) if return_sign else out
        w = _tf._t(w, logx.dtype)
        ax = _tf._axis(axis)
        log_absw_x = logx + _torch.log(_torch.abs(w))
        max_log = _torch.amax(log_absw_x, dim=ax, keepdim=True) if ax is not None else log_absw_x.max()
        max_log = _torch.where(_torch.isinf(max_log), _torch.zeros_like(max_log), max_log)
        wx_over_max = _torch.sign(w) * _torch.exp(log_absw_x - max_log)
        s = wx_over_max.sum(dim=ax, keepdim=keep_dims) if ax is not None else wx_over_max.sum()
        if not keep_dims and ax is not None:
            max_log = max_log.squeeze(ax)
        sgn = _torch.sign(s)
        lswe = max_log + _torch.log(sgn * s)
        return (lswe, sgn) if return_sign else lswe


math = _Math()


class _Normal:
    def __init__(self, loc, scale):
        self.loc, self.scale = _tf._t(loc), _tf._t(scale)

    def prob(self, x):
        x = _tf._t(x, self.loc.dtype)
        z = (x - self.loc) / self.scale
        return _torch.exp(-0.5 * z * z) / (self.scale * (2 * 3.141592653589793) ** 0.5)

    def log_prob(self, x):
        return _torch.log(self.prob(x))


class _Distributions:
    Normal = _Normal


distributions = _Distributions()
