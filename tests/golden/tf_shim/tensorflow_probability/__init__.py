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


# ---- the distributions the reference's TARGETS are written with (student_t_mixture.py:40-44, planar_robot.py:34,46) --------
# Closed forms as documented by TFP, written with differentiable torch ops so that the reference's GradientTape
# (sample_selector.py:73-77) works through them.
class _Categorical:
    def __init__(self, logits=None, probs=None):
        logits = _tf._t(logits) if logits is not None else _torch.log(_tf._t(probs))
        self.logits = logits - _torch.logsumexp(logits, dim=-1, keepdim=True)


class _MultivariateNormalDiag:
    def __init__(self, loc=None, scale_diag=None):
        if isinstance(scale_diag, (list, tuple)):      # Python floats become float32 constants in TF (planar_robot.py:46)
            scale_diag = _torch.tensor(scale_diag, dtype=_torch.float32).to(_torch.float64)
        self.scale_diag = _tf._t(scale_diag)
        self.loc = _tf._t(loc, self.scale_diag.dtype)

    def log_prob(self, x):
        x = _tf._t(x, self.loc.dtype)
        z = (x - self.loc) / self.scale_diag
        d = self.loc.shape[-1]
        return (-0.5 * _torch.sum(z * z, dim=-1) - _torch.sum(_torch.log(self.scale_diag), dim=-1)
                - 0.5 * d * _torch.log(_torch.tensor(2 * _torch.pi, dtype=self.loc.dtype)))


class _StudentT:
    """Scalar Student-t, batch of (df, loc, scale)."""
    def __init__(self, df, loc, scale):
        self.loc, self.scale = _tf._t(loc), _tf._t(scale)
        self.df = _tf._t(float(df), self.loc.dtype)

    def log_prob(self, x):
        x = _tf._t(x, self.loc.dtype)
        y = (x - self.loc) / self.scale
        df = self.df
        return (_torch.lgamma(0.5 * (df + 1)) - _torch.lgamma(0.5 * df) - 0.5 * _torch.log(df * _torch.pi)
                - _torch.log(self.scale) - 0.5 * (df + 1) * _torch.log1p(y * y / df))


class _MultivariateStudentTLinearOperator:
    """log p(x) = lgamma((df+d)/2) - lgamma(df/2) - d/2 log(df pi) - log|det scale| - (df+d)/2 log1p(|scale^-1 (x-loc)|^2 / df)."""
    def __init__(self, df, loc, scale):
        self.loc = _tf._t(loc)
        self.scale = scale                       # tf.linalg.LinearOperatorLowerTriangular
        self.df = _tf._t(float(df), self.loc.dtype)

    def log_prob(self, x):
        x = _tf._t(x, self.loc.dtype)
        L = self.scale.tril
        d = self.loc.shape[-1]
        diff = (x - self.loc).unsqueeze(-1)                                   # [..., K, d, 1] by broadcasting
        z = _torch.linalg.solve_triangular(L.expand(diff.shape[:-2] + L.shape[-2:]), diff, upper=False).squeeze(-1)
        maha = _torch.sum(z * z, dim=-1)
        logdet = _torch.sum(_torch.log(_torch.abs(_torch.diagonal(L, dim1=-2, dim2=-1))), dim=-1)
        df = self.df
        return (_torch.lgamma(0.5 * (df + d)) - _torch.lgamma(0.5 * df) - 0.5 * d * _torch.log(df * _torch.pi) - logdet
                - 0.5 * (df + d) * _torch.log1p(maha / df))


class _MixtureSameFamily:
    def __init__(self, mixture_distribution, components_distribution):
        self.mix, self.comp = mixture_distribution, components_distribution

    def log_prob(self, x):
        x = _tf._t(x, self.mix.logits.dtype)
        lp = self.comp.log_prob(x.unsqueeze(-1) if isinstance(self.comp, _StudentT) else x.unsqueeze(-2))   # [..., K]
        return _torch.logsumexp(lp + self.mix.logits, dim=-1)


_Distributions.Categorical = _Categorical
_Distributions.MultivariateNormalDiag = _MultivariateNormalDiag
_Distributions.StudentT = _StudentT
_Distributions.MultivariateStudentTLinearOperator = _MultivariateStudentTLinearOperator
_Distributions.MixtureSameFamily = _MixtureSameFamily
