"""A stand-in for the part of the TensorFlow 2 API that OlegArenz/gmmvi's hot path uses, on top of torch CPU tensors.

TEST INFRASTRUCTURE ONLY.  TensorFlow cannot be installed in this image (no wheel, no network), so the reference's own
Python sources (/root/reference/src/gmmvi/{models,optimization}/...) are executed UNMODIFIED against this module by
tests/golden/make_reference_golden.py to produce the golden vectors under tests/golden/reference_*.npz.  Each function
follows the documented semantics of the TF op of the same name that matter to the reference (eager mode; tf.function is
the identity, Python control flow on tensors runs eagerly like autograph's result; a failed Cholesky yields NaNs instead
of raising, as TF >= 2.5 does on CPU; unique ops keep first-occurrence order).  Floating point "tf.float32" maps to the
dtype chosen with set_float_dtype(): float64 for the logic pin (compared with the oracle's fp64 mode to ~1e-9) or float32
to mimic the reference's precision.  tf.float32.min / .max stay the float32 limits (they are constants of the algorithm,
e.g. the initial reward history, gmm_wrapper.py:72).  Random draws come from hooks so that the generator script can feed
the same noise to the oracle."""
from __future__ import annotations

import builtins
import math as _math

import numpy as _np
import torch as _torch

_FLOAT = _torch.float64
_INT = _torch.int64


def set_float_dtype(dt):
    global _FLOAT
    _FLOAT = dt


class DType:
    def __init__(self, name, kind):
        self.name, self.kind = name, kind

    @property
    def torch(self):
        return {"f": _FLOAT, "i": _INT, "b": _torch.bool}[self.kind]

    @property
    def min(self):
        return float(_np.finfo(_np.float32).min) if self.kind == "f" else int(_np.iinfo(_np.int32).min)

    @property
    def max(self):
        return float(_np.finfo(_np.float32).max) if self.kind == "f" else int(_np.iinfo(_np.int32).max)

    def __repr__(self):
        return f"tf.{self.name}"


float32 = DType("float32", "f")
float64 = DType("float64", "f")
int32 = DType("int32", "i")
int64 = DType("int64", "i")
bool = DType("bool", "b")  # noqa: A001
Tensor = _torch.Tensor


def _dt(dtype):
    if dtype is None:
        return None
    if isinstance(dtype, DType):
        return dtype.torch
    if dtype in (_torch.float32, _torch.float64):
        return _FLOAT
    return dtype


def _t(x, dtype=None):
    """Anything -> torch tensor (python floats / float arrays become the current float dtype, ints int64)."""
    dtype = _dt(dtype)
    if isinstance(x, _torch.Tensor):
        if dtype is not None and x.dtype != dtype:
            return x.to(dtype)
        if dtype is None and x.dtype in (_torch.float32, _torch.float64) and x.dtype != _FLOAT:
            return x.to(_FLOAT)
        return x
    if isinstance(x, (list, tuple)) and len(x) > 0 and any(isinstance(e, _torch.Tensor) for e in x):
        return _t(_torch.stack([_t(e) for e in x]), dtype)
    a = _np.asarray(x)
    if dtype is None:
        dtype = _FLOAT if a.dtype.kind == "f" else (_torch.bool if a.dtype.kind == "b" else _INT)
    return _torch.as_tensor(a).to(dtype)


class Variable(_torch.Tensor):
    """tf.Variable: a tensor whose value (and, with shape=[None, ...], leading dimension) can be re-assigned."""

    @staticmethod
    def __new__(cls, initial_value, shape=None, dtype=None, trainable=None, name=None, **_):
        t = _t(initial_value, dtype).detach().clone()
        return _torch.Tensor._make_subclass(cls, t)

    def __init__(self, *a, **k):
        pass

    __torch_function__ = _torch._C._disabled_torch_function_impl

    def assign(self, value):
        self.data = _t(value, self.dtype).detach().clone().reshape(_t(value).shape)
        return self

    def assign_add(self, value):
        self.data = (self.data + _t(value, self.dtype)).detach()
        return self

    def value(self):
        return self.data.clone()

    def __deepcopy__(self, memo):
        return Variable(self.data)


_orig_numpy = _torch.Tensor.numpy


def _numpy(self, *a, **k):
    return _orig_numpy(self.detach(), *a, **k)


_torch.Tensor.numpy = _numpy


def function(func=None, **_):
    if func is None:
        return lambda f: f
    return func


class TensorSpec:
    def __init__(self, shape=None, dtype=None, name=None):
        self.shape, self.dtype = shape, dtype


class TensorArray:
    def __init__(self, dtype=None, size=0, dynamic_size=False, infer_shape=True, clear_after_read=None, element_shape=None,
                 **_):
        self._items = [None] * int(size)
        self._dtype = dtype

    def write(self, index, value):
        index = int(index)
        while len(self._items) <= index:
            self._items.append(None)
        self._items[index] = _t(value, self._dtype)
        return self

    def read(self, index):
        return self._items[int(index)]

    def size(self):
        return len(self._items)

    def stack(self):
        if not self._items:
            return _torch.zeros((0,), dtype=_dt(self._dtype) or _FLOAT)
        return _torch.stack(self._items)

    def concat(self):
        return _torch.cat([i for i in self._items], dim=0)

    def gather(self, indices):
        return _torch.stack([self._items[int(i)] for i in indices])


class GradientTape:
    def __init__(self, persistent=False, watch_accessed_variables=True):
        self._watched = []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def watch(self, x):
        if not x.requires_grad:
            x.requires_grad_(True)
            self._watched.append(x)

    def gradient(self, target, sources):
        g = _torch.autograd.grad(target, sources, grad_outputs=_torch.ones_like(target), allow_unused=True)[0]
        for x in self._watched:
            x.requires_grad_(False)
        return None if g is None else g.detach()


# ---------------------------------------------------------------------------------------------- creation / shape
def constant(value, dtype=None, shape=None, name=None):
    return _t(value, dtype).clone()


def convert_to_tensor(value, dtype=None, **_):
    return _t(value, dtype)


def zeros(shape, dtype=float32, **_):
    return _torch.zeros(_shape(shape), dtype=_dt(dtype))


def ones(shape, dtype=float32, **_):
    return _torch.ones(_shape(shape), dtype=_dt(dtype))


def zeros_like(x, dtype=None):
    return _torch.zeros_like(_t(x), dtype=_dt(dtype))


def ones_like(x, dtype=None):
    return _torch.ones_like(_t(x), dtype=_dt(dtype))


def eye(n, dtype=float32, **_):
    return _torch.eye(int(n), dtype=_dt(dtype))


def _shape(shape):
    if isinstance(shape, _torch.Tensor):
        return tuple(int(s) for s in shape.reshape(-1))
    if isinstance(shape, (int, _np.integer)):
        return (int(shape),)
    return tuple(int(s) for s in shape)


def shape(x, out_type=None):
    return _torch.tensor(list(_t(x).shape), dtype=_INT)


def size(x, out_type=None):
    return _torch.tensor(_t(x).numel(), dtype=_INT)


def rank(x):
    return _torch.tensor(_t(x).dim(), dtype=_INT)


def range(*args, start=None, limit=None, delta=None, dtype=None, **_):  # noqa: A001
    if start is not None or limit is not None:
        args = tuple(a for a in (start if start is not None else 0, limit, delta) if a is not None)
    vals = [int(a) if not (isinstance(a, float)) else a for a in args]
    return _torch.arange(*vals, dtype=_dt(dtype) or _INT)


def cast(x, dtype):
    return _t(x).to(_dt(dtype))


def reshape(x, shape):
    return _t(x).reshape(_shape(shape))


def squeeze(x, axis=None):
    x = _t(x)
    return x.squeeze() if axis is None else x.squeeze(axis)


def expand_dims(x, axis):
    return _t(x).unsqueeze(axis)


def transpose(x, perm=None):
    x = _t(x)
    if perm is None:
        return x.permute(*reversed(builtins.range(x.dim())))
    return x.permute(*perm)


def concat(values, axis):
    return _torch.cat([_t(v) for v in values], dim=axis)


def stack(values, axis=0):
    return _torch.stack([_t(v) for v in values], dim=axis)


def gather(params, indices, axis=0, **_):
    idx = _t(indices).long()
    return _torch.index_select(_t(params), axis, idx.reshape(-1)).reshape(
        tuple(_t(params).shape[:axis]) + tuple(idx.shape) + tuple(_t(params).shape[axis + 1:]))


def repeat(x, repeats, axis=None):
    return _torch.repeat_interleave(_t(x), _t(repeats).long(), dim=axis)


def where(cond, x=None, y=None):
    if x is None:
        return _torch.nonzero(_t(cond))
    return _torch.where(_t(cond), _t(x), _t(y))


def scatter_nd(indices, updates, shape):
    out = _torch.zeros(_shape(shape), dtype=_t(updates).dtype)
    idx = _t(indices).long()
    out.index_put_(tuple(idx[..., d] for d in builtins.range(idx.shape[-1])), _t(updates), accumulate=True)
    return out


def tensor_scatter_nd_update(tensor, indices, updates):
    out = _t(tensor).clone()
    idx = _t(indices).long()
    out.index_put_(tuple(idx[..., d] for d in builtins.range(idx.shape[-1])), _t(updates, out.dtype))
    return out


def sort(x, axis=-1, direction="ASCENDING"):
    return _torch.sort(_t(x), dim=axis, descending=direction != "ASCENDING").values


def argmax(x, axis=None, output_type=None):
    x = _t(x)
    if x.dtype == _torch.bool:
        x = x.to(_INT)          # first maximal index like TF: torch.argmax returns the first occurrence on CPU
    return _torch.argmax(x, dim=axis)


def cumsum(x, axis=0):
    return _torch.cumsum(_t(x), dim=axis)


def _first_occurrence_unique(x):
    a = _t(x).numpy()
    _, first, inv, counts = _np.unique(a, return_index=True, return_inverse=True, return_counts=True)
    order = _np.argsort(first)                 # unique values in order of first occurrence
    rank_of = _np.empty_like(order)
    rank_of[order] = _np.arange(len(order))
    return a[_np.sort(first)], rank_of[inv], counts[order]


def unique(x, out_idx=None):
    y, idx, _ = _first_occurrence_unique(x)
    return _t(y, _t(x).dtype), _t(idx, _INT)


def unique_with_counts(x, out_idx=None):
    y, idx, c = _first_occurrence_unique(x)
    return _t(y, _t(x).dtype), _t(idx, _INT), _t(c, _INT)


# ---------------------------------------------------------------------------------------------- reductions / math
def _axis(axis):
    if axis is None:
        return None
    if isinstance(axis, _torch.Tensor):
        return int(axis)
    return tuple(axis) if isinstance(axis, (list, tuple)) else int(axis)


def reduce_sum(x, axis=None, keepdims=False):
    x = _t(x)
    return x.sum() if axis is None else x.sum(dim=_axis(axis), keepdim=keepdims)


def reduce_mean(x, axis=None, keepdims=False):
    x = _t(x)
    return x.mean() if axis is None else x.mean(dim=_axis(axis), keepdim=keepdims)


def reduce_max(x, axis=None, keepdims=False):
    x = _t(x)
    return x.max() if axis is None else _torch.amax(x, dim=_axis(axis), keepdim=keepdims)


def reduce_min(x, axis=None, keepdims=False):
    x = _t(x)
    return x.min() if axis is None else _torch.amin(x, dim=_axis(axis), keepdim=keepdims)


def reduce_any(x, axis=None):
    x = _t(x)
    return x.any() if axis is None else x.any(dim=_axis(axis))


def reduce_all(x, axis=None):
    x = _t(x)
    return x.all() if axis is None else x.all(dim=_axis(axis))


def reduce_logsumexp(x, axis=None, keepdims=False):
    x = _t(x)
    if axis is None:
        return _torch.logsumexp(x.reshape(-1), dim=0)
    return _torch.logsumexp(x, dim=_axis(axis), keepdim=keepdims)


def exp(x):
    return _torch.exp(_t(x))


def square(x):
    return _torch.square(_t(x))


def abs(x):  # noqa: A001
    return _torch.abs(_t(x))


def floor(x):
    return _torch.floor(_t(x))


def maximum(x, y):
    x, y = _t(x), _t(y)
    dt = _torch.promote_types(x.dtype, y.dtype)
    return _torch.maximum(x.to(dt), y.to(dt))


def minimum(x, y):
    x, y = _t(x), _t(y)
    dt = _torch.promote_types(x.dtype, y.dtype)
    return _torch.minimum(x.to(dt), y.to(dt))


def norm(x, ord="euclidean", axis=None, keepdims=False):
    x = _t(x)
    return _torch.linalg.vector_norm(x) if axis is None else _torch.linalg.vector_norm(x, dim=_axis(axis), keepdim=keepdims)


def assert_equal(x, y, message=None, **_):
    if not builtins.bool(_torch.all(_t(x) == _t(y))):
        raise AssertionError(message or "tf.assert_equal failed")


def print(*args, **kwargs):  # noqa: A001
    pass


class _Math:
    log = staticmethod(lambda x: _torch.log(_t(x)))
    exp = staticmethod(exp)
    sqrt = staticmethod(lambda x: _torch.sqrt(_t(x)))
    square = staticmethod(square)
    abs = staticmethod(abs)
    sign = staticmethod(lambda x: _torch.sign(_t(x)))
    floor = staticmethod(floor)
    maximum = staticmethod(maximum)
    minimum = staticmethod(minimum)
    is_nan = staticmethod(lambda x: _torch.isnan(_t(x)))
    is_inf = staticmethod(lambda x: _torch.isinf(_t(x)))
    is_finite = staticmethod(lambda x: _torch.isfinite(_t(x)))
    pow = staticmethod(lambda x, y: _torch.pow(_t(x), _t(y) if isinstance(y, _torch.Tensor) else y))
    sin = staticmethod(lambda x: _torch.sin(_t(x)))
    cos = staticmethod(lambda x: _torch.cos(_t(x)))
    reduce_logsumexp = staticmethod(reduce_logsumexp)
    reduce_sum = staticmethod(reduce_sum)
    reduce_max = staticmethod(reduce_max)
    cumsum = staticmethod(cumsum)
    argmax = staticmethod(argmax)


math = _Math()


# ---------------------------------------------------------------------------------------------- linear algebra
def _cholesky(a):
    a = _t(a)
    L, info = _torch.linalg.cholesky_ex(a)
    L = _torch.tril(L)
    bad = info > 0
    if builtins.bool(bad.any()):                 # TF (CPU, >= 2.5) returns NaNs for a matrix that is not positive definite
        L = _torch.where(bad.reshape(bad.shape + (1, 1)) if a.dim() > 2 else bad, _torch.full_like(L, float("nan")), L)
    return L


class _Linalg:
    cholesky = staticmethod(_cholesky)
    inv = staticmethod(lambda a: _torch.linalg.inv(_t(a)))
    diag_part = staticmethod(lambda a: _torch.diagonal(_t(a), dim1=-2, dim2=-1))
    tensor_diag_part = staticmethod(lambda a: _torch.diagonal(_t(a), dim1=-2, dim2=-1))
    diag = staticmethod(lambda v: _torch.diag_embed(_t(v)))
    matvec = staticmethod(lambda a, b, **_: (_t(a) @ _t(b).unsqueeze(-1)).squeeze(-1))
    matmul = staticmethod(lambda a, b, **_: _t(a) @ _t(b))
    det = staticmethod(lambda a: _torch.linalg.det(_t(a)))
    logdet = staticmethod(lambda a: _torch.logdet(_t(a)))
    eye = staticmethod(eye)
    norm = staticmethod(norm)

    @staticmethod
    def triangular_solve(matrix, rhs, lower=True, adjoint=False):
        m, r = _t(matrix), _t(rhs)
        if adjoint:
            m = m.transpose(-1, -2)
            lower = not lower
        return _torch.linalg.solve_triangular(m, r, upper=not lower)

    @staticmethod
    def cholesky_solve(chol, rhs):
        return _torch.cholesky_solve(_t(rhs), _t(chol), upper=False)

    @staticmethod
    def solve(matrix, rhs, adjoint=False):
        m = _t(matrix)
        return _torch.linalg.solve(m.transpose(-1, -2) if adjoint else m, _t(rhs))


class _LinearOperatorLowerTriangular:
    """tf.linalg.LinearOperatorLowerTriangular: only the lower triangle of `tril` is used."""
    def __init__(self, tril, **_):
        self.tril = _torch.tril(_t(tril))


_Linalg.LinearOperatorLowerTriangular = _LinearOperatorLowerTriangular
linalg = _Linalg()
matmul = _Linalg.matmul


# ---------------------------------------------------------------------------------------------- random (hooked)
class _Random:
    """normal_hook(shape) / uniform_hook(shape) / shuffle_hook(n) are set by the generator script."""
    normal_hook = None
    uniform_hook = None
    shuffle_hook = None

    def normal(self, shape, mean=0.0, stddev=1.0, dtype=float32, seed=None):
        if self.normal_hook is None:
            raise RuntimeError("tf.random.normal: no noise hook installed")
        return _t(self.normal_hook(_shape(shape)), dtype) * stddev + mean

    def uniform(self, shape, minval=0.0, maxval=1.0, dtype=float32, seed=None):
        if self.uniform_hook is None:
            raise RuntimeError("tf.random.uniform: no hook installed")
        return _t(self.uniform_hook(_shape(shape)), dtype) * (maxval - minval) + minval

    def shuffle(self, value, seed=None):
        if self.shuffle_hook is None:
            raise RuntimeError("tf.random.shuffle: no hook installed")
        v = _t(value)
        return v[_t(self.shuffle_hook(v.shape[0])).long()]

    def set_seed(self, seed):
        pass


random = _Random()
