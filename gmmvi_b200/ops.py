"""Thin torch-tensor wrappers over the C ABI (include/gmmvi_b200.h).

torch owns every buffer (CUDA, contiguous, fp32 / int32); the kernels run on torch's current stream.
No op here has a CPU or PyTorch fallback: a non-CUDA tensor is an error.
"""
from __future__ import annotations

from typing import Optional, Tuple

import os
import weakref

import torch

from . import _lib

LAUNCHES = 0   # number of C-ABI calls issued
KERNELS = 0    # number of CUDA kernels those calls launched (bench.py's "gpu_launches")

# kernels launched per C-ABI call (memsets are not counted)
_KERNELS_PER_CALL = {
    "gvi_stein_full_f32": 12,         # absmax x2, rowmax x2, transposes x2, stein_tc, gsum, split x2 + bgemm(P M), finalize
}


def kernel_launches() -> int:
    return KERNELS


USE_TENSOR_CORES = os.environ.get("GMMVI_B200_TC", "1") != "0"
_SPLIT_CACHE = []       # [(key, linv, hi, lo)]


# which tensor-core log-density kernel: "h16" (2 x fp16 split, resident factor), "tf32" (3xTF32, streamed)
TC_KIND = os.environ.get("GMMVI_B200_TC_KIND", "h16")
H16_MIN_DIM = 32        # below this the SIMT kernel wins (the fp16 operand is padded to 64 columns)


def logdens_kernel_kind(D: int) -> str:
    if USE_TENSOR_CORES and TC_KIND == "h16" and D >= H16_MIN_DIM and _lib.lib().gvi_logdens_full_h16_supported(int(D)):
        return "h16"
    if USE_TENSOR_CORES and _lib.lib().gvi_logdens_full_tc_supported(int(D)):
        return "tf32"
    return "simt"


def logdens_kernel_name(D: int = 256) -> str:
    h16 = ("gvi::h16::h16t::logdens_h16t_kernel (tcgen05 kind::f16, 2 x fp16 split, A operand in TMEM via tcgen05.st, "
           "resident Linv in shared memory by TMA)" if D > 64 and os.environ.get("GMMVI_B200_H16_A") != "smem" else
           "gvi::h16::logdens_h16_kernel (tcgen05 kind::f16, 2 x fp16 split, resident Linv, TMA + TMEM)")
    return {"h16": h16,
            "tf32": "gvi::tc::tc_logdens_kernel (tcgen05 kind::tf32, 3xTF32 split, TMA + TMEM)",
            "simt": "gvi::logdens_full_kernel (SIMT fp32 tile engine)"}[logdens_kernel_kind(int(D))]


_H16_CACHE = []         # [(key, linv, hi, lo, tmax)]
_ABSMAX_CACHE = []      # [(key, tensor, out)]


# Split operands of STATIC parameter buffers (a CUDA-graph runner's, optimization/graphed.py):
# {source address: [hi, lo, tmax, valid, C-ABI call, weak reference to the source tensor]}.  The split of such a buffer is written into its registered operand
# buffers, and while `valid` is set (the operands were produced from the buffer's current content) it is not recomputed;
# whoever rewrites the buffer calls invalidate_split().
_SPLIT_STATIC = {}
_SPLIT_CALL = {"lower": "gvi_split_h16_f32", "full": "gvi_split_h16_full_f32"}


def split_registered(src):
    """The registry entry of exactly this tensor object (an address can be recycled by the allocator once a graph runner
    and its buffers are gone: a stale entry must never be served to the new owner of the address)."""
    e = _SPLIT_STATIC.get(src.data_ptr())
    if e is not None and e[5]() is not src:
        del _SPLIT_STATIC[src.data_ptr()]
        return None
    return e


def register_split_buffers(src, kind="lower", valid=False):
    """Allocate operand buffers for the static tensor `src` [K, D, D] (kind: "lower" = inverse factors, "full" =
    precisions); no-op when the tensor-core kernels do not take this dimension."""
    K, D, _ = src.shape
    lib = _lib.lib()
    ok = lib.gvi_logdens_full_h16_supported(int(D)) if kind == "lower" else lib.gvi_mixture_grad_full_h16_supported(int(D))
    if not (USE_TENSOR_CORES and ok):
        return None
    for ptr in [ptr for ptr, old in _SPLIT_STATIC.items() if old[5]() is None]:      # buffers of runners that are gone
        del _SPLIT_STATIC[ptr]
    Dp = lib.gvi_h16_padded_dim(D)
    hi = torch.empty((K, Dp, Dp), device=src.device, dtype=torch.float16)
    e = [hi, torch.empty_like(hi), torch.empty(K, device=src.device, dtype=torch.float32), bool(valid), _SPLIT_CALL[kind],
         weakref.ref(src)]
    _SPLIT_STATIC[src.data_ptr()] = e
    return e


def invalidate_split(src):
    """`src` was rewritten in place (raw-pointer kernel / collective: torch's version counter does not see it): its
    registered split operands are stale, and so is anything memoised on its address."""
    e = split_registered(src)
    if e is not None:
        e[3] = False
        clear_memo()


def clear_split_registry():
    _SPLIT_STATIC.clear()


def _split_static(src):
    e = split_registered(src)
    if e is None:
        return None
    K, D, _ = src.shape
    if e[0].shape[0] != K:
        return None
    if not e[3]:
        _call(e[4], src.data_ptr(), K, D, e[0].data_ptr(), e[1].data_ptr(), e[2].data_ptr(), _stream())
        e[3] = True
    return e[0], e[1], e[2]


def split_static_now(src):
    """Compute the registered operands of `src` if they are not valid (eagerly, before a capture)."""
    _split_static(src)


def split_h16(linv):
    """Zero-padded, power-of-two scaled fp16 (hi, lo) copies of the inverse Cholesky factors + tmax[K]."""
    st = _split_static(linv)
    if st is not None:
        return st
    key = (linv.data_ptr(), linv._version, tuple(linv.shape))
    for k_, _, hi, lo, tmax in _H16_CACHE:
        if k_ == key:
            return hi, lo, tmax
    K, D, _ = linv.shape
    Dp = _lib.lib().gvi_h16_padded_dim(D)
    hi = torch.empty((K, Dp, Dp), device=linv.device, dtype=torch.float16)
    lo = torch.empty_like(hi)
    tmax = torch.empty(K, device=linv.device, dtype=torch.float32)
    _call("gvi_split_h16_f32", linv.data_ptr(), K, D, hi.data_ptr(), lo.data_ptr(), tmax.data_ptr(), _stream())
    _H16_CACHE.insert(0, (key, linv, hi, lo, tmax))
    del _H16_CACHE[3:]
    return hi, lo, tmax


def group_absmax(t, group=1):
    """out[g] = max |t[r, c]| over the rows r in [g * group, (g + 1) * group) (cached per buffer version)."""
    key = (t.data_ptr(), t._version, tuple(t.shape), group)
    for k_, _, out in _ABSMAX_CACHE:
        if k_ == key:
            return out
    out = torch.empty((t.shape[0] + group - 1) // group, device=t.device, dtype=torch.float32)
    _call("gvi_group_absmax_f32", t.data_ptr(), t.shape[0], t.shape[1], group, out.data_ptr(), _stream())
    _ABSMAX_CACHE.insert(0, (key, t, out))
    del _ABSMAX_CACHE[6:]
    return out


def split_tf32(linv):
    """(hi, lo) TF32 split of the inverse Cholesky factors, cached per buffer."""
    key = (linv.data_ptr(), linv._version, tuple(linv.shape))
    for k_, _, hi, lo in _SPLIT_CACHE:
        if k_ == key:
            return hi, lo
    hi, lo = torch.empty_like(linv), torch.empty_like(linv)
    _call("gvi_split_tf32_f32", linv.data_ptr(), linv.numel(), hi.data_ptr(), lo.data_ptr(), _stream())
    _SPLIT_CACHE.insert(0, (key, linv, hi, lo))
    del _SPLIT_CACHE[3:]
    return hi, lo


def _chk(t: torch.Tensor, name: str, dtype=torch.float32) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor, got {type(t)}")
    if not t.is_cuda:
        raise _lib.GmmviLibraryError(f"{name}: gmmvi_b200 kernels need CUDA tensors (got {t.device}); there is no CPU path")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _call(name, *args, kernels=None):
    global LAUNCHES, KERNELS
    LAUNCHES += 1
    KERNELS += kernels if kernels is not None else _KERNELS_PER_CALL.get(name, 1)
    _lib.check(getattr(_lib.lib(), name)(*args), name)


# ------------------------------------------------------------------------------------------------
def _fits(t, shape):
    return t is not None and tuple(t.shape) == tuple(shape) and t.dtype == torch.float32 and t.is_contiguous()


def prepare_full(chol: torch.Tensor, want_prec: bool = True, out=None):
    """chol[K,D,D] -> (linv[K,D,D], prec[K,D,D] | None, cst[K], ok[K] int32).  `out` = (linv, prec, cst) buffers to
    write into (used when their shapes fit; a graph runner passes its static buffers)."""
    chol = _chk(chol, "chol")
    K, D, _ = chol.shape
    o = out if out is not None else (None, None, None)
    if out is not None:
        clear_memo()        # results written into caller-owned buffers: the address / version keys of the memo do not see it
    linv = o[0] if _fits(o[0], chol.shape) else torch.empty_like(chol)
    prec = (o[1] if _fits(o[1], chol.shape) else torch.empty_like(chol)) if want_prec else None
    cst = o[2] if _fits(o[2], (K,)) else torch.empty(K, device=chol.device, dtype=torch.float32)
    ok = torch.empty(K, device=chol.device, dtype=torch.int32)
    nbytes = _lib.lib().gvi_prepare_full_workspace(K, D)
    ws = torch.empty(max(nbytes, 8) // 8, device=chol.device, dtype=torch.float64)
    _call("gvi_prepare_full_f32", chol.data_ptr(), K, D, linv.data_ptr(), _ptr(prec), cst.data_ptr(), ok.data_ptr(),
          ws.data_ptr(), nbytes, _stream())
    return linv, prec, cst, ok


_LOGDENS_MEMO = []      # [(key, keep_alive_tensors, lq)] newest first, at most 2 entries


def _memo_key(*tensors):
    return tuple((t.data_ptr(), t._version, tuple(t.shape)) for t in tensors)


def logdens_full(X, means, linv, cst, out=None, memo=True, tensor_cores=None):
    """lq[K,N].  Results are memoised on (buffer address, torch version counter, shape) of the four operands:
    inside one iteration the background density of the sample database and the Stein estimator evaluate the same
    components on the same samples (no-reuse configuration), and the second request is served from the first."""
    X, means, linv, cst = _chk(X, "X"), _chk(means, "means"), _chk(linv, "linv"), _chk(cst, "cst")
    N, D = X.shape
    K = means.shape[0]
    key = _memo_key(X, means, linv, cst) if (memo and out is None) else None
    if key is not None:
        for k_, _, lq_ in _LOGDENS_MEMO:
            if k_ == key:
                return lq_
    lq = out if out is not None else torch.empty((K, N), device=X.device, dtype=torch.float32)
    if tensor_cores is None:
        kind = logdens_kernel_kind(D)
    elif tensor_cores is False:
        kind = "simt"
    else:
        kind = tensor_cores if isinstance(tensor_cores, str) else logdens_kernel_kind(D)
        if kind == "tf32" and not _lib.lib().gvi_logdens_full_tc_supported(D):
            kind = "simt"
        if kind == "h16" and not _lib.lib().gvi_logdens_full_h16_supported(D):
            kind = "simt"
    if N == 0 or K == 0:
        kind = "simt"
    if kind == "h16":
        hi, lo, tmax = split_h16(linv)
        _call("gvi_logdens_full_h16_f32", X.data_ptr(), group_absmax(X, 128).data_ptr(), N, D, means.data_ptr(),
              group_absmax(means, 1).data_ptr(), hi.data_ptr(), lo.data_ptr(), tmax.data_ptr(), cst.data_ptr(), K,
              lq.data_ptr(), _stream())
    elif kind == "tf32":
        hi, lo = split_tf32(linv)
        _call("gvi_logdens_full_tc_f32", X.data_ptr(), N, D, means.data_ptr(), hi.data_ptr(), lo.data_ptr(),
              cst.data_ptr(), K, lq.data_ptr(), _stream())
    else:
        _call("gvi_logdens_full_f32", X.data_ptr(), N, D, means.data_ptr(), linv.data_ptr(), cst.data_ptr(), K,
              lq.data_ptr(), _stream())
    if key is not None:
        _LOGDENS_MEMO.insert(0, (key, (X, means, linv, cst), lq))
        del _LOGDENS_MEMO[2:]
    return lq


def clear_memo():
    del _LOGDENS_MEMO[:]


def clear_caches():
    """Drop the log-density memo and every derived-operand cache.  The caches are keyed on (address, torch version
    counter, shape); a CUDA-graph replay rewrites its static buffers without touching the version counters, so the
    graph runner clears them after every replay (optimization/graphed.py)."""
    del _LOGDENS_MEMO[:]
    del _H16_CACHE[:]
    del _SPLIT_CACHE[:]
    del _P16_CACHE[:]
    del _ABSMAX_CACHE[:]


def logdens_diag(X, means, stds):
    X, means, stds = _chk(X, "X"), _chk(means, "means"), _chk(stds, "stds")
    N, D = X.shape
    K = means.shape[0]
    lq = torch.empty((K, N), device=X.device, dtype=torch.float32)
    _call("gvi_logdens_diag_f32", X.data_ptr(), N, D, means.data_ptr(), stds.data_ptr(), K, lq.data_ptr(), _stream())
    return lq


def mixture_lse(lq, logw):
    lq, logw = _chk(lq, "lq"), _chk(logw, "logw")
    K, N = lq.shape
    out = torch.empty(N, device=lq.device, dtype=torch.float32)
    _call("gvi_mixture_lse_f32", lq.data_ptr(), logw.data_ptr(), K, N, out.data_ptr(), _stream())
    return out


_P16_CACHE = []         # [(key, prec, hi, lo, tmax)]
TC_MIXGRAD = os.environ.get("GMMVI_B200_TC_MIXGRAD", "1") != "0"


def split_h16_full(prec):
    """Zero-padded, power-of-two scaled fp16 (hi, lo) copies of full matrices [K, D, D] (the precisions) + tmax[K]."""
    st = _split_static(prec)
    if st is not None:
        return st
    key = (prec.data_ptr(), prec._version, tuple(prec.shape))
    for k_, _, hi, lo, tmax in _P16_CACHE:
        if k_ == key:
            return hi, lo, tmax
    K, D, _ = prec.shape
    Dp = _lib.lib().gvi_h16_padded_dim(D)
    hi = torch.empty((K, Dp, Dp), device=prec.device, dtype=torch.float16)
    lo = torch.empty_like(hi)
    tmax = torch.empty(K, device=prec.device, dtype=torch.float32)
    _call("gvi_split_h16_full_f32", prec.data_ptr(), K, D, hi.data_ptr(), lo.data_ptr(), tmax.data_ptr(), _stream())
    _P16_CACHE.insert(0, (key, prec, hi, lo, tmax))
    del _P16_CACHE[3:]
    return hi, lo, tmax


def mixture_grad_full(X, means, prec, lq, logw, logq, tensor_cores=None):
    X, means, prec = _chk(X, "X"), _chk(means, "means"), _chk(prec, "prec")
    lq, logw, logq = _chk(lq, "lq"), _chk(logw, "logw"), _chk(logq, "logq")
    N, D = X.shape
    K = means.shape[0]
    grad = torch.empty_like(X)
    nbytes = _lib.lib().gvi_mixture_grad_full_workspace(N, K)
    ws = torch.empty(max(nbytes, 4) // 4, device=X.device, dtype=torch.int32)
    use_tc = (USE_TENSOR_CORES and TC_MIXGRAD) if tensor_cores is None else bool(tensor_cores)
    if use_tc and N > 0 and K > 0 and _lib.lib().gvi_mixture_grad_full_h16_supported(int(D)):
        hi, lo, tmaxp = split_h16_full(prec)
        _call("gvi_mixture_grad_full_h16_f32", X.data_ptr(), group_absmax(X, 128).data_ptr(), N, D, means.data_ptr(),
              group_absmax(means, 1).data_ptr(), hi.data_ptr(), lo.data_ptr(), tmaxp.data_ptr(), lq.data_ptr(),
              logw.data_ptr(), logq.data_ptr(), K, grad.data_ptr(), ws.data_ptr(), nbytes, _stream(), kernels=2)
        return grad
    _call("gvi_mixture_grad_full_f32", X.data_ptr(), N, D, means.data_ptr(), prec.data_ptr(), lq.data_ptr(),
          logw.data_ptr(), logq.data_ptr(), K, grad.data_ptr(), ws.data_ptr(), nbytes, _stream(), kernels=2)
    return grad


def mixture_grad_diag(X, means, stds, lq, logw, logq):
    X, means, stds = _chk(X, "X"), _chk(means, "means"), _chk(stds, "stds")
    lq, logw, logq = _chk(lq, "lq"), _chk(logw, "logw"), _chk(logq, "logq")
    N, D = X.shape
    K = means.shape[0]
    grad = torch.empty_like(X)
    _call("gvi_mixture_grad_diag_f32", X.data_ptr(), N, D, means.data_ptr(), stds.data_ptr(), lq.data_ptr(),
          logw.data_ptr(), logq.data_ptr(), K, grad.data_ptr(), _stream())
    return grad


def importance_weights(lq, bg, rel_map=None, self_normalized=True, rho=None, want_W=False, want_dot=False,
                       want_ess=False, want_active=False):
    """Returns dict with the requested of W[K,N], dot[K], ess[K], active[K, ceil(N/128)] (uint8).
    self_normalized: True / 1 = softmax over the samples, normalised twice like the reference; False / 0 = the Stein
    convention exp(lw) / N (ng_estimator.py:155); 2 = plain exp(lw), what MORE hands to its regression (:356)."""
    lq = _chk(lq, "lq")
    K, N = lq.shape
    bg = _chk(bg, "bg") if bg is not None else None
    rel_map = _chk(rel_map, "rel_map", torch.int32) if rel_map is not None else None
    rho = _chk(rho, "rho") if rho is not None else None
    dev = lq.device
    W = torch.empty((K, N), device=dev, dtype=torch.float32) if want_W else None
    dot = torch.empty(K, device=dev, dtype=torch.float32) if want_dot else None
    ess = torch.empty(K, device=dev, dtype=torch.float32) if want_ess else None
    active = torch.empty((K, (N + 127) // 128), device=dev, dtype=torch.uint8) if want_active else None
    _call("gvi_importance_weights_f32", lq.data_ptr(), _ptr(bg), _ptr(rel_map), K, N, int(self_normalized),
          _ptr(rho), _ptr(W), _ptr(dot), _ptr(ess), _ptr(active), _stream())
    return dict(W=W, dot=dot, ess=ess, active=active)


def importance_weights_sharded(lq, bg, shard, self_normalized=True, rho=None, want_W=False, want_dot=False,
                               want_active=False, n_total=None):
    """Sample-sharded version of `importance_weights`: lq[K, N_local], bg[N_local]; the row maxima and sums of
    exponentials are all-reduced over `shard` (gmmvi_b200.distributed.ShardContext) so that the weights are
    normalised over ALL samples of the iteration.  `dot` is the global sum."""
    lq, bg = _chk(lq, "lq"), _chk(bg, "bg")
    K, N = lq.shape
    dev = lq.device
    rho = _chk(rho, "rho") if rho is not None else None
    f = lambda: torch.empty(K, device=dev, dtype=torch.float32)
    m = f()
    _call("gvi_row_max_f32", lq.data_ptr(), bg.data_ptr(), K, N, m.data_ptr(), _stream())
    W = torch.empty((K, N), device=dev, dtype=torch.float32) if want_W else None
    dot = f() if want_dot else None
    active = torch.empty((K, (N + 127) // 128), device=dev, dtype=torch.uint8) if want_active else None
    if self_normalized:
        # One all-gather of the per-rank (max, sum of exponentials) pairs replaces the MAX and the first SUM
        # all-reduce: every rank merges the pairs itself (same bits everywhere).
        m_safe = torch.where(torch.isfinite(m), m, torch.zeros_like(m)).contiguous()    # rows without a finite entry
        s = f()
        _call("gvi_row_sumexp_f32", lq.data_ptr(), bg.data_ptr(), K, N, m_safe.data_ptr(), s.data_ptr(), _stream())
        pairs = shard.all_gather_rows(torch.stack([m, m_safe, s]).unsqueeze(0).contiguous(), shard.world)  # [P, 3, K]
        m = pairs[:, 0].max(dim=0).values
        m = torch.where(torch.isfinite(m), m, torch.zeros_like(m)).contiguous()
        term = torch.where(pairs[:, 2] > 0, pairs[:, 2] * torch.exp(pairs[:, 1] - m), torch.zeros_like(pairs[:, 2]))
        lse = (m + torch.log(term.sum(dim=0))).contiguous()
        s2 = f()
        _call("gvi_row_sumexp_f32", lq.data_ptr(), bg.data_ptr(), K, N, lse.data_ptr(), s2.data_ptr(), _stream())
        if want_W or not want_dot:
            shard.all_reduce_sum_(s2)
            scale = (1.0 / s2).contiguous()
            _call("gvi_importance_weights_ext_f32", lq.data_ptr(), bg.data_ptr(), K, N, lse.data_ptr(),
                  scale.data_ptr(), m.data_ptr(), _ptr(rho), _ptr(W), _ptr(dot), _ptr(active), _stream())
            if dot is not None:
                shard.all_reduce_sum_(dot)
        else:
            # only the expected log-ratios are wanted: the second normaliser (the reference normalises twice) and
            # the un-normalised dot products travel in ONE all-reduce
            _call("gvi_importance_weights_ext_f32", lq.data_ptr(), bg.data_ptr(), K, N, lse.data_ptr(), None,
                  m.data_ptr(), _ptr(rho), None, dot.data_ptr(), _ptr(active), _stream())
            both = torch.stack([s2, dot]).contiguous()
            shard.all_reduce_sum_(both)
            dot = (both[1] / both[0]).contiguous()
    else:
        import math
        shard.all_reduce_max_(m)
        m = torch.where(torch.isfinite(m), m, torch.zeros_like(m)).contiguous()
        lse = torch.full((K,), math.log(float(n_total)), device=dev, dtype=torch.float32)
        _call("gvi_importance_weights_ext_f32", lq.data_ptr(), bg.data_ptr(), K, N, lse.data_ptr(), None,
              m.data_ptr(), _ptr(rho), _ptr(W), _ptr(dot), _ptr(active), _stream())
        if dot is not None:
            shard.all_reduce_sum_(dot)
    return dict(W=W, dot=dot, ess=None, active=active)


def stein_full(X, means, prec, W, active, G, symmetrize=True):
    X, means, prec, W, G = _chk(X, "X"), _chk(means, "means"), _chk(prec, "prec"), _chk(W, "W"), _chk(G, "G")
    if active is not None:
        active = _chk(active, "active", torch.uint8)
    N, D = X.shape
    K = means.shape[0]
    Hneg = torch.empty((K, D, D), device=X.device, dtype=torch.float32)
    gneg = torch.empty((K, D), device=X.device, dtype=torch.float32)
    nbytes = _lib.lib().gvi_stein_full_workspace(N, K, D)
    ws = torch.empty(max(nbytes, 4) // 4, device=X.device, dtype=torch.float32)
    _call("gvi_stein_full_f32", X.data_ptr(), N, D, means.data_ptr(), prec.data_ptr(), W.data_ptr(), _ptr(active),
          G.data_ptr(), K, int(bool(symmetrize)), Hneg.data_ptr(), gneg.data_ptr(), ws.data_ptr(), nbytes, _stream())
    return Hneg, gneg


def stein_stats_full(X, means, W, active, G):
    """Raw Stein statistics of this rank's samples -> (M[K,D,D] = sum_n w (x - mu) g^T, gneg[K,D] = -sum_n w g)."""
    X, means, W, G = _chk(X, "X"), _chk(means, "means"), _chk(W, "W"), _chk(G, "G")
    if active is not None:
        active = _chk(active, "active", torch.uint8)
    N, D = X.shape
    K = means.shape[0]
    M = torch.empty((K, D, D), device=X.device, dtype=torch.float32)
    gneg = torch.empty((K, D), device=X.device, dtype=torch.float32)
    nbytes = _lib.lib().gvi_stein_stats_full_workspace(N, K, D)
    ws = torch.empty(max(nbytes, 4) // 4, device=X.device, dtype=torch.float32)
    _call("gvi_stein_stats_full_f32", X.data_ptr(), N, D, means.data_ptr(), W.data_ptr(), _ptr(active), G.data_ptr(), K,
          M.data_ptr(), gneg.data_ptr(), ws.data_ptr(), nbytes, _stream(), kernels=8)
    return M, gneg


def stein_finalize_full(prec, M, symmetrize=True):
    """Hneg[k] = -sym(P_k M_k) (or the un-symmetrised form of the standard-IW branch) for the given components."""
    prec, M = _chk(prec, "prec"), _chk(M, "M")
    K, D, _ = M.shape
    Hneg = torch.empty_like(M)
    nbytes = _lib.lib().gvi_stein_finalize_full_workspace(K, D)
    ws = torch.empty(max(nbytes, 4) // 4, device=M.device, dtype=torch.float32)
    _call("gvi_stein_finalize_full_f32", prec.data_ptr(), M.data_ptr(), K, D, int(bool(symmetrize)), Hneg.data_ptr(),
          ws.data_ptr(), nbytes, _stream(), kernels=4)
    return Hneg


def stein_diag(X, means, stds, W, G):
    X, means, stds, W, G = _chk(X, "X"), _chk(means, "means"), _chk(stds, "stds"), _chk(W, "W"), _chk(G, "G")
    N, D = X.shape
    K = means.shape[0]
    Hneg = torch.empty((K, D), device=X.device, dtype=torch.float32)
    gneg = torch.empty((K, D), device=X.device, dtype=torch.float32)
    nbytes = _lib.lib().gvi_stein_diag_workspace(N, K, D)
    ws = torch.empty(max(nbytes, 4) // 4, device=X.device, dtype=torch.float32)
    _call("gvi_stein_diag_f32", X.data_ptr(), N, D, means.data_ptr(), stds.data_ptr(), W.data_ptr(), G.data_ptr(), K,
          Hneg.data_ptr(), gneg.data_ptr(), ws.data_ptr(), nbytes, _stream(), kernels=4)
    return Hneg, gneg


_MORE_BUDGET = None     # last default workspace budget (not queried again while a CUDA graph is being captured)


def more_fit(regularizers, samples, rewards, weights, means, linv, memory_budget_bytes=None):
    """MORE: per-component weighted quadratic regression -> (reward_quad[K,D,D], reward_lin[K,D], ok[K] int32).
    Components are processed in chunks that fit `memory_budget_bytes` of workspace (C3: 0.2 GB per component); the
    default is half of the device memory that is free right now, at most 32 GiB."""
    global _MORE_BUDGET
    if memory_budget_bytes is None:
        if _MORE_BUDGET is None or not torch.cuda.is_current_stream_capturing():
            free, _ = torch.cuda.mem_get_info(samples.device)
            _MORE_BUDGET = min(32 << 30, free // 2)
        memory_budget_bytes = _MORE_BUDGET
    X, y, W = _chk(samples, "samples"), _chk(rewards, "rewards"), _chk(weights, "weights")
    means, linv, l2 = _chk(means, "means"), _chk(linv, "linv"), _chk(regularizers, "regularizers")
    N, D = X.shape
    K = means.shape[0]
    dev = X.device
    per = _lib.lib().gvi_more_workspace(1, N, D)
    chunk = int(max(1, min(K, memory_budget_bytes // max(per, 1))))
    nbytes = _lib.lib().gvi_more_workspace(chunk, N, D)
    ws = torch.empty(max(nbytes, 4) // 4, device=dev, dtype=torch.float32)
    quad = torch.empty((K, D, D), device=dev, dtype=torch.float32)
    lin = torch.empty((K, D), device=dev, dtype=torch.float32)
    ok = torch.ones(K, device=dev, dtype=torch.int32)
    F = D * (D + 1) // 2 + D + 1
    panels = (F + 127) // 128
    _call("gvi_more_fit_f32", X.data_ptr(), N, D, means.data_ptr(), linv.data_ptr(), W.data_ptr(), y.data_ptr(),
          l2.data_ptr(), K, chunk, quad.data_ptr(), lin.data_ptr(), ok.data_ptr(), ws.data_ptr(), nbytes, _stream(),
          kernels=((K + chunk - 1) // chunk) * (9 + 4 * panels + (2 if _lib.lib().gvi_more_tensor_cores() == 2 else 0)))
    return quad, lin, ok


def bgemm_h16(A, B, C, alpha=1.0, beta=0.0, kseg_kblocks=0, lower_only=False):
    """In place C[b] = alpha * A[b] B[b]^T + beta * C[b] in the 2 x fp16 split precision (gvi_tc_bgemm_h16_f32): A [b, M, K],
    B [b, N, K] with K % 8 == 0, one power-of-two scale per batch entry and operand; kseg_kblocks counts blocks of 64."""
    A, B = _chk(A, "A"), _chk(B, "B")
    if not C.is_contiguous():
        raise ValueError("bgemm_h16: C is updated in place and must be contiguous")
    C = _chk(C, "C")
    batch, M, Kd = A.shape
    N = B.shape[1]
    if tuple(C.shape) != (batch, M, N) or B.shape[2] != Kd:
        raise ValueError(f"bgemm_h16: shapes {tuple(A.shape)} {tuple(B.shape)} {tuple(C.shape)} do not match")
    nbytes = _lib.lib().gvi_tc_bgemm_h16_workspace(batch, M, N, Kd)
    ws = torch.empty(nbytes // 4 + 4, device=A.device, dtype=torch.float32)
    _call("gvi_tc_bgemm_h16_f32", batch, M, N, Kd, float(alpha), A.data_ptr(), B.data_ptr(), C.data_ptr(), float(beta),
          int(kseg_kblocks), int(bool(lower_only)), ws.data_ptr(), nbytes, _stream(), kernels=6)
    return C


def bgemm_ex(A, B, C, transA=False, transB=False, alpha=1.0, beta=0.0, kseg_kblocks=0, lower_only=False):
    """In place C[b] = alpha * op(A[b]) op(B[b]) + beta * C[b] on the tcgen05 tensor cores (3xTF32) with the reduction
    cut into round-to-nearest accumulated segments of `kseg_kblocks` x 32 and, with lower_only, only the tiles that
    touch the lower triangle written (gvi_tc_bgemm_ex_f32; the MORE normal equations and Cholesky trailing updates)."""
    A, B = _chk(A, "A"), _chk(B, "B")
    if not C.is_contiguous():
        raise ValueError("bgemm_ex: C is updated in place and must be contiguous")
    C = _chk(C, "C")
    batch = A.shape[0]
    M, Kd = (A.shape[2], A.shape[1]) if transA else (A.shape[1], A.shape[2])
    N = B.shape[1] if transB else B.shape[2]
    if tuple(C.shape) != (batch, M, N):
        raise ValueError(f"bgemm_ex: C has shape {tuple(C.shape)}, expected {(batch, M, N)}")
    nbytes = _lib.lib().gvi_tc_bgemm_workspace(batch, M, N, Kd)
    ws = torch.empty(nbytes // 4 + 4, device=A.device, dtype=torch.float32)
    _call("gvi_tc_bgemm_ex_f32", int(transA), int(transB), batch, M, N, Kd, float(alpha), A.data_ptr(), A.shape[2],
          A.shape[1] * A.shape[2], B.data_ptr(), B.shape[2], B.shape[1] * B.shape[2], C.data_ptr(), N, M * N,
          float(beta), int(kseg_kblocks), int(bool(lower_only)), ws.data_ptr(), nbytes, _stream(), kernels=3)
    return C


UPDATE_MODES = {"trust-region": 0, "direct": 1, "iBLR": 2}
last_update_evals = None      # int32[K]: KL evaluations per component of the most recent full-covariance update


def LAST_UPDATE_EVALS_PTR(K, dev):
    global last_update_evals
    last_update_evals = torch.zeros(K, device=dev, dtype=torch.int32)
    return last_update_evals.data_ptr()


def update_components(mode: str, diagonal: bool, means, chols, Hneg, gneg, stepsizes, last_etas=None,
                      num_updates=None, temperature: float = 1.0):
    """-> (new_means, new_chols, success[K] int32, etas[K], kls[K])."""
    m = UPDATE_MODES[mode]
    means, chols, Hneg, gneg = _chk(means, "means"), _chk(chols, "chols"), _chk(Hneg, "Hneg"), _chk(gneg, "gneg")
    stepsizes = _chk(stepsizes, "stepsizes")
    last_etas = _chk(last_etas, "last_etas") if last_etas is not None else None
    num_updates = _chk(num_updates, "num_updates") if num_updates is not None else None
    K, D = means.shape
    dev = means.device
    om, oc = torch.empty_like(means), torch.empty_like(chols)
    succ = torch.empty(K, device=dev, dtype=torch.int32)
    etas = torch.empty(K, device=dev, dtype=torch.float32)
    kls = torch.empty(K, device=dev, dtype=torch.float32)
    if diagonal:
        _call("gvi_update_diag_f32", m, means.data_ptr(), chols.data_ptr(), Hneg.data_ptr(), gneg.data_ptr(),
              stepsizes.data_ptr(), _ptr(last_etas), _ptr(num_updates), K, D, float(temperature), om.data_ptr(),
              oc.data_ptr(), succ.data_ptr(), etas.data_ptr(), kls.data_ptr(), _stream())
    else:
        nbytes = _lib.lib().gvi_update_full_workspace(K, D)
        ws = torch.empty(max(nbytes, 4) // 4, device=dev, dtype=torch.float32)
        _call("gvi_update_full_f32", m, means.data_ptr(), chols.data_ptr(), Hneg.data_ptr(), gneg.data_ptr(),
              stepsizes.data_ptr(), _ptr(last_etas), _ptr(num_updates), K, D, float(temperature), om.data_ptr(),
              oc.data_ptr(), succ.data_ptr(), etas.data_ptr(), kls.data_ptr(), LAST_UPDATE_EVALS_PTR(K, dev),
              ws.data_ptr(), nbytes, _stream(),
              kernels=6 if m == 2 else 5)     # mirror + 2 (3) bgemm + vectors + update
    return om, oc, succ, etas, kls


def update_components_general(mode: str, means, chols, prec, Hneg, gneg, stepsizes, num_updates=None):
    """Direct / iBLR update for a non-symmetric -E[H] (gvi_update_full_general_f32) -> (new_means, new_chols, success)."""
    m = UPDATE_MODES[mode]
    means, chols, prec, Hneg, gneg = (_chk(means, "means"), _chk(chols, "chols"), _chk(prec, "prec"), _chk(Hneg, "Hneg"),
                                      _chk(gneg, "gneg"))
    stepsizes = _chk(stepsizes, "stepsizes")
    num_updates = _chk(num_updates, "num_updates") if num_updates is not None else None
    K, D = means.shape
    om, oc = torch.empty_like(means), torch.empty_like(chols)
    succ = torch.empty(K, device=means.device, dtype=torch.int32)
    nbytes = _lib.lib().gvi_update_full_general_workspace(K, D)
    ws = torch.empty(max(nbytes, 4) // 4, device=means.device, dtype=torch.float32)
    _call("gvi_update_full_general_f32", m, means.data_ptr(), chols.data_ptr(), prec.data_ptr(), Hneg.data_ptr(),
          gneg.data_ptr(), stepsizes.data_ptr(), _ptr(num_updates), K, D, om.data_ptr(), oc.data_ptr(), succ.data_ptr(),
          ws.data_ptr(), nbytes, _stream(), kernels=4 if m == 2 else 1)
    return om, oc, succ


def gauss_kernel_sum(X, Y, w):
    """sum_{i,j} exp(-sum_d w[d] (X[i,d] - Y[j,d])^2) -> 0-d float32 tensor (MMD evaluation, mmd.py:38-56)."""
    X, Y, w = _chk(X, "X"), _chk(Y, "Y"), _chk(w, "w")
    n1, D = X.shape
    n2 = Y.shape[0]
    npart = _lib.lib().gvi_gauss_kernel_sum_partials(n1, n2)
    partial = torch.zeros(max(npart, 1), device=X.device, dtype=torch.float64)
    _call("gvi_gauss_kernel_sum_f32", X.data_ptr(), n1, Y.data_ptr(), n2, D, w.data_ptr(), partial.data_ptr(), _stream())
    return partial.sum().to(torch.float32)


def cholesky(covs):
    """Lower Cholesky factors of covs[K,D,D] (fp64 arithmetic on the device, fp32 result; NaN factor when not positive
    definite) -> (L[K,D,D], ok[K] int32).  Construction-time only: models/full_cov_gmm.py:23, :67."""
    covs = _chk(covs, "covs")
    K, D, _ = covs.shape
    L = torch.empty_like(covs)
    ok = torch.empty(K, device=covs.device, dtype=torch.int32)
    nbytes = _lib.lib().gvi_cholesky_workspace(K, D)
    ws = torch.empty(max(nbytes, 8) // 8, device=covs.device, dtype=torch.float64)
    _call("gvi_cholesky_f32", covs.data_ptr(), K, D, L.data_ptr(), ok.data_ptr(), ws.data_ptr(), nbytes, _stream())
    return L, ok


def planar_robot(theta, prior_stds, goals, likelihood_std: float, want_grad: bool = True):
    """PlanarRobot.log_density (planar_robot.py:49-66) and its gradient -> (lnpdf[N], grad[N,D] | None)."""
    theta, prior_stds, goals = _chk(theta, "theta"), _chk(prior_stds, "prior_stds"), _chk(goals, "goals")
    N, D = theta.shape
    lnpdf = torch.empty(N, device=theta.device, dtype=torch.float32)
    grad = torch.empty_like(theta) if want_grad else None
    _call("gvi_planar_robot_f32", theta.data_ptr(), N, D, prior_stds.data_ptr(), None, goals.data_ptr(), goals.shape[0],
          float(likelihood_std), lnpdf.data_ptr(), _ptr(grad), _stream())
    return lnpdf, grad


def tridiag(B, h):
    """Householder tridiagonalisation of the symmetric matrices S[i][j] = B[k][min(i,j)][max(i,j)] (D <= 256):
    -> (d[K,D], e[K,D] with e[:, D-1] = 0, hp[K,D] = P^T h)."""
    B, h = _chk(B, "B"), _chk(h, "h")
    K, D = h.shape
    d, e, hp = (torch.empty((K, D), device=B.device, dtype=torch.float32) for _ in range(3))
    _call("gvi_tridiag_f32", B.data_ptr(), h.data_ptr(), K, D, d.data_ptr(), e.data_ptr(), hp.data_ptr(), _stream())
    return d, e, hp


def weight_update(trust_region: bool, logw, elr, stepsize, temperature: float = 1.0):
    """stepsize: device scalar tensor (or python float).  -> (new_log_weights (un-normalised result of the
    reference's search), info[2] = (kl, eta))."""
    logw, elr = _chk(logw, "logw"), _chk(elr, "elr")
    if not isinstance(stepsize, torch.Tensor):
        stepsize = torch.tensor([float(stepsize)], device=logw.device, dtype=torch.float32)
    stepsize = _chk(stepsize.reshape(1), "stepsize")
    K = logw.shape[0]
    out = torch.empty_like(logw)
    info = torch.empty(2, device=logw.device, dtype=torch.float32)
    _call("gvi_weight_update_f32", int(bool(trust_region)), logw.data_ptr(), elr.data_ptr(), K, stepsize.data_ptr(),
          float(temperature), out.data_ptr(), info.data_ptr(), _stream())
    return out, info


def fill_normal(rows: int, D: int, seed: int, subsequence: int = 0, row_offset: int = 0, device="cuda"):
    out = torch.empty((rows, D), device=device, dtype=torch.float32)
    if not out.is_cuda:
        raise _lib.GmmviLibraryError("fill_normal: needs a CUDA device")
    with torch.cuda.device(out.device):
        if isinstance(subsequence, tuple):          # (device counter, offset): rng device mode (CUDA-graph capture)
            counter, off = subsequence
            _call("gvi_fill_normal_dev_f32", out.data_ptr(), rows, D, seed & (2 ** 64 - 1), counter.data_ptr(),
                  int(off), row_offset, _stream())
        else:
            _call("gvi_fill_normal_f32", out.data_ptr(), rows, D, seed & (2 ** 64 - 1), subsequence & (2 ** 64 - 1),
                  row_offset, _stream())
    return out


def sample_components(diagonal: bool, eps, offsets, means, chols, max_rows_per_component: int):
    """eps[N,D] noise, offsets[K+1] int32 device prefix sums -> (X[N,D], mapping[N] int32)."""
    eps, means, chols = _chk(eps, "eps"), _chk(means, "means"), _chk(chols, "chols")
    offsets = _chk(offsets, "offsets", torch.int32)
    N, D = eps.shape
    K = means.shape[0]
    X = torch.empty_like(eps)
    mapping = torch.empty(N, device=eps.device, dtype=torch.int32)
    _call("gvi_sample_f32", int(bool(diagonal)), eps.data_ptr(), offsets.data_ptr(), means.data_ptr(),
          chols.data_ptr(), K, D, int(max_rows_per_component), X.data_ptr(), mapping.data_ptr(), _stream())
    return X, mapping


def bgemm(A, B, transA=False, transB=False, alpha=1.0, tensor_cores=None, out=None):
    """Batched C[b] = alpha * op(A[b]) op(B[b]) for 3-D tensors (or 2-D, batch 1).  tensor_cores=True forces the
    tcgen05 3xTF32 kernel, False the SIMT engine, None picks the tensor cores when the shape allows."""
    A, B = _chk(A, "A"), _chk(B, "B")
    squeeze = A.dim() == 2
    if squeeze:
        A, B = A.unsqueeze(0), B.unsqueeze(0)
    batch = A.shape[0]
    M, Kd = (A.shape[2], A.shape[1]) if transA else (A.shape[1], A.shape[2])
    N = B.shape[1] if transB else B.shape[2]
    if out is not None:
        clear_memo()
    Cc = out if (_fits(out, (batch, M, N)) and not squeeze) else torch.empty((batch, M, N), device=A.device, dtype=torch.float32)
    use_tc = tensor_cores
    if use_tc is None:
        use_tc = USE_TENSOR_CORES and M >= 64 and N >= 32 and bool(_lib.lib().gvi_tc_bgemm_supported(M, N, Kd))
    if use_tc:
        nbytes = _lib.lib().gvi_tc_bgemm_workspace(batch, M, N, Kd)
        ws = torch.empty(nbytes // 4 + 4, device=A.device, dtype=torch.float32)
        _call("gvi_tc_bgemm_f32", int(transA), int(transB), batch, M, N, Kd, float(alpha), A.data_ptr(), A.shape[2],
              A.shape[1] * A.shape[2], B.data_ptr(), B.shape[2], B.shape[1] * B.shape[2], Cc.data_ptr(), N, M * N,
              ws.data_ptr(), nbytes, _stream(), kernels=3)
    else:
        _call("gvi_bgemm_f32", int(transA), int(transB), batch, M, N, Kd, float(alpha), A.data_ptr(), A.shape[2],
              A.shape[1] * A.shape[2], B.data_ptr(), B.shape[2], B.shape[1] * B.shape[2], Cc.data_ptr(), N, M * N,
              _stream())
    return Cc[0] if squeeze else Cc
