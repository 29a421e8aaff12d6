"""SampleDB on device (mirror of optimization/sample_db.py:4-228)."""
from __future__ import annotations

import torch

from .. import ops


class _Growable:
    """Row-appendable device tensor: `view` is buf[:n]; appending copies only the new rows (capacity doubles when it
    runs out), where torch.cat re-copies -- and re-allocates -- the whole database every iteration (the reference's
    tf.concat does the same: SURVEY.md section 8, N2)."""

    def __init__(self, t: torch.Tensor):
        self.buf, self.n = t, int(t.shape[0])

    @property
    def view(self) -> torch.Tensor:
        return self.buf[:self.n]

    def append(self, x: torch.Tensor):
        m = int(x.shape[0])
        if self.n + m > self.buf.shape[0]:
            cap = max(2 * int(self.buf.shape[0]), self.n + m, 1024)
            nb = torch.empty((cap,) + tuple(self.buf.shape[1:]), device=self.buf.device, dtype=self.buf.dtype)
            nb[:self.n].copy_(self.buf[:self.n])
            self.buf = nb
        self.buf[self.n:self.n + m].copy_(x)
        self.n += m


def _stored(name):
    def get(self):
        return self._store[name].view

    def set_(self, value):
        self._store[name] = _Growable(value)
    return property(get, set_)


class SampleDB:
    # every stored array is a growable buffer behind a plain tensor attribute (reads see the filled rows)
    samples, means, chols, inv_chols = _stored("samples"), _stored("means"), _stored("chols"), _stored("inv_chols")
    consts, target_lnpdfs, target_grads, mapping = (_stored("consts"), _stored("target_lnpdfs"),
                                                    _stored("target_grads"), _stored("mapping"))

    def __init__(self, dim, diagonal_covariances, keep_samples, max_samples=None, device="cuda"):
        """optimization/sample_db.py:30-46."""
        self._dim = dim
        self._store = {}
        self.diagonal_covariances = diagonal_covariances
        self.keep_samples = keep_samples
        self.max_samples = max_samples
        self.device = torch.device(device)
        z = lambda *s: torch.zeros(s, device=self.device)
        cshape = (0, dim) if diagonal_covariances else (0, dim, dim)
        self.samples = z(0, dim)
        self.means = z(0, dim)
        self.chols = z(*cshape)
        self.inv_chols = z(*cshape)
        self.consts = z(0)                       # log-normalisers of the stored Gaussians (full cov)
        self.target_lnpdfs = z(0)
        self.target_grads = z(0, dim)
        self.mapping = torch.zeros(0, device=self.device, dtype=torch.int32)
        self.num_samples_written = 0
        self._last_batch = None                  # (#samples, #components) of the most recent add_samples call
        self.count_override = None               # global per-component counts when samples are sharded over GPUs

    @staticmethod
    def build_from_config(config, num_dimensions, device="cuda"):
        """optimization/sample_db.py:48-62."""
        return SampleDB(num_dimensions, config["model_initialization"]["use_diagonal_covs"],
                        config["use_sample_database"], config["max_database_size"], device=device)

    def remove_every_nth_sample(self, N: int):
        """optimization/sample_db.py:64-79."""
        self._last_batch = None
        self.samples = self.samples[::N].contiguous()
        self.target_lnpdfs = self.target_lnpdfs[::N].contiguous()
        self.target_grads = self.target_grads[::N].contiguous()
        mapping = self.mapping[::N]
        # tf.unique returns the values in FIRST-OCCURRENCE order (sample_db.py:75).  The mapping of the component-based
        # selector is non-decreasing (sorted order would do), the mixture-based selector stores draw-order indices
        # (models/gmm.py:155-163), so the order is rebuilt explicitly: rank the sorted unique values by the position of
        # their first occurrence.
        vals, inv = torch.unique(mapping, sorted=True, return_inverse=True)
        pos = torch.arange(mapping.shape[0], device=mapping.device)
        first = torch.full((vals.shape[0],), mapping.shape[0], device=mapping.device, dtype=pos.dtype)
        first.scatter_reduce_(0, inv, pos, reduce="amin")
        order = torch.argsort(first)                       # unique values in first-occurrence order
        rank = torch.empty_like(order)
        rank[order] = torch.arange(order.shape[0], device=order.device)
        self.mapping = rank[inv].to(torch.int32).contiguous()
        used = vals[order].long()
        self.means = self.means[used].contiguous()
        self.chols = self.chols[used].contiguous()
        self.inv_chols = self.inv_chols[used].contiguous()
        if not self.diagonal_covariances:
            self.consts = self.consts[used].contiguous()

    def _invert(self, chols, prepared=None):
        if self.diagonal_covariances:
            return 1.0 / chols, None
        if prepared is not None:
            return prepared[0], prepared[2]
        linv, _, cst, _ = ops.prepare_full(chols, want_prec=False)
        return linv, cst

    def add_samples(self, samples, means, chols, target_lnpdfs, target_grads, mapping, prepared=None):
        """optimization/sample_db.py:82-135.  `prepared` = model.prepared() lets the caller share the already
        inverted Cholesky factors (tf.linalg.inv(chols), :121,132) instead of recomputing them."""
        if self.max_samples is not None and samples.shape[0] + self.samples.shape[0] > self.max_samples:
            self.remove_every_nth_sample(2)
        self.num_samples_written += int(samples.shape[0])
        self._last_batch = (int(samples.shape[0]), int(means.shape[0]))
        inv, cst = self._invert(chols, prepared)
        mapping = mapping.to(torch.int32)
        if self.keep_samples:
            st = self._store
            st["mapping"].append(mapping + int(self.chols.shape[0]))
            st["means"].append(means)
            st["chols"].append(chols)
            st["inv_chols"].append(inv)
            if cst is not None:
                st["consts"].append(cst)
            st["samples"].append(samples)
            st["target_lnpdfs"].append(target_lnpdfs)
            st["target_grads"].append(target_grads)
        else:
            self.mapping, self.means, self.chols, self.inv_chols = mapping, means, chols, inv
            self.consts = cst if cst is not None else self.consts
            self.samples, self.target_lnpdfs, self.target_grads = samples, target_lnpdfs, target_grads

    def get_random_sample(self, N: int, permutation=None):
        """optimization/sample_db.py:137-152.  `permutation` (optional, a permutation of range(#samples)) replaces the
        device generator's shuffle: the parity tests inject the one the reference drew."""
        if permutation is None:
            idx = torch.randperm(self.samples.shape[0], device=self.device)[:N]
        else:
            idx = torch.as_tensor(permutation, dtype=torch.long, device=self.device)[:N]
        return self.samples[idx], self.target_lnpdfs[idx]

    def gaussian_log_pdf(self, mean, chol, inv_chol, x):
        """optimization/sample_db.py:154-162 (single stored Gaussian)."""
        if self.diagonal_covariances:
            return ops.logdens_diag(x, mean.reshape(1, -1).contiguous(), chol.reshape(1, -1).contiguous())[0]
        D = self._dim
        cst = (-0.5 * D * 1.8378770664093453 - torch.sum(torch.log(torch.diagonal(chol)))).reshape(1)
        return ops.logdens_full(x, mean.reshape(1, -1).contiguous(), inv_chol.reshape(1, D, D).contiguous(), cst)[0]

    def evaluate_background(self, weights, means, chols, inv_chols, samples, consts=None):
        """optimization/sample_db.py:164-192: log of the mixture the samples were drawn from.  The reference
        accumulates the components sequentially with pairwise logsumexp; here one batched log-density pass is
        followed by a single logsumexp over components (weights of 0 contribute -inf)."""
        if self.diagonal_covariances:
            lq = ops.logdens_diag(samples, means, chols)
        else:
            if consts is None:
                D = self._dim
                consts = -0.5 * D * 1.8378770664093453 - torch.sum(torch.log(torch.diagonal(chols, dim1=1, dim2=2)), 1)
            lq = ops.logdens_full(samples, means, inv_chols, consts.contiguous())
        return ops.mixture_lse(lq, torch.log(weights).contiguous()), lq

    def get_newest_samples(self, N):
        """optimization/sample_db.py:195-228 -> (bg[N'], samples, mapping, target_lnpdfs, target_grads)."""
        D, dev = self._dim, self.device
        S = int(self.samples.shape[0])
        N = int(N)
        if S == 0 or N == 0:
            z = lambda *s: torch.zeros(s, device=dev)
            return z(0), z(0, D), torch.zeros(0, device=dev, dtype=torch.int32), z(0), z(0, D)
        start = max(0, S - N)
        X = self.samples[start:]
        amap = self.mapping[start:]
        M = int(self.means.shape[0])
        if self.keep_samples and M > 0 and self._last_batch is not None and S - start == self._last_batch[0]:
            # exactly the batch stored last (every no-reuse iteration): its components are the last ones stored -- known
            # on the host, no device read.  (The reference derives them from unique(mapping), sample_db.py:221.)
            lo, hi = M - self._last_batch[1], M - 1
        elif self.keep_samples and M > 0:
            lo, hi = int(amap.min().item()), int(amap.max().item())      # mapping is non-decreasing
        else:
            lo, hi = 0, M - 1
        if self.count_override is not None:      # sharded: this rank only holds a slice of the iteration's samples
            count = self.count_override.to(torch.float32)
        else:
            count = torch.zeros(hi - lo + 1, device=dev, dtype=torch.float32)
            count.scatter_add_(0, (amap - lo).long(), torch.ones(amap.shape[0], device=dev))
        weight = count / torch.sum(count)
        sl = slice(lo, hi + 1)
        consts = None if self.diagonal_covariances else self.consts[sl]
        bg, lq = self.evaluate_background(weight, self.means[sl], self.chols[sl], self.inv_chols[sl],
                                          X.contiguous(), consts)
        return bg, X, amap, self.target_lnpdfs[start:], self.target_grads[start:]
