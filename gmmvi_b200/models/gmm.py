"""GMM base class on torch CUDA tensors + sm_100a kernels.

Mirrors the public surface of the reference's `gmmvi.models.gmm.GMM` (models/gmm.py:5-418): same method
names, argument meaning and return shapes; tensors are CUDA `torch.Tensor`s (fp32, int32 indices).
State: log_weights[K], means[K,D], chol_cov[K,D,D] (full) or [K,D] std-devs (diagonal).
"""
from __future__ import annotations

from math import pi
from typing import Optional, Tuple

import torch

from .. import ops, rng


def _as_param(x, device):
    return torch.as_tensor(x, dtype=torch.float32).to(device).contiguous().clone()


class GMM:
    def __init__(self, log_weights: torch.Tensor, means: torch.Tensor, chol_covs: torch.Tensor):
        # models/gmm.py:27-34
        self.diagonal_covs = chol_covs.dim() == 2
        self.num_dimensions = int(means.shape[1])
        self._const_log_det = 0.5 * self.num_dimensions * torch.log(torch.tensor(2 * pi)).item()
        self._version = 0
        self.shard = None              # gmmvi_b200.distributed.ShardContext when samples are sharded over GPUs
        self._prepared = None          # (version, linv, prec, cst) for full covariances
        self._chol_work = None         # in-flight all-gather of chol_cov (sharded component update)
        self._local_chol = None        # (version, a, b, chol[a:b]) of the most recent sharded update
        self._static_out = None        # {"chol" | "linv" | "prec" | "cst": buffer} a graph runner wants results written to
        self.log_weights = log_weights
        self._means = means
        self._chol_cov = chol_covs
        self.replace_weights(self.log_weights)

    # ---- parameter access; assignments invalidate the derived-operand cache -----------------------
    @property
    def means(self) -> torch.Tensor:
        return self._means

    @means.setter
    def means(self, v):
        self._means = v
        self._version += 1

    @property
    def chol_cov(self) -> torch.Tensor:
        if self._chol_work is not None:          # asynchronous all-gather of the sharded update: first use waits
            self._chol_work.wait()
            self._chol_work = None
        return self._chol_cov

    @chol_cov.setter
    def chol_cov(self, v):
        if self._chol_work is not None:
            self._chol_work.wait()
        self._chol_work = None
        self._local_chol = None
        self._chol_cov = v
        self._version += 1

    def set_components_sharded(self, a, b, means_local, chol_local, extras_local=()):
        """Result of a component-sharded update: this rank computed the components [a, b).  Three collectives distribute
        it: ONE small all-gather of (means | log-normalisers | extras) rows, ONE all-gather of the inverse factors (the
        next log-density pass of every rank needs all of them at once) and an ASYNCHRONOUS all-gather of the factors
        themselves, which in steady state nobody waits for: a rank samples from / updates only its own components
        (`local_chol`), the precisions are formed from the gathered inverses.  `extras_local`: [b - a] vectors
        (success, eta, KL) that travel with the small gather; returns them gathered, as float32 [K] each."""
        K = self.num_components
        shard = self.shard
        so = self._static_out or {}
        if not self.diagonal_covs:
            l_loc, _, c_loc = ops.prepare_full(chol_local, want_prec=False)[:3]
        cols = [means_local] + ([c_loc.unsqueeze(1)] if not self.diagonal_covs else []) \
            + [e.to(torch.float32).unsqueeze(1) for e in extras_local]
        packed = shard.all_gather_rows(torch.cat(cols, dim=1).contiguous(), K)
        D = self.num_dimensions
        # the inverse factors FIRST: NCCL runs a communicator's collectives in issue order, and the weight-update pass
        # waits for this gather, not for the factors' (which nobody waits for in steady state)
        linv = None if self.diagonal_covs else shard.all_gather_rows(l_loc, K, out=so.get("linv"))
        self.means = packed[:, :D].contiguous()
        chol_full, chol_work = shard.all_gather_rows_async(chol_local, K, out=so.get("chol"))
        self.chol_cov = chol_full               # (the setter waits for a previous in-flight gather)
        self._chol_work = chol_work
        self._local_chol = (self._version, a, b, chol_local)
        c = D
        if not self.diagonal_covs:
            ops.invalidate_split(linv)
            self._prepared = (self._version, linv, None, packed[:, c].contiguous())
            c += 1
        return [packed[:, c + i].contiguous() for i in range(len(extras_local))]

    @property
    def device(self):
        return self._means.device

    def local_chol(self, a: int, b: int) -> torch.Tensor:
        """Cholesky factors of the components [a, b) WITHOUT waiting for an in-flight all-gather when they are the rows
        this rank computed itself in the most recent sharded update."""
        lc = self._local_chol
        if lc is not None and lc[0] == self._version and (lc[1], lc[2]) == (a, b):
            return lc[3]
        return self.chol_cov[a:b].contiguous()

    @property
    def chol_cov_handle(self) -> torch.Tensor:
        """The factor tensor for STORAGE only (sample database): does not wait for an in-flight all-gather; a reader
        must go through `chol_cov`."""
        return self._chol_cov

    def prepared(self, need_prec: bool = True):
        """(linv, prec, cst): inverse Cholesky factors, precisions and log-normalisers of the current
        full-covariance components (one `gvi_prepare_full_f32` per parameter change).  In sharded runs every rank
        inverts the factors of its own components and ONE all-gather distributes `linv` (+ the [K] normalisers); the
        precisions are formed locally from the gathered `linv` (a tensor-core batched GEMM, cheaper than moving another
        K D^2 floats) the first time a caller asks for them (`need_prec`)."""
        if self._prepared is None or self._prepared[0] != self._version:
            rng_ = self.shard.component_range(self.num_components) if self.shard is not None else None
            so = self._static_out or {}
            if rng_ is None:
                linv, prec, cst, _ = ops.prepare_full(self.chol_cov, want_prec=True,
                                                      out=(so.get("linv"), so.get("prec"), so.get("cst")))
                ops.invalidate_split(linv)
                ops.invalidate_split(prec)
            else:       # components sharded over the ranks
                a, b = rng_
                K = self.num_components
                l_loc, _, c_loc = ops.prepare_full(self.local_chol(a, b), want_prec=False)[:3]
                linv = self.shard.all_gather_rows(l_loc, K, out=so.get("linv"))
                ops.invalidate_split(linv)
                cst = self.shard.all_gather_rows(c_loc, K)
                prec = None
            self._prepared = (self._version, linv, prec, cst)
        if need_prec and self._prepared[2] is None:
            v, linv, _, cst = self._prepared
            prec = ops.bgemm(linv, linv, transA=True, out=(self._static_out or {}).get("prec"))       # P = L^-T L^-1
            ops.invalidate_split(prec)
            self._prepared = (v, linv, prec, cst)
        return self._prepared[1:]

    # ---- abstract per-family pieces ---------------------------------------------------------------
    def sample_from_component(self, index: int, num_samples: int, noise: Optional[torch.Tensor] = None):
        """models/gmm.py:36 -- draws `num_samples` from component `index` -> [num_samples, D].
        `noise` ([num_samples, D], the transpose of the reference's (D, n) draw) may be injected."""
        n = int(num_samples)
        D = self.num_dimensions
        if noise is None:
            noise = ops.fill_normal(n, D, rng.seed(), rng.device_counter() or rng.next_subsequence(), 0, self.device)
        offsets = torch.tensor([0, n], device=self.device, dtype=torch.int32)
        X, _ = ops.sample_components(self.diagonal_covs, noise, offsets, self._means[index:index + 1].contiguous(),
                                     self.chol_cov[index:index + 1].contiguous(), n)
        return X

    def component_log_density(self, index: int, samples: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    def component_log_densities(self, samples: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    def component_marginal_log_densities(self, samples: torch.Tensor, dimension: int) -> torch.Tensor:
        raise NotImplementedError

    def gaussian_entropy(self, chol: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    def add_component(self, initial_weight, initial_mean, initial_cov):
        raise NotImplementedError

    def _mixture_grad(self, samples, lq, logq, logw=None, index=None):
        raise NotImplementedError

    @property
    def covs(self) -> torch.Tensor:
        raise NotImplementedError

    # ---- common functionality -----------------------------------------------------------------------
    def sample_categorical(self, num_samples: int, uniforms: Optional[torch.Tensor] = None) -> torch.Tensor:
        """models/gmm.py:124-137 (all-False argmax -> component 0, quirk 5)."""
        thresholds = torch.cumsum(self.weights, 0).unsqueeze(0)
        eps = uniforms.reshape(-1, 1) if uniforms is not None else torch.rand((int(num_samples), 1), device=self.device)
        return torch.argmax((eps < thresholds).to(torch.int32), dim=-1).to(torch.int32)

    def sample(self, num_samples: int, uniforms: Optional[torch.Tensor] = None,
               noise: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """models/gmm.py:139-163: samples grouped by component, indices in draw order (quirk 4).  `uniforms` [n] and
        `noise` [n, D] (rows in component order) inject the random draws (parity tests)."""
        sampled_components = self.sample_categorical(num_samples, uniforms)
        counts = torch.zeros(self.num_components, device=self.device, dtype=torch.int32)
        counts.scatter_add_(0, sampled_components.long(), torch.ones_like(sampled_components))
        samples, _ = self.sample_from_components_no_shuffle(counts, noise=noise)
        return samples, sampled_components

    @property
    def weights(self) -> torch.Tensor:
        return torch.exp(self.log_weights)

    def replace_weights(self, new_log_weights: torch.Tensor):
        """models/gmm.py:173-181 (normalises)."""
        v = torch.as_tensor(new_log_weights, dtype=torch.float32, device=self.device)
        self.log_weights = (v - torch.logsumexp(v, 0)).contiguous()

    def log_densities_also_individual(self, samples: torch.Tensor):
        """models/gmm.py:183-201 -> (log q(x)[N], component log densities [K,N])."""
        lq = self.component_log_densities(samples)
        return ops.mixture_lse(lq, self.log_weights), lq

    def log_density(self, samples: torch.Tensor) -> torch.Tensor:
        """models/gmm.py:203-216."""
        return self.log_densities_also_individual(samples)[0]

    def marginal_log_density(self, samples: torch.Tensor, dimension: int) -> torch.Tensor:
        """models/gmm.py:218-236."""
        lq = self.component_marginal_log_densities(samples, dimension).contiguous()
        return ops.mixture_lse(lq, self.log_weights)

    def density(self, samples: torch.Tensor) -> torch.Tensor:
        return torch.exp(self.log_density(samples))

    def component_entropies(self) -> torch.Tensor:
        """models/gmm.py:249-260."""
        D = self.num_dimensions
        chol = self.chol_cov          # the property waits for an in-flight all-gather of the sharded update
        diag = chol if self.diagonal_covs else torch.diagonal(chol, dim1=1, dim2=2)
        return 0.5 * D * (torch.log(torch.tensor(2 * pi)).item() + 1) + torch.sum(torch.log(diag), dim=1)

    def get_average_entropy(self) -> torch.Tensor:
        """models/gmm.py:262-272."""
        return torch.sum(self.weights * self.component_entropies())

    def log_density_and_grad(self, samples: torch.Tensor):
        """models/gmm.py:274-300 -> (log q[N], grad[N,D], component log densities[K,N]).  The reference
        back-propagates through the triangular solve; here the gradient is the analytic
        -sum_k r_kn Sigma_k^-1 (x_n - mu_k) evaluated by one fused kernel."""
        logq, lq = self.log_densities_also_individual(samples)
        return logq, self._mixture_grad(samples, lq, logq), lq

    def component_log_density_and_grad(self, index: int, samples: torch.Tensor):
        """models/gmm.py:302-321 -> (component log density [N], its gradient [N,D])."""
        lq = self.component_log_density(index, samples)
        zero = torch.zeros(1, device=self.device)
        grad = self._mixture_grad(samples, lq.unsqueeze(0).contiguous(), lq, logw=zero, index=int(index))
        return lq, grad

    @property
    def num_components(self) -> int:
        return int(self.log_weights.shape[0])

    def _offsets(self, samples_per_component):
        n = torch.as_tensor(samples_per_component, device=self.device).to(torch.int32)
        offsets = torch.zeros(n.shape[0] + 1, device=self.device, dtype=torch.int32)
        offsets[1:] = torch.cumsum(n, 0)
        return n, offsets

    def sample_from_components_no_shuffle(self, samples_per_component, noise: Optional[torch.Tensor] = None,
                                          total: Optional[int] = None, max_per_component: Optional[int] = None,
                                          row_offset: int = 0, component_range=None):
        """models/gmm.py:361-386 -> (samples[N,D] in component order, mapping[N] int32).
        `total` / `max_per_component` let a caller that already knows them avoid the host sync;
        `noise` injects the standard-normal draws ([N,D]); `component_range` = (a, b) promises that only the components
        [a, b) have a non-zero count (sharded runs)."""
        n, offsets = self._offsets(samples_per_component)
        if total is None or max_per_component is None:
            total, max_per_component = int(offsets[-1].item()), int(n.max().item()) if n.numel() else 0
        D = self.num_dimensions
        if noise is None:
            noise = ops.fill_normal(total, D, rng.seed(), rng.device_counter() or rng.next_subsequence(), row_offset, self.device)
        if total == 0:
            return torch.zeros((0, D), device=self.device), torch.zeros(0, device=self.device, dtype=torch.int32)
        if component_range is not None:
            # sharded iteration whose rows all belong to the components [a, b) this rank updated itself: their factors
            # are at hand, the all-gather of the other ranks' factors need not have finished
            a, b = component_range
            X, mapping = ops.sample_components(self.diagonal_covs, noise, (offsets[a:b + 1] - offsets[a]).contiguous(),
                                               self._means[a:b].contiguous(), self.local_chol(a, b), max_per_component)
            return X, mapping + a
        return ops.sample_components(self.diagonal_covs, noise, offsets, self._means, self.chol_cov,
                                     max_per_component)

    def sample_from_components(self, samples_per_component) -> torch.Tensor:
        """models/gmm.py:340-359 (shuffled)."""
        samples, _ = self.sample_from_components_no_shuffle(samples_per_component)
        return samples[torch.randperm(samples.shape[0], device=self.device)]

    def remove_component(self, idx: int):
        """models/gmm.py:388-399."""
        idx = int(idx)
        keep = [i for i in range(self.num_components) if i != idx]
        sel = torch.tensor(keep, device=self.device, dtype=torch.long)
        self.replace_weights(self.log_weights[sel])
        self.means = self._means[sel].contiguous()
        self.chol_cov = self.chol_cov[sel].contiguous()

    def replace_components(self, new_means, new_chols):
        """models/gmm.py:401-418."""
        if isinstance(new_means, (list, tuple)):
            new_means = torch.stack(list(new_means), 0)
        if isinstance(new_chols, (list, tuple)):
            new_chols = torch.stack(list(new_chols), 0)
        self.means = new_means.contiguous()
        self.chol_cov = new_chols.contiguous()
