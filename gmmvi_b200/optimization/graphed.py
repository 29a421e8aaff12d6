"""One SAMTRON iteration as a CUDA graph.

The reference wraps `train_iter` in `tf.function` (optimization/gmmvi.py:99-103) so that TensorFlow runs it as ONE graph
instead of op by op.  The equivalent here is a captured CUDA graph: at D = 10 ... 20 (BASELINE configurations C1 / C2)
and on 8 GPUs (C5: ~4 ms of kernels per rank) the iteration is bound by the ~170 kernel launches and the Python between
them, not by the GPU.  `GraphedIteration` captures `select samples -> NG estimate -> component update -> weight update`
once per number of components and replays it with a single launch; the host-side parts of the algorithm (adding /
deleting components, metrics, the growing sample database) stay outside and run eagerly between replays.

What makes an iteration replayable:
  * every tensor that lives across iterations (mixture parameters, per-component learner state, stepsizes, reward /
    weight histories, the inverse factors derived from the parameters) sits in a STATIC buffer: the captured body reads
    the static buffers, computes new values into graph-private memory, and ends by copying them back;
  * the noise generator's draw counter is a device scalar the graph increments (gvi_fill_normal_dev_f32), so replays
    draw the subsequences an eager run would have drawn -- results are bit-identical to eager iterations;
  * no host synchronisation inside: the body is the no-reuse iteration (`select_samples_deferred`), whose shapes depend
    only on the number of components; the sample database is appended to after the replay;
  * NCCL collectives of a sharded run are captured like any other kernel.
Supported: component-based selector with ratio_reused_samples_to_desired = 0, any estimator / updater / stepsize rule.
A change of the number of components drops the graph; it is captured again once K has been stable for an iteration."""
from __future__ import annotations

import torch

from .. import ops, rng


def _slots(gmmvi):
    """(object, attribute) of every tensor that persists from one iteration to the next."""
    w = gmmvi.model
    gmm = w.model if hasattr(w, "model") else w
    out = [(gmm, "log_weights"), (gmm, "_means"), (gmm, "_chol_cov")]
    if w is not gmm:
        out += [(w, n) for n in ("l2_regularizers", "last_log_etas", "num_received_updates", "stepsizes", "reward_history",
                                 "weight_history")]
    for adapter in (gmmvi.weight_stepsize_adapter, gmmvi.component_stepsize_adapter):
        for n in ("stepsize", "elbo_history", "num_weight_updates"):
            if isinstance(getattr(adapter, n, None), torch.Tensor):
                out.append((adapter, n))
    return gmm, out


class GraphedIteration:
    def __init__(self, gmmvi, noise_buffer=None):
        from .gmmvi_modules.sample_selector import VipsSampleSelector
        sel = gmmvi.sample_selector
        if not isinstance(sel, VipsSampleSelector) or sel.reused_samples_per_component != 0:
            raise NotImplementedError("CUDA-graph iterations need the component-based selector without sample reuse")
        self.gmmvi = gmmvi
        self.noise_buffer = noise_buffer            # static [N, D] buffer the caller fills before each replay (optional)
        self.graph = None
        self.num_components = None
        self.replays = 0

    # ------------------------------------------------------------------------------------------------------------
    def _install(self, gmm, slots, statics, prep):
        for (obj, name), s in zip(slots, statics):
            setattr(obj, name, s)
        gmm._version += 1
        if gmm._chol_work is not None:
            gmm._chol_work.wait()
            gmm._chol_work = None
        gmm._local_chol = None
        if gmm.shard is not None:
            r = gmm.shard.component_range(gmm.num_components)
            if r is not None:
                gmm._local_chol = (gmm._version, r[0], r[1], gmm._chol_cov[r[0]:r[1]])
        gmm._prepared = None if prep is None else (gmm._version, prep[0], prep[1], prep[2])

    def capture(self):
        g = self.gmmvi
        gmm, slots = _slots(g)
        self.num_components = gmm.num_components
        dev = gmm.device
        full = not gmm.diagonal_covs
        statics = [getattr(o, n).detach().clone().contiguous() for o, n in slots]
        prep = [t.detach().clone().contiguous() for t in gmm.prepared(need_prec=True)] if full else None
        self.counter = torch.tensor([rng._state["subsequence"]], device=dev, dtype=torch.int64)
        self._install(gmm, slots, statics, prep)
        ops.clear_caches()
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        written0 = g.sample_db.num_samples_written
        kernels0, calls0 = ops.KERNELS, ops.LAUNCHES
        # Python's cyclic collector must not run inside the capture: it may destroy an older, unreachable CUDA graph
        # (GMMVI <-> GraphedIteration is a cycle) and freeing its memory pool invalidates the capture in progress.
        import gc
        gc_was_enabled = gc.isenabled()
        gc.disable()                    # (no gc.collect() here: a full collection costs tens of ms per capture)
        try:
            self._capture(g, gmm, slots, statics, prep, full)
        finally:
            if gc_was_enabled:
                gc.enable()
        ops.clear_caches()
        # kernels one replay launches (ops.kernel_launches() stays a count of what ran on the device)
        self.kernels, self.calls = ops.KERNELS - kernels0, ops.LAUNCHES - calls0
        ops.KERNELS, ops.LAUNCHES = kernels0, calls0
        self.samples_per_iteration = g.sample_db.num_samples_written - written0
        # the capture run executed nothing: the first replay performs the iteration that was captured
        g.sample_db.num_samples_written = written0
        return self

    def _capture(self, g, gmm, slots, statics, prep, full):
        # capture_begin / capture_end by hand instead of the torch.cuda.graph context: the context empties the caching
        # allocator (a cudaFree of every cached block, ~0.2 s) on every entry, which an adaptive run that captures again
        # after each change of the number of components cannot afford.  All graphs of one GMMVI share a memory pool.
        live = [x for x in ((g._graph or {}).values() if isinstance(g._graph, dict) else []) if x.graph is not None]
        live += [x for x in (getattr(g, "_graph_retired", None) or []) if x.graph is not None]
        if getattr(g, "_graph_pool", None) is None or not live:      # a pool only exists while some graph uses it
            g._graph_pool = torch.cuda.graph_pool_handle()
        if getattr(g, "_graph_stream", None) is None:
            g._graph_stream = torch.cuda.Stream(gmm.device)
        stream = g._graph_stream
        stream.wait_stream(torch.cuda.current_stream(gmm.device))
        with torch.cuda.stream(stream):
            self.graph.capture_begin(pool=g._graph_pool, capture_error_mode="thread_local")
            try:
                self._captured_region(g, gmm, slots, statics, prep, full)
            finally:
                self.graph.capture_end()
        torch.cuda.current_stream(gmm.device).wait_stream(stream)

    def _captured_region(self, g, gmm, slots, statics, prep, full):
        if True:
            rng.begin_device_mode(self.counter)
            try:
                self.payload = self._body()
            finally:
                self.draws = rng.end_device_mode()
            self.counter.add_(self.draws)
            if full:
                for s, new in zip(prep, gmm.prepared(need_prec=True)):
                    if new is not s:
                        s.copy_(new)
            _ = gmm.chol_cov                          # waits for an in-flight all-gather of a sharded update
            for (obj, name), s in zip(slots, statics):
                new = getattr(obj, name)
                if new is not s:
                    s.copy_(new)
            self._install(gmm, slots, statics, prep)

    def _body(self):
        g = self.gmmvi
        out, payload = g.sample_selector.select_samples_deferred(noise=self.noise_buffer)
        if payload is None:
            # use_sample_database = False: the database is REPLACED by the iteration's samples (sample_db.py:125-135)
            samples, mapping, bg, lnpdfs, grads = out
            gmm = g.model
            chols = gmm.chol_cov if gmm.shard is None else gmm.chol_cov_handle
            g.sample_db.add_samples(samples, gmm.means, chols, lnpdfs, grads, mapping, prepared=g.sample_selector._prepared())
        else:
            g.sample_db.num_samples_written += int(out[0].shape[0])
        g._run_updates(*out)
        g.num_updates -= 1                             # counted by replay()
        return payload

    def replay(self):
        g = self.gmmvi
        self.graph.replay()
        self.replays += 1
        ops.clear_caches()                  # the replay rewrote the static buffers behind the caches' keys
        ops.KERNELS += self.kernels
        ops.LAUNCHES += self.calls
        rng.advance(self.draws)
        g.num_updates += 1
        if self.payload is not None:
            # use_sample_database = True: append what the graph drew (thinning and growth are host-side bookkeeping)
            samples, means, chols, lnpdfs, grads, mapping, prepared = self.payload
            g.sample_db.add_samples(samples, means, chols, lnpdfs, grads, mapping, prepared=prepared)
        else:
            g.sample_db.num_samples_written += self.samples_per_iteration
