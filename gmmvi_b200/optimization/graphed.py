"""One SAMTRON iteration as a CUDA graph.

The reference wraps `train_iter` in `tf.function` (optimization/gmmvi.py:99-103) so that TensorFlow runs it as ONE graph
instead of op by op.  The equivalent here is a captured CUDA graph: at D = 10 ... 20 (BASELINE configurations C1 / C2)
and on 8 GPUs (C5: ~4 ms of kernels per rank) the iteration is bound by the ~170 kernel launches and the Python between
them, not by the GPU.  `GraphedIteration` captures `select samples -> NG estimate -> component update -> weight update`
once per number of components and replays it with a single launch; the host-side parts of the algorithm (adding /
deleting components, metrics, the growing sample database) stay outside and run eagerly between replays.

What makes an iteration replayable:
  * every tensor that lives across iterations (mixture parameters, per-component learner state, stepsizes, reward /
    weight histories, the inverse factors / precisions derived from the parameters and their fp16 split operands) sits
    in a STATIC buffer (`GraphState`, shared by all graphs of one number of components): the captured body reads the
    static buffers and ends by writing the new values back -- the big ones (inverse factors, precisions, gathered
    Cholesky factors, split operands) are produced directly in place, the rest is copied;
  * the noise generator's draw counter is a device scalar the graph increments (gvi_fill_normal_dev_f32), so replays
    draw the subsequences an eager run would have drawn -- results are bit-identical to eager iterations;
  * no host synchronisation inside: the body is the no-reuse iteration (`select_samples_deferred`), whose shapes depend
    only on the number of components; the sample database is appended to after the replay;
  * NCCL collectives of a sharded run are captured like any other kernel.
Supported: both selectors with ratio_reused_samples_to_desired = 0, any estimator / updater / stepsize rule.
A change of the number of components drops the graphs; they are captured again once K has been stable for a while."""
from __future__ import annotations

import gc

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


class GraphCaptureError(RuntimeError):
    """The iteration could not be captured (e.g. a user-supplied target synchronises with the host inside
    `log_density`); the caller falls back to op-by-op iterations."""


class GraphState:
    """Static buffers of one number of components, shared by the graphs captured for it."""

    def __init__(self, gmmvi):
        self.gmm, self.slots = _slots(gmmvi)
        gmm = self.gmm
        self.num_components = gmm.num_components
        self.full = not gmm.diagonal_covs
        if gmm._chol_work is not None:
            _ = gmm.chol_cov                        # finish an in-flight all-gather before cloning
        self.statics = [getattr(o, n).detach().clone().contiguous() for o, n in self.slots]
        self.prep = [t.detach().clone().contiguous() for t in gmm.prepared(need_prec=True)] if self.full else None
        self.counter = torch.zeros(1, device=gmm.device, dtype=torch.int64)
        self.host_subsequence = None                # value of the device counter, tracked on the host
        ops.clear_split_registry()
        self.install()

    def chol_static(self):
        return self.statics[2]

    def sync_from_model(self):
        """Bring the static buffers up to date with whatever eager code did to the model since the last replay."""
        gmm = self.gmm
        dirty = False
        for (obj, name), s in zip(self.slots, self.statics):
            cur = getattr(obj, name) if name != "_chol_cov" else gmm.chol_cov
            if cur is not s:
                s.copy_(cur)
                dirty = True
        if self.full:
            p = gmm._prepared
            if dirty or p is None or p[0] != gmm._version or p[1] is not self.prep[0]:
                gmm._prepared = None if dirty else gmm._prepared
                for s, new in zip(self.prep, gmm.prepared(need_prec=True)):
                    if new is not s:
                        s.copy_(new)
                for t in self.prep[:2]:
                    ops.invalidate_split(t)
        self.install()
        if self.full:
            # a replay starts from VALID split operands (no split kernel is captured at the start of the body)
            for t in self.prep[:2]:
                ops.split_static_now(t)

    def install(self):
        """Point the model / learner attributes at the static buffers and declare the derived operands valid for them."""
        gmm = self.gmm
        for (obj, name), s in zip(self.slots, self.statics):
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
        if self.full:
            gmm._prepared = (gmm._version, self.prep[0], self.prep[1], self.prep[2])
            # results that can be produced in place: the next update gathers / prepares straight into the static buffers
            gmm._static_out = {"chol": self.chol_static(), "linv": self.prep[0], "prec": self.prep[1], "cst": self.prep[2]}
            for t, kind in zip(self.prep[:2], ("lower", "full")):
                if ops.split_registered(t) is None:
                    ops.register_split_buffers(t, kind, valid=False)
        else:
            gmm._prepared = None
            gmm._static_out = None

    def set_counter(self):
        if self.host_subsequence != rng._state["subsequence"]:      # eager draws advanced the host counter
            self.host_subsequence = rng._state["subsequence"]
            self.counter.fill_(self.host_subsequence)

    def release(self):
        self.gmm._static_out = None
        ops.clear_split_registry()


class GraphedIteration:
    def __init__(self, gmmvi, noise_buffer=None):
        sel = gmmvi.sample_selector
        if not hasattr(sel, "select_samples_deferred") or sel.reused_samples_per_component != 0:
            raise NotImplementedError("CUDA-graph iterations need a selector without sample reuse "
                                      "(ratio_reused_samples_to_desired = 0)")
        self.gmmvi = gmmvi
        self.noise_buffer = noise_buffer            # static [N, D] buffer the caller fills before each replay (optional)
        self.graph = None
        self.num_components = None
        self.replays = 0

    def capture(self):
        g = self.gmmvi
        st = getattr(g, "_graph_state", None)
        K = (g.model.model if hasattr(g.model, "model") else g.model).num_components
        if st is None or st.num_components != K:
            st = g._graph_state = GraphState(g)
        else:
            st.sync_from_model()
        self.state = st
        gmm = st.gmm
        self.num_components = K
        if st.full:                                  # the split operands of the current factors, computed once, eagerly
            for t in st.prep[:2]:
                ops.split_static_now(t)
        st.set_counter()
        ops.clear_caches()
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        written0 = g.sample_db.num_samples_written
        kernels0, calls0 = ops.KERNELS, ops.LAUNCHES
        # Python's cyclic collector must not run inside the capture: it may destroy an older, unreachable CUDA graph
        # (GMMVI <-> GraphedIteration is a cycle) and freeing its memory pool invalidates the capture in progress.
        gc_was_enabled = gc.isenabled()
        gc.disable()                    # (no gc.collect() here: a full collection costs tens of ms per capture)
        updates0 = g.num_updates
        try:
            self._capture(g, st)
        except Exception as e:
            # nothing was executed: put the host-side counters back, leave the model on the (valid) static buffers
            g.sample_db.num_samples_written, g.num_updates = written0, updates0
            ops.KERNELS, ops.LAUNCHES = kernels0, calls0
            self.graph = None
            st.install()
            ops.clear_caches()
            for t in (st.prep[:2] if st.full else ()):
                ops.invalidate_split(t)
            torch.cuda.synchronize()
            raise GraphCaptureError(f"CUDA-graph capture of the iteration failed: {type(e).__name__}: {e}") from e
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

    def _capture(self, g, st):
        # capture_begin / capture_end by hand instead of the torch.cuda.graph context: the context empties the caching
        # allocator (a cudaFree of every cached block, ~0.2 s) on every entry, which an adaptive run that captures again
        # after each change of the number of components cannot afford.  All graphs of one GMMVI share a memory pool.
        gmm = st.gmm
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
                self._captured_region(g, st)
            except BaseException:
                try:
                    self.graph.capture_end()          # leave capture mode; an invalidated capture raises again here
                except Exception:
                    pass
                # torch's CUDA generator stays in "capturing" state when capture_end fails (every later torch.randn would
                # raise "Offset increment outside graph capture"): a trivial capture that succeeds puts it back
                try:
                    dummy = torch.cuda.CUDAGraph()
                    dummy.capture_begin(capture_error_mode="thread_local")
                    st.counter.add_(0)
                    dummy.capture_end()
                    del dummy
                except Exception:
                    pass
                raise
            self.graph.capture_end()
        torch.cuda.current_stream(gmm.device).wait_stream(stream)

    def _captured_region(self, g, st):
        gmm = st.gmm
        rng.begin_device_mode(st.counter)
        try:
            self.payload = self._body()
        finally:
            self.draws = rng.end_device_mode()
        st.counter.add_(self.draws)
        if st.full:
            for s, new in zip(st.prep, gmm.prepared(need_prec=True)):
                if new is not s:                          # produced elsewhere (shape did not fit the in-place path)
                    s.copy_(new)
                    ops.invalidate_split(s)
            # The next replay starts from these buffers WITHOUT splitting them again (no split kernel is captured at the
            # start of the body: the operands were valid then), so every operand rewritten during this iteration must be
            # brought up to date here; the inverse factors already were, for the weight-update pass.
            for s in st.prep[:2]:
                ops.split_static_now(s)
        _ = gmm.chol_cov                                  # waits for an in-flight all-gather of a sharded update
        for (obj, name), s in zip(st.slots, st.statics):
            new = getattr(obj, name)
            if new is not s:
                s.copy_(new)
        st.install()

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
        st = self.state
        if any(getattr(o, n) is not s for (o, n), s in zip(st.slots, st.statics)):
            st.sync_from_model()            # eager code replaced parameters / learner state since the last replay
        st.set_counter()
        self.graph.replay()
        st.host_subsequence += self.draws
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
