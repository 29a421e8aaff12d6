"""CPU tests of the host side: C-ABI symbols, loud failure without the extension, config semantics, dispatch
errors, and the multi-rank sharding protocol on gloo (world_size 2)."""
import ctypes
import os
import re
import socket

import numpy as np
import pytest
import torch

import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "gmmvi_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gvi_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from gmmvi_b200 import _lib
    names = header_symbols()
    assert len(names) >= 20
    handle = ctypes.CDLL(_lib.LIB_PATH)          # loads without a GPU (no compute calls are made)
    for n in names:
        assert hasattr(handle, n), f"{n} declared in include/gmmvi_b200.h but not exported"
    assert set(names) == set(_lib.exported_symbols()), set(names) ^ set(_lib.exported_symbols())
    assert _lib.lib().gvi_version() >= 100
    assert _lib.lib().gvi_logdens_full_tc_supported(256) == 1 and _lib.lib().gvi_logdens_full_tc_supported(100) == 0


def test_missing_extension_fails_loudly(monkeypatch):
    from gmmvi_b200 import _lib
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libgmmvi_b200.so")
    monkeypatch.setattr(_lib, "_lib", None)
    with pytest.raises(_lib.GmmviLibraryError, match="no CPU or PyTorch fallback"):
        _lib.lib()


def test_cpu_tensors_are_rejected_not_silently_computed():
    from gmmvi_b200 import _lib, ops
    with pytest.raises(_lib.GmmviLibraryError):
        ops.mixture_lse(torch.zeros(2, 3), torch.zeros(2))
    with pytest.raises(_lib.GmmviLibraryError):
        ops.logdens_diag(torch.zeros(4, 3), torch.zeros(2, 3), torch.ones(2, 3))


def test_bad_arguments_return_error_codes_without_a_gpu():
    from gmmvi_b200 import _lib
    lib = _lib.lib()
    assert lib.gvi_logdens_full_f32(None, -1, 4, None, None, None, 1, None, None) == -1
    assert b"bad sizes" in lib.gvi_last_error()
    assert lib.gvi_update_full_f32(7, None, None, None, None, None, None, None, 1, 4, 1.0, None, None, None, None,
                                   None, None, None, 0, None) == -1
    assert b"unknown mode" in lib.gvi_last_error()
    assert lib.gvi_update_diag_f32(1, None, None, None, None, None, None, None, 1, 4, 1.0, None, None, None, None,
                                   None, None) == -1
    assert lib.gvi_update_full_workspace(2, 4) > 0 and lib.gvi_prepare_full_workspace(0, 4) == 0


def test_config_assembly_matches_reference_semantics():
    from gmmvi_b200.configs import (get_default_algorithm_config, get_default_config, get_default_experiment_config,
                                    update_config)
    algo = get_default_algorithm_config("SAMTRON")
    assert algo["ng_estimator_type"] == "Stein" and algo["num_component_adapter_type"] == "adaptive"
    assert algo["sample_selector_type"] == "component-based" and algo["ng_based_updater_type"] == "trust-region"
    assert algo["component_stepsize_adapter_type"] == "improvement-based"
    assert algo["weight_updater_type"] == "trust-region"
    assert algo["weight_stepsize_adapter_type"] == "improvement_based"          # underscore (reference quirk 14)
    zep = get_default_algorithm_config("zepyfux")                                # letters are case-insensitive
    assert zep["ng_estimator_type"] == "MORE" and zep["ng_estimator_config"]["initial_l2_regularizer"] == 1e-12
    env = get_default_experiment_config("stm20")
    merged = update_config(env, {"model_initialization": {"num_initial_components": 45}, "start_seed": 3})
    assert merged["model_initialization"]["num_initial_components"] == 45
    assert merged["model_initialization"]["prior_scale"] == 100.0              # nested dicts merge
    assert merged["start_seed"] == 3 and env["start_seed"] == 10000             # top level is a copy
    m2 = update_config(algo, {"num_component_adapter_config": {"thresholds_for_add_heuristic": [1.0]}})
    assert m2["num_component_adapter_config"]["thresholds_for_add_heuristic"] == [1.0]   # lists are replaced
    full = get_default_config("SAMTRON", "planar_robot_4")
    assert full["environment_name"] == "PlanarRobot4" and len(full["model_initialization"]["prior_scale"]) == 10
    with pytest.raises(FileNotFoundError):
        get_default_experiment_config("nope")


@pytest.mark.parametrize("key,cls_path", [
    ("ng_estimator_type", "ng_estimator.NgEstimator"),
    ("ng_based_updater_type", "ng_based_component_updater.NgBasedComponentUpdater"),
    ("weight_updater_type", "weight_updater.WeightUpdater"),
    ("sample_selector_type", "sample_selector.SampleSelector"),
    ("component_stepsize_adapter_type", "component_stepsize_adaptation.ComponentStepsizeAdaptation"),
    ("weight_stepsize_adapter_type", "weight_stepsize_adaptation.WeightStepsizeAdaptation"),
    ("num_component_adapter_type", "component_adaptation.ComponentAdaptation"),
])
def test_unknown_module_type_raises_value_error(key, cls_path):
    import importlib
    from gmmvi_b200.configs import get_default_config
    mod, cls = cls_path.split(".")
    C = getattr(importlib.import_module(f"gmmvi_b200.optimization.gmmvi_modules.{mod}"), cls)
    cfg = get_default_config("SAMTRON", "stm20")
    cfg[key] = "no-such-type"

    class Dummy:
        device = "cpu"
    args = {"ng_estimator_type": (cfg, 1.0, Dummy()), "sample_selector_type": (cfg, Dummy(), None, None),
            "num_component_adapter_type": (cfg, Dummy(), None, None, 0.0, 1.0)}.get(key, (cfg, Dummy()))
    with pytest.raises((ValueError, KeyError)):
        C.build_from_config(*args)


def test_shard_index_arithmetic():
    from gmmvi_b200.distributed import ShardContext
    counts = [128] * 512
    seen = []
    for world in (1, 2, 4, 8):
        rows = 0
        for r in range(world):
            sc = ShardContext(r, world)
            local, lo = sc.local_counts(counts)
            assert lo == sc.row_range(sum(counts))[0] and sum(local) == 65536 // world
            rows += sum(local)
            seen.append((world, r, lo))
        assert rows == 65536
    # ragged: coverage and order are preserved for any split
    counts = [5, 0, 7, 3, 1]
    for world in (2, 3, 5, 7):
        tot = np.zeros(len(counts), int)
        prev_hi = 0
        for r in range(world):
            sc = ShardContext(r, world)
            lo, hi = sc.row_range(sum(counts))
            assert lo == prev_hi
            prev_hi = hi
            tot += np.array(sc.local_counts(counts)[0])
        assert prev_hi == sum(counts) and tot.tolist() == counts
    assert ShardContext(1, 4).component_range(512) == (128, 256) and ShardContext(1, 3).component_range(512) is None


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _sharded_worker(rank, world, port, K, N, lq, bg, rho, out_dir):
    """Mirrors ops.importance_weights_sharded + the Stein / update exchange with torch CPU ops over gloo."""
    import torch.distributed as dist
    from gmmvi_b200.distributed import ShardContext
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    sc = ShardContext(rank, world)
    lo, hi = sc.row_range(N)
    lw = torch.as_tensor(lq[:, lo:hi] - bg[None, lo:hi])
    m = sc.all_reduce_max_(lw.max(dim=1).values.clone())
    s = sc.all_reduce_sum_(torch.exp(lw - m[:, None]).sum(1))
    lse = m + torch.log(s)
    s2 = sc.all_reduce_sum_(torch.exp(lw - lse[:, None]).sum(1))
    W = torch.exp(lw - lse[:, None]) / s2[:, None]
    dot = sc.all_reduce_sum_(W @ torch.as_tensor(rho[lo:hi]))
    # component-sharded "update" followed by the all-gather of the new parameters
    a, b = sc.component_range(K)
    mine = torch.arange(a, b, dtype=torch.float64)[:, None] * torch.ones(1, 3, dtype=torch.float64) + 0.5
    gathered = sc.all_gather_rows(mine, K)
    # round 2: gathers into a caller-owned (static) buffer when it fits, into a fresh one when it does not; the
    # asynchronous variant; and the reduce-scatter of raw statistics by component
    static = torch.full((K, 3), -1.0, dtype=torch.float64)
    into = sc.all_gather_rows(mine, K, out=static)
    wrong = sc.all_gather_rows(mine, K, out=torch.zeros((K + 1, 3), dtype=torch.float64))
    async_out, work = sc.all_gather_rows_async(mine, K, out=torch.empty((K, 3), dtype=torch.float64))
    work.wait()
    stats = torch.arange(K * 2, dtype=torch.float64).reshape(K, 2) * (rank + 1)
    full = sc.all_reduce_sum_(stats.clone())     # (gloo has no reduce_scatter_tensor: the all-reduce fallback + own rows)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), W=W.numpy(), dot=dot.numpy(), gathered=gathered.numpy(),
             lo=lo, hi=hi, into_is_static=into.data_ptr() == static.data_ptr(), into=into.numpy(), wrong=wrong.numpy(),
             async_out=async_out.numpy(), own_stats=full[a:b].numpy())
    dist.destroy_process_group()


def test_sharded_importance_weights_match_unsharded_gloo(tmp_path):
    import torch.multiprocessing as mp
    K, N, world = 4, 37, 2
    rng = np.random.default_rng(0)
    lq = rng.standard_normal((K, N)) * 4 - 10
    bg = rng.standard_normal(N) - 9
    rho = rng.standard_normal(N)
    port = _free_port()
    mp.spawn(_sharded_worker, args=(world, port, K, N, lq, bg, rho, str(tmp_path)), nprocs=world, join=True)
    lw = lq - bg
    w = np.exp(lw - O.logsumexp(lw, axis=1, keepdims=True))
    w = w / w.sum(1, keepdims=True)
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    W = np.concatenate([p["W"] for p in parts], axis=1)
    assert parts[0]["lo"] == 0 and parts[0]["hi"] == parts[1]["lo"] and parts[1]["hi"] == N
    assert np.allclose(W, w, rtol=1e-12)
    for p in parts:
        assert np.allclose(p["dot"], w @ rho, rtol=1e-12)
        assert np.allclose(p["gathered"][:, 0], np.arange(K) + 0.5)
        assert bool(p["into_is_static"]) and np.array_equal(p["into"], p["gathered"])
        assert p["wrong"].shape == (K, 3) and np.array_equal(p["wrong"], p["gathered"])
        assert np.array_equal(p["async_out"], p["gathered"])
    c = K // world
    for r, p in enumerate(parts):            # sum over ranks of stats * (rank + 1) = 3 * stats for two ranks
        assert np.array_equal(p["own_stats"], 3.0 * np.arange(K * 2, dtype=np.float64).reshape(K, 2)[r * c:(r + 1) * c])


def test_rng_device_mode_bookkeeping():
    """gmmvi_b200/rng.py: draws inside a graph capture use (device counter, offset); replays advance the host counter."""
    from gmmvi_b200 import rng
    rng.set_seed(3)
    assert rng.next_subsequence() == 0 and rng.next_subsequence() == 1
    assert rng.device_counter() is None
    counter = object()
    rng.begin_device_mode(counter)
    assert rng.device_counter() == (counter, 0) and rng.device_counter() == (counter, 1)
    assert rng.end_device_mode() == 2 and rng.device_counter() is None
    assert rng.next_subsequence() == 2          # the capture itself did not consume host subsequences
    rng.advance(2)                              # one replay
    assert rng.next_subsequence() == 5


def test_default_configs_equal_the_reference_yml_files():
    """gmmvi_b200.configs against tests/golden/reference_configs.json, which tests/golden/make_reference_configs.py wrote by
    calling the reference's own configs/__init__.py on its yml files: every module letter, the codewords of the examples /
    BASELINE configurations, the shipped experiment configs, get_default_config and update_config."""
    import contextlib
    import io
    import json
    from gmmvi_b200 import configs as C
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_configs.json")) as f:
        ref = json.load(f)
    norm = lambda d: json.loads(json.dumps(d))
    with contextlib.redirect_stdout(io.StringIO()):
        for codeword, want in ref["algorithm"].items():
            assert norm(C.get_default_algorithm_config(codeword)) == want, codeword
        for name, want in ref["experiment"].items():
            assert norm(C.get_default_experiment_config(name)) == want, name
        assert norm(C.get_default_config("SAMTRON", "stm20")) == ref["merged"]["SAMTRON/stm20"]
        upd = C.update_config(C.get_default_algorithm_config("SAMTRON"),
                              {"sample_selector_config": {"desired_samples_per_component": 7}, "temperature": 0.5})
        assert norm(upd) == ref["merged"]["update"]


def test_thinning_reindexes_in_first_occurrence_order():
    """SampleDB.remove_every_nth_sample (sample_db.py:64-79): tf.unique returns first-occurrence order; with the
    mixture-based selector's draw-order mapping that differs from sorted order.  Pure index logic: runs on CPU tensors."""
    from gmmvi_b200.optimization.sample_db import SampleDB
    rng = np.random.default_rng(0)
    D, M, N = 3, 7, 40
    mapping = rng.integers(0, M, N).astype(np.int32)            # not monotone
    X = rng.standard_normal((N, D)).astype(np.float32)
    means = rng.standard_normal((M, D)).astype(np.float32)
    chols = np.tile(np.eye(D, dtype=np.float32), (M, 1, 1)) * np.arange(1, M + 1, dtype=np.float32)[:, None, None]
    db = SampleDB(D, False, True, 1000, device="cpu")
    t = torch.as_tensor
    db.samples, db.target_lnpdfs, db.target_grads = t(X), t(X[:, 0].copy()), t(X.copy())
    db.mapping, db.means, db.chols, db.inv_chols = t(mapping), t(means), t(chols), t(1.0 / np.maximum(chols, 1e-30) * (chols > 0))
    db.consts = t(np.arange(M, dtype=np.float32))
    odb = O.OracleSampleDB(D, False, True, 1000, np.float32)
    odb.samples, odb.target_lnpdfs, odb.target_grads = X, X[:, 0].copy(), X.copy()
    odb.mapping, odb.means, odb.chols, odb.inv_chols = mapping, means, chols, chols.copy()
    db.remove_every_nth_sample(2)
    odb.remove_every_nth_sample(2)
    assert np.array_equal(db.mapping.numpy(), odb.mapping)
    assert np.array_equal(db.means.numpy(), odb.means) and np.array_equal(db.chols.numpy(), odb.chols)
    assert np.array_equal(db.samples.numpy(), odb.samples)
    assert db.consts.shape[0] == odb.means.shape[0]


def test_bench_traffic_parser_reads_the_committed_ncu_summary():
    """bench.py's `roofline.traffic` comes from the newest committed ncu summary of the log-density kernel, not a literal."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    traffic, src = bench.ncu_dram_traffic()
    assert src is not None and src.startswith("profiles/r") and src.endswith("_ncu_h16t_logdens.txt")
    text = open(os.path.join(ROOT, src)).read()
    rd = float(re.search(r"dram__bytes_read\.sum\s+(\w+)\s+([\d.]+)", text).group(2))
    assert 1e8 < traffic < 1e10 and traffic > rd      # read + write, in bytes
    # every BASELINE configuration has a documented --config entry
    assert set(bench.CONFIG_DOC) == {"C1", "C2", "C3", "C3w", "C4d", "C4f"}
