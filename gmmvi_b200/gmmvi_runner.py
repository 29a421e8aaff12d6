"""GmmviRunner (mirror of gmmvi_runner.py:24-200): seeds the generators, builds target + model + GMMVI from one
config dict, times `train_iter`, returns the reference's metric dict and dumps the GMM as .npz.
matplotlib figures are out of scope."""
from __future__ import annotations

import os
from time import time

import numpy as np
import torch

from . import rng
from .experiments.setup_experiment import init_experiment
from .optimization.gmmvi import GMMVI


class GmmviRunner:
    def __init__(self, config, log_metrics_interval, device="cuda", use_cuda_graph=True):
        """gmmvi_runner.py:34-61.  `use_cuda_graph` (gmmvi_runner_config, default on): run the iteration as one captured
        CUDA graph whenever the configuration allows it -- what the reference's tf.function around train_iter
        (optimization/gmmvi.py:99-103) does for TensorFlow; results are identical to op-by-op iterations."""
        if "seed" not in config.keys():
            config["seed"] = config["start_seed"]
        rng.set_seed(config["seed"])
        self.wall_times = []
        self.config = config
        self.log_metrics_interval = log_metrics_interval
        target_distribution, initial_model = init_experiment(self.config, device=device)
        self.gmmvi = GMMVI.build_from_config(self.config, target_distribution, initial_model)
        if use_cuda_graph:
            try:
                self.gmmvi.enable_cuda_graph()
            except NotImplementedError:          # sample reuse / mixture-based selector: op-by-op iterations
                pass
        if "mmd_evaluation_config" in config.keys():
            # gmmvi_runner.py:45-54; `sample_dir` is looked up next to this file like in the reference, then as given
            from .experiments.evaluation.mmd import MMD
            rel = config["mmd_evaluation_config"]["sample_dir"]
            path = os.path.join(os.path.dirname(os.path.realpath(__file__)), rel)
            samples = np.load(path if os.path.exists(path) else rel)
            self.mmd = MMD(samples, config["mmd_evaluation_config"]["alpha"], device=device)
        else:
            self.mmd = None
        if "dump_gmm_path" not in self.config:
            self.dump_gmms = False
        else:
            self.dump_gmms = True
            self.dump_gmm_path = os.path.join(self.config["dump_gmm_path"], str(time()))
            os.makedirs(self.dump_gmm_path, exist_ok=True)

    @staticmethod
    def build_from_config(config: dict, device="cuda"):
        """gmmvi_runner.py:63-81."""
        return GmmviRunner(config=config, device=device, **config["gmmvi_runner_config"])

    def get_samples_and_entropy(self, num_samples):
        """gmmvi_runner.py:83-100."""
        test_samples = self.gmmvi.model.sample(num_samples)[0]
        entropy = -torch.mean(self.gmmvi.model.log_density(test_samples))
        return test_samples, entropy

    def get_cheap_metrics(self):
        """gmmvi_runner.py:102-117."""
        g = self.gmmvi
        return {"num_samples": g.sample_db.num_samples_written,
                "num_components": g.model.num_components,
                "max_weight": float(torch.max(g.model.weights).item()),
                "num_db_samples": int(g.sample_db.samples.shape[0]),
                "num_db_components": int(g.sample_db.means.shape[0])}

    def get_expensive_metrics(self):
        """gmmvi_runner.py:119-144."""
        g = self.gmmvi
        test_samples, entropy = self.get_samples_and_entropy(2000)
        mean_reward = torch.mean(g.sample_selector.target_uld(test_samples))
        elbo = mean_reward + g.temperature * entropy
        out = {"-elbo": float(-elbo.item()), "entropy": float(entropy.item()),
               "target_density": float(mean_reward.item()), "algo_time": float(np.sum(self.wall_times))}
        out.update(g.sample_selector.target_distribution.expensive_metrics(g.model, test_samples))
        if self.mmd is not None:
            out.update({"MMD:": float(self.mmd.compute_MMD(test_samples).item())})      # key as in the reference (:142)
        return out

    def iterate_and_log(self, n: int) -> dict:
        """gmmvi_runner.py:146-175 (wall-clock includes a device synchronisation so that it measures the work)."""
        output_dict = {}
        ts1 = time()
        self.gmmvi.train_iter()
        torch.cuda.synchronize()
        ts2 = time()
        output_dict.update({"walltime": ts2 - ts1})
        self.wall_times.append(ts2 - ts1)
        output_dict.update(self.get_cheap_metrics())
        if n % self.log_metrics_interval == 0:
            eval_dict = self.get_expensive_metrics()
            print("Checkpoint {:3d} | FEVALS: {:10d} | avg. sample logpdf: {:05.05f} | ELBO: {:05.05f}".format(
                n, output_dict["num_samples"], eval_dict["target_density"], -eval_dict["-elbo"]))
            print(f"{self.gmmvi.model.num_components} components\n")
            output_dict.update(eval_dict)
        return output_dict

    def _dump(self, path):
        m = self.gmmvi.model
        np.savez(path, weights=m.weights.cpu().numpy(), means=m.means.cpu().numpy(), covs=m.covs.cpu().numpy(),
                 timestamps=time(), fevals=self.gmmvi.sample_db.num_samples_written)

    def log_to_disk(self, n: int):
        """gmmvi_runner.py:177-190."""
        if self.dump_gmms and (n < 100 or n % 50 == 0):
            self._dump(self.dump_gmm_path + "/gmm_dump_" + str("%01d" % n) + ".npz")

    def finalize(self):
        """gmmvi_runner.py:192-200."""
        if self.dump_gmms:
            self._dump(self.dump_gmm_path + "/final_gmm_dump.npz")
