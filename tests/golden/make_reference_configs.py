"""tests/golden/reference_configs.json: the default configuration dictionaries as the reference's own
configs/__init__.py builds them from its yml files (every module letter, the codewords used by the examples / BASELINE
configurations, the experiment configs this repo ships) - tests/test_host_cpu.py compares gmmvi_b200.configs with them.
Usage (needs /root/reference):  python tests/golden/make_reference_configs.py"""
import contextlib
import io
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "tf_shim"))          # mergedeep stand-in
sys.path.insert(0, "/root/reference/src")
import importlib.util  # noqa: E402

spec = importlib.util.spec_from_file_location("ref_configs", "/root/reference/src/gmmvi/configs/__init__.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

LETTERS = "ZSAEPMIYTFDRUOXGN"
CODEWORDS = ["SAMTRON", "ZEPTFOX", "SEMYFUX", "SAMYROX", "ZAMTRUX", "SEPIDUG", "ZEMTRON"]
EXPERIMENTS = ["stm20", "stm300", "gmm20", "gmm100", "planar_robot_4"]

out = {"algorithm": {}, "experiment": {}, "merged": {}}
with contextlib.redirect_stdout(io.StringIO()):
    for c in list(LETTERS) + CODEWORDS:
        out["algorithm"][c] = ref.get_default_algorithm_config(c)
    for e in EXPERIMENTS:
        out["experiment"][e] = ref.get_default_experiment_config(e)
    out["merged"]["SAMTRON/stm20"] = ref.get_default_config("SAMTRON", "stm20")
    out["merged"]["update"] = ref.update_config(ref.get_default_algorithm_config("SAMTRON"),
                                                {"sample_selector_config": {"desired_samples_per_component": 7},
                                                 "temperature": 0.5})
with open(os.path.join(HERE, "reference_configs.json"), "w") as f:
    json.dump(out, f, indent=1, sort_keys=True)
print("wrote reference_configs.json")
