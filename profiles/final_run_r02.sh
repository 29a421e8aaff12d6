#!/bin/bash
# One GPU-box call that refreshes the round-2 evidence under profiles/ (copy from gpurun_out/ afterwards):
# GPU tests, the C5 bench line, the ncu launch list of the same command, the C1-C4 config lines with CPU baselines,
# the MORE trailing-update capture.   usage (repo root): bash profiles/final_run_r02.sh
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02_gpu_tests.log 2>&1; tail -2 gpurun_out/r02_gpu_tests.log
timeout 300 python bench.py > gpurun_out/r02_bench_c5.json 2> gpurun_out/r02_bench_c5.err; cut -c1-200 gpurun_out/r02_bench_c5.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-dense --no-graph > gpurun_out/r02_ncu_launches.log 2>&1; tail -1 gpurun_out/r02_ncu_launches.log | cut -c1-120
for c in C1 C2 C4d C4f C3 C3w; do
  extra=""; [ $c = C3w ] && extra="--no-cpu"
  timeout 400 python bench.py --config $c --steps 20 $extra > gpurun_out/r02_bench_$c.json 2> gpurun_out/r02_bench_$c.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_$c.json").read().strip().splitlines()[-1])
    cb = d.get("cpu_baseline") or {}
    print("$c", "ms/iter %.3f" % d["ms_per_step"], "e2e it/s %.1f" % d["e2e"]["value"], "K", d["config"]["components"],
          "captures", d["config"]["graph_captures_in_timed_region"], "cpu s/iter", cb.get("sec_per_iter"), cb.get("error"))
except Exception as e:
    print("$c failed", e)
PY
done
timeout 200 ncu --set full --clock-control none --import-source on -k regex:tc_bgemm_kernel -s 8 -c 1 -f -o gpurun_out/r02_ncu_more_trailing \
    python bench.py --config C3 --steps 1 --no-cpu --no-graph > gpurun_out/r02_ncu_more_trailing.log 2>&1
python profiles/ncu_summary.py gpurun_out/r02_ncu_more_trailing.ncu-rep 12 > gpurun_out/r02_ncu_more_trailing.txt 2>&1; rm -f gpurun_out/r02_ncu_more_trailing.ncu-rep
head -9 gpurun_out/r02_ncu_more_trailing.txt | cut -c1-140
