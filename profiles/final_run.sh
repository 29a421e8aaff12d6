#!/bin/bash
# One GPU-box call that refreshes the evidence under profiles/: GPU tests, the bench line, the ncu launch list of the
# same command, the opt-in blocked diagonal-block kernel of the MORE estimator, the other BASELINE configurations and
# the per-kernel budget of the C5 iteration.  usage (from the repo root): bash profiles/final_run.sh
mkdir -p gpurun_out
timeout 120 python -m pytest tests -m gpu -x -q > gpurun_out/f_tests.log 2>&1; tail -2 gpurun_out/f_tests.log
timeout 200 python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; cut -c1-160 gpurun_out/f_bench.json
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/f_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-dense > gpurun_out/f_ncu.log 2>&1; tail -1 gpurun_out/f_ncu.log | cut -c1-120
GMMVI_B200_MORE_POTRF=blocked timeout 80 python -m pytest tests -m gpu -x -q -k "more" > gpurun_out/f_tests_potrf.log 2>&1; tail -2 gpurun_out/f_tests_potrf.log
GMMVI_B200_MORE_POTRF=blocked timeout 60 python profiles/bench_configs.py 10 C3 --kernels 2>&1 | grep " ms" | head -4 > gpurun_out/f_c3_potrf.log; cat gpurun_out/f_c3_potrf.log
timeout 150 python profiles/bench_configs.py 20 > gpurun_out/f_configs.log 2>&1; grep "ms/iter" gpurun_out/f_configs.log
timeout 60 python profiles/prof_iter.py > gpurun_out/f_prof_iter.log 2>&1; grep -v Warn gpurun_out/f_prof_iter.log | head -16
