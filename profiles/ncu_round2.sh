#!/bin/bash
# ncu --set full captures of round 2 (one kernel launch each, after the plain run of the same command has exited 0).
# usage (on the GPU box, from the repo root): bash profiles/ncu_round2.sh ; summaries land in gpurun_out/ncu_r02_*.txt
mkdir -p gpurun_out
cap() {   # cap <tag> <kernel regex> <prof_kernels mode> [skip launches]
  timeout 100 python profiles/prof_kernels.py $3 1 > gpurun_out/ncu_r02_$1.plain 2>&1 || { echo "$1: plain run failed"; tail -3 gpurun_out/ncu_r02_$1.plain; return; }
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$2 -s ${4:-1} -c 1 -f -o gpurun_out/ncu_r02_$1 \
      python profiles/prof_kernels.py $3 1 > gpurun_out/ncu_r02_$1.log 2>&1
  { echo "# ncu --set full --clock-control none -k regex:$2 -s ${4:-1} -c 1 python profiles/prof_kernels.py $3 1"; grep -E "ms per launch|TFLOP|active blocks" gpurun_out/ncu_r02_$1.plain | sed 's/^/# plain run: /';
    python profiles/ncu_summary.py gpurun_out/ncu_r02_$1.ncu-rep 14; } > gpurun_out/ncu_r02_$1.txt 2>&1
  rm -f gpurun_out/ncu_r02_$1.ncu-rep
  head -12 gpurun_out/ncu_r02_$1.txt | cut -c1-150
}
cap h16t_logdens logdens_h16t logdens
cap stein_tc_full stein_tc_kernel stein_full
cap mixgrad_h16_dense mixgrad_h16 mixgrad
cap gsum2 stein_gsum2 gsum 0
cap logdens_small logdens_small small
cap logdens_diag2 logdens_diag2 diag
cap update_blocked update_full_blocked update
