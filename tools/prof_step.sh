#!/bin/bash
# Profiles one gpurun call: plain bench run, then the ncu launch list of the same command, then a full capture of the named kernels.
set -x
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/plain_bench.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"attn_tc|tc_gemm_tall|tc_wgrad|score_tc" -s 60 -c 24 -o gpurun_out/prof_step_r1 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
