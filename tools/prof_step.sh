#!/bin/bash
# One gpurun call: plain bench run, then the ncu launch list of the same command, then a full capture of the tensor-core kernels.
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/plain_bench.log 2>&1 || exit 1
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 1100 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_tc|tc_gemm_tall|tc_wgrad|score_tc|ce_bwd_tc" -s 420 -c 36 -o gpurun_out/prof_step_r1b python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
