#!/bin/bash
# One gpurun call: plain run, then the ncu launch list of ONE training step + ONE evaluation step (bench.py --profile-region
# brackets them in cudaProfilerStart/Stop), then a full capture of the hot kernels of the same region.
# Keep the captures small: gpurun brings back at most 64 MiB, and every profiled launch costs GPU-seconds.
TAG=${1:-r1d}
CMD="python bench.py --profile-region --no-cpu"
$CMD > gpurun_out/plain_profile_region.log 2>&1 || exit 1
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launch.log 2>&1
timeout 500 ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,sm__cycles_elapsed.max,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size,launch__waves_per_multiprocessor --section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section LaunchStats --section Occupancy --section SchedulerStats --clock-control none -k regex:"attn_tc|ffn_fused|tc_gemm_persist|tc_wgrad|score_tc_kernel|ce_bwd_tc|layernorm|embed_|attn_row|embgrad|dropout_cast|adam" -c 90 -o gpurun_out/prof_step_${TAG} $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out/
