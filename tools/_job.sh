mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_exact_topk.py tests/test_gpu_graphs.py -x -q > gpurun_out/t14.log 2>&1; tail -3 gpurun_out/t14.log
timeout 300 python tools/perf_exact.py 2048 500002 2>&1 | grep -i "candidate sweep\|uncert" | head -5
timeout 300 python tools/perf_exact.py 4096 250001 2>&1 | grep -i "candidate sweep\|uncert" | head -5
