mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_exact_topk.py tests/test_gpu_tc.py -x -q > gpurun_out/t9.log 2>&1; tail -3 gpurun_out/t9.log
timeout 300 python tools/perf_exact.py 8192 125001 2>&1 | grep -i "candidate sweep\|uncert\|whole" | head -8
timeout 300 python tools/perf_exact.py 2048 500002 2>&1 | grep -i "candidate sweep\|uncert\|whole" | head -4
timeout 300 python tools/perf_exact.py 1024 1000003 2>&1 | grep -i "candidate sweep\|uncert\|whole" | head -4
