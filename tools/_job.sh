mkdir -p gpurun_out
timeout 300 python tools/perf_exact.py 4096 250001 2>&1 | grep -i "candidate sweep\|uncert" | head -5
timeout 600 python -m pytest tests/test_gpu_exact_topk.py -x -q > gpurun_out/t15.log 2>&1; tail -2 gpurun_out/t15.log
