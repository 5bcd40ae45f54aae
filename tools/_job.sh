mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_models_bf16.py tests/test_gpu_c1.py -x -q > gpurun_out/t7.log 2>&1; tail -15 gpurun_out/t7.log
