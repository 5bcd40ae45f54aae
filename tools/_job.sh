./tools/microbench/tmem_read_bw
python -m pytest tests/test_gpu_tc_attention.py -x -q > gpurun_out/t16.log 2>&1; tail -3 gpurun_out/t16.log
