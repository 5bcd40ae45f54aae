mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_tc_attention.py -x -q > gpurun_out/t_2gpu_r2j.log 2>&1; tail -3 gpurun_out/t_2gpu_r2j.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --no-cpu > gpurun_out/bench_r2j_2gpu.json 2> gpurun_out/bench_r2j_2gpu.err; tail -c 600 gpurun_out/bench_r2j_2gpu.json
