mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 8 --no-cpu > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err; tail -c 700 gpurun_out/bench_8gpu.json
