mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/gputests_r2c.log 2>&1; tail -4 gpurun_out/gputests_r2c.log
timeout 400 python bench.py --no-cpu > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; tail -c 600 gpurun_out/bench_r2c.json
