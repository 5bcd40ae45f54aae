mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gputests_r2g.log 2>&1; tail -3 gpurun_out/gputests_r2g.log
timeout 400 python bench.py --no-cpu > gpurun_out/bench_r2g.json 2> gpurun_out/bench_r2g.err; python tools/show_bench.py gpurun_out/bench_r2g.json | grep "^eval\|^value\|^ms_per" | cut -c1-300
