python -m pytest tests -m gpu -x -q > gpurun_out/t17.log 2>&1; tail -5 gpurun_out/t17.log
python bench.py > gpurun_out/bench_r2k_1gpu.json 2> gpurun_out/bench_r2k_1gpu.err; tail -c 1500 gpurun_out/bench_r2k_1gpu.json | head -c 600; echo; tail -3 gpurun_out/bench_r2k_1gpu.err
