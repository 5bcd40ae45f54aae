mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gputests_r2i.log 2>&1; tail -3 gpurun_out/gputests_r2i.log
timeout 400 python bench.py --no-cpu --no-eval > gpurun_out/bench_r2i.json 2> gpurun_out/bench_r2i.err; python tools/show_bench.py gpurun_out/bench_r2i.json | grep "^value\|^ms_per\|^e2e" | cut -c1-200
ASME_B200_SIDE_WGRAD=0 timeout 400 python bench.py --no-cpu --no-eval > gpurun_out/bench_r2i_noside.json 2> /dev/null; python tools/show_bench.py gpurun_out/bench_r2i_noside.json | grep "^value\|^ms_per" | cut -c1-200
