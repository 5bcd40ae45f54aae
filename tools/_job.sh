mkdir -p gpurun_out
timeout 400 python bench.py --no-cpu > gpurun_out/bench_r2e.json 2> gpurun_out/bench_r2e.err; python tools/show_bench.py gpurun_out/bench_r2e.json | grep -v "^config\|^implementation" | cut -c1-600 | head -30
