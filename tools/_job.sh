mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gputests_r2d.log 2>&1; tail -3 gpurun_out/gputests_r2d.log
timeout 400 python bench.py > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err; tail -c 400 gpurun_out/bench_r2d.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r2d.json 2> gpurun_out/bench_ref_r2d.err; tail -c 300 gpurun_out/bench_ref_r2d.json
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
bash tools/prof_step.sh r2d
