mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gputests_r2j.log 2>&1; tail -2 gpurun_out/gputests_r2j.log
timeout 400 python bench.py > gpurun_out/bench_r2j.json 2> gpurun_out/bench_r2j.err; tail -c 700 gpurun_out/bench_r2j.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r2j.json 2> gpurun_out/bench_ref_r2j.err; tail -c 200 gpurun_out/bench_ref_r2j.json
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
bash tools/prof_step.sh r2j
