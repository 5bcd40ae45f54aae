python -m pytest tests/test_gpu_graphs.py tests/test_gpu_c1.py -x -q > gpurun_out/t20.log 2>&1; tail -2 gpurun_out/t20.log
python bench.py --no-cpu > gpurun_out/bench_r2o_1gpu.json 2> gpurun_out/bench_r2o_1gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r2o_1gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'])
print(d['eval']['value'], d['eval']['ms_per_step'], d['eval']['e2e']['value'], d['eval']['checks'])
PY
