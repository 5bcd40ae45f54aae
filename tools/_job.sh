python -m pytest tests -m gpu -x -q > gpurun_out/t19.log 2>&1; tail -2 gpurun_out/t19.log
python bench.py > gpurun_out/bench_r2n_1gpu.json 2> gpurun_out/bench_r2n_1gpu.err; tail -1 gpurun_out/bench_r2n_1gpu.err | cut -c1-200
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r2n_1gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'])
print(d['eval']['value'], d['eval']['ms_per_step'], d['eval']['kernel_ms_per_step'], d['eval']['e2e']['value'], d['eval']['e2e']['ms_per_step'], d['eval']['recall@10'], d['eval']['checks'])
PY
