python bench.py > gpurun_out/bench_r2m_1gpu.json 2> gpurun_out/bench_r2m_1gpu.err; grep "evaluation end to end" gpurun_out/bench_r2m_1gpu.err; tail -2 gpurun_out/bench_r2m_1gpu.err | cut -c1-200
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r2m_1gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'])
print(d['eval']['value'], d['eval']['ms_per_step'], d['eval']['e2e']['value'], d['eval']['e2e']['ms_per_step'], d['eval']['recall@10'], d['eval']['checks'])
PY
