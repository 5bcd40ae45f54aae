mkdir -p gpurun_out
timeout 400 python bench.py --no-cpu > gpurun_out/bench_r2h.json 2> gpurun_out/bench_r2h.err; tail -3 gpurun_out/bench_r2h.err | cut -c1-300; python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_r2h.json').read().strip().splitlines()[-1])
e=d['eval']
print('eval', e['value'], e['ms_per_step'], e.get('batch_4096'))
P
