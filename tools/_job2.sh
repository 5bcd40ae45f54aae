mkdir -p gpurun_out
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --no-cpu > gpurun_out/bench_r2o_2gpu.json 2> gpurun_out/bench_r2o_2gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r2o_2gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['eval']['value'], d['eval']['ms_per_step'], d['eval']['e2e']['value'], d['eval']['checks'])
PY
