mkdir -p gpurun_out
for i in 1 2; do
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --no-cpu > gpurun_out/bench_r2k_2gpu_$i.json 2> gpurun_out/bench_r2k_2gpu_$i.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r2k_2gpu_$i.json').read().strip().splitlines()[-1])
print($i, d['value'], d['ms_per_step'], d['eval']['value'], d['eval']['ms_per_step'], d['eval'].get('kernel_ms_per_step'))
PY
done
