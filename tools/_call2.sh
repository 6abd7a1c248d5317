mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --no-other-configs > gpurun_out/r2E_bench2.json 2> gpurun_out/r2E_bench2.err; echo "bench2 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2E_bench2.json').read().strip().splitlines()[-1])
print('N2 ms/step', round(d['ms_per_step'],2), 'partition_check', d.get('partition_check'))
PY
tail -3 gpurun_out/r2E_bench2.err
