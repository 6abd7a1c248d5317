mkdir -p gpurun_out
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2F_bench8.json 2> gpurun_out/r2F_bench8.err ) 2>&1 | grep real; echo "bench8 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2F_bench8.json').read().strip().splitlines()[-1])
    print('N8 ms/step', round(d['ms_per_step'],2), 'value', d['value'], 'e2e', round(d['e2e']['ms_per_step'],1), 'c4', d.get('c4',{}).get('ms_per_step'), d.get('c4',{}).get('error'))
    print('partition_check', d.get('partition_check'))
except Exception as e:
    print('parse failed', e)
PY
tail -3 gpurun_out/r2F_bench8.err
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2F_bench4.json 2> gpurun_out/r2F_bench4.err ) 2>&1 | grep real; echo "bench4 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2F_bench4.json').read().strip().splitlines()[-1])
    print('N4 ms/step', round(d['ms_per_step'],2), 'value', d['value'], 'e2e', round(d['e2e']['ms_per_step'],1), 'c4', d.get('c4',{}).get('ms_per_step'), 'pcheck ok', d.get('partition_check',{}).get('ok'))
except Exception as e:
    print('parse failed', e)
PY
