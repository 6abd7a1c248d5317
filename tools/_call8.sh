mkdir -p gpurun_out
nvidia-smi -L | wc -l
( time timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2x_bench8.json 2> gpurun_out/r2x_bench8.err ) 2>&1 | tail -3; echo "bench8 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2x_bench8.json').read().strip().splitlines()[-1])
    print('N8 ms/step', round(d['ms_per_step'],2), 'value', d['value'], 'e2e', round(d['e2e']['ms_per_step'],1), 'c4', d.get('c4',{}).get('ms_per_step'), d.get('c4',{}).get('error'))
    print('transport', d['config'].get('transport'))
except Exception as e:
    print('parse failed', e)
PY
tail -5 gpurun_out/r2x_bench8.err
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 tests/multigpu_check.py > gpurun_out/r2x_mg8.log 2>&1 ) 2>&1 | tail -3; echo "mg8 rc=$?"; tail -12 gpurun_out/r2x_mg8.log
