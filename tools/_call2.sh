mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 tests/multigpu_check.py > gpurun_out/r2D_mg2.log 2>&1; echo "mg2 rc=$?"; tail -6 gpurun_out/r2D_mg2.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2D_bench2.json 2> gpurun_out/r2D_bench2.err; echo "bench2 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2D_bench2.json').read().strip().splitlines()[-1])
print('N2 ms/step', round(d['ms_per_step'],2), 'value', d['value'], 'e2e', round(d['e2e']['ms_per_step'],1), 'c4', d.get('c4',{}).get('ms_per_step'))
PY
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | cut -c1-300
