# Multi-GPU validation (gpurun --gpus N -- 'bash tools/gpu_validate_multi.sh N'): partitioned-vs-single parity of four
# space/dimension cases, then the bench line at N GPUs (C3 weak scaling, C4 strong scaling, partition_check).
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 tests/multigpu_check.py > gpurun_out/validate_mg$N.log 2>&1; echo "multigpu_check rc=$?"; grep -E "relerr|MULTIGPU_CHECK" gpurun_out/validate_mg$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/validate_bench$N.json 2> gpurun_out/validate_bench$N.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open('gpurun_out/validate_bench$N.json').read().strip().splitlines()[-1])
print('N=$N ms/step', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['ms_per_step'], 1), 'c4', d.get('c4', {}).get('ms_per_step'),
      'partition_check', d.get('partition_check', {}).get('ok'))
PY
