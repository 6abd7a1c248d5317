mkdir -p gpurun_out
B="python bench.py --steps 6 --warmup 3 --no-other-configs --no-cpu-baseline --no-parity"
for m in ce ce2 sm:8 sm:16 sm:32 sm:64; do
  SG_MIRROR_MODE=$m $B > gpurun_out/r2w_e2e_$m.json 2> gpurun_out/r2w_e2e_$m.err; echo "$m rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2w_e2e_$m.json').read().strip().splitlines()[-1])
print('$m', 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'agg', round(d['e2e']['aggregate_d2h_GBs'],1), 'raw', round(d['e2e']['raw_concurrent_d2h_GBs'],1))
PY
done
