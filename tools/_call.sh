mkdir -p gpurun_out
python -m pytest tests/test_visco_gpu.py tests/test_problem_gpu.py -x -q > gpurun_out/r2z_pytest_visco.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2z_pytest_visco.log
python tools/bench_visco.py --sweep > gpurun_out/r2z_visco_sweep.jsonl 2> gpurun_out/r2z_visco_sweep.err; echo "sweep rc=$?"
python - <<PY
import json
for l in open('gpurun_out/r2z_visco_sweep.jsonl'):
    try: d=json.loads(l)
    except Exception: continue
    print('d',d['dim'],'N',d['terms'],'corr' if d.get('corrected') else 'ref','ms',d['ms_median'],'frac',d['frac_of_peak'])
PY
