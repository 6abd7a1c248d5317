# Full single-GPU validation on a B200 box (run through gpurun from the repository root):
#   pytest -m gpu, smoke(), the default bench line with the driver's arguments, the reference arm.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/validate_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/validate_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/validate_smoke.log 2>&1; echo "smoke rc=$?"; tail -7 gpurun_out/validate_smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/validate_bench.json 2> gpurun_out/validate_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/validate_ref.json 2> gpurun_out/validate_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/validate_bench.json').read().strip().splitlines()[-1])
print('ms/step', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['ms_per_step'], 1), 'roofline', round(d['roofline']['frac'], 3),
      'parity', d['parity_check']['ok'])
for k, v in d['other_configs'].items():
    print(' ', k, {n: x['frac'] for n, x in v.items()} if 'C5' in k else (v.get('ms_per_step'), v.get('error')))
print('  c4', d['c4'].get('ms_per_step'))
PY
