mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q > gpurun_out/r2A_pytest.log 2>&1 ) 2>&1 | grep real; echo "pytest rc=$?"; tail -3 gpurun_out/r2A_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2A_smoke.log 2>&1; echo "smoke rc=$?"; tail -6 gpurun_out/r2A_smoke.log
( time python bench.py --steps 20 --warmup 5 > gpurun_out/r2A_bench.json 2> gpurun_out/r2A_bench.err ) 2>&1 | grep real; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2A_bench.json').read().strip().splitlines()[-1])
print('ms/step',round(d['ms_per_step'],2),'cheb',round(d['roofline_cheb_step']['avg_launch_ms']*1e3,1),round(d['roofline_cheb_step']['frac'],3),'apply',round(d['roofline_apply']['avg_launch_ms']*1e3,1),round(d['roofline_apply']['frac'],3),'visco',round(d['roofline_visco']['frac'],3),'parity',d['parity_check']['ok'],'e2e',round(d['e2e']['ms_per_step'],1),d['e2e'].get('raw_concurrent_d2h_GBs'))
for k,v in d['other_configs'].items():
    if 'C5' in k: print(k, {n:x['frac'] for n,x in v.items()} if 'error' not in v else v)
    else: print(k, v.get('ms_per_step'), v.get('pcg_its_per_step'), v.get('mech_pcg_its_per_step'), v.get('equilibrium_ok'), v.get('error'))
print('c4', d['c4'].get('ms_per_step'), d['c4'].get('setup_s'))
print('cpu', d['cpu_baseline'])
PY
( time python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2A_ref.json 2> gpurun_out/r2A_ref.err ) 2>&1 | grep real; echo "ref rc=$?"; cut -c1-400 gpurun_out/r2A_ref.json
