mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2u_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2u_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2u_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r2u_smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err; echo "bench rc=$?"
CMD="python bench.py --steps 2 --warmup 1 --no-other-configs --no-cpu-baseline --no-parity"
$CMD > gpurun_out/r2u_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2u_c3_launches.csv $CMD > gpurun_out/r2u_ncu1.log 2>&1
$CMD > gpurun_out/r2u_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k "regex:dg_cheb_step|dg_class_apply|k_update_xr_cheb|k_axpy_p" -s 40 -c 8 -o gpurun_out/r2u_prof $CMD > gpurun_out/r2u_ncu2.log 2>&1
ls -la gpurun_out/r2u_prof* 
python tools/launch_list_summary.py gpurun_out/r2u_c3_launches.csv | head -14
python -c "
import json
d=json.loads(open('gpurun_out/r2u_bench.json').read().strip().splitlines()[-1])
print('ms/step',round(d['ms_per_step'],2),'cheb',round(d['roofline_cheb_step']['avg_launch_ms']*1e3,1),round(d['roofline_cheb_step']['frac'],3),'apply',round(d['roofline_apply']['avg_launch_ms']*1e3,1),round(d['roofline_apply']['frac'],3),'visco',round(d['roofline_visco']['frac'],3),'parity',d['parity_check']['ok'],'e2e',round(d['e2e']['ms_per_step'],1),d['e2e'].get('raw_concurrent_d2h_GBs'))
for k,v in d['other_configs'].items(): print(k, v.get('ms_per_step'), v.get('pcg_its_per_step'), v.get('error'))
print('c4', d['c4'].get('ms_per_step'), d['c4'].get('setup_s'))
"
