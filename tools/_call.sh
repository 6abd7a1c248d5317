set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2e_pytest.log
for sp in 1 0; do
SG_SERPENTINE=$sp python bench.py --steps 10 --warmup 3 --no-other-configs --no-cpu-baseline --no-parity > gpurun_out/r2e_bench_sp$sp.json 2> gpurun_out/r2e_bench_sp$sp.err; echo "bench rc=$?"
done
python -c "
import json
for f in ('gpurun_out/r2e_bench_sp1.json','gpurun_out/r2e_bench_sp0.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f,'ms/step',d['ms_per_step'],'cheb',d['roofline_cheb_step']['avg_launch_ms'],d['roofline_cheb_step']['frac'],'apply',d['roofline_apply']['avg_launch_ms'],d['roofline_apply']['frac'],'visco',d['roofline_visco']['avg_launch_ms'])
"
