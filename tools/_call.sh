mkdir -p gpurun_out
L=fem_glass_tempering_b200/lib/libsurroglas_b200.so
for v in head new head new; do
cp tools/_ab/lib_$v.so $L
echo "== $v"
python bench.py --workload C2_plate2d_CG2_1M_qp --steps 20 --warmup 5 --no-other-configs --no-cpu-baseline --no-parity 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C2 ms/step', round(d['ms_per_step'],3), 'its', d['config'].get('pcg_its_per_step'))"
done
