mkdir -p gpurun_out
SG_PERSIST_TIMING=1 python bench.py --workload C2_plate2d_CG2_1M_qp --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/r2s_C2t.err
grep "persistent PCG" gpurun_out/r2s_C2t.err | tail -1
python bench.py --workload C2_plate2d_CG2_1M_qp --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2s_C2.json 2> gpurun_out/r2s_C2.err; echo "rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r2s_C2.json').read().strip().splitlines()[-1])
print('ms/step',round(d['ms_per_step'],3),'its',d['config']['pcg_its_per_step'],'launches',d['gpu_launches'])
"
