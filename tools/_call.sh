mkdir -p gpurun_out
for v in "A:" "B:SG_FORCE_HALO_VARIANT=1"; do
name=${v%%:*}; envs=${v#*:}
env $envs python bench.py --steps 10 --warmup 3 --no-other-configs --no-cpu-baseline --no-parity > gpurun_out/r2m_bench1_$name.json 2> gpurun_out/r2m_bench1_$name.err; echo "bench1 $name rc=$?"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multigpu_check.py > gpurun_out/r2m_mgcheck.log 2>&1; echo "mgcheck rc=$?"
tail -1 gpurun_out/r2m_mgcheck.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 10 --warmup 3 --no-other-configs > gpurun_out/r2m_bench2.json 2> gpurun_out/r2m_bench2.err; echo "bench2 rc=$?"
python -c "
import json
for f in ('gpurun_out/r2m_bench1_A.json','gpurun_out/r2m_bench1_B.json','gpurun_out/r2m_bench2.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f,'ms/step',round(d['ms_per_step'],2),'cheb',round(d['roofline_cheb_step']['avg_launch_ms']*1e3,1),'apply',round(d['roofline_apply']['avg_launch_ms']*1e3,1),'launches',d['gpu_launches'])
"
