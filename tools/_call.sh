set -x
mkdir -p gpurun_out
python tools/parity_probe.py > gpurun_out/r2a_parity_probe.jsonl 2> gpurun_out/r2a_parity_probe.err
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/probe_table_read tools/probe_table_read.cu && /tmp/probe_table_read > gpurun_out/r2a_probe_table_read.txt 2>&1
python bench.py --workload C4_plate3d_CG2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_c4.json 2> gpurun_out/r2a_c4.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2a_c4_launches.csv python bench.py --workload C4_plate3d_CG2 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2a_c4_ncu.log 2>&1
python bench.py --workload C2_plate2d_CG2_1M_qp --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_c2.json 2> gpurun_out/r2a_c2.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/r2a_c2_launches.csv python bench.py --workload C2_plate2d_CG2_1M_qp --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2a_c2_ncu.log 2>&1
tail -3 gpurun_out/r2a_parity_probe.jsonl
cat gpurun_out/r2a_probe_table_read.txt
