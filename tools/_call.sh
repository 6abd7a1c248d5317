set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"
( time python bench.py --steps 10 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err ) 2> gpurun_out/r2b_bench.time; echo "bench rc=$?"
( time python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r2b_ref.json 2> gpurun_out/r2b_ref.err ) 2> gpurun_out/r2b_ref.time
tail -5 gpurun_out/r2b_pytest.log
tail -5 gpurun_out/r2b_bench.err
cat gpurun_out/r2b_bench.time gpurun_out/r2b_ref.time
nvidia-smi topo -m > gpurun_out/r2b_topo.txt 2>&1
lscpu | head -30 > gpurun_out/r2b_lscpu.txt
