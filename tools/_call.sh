mkdir -p gpurun_out
( time python -m pytest tests/test_fullsize_gpu.py -x -q > gpurun_out/r2B_pytest_full.log 2>&1 ) 2>&1 | grep real; echo "pytest rc=$?"; tail -25 gpurun_out/r2B_pytest_full.log
