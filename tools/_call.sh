mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2C_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2C_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
