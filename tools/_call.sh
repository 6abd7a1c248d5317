mkdir -p gpurun_out
python examples/plate_mechanics.py 10 2>&1 | tail -12
python tools/mech_probe.py 160 160 8 2>&1 | head -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('its', d['pcg_its'], 'ms_mech', d['ms_mechanics'])"
python -m pytest tests/test_mechanics_gpu.py -x -q 2>&1 | tail -2
