mkdir -p gpurun_out
python -m pytest tests/test_mechanics_gpu.py -x -q 2>&1 | tail -12
python tools/mech_probe.py 48 48 8 160 160 8 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print(d['cells'], d['physics'], 'its', d['pcg_its'], 'ms_mech', d['ms_mechanics'], 'ms_step', d['ms_per_step'])"
