python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python bench.py --ramp --gpus 1 --steps 10 --warmup 3 > gpurun_out/r2_ramp_n1.json 2> gpurun_out/r2_ramp_n1.err
python tools/show_bench.py gpurun_out/r2_ramp_n1.json | cut -c1-600; tail -3 gpurun_out/r2_ramp_n1.err
python - <<'PY'
import json
for l in open('gpurun_out/r2_ramp_n1.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['per_step'], d['config']['sites'])
PY
