# round-2 experiment: tests, N=1 bench, pairwise alone numbers
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2b_gputest.log
tail -3 gpurun_out/r2b_gputest.log
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err
python tools/show_bench.py gpurun_out/r2b_bench_n1.json
python - <<'PY'
import json
for l in open('gpurun_out/r2b_bench_n1.json'):
    if l.startswith('{'):
        d=json.loads(l); print(json.dumps(d['rooflines']['pairwise'])); print(d['per_step']['cg_iterations'])
PY
