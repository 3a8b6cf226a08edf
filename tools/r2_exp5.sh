python -m pytest tests -m gpu -x -q 2>&1 | tail -12
for w in device_7.5nm crossbar_2x2; do
  python bench.py --gpus 1 --steps 10 --warmup 5 --no-cpu-baseline --workload $w > gpurun_out/r2_bench_${w}_n1.json 2> gpurun_out/r2_bench_${w}_n1.err
  python tools/show_bench.py gpurun_out/r2_bench_${w}_n1.json | cut -c1-420
  python - <<PY
import json
for l in open('gpurun_out/r2_bench_${w}_n1.json'):
    if l.startswith('{'):
        d=json.loads(l); print(json.dumps(d['rooflines']['spmv'])[:300]); print(d['per_step']['cg_iterations'], d['config'])
PY
  tail -3 gpurun_out/r2_bench_${w}_n1.err
done
