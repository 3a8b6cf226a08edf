for mg in 0 30 100; do
  DKMC_RESTART_MARGIN=$mg DKMC_SOLVE_TRACE=1 python bench.py --gpus 1 --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/r2e_margin_$mg.json 2> gpurun_out/r2e_margin_$mg.err
  echo "margin $mg"; python tools/show_bench.py gpurun_out/r2e_margin_$mg.json | cut -c1-330
  python - <<PY
import json
for l in open('gpurun_out/r2e_margin_$mg.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['per_step']['cg_iterations'])
PY
  grep "dkmc solve" gpurun_out/r2e_margin_$mg.err | sed -n 41,52p
done
