for cfg in 1e-12,1e-6 1e-8,1e-6 1e-6,1e-6 1e-4,1e-6 1e-8,1e-5 1e-6,1e-5 1e-6,1e-4 1e-6,1e-7; do
  rt=${cfg%,*}; ft=${cfg#*,}
  DKMC_REL_TOL=$rt DKMC_REFINE_TOL=$ft DKMC_SOLVE_TRACE=1 python bench.py --gpus 1 --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/r2f_tol_$cfg.json 2> gpurun_out/r2f_tol_$cfg.err
  echo "rel_tol $rt refine_tol $ft"; python tools/show_bench.py gpurun_out/r2f_tol_$cfg.json | cut -c60-230
  python - <<PY
import json
for l in open('gpurun_out/r2f_tol_$cfg.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['per_step']['cg_iterations'], d['per_step']['events'])
PY
  grep "dkmc solve" gpurun_out/r2f_tol_$cfg.err | tail -5 | tr '\n' ';'; echo
done
