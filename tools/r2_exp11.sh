for lean in 0 1; do
  DKMC_PW_LEAN=$lean python bench.py --gpus 1 --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/r2_lean_$lean.json 2> gpurun_out/r2_lean_$lean.err
  echo "lean $lean"; python tools/show_bench.py gpurun_out/r2_lean_$lean.json | cut -c1-330
done
python tools/time_dropin.py tiled_100k 10.0 2.5e-13 > gpurun_out/r02_dropin_100k.json 2> gpurun_out/r02_dropin_100k.err; cat gpurun_out/r02_dropin_100k.json | cut -c1-700; tail -2 gpurun_out/r02_dropin_100k.err
python tools/time_dropin.py tiled_1M 10.0 2.5e-13 > gpurun_out/r02_dropin_1M.json 2> gpurun_out/r02_dropin_1M.err; cat gpurun_out/r02_dropin_1M.json | cut -c1-700; tail -2 gpurun_out/r02_dropin_1M.err
