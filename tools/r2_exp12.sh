for cfg in "0,-1,0,-1" "0,-1,0,3" "0,3,0,3"; do
  DKMC_PCG_CPS=$cfg python bench.py --gpus 1 --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/r2_v32.json 2> gpurun_out/r2_v32.err
  echo "cfg $cfg"; python tools/show_bench.py gpurun_out/r2_v32.json | cut -c1-330
done
