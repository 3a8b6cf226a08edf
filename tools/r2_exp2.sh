# N=1: pairwise share sweep (blocks per SM, threads per block)
for sh in 3,128 3,96 2,128 4,64 2,160 5,64 2,96; do
  DKMC_PW_SHARE=$sh python bench.py --gpus 1 --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/r2c_share_$sh.json 2> gpurun_out/r2c_share_$sh.err
  echo "share $sh"; python tools/show_bench.py gpurun_out/r2c_share_$sh.json | cut -c1-330
done
