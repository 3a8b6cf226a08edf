n=$1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29577 bench.py --ramp --gpus $n --steps 10 --warmup 3 > gpurun_out/r2_ramp_n$n.json 2> gpurun_out/r2_ramp_n$n.err
python tools/show_bench.py gpurun_out/r2_ramp_n$n.json | cut -c1-500; grep -i "error" gpurun_out/r2_ramp_n$n.err | tail -3
if [ "$n" = "4" ]; then
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29578 bench.py --gpus 4 --steps 6 --warmup 3 --workload crossbar_2x2 > gpurun_out/r2_bench_crossbar_2x2_n4.json 2> gpurun_out/r2_bench_crossbar_2x2_n4.err
  python tools/show_bench.py gpurun_out/r2_bench_crossbar_2x2_n4.json | cut -c1-500
fi
