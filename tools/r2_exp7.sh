timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/check_multi_gpu.py --workload device_7.5nm --steps 2 --oracle 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tools/check_multi_gpu.py --workload crossbar_2x2 --steps 2 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -6
run() { n=$1; tag=$2; w=$3; shift 3
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $n --steps 6 --warmup 3 --workload $w > gpurun_out/r2_bench_${tag}.json 2> gpurun_out/r2_bench_${tag}.err
  python tools/show_bench.py gpurun_out/r2_bench_${tag}.json | cut -c1-500; grep -i "error" gpurun_out/r2_bench_${tag}.err | tail -2; }
run 2 crossbar_2x2_n2 crossbar_2x2
run 2 device_7.5nm_n2 device_7.5nm
