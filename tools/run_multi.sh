run() { # n tag env...
  n=$1; tag=$2; shift 2
  env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $n --steps 10 --warmup 5 > gpurun_out/r2_bench_${tag}.json 2> gpurun_out/r2_bench_${tag}.err
  python tools/show_bench.py gpurun_out/r2_bench_${tag}.json
  grep -E "pcg prof|rror" gpurun_out/r2_bench_${tag}.err | tail -2 | cut -c1-260
}
