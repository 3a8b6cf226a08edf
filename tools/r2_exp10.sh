python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 120 python bench.py --gpus 1 --steps 6 --warmup 3 --no-cpu-baseline --workload crossbar_2x2 > gpurun_out/r2_bench_crossbar_2x2_n1b.json 2> gpurun_out/r2_bench_crossbar_2x2_n1b.err
python tools/show_bench.py gpurun_out/r2_bench_crossbar_2x2_n1b.json | cut -c1-420
timeout 120 python bench.py --gpus 1 --steps 3 --warmup 0 --no-cpu-baseline > gpurun_out/r2_bench_1M_cold.json 2> gpurun_out/r2_bench_1M_cold.err
python tools/show_bench.py gpurun_out/r2_bench_1M_cold.json | cut -c1-420
