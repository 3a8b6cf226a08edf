python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_tiled_1M_n1.json 2> gpurun_out/r02_bench_tiled_1M_n1.err
python tools/show_bench.py gpurun_out/r02_bench_tiled_1M_n1.json | cut -c1-400
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
