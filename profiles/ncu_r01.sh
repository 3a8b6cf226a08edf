#!/bin/bash
# Round-1 ncu evidence, run on the GPU box:  gpurun --timeout 1500 -- bash profiles/ncu_r01.sh
# 1) plain run (must exit 0), 2) launch list of one warm KMC step, 3) one --set full capture per hot kernel.
# Kernels are serialised under ncu, so the side-stream overlap is not visible here: compare SHARES.
set -u
mkdir -p gpurun_out
python profiles/step_for_ncu.py > gpurun_out/step_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/step_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_r01.csv python profiles/step_for_ncu.py > gpurun_out/ncu_list.log 2>&1
for k in spmv_tile_kernel pairwise_cells_kernel cg_update_kernel cg_direction_kernel event_loop_kernel rate_rows_kernel assemble_kernel; do
  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$k -c 1 \
      -o gpurun_out/prof_r01_$k -f python profiles/step_for_ncu.py > gpurun_out/ncu_$k.log 2>&1
  tail -1 gpurun_out/ncu_$k.log
done
