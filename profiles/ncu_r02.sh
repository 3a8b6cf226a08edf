#!/bin/bash
# Round-2 ncu evidence, run on the GPU box:  gpurun --timeout 900 -- bash profiles/ncu_r02.sh
# 1) plain run (must exit 0), 2) launch list of one warm KMC step (+ one public SpMV launch), 3) one --set full
# capture per hot kernel.  Kernels are serialised under ncu, so the side-stream overlap is not visible here:
# compare SHARES.  The persistent PCG kernel is ONE launch per solve (first solve + restarts).
set -u
mkdir -p gpurun_out
python profiles/step_for_ncu.py > gpurun_out/step_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/step_plain.log; exit 1; }
tail -1 gpurun_out/step_plain.log | cut -c1-300
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_r02.csv python profiles/step_for_ncu.py > gpurun_out/ncu_list.log 2>&1
for k in pcg_persistent_kernel pairwise_cells_kernel spmv_tile_kernel event_loop_kernel rate_rows_kernel assemble_kernel; do
  timeout 240 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$k -c 1 \
      -o gpurun_out/prof_r02_$k -f python profiles/step_for_ncu.py > gpurun_out/ncu_$k.log 2>&1
  tail -1 gpurun_out/ncu_$k.log | cut -c1-200
  # the raw page as CSV travels back in any case; the reports themselves only for the two dominant kernels (64 MiB limit)
  ncu -i gpurun_out/prof_r02_$k.ncu-rep --page raw --csv > gpurun_out/prof_r02_$k.raw.csv 2>/dev/null
  case $k in pcg_persistent_kernel|pairwise_cells_kernel) ;; *) rm -f gpurun_out/prof_r02_$k.ncu-rep ;; esac
done
ls -la gpurun_out/prof_r02_* | awk '{print $5, $9}'
