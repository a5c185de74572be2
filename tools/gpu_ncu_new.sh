#!/bin/bash
# ncu --set full captures of the round-2 kernel variants (summarised on the box; the reports stay in /tmp)
TAG=${1:-r02g}
mkdir -p gpurun_out
for spec in "n25:k_mmarb:200:0" "n32:k_mmar:200:0" "n48:k_mmar2:100:0" "n40:k_mmar2:100:0"; do
  IFS=: read wl kern frames prof <<< "$spec"
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$kern -c 1 -f -o /tmp/prof_${wl}_$TAG \
      python tools/run_kernel.py --workload $wl --frames $frames --reps 1 > gpurun_out/ncu_full_${wl}_$TAG.log 2>&1
  python tools/ncu_summary.py /tmp/prof_${wl}_$TAG.ncu-rep > gpurun_out/ncu_${wl}_$TAG.txt 2>&1
  head -12 gpurun_out/ncu_${wl}_$TAG.txt
done
