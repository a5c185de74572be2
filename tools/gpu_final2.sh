#!/bin/bash
# after the validation run: refresh the sweep rows of the kernels that changed late (N = 25: k_mmarb column-wise, N = 100: k_mmar8)
# and take ncu --set full summaries of k_mmar8 (configs[2]) and the four-warp k_mmar2 (N = 60)
TAG=${1:-r02y}
mkdir -p gpurun_out
timeout 900 python tools/sweep.py --only-N 25 100 > gpurun_out/sweep_$TAG.jsonl 2> gpurun_out/sweep_$TAG.md
grep "^|" gpurun_out/sweep_$TAG.md | cut -c1-200
for spec in "c3:k_mmar8:30:2048" "n60:k_mmar2:60:2048"; do
  IFS=: read wl kern frames prof <<< "$spec"
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$kern -c 1 -f -o /tmp/prof_${wl}_$TAG \
      python tools/run_kernel.py --workload $wl --frames $frames --profiles $prof --reps 1 > gpurun_out/ncu_full_${wl}_$TAG.log 2>&1
  python tools/ncu_summary.py /tmp/prof_${wl}_$TAG.ncu-rep > gpurun_out/ncu_${wl}_$TAG.txt 2>&1
  head -36 gpurun_out/ncu_${wl}_$TAG.txt | tail -30
done
