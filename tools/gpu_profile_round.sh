#!/bin/bash
# One GPU session that refreshes everything profiles/ cites: tests, the default bench line, the per-workload
# bench lines, the ncu launch list of the bench command and one full ncu capture of the filter kernel.
TAG=${1:-v9}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > gpurun_out/gpu_tests_$TAG.log
cat gpurun_out/gpu_tests_$TAG.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c2_$TAG.json 2> gpurun_out/bench_c2_$TAG.err
cat gpurun_out/bench_c2_$TAG.json
for wl in n20big n25 n10 n50 c3; do
  timeout 600 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_${wl}_$TAG.json 2> gpurun_out/bench_${wl}_$TAG.err
  python -c "
import json,sys
d=json.load(open('gpurun_out/bench_${wl}_$TAG.json'))
print('$wl', 'frac=%.4f'%d['roofline']['frac'], 'value=%.4g'%d['value'], 'e2e=%.4g'%d['e2e']['value'], d['config']['plan'])"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2_$TAG.csv python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/ncu_launches_$TAG.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_mmar -c 1 -f -o gpurun_out/prof_c2_mmar_$TAG python tools/run_kernel.py --workload c2 --reps 2 > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log
timeout 900 python tools/sweep.py --only-N 100 > gpurun_out/sweep_n100_$TAG.jsonl 2> gpurun_out/sweep_n100_$TAG.md
tail -6 gpurun_out/sweep_n100_$TAG.md
