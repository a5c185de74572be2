#!/bin/bash
# Roofline fraction of the filter kernel for a list of workloads (no CPU legs):  tools/gpu_sizes.sh TAG n16 n24 ...
TAG=$1; shift
mkdir -p gpurun_out
for wl in "$@"; do
  timeout 300 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu --also '' > gpurun_out/bench_${wl}_$TAG.json 2> gpurun_out/bench_${wl}_$TAG.err || tail -5 gpurun_out/bench_${wl}_$TAG.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_${wl}_$TAG.json')); print('$wl frac %.4f value %.4g ms %.3f'%(d['roofline']['frac'], d['value'], d['ms_per_step']), d['detail']['plan'])"
done | tee gpurun_out/sizes_$TAG.txt
