#!/bin/bash
# k_mmarb final (column-wise border products): parity + throughput where selected, and where it is not (BILDK_MMARB=2)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "border or external_force or against_c_oracle or register_chained_kernel" 2>&1 | tail -3
run() {  # label, workload, env...
  local label=$1 wl=$2; shift 2
  env "$@" timeout 300 python bench.py --workload $wl --steps 4 --warmup 3 --no-cpu --also '' > /tmp/b.json 2> /tmp/b.err || tail -3 /tmp/b.err
  python -c "
import json; d=json.load(open('/tmp/b.json')); print('$label $wl frac %.4f ms %.3f'%(d['roofline']['frac'], d['ms_per_step']), d['detail']['plan'])"
}
{
run default n25 A=1
run default n26 A=1
run default n17 A=1
run default n18 A=1
run forced-mmarb n18 BILDK_MMARB=2
run default n10 A=1
run forced-mmarb n10 BILDK_MMARB=2
} | tee gpurun_out/exp_border4.txt
timeout 600 python tools/bench_dataset.py --n-traj 64 --check 0 --profile > gpurun_out/dataset_prof.json 2> gpurun_out/dataset_prof.txt; tail -c 600 gpurun_out/dataset_prof.json
