#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "eight_warp and (105 or 108) or (against_c_oracle and (108 or 110))" 2>&1 | tail -3
run() {  # label, workload, env...
  local label=$1 wl=$2; shift 2
  env "$@" timeout 120 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu --also '' > /tmp/b.json 2> /tmp/b.err || { tail -3 /tmp/b.err; echo "$label $wl FAILED/timeout"; return; }
  python -c "
import json; d=json.load(open('/tmp/b.json')); print('$label $wl frac %.4f ms %.3f'%(d['roofline']['frac'], d['ms_per_step']), d['detail']['plan'])"
}
{
run mmact n108 BILDK_MMAR8=0
run mmar8 n108 A=1
} | tee gpurun_out/exp_mmar8c.txt
