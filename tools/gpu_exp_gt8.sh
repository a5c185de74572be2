#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "four_warp or against_c_oracle or two_warp" 2>&1 | tail -4
run() {  # label, workload, env...
  local label=$1 wl=$2; shift 2
  env "$@" timeout 300 python bench.py --workload $wl --steps 4 --warmup 3 --no-cpu --also '' > /tmp/b.json 2> /tmp/b.err || tail -3 /tmp/b.err
  python -c "
import json; d=json.load(open('/tmp/b.json')); print('$label $wl frac %.4f ms %.3f'%(d['roofline']['frac'], d['ms_per_step']), d['detail']['plan'])"
}
{
run mmac n60 BILDK_MMAR2=0
run mmar2x4-maxf3 n60 A=1
run mmar2x4-maxf3-fpc1 n60 BILDK_FPC2=1
run mmar2x4-maxf2 n60 BILDK_MMAR2_MAXF=2
run mmar2x4-maxf2-fpc1 n60 BILDK_MMAR2_MAXF=2 BILDK_FPC2=1
run mmac n64 BILDK_MMAR2=0
run mmar2x4-maxf3 n64 A=1
run mmar2x4-maxf2 n64 BILDK_MMAR2_MAXF=2
run mmar2x4-maxf2-fpc1 n64 BILDK_MMAR2_MAXF=2 BILDK_FPC2=1
} | tee gpurun_out/exp_gt8.txt
