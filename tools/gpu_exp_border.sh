#!/bin/bash
# experiment: k_mmarb (border rows / columns in DFMAs) against k_mmar; placement of the border products (ORD)
mkdir -p gpurun_out
run() {  # label, workload, env...
  local label=$1 wl=$2; shift 2
  env "$@" timeout 300 python bench.py --workload $wl --steps 4 --warmup 3 --no-cpu --also '' > /tmp/b.json 2> /tmp/b.err || tail -3 /tmp/b.err
  python -c "
import json; d=json.load(open('/tmp/b.json')); print('$label $wl frac %.4f ms %.3f'%(d['roofline']['frac'], d['ms_per_step']), d['detail']['plan'])"
}
{
run mmar n25 BILDK_MMARB=0
run mmarb-ord0 n25 BILDK_MMARB_ORD=0
run mmarb-ord1 n25 BILDK_MMARB_ORD=1
run mmarb-ord2 n25 BILDK_MMARB_ORD=2
run mmar n17 BILDK_MMARB=0
run mmarb-ord0 n17 BILDK_MMARB_ORD=0
run mmarb-ord1 n17 BILDK_MMARB_ORD=1
run mmarb-ord2 n17 BILDK_MMARB_ORD=2
run mmarb-ord0 n26 BILDK_MMARB_ORD=0
run mmarb-ord0 n18 BILDK_MMARB_ORD=0
} | tee gpurun_out/exp_border2.txt
