#!/bin/bash
# experiment: register budgets / filters per CTA of the register-chained kernels
mkdir -p gpurun_out
run() {  # label, workload, env...
  local label=$1 wl=$2; shift 2
  env "$@" timeout 300 python bench.py --workload $wl --steps 4 --warmup 3 --no-cpu --also '' > /tmp/b.json 2> /tmp/b.err || tail -3 /tmp/b.err
  python -c "
import json; d=json.load(open('/tmp/b.json')); print('$label $wl frac %.4f ms %.3f'%(d['roofline']['frac'], d['ms_per_step']), d['detail']['plan'])"
}
{
run default n40 A=1
run fpc4 n40 BILDK_FPC2=4
run fpc2 n40 BILDK_FPC2=2
run maxf6-fpc6 n40 BILDK_MMAR2_MAXF=6 BILDK_FPC2=6
run maxf6-fpc3 n40 BILDK_MMAR2_MAXF=6 BILDK_FPC2=3
run maxf6-fpc2 n40 BILDK_MMAR2_MAXF=6 BILDK_FPC2=2
run maxf5-fpc5 n40 BILDK_MMAR2_MAXF=5 BILDK_FPC2=5
run default n36 A=1
run fpc4 n36 BILDK_FPC2=4
run maxf6-fpc6 n36 BILDK_MMAR2_MAXF=6 BILDK_FPC2=6
run maxf6-fpc3 n36 BILDK_MMAR2_MAXF=6 BILDK_FPC2=3
run default n48 A=1
run maxf5-fpc5 n48 BILDK_MMAR2_MAXF=5 BILDK_FPC2=5
run fpc2 n48 BILDK_FPC2=2
run default n24 A=1
run nb3 n24 BILDK_MMAR_NB=3
run nb5 n24 BILDK_MMAR_NB=5
run default n32 A=1
run nb4 n32 BILDK_MMAR_NB=4
run default n16 A=1
run nb5 n16 BILDK_MMAR_NB=5
} | tee gpurun_out/exp_mx.txt
