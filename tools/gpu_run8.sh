mkdir -p gpurun_out
rm -f gpurun_out/mmac_help.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "against_c_oracle or kernel_variants or golden" 2>&1 | tail -3
BILDK_MMAC_HELPERS=0 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "against_c_oracle" 2>&1 | tail -1
run() {  # env-string workload
  env $1 timeout 600 python bench.py --workload $2 --steps 2 --warmup 3 --no-cpu 2>gpurun_out/err_$2.log | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print('$1', '$2', 'frac=%.4f' % d['roofline']['frac'], 'ms=%.3f' % d['roofline']['kernel_ms'], d['config']['plan'][:100])
" >> gpurun_out/mmac_help.log
}
run "BILDK_MMAC_HELPERS=0" c3
run "BILDK_MMAC_HELPERS=1" c3
cat gpurun_out/mmac_help.log
