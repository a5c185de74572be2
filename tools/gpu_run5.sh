mkdir -p gpurun_out
rm -f gpurun_out/mmar_bench5.log
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/gpu_tests5.log
cat gpurun_out/gpu_tests5.log
run() {  # env-string workload
  env $1 timeout 300 python bench.py --workload $2 --steps 5 --warmup 3 --no-cpu 2>gpurun_out/err_$2.log | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print('$1', '$2', 'frac=%.4f' % d['roofline']['frac'], 'ms=%.3f' % d['roofline']['kernel_ms'], 'e2e=%.4g' % d['e2e']['value'], d['config']['plan'])
" >> gpurun_out/mmar_bench5.log
}
for nb in 4 5 7; do run "BILDK_MMAR_NB=$nb" c2; run "BILDK_MMAR_NB=$nb" n20big; run "BILDK_MMAR_NB=$nb" n10; done
for nb in 3 4 5; do run "BILDK_MMAR_NB=$nb" n25; done
cat gpurun_out/mmar_bench5.log
