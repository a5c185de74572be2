#!/bin/bash
# Round-2 GPU session: tests, the default bench line (N=50 target + sub-records) with its reference arm, the ncu launch
# list of that command, ncu --set full captures of the kernels behind the headline numbers (summarised ON THE BOX by
# tools/ncu_summary.py - gpurun_out/ may carry at most 64 MiB back, one .ncu-rep is 15-20 MB), and the configs[4] sweep.
#   tools/gpu_round2.sh TAG [steps...]      steps: tests bench ref launches ncu exp sweep (default: all but exp)
TAG=${1:-r02a}; shift
STEPS=${@:-tests bench ref launches ncu sweep}
mkdir -p gpurun_out
has() { [[ " $STEPS " == *" $1 "* ]]; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.txt 2>&1
nproc > gpurun_out/nproc_$TAG.txt
if has tests; then
  timeout 1200 python -m pytest tests -x -q -m gpu --durations=8 > gpurun_out/gpu_tests_$TAG.log 2>&1
  tail -15 gpurun_out/gpu_tests_$TAG.log
fi
if has bench; then
  timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_$TAG.json"))
print("bench: value %.4g e2e %.4g frac %.3f cpu %.4g (%s) parity %s" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["cpu_baseline"]["value"], d["cpu_baseline"]["implementation"], d["parity"]))
for s in d.get("also", []):
    print("  also", s["name"], "value %.4g" % s["value"], "frac", s.get("roofline", {}).get("frac"), "cpu", s.get("cpu_baseline", {}).get("value"), s.get("parity"))
PY
  tail -3 gpurun_out/bench_$TAG.err
fi
if has ref; then
  timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
  tail -c 400 gpurun_out/bench_ref_$TAG.json
fi
if has launches; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
      python bench.py --steps 5 --warmup 3 --no-cpu --also '' > gpurun_out/ncu_launches_$TAG.log 2>&1
  tail -3 gpurun_out/launches_$TAG.csv
fi
if has ncu; then
  for spec in "n50:k_mmar2:100" "n50small:k_mmar2:100" "n200:k_mmag2:20" "c2:k_mmar:500" "c3:k_mmar8:40"; do
    IFS=: read wl kern frames <<< "$spec"
    timeout 600 ncu --set full --import-source on --clock-control none -k regex:$kern -c 1 -f -o /tmp/prof_${wl}_$TAG \
        python tools/run_kernel.py --workload $wl --frames $frames --reps 1 > gpurun_out/ncu_full_${wl}_$TAG.log 2>&1
    python tools/ncu_summary.py /tmp/prof_${wl}_$TAG.ncu-rep > gpurun_out/ncu_${wl}_$TAG.txt 2>&1
    ncu -i /tmp/prof_${wl}_$TAG.ncu-rep --page raw --csv > gpurun_out/ncu_${wl}_${TAG}_raw.csv 2>/dev/null
    head -40 gpurun_out/ncu_${wl}_$TAG.txt
  done
  cp /tmp/prof_n50_$TAG.ncu-rep gpurun_out/ 2>/dev/null     # one full report travels back (source view of the headline kernel)
fi
if has exp; then
  # experiments: k_mma2 filters per CTA (one full wave each), the DFMA tile kernel at N = 10
  for f in 1 2 3 4 5 6; do
    BILDK_FPC2=$f python tools/run_kernel.py --workload n50 --frames 200 --profiles $((148 * f)) --reps 3 2>&1 | tail -2 | sed "s/^/FPC2=$f one wave: /"
  done
  for f in 3 4 5 6; do
    BILDK_FPC2=$f python tools/run_kernel.py --workload n50 --frames 200 --profiles 1024 --reps 3 2>&1 | tail -2 | sed "s/^/FPC2=$f P=1024: /"
  done
  python tools/run_kernel.py --workload n10 --frames 200 --reps 3 2>&1 | tail -2 | sed "s/^/n10 default: /"
  BILDK_KERNEL=tile python tools/run_kernel.py --workload n10 --frames 200 --reps 3 2>&1 | tail -2 | sed "s/^/n10 tile: /"
  python tools/run_kernel.py --workload c2 --reps 5 2>&1 | tail -3 | sed "s/^/c2: /"
fi > gpurun_out/exp_$TAG.txt 2>&1
has exp && cat gpurun_out/exp_$TAG.txt
if has sweep; then
  timeout 1200 python tools/sweep.py > gpurun_out/sweep_$TAG.jsonl 2> gpurun_out/sweep_$TAG.md
  tail -24 gpurun_out/sweep_$TAG.md
fi
du -sh gpurun_out
if has dataset; then
  timeout 900 python tools/bench_dataset.py --n-traj ${NTRAJ:-128} --check 2 > gpurun_out/dataset_${NTRAJ:-128}traj_1gpu_$TAG.json 2> gpurun_out/dataset_$TAG.err
  cat gpurun_out/dataset_${NTRAJ:-128}traj_1gpu_$TAG.json; tail -3 gpurun_out/dataset_$TAG.err
fi
if has small; then
  timeout 600 python bench.py --workload n50small --steps 10 --warmup 3 --no-cpu --also '' > gpurun_out/bench_n50small_$TAG.json 2> gpurun_out/bench_n50small_$TAG.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_n50small_$TAG.json')); print('n50small frac %.3f value %.4g'%(d['roofline']['frac'], d['value']), d['detail']['plan'])"
fi
