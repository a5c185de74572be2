"""
BASELINE.json configs[4]: throughput sweep of the batched logL over N x T x batch, next to the reference
Cython path on all host cores (bounded sample per point).  One JSON object per point + a markdown table.

    python tools/sweep.py [--quick] > profiles/sweep.jsonl
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--quick", action="store_true")
ap.add_argument("--cpu-seconds", type=float, default=2.0, help="wall seconds on all cores per point and CPU implementation")
ap.add_argument("--only-N", type=int, nargs="*", default=None, help="restrict the sweep to these polymer sizes")
a = ap.parse_args()

Ns = [10, 25, 50, 100, 200] if not a.only_N else list(a.only_N)
Ts = [100, 1000]
Ps = [1024, 65536]
points = [(N, T, P) for N in Ns for T in Ts for P in Ps]
if a.quick:
    points = [(N, 100, 1024) for N in Ns]

# CPU legs first (fork pools must precede CUDA initialisation): the reference's compiled .pyx AND its pure-Python twin
# (MSRouse_logL_py.py, the faster CPU path for N >= 50) on all host cores; the speed-up is quoted against the faster one
sys.path.insert(0, os.path.join(bench.ROOT, "oracle"))
import kalman_oracle as ko  # noqa: E402
cpu = {}
for N, T, P in points:
    key = (N, T)
    if key in cpu:
        continue
    wl = dict(N=N, T=T, P=2048, p_nan=0.0)
    model, traj, ss, thetas = bench.make_inputs(wl, 0)
    r = bench.cpu_arm(wl, model, traj, ss, thetas, budget_s=a.cpu_seconds)
    best = r["impls"][r["best"]]
    pyx = r["impls"].get("cython_pyx", best)
    cpu[key] = dict(frame_steps_per_s=best["frame_steps_per_s"], cores=best["cores"], kind=r["kind"], impl=r["best"], n=pyx["n"], logL=pyx["logL"],
                    ss=ss[:pyx["n"]], thetas=thetas[:pyx["n"]],
                    all={k: v["frame_steps_per_s"] for k, v in r["impls"].items()})

import torch  # noqa: E402
from bild_b200 import _lib  # noqa: E402
from bild_b200.engine import st_to_runs  # noqa: E402
lib = _lib.load()
peak = max(bench.measure_fp64_peak(0))
dev = torch.device("cuda", 0)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
rows = []
for N, T, P in points:
    if N == 200 and P > 1024 and T > 100:
        rows.append(dict(N=N, T=T, P=P, skipped="not run: 2.1e15 flop per pass (about 1.5 minutes at the measured N=200 rate; three passes per point)"))
        print(json.dumps(rows[-1]), flush=True)
        continue
    wl = dict(N=N, T=T, P=P, p_nan=0.0)
    model, traj, ss, thetas = bench.make_inputs(wl, 0)
    th = model._handle(traj)
    starts, rstates = st_to_runs(ss, thetas, T)
    d_s, d_r = torch.from_numpy(starts).to(dev), torch.from_numpy(rstates).to(dev)
    d_o = torch.empty(P, dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream()

    def run():
        model.engine.logl_runs_device(th, P, starts.shape[1], d_s.data_ptr(), d_r.data_ptr(), d_o.data_ptr(), st.cuda_stream)

    run()
    torch.cuda.synchronize()
    times = []
    reps = 3 if N < 200 else 1
    if N == 200 and P > 1024:
        reps = 1
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); run(); e1.record(st)
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = float(np.mean(times))
    t0 = time.perf_counter()
    host = model.logL_st_batch(ss, thetas, traj)
    e2e_s = time.perf_counter() - t0
    c = cpu[(N, T)]
    # parity on the CPU sample: same model parameters and seed -> same trajectory; evaluate the CPU sample's profiles
    chk = model.logL_st_batch(c["ss"], c["thetas"], traj)
    want, note = c["logL"], ""
    if not np.all(np.isfinite(want)):
        # the reference .pyx returns NaN here (N = 200: its dsymv calls with beta = 0 write into np.empty buffers,
        # pyx:168, 199, and scipy's OpenBLAS propagates the uninitialised NaNs for large N); check against the C oracle
        m = min(4, len(chk))
        arrs = ko.model_arrays(model.models)
        s2o, cio = ko.noise_to_s2_cind(model._get_noise(traj))
        want = ko.logl_c(*arrs, model.measurement, traj[:], s2o, cio, np.array([ko.st2states(c["ss"][i], c["thetas"][i], T) for i in range(m)]))
        chk = chk[:m]
        note = "reference returned NaN; parity vs C oracle"
    rel = float(np.max(np.abs(chk - want) / np.maximum(1, np.abs(want))))
    fl = bench.flops_per_eval(N, 3, 1, T, T) * P
    row = dict(N=N, T=T, P=P, kernel_ms=ms, frame_steps_per_s=P * (T - 1) / (ms * 1e-3), evals_per_s=P / (ms * 1e-3),
               e2e_frame_steps_per_s=P * (T - 1) / e2e_s, tflops=fl / (ms * 1e-3) * 1e-12, frac_of_fp64_peak=fl / (ms * 1e-3) * 1e-12 / peak,
               cpu_frame_steps_per_s=c["frame_steps_per_s"], cpu_cores=c["cores"], cpu_kind=c["kind"], cpu_impl=c["impl"], cpu_all=c["all"], cpu_sample=c["n"],
               speedup_vs_cpu=P * (T - 1) / (ms * 1e-3) / c["frame_steps_per_s"], max_rel_err_vs_cpu=rel,
               plan=th.describe_plan(P), peak_tflops=peak, note=note)
    rows.append(row)
    print(json.dumps(row), flush=True)

print("\n| N | T | P | kernel ms | frame-steps/s | TFLOP/s (alg.) | of FP64 peak | CPU ref, faster of .pyx / twin (all cores) | speed-up | max rel err | kernel |", file=sys.stderr)
print("|---|---|---|---|---|---|---|---|---|---|---|", file=sys.stderr)
for r in rows:
    if "skipped" in r:
        print(f"| {r['N']} | {r['T']} | {r['P']} | - | - | - | - | - | - | - | {r['skipped']} |", file=sys.stderr)
    else:
        print(f"| {r['N']} | {r['T']} | {r['P']} | {r['kernel_ms']:.2f} | {r['frame_steps_per_s']:.3g} | {r['tflops']:.1f} | {r['frac_of_fp64_peak']:.2f} | "
              f"{r['cpu_frame_steps_per_s']:.3g} ({r['cpu_impl']}, {r['cpu_cores']} cores) | {r['speedup_vs_cpu']:.0f}x | {r['max_rel_err_vs_cpu']:.1e} | {r['plan'].split(' threads')[0]} |", file=sys.stderr)
