"""Tiny driver for profiling: runs the filter kernel a few times on one workload (no CPU leg, no torch)."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c2")
ap.add_argument("--profiles", type=int, default=0)
ap.add_argument("--frames", type=int, default=0)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
wl = dict(bench.WORKLOADS[a.workload])
if a.profiles:
    wl["P"] = a.profiles
if a.frames:
    wl["T"] = a.frames
model, traj, ss, thetas = bench.make_inputs(wl, 0)
for i in range(a.reps):
    t0 = time.perf_counter()
    out = model.logL_st_batch(ss, thetas, traj)
    dt = time.perf_counter() - t0
    print(f"rep {i}: {dt*1e3:.3f} ms  {wl['P']*(wl['T']-1)/dt:.4g} frame-steps/s  sum={out.sum():.6f}", flush=True)
print(model._handle(traj).describe_plan(wl["P"]))
