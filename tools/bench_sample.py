"""
BASELINE.json configs[0]: wall seconds of one full `bild.sample` run (MultiStateRouse N=20, d=3, T=100, defaults) on
the trajectory recorded in tests/golden/sample_runs.npz, same seed as the reference run that produced the golden.

    python tools/bench_sample.py                    # this engine (needs a GPU)
    python tools/bench_sample.py --impl reference   # the unmodified reference with its own compiled .pyx
                                                    # (only where /root/reference exists: the build container)
"""
import argparse
import json
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
runs = np.load(os.path.join(ROOT, "tests", "golden", "sample_runs.npz"))

if a.impl == "reference":
    sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))
    sys.path.insert(0, "/root/reference")
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import bild
        import bild.models as bm
    import kalman_oracle as ko
    import noctiluca as nl
    bm.MSRouse_logL = ko.ref_cython()
    make_traj = lambda: nl.Trajectory(runs["c1_x"], localization_error=[0.3] * 3)   # noqa: E731
else:
    sys.path.insert(0, ROOT)
    import bild_b200 as bild
    make_traj = lambda: bild.Trajectory(runs["c1_x"], localization_error=[0.3] * 3)   # noqa: E731

walls = []
for rep in range(a.reps + 1):          # first pass = warm-up (library load, handle creation)
    model = bild.models.MultiStateRouse(20, 1, 5, d=3, localization_error=0.3)
    traj = make_traj()
    np.random.seed(1234)
    t0 = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = bild.sample(traj, model)
    walls.append(time.perf_counter() - t0)
ok = bool(np.array_equal(res.k, runs["c1_k"]) and np.array_equal(res.log["k"], runs["c1_logk"]))
print(json.dumps({"metric": "bild.sample wall seconds", "impl": a.impl, "value": float(np.median(walls[1:])), "unit": "s",
                  "higher_is_better": False, "walls_s": [round(w, 4) for w in walls], "amis_steps": int(len(res.log["k"])),
                  "logL_evaluations": int(sum(len(s["logLs"]) for smp in res.samplers for s in smp.samples)),
                  "same_step_sequence_as_golden": ok, "host_cores_used": 1,
                  "config": {"workload": "configs[0]: bild.sample, MultiStateRouse N=20 d=3 T=100, defaults"}}))
