"""
BASELINE.json configs[3]: a dataset of synthetic trajectories (T=300, N=50) through the FULL bild.sample
scheme, likelihood batches of all trajectories fused per round (bild_b200.dataset.sample_many), trajectories
partitioned across ranks.  Prints one JSON line per run (rank 0).

    python tools/bench_dataset.py --n-traj 128 [--N 50 --T 300 --check 2]
    python -m torch.distributed.run --nproc-per-node 8 ... tools/bench_dataset.py --n-traj 1024
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bild_b200 as bild  # noqa: E402
from bild_b200.dataset import sample_many, store_claimer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n-traj", type=int, default=128)
ap.add_argument("--N", type=int, default=50)
ap.add_argument("--T", type=int, default=300)
ap.add_argument("--check", type=int, default=2, help="re-run this many trajectories one by one and compare")
ap.add_argument("--profile", action="store_true", help="cProfile the driver on rank 0 and print the top of the list to stderr")
ap.add_argument("--static", action="store_true", help="static round-robin partition instead of dynamic claims from a shared counter")
ap.add_argument("--max-active", type=int, default=0, help="concurrent state machines per rank (default: 64 dynamic / all static)")
ap.add_argument("--no-fuse-amis", action="store_true", help="one synchronous bildk_amis_step per sampler step instead of riding on the likelihood launch")
ap.add_argument("--schedule", default="priority", choices=["priority", "rounds"], help="dataset driver scheduling (bild_b200/dataset.py)")
a = ap.parse_args()
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))

model = bild.models.MultiStateRouse(a.N, 1, 5, d=3, localization_error=0.3, device=local)
np.random.seed(685441950)
trajs, truths = [], []
for i in range(a.n_traj):
    truth = (np.cumsum(np.random.rand(a.T) < 5.0 / a.T) % 2).astype(int)
    truths.append(truth)
    trajs.append(model.trajectory_from_loopingprofile(bild.Loopingprofile(truth)))
seeds = [1000 + i for i in range(a.n_traj)]

if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dist.barrier()
from bild_b200 import _lib  # noqa: E402
tic = time.perf_counter()
model.engine          # CUDA context, library load and model upload happen before the timed region (reported as setup_s)
setup_s = time.perf_counter() - tic
if world > 1:
    dist.barrier()
l0 = _lib.load().bildk_launch_count()
t0 = time.perf_counter()
claim = store_claimer(len(trajs)) if (world > 1 and not a.static) else None
prof = None
if a.profile and rank == 0:
    import cProfile
    prof = cProfile.Profile()
    prof.enable()
res, stats = sample_many(trajs, model, seeds=seeds, rank=rank, world=world, claim=claim, max_active=a.max_active or None, fuse_amis=not a.no_fuse_amis, schedule=a.schedule)
if prof is not None:
    import io
    import pstats
    prof.disable()
    buf = io.StringIO()
    pstats.Stats(prof, stream=buf).sort_stats("tottime").print_stats(30)
    print(buf.getvalue(), file=sys.stderr)
wall = time.perf_counter() - t0
launches = _lib.load().bildk_launch_count() - l0
summary = np.array([wall, stats["frame_steps"], stats["profiles"], stats["rounds"], len(res),
                    sum(int(np.array_equal(res[i].best_profile()[:], truths[i])) for i in res), -wall], dtype=np.float64)
if world > 1:
    t = torch.from_numpy(summary).cuda()
    tmax = t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)       # the one collective of the dataset run
    summary = t.cpu().numpy()
    summary[0] = float(tmax[0])
    summary[3] = float(tmax[3])
    summary[6] = float(tmax[6])          # max of -wall = -(fastest rank's wall)

per_rank = None
if world > 1:
    mine = torch.tensor([wall, stats["t_host_lanes"], stats["t_gpu"], stats["t_pack"], stats["launches"], stats["profiles"], len(res),
                         max((len(r.log["k"]) for r in res.values()), default=0)], dtype=torch.float64).cuda()
    allr = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allr, mine)
    per_rank = [dict(zip(("wall_s", "t_host_lanes", "t_gpu", "t_pack", "launches", "profiles", "trajectories", "longest_run_steps"),
                         [round(float(v), 3) for v in t.cpu()])) for t in allr]
check = {}
if rank == 0 and a.check:
    worst = 0.0
    same = True
    for i in sorted(res)[:a.check]:
        np.random.seed(seeds[i])
        solo = bild.sample(trajs[i], model)
        same &= bool(np.array_equal(solo.log["k"], res[i].log["k"]))
        worst = max(worst, float(np.max(np.abs(solo.evidence - res[i].evidence))))
    check = {"n": a.check, "identical_k_sequence": same, "max_abs_dlogE_vs_one_by_one": worst}
if rank == 0:
    print(json.dumps({
        "metric": "bild.sample dataset wall seconds", "value": summary[0], "unit": "s", "higher_is_better": False,
        "n_gpus": world, "config": {"workload": f"configs[3]: {a.n_traj} trajectories, N={a.N}, T={a.T}, full bild.sample, defaults",
                                    "parallelism": f"trajectories partitioned over {world} rank(s), fused likelihood batches"},
        "frame_steps": summary[1], "frame_steps_per_s": summary[1] / summary[0], "profiles": summary[2],
        "fused_rounds_max": summary[3], "rank_wall_min_s": -summary[6], "rank_imbalance": summary[0] / max(-summary[6], 1e-9) - 1.0,
        "partition": "static round-robin" if (a.static or world == 1) else "dynamic (shared counter)", "launches_rank0": int(launches), "setup_s_rank0": round(setup_s, 3), "schedule": a.schedule, "amis_bookkeeping": "synchronous per step" if a.no_fuse_amis else "fused into the likelihood launch", "trajectories": int(summary[4]),
        "truth_recovered_exactly": int(summary[5]), "check": check,
        "rank0_wall_split_s": {k: round(stats[k], 3) for k in ("t_host_lanes", "t_pack", "t_gpu", "t_submit")}, "per_rank": per_rank, "heaviest_trajectories_rank0": sorted(stats["lanes"], key=lambda t: -t[2])[:5], "data": "synthetic", "dtype": "f64"}), flush=True)
if world > 1:
    dist.destroy_process_group()
