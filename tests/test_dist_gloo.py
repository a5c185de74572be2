"""
CPU tests of the N > 1 path (gloo, world_size 2): contiguous sharding of an AMIS batch, one all-gather of
logL, and bit-identical sampler state on every rank.  The likelihood is answered by the oracle-backed test
double (no GPU here); the sharding / collective code is the product's (bild_b200/dist.py).
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds():
    from bild_b200.dist import shard_bounds
    assert list(shard_bounds(10, 3)) == [0, 4, 7, 10]
    assert list(shard_bounds(2, 4)) == [0, 1, 2, 2, 2]
    assert list(shard_bounds(0, 2)) == [0, 0, 0]
    for P in (1, 7, 100, 4096):
        for w in (1, 2, 8):
            b = shard_bounds(P, w)
            assert b[0] == 0 and b[-1] == P and np.all(np.diff(b) >= 0) and np.max(np.diff(b)) - np.min(np.diff(b)) <= 1


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bild_b200 as bild
    from test_amis_host import OracleBackedRouse
    model = OracleBackedRouse(8, 1, 5, d=2, localization_error=0.3)
    np.random.seed(11)
    traj = model.trajectory_from_loopingprofile(bild.Loopingprofile([0] * 15 + [1] * 15))
    calls = []
    orig = model.logL_st_batch

    def local(a, b, traj_=traj):
        calls.append(len(a))
        return OracleBackedRouse.logL_st_batch(model, a, b, traj_)

    model._logL_st_local = lambda a, b, t: local(a, b)
    bild.models.MultiStateRouse.shard_over(model, None, "cpu")
    # route the sampler's batch through the sharder
    model.logL_st_batch = lambda ss, th, t: model._sharder(lambda a, b: local(a, b), np.asarray(ss), np.asarray(th))
    np.random.seed(5)
    smp = bild.amis.FixedkSampler(traj, model, k=3, N=33)
    for _ in range(3):
        smp.step()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), ev=np.array(smp.evidences), logL=smp.samples[-1]["logLs"],
             ss=smp.samples[-1]["ss"], calls=np.array(calls))
    dist.destroy_process_group()


def test_sharded_sampler_gloo_world2(tmp_path):
    import torch.multiprocessing as mp
    port = 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = np.load(tmp_path / "rank0.npz")
    r1 = np.load(tmp_path / "rank1.npz")
    # both ranks hold the identical ensemble and evidences (bitwise), each evaluated only its block
    assert np.array_equal(r0["ev"], r1["ev"]) and np.array_equal(r0["logL"], r1["logL"]) and np.array_equal(r0["ss"], r1["ss"])
    assert list(r0["calls"]) == [17, 17, 17] and list(r1["calls"]) == [16, 16, 16]

    # and they equal the unsharded single-process run
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import bild_b200 as bild
    from test_amis_host import OracleBackedRouse
    model = OracleBackedRouse(8, 1, 5, d=2, localization_error=0.3)
    np.random.seed(11)
    traj = model.trajectory_from_loopingprofile(bild.Loopingprofile([0] * 15 + [1] * 15))
    np.random.seed(5)
    smp = bild.amis.FixedkSampler(traj, model, k=3, N=33)
    for _ in range(3):
        smp.step()
    assert np.array_equal(np.array(smp.evidences), r0["ev"])
    assert np.array_equal(smp.samples[-1]["logLs"], r0["logL"])


def _claim_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bild_b200 as bild
    from bild_b200.dataset import sample_many, store_claimer
    from test_amis_host import OracleBackedRouse
    model = OracleBackedRouse(8, 1, 5, d=2, localization_error=0.3)
    np.random.seed(21)
    trajs = [model.trajectory_from_loopingprofile(bild.Loopingprofile([0] * a + [1] * b + [0] * c))
             for a, b, c in [(8, 9, 7), (12, 10, 0), (5, 5, 9), (20, 0, 0), (6, 8, 6), (9, 9, 4), (3, 14, 5)]]
    kw = dict(init_runs=3, sampler_kw={"N": 15, "max_fcomplete": 40}, k_max=3, certainty_in_k=0.9)
    dist.barrier()      # both ranks start claiming together (imports and model set-up take different times per process)
    res, stats = sample_many(trajs, model, seeds=list(range(300, 307)), claim=store_claimer(len(trajs)), max_active=2, **kw)
    np.savez(os.path.join(out_dir, f"claim{rank}.npz"), idx=np.array(sorted(res), dtype=int),
             **{f"ev{i}": res[i].evidence for i in res})
    dist.barrier()
    dist.destroy_process_group()


def test_dynamic_trajectory_assignment_gloo_world2(tmp_path):
    """`store_claimer`: every trajectory is claimed by exactly one rank; the results do not depend on the rank."""
    import torch.multiprocessing as mp
    port = 29950 + os.getpid() % 300
    mp.spawn(_claim_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r = [np.load(tmp_path / f"claim{i}.npz") for i in range(2)]
    got = sorted(list(r[0]["idx"]) + list(r[1]["idx"]))
    assert got == list(range(7)) and len(r[0]["idx"]) >= 1 and len(r[1]["idx"]) >= 1
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import bild_b200 as bild
    from test_amis_host import OracleBackedRouse
    model = OracleBackedRouse(8, 1, 5, d=2, localization_error=0.3)
    np.random.seed(21)
    trajs = [model.trajectory_from_loopingprofile(bild.Loopingprofile([0] * a + [1] * b + [0] * c))
             for a, b, c in [(8, 9, 7), (12, 10, 0), (5, 5, 9), (20, 0, 0), (6, 8, 6), (9, 9, 4), (3, 14, 5)]]
    kw = dict(init_runs=3, sampler_kw={"N": 15, "max_fcomplete": 40}, k_max=3, certainty_in_k=0.9)
    for i in range(7):
        np.random.seed(300 + i)
        solo = bild.sample(trajs[i], model, **kw)
        owner = r[0] if i in r[0]["idx"] else r[1]
        assert np.array_equal(solo.evidence, owner[f"ev{i}"])
