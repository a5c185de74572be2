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
