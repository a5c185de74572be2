"""
N > 1 on hardware: the PRODUCT's sharded evaluator (`MultiStateRouse.shard_over` -> bild_b200.dist.ShardedEvaluator)
over NCCL with two ranks, one GPU each.  Skipped on a box with a single GPU (run it with `gpurun --gpus 2`).
"""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_devices():
    try:
        from bild_b200 import _lib
        return int(_lib.load().bildk_device_count())
    except Exception:
        return 0


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import bild_b200 as bild
    model = bild.models.MultiStateRouse(20, 1, 5, d=3, localization_error=0.3, device=rank)
    np.random.seed(3)
    traj = model.trajectory_from_loopingprofile(bild.Loopingprofile((np.arange(120) // 30) % 2), missing_frames=0.1)
    rng = np.random.default_rng(4)
    P, K1 = 1001, 6                                    # odd size: ranks own 501 and 500 profiles
    ss = rng.dirichlet(np.ones(K1), size=P)
    thetas = (rng.integers(0, 2, size=(P, 1)) + np.arange(K1)[None, :]) % 2
    plain = model.logL_st_batch(ss, thetas, traj)      # unsharded, this rank's GPU
    model.shard_over(device=f"cuda:{rank}")
    sharded = model.logL_st_batch(ss, thetas, traj)    # block per rank + one NCCL all-gather
    # a sharded AMIS sampler: identical state on both ranks (replicated host RNG, gathered likelihoods)
    np.random.seed(9)
    smp = bild.amis.FixedkSampler(traj, model, k=3, N=101)
    for _ in range(3):
        smp.step()
    np.savez(os.path.join(out_dir, f"nccl{rank}.npz"), plain=plain, sharded=sharded, ev=np.array(smp.evidences),
             logL=smp.samples[-1]["logLs"])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(_n_devices() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_sharded_evaluator_nccl_world2(tmp_path):
    import torch.multiprocessing as mp
    port = 29700 + os.getpid() % 200
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (np.load(tmp_path / f"nccl{i}.npz") for i in range(2))
    assert np.array_equal(r0["sharded"], r0["plain"]) and np.array_equal(r1["sharded"], r1["plain"])   # same kernels, same bits
    assert np.array_equal(r0["sharded"], r1["sharded"])
    assert np.array_equal(r0["ev"], r1["ev"]) and np.array_equal(r0["logL"], r1["logL"])
