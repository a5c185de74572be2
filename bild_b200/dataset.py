"""
Dataset runs: many trajectories through the full `sample` scheme with their likelihood batches fused.

The reference analyses a dataset by calling ``bild.sample`` once per trajectory, one after the other
(/root/reference/bild/core.py:22 has no dataset entry point); each call is a data-dependent sequential state
machine (core.py:202-229) that needs a likelihood batch of only ``sampler_kw['N']`` (default 100) profiles
per step - far too few to fill a B200.  `sample_many` runs all state machines concurrently and fuses the
pending batches of all trajectories into ONE multi-trajectory launch per round (C ABI
``bildk_logl_runs_multi``): SURVEY.md section 8(f) rank 2, BASELINE.json configs[3].

Determinism: every trajectory owns its numpy RNG stream.  The state machines are generators (`core.sample_gen`
yields every likelihood batch it needs); exactly one runs at a time, in a fixed round-robin order, and the global
``np.random`` state is swapped at every hand-over, so trajectory ``i`` consumes exactly the random numbers it would
consume in ``np.random.seed(seeds[i]); sample(trajs[i], model, ...)`` run on its own - the results are identical to
the sequential runs, independent of how many trajectories share a launch or a GPU.  (A first version ran every
`sample` in its own thread and handed over with semaphores: the hand-overs cost as much as the AMIS bookkeeping.)

Multi-GPU: trajectories are partitioned across ranks; no data-path communication during sampling.  The partition is
static round-robin (`rank`, `world`) or, with ``claim=store_claimer(...)``, DYNAMIC: every rank keeps `max_active` state
machines running and claims the next trajectory indices from a shared counter (a `torch.distributed` store - control
plane only) whenever one finishes.  The cost of a trajectory is data dependent (20 to 700 AMIS rounds in the 1024-trajectory
run of round 1, where the static partition left the slowest rank 1.8x behind the fastest); a result does not depend on
the rank that computed it (private RNG stream per trajectory).

Overlap: two groups of state machines take turns; the fused launch of one group is submitted asynchronously
(``bildk_logl_runs_multi_submit``: pinned staging, no stream synchronisation) and runs while the host code of the other
group advances, in ONE thread.  (Round 1 tried this with a worker thread inside the GIL-free C call: the launch time
was hidden but the lane threads lost as much to GIL hand-overs as was gained.)
"""
import time

import numpy as np

from .core import sample_gen
from .engine import st_to_runs
from .trajectory import make_Trajectory

__all__ = ["sample_many", "store_claimer"]


class _Lane:
    """One trajectory's state machine: the `sample_gen` generator, its pending request and its private RNG state."""

    def __init__(self, idx, traj, seed):
        self.idx, self.traj, self.seed = idx, traj, seed
        self.gen = None
        self.request = None                      # (ss, thetas) waiting for evaluation
        self.amis = None                         # the AMIS bookkeeping that may ride on that launch (FusedAmisStep)
        self.answer = None
        self.result = None
        self.done = False
        self.rng_state = None

    def advance(self, model, sample_kw):
        """Run until the next likelihood request (stored in `request`) or the end (`result`, `done`)."""
        if self.gen is None:
            np.random.seed(self.seed)
            self.gen = sample_gen(self.traj, model, **sample_kw)
            send = None
        else:
            np.random.set_state(self.rng_state)
            send, self.answer = self.answer, None
        try:
            req = self.gen.send(send)
            ss, thetas = req
            self.request = (np.asarray(ss, dtype=float), np.asarray(thetas))
            self.amis = getattr(req, "amis", None)
        except StopIteration as stop:
            self.result = stop.value
            self.done = True
        self.rng_state = np.random.get_state()


def store_claimer(n_total, store=None, key="bild_b200/next_trajectory"):
    """
    ``claim(n) -> list of trajectory indices`` backed by an atomic counter in a `torch.distributed` key-value store
    (default: the store of the default process group).  Every index in ``range(n_total)`` is handed out exactly once
    across all ranks; an empty list means the dataset is used up.
    """
    if store is None:
        import torch.distributed as dist
        store = dist.distributed_c10d._get_default_store()

    def claim(n):
        hi = int(store.add(key, int(n)))
        return list(range(min(hi - n, n_total), min(hi, n_total)))
    return claim


def sample_many(trajs, model, seeds=None, rank=0, world=1, max_active=None, claim=None, fuse_amis=True, **sample_kw):
    """
    Run `sample` on every trajectory, fusing the likelihood batches of all concurrently active trajectories.

    Parameters
    ----------
    trajs : sequence of trajectories (anything `make_Trajectory` accepts)
    model : bild_b200.models.MultiStateRouse
    seeds : sequence of int, optional
        numpy seed of each trajectory's private RNG stream (default: ``range(len(trajs))``)
    rank, world : int
        static partition: this process handles trajectories ``rank, rank + world, ...`` (one process per GPU)
    max_active : int, optional
        upper bound on concurrently running state machines (default: all of this rank's; 64 with `claim`)
    claim : callable, optional
        dynamic partition: ``claim(n)`` returns up to ``n`` not yet assigned trajectory indices (`store_claimer`);
        overrides `rank` / `world`
    fuse_amis : bool
        let every sampler's device bookkeeping (`bildk_amis_step`) ride on the fused likelihood launch, in stream order
        behind the filter kernel, instead of one synchronous call per sampler step (same arithmetic, same bits)
    **sample_kw : forwarded to `sample` (dE, init_runs, sampler_kw, ...)

    Returns
    -------
    dict  trajectory index -> SamplingResults   (this rank's trajectories)
    stats : dict  launches, profiles and frame-steps evaluated
    """
    trajs = [make_Trajectory(t) for t in trajs]
    if seeds is None:
        seeds = list(range(len(trajs)))
    mine = list(range(rank, len(trajs), world)) if claim is None else []
    outer_rng = np.random.get_state()
    stats = {"launches": 0, "profiles": 0, "frame_steps": 0, "rounds": 0,
             "t_host_lanes": 0.0, "t_pack": 0.0, "t_gpu": 0.0}      # wall seconds: AMIS host code / run-length packing / fused launches
    pending = [_Lane(i, trajs[i], seeds[i]) for i in mine]
    limit = max_active or (len(pending) if claim is None else 64) or 1
    results = {}
    more = claim is not None
    # Two groups of state machines take turns when the model can launch asynchronously (C ABI bildk_logl_runs_multi_submit /
    # bildk_logl_wait: pinned staging, nothing in the submit path waits for the GPU): while the fused launch of one group
    # runs, the host code of the other group advances - single-threaded, no GIL hand-overs.
    n_groups = 2 if (hasattr(model, "logL_runs_multi_submit") and limit > 1) else 1
    per_group = -(-limit // n_groups)
    groups = [[] for _ in range(n_groups)]
    inflight = [[] for _ in range(n_groups)]     # per group: (pending batch, lanes, offsets)
    g = n_groups - 1
    try:
        while pending or more or any(groups) or any(inflight):
            g = (g + 1) % n_groups
            active = groups[g]
            # ---- answers of this group's launch (the other group's host code ran meanwhile)
            tic = time.perf_counter()
            for batch, lanes, offsets in inflight[g]:
                out = batch.wait() if hasattr(batch, "wait") else batch
                for ln, lo, hi in zip(lanes, offsets[:-1], offsets[1:]):
                    ln.answer = out[lo:hi]
                    ln.request = None
            inflight[g] = []
            stats["t_gpu"] += time.perf_counter() - tic
            # ---- top up this group (static list first, then the shared counter)
            room = per_group - len(active)
            if more and len(pending) < room:
                got = claim(room - len(pending))
                pending.extend(_Lane(i, trajs[i], seeds[i]) for i in got)
                more = bool(got)
            while pending and len(active) < per_group:
                active.append(pending.pop(0))
            # ---- let every lane of the group run (one at a time, fixed order) until it asks for likelihoods or finishes
            tic = time.perf_counter()
            for lane in active:
                if lane.request is None and not lane.done:
                    lane.advance(model, sample_kw)
            stats["t_host_lanes"] += time.perf_counter() - tic
            for lane in [ln for ln in active if ln.done]:
                results[lane.idx] = lane.result
                active.remove(lane)
            waiting = [ln for ln in active if ln.request is not None]
            if not waiting:
                continue
            # ---- fuse: one launch for all waiting trajectories that share a localisation error (the fused kernel takes
            #      ONE error structure per launch; with ``model.localization_error`` set that is a single batch); profiles
            #      with fewer runs are padded with empty runs (start = T), which vanish exactly like the empty slices
            #      of st2profile
            stats["rounds"] += 1
            by_noise = {}
            for ln in waiting:
                key = tuple(np.asarray(model._get_noise(ln.traj), dtype=float).ravel())
                by_noise.setdefault(key, []).append(ln)
            for lanes in by_noise.values():
                tic = time.perf_counter()
                K1 = max(ln.request[0].shape[1] for ln in lanes)
                starts, states, offsets = [], [], [0]
                for ln in lanes:
                    ss, thetas = ln.request
                    if thetas.size and (thetas.min() < 0 or thetas.max() >= model.nStates):
                        raise ValueError("state index out of range")
                    a, b = st_to_runs(ss, thetas, len(ln.traj))
                    if a.shape[1] < K1:
                        extra = K1 - a.shape[1]
                        a = np.concatenate([a, np.full((len(a), extra), len(ln.traj), dtype=a.dtype)], axis=1)
                        b = np.concatenate([b, np.repeat(b[:, -1:], extra, axis=1)], axis=1)
                    starts.append(a)
                    states.append(b)
                    offsets.append(offsets[-1] + len(a))
                    stats["frame_steps"] += len(a) * (len(ln.traj) - 1)
                all_starts, all_states = np.concatenate(starts), np.concatenate(states)
                stats["t_pack"] += time.perf_counter() - tic
                tic = time.perf_counter()
                if n_groups > 1 and len(inflight[g]) == 0:
                    batch = model.logL_runs_multi_submit([ln.traj for ln in lanes], offsets, all_starts, all_states,
                                                         amis=[ln.amis for ln in lanes] if fuse_amis else None)
                else:   # synchronous models, and the (rare) second localisation-error batch of a round: one slot per group
                    batch = model.logL_runs_multi([ln.traj for ln in lanes], offsets, all_starts, all_states)
                inflight[g].append((batch, lanes, offsets))
                stats["t_gpu"] += time.perf_counter() - tic
                stats["launches"] += 1
                stats["profiles"] += offsets[-1]
    finally:
        np.random.set_state(outer_rng)      # also when a lane or a launch raises: the caller's RNG stream is not ours to keep
    return results, stats
