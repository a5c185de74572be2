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

Scheduling (models that launch asynchronously): event driven, most-advanced trajectory first.  The state machines whose
likelihoods have arrived wait in a priority queue ordered by the number of AMIS steps they have taken; the driver advances
the first one (host code of ONE step), adds its request to the pending batch, and hands the pending batch to the library
whenever one of its two slots is free (``bildk_logl_runs_multi_submit``; finished batches are noticed with the
non-blocking ``bildk_logl_ready``).  The cost of a trajectory is data dependent with a heavy tail (20 to 950 AMIS steps in
the 1024-trajectory run); in the round-based scheme below every trajectory advances one step per round of ALL active
trajectories, so the longest one sets the wall time (8 GPUs: slowest rank 20.5 s, fastest 9.5 s).  Here a trajectory that
has come far runs at the speed of its own critical path (host step + one launch) while the others fill the gaps.

Round-based scheme (``schedule="rounds"``, and all synchronous models): two groups of state machines take turns; the fused launch of one group is submitted asynchronously
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
        self.steps = 0                           # likelihood batches answered so far (scheduling priority)
        self.host_s, self.n_advance = 0.0, 0     # host seconds / calls spent in this lane's state machine
        self.rng_state = None

    def advance(self, model, sample_kw):
        """Run until the next likelihood request (stored in `request`) or the end (`result`, `done`)."""
        tic = time.perf_counter()
        try:
            self._advance(model, sample_kw)
        finally:
            self.host_s += time.perf_counter() - tic
            self.n_advance += 1

    def _advance(self, model, sample_kw):
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


def _pack(lanes, model, stats):
    """Run-length arrays of the requests of `lanes` (one localisation error): (offsets, starts, states).  Profiles with
    fewer runs are padded with empty runs (start = T), which vanish exactly like the empty slices of st2profile."""
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
    return offsets, np.concatenate(starts), np.concatenate(states)


def _noise_key(model, traj):
    return tuple(np.asarray(model._get_noise(traj), dtype=float).ravel())


def _sample_many_priority(trajs, model, seeds, pending_lanes, claim, limit, fuse_amis, sample_kw, stats):
    """Event-driven driver (module docstring): most-advanced state machine first, at most two batches in flight."""
    import heapq
    results = {}
    more = claim is not None
    n_active = 0
    ready = []                  # heap of (-steps taken, trajectory index, lane): lanes whose answer has arrived / new lanes
    waiting = []                # lanes with a request that has not been submitted yet
    inflight = []               # (batch, lanes, offsets), oldest first; the library holds at most two
    while True:
        # ---- top up the set of running state machines (static list first, then the shared counter)
        room = limit - n_active
        if more and room > len(pending_lanes):
            got = claim(room - len(pending_lanes))
            pending_lanes.extend(_Lane(i, trajs[i], seeds[i]) for i in got)
            more = bool(got)
        while pending_lanes and n_active < limit:
            ln = pending_lanes.pop(0)
            ln.steps = 0
            heapq.heappush(ready, (0, ln.idx, ln))
            n_active += 1
        if n_active == 0:
            break
        # ---- answers: take every finished batch; block only when there is nothing else to do
        tic = time.perf_counter()
        while inflight:
            done = [i for i, item in enumerate(inflight) if item[0].ready()]      # the two slots may finish in either order
            if not done and (ready or (waiting and len(inflight) < 2)):
                break
            batch, lanes, offsets = inflight.pop(done[0] if done else 0)
            out = batch.wait()
            for ln, lo, hi in zip(lanes, offsets[:-1], offsets[1:]):
                ln.answer = out[lo:hi]
                ln.request = None
                heapq.heappush(ready, (-ln.steps, ln.idx, ln))
        stats["t_gpu"] += time.perf_counter() - tic
        # ---- host code of ONE step of the most advanced ready lane
        if ready:
            tic = time.perf_counter()
            _, _, lane = heapq.heappop(ready)
            lane.advance(model, sample_kw)
            lane.steps += 1
            stats["t_host_lanes"] += time.perf_counter() - tic
            if lane.done:
                results[lane.idx] = lane.result
                stats["lanes"].append((lane.idx, lane.n_advance, round(lane.host_s, 4)))
                n_active -= 1
            else:
                waiting.append(lane)
        # ---- a free slot takes everything that is waiting (one localisation error per launch)
        if waiting and len(inflight) < 2:
            tic = time.perf_counter()
            key = _noise_key(model, waiting[0].traj)
            lanes = [ln for ln in waiting if _noise_key(model, ln.traj) == key] if model.localization_error is None else waiting
            waiting = [ln for ln in waiting if ln not in lanes] if len(lanes) < len(waiting) else []
            offsets, starts, states = _pack(lanes, model, stats)
            stats["t_pack"] += time.perf_counter() - tic
            tic = time.perf_counter()
            batch = model.logL_runs_multi_submit([ln.traj for ln in lanes], offsets, starts, states,
                                                 amis=[ln.amis for ln in lanes] if fuse_amis else None)
            inflight.append((batch, lanes, offsets))
            stats["t_gpu"] += time.perf_counter() - tic
            stats["t_submit"] += time.perf_counter() - tic
            stats["launches"] += 1
            stats["rounds"] += 1
            stats["profiles"] += offsets[-1]
    return results


def sample_many(trajs, model, seeds=None, rank=0, world=1, max_active=None, claim=None, fuse_amis=True, schedule="priority",
                **sample_kw):
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
    schedule : "priority" (default) or "rounds"
        for models that launch asynchronously: event driven, most-advanced trajectory first - or the round-based scheme
        of two alternating groups (module docstring); synchronous models always run in rounds
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
             "t_host_lanes": 0.0, "t_pack": 0.0, "t_gpu": 0.0, "t_submit": 0.0,   # wall seconds: AMIS host code / run-length packing / fused launches
             "lanes": []}                                            # per finished trajectory: (index, likelihood batches, host seconds)
    pending = [_Lane(i, trajs[i], seeds[i]) for i in mine]
    limit = max_active or (len(pending) if claim is None else 64) or 1
    if hasattr(model, "logL_runs_multi_submit") and schedule == "priority":
        try:
            return _sample_many_priority(trajs, model, seeds, pending, claim, limit, fuse_amis, sample_kw, stats), stats
        finally:
            np.random.set_state(outer_rng)
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
                stats["lanes"].append((lane.idx, lane.n_advance, round(lane.host_s, 4)))
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
                by_noise.setdefault(_noise_key(model, ln.traj), []).append(ln)
            for lanes in by_noise.values():
                tic = time.perf_counter()
                offsets, all_starts, all_states = _pack(lanes, model, stats)
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
