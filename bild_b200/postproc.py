"""
Local optimisation of the boundaries of an inferred profile.

Mirror of /root/reference/bild/postproc.py.  The reference evaluates the 2k+1 candidate profiles of one
pass with 2k+1 separate ``model.logL`` calls (postproc.py:36-59); here they form ONE batch when the model
offers ``logL_batch`` (the GPU engine), which is the second caller of the hot path.
"""
import numpy as np

__all__ = ["logLR_boundaries", "BoundaryEliminationError", "optimize_boundary"]


def _logL_many(model, profiles_states, traj):
    if hasattr(model, "logL_batch"):
        return np.asarray(model.logL_batch(np.asarray(profiles_states, dtype=np.int32), traj), dtype=float)
    from .util import Loopingprofile
    return np.array([model.logL(Loopingprofile(st), traj) for st in profiles_states])


def logLR_boundaries(profile, traj, model):
    """``(k, 2)`` log-likelihood ratios for moving each of the k boundaries one frame left (``[:, 0]``) /
    right (``[:, 1]``); empty array for a profile without boundaries."""
    base = np.asarray(profile.state)
    boundaries = np.nonzero(np.diff(base))[0]          # boundary between frames b and b+1
    if len(boundaries) == 0:
        return np.array([])
    cand = np.repeat(base[None, :], 2 * len(boundaries) + 1, axis=0)
    for i, b in enumerate(boundaries):
        cand[2 * i, b] = base[b + 1]                   # move left
        cand[2 * i + 1, b + 1] = base[b]               # move right
    ll = _logL_many(model, cand, traj)
    return ll[:-1].reshape(len(boundaries), 2) - ll[-1]


class BoundaryEliminationError(Exception):
    pass


def optimize_boundary(profile, traj, model, max_iteration=10000):
    """
    Greedy ascent: repeatedly apply the single one-frame boundary move with the largest likelihood gain
    until none improves.  Raises `BoundaryEliminationError` if the best move would delete a boundary
    (sampling was probably not extensive enough) and ``RuntimeError`` after ``max_iteration`` moves.
    """
    cur = profile.copy()
    T = len(traj)
    for _ in range(max_iteration):
        logLR = logLR_boundaries(cur, traj, model)
        if len(logLR) == 0:
            break
        i, j = np.unravel_index(np.argmax(logLR), logLR.shape)
        if not logLR[i, j] > 0:
            break
        b = np.nonzero(np.diff(cur.state))[0][i]
        if ((j == 0 and (b == 0 or cur[b - 1] == cur[b + 1]))
                or (j == 1 and (b == T - 2 or cur[b + 2] == cur[b]))):
            raise BoundaryEliminationError(f"Trying to abolish boundary at {b}")
        cur[b + j] = cur[b + (1 - j)]
    else:
        raise RuntimeError(f"Exceeded max_iteration = {max_iteration}")
    return cur
