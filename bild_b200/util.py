"""
Looping profiles and small helpers.

API-compatible with /root/reference/bild/util.py (``Loopingprofile``, ``state_probabilities``); the
semantics that define the kernel's input format are util.py:15-23: ``profile[t]`` is the state used
to propagate TO frame ``t``; ``profile[0]`` selects the steady-state ensemble the trajectory starts in.
"""
import numpy as np

__all__ = ["Loopingprofile", "state_probabilities"]


class Loopingprofile:
    """Integer state per frame.  ``len``, item get/set, ``==``, ``copy``, switch counting, intervals."""

    __hash__ = None

    def __init__(self, states=None):
        self.state = np.zeros(0, dtype=int) if states is None else np.asarray(states, dtype=int)

    def copy(self):
        dup = Loopingprofile()
        dup.state = self.state.copy()
        return dup

    def __len__(self):
        return self.state.shape[0]

    def __getitem__(self, key):
        return self.state[key]

    def __setitem__(self, key, val):
        val = np.asarray(val)
        # refuse floats instead of silently truncating them: that is almost always a bug upstream
        assert np.issubdtype(val.dtype, np.integer)
        self.state[key] = val

    def __eq__(self, other):
        try:
            return len(self) == len(other) and bool(np.all(self.state == other.state))
        except Exception:
            return False

    def count_switches(self):
        """Number of frames whose state differs from the previous frame's."""
        return int(np.count_nonzero(self.state[1:] != self.state[:-1]))

    def runs(self):
        """Run-length code: ``(starts, states)`` with ``starts[0] == 0`` (what the GPU consumes)."""
        if len(self) == 0:
            return np.zeros(0, dtype=int), np.zeros(0, dtype=int)
        starts = np.concatenate([[0], np.flatnonzero(self.state[1:] != self.state[:-1]) + 1])
        return starts, self.state[starts]

    def intervals(self):
        """``[(start, end, state), ...]``; ``start``/``end`` are ``None`` at the open ends."""
        starts, states = self.runs()
        edges = [None] + [int(s) for s in starts[1:]] + [None]
        return [(edges[i], edges[i + 1], states[i]) for i in range(len(states))]

    def plottable(self):
        """``t, y`` arrays tracing the profile as a step function (frame ``i`` spans ``(i-1, i]``)."""
        starts, states = self.runs()
        ends = np.concatenate([starts[1:], [len(self)]])
        t = np.stack([starts, ends], axis=-1).ravel() - 1
        y = np.repeat(states, 2)
        return t, y


def state_probabilities(profiles, nStates=None):
    """Fraction of ``profiles`` in each state at each frame: ``(nStates, T)``."""
    stack = np.array([p[:] for p in profiles])
    if nStates is None:
        nStates = int(stack.max()) + 1
    occupancy = stack[None, :, :] == np.arange(nStates)[:, None, None]
    return occupancy.sum(axis=1) / stack.shape[0]
