"""
Minimal trajectory container.

The reference uses ``noctiluca.Trajectory`` (third-party, not vendored).  The in-tree code only
touches a small surface of it (/root/reference/bild/models.py:347-350, 442, 467;
src/MSRouse_logL.pyx:171-178; core.py:111), which is what this class provides.  Any object with the
same surface - in particular a real ``noctiluca.Trajectory`` - is accepted everywhere instead.
"""
import numpy as np

__all__ = ["Trajectory", "make_Trajectory"]


class Trajectory:
    """``(T, d)`` float64 positions, NaN rows = missing frames, optional per-dimension localisation error."""

    def __init__(self, data, localization_error=None, **meta):
        data = np.array(data, dtype=float)
        if data.ndim == 1:
            data = data[:, None]
        if data.ndim != 2:
            raise ValueError("trajectory data must be (T,) or (T, d)")
        self.data = data
        self.localization_error = None if localization_error is None else np.asarray(localization_error, dtype=float)
        self.meta = dict(meta)

    def __len__(self):
        return self.data.shape[0]

    @property
    def T(self):
        return self.data.shape[0]

    @property
    def d(self):
        return self.data.shape[1]

    def __getitem__(self, key):
        return self.data[key]

    def abs(self):
        """Trajectory of Euclidean norms, shape ``(T, 1)``."""
        return Trajectory(np.linalg.norm(self.data, axis=1), **self.meta)

    def valid_frames(self):
        return ~np.any(np.isnan(self.data), axis=1)

    def count_valid_frames(self):
        return int(np.count_nonzero(self.valid_frames()))


def make_Trajectory(obj, **kwargs):
    """Pass trajectory-like objects through, wrap arrays."""
    if hasattr(obj, "localization_error") and hasattr(obj, "__len__") and hasattr(obj, "__getitem__"):
        return obj
    return Trajectory(obj, **kwargs)
