"""
ORACLE SHIM (test infrastructure): the slice of ``noctiluca`` the reference touches
(models.py:17, 347-350; core.py:9, 111; MSRouse_logL.pyx:171-178; tests/test_bild.py:125).
"""
import numpy as np


class Trajectory:
    def __init__(self, data, localization_error=None, **meta):
        data = np.array(data, dtype=float)
        if data.ndim == 1:
            data = data[:, None]
        self.data = data
        self.localization_error = None if localization_error is None else np.asarray(localization_error, dtype=float)
        self.meta = dict(meta)

    def __len__(self):
        return self.data.shape[0]

    @property
    def T(self):
        return len(self)

    @property
    def d(self):
        return self.data.shape[1]

    def __getitem__(self, key):
        return self.data[key]

    def abs(self):
        return Trajectory(np.sqrt(np.sum(self.data ** 2, axis=1)), localization_error=None, **self.meta)

    def count_valid_frames(self):
        return int(np.count_nonzero(~np.any(np.isnan(self.data), axis=1)))


def make_Trajectory(x, **kw):
    if isinstance(x, Trajectory) or (hasattr(x, "localization_error") and hasattr(x, "__len__")):
        return x
    return Trajectory(x, **kw)
