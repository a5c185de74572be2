"""ORACLE SHIM (test infrastructure): import-time surface of ``bayesmsd`` (models.py:21-22); the
GenericGaussianModel that uses it is out of scope."""
from . import gp, deco  # noqa: F401
