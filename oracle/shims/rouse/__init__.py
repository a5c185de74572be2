"""
ORACLE SHIM (test infrastructure): stand-in for the third-party ``rouse`` package, which the reference
imports (models.py:16) but which is not installed here and not vendored under /root/reference.
Only visible to processes that put oracle/shims on sys.path (oracle/make_golden.py, tests).
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from rouse_oracle import Model, twoLocusMSD  # noqa: E402,F401
sys.path.pop(0)
