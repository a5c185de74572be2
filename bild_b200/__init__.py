"""
bild_b200 - B200-native likelihood engine for BILD (Bayesian Inference of Looping Dynamics).

Drop-in for the hot path of OpenTrajectoryAnalysis/bild: batched multi-state-Rouse Kalman-filter
log-likelihoods on sm_100a behind the reference's own Python API.
"""
from . import rouse  # noqa: F401
from .engine import RouseEngine, TrajectoryHandle, st_to_runs, states_to_runs  # noqa: F401

__version__ = "0.1.0"
