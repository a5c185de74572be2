"""
bild_b200 - B200-native likelihood engine for BILD (Bayesian Inference of Looping Dynamics).

Drop-in for the hot path of OpenTrajectoryAnalysis/bild: ``import bild_b200 as bild`` gives the same
public surface (/root/reference/bild/__init__.py:12-17: ``util``, ``Loopingprofile``, ``models``,
``amis``, ``postproc``, ``sample``, ``SamplingResults``) with the multi-state-Rouse Kalman-filter
likelihood evaluated in batches by hand-written sm_100a kernels (bild_b200/csrc) through a C ABI
(include/bild_b200.h).  No CPU fallback.
"""
from . import util
from .util import Loopingprofile
from . import rouse
from . import models
from . import amis
from . import choicesampler
from . import postproc
from .core import sample, SamplingResults
from .trajectory import Trajectory, make_Trajectory
from .engine import RouseEngine, TrajectoryHandle, st_to_runs, states_to_runs

__version__ = "0.1.0"
__all__ = ["util", "Loopingprofile", "rouse", "models", "amis", "choicesampler", "postproc", "sample",
           "SamplingResults", "Trajectory", "make_Trajectory", "RouseEngine", "TrajectoryHandle"]
