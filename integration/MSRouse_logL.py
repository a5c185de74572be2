"""
ctypes binding of libbild_b200.so for the REFERENCE's native plugin slot.

The reference resolves its likelihood at /root/reference/bild/cython_imports.py:3-7
(``from .bin.MSRouse_logL import MSRouse_logL``, built from bild/src/MSRouse_logL.pyx by setup.py:41-46) and calls
it from ``MultiStateRouse.logL`` (bild/models.py:278).  Dropping THIS file into the reference tree as
``bild/bin/MSRouse_logL.py`` (next to an empty ``bild/bin/__init__.py``) puts the B200 engine into that slot:
``MSRouse_logL(model, profile, traj) -> float`` keeps the signature of pyx:95, nothing else in the reference
changes.  `install_batched_hook` additionally routes ``FixedkSampler.logL`` (bild/amis.py:717-739) through ONE
launch per batch (the reference's loop calls the slot once per profile).

Nothing here imports bild_b200: the binding talks to the C ABI (include/bild_b200.h) only.  The library is looked
up in $BILD_B200_LIB, else next to this file, else in the bild_b200 package directory of this repository.
"""
import ctypes
import os
import weakref

import numpy as np

__all__ = ["MSRouse_logL", "MSRouse_logL_st_batch", "install_batched_hook"]


def _find_library():
    here = os.path.dirname(os.path.abspath(__file__))
    cands = [os.environ.get("BILD_B200_LIB"), os.path.join(here, "libbild_b200.so")]
    up = here
    for _ in range(6):                                   # .../baseline/_ref*/bild/bin -> repository root
        up = os.path.dirname(up)
        cands.append(os.path.join(up, "bild_b200", "libbild_b200.so"))
    for c in cands:
        if c and os.path.exists(c):
            return c
    raise ImportError("libbild_b200.so not found (set BILD_B200_LIB); there is no CPU fallback")


_lib = ctypes.CDLL(_find_library())
_dp = ctypes.POINTER(ctypes.c_double)
_lib.bildk_last_error.restype = ctypes.c_char_p


def _ck(rc):
    if rc:
        raise (ValueError if rc == -1 else RuntimeError)(_lib.bildk_last_error().decode())


def _p(a, t=_dp):
    return a.ctypes.data_as(t)


_models = {}     # id(model) -> (weakref, handle, [dynamics dicts], measurement copy)
_trajs = {}      # (id(model), id(traj)) -> (weakref, handle, fingerprint)
DEVICE = int(os.environ.get("BILD_B200_DEVICE", os.environ.get("LOCAL_RANK", 0)))


def _model_handle(model):
    """pyx:150-160, hoisted: once per model (rebuilt when a state's dynamics or the measurement vector changed)."""
    for m in model.models:
        m.check_dynamics()                                                           # pyx:152-153
    dyn = [m._dynamics for m in model.models]
    w = np.ascontiguousarray(model.measurement, dtype=float)
    hit = _models.get(id(model))
    if hit is not None and hit[0]() is model and len(hit[2]) == len(dyn) and all(a is b for a, b in zip(hit[2], dyn)) \
            and np.array_equal(hit[3], w):
        return hit[1]
    stack = lambda key: np.ascontiguousarray([m._dynamics[key] for m in model.models], dtype=float)   # noqa: E731
    ss = [m.steady_state() for m in model.models]                                   # pyx:160, every state
    B, G, Sig = stack("B"), stack("G"), stack("Sig")
    M0 = np.ascontiguousarray([s[0] for s in ss], dtype=float)
    C0 = np.ascontiguousarray([s[1] for s in ss], dtype=float)
    S, N, d = G.shape
    h = ctypes.c_void_p()
    _ck(_lib.bildk_model_create(N, d, S, _p(B), _p(G), _p(Sig), _p(M0), _p(C0), _p(w), DEVICE, ctypes.byref(h)))
    for key in [k for k in _trajs if k[0] == id(model)]:
        del _trajs[key]
    _models[id(model)] = (weakref.ref(model), h, dyn, w.copy())
    return h


def _traj_handle(model, traj):
    """pyx:144-147, 174-178: once per (model, trajectory, localisation error)."""
    mh = _model_handle(model)
    err = np.asarray(model._get_noise(traj), dtype=float)                            # models.py:255-263 (may raise ValueError)
    x = np.ascontiguousarray(traj[:], dtype=float)
    fp = (hash(x.tobytes()), hash(err.tobytes()), x.shape)
    key = (id(model), id(traj))
    hit = _trajs.get(key)
    if hit is not None and hit[2] == fp:
        return hit[1]
    uniq, cind = np.unique(err, return_inverse=True)                                # pyx:145-147
    s2 = np.ascontiguousarray(uniq * uniq)
    cind = np.ascontiguousarray(cind, dtype=np.uint32)
    h = ctypes.c_void_p()
    _ck(_lib.bildk_traj_create(mh, len(x), _p(x), len(s2), _p(s2), _p(cind, ctypes.POINTER(ctypes.c_uint32)), ctypes.byref(h)))
    _trajs[key] = (None, h, fp)
    return h


def MSRouse_logL(model, profile, traj):
    """Same signature and meaning as /root/reference/bild/src/MSRouse_logL.pyx:95 (a batch of one)."""
    st = np.ascontiguousarray(profile[:], dtype=np.int32)[None, :]
    if st.shape[1] != len(traj):
        raise ValueError("profile and trajectory lengths differ")
    out = np.empty(1)
    _ck(_lib.bildk_logl_states(_traj_handle(model, traj), 1, _p(st, ctypes.POINTER(ctypes.c_int32)), _p(out)))
    return float(out[0])


def MSRouse_logL_st_batch(model, ss, thetas, traj):
    """The whole batch bild/amis.py:735-739 loops over, in one launch; (s, theta) -> run-length profiles inside the library
    with the arithmetic of amis.py:687-688."""
    ss = np.ascontiguousarray(ss, dtype=np.float64)
    thetas = np.ascontiguousarray(thetas, dtype=np.int64)
    out = np.empty(len(ss))
    if len(ss):
        _ck(_lib.bildk_logl_st(_traj_handle(model, traj), ss.shape[0], ss.shape[1], _p(ss),
                               _p(thetas, ctypes.POINTER(ctypes.c_int64)), _p(out)))
    return out


def install_batched_hook(bild_module):
    """
    The three-line change of INTEGRATION.md as a runtime patch of the imported reference package: `FixedkSampler.logL`
    hands the whole ``(ss, thetas)`` batch to the engine when the model is a `MultiStateRouse`, and falls through to the
    reference's own loop for every other model.
    """
    amis, models = bild_module.amis, bild_module.models
    original = amis.FixedkSampler.logL

    def logL(self, ss, thetas):
        if isinstance(self.model, models.MultiStateRouse):
            return MSRouse_logL_st_batch(self.model, ss, thetas, self.traj)
        return original(self, ss, thetas)

    amis.FixedkSampler.logL = logL
    return original
