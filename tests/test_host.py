"""
CPU tests of the host logic and of the C-ABI library surface (no compute calls - there is no GPU
here and no CPU fallback by design).
"""
import ctypes
import os
import re

import numpy as np
import pytest

import kalman_oracle as ko
from helpers import ROOT


# ------------------------------------------------------------------ C ABI surface
def test_library_exports_every_declared_symbol():
    from bild_b200 import _lib
    header = open(os.path.join(ROOT, "include", "bild_b200.h")).read()
    declared = set(re.findall(r"\b(bildk_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/bild_b200.h but not exported"
    assert declared == set(_lib.SYMBOLS), "ctypes prototypes and header diverge"
    assert _lib.load().bildk_version() == 1000


def test_no_cpu_fallback_without_device():
    from bild_b200 import _lib
    from bild_b200.engine import RouseEngine
    lib = _lib.load()
    if lib.bildk_device_count() > 0:
        pytest.skip("a GPU is present")
    eye = np.eye(4)[None]
    with pytest.raises(_lib.BildkError, match="no CUDA device"):
        RouseEngine(eye, np.zeros((1, 4, 1)), eye, np.zeros((1, 4, 1)), eye, np.array([-1.0, 0, 0, 1]))


def test_argument_validation_precedes_device_use():
    from bild_b200.engine import RouseEngine
    eye = np.eye(4)[None]
    with pytest.raises(ValueError):
        RouseEngine(eye, np.zeros((1, 4, 1)), np.eye(3)[None], np.zeros((1, 4, 1)), eye, np.ones(4))
    bad = eye.copy(); bad[0, 0, 0] = np.nan
    with pytest.raises(ValueError, match="non-finite"):
        RouseEngine(bad, np.zeros((1, 4, 1)), eye, np.zeros((1, 4, 1)), eye, np.ones(4))


def test_product_does_not_import_oracle():
    """The product path must never route through the oracle (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "bild_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "kalman_oracle" not in src and "rouse_oracle" not in src and "oracle/" not in src.replace("# oracle/", ""), f


# ------------------------------------------------------------------ profile coding
def test_st_to_runs_matches_reference_st2profile(golden_dir):
    from bild_b200.engine import st_to_runs
    g = np.load(os.path.join(golden_dir, "st2profile.npz"))
    for i in range(len(g["k1"])):
        k1, T = int(g["k1"][i]), int(g["T"][i])
        want = g["states_flat"][g["offsets"][i]:g["offsets"][i + 1]]
        starts, rst = st_to_runs(g["ss"][i, :k1], g["thetas"][i, :k1], T)
        got = np.empty(T, dtype=int)
        bounds = list(starts[0]) + [T]
        for r in range(k1):          # later runs overwrite nothing: [start_r, start_{r+1}) are disjoint
            got[bounds[r]:max(bounds[r], bounds[r + 1])] = rst[0, r]
        got[bounds[-2]:] = rst[0, -1]
        assert np.array_equal(got, want), (i, g["ss"][i, :k1], g["thetas"][i, :k1])
        assert np.array_equal(ko.st2states(g["ss"][i, :k1], g["thetas"][i, :k1], T), want)


def test_st2profile_probe_cases():
    """SURVEY.md appendix A / tests/test_amis.py:199-202."""
    from bild_b200.engine import st_to_runs

    def expand(s, th, T=6):
        starts, rst = st_to_runs(np.array(s, dtype=float), np.array(th), T)
        out = np.empty(T, dtype=int)
        b = list(starts[0]) + [T]
        for r in range(len(s)):
            out[b[r]:b[r + 1]] = rst[0, r]
        return "".join(map(str, out))

    assert expand([.25, .5, .25], [0, 1, 0]) == "001100"
    assert expand([0, .5, .5], [0, 1, 0]) == "011000"
    assert expand([.5, 0, .5], [0, 1, 0]) == "000000"
    assert expand([.5, .5, 0], [0, 1, 0]) == "000111"
    assert expand([.999, .0005, .0005], [0, 1, 0]) == "000000"


def test_states_to_runs_roundtrip():
    from bild_b200.engine import states_to_runs
    rng = np.random.default_rng(0)
    for T in (1, 2, 17, 200):
        states = rng.integers(0, 3, size=(23, T))
        states[0] = 1                                  # a constant profile
        starts, rst = states_to_runs(states)
        K1 = starts.shape[1]
        for p in range(len(states)):
            b = list(starts[p]) + [T]
            rec = np.empty(T, dtype=int)
            for r in range(K1):
                rec[b[r]:max(b[r], b[r + 1])] = rst[p, r]
            assert np.array_equal(rec, states[p])
        assert starts[:, 0].max() == 0


# ------------------------------------------------------------------ util / trajectory / models (host side)
def test_loopingprofile_api():
    """Behaviour pinned by /root/reference/tests/test_bild.py:51-107."""
    from bild_b200.util import Loopingprofile, state_probabilities
    lp = Loopingprofile()
    assert len(lp) == 0
    prof = Loopingprofile([0, 0, 0, 1, 1, 0, 3, 3])
    assert len(prof) == 8 and prof.count_switches() == 3
    assert prof.copy() == prof and prof.copy() is not prof
    assert not (prof == Loopingprofile([0, 0, 0, 1, 1, 0, 3, 2]))
    assert not (prof == Loopingprofile([0, 1]))
    assert prof.intervals() == [(None, 3, 0), (3, 5, 1), (5, 6, 0), (6, None, 3)]
    t, y = prof.plottable()
    assert np.array_equal(t, [-1, 2, 2, 4, 4, 5, 5, 7]) and np.array_equal(y, [0, 0, 1, 1, 0, 0, 3, 3])
    prof[2] = 1
    assert prof[2] == 1
    with pytest.raises(AssertionError):
        prof[2] = 1.5
    sp = state_probabilities([Loopingprofile([0, 0, 1]), Loopingprofile([0, 1, 1])])
    assert np.array_equal(sp, [[1, .5, 0], [0, .5, 1]])
    assert state_probabilities([Loopingprofile([0, 0])], nStates=3).shape == (3, 2)


def test_models_host_side():
    from bild_b200.models import FactorizedModel, MultiStateModel, MultiStateRouse
    from bild_b200.trajectory import Trajectory
    from bild_b200.util import Loopingprofile
    import scipy.stats
    np.random.seed(12)
    traj = Trajectory(np.array([1, 2, np.nan, 4]), localization_error=[0.5])
    model = MultiStateRouse(20, 1, 5, d=1)
    assert model.nStates == 2 and model.d == 1
    assert np.array_equal(model.transitions, [[False, True], [True, False]])
    assert np.array_equal(model.measurement[[0, -1]], [-1, 1]) and not model.measurement[1:-1].any()
    assert len(MultiStateModel.initial_loopingprofile(model, traj)) == 4
    assert np.array_equal(model._get_noise(traj), [0.5])
    t2 = Trajectory(np.array([1, 2, np.nan, 4]))
    with pytest.raises(ValueError):
        model._get_noise(t2)
    model = MultiStateRouse(20, 1, 5, d=1, localization_error=0.5)
    assert np.array_equal(model.initial_loopingprofile(traj).state, [1, 0, 0, 0])       # test_bild.py:150-151
    tr = model.trajectory_from_loopingprofile(Loopingprofile([0, 0, 0, 1, 1, 1]), localization_error=0.1)
    assert len(tr) == 6 and tr.d == 1
    tr = model.trajectory_from_loopingprofile(Loopingprofile(np.ones(20, dtype=int)), localization_error=0.1, missing_frames=0.9)
    assert tr.count_valid_frames() < 18
    tr = model.trajectory_from_loopingprofile(Loopingprofile(np.ones(20, dtype=int)), missing_frames=12)
    assert tr.count_valid_frames() == 8
    fm = FactorizedModel([scipy.stats.maxwell(scale=1), scipy.stats.maxwell(scale=4)], d=1)
    v = fm.logL(Loopingprofile([1, 1, 0, 0]), traj)
    assert -100 < v < 0
    assert np.array_equal(fm.initial_loopingprofile(traj).state, [0, 0, 1, 1])          # test_bild.py:183-186
    fm.clear_memo()
    assert fm.logL(Loopingprofile([1, 1, 0, 0]), traj) == v
    assert len(fm.trajectory_from_loopingprofile(Loopingprofile([0, 0, 0, 1, 1, 1]))) == 6


def test_register_chained_kernel_tables():
    """Host tables of k_mmar (bildk_debug_tables, no GPU): for every admissible (GT, r, ncols) the mean rows and the
    zero row are distinct spare rows of the last tile-row block, the covariance rows sit in the even B-fragment slots,
    and the placement keeps the 128-bit (lane pairs: rows of different parity) and 64-bit (lane quads: rows distinct
    mod 4) fragment loads free of bank conflicts for the d = 3 shapes of the BASELINE workloads (with fewer mean columns
    several unused slots share the single zero row, which cannot differ in parity from all of their partners)."""
    from bild_b200 import _lib
    lib = _lib.load()
    out = np.zeros(12, dtype=np.uint8)
    for GT in (1, 2, 3, 4):
        base = 8 * (GT - 1)
        for r in (1, 2, 3, 4):
            for ncols in (1, 2, 3, 4):
                assert lib.bildk_debug_tables(0, GT, r, ncols, _lib.ptr(out, _lib.c_uint8_p)) == 12
                lastrow, mrow = out[:8].astype(int), out[8:].astype(int)
                spare = set(range(base + r, base + 8))
                m = list(mrow[:ncols])
                assert len(set(m)) == ncols and set(m) <= spare                      # mean rows: distinct spare rows
                for i in range(4):
                    assert lastrow[2 * i + 1] == (m[i] if i < ncols else lastrow[2 * i + 1])
                    if i < r:
                        assert lastrow[2 * i] == base + i                             # covariance rows in the even slots
                unused = [lastrow[2 * i] for i in range(r, 4)] + [lastrow[2 * i + 1] for i in range(ncols, 4)]
                if unused:
                    z = set(unused)
                    assert len(z) == 1 and z <= spare and not (z & set(m))          # one zero row, never a mean row
                # bank conflicts (row stride == 8 mod 16 doubles, column bit 2 flipped by (row >> 1) & 1)
                pair_conf = sum(1 for i in range(4) if lastrow[2 * i] != lastrow[2 * i + 1] and (lastrow[2 * i] - lastrow[2 * i + 1]) % 2 == 0)
                quad_conf = sum(1 for h in range(2) for a in range(4) for b in range(a + 1, 4)
                                if lastrow[4 * h + a] != lastrow[4 * h + b] and (lastrow[4 * h + a] - lastrow[4 * h + b]) % 4 == 0)
                if (r, ncols) in ((4, 3), (2, 3), (1, 3)):    # shapes of the BASELINE workloads (d = 3; N = 20, 10 / 50, 25)
                    assert pair_conf == 0 and quad_conf == 0, (GT, r, ncols, lastrow)
    # mean in an extra row block (r = N - 8 (GT - 1) in 5..8; k_mmar GT <= 4, k_mmar2 GT 5..8): zero row 8 GT in the even
    # slots, mean row q = 8 GT + 2 q + 1 in the odd ones - the two rows of every lane pair differ in parity
    for GT in range(1, 14):
        for r in (5, 6, 7, 8):
            for ncols in (1, 2, 3, 4):
                assert lib.bildk_debug_tables(0, GT, r, ncols, _lib.ptr(out, _lib.c_uint8_p)) == 12
                lastrow, mrow = out[:8].astype(int), out[8:].astype(int)
                assert all(lastrow[2 * i] == 8 * GT for i in range(4))
                assert [int(v) for v in mrow[:ncols]] == [8 * GT + 2 * q + 1 for q in range(ncols)]
                assert all(lastrow[2 * i + 1] == (mrow[i] if i < ncols else 8 * GT) for i in range(4))
                assert all(8 * GT <= v < 8 * GT + 8 for v in lastrow)
    assert lib.bildk_debug_tables(0, 14, 1, 1, _lib.ptr(out, _lib.c_uint8_p)) < 0      # out of the kernels' range
    assert lib.bildk_debug_tables(0, 3, 9, 1, _lib.ptr(out, _lib.c_uint8_p)) < 0


def test_tile_slot_tables():
    """Host tables of k_mmact: every upper tile (ti <= c) exactly once, at most MAXS = 7 slots and two segments
    (consecutive tile rows of one column) per warp, tile (0, GT-1) first on warp 0, the four schedulers (warp % 4)
    within three tiles of each other."""
    from bild_b200 import _lib
    lib = _lib.load()
    out = np.zeros(288, dtype=np.uint8)
    for GT in range(8, 15):
        assert lib.bildk_debug_tables(1, GT, 0, 0, _lib.ptr(out, _lib.c_uint8_p)) == 288
        t = out.reshape(16, 18).astype(int)
        nw = 8 if GT == 8 else (12 if GT == 9 else 16)
        seen = set()
        sched = [0, 0, 0, 0]
        for w in range(16):
            n, na = t[w, 0], t[w, 1]
            if w >= nw:
                assert n == 0
                continue
            assert 1 <= n <= 7 and 1 <= na <= n
            slots = [(t[w, 2 + 2 * i], t[w, 3 + 2 * i]) for i in range(n)]
            for seg in (slots[:na], slots[na:]):
                if seg:
                    assert len({c for _, c in seg}) == 1                              # one column per segment
                    assert [ti for ti, _ in seg] == list(range(seg[0][0], seg[0][0] + len(seg)))   # consecutive rows
            for ti, c in slots:
                assert 0 <= ti <= c < GT and (ti, c) not in seen
                seen.add((ti, c))
            sched[w % 4] += n
        assert len(seen) == GT * (GT + 1) // 2
        assert (t[0, 2], t[0, 3]) == (0, GT - 1)
        assert max(sched) - min(sched) <= 3, (GT, sched)      # e.g. GT = 13: 24 / 23 / 23 / 21 of 91 tiles


# ------------------------------------------------------------------ round-2 host-side behaviour
def test_engine_rebuilt_when_model_edited(monkeypatch):
    """The reference re-reads the dynamics and the measurement vector on every call (pyx:150-160): editing either
    between two likelihood calls must invalidate the GPU copy (no device needed: the engine factory is stubbed)."""
    import bild_b200 as bild
    from bild_b200 import engine as eng_mod
    built = []

    class Dummy:
        device = 0

    def fake(models, measurement, device=0):
        built.append((models[1]._dynamics["B"][0, 0], float(measurement[-1])))
        return Dummy()

    monkeypatch.setattr(eng_mod.RouseEngine, "from_models", staticmethod(fake))
    model = bild.models.MultiStateRouse(6, 1, 5, d=2, localization_error=0.3)
    e1 = model.engine
    assert model.engine is e1 and len(built) == 1
    model.models[1].k = 2.0                         # check_dynamics() notices the stale k and refreshes -> new dynamics dict
    e2 = model.engine
    assert e2 is not e1 and len(built) == 2 and built[1][0] != built[0][0]
    model.models[0].update_dynamics()               # explicit refresh: new dict, engine follows
    assert model.engine is not e2 and len(built) == 3
    e3 = model.engine
    model.measurement[-1] = 0.5
    assert model.engine is not e3 and built[-1][1] == 0.5
    assert model.engine is model.engine


def test_state_arrays_refuse_values_that_would_wrap():
    from bild_b200.engine import st_to_runs, states_to_runs
    with pytest.raises(ValueError):
        st_to_runs(np.array([[0.5, 0.5]]), np.array([[0, 256]]), 10)
    with pytest.raises(ValueError):
        states_to_runs(np.array([[0, 0, 300, 300]]))
    with pytest.raises(ValueError):
        states_to_runs(np.array([[0, -1, 0, 0]]))
    a, b = states_to_runs(np.array([[0, 0, 255, 255]]))
    assert b.tolist() == [[0, 255]] and a.tolist() == [[0, 2]]
