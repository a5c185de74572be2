"""
ORACLE - TEST INFRASTRUCTURE.  Stages the UNMODIFIED reference package where the GPU box can import it.

/root/reference does not exist on the GPU box; `baseline/_ref*/` is git-ignored but travels with the gpurun snapshot
(like our own built .so files).  Run by `__graft_entry__.build()` in the build container:

  baseline/_ref/bild/        copy of /root/reference/bild, plus  bin/__init__.py  and  bin/MSRouse_logL<EXT>.so
                             = the reference's own .pyx compiled by oracle/Makefile  ->  the reference AS SHIPPED
                             (its plugin slot cython_imports.py:3-7 filled with its own Cython code).  This is the
                             CPU arm of bench.py (`--impl reference`, `cpu_baseline`, the `bild.sample` wall).
  baseline/_ref_b200/bild/   the same unmodified copy, with  bin/MSRouse_logL.py = integration/MSRouse_logL.py
                             (the ctypes binding of libbild_b200.so) in that slot -> tests/test_gpu_integration.py.

The three third-party packages the reference imports but does not vendor (rouse, noctiluca, bayesmsd) come from
oracle/shims/ (ours, tracked) - put `oracle/shims` on sys.path before importing either tree.
No reference file is modified and none enters the git history.
"""
import glob
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"


def stage(verbose=False):
    if not os.path.isdir(os.path.join(REF, "bild")):
        return False
    built = sorted(glob.glob(os.path.join(HERE, "_ref", "MSRouse_logL*.so")))
    for name, plug in (("_ref", built[0] if built else None),
                       ("_ref_b200", os.path.join(ROOT, "integration", "MSRouse_logL.py"))):
        dst = os.path.join(ROOT, "baseline", name)
        pkg = os.path.join(dst, "bild")
        if os.path.isdir(pkg):
            shutil.rmtree(pkg)
        os.makedirs(dst, exist_ok=True)
        shutil.copytree(os.path.join(REF, "bild"), pkg, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        os.chmod(pkg, 0o755)
        for d, _, files in os.walk(pkg):
            os.chmod(d, 0o755)
            for f in files:
                os.chmod(os.path.join(d, f), 0o644)
        os.makedirs(os.path.join(pkg, "bin"), exist_ok=True)
        open(os.path.join(pkg, "bin", "__init__.py"), "w").close()
        if plug:
            shutil.copy(plug, os.path.join(pkg, "bin", os.path.basename(plug)))
        if verbose:
            print("staged", pkg, "<-", plug)
    return True


def import_reference(which="_ref"):
    """Import the staged reference package (fresh) and return the module; `which` = "_ref" or "_ref_b200"."""
    import importlib
    import warnings
    base = os.path.join(ROOT, "baseline", which)
    if not os.path.isdir(os.path.join(base, "bild")):
        raise ImportError(f"{base}/bild is not staged - run `python -c 'import __graft_entry__ as g; g.build()'` in the build container")
    for name in [m for m in sys.modules if m == "bild" or m.startswith("bild.")]:
        del sys.modules[name]
    shims = os.path.join(HERE, "shims")
    for p in (shims, base):
        if p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, shims)
    sys.path.insert(0, base)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mod = importlib.import_module("bild")
    sys.path.remove(base)
    return mod


if __name__ == "__main__":
    print("staged" if stage(verbose=True) else "no /root/reference here: nothing staged")
