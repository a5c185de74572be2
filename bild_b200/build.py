"""
Build libbild_b200.so (CUDA kernels + C ABI) in-tree for sm_100a.

    python -m bild_b200.build [--force]

nvcc cross-compiles without a GPU.  The library sits next to this file so that it travels with a
snapshot of the repository; it is git-ignored.  The sources are several translation units (csrc/bildk.cu with the C ABI and most
kernels, csrc/bildk_tu_*.cu with the launchers of the register-chained kernel families) compiled in parallel and linked.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# translation units: the C ABI + most kernels, and one unit per family of register-chained kernels (their unrolled template
# instantiations are where NVVM spends its time; the units compile in parallel)
SRC = [os.path.join(CSRC, f) for f in ("bildk.cu", "bildk_tu_mmar.cu", "bildk_tu_mmar2.cu", "bildk_tu_mmar8.cu")]
# (source, object name, extra flags): k_mmar8 is compiled once per instantiation (NVVM needs about a minute for each)
UNITS = [(SRC[0], "bildk", []), (SRC[1], "bildk_tu_mmar", []), (SRC[2], "bildk_tu_mmar2", [])] + \
        [(SRC[3], f"bildk_tu_mmar8_{gt}_{mx}", [f"-DBILDK_MMAR8_GT={gt}", f"-DBILDK_MMAR8_MX={mx}"]) for gt in (13, 12, 11, 10) for mx in (1, 0)]
DEPS = SRC + [os.path.join(CSRC, f) for f in ("bildk_kernels.cuh", "bildk_mma.cuh", "bildk_mmar.cuh", "bildk_mmar2.cuh", "bildk_mmarb.cuh",
                                              "bildk_mmag2.cuh", "bildk_mmact.cuh", "bildk_amis.cuh", "bildk_launch.h")] + \
    [os.path.join(os.path.dirname(HERE), "include", "bild_b200.h")]
OUT = os.path.join(HERE, "libbild_b200.so")
OBJ_DIR = os.path.join(HERE, "csrc", "_obj")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC"]


# roofline micro-benchmark (measurement tooling, not the product): tools/fp64_peak.cu -> tools/libfp64peak.so
PEAK_SRC = os.path.join(os.path.dirname(HERE), "tools", "fp64_peak.cu")
PEAK_OUT = os.path.join(os.path.dirname(HERE), "tools", "libfp64peak.so")


def build_peak(force=False):
    if not force and os.path.exists(PEAK_OUT) and os.path.getmtime(PEAK_OUT) >= os.path.getmtime(PEAK_SRC):
        return PEAK_OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
           "-DFP64_PEAK_NO_MAIN", "-o", PEAK_OUT, PEAK_SRC]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    return PEAK_OUT


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(f) > t for f in DEPS)


def build(force=False, verbose=False):
    build_peak(force)
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(OBJ_DIR, exist_ok=True)
    threads = max(1, min(16, os.cpu_count() or 1) // 4)
    procs = []
    for src, name, extra in UNITS:   # all translation units at once; every nvcc splits its own optimisation over a few threads
        obj = os.path.join(OBJ_DIR, name + ".o")
        cmd = [nvcc] + NVCC_FLAGS + extra + ["--split-compile", str(threads)] + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
        procs.append((cmd, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    for cmd, obj, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0:
            for _, _, other in procs:
                if other.poll() is None:
                    other.kill()
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + out)
        if verbose:
            print(out)
        objs.append(obj)
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-o", OUT] + objs
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc (link) failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    for obj in objs:   # the objects are not needed again (and would travel with every snapshot of the repository)
        os.remove(obj)
    os.rmdir(OBJ_DIR)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
