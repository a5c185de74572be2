"""
Build libbild_b200.so (CUDA kernels + C ABI) in-tree for sm_100a.

    python -m bild_b200.build [--force]

nvcc cross-compiles without a GPU.  The library sits next to this file so that it travels with a
snapshot of the repository; it is git-ignored.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = [os.path.join(HERE, "csrc", "bildk.cu")]
DEPS = SRC + [os.path.join(HERE, "csrc", "bildk_kernels.cuh"), os.path.join(HERE, "csrc", "bildk_mma.cuh"),
              os.path.join(HERE, "csrc", "bildk_mmar.cuh"), os.path.join(HERE, "csrc", "bildk_mmar2.cuh"), os.path.join(HERE, "csrc", "bildk_mmarb.cuh"), os.path.join(HERE, "csrc", "bildk_mmag2.cuh"), os.path.join(HERE, "csrc", "bildk_mmact.cuh"),
              os.path.join(HERE, "csrc", "bildk_amis.cuh"),
              os.path.join(os.path.dirname(HERE), "include", "bild_b200.h")]
OUT = os.path.join(HERE, "libbild_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "--use_fast_math=false"]


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
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")]
    # one translation unit, ~40 kernel instantiations: let nvcc optimise them in parallel (3.7 min -> 1 min on 8 cores)
    flags += ["--split-compile", str(min(16, os.cpu_count() or 1))]
    cmd = [nvcc] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + SRC
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
