"""Build libgsn.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

The library is plain CUDA C++ behind ``extern "C"`` entry points (include/gsn.h);
it does not link against torch.  nvcc cross-compiles without a GPU, so this runs
in the build container as well as on the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_DIR = os.path.dirname(PKG_DIR)
LIB_PATH = os.path.join(PKG_DIR, "libgsn.so")
OBJ_DIR = os.path.join(PKG_DIR, "csrc", "build")
CSRC = os.path.join(PKG_DIR, "csrc")
HEADERS = [
    os.path.join(CSRC, "gsn_device.cuh"),
    os.path.join(CSRC, "gsn_chol_dmma.cuh"),
    os.path.join(CSRC, "gsn_launch.h"),
    os.path.join(CSRC, "gsn_predict.cuh"),
    os.path.join(CSRC, "gsn_cluster.cuh"),
    os.path.join(CSRC, "gsn_onchip.cuh"),
    os.path.join(REPO_DIR, "include", "gsn.h"),
]
SOURCES = [os.path.join(CSRC, "gsn_api.cu"), os.path.join(CSRC, "gsn_kernels.cu"), os.path.join(CSRC, "gsn_predict.cu"),
           os.path.join(CSRC, "gsn_cluster.cu"), os.path.join(CSRC, "gsn_onchip.cu")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    # no implicit multiply-add contraction: every fused operation in the kernels is an explicit fma(), so the two
    # launch shapes (separately compiled instantiations) produce bit-identical results, and the unfused plugin
    # formulas round like the reference's numpy/XLA arithmetic
    "-fmad=false",
    "-Xcompiler", "-fPIC",
]
# translation units: the C ABI, four slices of likelihood-kernel instantiations ({shape} x {kinds}), two of the
# one-theta-per-cluster kernel and two of the prediction kernel, built in parallel
UNITS = [
    ("gsn_api", "gsn_api.cu", []),
    ("gsn_wide_a", "gsn_kernels.cu", ["-DGSN_SLICE_CFG=CfgWide", "-DGSN_SLICE_NAME=launch_logl_wide_a", "-DGSN_SLICE_LO=0", "-DGSN_SLICE_HI=4",
                                      "-DGSN_PHASE_EXPORT"]),
    ("gsn_wide_b", "gsn_kernels.cu", ["-DGSN_SLICE_CFG=CfgWide", "-DGSN_SLICE_NAME=launch_logl_wide_b", "-DGSN_SLICE_LO=5", "-DGSN_SLICE_HI=10"]),
    ("gsn_narrow_a", "gsn_kernels.cu", ["-DGSN_SLICE_CFG=CfgNarrow", "-DGSN_SLICE_NAME=launch_logl_narrow_a", "-DGSN_SLICE_LO=0", "-DGSN_SLICE_HI=4"]),
    ("gsn_narrow_b", "gsn_kernels.cu", ["-DGSN_SLICE_CFG=CfgNarrow", "-DGSN_SLICE_NAME=launch_logl_narrow_b", "-DGSN_SLICE_LO=5", "-DGSN_SLICE_HI=10"]),
    ("gsn_tiny_a", "gsn_kernels.cu", ["-DGSN_SLICE_CFG=CfgTiny", "-DGSN_SLICE_NAME=launch_logl_tiny_a", "-DGSN_SLICE_LO=0", "-DGSN_SLICE_HI=4"]),
    ("gsn_tiny_b", "gsn_kernels.cu", ["-DGSN_SLICE_CFG=CfgTiny", "-DGSN_SLICE_NAME=launch_logl_tiny_b", "-DGSN_SLICE_LO=5", "-DGSN_SLICE_HI=10"]),
    ("gsn_cluster_a", "gsn_cluster.cu", ["-DGSN_SLICE_NAME=launch_logl_cluster_a", "-DGSN_SLICE_LO=0", "-DGSN_SLICE_HI=4"]),
    ("gsn_cluster_b", "gsn_cluster.cu", ["-DGSN_SLICE_NAME=launch_logl_cluster_b", "-DGSN_SLICE_LO=5", "-DGSN_SLICE_HI=10"]),
    ("gsn_onchip_a", "gsn_onchip.cu", ["-DGSN_SLICE_NAME=launch_logl_onchip_a", "-DGSN_SLICE_LO=0", "-DGSN_SLICE_HI=4"]),
    ("gsn_onchip_b", "gsn_onchip.cu", ["-DGSN_SLICE_NAME=launch_logl_onchip_b", "-DGSN_SLICE_LO=5", "-DGSN_SLICE_HI=10"]),
    ("gsn_predict_a", "gsn_predict.cu", ["-DGSN_SLICE_NAME=launch_predict_a", "-DGSN_SLICE_LO=0", "-DGSN_SLICE_HI=4"]),
    ("gsn_predict_b", "gsn_predict.cu", ["-DGSN_SLICE_NAME=launch_predict_b", "-DGSN_SLICE_LO=5", "-DGSN_SLICE_HI=10"]),
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libgsn.so cannot be built")


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(f) > t for f in SOURCES + HEADERS)


def build_lib(force: bool = False, verbose: bool = False) -> str:
    """Compile libgsn.so if it is missing or older than its sources."""
    if not force and not is_stale():
        return LIB_PATH
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    extra = os.environ.get("GSN_EXTRA_NVCC_FLAGS", "").split()   # experiments (e.g. -DGSN_ONCHIP_TRACE=20); never set in shipped builds
    procs = []
    for name, src, defs in UNITS:
        cmd = [nvcc, *NVCC_FLAGS, *extra, *defs, "-c", "-o", os.path.join(OBJ_DIR, name + ".o"), os.path.join(CSRC, src)]
        if verbose:
            cmd[1:1] = ["-Xptxas", "-v"]
            print(" ".join(cmd))
        procs.append((name, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for name, proc in procs:
        out, _ = proc.communicate()
        if proc.returncode != 0:
            raise RuntimeError(f"nvcc failed on {name} ({proc.returncode}):\n{out}")
        if verbose:
            print(out)
    objs = [os.path.join(OBJ_DIR, name + ".o") for name, _, _ in UNITS]
    res = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB_PATH, *objs], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed ({res.returncode}):\n{res.stdout}\n{res.stderr}")
    return LIB_PATH


if __name__ == "__main__":
    import sys
    print(build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv))
