"""CPU: the C-ABI library builds for sm_100a without a GPU, loads, exports every symbol that
include/gsn.h declares, and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, has_cuda


def header_functions():
    text = open(os.path.join(ROOT, "include", "gsn.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gsn_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    from gaussn_b200 import _engine
    assert sorted(_engine.EXPORTED_SYMBOLS) == header_functions()


def test_library_exports_every_declared_symbol(libgsn):
    for name in header_functions():
        assert hasattr(libgsn, name), f"libgsn.so does not export {name}"
    assert libgsn.gsn_abi_version() == 1


def test_built_for_sm100a(libgsn):
    import subprocess
    from gaussn_b200 import build
    out = subprocess.run(["cuobjdump", "-lelf", build.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_nparams_tables(libgsn):
    from oracle import gp_oracle as o
    from gaussn_b200 import kernels, meanfuncs, lensingmodels
    for name, n in o.KERNEL_NPARAMS.items():
        assert libgsn.gsn_kernel_nparams(getattr(kernels, name).kind) == n, name
    for name, n in o.MEAN_NPARAMS.items():
        assert libgsn.gsn_mean_nparams(getattr(meanfuncs, name).kind) == n, name
    for name, n in o.LENS_NPARAMS_PER_IMAGE.items():
        assert libgsn.gsn_lens_nparams(getattr(lensingmodels, name).kind, 3) == 2 * n, name
    assert libgsn.gsn_kernel_nparams(99) == -1 and libgsn.gsn_lens_nparams(1, 0) == -1


def test_invalid_arguments_are_rejected_before_any_cuda_call(libgsn):
    from gaussn_b200 import _engine
    h = ctypes.c_void_p()
    x = np.zeros(4)
    idx = np.array([0, 4], dtype=np.int64)
    dp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    ip = idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))
    rc = libgsn.gsn_problem_create(ctypes.byref(h), 0, 4, dp(x), dp(x), dp(x), 1, 1, ip, 99, 0, 0, -1)
    assert rc == -1 and b"kernel kind" in libgsn.gsn_last_error()
    bad = np.array([0, 3], dtype=np.int64)
    rc = libgsn.gsn_problem_create(ctypes.byref(h), 0, 4, dp(x), dp(x), dp(x), 1, 1,
                                   bad.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), 0, 0, 0, -1)
    assert rc == -1 and b"indices" in libgsn.gsn_last_error()
    rc = libgsn.gsn_problem_create(ctypes.byref(h), 0, 0, dp(x), dp(x), dp(x), 1, 1, ip, 0, 0, 0, -1)
    assert rc == -1
    assert libgsn.gsn_query(None, _engine.Q_N) == -1


@pytest.mark.skipif(has_cuda(), reason="checks behaviour on a machine without a GPU")
def test_no_cpu_fallback_without_a_gpu(libgsn):
    from gaussn_b200 import _engine, gausSN, kernels, meanfuncs
    with pytest.raises(_engine.GsnError, match="no CUDA device"):
        kernels.ExpSquaredKernel([1.0, 2.0]).covariance(np.linspace(0, 1, 5))
    with pytest.raises(_engine.GsnError, match="no CUDA device"):
        meanfuncs.Sin([1.0, 2.0, 0.0]).mean(np.linspace(0, 1, 5))
    gp = gausSN.GP(kernels.ExpSquaredKernel([1.0, 2.0]), meanfuncs.UniformMean([0.0]))
    x = np.linspace(0, 1, 8)
    with pytest.raises(_engine.GsnError, match="no CUDA device"):
        gp.loglikelihood(x, x, 0.1 + 0 * x, [1.0, 2.0], [0.0], None)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under gaussn_b200/ may import or execute it."""
    pkg = os.path.join(ROOT, "gaussn_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "gp_oracle" not in text, f
