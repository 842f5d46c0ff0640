"""ctypes binding of libgsn.so (include/gsn.h) -- the only way Python reaches the GPU.

There is no CPU implementation behind these calls: if the library is missing, or
no CUDA device is present, they raise.  PyTorch is used by callers for device
memory and streams only; nothing here imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

from . import build as _build

# --- enums (values fixed by include/gsn.h) ---------------------------------
K_EXPSQUARED, K_EXPONENTIAL, K_CONSTANT, K_DOTPRODUCT, K_MATERN32, K_MATERN52, K_RATQUAD, K_GIBBS, K_OU, K_PERIODIC, K_JJ = range(11)
M_UNIFORM, M_SIN, M_GAUSSIAN, M_EXPFUNCTION, M_BAZIN2009, M_KARPENKA2012, M_VILLAR2019, M_HOST = range(8)
L_NONE, L_CONSTANT, L_SIGMOID, L_SINUSOIDAL = range(4)
FLAG_NEG_INF = 1
Q_N, Q_NPARAMS, Q_NPARAMS_KERNEL, Q_NPARAMS_MEAN, Q_NPARAMS_LENS, Q_WORKSPACE_BYTES, Q_GRID, Q_DEVICE, Q_LAUNCHES, Q_NPAD, Q_PATH = range(11)
Q_CLUSTER2_MAX, Q_CLUSTER4_MAX, Q_CLUSTER8_MAX = 11, 12, 13
Q_ONCHIP, Q_ONCHIP_WARPS, Q_ONCHIP_SMEM, Q_ONCHIP_REG = 14, 15, 16, 17
PATH_WIDE, PATH_NARROW, PATH_CLUSTER, PATH_ONCHIP = 1, 2, 3, 4
ABI_VERSION = 1

EXPORTED_SYMBOLS = (
    "gsn_problem_create", "gsn_problem_destroy", "gsn_problem_set_theta_map", "gsn_logl_batch",
    "gsn_logl_batch_host", "gsn_logl_batch_hostmean", "gsn_lens_batch", "gsn_query", "gsn_covariance",
    "gsn_mean", "gsn_lens", "gsn_kernel_nparams", "gsn_mean_nparams", "gsn_lens_nparams",
    "gsn_abi_version", "gsn_last_error", "gsn_item_order", "gsn_logl_batch_gather",
    "gsn_predict_batch", "gsn_predict_batch_host", "gsn_predict_item_order", "gsn_problem_set_data",
    "gsn_measure_fp64_peak", "gsn_device_count", "gsn_onchip_slot_table",
)
MAX_ORDER = 8128   # largest padded order of a factorisation (32 * 254 block rows), also for n_rows + M of a prediction


class GsnError(RuntimeError):
    """A libgsn call returned a negative status."""


_lib = None
_lib_lock = threading.Lock()
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)


def lib_path() -> str:
    return os.environ.get("GSN_LIB", _build.LIB_PATH)   # GSN_LIB: tuning experiments with alternative builds


def load():
    """Load libgsn.so (never builds implicitly unless GSN_AUTOBUILD=1)."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        path = lib_path()
        if not os.path.exists(path):
            if os.environ.get("GSN_AUTOBUILD") == "1":
                _build.build_lib()
            else:
                raise GsnError(
                    f"{path} not found: build the CUDA library first "
                    "(python -c 'import __graft_entry__ as g; g.build()' or python -m gaussn_b200.build). "
                    "gaussn_b200 has no CPU fallback.")
        lib = C.CDLL(path)
        lib.gsn_abi_version.restype = C.c_int32
        if lib.gsn_abi_version() != ABI_VERSION:
            raise GsnError(f"libgsn ABI {lib.gsn_abi_version()} != binding ABI {ABI_VERSION}; rebuild")
        lib.gsn_last_error.restype = C.c_char_p
        lib.gsn_problem_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int64, _dp, _dp, _dp, C.c_int32, C.c_int32,
                                           _lp, C.c_int32, C.c_int32, C.c_int32, C.c_int32]
        lib.gsn_problem_destroy.argtypes = [C.c_void_p]
        lib.gsn_problem_set_theta_map.argtypes = [C.c_void_p, _ip, _dp]
        lib.gsn_problem_set_data.argtypes = [C.c_void_p, _dp, _dp, _dp]
        lib.gsn_logl_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        lib.gsn_logl_batch_gather.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.POINTER(C.c_void_p), C.c_int32, C.c_int64,
                                              C.c_void_p, C.c_uint32, C.c_void_p]
        lib.gsn_logl_batch_host.argtypes = [C.c_void_p, _dp, C.c_int64, C.c_int64, _dp, _ip, C.c_uint32]
        lib.gsn_logl_batch_hostmean.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                                C.c_void_p, C.c_uint32, C.c_void_p]
        lib.gsn_lens_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.gsn_predict_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int64,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.gsn_predict_batch_host.argtypes = [C.c_void_p, _dp, C.c_int64, C.c_int64, C.c_int64, C.c_int64, _dp, C.c_int64,
                                               _dp, _dp, _dp, _dp, _ip]
        lib.gsn_device_count.restype = C.c_int32
        lib.gsn_measure_fp64_peak.argtypes = [C.c_int, C.c_double, _dp]
        lib.gsn_onchip_slot_table.argtypes = [C.c_int32, C.c_int32, C.POINTER(C.c_uint8), C.c_int64]
        lib.gsn_onchip_slot_table.restype = C.c_int64
        lib.gsn_query.argtypes = [C.c_void_p, C.c_int32]
        lib.gsn_query.restype = C.c_int64
        lib.gsn_covariance.argtypes = [C.c_int, C.c_int32, _dp, C.c_int32, _dp, C.c_int64, _dp, C.c_int64, _dp]
        lib.gsn_mean.argtypes = [C.c_int, C.c_int32, _dp, C.c_int32, _dp, C.c_int64, _dp]
        lib.gsn_lens.argtypes = [C.c_int, C.c_int32, _dp, C.c_int32, _dp, C.c_int64, C.c_int32, C.c_int32, _lp, _dp, _dp]
        for name in ("gsn_kernel_nparams", "gsn_mean_nparams"):
            getattr(lib, name).argtypes = [C.c_int32]
            getattr(lib, name).restype = C.c_int32
        lib.gsn_lens_nparams.argtypes = [C.c_int32, C.c_int32]
        lib.gsn_lens_nparams.restype = C.c_int32
        lib.gsn_item_order.argtypes = [C.c_int64, _ip, C.c_int32, C.c_int32, C.POINTER(C.c_uint32), C.c_int64]
        lib.gsn_item_order.restype = C.c_int64
        lib.gsn_predict_item_order.argtypes = [C.c_int64, C.c_int64, C.c_int32, C.POINTER(C.c_uint32), C.c_int64]
        lib.gsn_predict_item_order.restype = C.c_int64
        _lib = lib
        return lib


def _check(rc: int):
    if rc != 0:
        raise GsnError(f"libgsn error {rc}: {load().gsn_last_error().decode()}")


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_dp)


def default_device() -> int:
    """Device used when none is given: GSN_DEVICE if set; else LOCAL_RANK (torchrun) folded onto the devices this process
    can see (launchers that expose one GPU per rank through CUDA_VISIBLE_DEVICES leave every rank with device 0); else 0."""
    if os.environ.get("GSN_DEVICE"):
        return int(os.environ["GSN_DEVICE"])
    rank = os.environ.get("LOCAL_RANK")
    if rank:
        n = int(load().gsn_device_count())
        return int(rank) % n if n > 0 else int(rank)
    return 0


def onchip_slot_table(tile_rows: int, lag: int):
    """(slots, table) of the on-chip shape's recycled tile storage: tile (i, k) lives in slot table[i (i + 1) / 2 + k]."""
    import numpy as np
    n = tile_rows * (tile_rows + 1) // 2
    tab = np.zeros(n, dtype=np.uint8)
    slots = load().gsn_onchip_slot_table(tile_rows, lag, tab.ctypes.data_as(C.POINTER(C.c_uint8)), n)
    if slots < 0:
        raise GsnError(f"gsn_onchip_slot_table({tile_rows}, {lag}): bad arguments")
    return int(slots), tab


def measure_fp64_peak(device=None, seconds=0.3) -> float:
    """FP64 tensor-pipe rate of the device in TFLOP/s (independent DMMA chains on every SM)."""
    out = C.c_double(0.0)
    _check(load().gsn_measure_fp64_peak(default_device() if device is None else device, float(seconds), C.byref(out)))
    return float(out.value)


# --- single-call plugin evaluations ----------------------------------------
def covariance(kind: int, params, x, x_prime=None, device=None) -> np.ndarray:
    lib = load()
    x = _f64(x).ravel()
    xp = None if x_prime is None else _f64(x_prime).ravel()
    p = _f64(params if params is not None else []).ravel()
    n, m = len(x), (len(x) if xp is None else len(xp))
    out = np.empty((n, m), dtype=np.float64)
    _check(lib.gsn_covariance(default_device() if device is None else device, kind, _ptr(p) if len(p) else None, len(p),
                              _ptr(x), n, None if xp is None else _ptr(xp), m, _ptr(out)))
    return out


def mean(kind: int, params, t, device=None) -> np.ndarray:
    lib = load()
    t = _f64(t).ravel()
    p = _f64(params).ravel()
    out = np.empty(len(t), dtype=np.float64)
    _check(lib.gsn_mean(default_device() if device is None else device, kind, _ptr(p), len(p), _ptr(t), len(t), _ptr(out)))
    return out


def lens(kind: int, params, x, n_bands: int, n_images: int, indices, device=None):
    lib = load()
    x = _f64(x).ravel()
    p = _f64(params if params is not None else []).ravel()
    idx = np.ascontiguousarray(np.asarray(indices, dtype=np.int64))
    xs = np.empty(len(x), dtype=np.float64)
    b = np.empty(len(x), dtype=np.float64)
    _check(lib.gsn_lens(default_device() if device is None else device, kind, _ptr(p) if len(p) else None, len(p), _ptr(x),
                        len(x), n_bands, n_images, idx.ctypes.data_as(_lp), _ptr(xs), _ptr(b)))
    return xs, b


def item_order(N: int, band_of_row=None, n_bands: int = 1, masked: bool = False):
    """Work-item order of one evaluation as [(c, j, first_col), ...] (no GPU needed)."""
    lib = load()
    band = None if band_of_row is None else np.ascontiguousarray(np.asarray(band_of_row, dtype=np.int32))
    bp = None if band is None else band.ctypes.data_as(_ip)
    n = lib.gsn_item_order(N, bp, n_bands, int(masked), None, 0)
    if n < 0:
        raise GsnError(lib.gsn_last_error().decode())
    out = np.empty(n, dtype=np.uint32)
    lib.gsn_item_order(N, bp, n_bands, int(masked), out.ctypes.data_as(C.POINTER(C.c_uint32)), n)
    return [(int((w >> 12) & 0xfff), int(w & 0xfff), int(w >> 24)) for w in out]


def predict_item_order(n_rows: int, M: int, want_cov: bool = True, cluster: bool = False):
    """Work-item order of one prediction as [(c, j, is_trailing), ...] (no GPU needed); ``cluster`` selects the
    column-major order used when one theta runs on a thread-block cluster."""
    lib = load()
    want_cov = int(bool(want_cov)) | (2 if cluster else 0)
    n = lib.gsn_predict_item_order(n_rows, M, int(want_cov), None, 0)
    if n < 0:
        raise GsnError(lib.gsn_last_error().decode())
    out = np.empty(n, dtype=np.uint32)
    lib.gsn_predict_item_order(n_rows, M, int(want_cov), out.ctypes.data_as(C.POINTER(C.c_uint32)), n)
    return [(int((w >> 12) & 0x7ff), int(w & 0xfff), bool(w & (1 << 23))) for w in out]


# --- problem handle ---------------------------------------------------------
class Problem:
    """Owns one ``gsn_problem``: a data set plus a plugin triple on one GPU."""

    def __init__(self, x, y, yerr, n_bands, n_images, indices, kernel_kind, mean_kind, lens_kind,
                 n_mean_params=-1, device=None):
        self._lib = load()
        self._h = C.c_void_p()
        self.device = default_device() if device is None else int(device)
        x, y, yerr = _f64(x).ravel(), _f64(y).ravel(), _f64(yerr).ravel()
        if not (len(x) == len(y) == len(yerr)):
            raise ValueError("x, y and yerr must have the same length")
        idx = np.ascontiguousarray(np.asarray(indices, dtype=np.int64))
        if len(idx) != n_bands * n_images + 1:
            raise ValueError("indices must have n_bands*n_images + 1 entries")
        _check(self._lib.gsn_problem_create(C.byref(self._h), self.device, len(x), _ptr(x), _ptr(y), _ptr(yerr), n_bands,
                                            n_images, idx.ctypes.data_as(_lp), kernel_kind, mean_kind, lens_kind, n_mean_params))
        self.N = len(x)
        self.n_params = int(self.query(Q_NPARAMS))
        self.n_kernel = int(self.query(Q_NPARAMS_KERNEL))
        self.n_mean = int(self.query(Q_NPARAMS_MEAN))
        self.n_lens = int(self.query(Q_NPARAMS_LENS))
        self.n_cols = self.n_params

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.gsn_problem_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def query(self, what: int) -> int:
        return int(self._lib.gsn_query(self._h, what))

    def set_data(self, x, y, yerr):
        """New observations of the same length and band / image structure, in place (no new handle)."""
        x, y, yerr = _f64(x).ravel(), _f64(y).ravel(), _f64(yerr).ravel()
        if not (len(x) == len(y) == len(yerr) == self.N):
            raise ValueError(f"set_data needs arrays of length {self.N}")
        _check(self._lib.gsn_problem_set_data(self._h, _ptr(x), _ptr(y), _ptr(yerr)))

    def set_theta_map(self, col_of_param=None, fixed_values=None):
        """``col_of_param[k]`` = column of the caller's theta feeding slot k, or -1
        (slot held at ``fixed_values[k]``).  ``None`` restores the identity map."""
        if col_of_param is None:
            _check(self._lib.gsn_problem_set_theta_map(self._h, None, None))
            self.n_cols = self.n_params
            return
        cols = np.ascontiguousarray(np.asarray(col_of_param, dtype=np.int32))
        fixed = _f64(np.zeros(self.n_params) if fixed_values is None else fixed_values)
        if len(cols) != self.n_params or len(fixed) != self.n_params:
            raise ValueError(f"theta map needs {self.n_params} entries")
        _check(self._lib.gsn_problem_set_theta_map(self._h, cols.ctypes.data_as(_ip), _ptr(fixed)))
        self.n_cols = int(cols.max()) + 1 if len(cols) else 0

    # host buffers in, host buffers out (H2D / D2H inside the call)
    def logl_host(self, theta, flags: int = 0, return_info: bool = False):
        theta = _f64(theta)
        if theta.ndim == 1:
            theta = theta.reshape(1, -1)
        B, ld = theta.shape
        out = np.empty(B, dtype=np.float64)
        info = np.empty(B, dtype=np.int32)
        _check(self._lib.gsn_logl_batch_host(self._h, _ptr(theta) if theta.size else None, B, ld, _ptr(out),
                                             info.ctypes.data_as(_ip), flags))
        return (out, info) if return_info else out

    # device pointers (torch tensors' data_ptr()), asynchronous on `stream`
    def logl_device(self, theta_ptr: int, B: int, ld: int, logl_ptr: int, info_ptr: int = 0, flags: int = 0, stream: int = 0):
        _check(self._lib.gsn_logl_batch(self._h, C.c_void_p(theta_ptr), B, ld, C.c_void_p(logl_ptr),
                                        C.c_void_p(info_ptr) if info_ptr else None, flags, C.c_void_p(stream) if stream else None))

    def logl_device_gather(self, theta_ptr: int, B: int, ld: int, out_ptrs, out_offset: int, info_ptr: int = 0, flags: int = 0,
                           stream: int = 0):
        """Results go to index out_offset + i of every buffer in `out_ptrs` (own and NVLink-peer device pointers)."""
        arr = (C.c_void_p * len(out_ptrs))(*[C.c_void_p(int(p)) for p in out_ptrs])
        _check(self._lib.gsn_logl_batch_gather(self._h, C.c_void_p(theta_ptr), B, ld, arr, len(out_ptrs), out_offset,
                                               C.c_void_p(info_ptr) if info_ptr else None, flags, C.c_void_p(stream) if stream else None))

    def logl_device_hostmean(self, theta_ptr: int, B: int, ld: int, mean_ptr: int, logl_ptr: int, info_ptr: int = 0,
                             flags: int = 0, stream: int = 0):
        _check(self._lib.gsn_logl_batch_hostmean(self._h, C.c_void_p(theta_ptr), B, ld, C.c_void_p(mean_ptr), C.c_void_p(logl_ptr),
                                                 C.c_void_p(info_ptr) if info_ptr else None, flags,
                                                 C.c_void_p(stream) if stream else None))

    def lens_device(self, theta_ptr: int, B: int, ld: int, shifted_ptr: int, b_ptr: int, stream: int = 0):
        _check(self._lib.gsn_lens_batch(self._h, C.c_void_p(theta_ptr), B, ld, C.c_void_p(shifted_ptr), C.c_void_p(b_ptr),
                                        C.c_void_p(stream) if stream else None))

    # GP.predict for a batch of theta: rows [row0, row0 + n_rows) are conditioned on, x_prime are the new epochs
    def predict_host(self, theta, row0: int, n_rows: int, x_prime, mean_v=None, mean_u=None, want_cov: bool = True):
        theta = _f64(theta)
        if theta.ndim == 1:
            theta = theta.reshape(1, -1)
        B, ld = theta.shape
        xp = _f64(x_prime).ravel()
        M = len(xp)
        mv = None if mean_v is None else _f64(mean_v).reshape(B, n_rows)
        mu = None if mean_u is None else _f64(mean_u).reshape(B, M)
        e = np.empty((B, M), dtype=np.float64)
        v = np.empty((B, M, M), dtype=np.float64) if want_cov else None
        info = np.empty(B, dtype=np.int32)
        _check(self._lib.gsn_predict_batch_host(self._h, _ptr(theta) if theta.size else None, B, ld, row0, n_rows, _ptr(xp), M,
                                                None if mv is None else _ptr(mv), None if mu is None else _ptr(mu), _ptr(e),
                                                None if v is None else _ptr(v), info.ctypes.data_as(_ip)))
        return e, v, info

    def predict_device(self, theta_ptr: int, B: int, ld: int, row0: int, n_rows: int, xprime_ptr: int, M: int, exp_ptr: int,
                       var_ptr: int = 0, info_ptr: int = 0, mean_v_ptr: int = 0, mean_u_ptr: int = 0, stream: int = 0):
        vp = lambda q: C.c_void_p(q) if q else None
        _check(self._lib.gsn_predict_batch(self._h, vp(theta_ptr), B, ld, row0, n_rows, vp(xprime_ptr), M, vp(mean_v_ptr),
                                           vp(mean_u_ptr), vp(exp_ptr), vp(var_ptr), vp(info_ptr), vp(stream)))
