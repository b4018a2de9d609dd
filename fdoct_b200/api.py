"""ctypes binding of the C ABI in ``include/abcoct.h`` (the product path; no oracle, no CPU fallback).

The reference is compiled C++ with no Python layer, so this module is only the thin host shim the tests and
``bench.py`` drive the library through: same names and argument meaning as the C entry points, errors raised
as :class:`AbcoctError` carrying the C status code and ``abcoct_last_error`` text.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_STATE, ERR_IO = 0, -1, -2, -3, -4, -5
INI_BSCANFFT, INI_SPINJ, INI_SPINJNT, INI_DARK, INI_PEAK, INI_WEBCAM, INI_SIM = range(7)


class AbcoctError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"abcoct error {code}: {msg}")
        self.code = code


class Params(C.Structure):
    """``abcoct_params`` (include/abcoct.h); field-for-field."""

    _fields_ = [
        ("w", C.c_uint32), ("h", C.c_uint32), ("bpp", C.c_uint32), ("binx", C.c_uint32), ("biny", C.c_uint32),
        ("averages", C.c_uint32), ("numfftpoints", C.c_uint32), ("numdisplaypoints", C.c_uint32),
        ("lambdamin", C.c_double), ("lambdamax", C.c_double), ("mediann", C.c_int32), ("movavgn", C.c_int32),
        ("fft_multiplier", C.c_uint32), ("rowwisenormalize", C.c_uint8), ("donotnormalize", C.c_uint8),
        ("variant", C.c_uint8), ("weight_mode", C.c_uint8), ("bscanthreshold", C.c_double),
        ("clampupper", C.c_uint8), ("bandpassfilter", C.c_uint8), ("lowpassfilter", C.c_uint8), ("output_rebin", C.c_uint8),
        ("bscanbinx", C.c_uint8), ("bscanbiny", C.c_uint8), ("channelnum", C.c_uint8), ("reserved", C.c_uint8 * 1),
        ("clamp_db", C.c_double),
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_ if n != "reserved"}


class Info(C.Structure):
    _fields_ = [
        ("opw", C.c_uint32), ("oph", C.c_uint32), ("M", C.c_uint32), ("N", C.c_uint32), ("D", C.c_uint32),
        ("averages", C.c_uint32), ("fft_threads", C.c_uint32), ("fft_radix", C.c_uint32 * 3),
        ("groups_per_cta", C.c_uint32), ("ctas_per_sm", C.c_uint32), ("smem_bytes", C.c_uint32),
        ("regs_per_thread", C.c_uint32), ("ngpu", C.c_uint32), ("sm_count", C.c_uint32),
        ("kernel_launches", C.c_uint64), ("last_recon_ms", C.c_double), ("last_norm_ms", C.c_double),
        ("kernel_kind", C.c_uint32), ("slots_per_warp", C.c_uint32),
    ]


class Outputs(C.Structure):
    """abcoct_outputs (include/abcoct.h): nullable image pointers of the *_ex entry points."""
    _fields_ = [("bscan_u8", C.c_void_p), ("bscan_db", C.c_void_p), ("bscan_lin", C.c_void_p), ("bscan_bgr", C.c_void_p),
                ("jsub_u8", C.c_void_p), ("jsub_bgr", C.c_void_p), ("reserved", C.c_void_p * 2)]


# name -> (dtype, trailing shape) of every image an *_ex call can produce
OUTPUT_KINDS = {"bscan_u8": (np.uint8, ()), "bscan_db": (np.float32, ()), "bscan_lin": (np.float32, ()),
                "bscan_bgr": (np.uint8, (3,)), "jsub_u8": (np.uint8, ()), "jsub_bgr": (np.uint8, (3,))}

EXPORTS = [
    "abcoct_params_default", "abcoct_params_from_ini", "abcoct_create", "abcoct_destroy", "abcoct_last_error",
    "abcoct_set_background", "abcoct_set_pishift", "abcoct_set_dark", "abcoct_set_calibration_from_frames",
    "abcoct_compose_dark_background", "abcoct_get_calibration", "abcoct_set_threshold", "abcoct_set_clampupper", "abcoct_set_averages",
    "abcoct_build_tables", "abcoct_get_tables", "abcoct_get_window", "abcoct_process_bscans",
    "abcoct_process_bscans_device", "abcoct_timing_reset", "abcoct_timing_read", "abcoct_debug_linearised", "abcoct_host_alloc", "abcoct_host_free", "abcoct_get_info",
    "abcoct_set_jscan", "abcoct_process_bscans_ex", "abcoct_process_bscans_device_ex",
]

_lib = None


def lib() -> C.CDLL:
    """Load (building first if stale) the in-tree ``libabcoct.so``. Raises if it cannot be built or loaded."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if os.environ.get("ABCOCT_NO_BUILD") != "1":
        try:
            path = _build.build()
        except Exception:
            if not os.path.exists(path):
                raise
    L = C.CDLL(path)
    vp, sz, i32 = C.c_void_p, C.c_size_t, C.c_int
    L.abcoct_params_default.argtypes = [C.POINTER(Params)]
    L.abcoct_params_default.restype = None
    L.abcoct_params_from_ini.argtypes = [C.c_char_p, i32, C.POINTER(Params)]
    L.abcoct_create.argtypes = [C.POINTER(Params), C.POINTER(i32), i32, C.POINTER(vp)]
    L.abcoct_destroy.argtypes = [vp]
    L.abcoct_destroy.restype = None
    L.abcoct_last_error.argtypes = [vp]
    L.abcoct_last_error.restype = C.c_char_p
    for n in ("abcoct_set_background", "abcoct_set_pishift", "abcoct_set_dark"):
        getattr(L, n).argtypes = [vp, vp, sz]
    L.abcoct_set_calibration_from_frames.argtypes = [vp, i32, vp, sz, sz]
    L.abcoct_set_threshold.argtypes = [vp, C.c_double]
    L.abcoct_set_clampupper.argtypes = [vp, i32]
    L.abcoct_set_averages.argtypes = [vp, C.c_uint32]
    L.abcoct_compose_dark_background.argtypes = [vp]
    L.abcoct_get_calibration.argtypes = [vp, i32, vp, sz]
    L.abcoct_build_tables.argtypes = [C.POINTER(Params), vp, vp, vp]
    L.abcoct_get_tables.argtypes = [vp, vp, vp]
    L.abcoct_get_window.argtypes = [vp, vp]
    L.abcoct_process_bscans.argtypes = [vp, vp, sz, sz, vp, vp]
    L.abcoct_process_bscans_device.argtypes = [vp, i32, vp, sz, sz, vp, vp, vp]
    L.abcoct_set_jscan.argtypes = [vp, vp, sz]
    L.abcoct_process_bscans_ex.argtypes = [vp, vp, sz, sz, C.POINTER(Outputs)]
    L.abcoct_process_bscans_device_ex.argtypes = [vp, i32, vp, sz, sz, C.POINTER(Outputs), vp]
    L.abcoct_timing_reset.argtypes = [vp]
    L.abcoct_timing_read.argtypes = [vp, i32, C.POINTER(C.c_uint32), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.abcoct_debug_linearised.argtypes = [vp, vp, sz, vp]
    L.abcoct_host_alloc.argtypes = [sz]
    L.abcoct_host_alloc.restype = vp
    L.abcoct_host_free.argtypes = [vp]
    L.abcoct_host_free.restype = None
    L.abcoct_get_info.argtypes = [vp, C.POINTER(Info)]
    _lib = L
    return L


def default_params(**kw) -> Params:
    p = Params()
    lib().abcoct_params_default(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def params_from_ini(path: str, flavour: int = INI_BSCANFFT) -> Params:
    p = Params()
    rc = lib().abcoct_params_from_ini(os.fsencode(path), flavour, C.byref(p))
    if rc != OK:
        raise AbcoctError(rc, f"cannot read ini file {path}")
    return p


def build_tables(p: Params):
    """Host-only table precompute: (nearestkindex int32[N], fractionalk f64[N], barthannwin f64[opw])."""
    nk = np.empty(p.numfftpoints, dtype=np.int32)
    fr = np.empty(p.numfftpoints, dtype=np.float64)
    win = np.empty(p.w // max(p.binx, 1), dtype=np.float64)
    rc = lib().abcoct_build_tables(C.byref(p), nk.ctypes.data, fr.ctypes.data, win.ctypes.data)
    if rc != OK:
        raise AbcoctError(rc, "abcoct_build_tables rejected the parameters")
    return nk, fr, win


class PinnedArray:
    """A numpy view over ``abcoct_host_alloc`` memory (the caller-side pinned ring)."""

    def __init__(self, shape, dtype):
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(shape)) * self.dtype.itemsize
        self.ptr = lib().abcoct_host_alloc(max(self.nbytes, 16))
        if not self.ptr:
            raise AbcoctError(ERR_CUDA, "abcoct_host_alloc failed")
        buf = (C.c_uint8 * max(self.nbytes, 16)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            lib().abcoct_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """``abcoct_ctx`` owner. Mirrors the C entry points one-to-one."""

    def __init__(self, params: Params, gpu_ids=None, ngpu: int = 1):
        self._h = C.c_void_p()
        self.params = params
        ids = None
        if gpu_ids is not None:
            ngpu = len(gpu_ids)
            ids = (C.c_int * ngpu)(*gpu_ids)
        rc = lib().abcoct_create(C.byref(params), ids, ngpu, C.byref(self._h))
        if rc != OK:
            raise AbcoctError(rc, (lib().abcoct_last_error(None) or b"").decode())
        self.opw = params.w // params.binx
        self.oph = params.h // params.biny
        self.D = params.numdisplaypoints
        self.A = params.averages

    def _check(self, rc):
        if rc != OK:
            raise AbcoctError(rc, (lib().abcoct_last_error(self._h) or b"").decode())

    def close(self):
        if self._h:
            lib().abcoct_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # calibration -----------------------------------------------------------------------
    def _cal(self, fn, a):
        if a is None:
            self._check(fn(self._h, None, 0))
            return
        a = np.ascontiguousarray(a, dtype=np.float64)
        assert a.shape == (self.oph, self.opw), a.shape
        self._check(fn(self._h, a.ctypes.data, self.opw))

    def set_background(self, yb):
        self._cal(lib().abcoct_set_background, yb)

    def set_pishift(self, yp):
        self._cal(lib().abcoct_set_pishift, yp)

    def set_dark(self, yd):
        self._cal(lib().abcoct_set_dark, yd)

    def set_calibration_from_frames(self, which: int, frames: np.ndarray):
        frames = np.ascontiguousarray(frames)
        self._check_frames(frames)
        self._check(lib().abcoct_set_calibration_from_frames(self._h, which, frames.ctypes.data, frames.shape[0], 0))

    def set_threshold(self, thr: float):
        self._check(lib().abcoct_set_threshold(self._h, thr))

    def set_clampupper(self, on: bool):
        self._check(lib().abcoct_set_clampupper(self._h, int(on)))

    def set_averages(self, averages: int):
        self._check(lib().abcoct_set_averages(self._h, averages))
        self.A = averages

    def get_calibration(self, which: int) -> np.ndarray:
        out = np.empty((self.oph, self.opw), dtype=np.float64)
        self._check(lib().abcoct_get_calibration(self._h, which, out.ctypes.data, 0))
        return out

    def compose_dark_background(self):
        self._check(lib().abcoct_compose_dark_background(self._h))

    def tables(self):
        nk = np.empty(self.params.numfftpoints, dtype=np.int32)
        fr = np.empty(self.params.numfftpoints, dtype=np.float64)
        self._check(lib().abcoct_get_tables(self._h, nk.ctypes.data, fr.ctypes.data))
        win = np.empty(self.opw, dtype=np.float64)
        self._check(lib().abcoct_get_window(self._h, win.ctypes.data))
        return nk, fr, win

    # the hot path ----------------------------------------------------------------------
    def _check_frames(self, frames: np.ndarray):
        """(nframes, h, w) uint8 / uint16, or (nframes, h, w, 3) uint8 interleaved BGR when channelnum >= 3 (BscanFFTwebcam.cpp:1021)."""
        assert frames.flags.c_contiguous
        assert frames.dtype == (np.uint8 if self.params.bpp == 8 else np.uint16), frames.dtype
        if self.params.channelnum >= 3:
            assert frames.ndim == 4 and frames.shape[1:] == (self.params.h, self.params.w, 3), frames.shape
        else:
            assert frames.ndim == 3 and frames.shape[1:] == (self.params.h, self.params.w), frames.shape

    def process_bscans(self, frames: np.ndarray, want_db: bool = False, out8: np.ndarray | None = None,
                       outdb: np.ndarray | None = None):
        """Host-buffer call (abcoct_process_bscans). frames: (nframes, h, w) uint16, C-contiguous."""
        self._check_frames(frames)
        nframes = frames.shape[0]
        nB = nframes // max(self.A, 1)
        if out8 is None:
            out8 = np.empty((nB, self.D, self.oph), dtype=np.uint8)
        if want_db and outdb is None:
            outdb = np.empty((nB, self.D, self.oph), dtype=np.float32)
        self._check(lib().abcoct_process_bscans(self._h, frames.ctypes.data, nframes, 0, out8.ctypes.data,
                                                outdb.ctypes.data if outdb is not None else None))
        return (out8, outdb) if want_db else out8

    def process_bscans_device(self, d_frames: int, nframes: int, d_out8: int, d_outdb: int | None = None,
                              stream: int | None = None, gpu_index: int = 0, stride_bytes: int = 0):
        """Device-pointer call (abcoct_process_bscans_device); pointers are raw integers (e.g. tensor.data_ptr())."""
        self._check(lib().abcoct_process_bscans_device(self._h, gpu_index, d_frames, nframes, stride_bytes, d_out8,
                                                       d_outdb, stream))

    def set_jscan(self, jscan: np.ndarray | None):
        """Key 'j' / 'c' (BscanFFT.cpp:1292-1303): keep a linear B-scan (D x oph float32) as the lock-in reference, None = off."""
        if jscan is None:
            self._check(lib().abcoct_set_jscan(self._h, None, 0))
            return
        j = np.ascontiguousarray(jscan, dtype=np.float32)
        assert j.shape == (self.D, self.oph), j.shape
        self._check(lib().abcoct_set_jscan(self._h, j.ctypes.data, 0))

    def process_bscans_ex(self, frames: np.ndarray, want=("bscan_u8",)) -> dict:
        """abcoct_process_bscans_ex: returns {name: array} for the requested images (see OUTPUT_KINDS); bscan_u8 always."""
        self._check_frames(frames)
        nframes = frames.shape[0]
        nB = nframes // max(self.A, 1)
        res, o = {}, Outputs()
        for name in set(want) | {"bscan_u8"}:
            dt, tail = OUTPUT_KINDS[name]
            res[name] = np.empty((nB, self.D, self.oph) + tail, dtype=dt)
            setattr(o, name, res[name].ctypes.data)
        self._check(lib().abcoct_process_bscans_ex(self._h, frames.ctypes.data, nframes, 0, C.byref(o)))
        return res

    def process_bscans_device_ex(self, d_frames: int, nframes: int, d_out: dict, stream: int | None = None, gpu_index: int = 0,
                                 stride_bytes: int = 0):
        """abcoct_process_bscans_device_ex; d_out maps output names to raw device pointers."""
        o = Outputs()
        for name, ptr in d_out.items():
            assert name in OUTPUT_KINDS, name
            setattr(o, name, ptr)
        self._check(lib().abcoct_process_bscans_device_ex(self._h, gpu_index, d_frames, nframes, stride_bytes, C.byref(o), stream))

    def debug_linearised(self, frame: np.ndarray) -> np.ndarray:
        """data_ylin of one frame (abcoct_debug_linearised): float32 [oph, numfftpoints]."""
        frame = np.ascontiguousarray(frame)
        assert frame.shape == (self.params.h, self.params.w)
        out = np.empty((self.oph, self.params.numfftpoints), dtype=np.float32)
        self._check(lib().abcoct_debug_linearised(self._h, frame.ctypes.data, 0, out.ctypes.data))
        return out

    def timing_reset(self):
        self._check(lib().abcoct_timing_reset(self._h))

    def timing_read(self, gpu_index: int = 0):
        """(chunks timed, summed recon-kernel ms, summed normalise-kernel ms) since timing_reset."""
        n, r, m = C.c_uint32(), C.c_double(), C.c_double()
        self._check(lib().abcoct_timing_read(self._h, gpu_index, C.byref(n), C.byref(r), C.byref(m)))
        return n.value, r.value, m.value

    def info(self) -> Info:
        i = Info()
        self._check(lib().abcoct_get_info(self._h, C.byref(i)))
        return i
