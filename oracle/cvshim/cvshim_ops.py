"""NumPy / cv2 back end of oracle/cvshim/opencv2/opencv.hpp.  TEST INFRASTRUCTURE ONLY (see oracle/README in DESIGN.md section 2).

Every function here is the Python spelling of ONE OpenCV C++ call the reference's processing block makes; the arithmetic is done
by OpenCV itself (cv2), or - for Mat::convertTo and the MatExpr scalings, which cv2 does not export - by the one IEEE operation
OpenCV performs per element (saturate_cast<T>(src * alpha + beta))."""
import cv2
import numpy as np

_DEPTH = {0: np.uint8, 1: np.int8, 2: np.uint16, 3: np.int16, 4: np.int32, 5: np.float32, 6: np.float64}


def aligned_empty(shape, dtype, align=64):
    """cv::Mat buffers are 64-byte aligned (cv::fastMalloc); cv2's IPP-backed kernels differ by one f32 ulp between alignments."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    raw = np.empty(n + align, dtype=np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off:off + n].view(dtype).reshape(shape)


def aligned_copy(a):
    out = aligned_empty(a.shape, a.dtype)
    out[...] = a
    return out


def zeros(rows, cols, depth, cn):
    out = aligned_empty((rows, cols) if cn == 1 else (rows, cols, cn), _DEPTH[depth])
    out[...] = 0
    return out


def depth_of(a):
    for k, v in _DEPTH.items():
        if a.dtype == v:
            return k
    raise TypeError(a.dtype)


def convert(a, depth, alpha, beta):
    """Mat::convertTo: dst = saturate_cast<T>(src * alpha + beta), evaluated in double (float for 8/16-bit sources to a float
    target makes no difference at alpha = 1, beta = 0, the only integer-source case the block has)."""
    t = _DEPTH[depth]
    if alpha == 1.0 and beta == 0.0:
        if np.issubdtype(t, np.integer) and not np.issubdtype(a.dtype, np.integer):
            info = np.iinfo(t)
            return np.clip(np.rint(a), info.min, info.max).astype(t)  # cvRound = round half to even, then saturate
        return a.astype(t)
    x = a.astype(np.float64) * alpha
    if beta != 0.0:
        x = x + beta
    if np.issubdtype(t, np.integer):
        info = np.iinfo(t)
        return np.clip(np.rint(x), info.min, info.max).astype(t)
    return x.astype(t)


def copy(a):
    return aligned_copy(np.asarray(a))


def assign(dst, src):
    np.copyto(dst, src)


def view(a, y, x, h, w):
    return a[y:y + h, x:x + w]


def subtract(a, b):
    return cv2.subtract(a, b)


def add(a, b):
    return cv2.add(a, b)


def add_scalar(a, s):
    return cv2.add(a, float(s))


def divide(a, b):
    return cv2.divide(a, b)


def divide_scalar_by(s, a):
    return cv2.divide(float(s), a)


def max_scalar(a, s):
    return cv2.max(a, float(s))


def normalize(a, lo, hi, norm_type):
    return cv2.normalize(a, None, lo, hi, norm_type)


def mean0(a):
    return float(cv2.mean(a)[0])


def multiply(a, b):
    return cv2.multiply(a, b)


def merge(planes):
    return aligned_copy(cv2.merge(list(planes)))


def split(a):
    return [aligned_copy(p) for p in cv2.split(a)]


def _dft_out_shape(a, flags):
    if flags & cv2.DFT_REAL_OUTPUT:
        return a.shape[:2]
    if a.ndim == 2 and (flags & cv2.DFT_COMPLEX_OUTPUT):
        return a.shape + (2,)
    return a.shape


def dft(a, flags):
    return cv2.dft(aligned_copy(a), aligned_empty(_dft_out_shape(a, flags), a.dtype), flags=flags)


def magnitude(x, y):
    return cv2.magnitude(aligned_copy(x), aligned_copy(y), aligned_empty(x.shape, x.dtype))


def accumulate(src, dst):
    cv2.accumulate(src, dst)


def transpose(a):
    return cv2.transpose(a)


def log(a):
    return cv2.log(a)


def copy_make_border(a, t, b, l, r, btype, value):
    return cv2.copyMakeBorder(a, t, b, l, r, btype, value=float(value))


def median_blur(a, k):
    return cv2.medianBlur(a, k)


def resize(a, fx, fy, interp):
    return cv2.resize(a, None, fx=fx, fy=fy, interpolation=interp)


def set_all(a, v):
    a[...] = v


def threshold(a, thresh, maxval, ttype):
    return cv2.threshold(a, thresh, maxval, ttype)[1]


def apply_color_map(a, colormap):
    return cv2.applyColorMap(np.ascontiguousarray(a), colormap)


def opencv_version():
    return cv2.__version__
