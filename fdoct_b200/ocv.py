"""The reference's raw matrix file format `.ocv` (matwrite / matread, BscanFFTspinj.cpp:672-716): four native ints
`rows, cols, type, channels` followed by the row-major element data.  `type` is the OpenCV type code
(depth + ((channels - 1) << 3)); `spectrum.ocv` holds data_yb as CV_64F (BscanFFTspinj.cpp:1788-1789) and
`bscan%03d.ocv` holds bscandb as CV_64F (BscanFFTspinj.cpp:2042)."""
from __future__ import annotations

import numpy as np

_DEPTHS = {0: np.uint8, 1: np.int8, 2: np.uint16, 3: np.int16, 4: np.int32, 5: np.float32, 6: np.float64}
_CODES = {np.dtype(v): k for k, v in _DEPTHS.items()}


def write_ocv(path: str, mat: np.ndarray) -> None:
    mat = np.ascontiguousarray(mat)
    if mat.ndim == 2:
        mat = mat[:, :, None]
    if mat.ndim != 3 or mat.dtype not in _CODES or not 1 <= mat.shape[2] <= 4:
        raise ValueError(f"cannot store {mat.dtype}{mat.shape} as .ocv")
    rows, cols, ch = mat.shape
    typ = _CODES[mat.dtype] + ((ch - 1) << 3)
    with open(path, "wb") as f:
        np.array([rows, cols, typ, ch], dtype=np.int32).tofile(f)
        mat.tofile(f)


def read_ocv(path: str) -> np.ndarray:
    with open(path, "rb") as f:
        hdr = np.fromfile(f, dtype=np.int32, count=4)
        if hdr.size != 4:
            raise ValueError(f"{path}: truncated .ocv header")
        rows, cols, typ, ch = (int(x) for x in hdr)
        depth, ch_t = typ & 7, (typ >> 3) + 1
        if depth not in _DEPTHS or rows < 0 or cols < 0 or ch_t != ch:
            raise ValueError(f"{path}: bad .ocv header {hdr.tolist()}")
        data = np.fromfile(f, dtype=_DEPTHS[depth], count=rows * cols * ch)
    if data.size != rows * cols * ch:
        raise ValueError(f"{path}: truncated .ocv data")
    return data.reshape(rows, cols) if ch == 1 else data.reshape(rows, cols, ch)
