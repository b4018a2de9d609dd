"""CPU oracle for the ABC-OCT per-frame B-scan reconstruction block.

TEST INFRASTRUCTURE ONLY.  Nothing under ``fdoct_b200/`` imports this module;
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may.  The product path is the CUDA
library behind ``include/abcoct.h`` and it fails loudly when that library is
missing.

What this is
------------
A statement-by-statement restatement of the reference's inline processing block
(all citations are ``file:line`` into ``/root/reference``):

* lambda->k tables ................ BscanFFT.cpp:615-698
* Bartlett-Hann window ............ BscanFFT.cpp:936-944
* median / binning / to-double .... BscanFFT.cpp:953-958, 987-991
* smoothmovavg .................... BscanFFT.cpp:247-304
* normalizerows ................... BscanFFT.cpp:88-97
* dark subtract (DARK variant) .... BscanDark.cpp:1269 ; yb synthesis BscanDark.cpp:996
* (y - yp) / yb ................... BscanFFT.cpp:1125-1132
* DC removal + apodisation ........ BscanFFT.cpp:1135-1143
* zeropadrowwise (Fourier upsample) BscanFFT.cpp:180-245 (band-pass: BscanDark.cpp:218-236)
* slopes + gather-lerp ............ BscanFFT.cpp:1151-1177
* row inverse DFT, magnitude ...... BscanFFT.cpp:1181-1190
* crop + accumulate ............... BscanFFT.cpp:1193-1209
* finalise (dB, mask, clamp, u8) .. BscanFFT.cpp:1220-1255

The reference's third-party arithmetic (OpenCV ``dft``, ``resize``,
``medianBlur``, ``normalize``, ``mean``, ``log``, ``magnitude``) is NOT
re-implemented: the same OpenCV kernels are called through
``opencv-python-headless`` (4.13.0 in this image; the reference pins no OpenCV
version - README.md:18 just installs ``libopencv-dev``).

PINNED (round 2): the reference ships no tests and no golden outputs and its executables cannot be built here (no
OpenCV C++ headers, no camera SDKs) - but its processing block can.  oracle/build_ref.py cuts the block, the helper
functions, the table precompute, the window loop and the frame ingest out of /root/reference/BscanFFT.cpp and
BscanDark.cpp and compiles them verbatim against oracle/cvshim (a stand-in for the OpenCV headers that forwards every
OpenCV call to the same kernels through cv2) into oracle/_ref/.  ``strict=True`` of this module equals those modules
BIT FOR BIT (tables, display bytes, f64 dB image, calibration captures of both key handlers, the DARK composition, the
J0 lock-in display, the JET images, the webcam channel sum, BscanFFTspinjnt's re-binning) on every configuration of
tests/test_oracle_pinned.py, and reproduces the vectors they wrote (tests/golden/ref_*.npz).  Also kept: the physics known-answer test on the reference's own
fixtures ``Matlab files/imgi.png`` / ``backg.png`` (tests/test_oracle.py).

Two execution modes give the same f64 intermediates to <= 1e-12 relative (single f32 ulps can flip behind the f32 DFT):
``strict=True`` issues the per-row OpenCV calls exactly like the reference's C++
loops; ``strict=False`` vectorises those loops with NumPy so that the CPU
baseline is not handicapped by the Python interpreter.
"""
from __future__ import annotations

import dataclasses
import math

import cv2
import numpy as np

PI = 3.141592653589793  # BscanFFT.cpp:609


def aligned_empty(shape, dtype, align: int = 64) -> np.ndarray:
    """An array whose data pointer is `align`-byte aligned, like every cv::Mat buffer (cv::fastMalloc, CV_MALLOC_ALIGN = 64).
    cv2's IPP-backed kernels (magnitude, dft) pick their peel / vector split from the pointer alignment and differ by one f32 ulp
    between alignments (measured: cv2.magnitude into a destination at 4 mod 16 bytes), so the f32 stages run on buffers aligned
    the way the C++ reference's are."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    raw = np.empty(n + align, dtype=np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off : off + n].view(dtype).reshape(shape)


def aligned_copy(a: np.ndarray) -> np.ndarray:
    out = aligned_empty(a.shape, a.dtype)
    out[...] = a
    return out


@dataclasses.dataclass
class Params:
    """Mirror of ``abcoct_params`` (include/abcoct.h) / the .ini fields the block reads."""

    w: int
    h: int
    bpp: int = 16
    binx: int = 1
    biny: int = 1
    averages: int = 1
    numfftpoints: int = 1024
    numdisplaypoints: int = 512
    lambdamin: float = 816e-9  # BscanFFT.cpp:381
    lambdamax: float = 884e-9  # BscanFFT.cpp:382
    mediann: int = 0
    movavgn: int = 0
    fft_multiplier: int = 1
    rowwisenormalize: bool = False
    donotnormalize: bool = True
    variant: int = 0  # 0 = FFT (BscanFFT.cpp), 1 = DARK (BscanDark.cpp)
    weight_mode: int = 0  # 0 = reference quirk fractionalk[nearestkindex[q]], 1 = corrected fractionalk[q]
    bscanthreshold: float = -30.0  # BscanFFT.cpp:385
    clampupper: bool = False
    clamp_db: float = 50.0  # BscanFFT.cpp:1252 (30.0 in BscanFFTspinjnt.cpp:1886)
    bandpassfilter: bool = False  # BscanDark.cpp:218-236, only inside zeropadrowwise
    lowpassfilter: bool = False  # BscanDark.cpp:1070-1074: lpfilter on the captured calibration frames
    output_rebin: bool = False  # BscanFFTspinjnt.cpp:1856-1862: re-bin the LINEAR B-scan before the log (that variant only)
    bscanbinx: int = 1  # BscanFFTspinjnt.cpp:795
    bscanbiny: int = 1  # BscanFFTspinjnt.cpp:797
    channelnum: int = 0  # BscanFFTwebcam.cpp:412, 1019: >= 3 sums the three channels of a BGR frame, scaled by 1/765, into a CV_64F mraw

    @property
    def opw(self) -> int:
        return self.w // self.binx  # BscanFFT.cpp:545

    @property
    def oph(self) -> int:
        return self.h // self.biny  # BscanFFT.cpp:546

    @property
    def M(self) -> int:
        return self.fft_multiplier * self.opw


# --------------------------------------------------------------------------- tables
def build_tables(p: Params):
    """lambda->k resampling tables, BscanFFT.cpp:615-698.

    Returns dict(lambdas, k, klinear, diffk, nearestkindex[int32], fractionalk).
    Every expression keeps the reference's evaluation order (all f64).
    """
    opw, m, N = p.opw, p.fft_multiplier, p.numfftpoints
    M = m * opw
    deltalambda = (p.lambdamax - p.lambdamin) / opw  # :615 (data_y.cols is int -> double)
    it = np.arange(M, dtype=np.float64)
    lambdas = p.lambdamin + it * deltalambda / float(m)  # :641
    k = cv2.divide(2 * PI, lambdas.reshape(-1, 1)).reshape(-1)  # :644 (scalar / Mat)
    kmin = 2 * PI / (p.lambdamax - deltalambda)  # :645
    kmax = 2 * PI / p.lambdamin  # :646
    deltak = (kmax - kmin) / N  # :647
    klinear = kmin + (np.arange(N, dtype=np.float64) + 1.0) * deltak  # :652
    diffk = np.zeros(M, dtype=np.float64)
    diffk[1:] = k[:-1] - k[1:]  # :667
    diffk[0] = diffk[1]  # :671
    # :673-690 first i with k[i] < klinear[f] (strict); stays 0 if none.
    # k is strictly decreasing, so the first such i is a searchsorted on the reversed array.
    kr = k[::-1]  # increasing
    cnt_less = np.searchsorted(kr, klinear, side="left")  # number of k values strictly < klinear[f]
    nk = np.where(cnt_less > 0, M - cnt_less, 0).astype(np.int32)
    fractionalk = (klinear - k[nk]) / diffk[nk]  # :695
    return dict(lambdas=lambdas, k=k, klinear=klinear, diffk=diffk, nearestkindex=nk, fractionalk=fractionalk,
                kmin=kmin, kmax=kmax, deltak=deltak)


def build_tables_linear_scan(p: Params):
    """Same as build_tables but with the reference's literal O(N*M) scan (:673-690); small sizes only."""
    t = build_tables(p)
    k, klinear = t["k"], t["klinear"]
    nk = np.zeros(p.numfftpoints, dtype=np.int32)
    for f in range(p.numfftpoints):
        for i in range(p.M):
            if k[i] < klinear[f]:
                nk[f] = i
                break
    return nk


def barthann_window(opw: int) -> np.ndarray:
    """Modified Bartlett-Hann window, BscanFFT.cpp:936-944 (x = float(p)/float(opw-1) in f32, rest f64)."""
    nn = np.arange(opw, dtype=np.float32)
    NN = np.float32(opw - 1)
    x = (nn / NN).astype(np.float64)  # f32 division, then promoted
    return 0.62 - 0.48 * np.abs(x - 0.5) + 0.38 * np.cos(2 * PI * (x - 0.5))


# --------------------------------------------------------------------------- helpers
def smoothmovavg(sm: np.ndarray, sn: int) -> np.ndarray:
    """BscanFFT.cpp:247-304: 2n+1 taps, centre weight 2, out-of-range taps replaced by centre, / 2 / (n+1)."""
    rows, cols = sm.shape
    out = np.empty_like(sm)
    j = np.arange(cols)
    for si in range(rows):
        src = sm[si]
        ssum = np.zeros(cols, dtype=np.float64)
        for sk in range(-sn, sn + 1):
            idx = j + sk
            ok = (idx > -1) & (idx < cols)
            ssum = ssum + np.where(ok, src[np.clip(idx, 0, cols - 1)], src)
        ssum = ssum + src
        out[si] = ssum / 2 / (sn + 1)
    return out


def normalizerows(src: np.ndarray, lo: float, hi: float) -> np.ndarray:
    """BscanFFT.cpp:88-97."""
    dst = np.empty_like(src)
    for ii in range(src.shape[0]):
        dst[ii : ii + 1] = cv2.normalize(src[ii : ii + 1], None, lo, hi, cv2.NORM_MINMAX)
    return dst


def _swap_halves(a: np.ndarray) -> np.ndarray:
    cx = a.shape[1] // 2
    out = a.copy()
    out[:, :cx] = a[:, cx : 2 * cx]
    out[:, cx : 2 * cx] = a[:, :cx]
    return out


def zeropadrowwise(sm: np.ndarray, sn: int, bandpassfilter: bool = False) -> np.ndarray:
    """BscanFFT.cpp:180-245 (BscanDark.cpp:169-254 with the band-pass block 218-236). Returns f64."""
    numcols = sm.shape[1]
    newnumcols = numcols * sn
    orig = sm.astype(np.float32)  # :209
    ft = cv2.dft(aligned_copy(orig), aligned_empty(orig.shape + (2,), np.float32), flags=cv2.DFT_SCALE | cv2.DFT_COMPLEX_OUTPUT | cv2.DFT_ROWS)  # :211
    ft = _swap_halves(ft)  # :215-227
    if bandpassfilter:
        cols = ft.shape[1]
        dcl = cols // 2 - int(math.floor(cols / 10))
        dcr = cols // 2 + int(math.floor(cols / 10))
        ft[:, 0:dcl] = 0
        ft[:, dcr : dcr + dcl] = 0
        dcvals = 3
        dcl2 = cols // 2 - dcvals
        ft[:, dcl2 : dcl2 + 2 * dcvals] = 0
    pad = int(math.floor((newnumcols - numcols) / 2))
    ftzp = cv2.copyMakeBorder(ft, 0, 0, pad, pad, cv2.BORDER_CONSTANT, value=0.0)  # :229
    ftzp = _swap_halves(ftzp)  # :233-239
    inv = cv2.dft(aligned_copy(ftzp), aligned_empty(ftzp.shape[:2], np.float32), flags=cv2.DFT_INVERSE | cv2.DFT_REAL_OUTPUT | cv2.DFT_ROWS)  # :241
    return inv.astype(np.float64)  # :242


def lpfilter(sm: np.ndarray) -> np.ndarray:
    """BscanDark.cpp:119-167: row-wise FFT-domain low-pass (keeps the centre 20 % of the shifted spectrum). Returns f64."""
    orig = sm.astype(np.float32)
    ft = cv2.dft(aligned_copy(orig), aligned_empty(orig.shape + (2,), np.float32), flags=cv2.DFT_SCALE | cv2.DFT_COMPLEX_OUTPUT | cv2.DFT_ROWS)
    ft = _swap_halves(ft)
    cols = ft.shape[1]
    dcl = cols // 2 - int(math.floor(cols / 10))
    dcr = cols // 2 + int(math.floor(cols / 10))
    ft[:, 0:dcl] = 0
    ft[:, dcr : dcr + dcl] = 0
    ft = _swap_halves(ft)
    inv = cv2.dft(aligned_copy(ft), aligned_empty(ft.shape[:2], np.float32), flags=cv2.DFT_INVERSE | cv2.DFT_REAL_OUTPUT | cv2.DFT_ROWS)
    return inv.astype(np.float64)


def bin_frame(mraw: np.ndarray, p: Params) -> np.ndarray:
    """medianBlur + INTER_AREA binning on the integer frame, BscanFFT.cpp:953-958 (x/y: BscanFFTspinjnt.cpp:1553)."""
    if p.channelnum >= 3:  # BscanFFTwebcam.cpp:1021-1037: mraw (h x w x 3, u8) -> CV_64F sum of the channels * 0.00130718954
        assert mraw.ndim == 3 and mraw.shape[2] == 3 and mraw.dtype == np.uint8
        s64 = mraw[:, :, 0].astype(np.float64)
        s64 = s64 + mraw[:, :, 1].astype(np.float64)
        s64 = s64 + mraw[:, :, 2].astype(np.float64)
        mraw = s64 * 0.00130718954
        if p.mediann > 0:
            raise ValueError("cv::medianBlur does not take CV_64F: the reference throws with mediann > 0 and channelnum >= 3")
    m = cv2.medianBlur(mraw, p.mediann) if p.mediann > 0 else mraw
    if p.binx == 1 and p.biny == 1:
        return m.copy()
    return cv2.resize(m, None, fx=1.0 / p.binx, fy=1.0 / p.biny, interpolation=cv2.INTER_AREA)


# --------------------------------------------------------------------------- the block
class Oracle:
    """State the reference keeps between frames: tables, window, yb/yp/yd, the accumulator."""

    def __init__(self, p: Params, strict: bool = False):
        if p.numfftpoints < p.M:
            raise ValueError("numfftpoints < multiplier*opw is out-of-bounds in the reference (BscanFFT.cpp:1170)")
        self.p = p
        self.strict = strict
        self.t = build_tables(p)
        self.win = barthann_window(p.opw).reshape(1, -1)
        self.yb = None  # reference default is zeros (BscanFFT.cpp:562) -> division by zero; we require it
        self.yp = np.zeros((p.oph, p.opw), dtype=np.float64)  # :563
        self.yd = np.zeros((p.oph, p.opw), dtype=np.float64)
        nk = self.t["nearestkindex"]
        self.wq = self.t["fractionalk"][nk] if p.weight_mode == 0 else self.t["fractionalk"]  # :1170 quirk
        self.reset()

    def reset(self):
        p = self.p
        self.acc = np.zeros((p.oph, p.numdisplaypoints), dtype=np.float64)  # bscantransposed :932
        self.indextemp = 0

    # calibration ------------------------------------------------------------------
    def set_background(self, yb):
        self.yb = np.asarray(yb, dtype=np.float64)

    def set_pishift(self, yp):
        self.yp = np.zeros_like(self.yp) if yp is None else np.asarray(yp, dtype=np.float64)

    def set_dark(self, yd):
        self.yd = np.asarray(yd, dtype=np.float64)

    def calib_mean_of_frames(self, frames) -> np.ndarray:
        """Mean of A binned frames, BscanFFT.cpp:1041-1062 (accumulate, then / averagestoggle)."""
        acc = np.zeros((self.p.oph, self.p.opw), dtype=np.float64)
        for f in frames:
            acc += bin_frame(f, self.p).astype(np.float64)
        return acc * (1.0 / len(frames))  # Mat / double == Mat * (1/double) in OpenCV

    def calib_capture(self, frames, lowpass: bool = False) -> np.ndarray:
        """The full capture on keys b / o / r / t: accumulate data_y (after median, binning, convertTo, smoothmovavg) over the
        frames, then normalizerows / normalize to [0.0001, 1] or divide by the count - with the reference's if / if-else
        structure (BscanFFT.cpp:1041-1057; BscanDark.cpp:1045-1067, 1107-1114, 1180-1187) - and optionally lpfilter
        (BscanDark.cpp:1070-1074, 1145-1149, 1218-1222)."""
        p = self.p
        acc = np.zeros((p.oph, p.opw), dtype=np.float64)
        for f in frames:
            y = bin_frame(f, p).astype(np.float64)
            if p.movavgn > 0:
                y = smoothmovavg(y, p.movavgn)
            acc += y
        out = acc
        if p.rowwisenormalize:
            out = normalizerows(out, 0.0001, 1)
        if not p.donotnormalize:
            out = cv2.normalize(out, None, 0.0001, 1, cv2.NORM_MINMAX)
        else:
            out = out * (1.0 / len(frames))
        if lowpass:
            out = lpfilter(out)
        return out

    def calib_capture_pishift(self, frame) -> np.ndarray:
        """Key p: data_y.copyTo(data_yp) of ONE frame (after median, binning, convertTo, smoothmovavg), then normalised like
        data_y itself - normalizerows(.., 0, 1) / normalize(.., 0, 1, NORM_MINMAX) - BscanFFT.cpp:1081, 1092-1096."""
        p = self.p
        y = bin_frame(frame, p).astype(np.float64)
        if p.movavgn > 0:
            y = smoothmovavg(y, p.movavgn)
        if p.rowwisenormalize:
            y = normalizerows(y, 0, 1)
        if not p.donotnormalize:
            y = cv2.normalize(y, None, 0, 1, cv2.NORM_MINMAX)
        return y

    # per-frame stages ------------------------------------------------------------
    def linearised(self, mraw: np.ndarray, dump: dict | None = None) -> np.ndarray:
        """Everything up to data_ylin (f64, oph x N): BscanFFT.cpp:953-1177."""
        p = self.p
        opm = bin_frame(mraw, p)
        y = opm.astype(np.float64)  # :987
        if p.movavgn > 0:
            y = smoothmovavg(y, p.movavgn)  # :990
        if p.variant == 1:
            y = y - self.yd  # BscanDark.cpp:1269
        if p.rowwisenormalize:
            y = normalizerows(y, 0, 1)  # :1126
        if not p.donotnormalize:
            y = cv2.normalize(y, None, 0, 1, cv2.NORM_MINMAX)  # :1128
        y = cv2.divide(y - self.yp, self.yb)  # :1132
        if self.strict:
            for r in range(y.shape[0]):  # :1135-1143
                meanval = cv2.mean(y[r : r + 1])[0]
                y[r : r + 1] = cv2.multiply(y[r : r + 1] - meanval, self.win)
        else:
            y = (y - y.mean(axis=1, keepdims=True)) * self.win
        if dump is not None:
            dump["apodised"] = y.copy()
        if p.fft_multiplier > 1:
            y = zeropadrowwise(y, p.fft_multiplier, p.bandpassfilter)  # :1146
            if dump is not None:
                dump["upsampled"] = y.copy()
        nk = self.t["nearestkindex"]
        N = p.numfftpoints
        slopes = np.empty_like(y)
        slopes[:, 1:] = y[:, 1:] - y[:, :-1]  # :1156
        slopes[:, 0] = slopes[:, 1]  # :1161
        ylin = np.zeros((y.shape[0], N), dtype=np.float64)  # cols 0 and N-1 never written (:1164)
        q = np.arange(1, N - 1)
        i = nk[q]
        ylin[:, q] = y[:, i] + self.wq[q] * slopes[:, i]  # :1169-1171
        if dump is not None:
            dump["ylin"] = ylin
        return ylin

    def magnitude(self, ylin: np.ndarray) -> np.ndarray:
        """Row inverse DFT (unscaled, f32) + magnitude: BscanFFT.cpp:1181-1190. Returns f32 oph x N."""
        re = ylin.astype(np.float32)  # Mat_<float>(data_ylin)
        c = aligned_empty(re.shape + (2,), np.float32)
        c[..., 0] = re
        c[..., 1] = 0
        c = cv2.dft(c, aligned_empty(c.shape, np.float32), flags=cv2.DFT_ROWS | cv2.DFT_INVERSE)
        return cv2.magnitude(aligned_copy(c[..., 0]), aligned_copy(c[..., 1]), aligned_empty(re.shape, np.float32))

    def finalise(self, acc: np.ndarray, averages: int):
        """BscanFFT.cpp:1220-1255. Returns (bscan linear f64, bscandb f64, bscandisp u8), all D x oph."""
        p = self.p
        bscan = acc.T * (1.0 / averages)  # :1220-1221
        bscan = bscan + 0.00001  # :1222
        if p.output_rebin:
            bscan = spinjnt_rebin(bscan, p.bscanbinx, p.bscanbiny, p.binx, p.biny)
        bscanlog = cv2.log(bscan)  # :1235
        bscandb = bscanlog * (20.0 * (1.0 / 2.303))  # :1237
        bscandb[1] = bscandb[4]  # :1239
        bscandb[0] = bscandb[4]  # :1240
        disp = np.maximum(bscandb[: p.numdisplaypoints], p.bscanthreshold)  # :1243-1247
        if p.clampupper:
            disp = disp.copy()
            disp[5, 5] = p.clamp_db  # :1252
        disp = cv2.normalize(disp, None, 0, 1, cv2.NORM_MINMAX)  # :1254
        disp8 = np.clip(np.rint(disp * 255.0), 0, 255).astype(np.uint8)  # :1255 saturate(cvRound)
        return bscan, bscandb, disp8

    def push_frame(self, mraw: np.ndarray, dump: dict | None = None):
        """One iteration of the reference's while(1) body. Returns None or (bscandb, bscandisp)."""
        p = self.p
        ylin = self.linearised(mraw, dump)
        mag = self.magnitude(ylin)
        if dump is not None:
            dump["mag"] = mag
        if self.indextemp < p.averages:  # :1193
            self.acc += mag[:, : p.numdisplaypoints].astype(np.float64)
            self.indextemp += 1
        if self.indextemp >= p.averages:  # :1211
            bscan, db, disp = self.finalise(self.acc, p.averages)
            if dump is not None:
                dump["bscan"] = bscan
            self.reset()
            return db, disp
        return None

    def process_bscans(self, frames: np.ndarray, want_linear: bool = False):
        """frames: (nframes, h, w) integer array; nframes % averages == 0.  Returns (u8 [nB,D,oph], dB f64 [nB,D,oph]) and,
        with want_linear, the linear `bscan` f64 [nB,D,oph] (BscanFFT.cpp:1220-1222) as a third element."""
        p = self.p
        assert frames.shape[0] % p.averages == 0
        nB = frames.shape[0] // p.averages
        out8 = np.empty((nB, p.numdisplaypoints, p.oph), dtype=np.uint8)
        outdb = np.empty((nB, p.numdisplaypoints, p.oph), dtype=np.float64)
        outlin = np.empty((nB, p.numdisplaypoints, p.oph), dtype=np.float64) if want_linear else None
        b = 0
        for f in frames:
            dump = {} if want_linear else None
            r = self.push_frame(f, dump)
            if r is not None:
                outdb[b], out8[b] = r
                if want_linear:
                    outlin[b] = dump["bscan"]
                b += 1
        return (out8, outdb, outlin) if want_linear else (out8, outdb)


def jlockin_display(bscan: np.ndarray, jscansave: np.ndarray, bscanthreshold: float) -> np.ndarray:
    """The 'Bscan subtracted' image of the J0 lock-in, BscanFFT.cpp:1225-1231 and 1257-1267.  bscan / jscansave: linear f64
    D x oph (finalise()[0]).  No DC-row mask and no clampupper on this image in the reference.  Returns u8 D x oph."""
    jdiff = np.asarray(bscan, dtype=np.float64) - np.asarray(jscansave, dtype=np.float64)  # :1227
    positivediff = cv2.max(jdiff, 0.0)  # makeonlypositive, :173-178, 1229
    positivediff = positivediff + 0.001  # :1230
    bscansublog = cv2.log(positivediff)  # :1260
    manual = bscansublog * (20.0 * (1.0 / 2.303))  # :1261
    manual = np.maximum(manual, bscanthreshold)  # :1264
    manual = cv2.normalize(manual, None, 0, 1, cv2.NORM_MINMAX)  # :1266
    return np.clip(np.rint(manual * 255.0), 0, 255).astype(np.uint8)  # :1267


def colormap_jet(img8: np.ndarray) -> np.ndarray:
    """applyColorMap(., COLORMAP_JET), BscanFFT.cpp:1268, 1284: u8 (...) -> BGR u8 (..., 3)."""
    a = np.ascontiguousarray(img8, dtype=np.uint8)
    return cv2.applyColorMap(a.reshape(-1, 1), cv2.COLORMAP_JET).reshape(a.shape + (3,))


def spinjnt_rebin(bscan: np.ndarray, bscanbinx: int, bscanbiny: int, binvaluex: int, binvaluey: int) -> np.ndarray:
    """BscanFFTspinjnt.cpp:1856-1862: re-binning of the linear B-scan in front of the log, active whenever any factor exceeds 1."""
    if not (bscanbinx > 1 or bscanbiny > 1 or binvaluex > 1 or binvaluey > 1):  # :1856
        return bscan
    multiplyfactor = bscanbinx * bscanbiny * binvaluex * binvaluey  # :835 (int)
    bscanbinned = cv2.resize(np.ascontiguousarray(bscan), None, fx=1.0 / bscanbinx, fy=1.0 / bscanbiny, interpolation=cv2.INTER_AREA)  # :1859
    # :1860 - int * Mat is a MatExpr scaling (one multiplication per element); fx carries binvaluey, the reference's quirk
    return cv2.resize(bscanbinned * float(multiplyfactor), None, fx=bscanbinx * binvaluey, fy=bscanbiny, interpolation=cv2.INTER_CUBIC)


def dark_background(yr, yd, ys):
    """BscanDark.cpp:996: data_yb = (data_yr - data_yd) + (data_ys - data_yd)."""
    return (yr - yd) + (ys - yd)
