"""Shared helpers for the parity tests: oracle <-> C-ABI parameter mapping and the tolerance definitions.

Tolerances (BASELINE.json north_star):
  * index / weight tables ............ bit-exact (int32 equal, f64 equal)
  * magnitude ........................ <= 1e-4 relative to max(|ref|, 1e-3 * A-scan max) against the oracle (OpenCV f32 DFT; the one
                                       full-size N = 2048 comparison uses 2e-3, see test_full_size_properties_c5), AND
                                       no further from an exact f64 evaluation than the reference's own f32 path is
                                       (tests/test_parity_gpu.py::test_accuracy_vs_exact_f64).  The floor is where two correct f32 FFTs stop
                                       agreeing to 1e-4 of the value: measured on B200 (profiles/r01_precision_probe.txt), OpenCV's f32 DFT
                                       deviates from the exact result by 5.8e-5 .. 6.7e-5 of max(|x|, 1e-3 * A-scan max) for N = 1024 .. 4096, the
                                       CUDA path by 4.3e-5 .. 5.2e-5, so a 1e-3 floor would test the reference's rounding noise, not parity.
  * 8-bit display image .............. +-1 LSB; exact 0 / 255 present like the reference's min-max normalise
  * noise scaling .................... the rounding noise of an f32 transform scales with the RMS of its spectrum, and the fused kernels pack
                                       two A-scans into one complex transform, so their floor follows the larger row of the pair
                                       (mag_err_pairwise).  With normalised calibration captures (rowwisenormalize / !donotnormalize: 1 / data_yb
                                       spans four decades, neighbouring rows differ by orders of magnitude, the mean removal cancels digits) the
                                       library therefore runs every stage up to data_ylin in f64 and one row per transform (round 2), and the
                                       sweep holds that regime to the same 1e-4 with the row's OWN floor (measured 2e-6 .. 8e-5; one case where
                                       two f32 transforms of a flat spectrum are 1.04e-4 apart is decided by their distances from the exact
                                       result, like test_accuracy_vs_exact_f64).
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
MAG_RTOL = 1e-4
MAG_FLOOR = float(os.environ.get("ABCOCT_TEST_MAG_FLOOR", "1e-3"))  # max(|ref|, 1e-3 * A-scan max)
DB_PER_NEPER = 20.0 * (1.0 / 2.303)  # BscanFFT.cpp:1237


def oracle_params(**kw):
    from oracle.abcoct_oracle import Params

    return Params(**kw)


def abi_params(op):
    """oracle Params -> ctypes abcoct_params (same field names)."""
    from fdoct_b200 import api

    return api.default_params(
        w=op.w, h=op.h, bpp=op.bpp, binx=op.binx, biny=op.biny, averages=op.averages, numfftpoints=op.numfftpoints,
        numdisplaypoints=op.numdisplaypoints, lambdamin=op.lambdamin, lambdamax=op.lambdamax, mediann=op.mediann,
        movavgn=op.movavgn, fft_multiplier=op.fft_multiplier, rowwisenormalize=int(op.rowwisenormalize),
        donotnormalize=int(op.donotnormalize), variant=op.variant, weight_mode=op.weight_mode,
        bscanthreshold=op.bscanthreshold, clampupper=int(op.clampupper), clamp_db=op.clamp_db,
        bandpassfilter=int(op.bandpassfilter), lowpassfilter=int(op.lowpassfilter), channelnum=int(op.channelnum),
        output_rebin=int(op.output_rebin), bscanbinx=op.bscanbinx, bscanbiny=op.bscanbiny)


def db_to_mag(db):
    """Invert bscandb = ln(bscan + 1e-5) * 20/2.303 (BscanFFT.cpp:1222-1237) back to the averaged magnitude."""
    return np.exp(np.asarray(db, dtype=np.float64) / DB_PER_NEPER) - 1e-5


def mag_err(g, r, floor=MAG_FLOOR):
    """Worst |g - r| / max(|r|, floor * max over the A-scan) of two magnitude images [nB, D, oph]."""
    colmax = np.abs(r).max(axis=-2, keepdims=True)
    den = np.maximum(np.abs(r), floor * colmax)
    return float((np.abs(g - r) / den).max())


def mag_err_pairwise(g, r, floor=MAG_FLOOR):
    """mag_err with the floor taken from the larger of the two A-scans that share one complex transform (rows 2p, 2p+1 - the
    two-for-one packing of the CUDA path): its f32 rounding noise scales with the larger row of the pair, whereas cv::dft
    transforms every row on its own.  Only matters when neighbouring A-scans differ by orders of magnitude (a normalised
    calibration frame with near-zero edge samples); for ordinary frames it equals mag_err."""
    colmax = np.abs(r).max(axis=-2, keepdims=True)  # [nB, 1, oph]
    oph = colmax.shape[-1]
    pm = colmax.copy()
    even = oph - (oph & 1)
    pair = np.maximum(colmax[..., 0:even:2], colmax[..., 1:even:2])
    pm[..., 0:even:2] = pair
    pm[..., 1:even:2] = pair
    den = np.maximum(np.abs(r), floor * pm)
    return float((np.abs(g - r) / den).max())


def mag_rel_err(db_got, db_ref, floor=MAG_FLOOR):
    """The same on dB images (rows 0,1 are copies of row 4 - the DC mask - and take part like any other row)."""
    return mag_err(db_to_mag(db_got), db_to_mag(db_ref), floor)


def assert_display_parity(got8, ref8, what=""):
    d = np.abs(got8.astype(np.int16) - ref8.astype(np.int16))
    assert d.max() <= 1, f"{what}: display differs by {d.max()} LSB at {np.argwhere(d > 1)[:5]}"
    # the reference's min-max normalise puts exact 0 and 255 in every B-scan; so must we, at the same places
    for b in range(ref8.shape[0]):
        assert got8[b].min() == ref8[b].min() and got8[b].max() == ref8[b].max(), what
        assert (got8[b][ref8[b] == 255] >= 254).all() and (got8[b][ref8[b] == 0] <= 1).all(), what
    return float((d > 0).mean())
