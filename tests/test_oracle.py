"""CPU tests of the oracle itself (the checker): committed golden vectors, the reference's literal loops vs the
vectorised forms, and a physics known-answer test on the reference's own fixture frames."""
import os

import numpy as np
import pytest

from util import GOLDEN, ROOT, oracle_params

from oracle.abcoct_oracle import (Oracle, barthann_window, build_tables, build_tables_linear_scan, smoothmovavg,
                                  zeropadrowwise)


def test_tables_match_literal_linear_scan():
    """searchsorted form == the reference's O(N*M) first-k[i]<klinear[f] scan (BscanFFT.cpp:673-690)."""
    for w, m, N in [(128, 1, 128), (160, 4, 640), (320, 2, 700), (96, 1, 128)]:
        p = oracle_params(w=w, h=4, numfftpoints=N, fft_multiplier=m, lambdamin=840.5e-9, lambdamax=859.5e-9)
        t = build_tables(p)
        assert np.array_equal(t["nearestkindex"], build_tables_linear_scan(p))
        nk, fr = t["nearestkindex"], t["fractionalk"]
        assert (np.diff(nk) <= 0).all() and nk.min() >= 1 and nk.max() <= p.M - 1
        assert (fr > 0).all() and (fr <= 1.0 + 1e-9).all()


def test_golden_wang128_regression_and_physics():
    """The committed vectors for the reference's fixture frames reproduce, and the two reflectors of
    Matlab files/wangOCTimg.m (50 um apart) appear as two peaks whose bin spacing matches z = n*pi/(kmax-kmin)."""
    z = np.load(os.path.join(GOLDEN, "wang128.npz"))
    p = oracle_params(w=128, h=96, numfftpoints=128, numdisplaypoints=64, lambdamin=816e-9, lambdamax=884e-9)
    o = Oracle(p, strict=True)
    o.set_background(z["bg"].astype(np.float64))
    d = {}
    db, disp = o.push_frame(z["img"], d)
    assert np.array_equal(disp, z["disp"])
    assert np.allclose(db, z["db"], rtol=0, atol=2e-5)
    assert np.array_equal(o.t["nearestkindex"], z["nk"]) and np.array_equal(o.t["fractionalk"], z["frac"])
    # physics: a reflector at depth ls in a sample of index ns = 1.38 (wangOCTimg.m:14, 43-46: phase 2 k ns ls) lands in
    # bin ns * ls * (kmax - kmin) / pi (depth per bin pi / (kmax - kmin), wangOCTrec4.m:200-202); row ii holds reflectors
    # at ii um (reflectivity 0.5) and ii + 50 um (0.25).
    t = o.t
    bins_per_m = 1.38 * (t["kmax"] - t["kmin"]) / np.pi
    for row in (30, 40, 70):
        col = db[6:, row]  # one A-scan, skip the masked DC rows
        pk = np.argsort(col)[::-1]
        first = int(pk[0])
        second = int(next(i for i in pk[1:] if abs(i - first) > 3))
        lo, hi = sorted((first + 6, second + 6))
        assert abs(lo - (row + 1) * 1e-6 * bins_per_m) <= 1.5, (row, lo, (row + 1) * 1e-6 * bins_per_m)
        assert abs(hi - (row + 51) * 1e-6 * bins_per_m) <= 1.5, (row, hi, (row + 51) * 1e-6 * bins_per_m)


@pytest.mark.parametrize("name", ["synth_fft_1280x32", "synth_dark_1280x16_a4", "synth_fft_1024x17_n2048_clamp"])
def test_golden_synth_regression(name):
    """Vectorised oracle == the committed outputs (generated with the strict per-row form)."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    w, h, N, D, A, variant, seed, clamp, wm = [int(x) for x in z["params"]]
    p = oracle_params(w=w, h=h, numfftpoints=N, numdisplaypoints=D, averages=A, variant=variant, lambdamin=840.5e-9,
                      lambdamax=859.5e-9, bscanthreshold=float(z["thr"]), clampupper=bool(clamp), weight_mode=wm)
    o = Oracle(p, strict=False)
    if "yd" in z.files:
        o.set_dark(z["yd"])
    o.set_background(z["yb"])
    out8, outdb = o.process_bscans(z["frames"])
    assert np.abs(out8.astype(int) - z["out8"].astype(int)).max() <= 1
    assert (out8 != z["out8"]).mean() < 1e-3
    assert np.allclose(outdb, z["outdb"], rtol=0, atol=2e-4)
    if clamp:
        # element (5,5) is forced to clamp_db before the min-max normalise (BscanFFT.cpp:1248-1254)
        v = np.maximum(outdb[0], p.bscanthreshold)
        v[5, 5] = p.clamp_db
        assert abs(int(out8[0, 5, 5]) - round(255 * (p.clamp_db - v.min()) / (v.max() - v.min()))) <= 1


def test_strict_and_vectorised_agree():
    from fdoct_b200 import synth

    p = oracle_params(w=256, h=6, numfftpoints=512, numdisplaypoints=128, averages=2, fft_multiplier=2, lambdamin=840.5e-9,
                      lambdamax=859.5e-9)
    frames = synth.make_frames(4, 256, 6, seed=5)
    yb = synth.make_background_frames(2, 256, 6, seed=6).mean(axis=0)
    outs = []
    for strict in (True, False):
        o = Oracle(p, strict=strict)
        o.set_background(yb)
        outs.append(o.process_bscans(frames))
    assert np.abs(outs[0][0].astype(int) - outs[1][0].astype(int)).max() <= 1
    assert np.allclose(outs[0][1], outs[1][1], rtol=0, atol=1e-6)


def test_window_and_helpers():
    w = barthann_window(640)
    assert w.shape == (640,) and abs(w[0]) < 1e-6 and abs(w[-1]) < 1e-6 and abs(w.max() - 1.0) < 1e-3
    a = np.arange(12, dtype=np.float64).reshape(2, 6)
    s = smoothmovavg(a, 1)  # 3 taps + centre again, /4; edge taps replaced by the centre sample
    assert np.allclose(s[0, 1:-1], a[0, 1:-1]) and np.isclose(s[0, 0], (0 + 0 + 1 + 0) / 4) and np.isclose(s[0, -1], (4 + 5 + 5 + 5) / 4)
    # Fourier upsample of a band-limited row: result = m * irfft(padded rfft)
    x = np.cos(2 * np.pi * 3 * np.arange(32) / 32)[None, :]
    up = zeropadrowwise(x, 2)
    assert up.shape == (1, 64) and np.allclose(up[0, ::2], x[0], atol=1e-5)


def test_upsample_closed_form_used_by_the_cuda_path():
    """zeropadrowwise == zero-padding the scaled spectrum with the Nyquist bin dropped (what DFT_REAL_OUTPUT does), and two real
    rows can share one complex transform each way: the closed form rowprep_kernel implements (prep_kernels.cu)."""
    rng = np.random.default_rng(0)
    for W, m, bp in [(32, 2, False), (640, 4, False), (1920, 2, False), (720, 4, True), (96, 3, False)]:
        M = m * W
        x = rng.normal(size=(2, W)) * 100 + 1000
        ref = zeropadrowwise(x, m, bp)
        z = x[0].astype(np.float32).astype(np.float64) + 1j * x[1].astype(np.float32).astype(np.float64)
        Z = np.fft.fft(z) / W
        lo, hi = (3, W // 10) if bp else (0, W // 2)
        Y = np.zeros(M, dtype=complex)
        for k in range(lo, hi):
            Y[k] = Z[k]
            if k >= 1:
                Y[M - k] = Z[W - k]
        out = np.fft.ifft(Y) * M
        scale = max(np.abs(ref).max(), 1e-30)
        assert np.abs(out.real - ref[0]).max() <= 2e-6 * scale and np.abs(out.imag - ref[1]).max() <= 2e-6 * scale, (W, m, bp)


def test_opencv_rules_the_cuda_path_reimplements():
    """The integer / rounding rules of the OpenCV calls on the path, pinned against cv2 itself: these are what
    bin_kernel, median_kernel, normalise_part and the host capture code re-implement (prep_kernels.cu, recon_kernel.cuh)."""
    import cv2

    rng = np.random.default_rng(9)
    for dt, hi in ((np.uint16, 65535), (np.uint8, 255)):
        img = rng.integers(0, hi + 1, size=(24, 36)).astype(dt)
        # INTER_AREA with integer factors: 2 x 2 -> (sum + 2) >> 2; anything else -> f32 sum * (1 / area), round half to even
        for bx, by in ((2, 2), (3, 3), (2, 1), (1, 2), (4, 4), (3, 2)):
            ref = cv2.resize(img, None, fx=1.0 / bx, fy=1.0 / by, interpolation=cv2.INTER_AREA)
            blocks = img.reshape(24 // by, by, 36 // bx, bx).astype(np.uint32).sum(axis=(1, 3))
            if (bx, by) == (2, 2):
                mine = (blocks + 2) >> 2
            else:
                mine = np.rint(blocks.astype(np.float32) * np.float32(1.0 / (bx * by)))
            assert np.array_equal(ref, np.clip(mine, 0, hi).astype(dt)), (dt, bx, by)
        # medianBlur replicates the border
        for k in (3, 5):
            ref = cv2.medianBlur(img, k)
            pad = np.pad(img, k // 2, mode="edge")
            win = np.lib.stride_tricks.sliding_window_view(pad, (k, k)).reshape(24, 36, k * k)
            assert np.array_equal(ref, np.sort(win, axis=2)[:, :, k * k // 2]), (dt, k)
    # normalize(NORM_MINMAX): dst = src * scale + (a - min * scale), scale = (b - a) / (max - min); flat image -> all a
    x = rng.normal(size=(5, 7)) * 40.0
    for a, b in ((0.0, 1.0), (0.0001, 1.0)):
        mn, mx = x.min(), x.max()
        sc = (b - a) / (mx - mn)
        assert np.allclose(cv2.normalize(x, None, a, b, cv2.NORM_MINMAX), x * sc + (a - mn * sc), rtol=0, atol=1e-15)
    assert (cv2.normalize(np.full((3, 3), 2.5), None, 0, 1, cv2.NORM_MINMAX) == 0).all()
    # convertTo(CV_8UC1, 255.0): round half to even, saturate (convertScaleAbs shares the saturate_cast<uchar>(double) path)
    v = np.array([[0.5 / 255, 1.5 / 255, 2.5 / 255, 254.5 / 255, 1.2, 0.0]])
    assert np.array_equal(cv2.convertScaleAbs(v, alpha=255.0), np.array([[0, 2, 2, 254, 255, 0]], np.uint8))
    # dft: DFT_ROWS | DFT_INVERSE is the unscaled e^{+i...} transform; DFT_SCALE forward divides by the row length
    z = (rng.normal(size=(3, 40)) + 1j * rng.normal(size=(3, 40))).astype(np.complex64)
    c = np.stack([z.real, z.imag], axis=-1).astype(np.float32)
    inv = cv2.dft(c, flags=cv2.DFT_ROWS | cv2.DFT_INVERSE)
    assert np.allclose(inv[..., 0] + 1j * inv[..., 1], np.fft.ifft(z, axis=1) * 40, rtol=0, atol=2e-4)
    fwd = cv2.dft(c[..., 0].copy(), flags=cv2.DFT_ROWS | cv2.DFT_SCALE | cv2.DFT_COMPLEX_OUTPUT)
    assert np.allclose(fwd[..., 0] + 1j * fwd[..., 1], np.fft.fft(z.real, axis=1) / 40, rtol=0, atol=1e-5)


def test_oracle_rejects_undefined_reference_behaviour():
    with pytest.raises(ValueError):
        Oracle(oracle_params(w=256, h=4, numfftpoints=128))  # N < M reads past fractionalk (BscanFFT.cpp:1170)


def test_jlockin_display_known_answer():
    """BscanFFT.cpp:1225-1231, 1257-1267 on a hand-made pair of linear images, against plain f64 arithmetic."""
    from oracle.abcoct_oracle import jlockin_display

    bscan = np.array([[5.0, 1.0, 0.5], [2.0, 2.0, 100.0]])
    jscan = np.array([[1.0, 3.0, 0.5], [1.0, 2.5, 0.0]])
    pos = np.maximum(bscan - jscan, 0.0) + 0.001  # [[4.001, .001, .001], [1.001, .001, 100.001]]
    db = np.maximum(20.0 * np.log(pos) / 2.303, -30.0)  # 0.001 -> -60 dB -> clamped to the threshold
    want = np.rint((db - db.min()) / (db.max() - db.min()) * 255.0).astype(np.uint8)
    got = jlockin_display(bscan, jscan, -30.0)
    assert np.array_equal(got, want)
    assert got[0, 1] == 0 and got[1, 2] == 255 and 0 < got[1, 0] < got[0, 0] < 255
    # the lock-in reference against itself: positivediff == 0.001 everywhere, a flat image, which cv::normalize maps to zeros
    assert (jlockin_display(bscan, bscan, -30.0) == 0).all()


def test_jet_table_in_the_cuda_sources_is_opencvs():
    """fdoct_b200/csrc/jet_lut.h (generated by tools/gen_jet_lut.py) must be cv2's COLORMAP_JET table, entry for entry."""
    import re

    import cv2
    from oracle.abcoct_oracle import colormap_jet

    txt = open(os.path.join(ROOT, "fdoct_b200", "csrc", "jet_lut.h")).read()
    body = txt[txt.index("= {") + 3:txt.rindex("}")]
    tab = np.array([int(x) for x in re.findall(r"\d+", body)], dtype=np.uint8).reshape(256, 3)
    assert np.array_equal(tab, colormap_jet(np.arange(256, dtype=np.uint8)))
    assert tuple(tab[0]) == (128, 0, 0) and tuple(tab[255]) == (0, 0, 128)  # BGR: dark blue ... dark red
    img = np.array([[0, 128], [255, 7]], dtype=np.uint8)
    assert np.array_equal(colormap_jet(img), cv2.applyColorMap(img, cv2.COLORMAP_JET))


def test_golden_consumers_regression():
    """The committed consumer fixture (linear bscan, J0 lock-in display, JET images) against a fresh run of the vectorised oracle."""
    from oracle.abcoct_oracle import colormap_jet, jlockin_display

    z = np.load(os.path.join(GOLDEN, "consumers_1024x24_a2.npz"))
    w, h, N, D, A, seed = [int(x) for x in z["params"]]
    thr = float(z["thr"])
    p = oracle_params(w=w, h=h, numfftpoints=N, numdisplaypoints=D, averages=A, lambdamin=840.5e-9, lambdamax=859.5e-9, bscanthreshold=thr)
    o = Oracle(p, strict=False)
    o.set_background(z["yb"])
    _, _, jscan = o.process_bscans(z["jframes"], want_linear=True)
    out8, _, lin = o.process_bscans(z["frames"], want_linear=True)
    assert np.allclose(lin, z["lin"], rtol=2e-6, atol=0) and np.allclose(jscan[0], z["jscan"], rtol=2e-6, atol=0)
    assert np.abs(out8.astype(int) - z["out8"].astype(int)).max() <= 1
    # the J0 display and the colour mapping are functions of the stored images alone: exact
    jsub = np.stack([jlockin_display(b.astype(np.float64), z["jscan"].astype(np.float64), thr) for b in z["lin"]])
    assert np.abs(jsub.astype(int) - z["jsub"].astype(int)).max() <= 1 and (jsub != z["jsub"]).mean() < 1e-3
    assert np.array_equal(colormap_jet(z["out8"]), z["bgr"]) and np.array_equal(colormap_jet(z["jsub"]), z["jbgr"])
    assert z["jsub"].min() == 0 and z["jsub"].max() == 255


def test_spinjnt_output_rebinning_shapes():
    """BscanFFTspinjnt.cpp:1856-1862.  Shipped shape (binvaluex = 2, the other three factors 1): both cv::resize calls are copies,
    the block is `bscan *= multiplyfactor`, a constant dB offset.  With a real resampling step the bicubic overshoot next to bright
    A-scans is negative and the reference's own log() gives NaN - why abcoct_create refuses those shapes (ABCOCT_ERR_UNSUPPORTED)."""
    import dataclasses

    from fdoct_b200 import synth
    from oracle.abcoct_oracle import Oracle, Params

    w, h = 512, 12
    base = Params(w=w, h=h, binx=2, biny=1, numfftpoints=256, numdisplaypoints=128, lambdamin=840.5e-9, lambdamax=859.5e-9)
    frames = synth.make_frames(2, w, h, seed=5)
    outs = {}
    for name, kw in (("plain", {}), ("shipped", dict(output_rebin=True)), ("resampled", dict(output_rebin=True, bscanbinx=2))):
        o = Oracle(dataclasses.replace(base, **kw))
        o.set_background(o.calib_mean_of_frames(synth.make_background_frames(2, w, h, seed=6)))
        with np.errstate(invalid="ignore"):
            outs[name] = o.process_bscans(frames)[1]
    assert np.allclose(outs["shipped"] - outs["plain"], 20.0 / 2.303 * np.log(2.0), atol=1e-9)
    assert outs["resampled"].shape == outs["plain"].shape and np.isnan(outs["resampled"]).any()
