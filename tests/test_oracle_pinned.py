"""The pin of the oracle: oracle/abcoct_oracle.py against the reference's OWN processing block, compiled verbatim.

oracle/build_ref.py cuts the block (and the helper functions, the table precompute, the window loop and the frame ingest) out of
/root/reference/BscanFFT.cpp and BscanDark.cpp at build time and compiles them unmodified against oracle/cvshim, a stand-in for the
OpenCV C++ headers that forwards every OpenCV call to the same OpenCV kernels through cv2.  The modules live in oracle/_ref/ (git-
ignored, built by __graft_entry__.build() wherever /root/reference exists, shipped to the GPU box as built files).

Bar: bit-exact.  Both sides run the same OpenCV kernels on the same inputs, so every difference would be a difference in the
restatement: tables, display bytes and the f64 dB image must be IDENTICAL."""
import os

import numpy as np
import pytest

from util import GOLDEN, oracle_params

from oracle import build_ref


def _ref(name="abcoct_ref"):
    build_ref.build()  # no-op when up to date or when the reference is not on this machine
    mod = build_ref.load(name)
    if mod is None:
        pytest.skip("oracle/_ref is not built (no /root/reference on this machine and no shipped module)")
    return mod


def ref_params(p):
    assert p.binx == p.biny
    return dict(w=p.w, h=p.h, averages=p.averages, binvalue=p.binx, numfftpoints=p.numfftpoints, numdisplaypoints=p.numdisplaypoints,
                movavgn=p.movavgn, clampupper=p.clampupper, lambdamin=p.lambdamin, lambdamax=p.lambdamax, mediann=p.mediann,
                fft_multiplier=p.fft_multiplier, bscanthreshold=p.bscanthreshold, rowwisenormalize=p.rowwisenormalize,
                donotnormalize=p.donotnormalize, bandpassfilter=p.bandpassfilter)


FFT_CASES = [
    dict(w=128, h=8, numfftpoints=128, numdisplaypoints=64),
    dict(w=256, h=12, numfftpoints=256, numdisplaypoints=100, averages=3),
    dict(w=640, h=9, numfftpoints=1280, numdisplaypoints=300, fft_multiplier=2, movavgn=1),
    dict(w=256, h=16, numfftpoints=256, numdisplaypoints=128, binx=2, biny=2, mediann=3),
    dict(w=512, h=10, numfftpoints=512, numdisplaypoints=256, mediann=5, movavgn=2, averages=2),
    dict(w=320, h=10, numfftpoints=512, numdisplaypoints=200, rowwisenormalize=True),
    dict(w=320, h=10, numfftpoints=512, numdisplaypoints=200, donotnormalize=False, clampupper=True, bscanthreshold=5.0),
    dict(w=300, h=7, numfftpoints=1000, numdisplaypoints=1000, fft_multiplier=3),  # D = N, odd multiplier
    dict(w=1280, h=6, numfftpoints=2048, numdisplaypoints=1024, pishift=True),
    dict(w=250, h=6, numfftpoints=250, numdisplaypoints=125, bpp=8),
    dict(w=2048, h=4, numfftpoints=2048, numdisplaypoints=1024),  # the north_star configuration's row shape
]


@pytest.mark.parametrize("kw", FFT_CASES)
def test_oracle_equals_compiled_bscanfft_block(kw):
    from fdoct_b200 import synth
    from oracle.abcoct_oracle import Oracle

    ref = _ref()
    kw = dict(kw)
    pishift = kw.pop("pishift", False)
    p = oracle_params(lambdamin=840.5e-9, lambdamax=859.5e-9, **kw)
    nB = 2
    frames = synth.make_frames(nB * p.averages, p.w, p.h, seed=4242 + p.w)
    if p.bpp == 8:
        frames = (frames >> 8).astype(np.uint8)
    for strict in (True, False):
        o = Oracle(p, strict=strict)
        bgf = synth.make_background_frames(2, p.w, p.h, seed=77)
        yb = o.calib_mean_of_frames((bgf >> 8).astype(np.uint8) if p.bpp == 8 else bgf)
        o.set_background(yb)
        yp = None
        if pishift:
            yp = 0.05 * yb
            o.set_pishift(yp)
        o8, odb = o.process_bscans(frames)
        if strict:
            r = ref.run_block(ref_params(p), frames, np.ascontiguousarray(yb), None if yp is None else np.ascontiguousarray(yp))
            assert np.array_equal(r["nearestkindex"].ravel(), o.t["nearestkindex"])
            assert np.array_equal(r["fractionalk"].ravel(), o.t["fractionalk"])
            assert np.array_equal(r["barthannwin"].ravel(), np.asarray(o.win).ravel())
            r8, rdb = np.stack(r["bscandisp"]), np.stack(r["bscandb"])
            assert r8.shape == o8.shape == (nB, p.numdisplaypoints, p.oph)
            assert np.array_equal(r8, o8), "display bytes differ from the compiled reference block"
            assert np.array_equal(rdb, odb), f"dB image differs from the compiled reference block by {np.abs(rdb - odb).max():.3g}"
        else:  # the vectorised mode (the CPU baseline): f64 sums in another order flip single f32 ulps in front of the f32 DFT
            assert np.abs(np.stack(r["bscandb"]) - odb).max() <= 1e-4 and np.abs(r8.astype(int) - o8.astype(int)).max() <= 1


DARK_CASES = [
    dict(w=256, h=8, numfftpoints=256, numdisplaypoints=128),
    dict(w=1280, h=6, numfftpoints=1280, numdisplaypoints=640, averages=4),
    dict(w=640, h=8, numfftpoints=2560, numdisplaypoints=500, fft_multiplier=4, bandpassfilter=True),
    dict(w=640, h=8, numfftpoints=1280, numdisplaypoints=500, fft_multiplier=2, movavgn=1, mediann=3),
    dict(w=320, h=10, numfftpoints=512, numdisplaypoints=200, rowwisenormalize=True),
    dict(w=320, h=10, numfftpoints=512, numdisplaypoints=200, donotnormalize=False, binx=2, biny=2),
]


@pytest.mark.parametrize("kw", DARK_CASES)
def test_oracle_equals_compiled_bscandark_block(kw):
    from fdoct_b200 import synth
    from oracle.abcoct_oracle import Oracle, dark_background

    ref = _ref("abcoct_ref_dark")
    p = oracle_params(variant=1, lambdamin=840.5e-9, lambdamax=859.5e-9, **kw)
    nB = 2
    frames = synth.make_frames(nB * p.averages, p.w, p.h, seed=99 + p.w, dark=True)
    o = Oracle(p, strict=True)
    yd = o.calib_mean_of_frames(synth.make_dark_frames(2, p.w, p.h, seed=5))
    yr = o.calib_mean_of_frames(synth.make_background_frames(2, p.w, p.h, seed=6, dark=True))
    yb = dark_background(yr, yd, yd + 0.02 * (yr - yd))
    o.set_dark(yd)
    o.set_background(yb)
    o8, odb = o.process_bscans(frames)
    r = ref.run_block(ref_params(p), frames, np.ascontiguousarray(yb), None, np.ascontiguousarray(yd))
    assert np.array_equal(r["nearestkindex"].ravel(), o.t["nearestkindex"])
    assert np.array_equal(r["fractionalk"].ravel(), o.t["fractionalk"])
    assert np.array_equal(np.stack(r["bscandisp"]), o8)
    assert np.array_equal(np.stack(r["bscandb"]), odb)


def test_oracle_lpfilter_equals_compiled():
    """lpfilter (BscanDark.cpp:119-167), applied to the captured calibration frames."""
    from oracle.abcoct_oracle import lpfilter

    ref = _ref("abcoct_ref_dark")
    rng = np.random.default_rng(3)
    for cols in (128, 250, 1280):
        a = rng.uniform(100.0, 4000.0, size=(5, cols))
        assert np.array_equal(ref.lpfilter(a), lpfilter(a))


def test_reference_generated_golden_vectors():
    """tests/golden/ref_*.npz were written by the compiled reference block itself (tests/golden/make_ref_golden.py); the oracle must
    reproduce them bit for bit - this is the check that also runs where oracle/_ref cannot be built."""
    from oracle.abcoct_oracle import Oracle

    names = sorted(f for f in os.listdir(GOLDEN) if f.startswith("ref_") and f.endswith(".npz"))
    assert names, "no reference-generated fixtures"
    for name in names:
        z = np.load(os.path.join(GOLDEN, name), allow_pickle=False)
        kw = {k[2:]: z[k].item() for k in z.files if k.startswith("p_")}
        p = oracle_params(**kw)
        o = Oracle(p, strict=True)
        if "yd" in z.files:
            o.set_dark(z["yd"])
        o.set_background(z["yb"])
        o8, odb = o.process_bscans(z["frames"])
        assert np.array_equal(o8, z["bscandisp"]), name
        assert np.array_equal(odb, z["bscandb"]), name
        assert np.array_equal(o.t["nearestkindex"], z["nearestkindex"]) and np.array_equal(o.t["fractionalk"], z["fractionalk"])
        # the product library's host-side precompute (abcoct_build_tables, no GPU needed) against the reference's own tables
        from fdoct_b200 import api
        from util import abi_params

        nk, fr, win = api.build_tables(abi_params(p))
        assert np.array_equal(nk, z["nearestkindex"]) and np.array_equal(fr, z["fractionalk"]) and np.array_equal(win, z["barthannwin"])


def test_oracle_consumers_equal_compiled_block():
    """The consumers of a finished B-scan inside the same block: the J0 lock-in display (BscanFFT.cpp:1225-1231, 1256-1268) and the
    JET colour images (:1267, 1286) - oracle.jlockin_display / colormap_jet against the compiled reference, bit for bit."""
    from fdoct_b200 import synth
    from oracle.abcoct_oracle import Oracle, colormap_jet, jlockin_display

    ref = _ref()
    for kw in (dict(w=256, h=12, numfftpoints=256, numdisplaypoints=100), dict(w=1280, h=10, numfftpoints=1280, numdisplaypoints=640, averages=2)):
        p = oracle_params(lambdamin=840.5e-9, lambdamax=859.5e-9, bscanthreshold=-20.0, **kw)
        A = p.averages
        scene1 = synth.make_frames(A, p.w, p.h, seed=31)
        scene2 = synth.make_frames(2 * A, p.w, p.h, seed=32)
        o = Oracle(p, strict=True)
        yb = o.calib_mean_of_frames(synth.make_background_frames(2, p.w, p.h, seed=33))
        o.set_background(yb)
        _, _, lin1 = o.process_bscans(scene1, want_linear=True)
        jscansave = np.ascontiguousarray(lin1[0])  # key 'j': bscan.copyTo(jscansave), BscanFFT.cpp:1294
        o8, odb, olin = o.process_bscans(scene2, want_linear=True)
        r = ref.run_block(ref_params(p), scene2, np.ascontiguousarray(yb), None, None, jscansave)
        assert np.array_equal(np.stack(r["bscandisp"]), o8) and np.array_equal(np.stack(r["bscandb"]), odb)
        for b in range(2):
            jd = jlockin_display(olin[b], jscansave, p.bscanthreshold)
            assert np.array_equal(r["bscandispmanual"][b], jd)
            assert np.array_equal(r["cmagImanual"][b], colormap_jet(jd))
            assert np.array_equal(r["cmagI"][b], colormap_jet(o8[b]))


@pytest.mark.parametrize("kw", [
    dict(w=256, h=10, numfftpoints=256, numdisplaypoints=100),
    dict(w=512, h=12, numfftpoints=256, numdisplaypoints=100, binx=2, biny=2, mediann=3, movavgn=2),
    dict(w=320, h=9, numfftpoints=512, numdisplaypoints=200, rowwisenormalize=True),
    dict(w=320, h=9, numfftpoints=512, numdisplaypoints=200, donotnormalize=False),
    dict(w=320, h=9, numfftpoints=512, numdisplaypoints=200, rowwisenormalize=True, donotnormalize=False, movavgn=1),
])
def test_oracle_captures_equal_compiled_key_handler(kw):
    """Keys 'b' and 'p' as the reference's frame loop handles them (BscanFFT.cpp:1000-1099, compiled verbatim): 'b' accumulates
    averagestoggle frames and finishes on the next one (normalise branches or / n), 'p' copies one frame and normalises it like
    data_y.  oracle.calib_capture / calib_capture_pishift must give the same doubles, bit for bit."""
    from fdoct_b200 import synth
    from oracle.abcoct_oracle import Oracle

    ref = _ref()
    nacc = 3
    p = oracle_params(lambdamin=840.5e-9, lambdamax=859.5e-9, averages=nacc, **kw)
    frames = synth.make_background_frames(nacc + 1, p.w, p.h, seed=71)  # the frame after the last accumulated one triggers the tail
    o = Oracle(p, strict=True)
    r = ref.run_block(dict(ref_params(p), keys=[1]), frames)  # 'b' before the first frame
    assert not r["capture_pending"]
    assert np.array_equal(r["data_yb"], o.calib_capture(frames[:nacc]))
    r = ref.run_block(dict(ref_params(p), keys=[2]), frames[:1])  # 'p'
    assert not r["capture_pending"]
    assert np.array_equal(r["data_yp"], o.calib_capture_pishift(frames[0]))


@pytest.mark.parametrize("kw", [
    dict(w=256, h=10, numfftpoints=256, numdisplaypoints=100),
    dict(w=640, h=8, numfftpoints=640, numdisplaypoints=200, lowpassfilter=True, movavgn=1),
    dict(w=320, h=9, numfftpoints=512, numdisplaypoints=200, rowwisenormalize=True, lowpassfilter=True),
    dict(w=320, h=9, numfftpoints=512, numdisplaypoints=200, donotnormalize=False, binx=2, biny=2),
])
def test_oracle_dark_calibration_flow_equals_compiled_key_handler(kw):
    """BscanDark's key handler compiled verbatim (BscanDark.cpp:993-1249): keys 'o' / 'r' / 't' capture the dark, reference-arm and
    sample-arm frames (accumulate, normalise branches or / n, lpfilter), 'b' composes data_yb = (yr - yd) + (ys - yd) (:996), 'p'
    takes the pi-shift frame; then B-scans are processed with that state.  oracle.calib_capture(lowpass) / dark_background /
    calib_capture_pishift and the processing must reproduce every array bit for bit."""
    from fdoct_b200 import synth
    from oracle.abcoct_oracle import Oracle, dark_background

    ref = _ref("abcoct_ref_dark")
    nacc = 2
    p = oracle_params(variant=1, lambdamin=840.5e-9, lambdamax=859.5e-9, averages=nacc, **kw)
    w, h = p.w, p.h
    dark = synth.make_dark_frames(nacc + 1, w, h, seed=81)
    refarm = synth.make_background_frames(nacc + 1, w, h, seed=82, dark=True)
    sample = (0.1 * synth.make_background_frames(nacc + 1, w, h, seed=83, dark=True) + 0.9 * dark).astype(np.uint16)
    scene = synth.make_frames(2 * nacc + 2, w, h, seed=84, dark=True)
    # key presses: 'o' with the dark frames, 'r', 't', then 'b' (composition) and 'p' on the first scene frames, then plain processing
    frames = np.concatenate([dark, refarm, sample, scene])
    keys = [0] * len(frames)
    keys[0], keys[nacc + 1], keys[2 * (nacc + 1)] = 3, 4, 5
    keys[3 * (nacc + 1)] = 1
    keys[3 * (nacc + 1) + 1] = 2
    r = ref.run_block(dict(ref_params(p), lowpassfilter=p.lowpassfilter, keys=keys), frames)
    assert not r["capture_pending"]
    o = Oracle(p, strict=True)
    yd = o.calib_capture(dark[:nacc], lowpass=p.lowpassfilter)
    yr = o.calib_capture(refarm[:nacc], lowpass=p.lowpassfilter)
    ys = o.calib_capture(sample[:nacc], lowpass=p.lowpassfilter)
    assert np.array_equal(r["data_yd"], yd) and np.array_equal(r["data_yr"], yr) and np.array_equal(r["data_ys"], ys)
    yb = dark_background(yr, yd, ys)
    assert np.array_equal(r["data_yb"], yb)
    yp = o.calib_capture_pishift(scene[1])
    assert np.array_equal(r["data_yp"], yp)
    # the B-scans made after the last key press: frames scene[2:], averaging restarts wherever indextemp stands - compare the last one
    o.set_dark(yd)
    o.set_background(yb)
    o.set_pishift(yp)
    nproc = len(frames)  # every frame went through the block; indextemp counts all of them
    last_start = (nproc // nacc - 1) * nacc
    assert last_start >= 3 * (nacc + 1) + 2  # the last B-scan was averaged entirely after the last key press
    o8, odb = o.process_bscans(frames[last_start:last_start + nacc])
    assert np.array_equal(r["bscandisp"][-1], o8[0]) and np.array_equal(r["bscandb"][-1], odb[0])


def test_oracle_webcam_channel_sum_equals_compiled():
    """BscanFFTwebcam.cpp:1018-1038 compiled verbatim: channelnum < 3 copies one 8-bit plane, channelnum >= 3 sums the three planes
    in CV_64F and scales by 0.00130718954 - oracle.bin_frame's restatement, bit for bit."""
    from oracle.abcoct_oracle import bin_frame

    ref = _ref()
    rng = np.random.default_rng(12)
    frame = rng.integers(0, 256, size=(24, 320, 3), dtype=np.uint8)
    for c in range(3):
        assert np.array_equal(ref.webcam_mraw(frame, c), frame[:, :, c])
    p = oracle_params(w=320, h=24, bpp=8, numfftpoints=512, numdisplaypoints=100, channelnum=3)
    got = ref.webcam_mraw(frame, 3)
    assert got.dtype == np.float64 and np.array_equal(got, bin_frame(frame, p))


def test_oracle_spinjnt_rebin_equals_compiled():
    """BscanFFTspinjnt.cpp:1856-1862 compiled verbatim against oracle.spinjnt_rebin: the shipped shape (a multiplication by
    multiplyfactor) and real resampling shapes, whose bicubic overshoot is negative in places - in the reference's own code."""
    from oracle.abcoct_oracle import spinjnt_rebin

    ref = _ref()
    rng = np.random.default_rng(5)
    bscan = np.exp(rng.normal(0.0, 2.5, size=(120, 48))) + 1e-5  # linear B-scan with speckle-like contrast
    for bx, by, vx, vy in ((1, 1, 2, 1), (2, 1, 1, 1), (2, 2, 2, 1), (3, 2, 1, 1), (1, 1, 1, 2), (1, 1, 1, 1)):
        got = ref.spinjnt_rebin(bscan, bx, by, vx, vy)
        want = spinjnt_rebin(bscan, bx, by, vx, vy)
        assert got.shape == want.shape and np.array_equal(got, want), (bx, by, vx, vy)
        if (bx, by, vx, vy) == (1, 1, 2, 1):
            assert np.array_equal(got, bscan * 2.0)  # both resizes are copies
        if bx > 1 or by > 1:
            assert (got <= 0).any()  # log() of these is NaN: why the library refuses those shapes
