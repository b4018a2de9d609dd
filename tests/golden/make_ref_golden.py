"""Writes tests/golden/ref_*.npz: inputs and outputs of the REFERENCE'S OWN processing block (BscanFFT.cpp / BscanDark.cpp cut out and
compiled verbatim by oracle/build_ref.py, OpenCV calls running in cv2 - see tests/test_oracle_pinned.py).  Run HERE, where
/root/reference exists:

    python tests/golden/make_ref_golden.py

Unlike the synth_*.npz fixtures (written by the oracle), nothing in these files went through oracle/abcoct_oracle.py."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from fdoct_b200 import synth  # noqa: E402
from oracle import build_ref  # noqa: E402

DEFAULTS = dict(bpp=16, binx=1, biny=1, averages=1, mediann=0, movavgn=0, fft_multiplier=1, rowwisenormalize=False, donotnormalize=True,
                variant=0, bscanthreshold=-30.0, clampupper=False, bandpassfilter=False, lambdamin=840.5e-9, lambdamax=859.5e-9)


def mean_binned(frames, b):
    """What key 'b' accumulates (BscanFFT.cpp:1041-1057 with donotnormalize): the mean of the INTER_AREA-binned frames."""
    import cv2

    acc = 0.0
    for f in frames:
        acc = acc + (cv2.resize(f, None, fx=1.0 / b, fy=1.0 / b, interpolation=cv2.INTER_AREA) if b > 1 else f).astype(np.float64)
    return acc * (1.0 / len(frames))


def case(name, nB=2, seed=1, **kw):
    p = dict(DEFAULTS, **kw)
    dark = p["variant"] == 1
    mod = build_ref.load("abcoct_ref_dark" if dark else "abcoct_ref")
    w, h, A, b = p["w"], p["h"], p["averages"], p["binx"]
    frames = synth.make_frames(nB * A, w, h, seed=seed, dark=dark)
    yd = None
    if dark:
        yd = mean_binned(synth.make_dark_frames(2, w, h, seed=seed + 2), b)
        yr = mean_binned(synth.make_background_frames(2, w, h, seed=seed + 1, dark=True), b)
        ys = yd + 0.02 * (yr - yd)
        yb = (yr - yd) + (ys - yd)  # BscanDark.cpp:996
    else:
        yb = mean_binned(synth.make_background_frames(2, w, h, seed=seed + 1), b)
    prm = dict(w=w, h=h, averages=A, binvalue=b, numfftpoints=p["numfftpoints"], numdisplaypoints=p["numdisplaypoints"], movavgn=p["movavgn"],
               clampupper=p["clampupper"], lambdamin=p["lambdamin"], lambdamax=p["lambdamax"], mediann=p["mediann"],
               fft_multiplier=p["fft_multiplier"], bscanthreshold=p["bscanthreshold"], rowwisenormalize=p["rowwisenormalize"],
               donotnormalize=p["donotnormalize"], bandpassfilter=p["bandpassfilter"])
    r = mod.run_block(prm, frames, np.ascontiguousarray(yb), None, None if yd is None else np.ascontiguousarray(yd))
    out = dict(frames=frames, yb=yb, bscandisp=np.stack(r["bscandisp"]), bscandb=np.stack(r["bscandb"]),
               nearestkindex=r["nearestkindex"].ravel(), fractionalk=r["fractionalk"].ravel(), barthannwin=r["barthannwin"].ravel())
    if yd is not None:
        out["yd"] = yd
    for k, v in p.items():
        out["p_" + k] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, out["bscandisp"].shape)


if __name__ == "__main__":
    build_ref.build()
    case("ref_fft_1280x12", w=1280, h=12, numfftpoints=1280, numdisplaypoints=640, seed=11)
    case("ref_fft_2048x8_a2_clamp", w=2048, h=8, numfftpoints=2048, numdisplaypoints=1024, averages=2, clampupper=True, bscanthreshold=8.0, seed=12)
    case("ref_fft_960x10_bin2_up2", w=960, h=20, numfftpoints=1024, numdisplaypoints=400, binx=2, biny=2, fft_multiplier=2, mediann=3, movavgn=1,
         seed=13)
    case("ref_dark_1280x8_a4", w=1280, h=8, numfftpoints=1280, numdisplaypoints=640, averages=4, variant=1, seed=14)
    case("ref_dark_640x8_up4_bandpass", w=640, h=8, numfftpoints=2560, numdisplaypoints=600, fft_multiplier=4, bandpassfilter=True, variant=1, seed=15)
