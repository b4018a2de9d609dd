"""Regenerate the committed golden fixtures (run HERE, where /root/reference exists):

    python tests/golden/make_golden.py

* wang128.npz   the reference's own fixtures `Matlab files/imgi.png` + `backg.png` (128 x 96, 16-bit; produced
                by `Matlab files/wangOCTimg.m`) pushed through the oracle with the reference's default
                lambdamin/lambdamax (BscanFFT.cpp:381-382, which are that generator's +-2 sigma range).
                Inputs are stored as arrays (the PNG files themselves are not copied), outputs are the oracle's.
* synth_*.npz   small seeded synthetic cases (fdoct_b200.synth) with the oracle's outputs, one per variant.
* consumers_*.npz  the consumers of a finished B-scan (BscanFFT.cpp:1220-1231, 1257-1268, 1284): the linear bscan, the J0 lock-in
                display against a second scene and the JET colour images, all from the oracle.
* *.ini         parameter files in the reference's positional layout (written by this script, not copied).
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from fdoct_b200 import synth  # noqa: E402
from oracle.abcoct_oracle import Oracle, Params, colormap_jet, dark_background, jlockin_display  # noqa: E402

REF = "/root/reference/Matlab files"


def wang128():
    img = cv2.imread(os.path.join(REF, "imgi.png"), cv2.IMREAD_UNCHANGED)
    bg = cv2.imread(os.path.join(REF, "backg.png"), cv2.IMREAD_UNCHANGED)
    assert img.shape == (96, 128) and img.dtype == np.uint16
    p = Params(w=128, h=96, numfftpoints=128, numdisplaypoints=64, lambdamin=816e-9, lambdamax=884e-9)
    o = Oracle(p, strict=True)
    o.set_background(bg.astype(np.float64))
    d = {}
    db, disp = o.push_frame(img, d)
    np.savez_compressed(os.path.join(HERE, "wang128.npz"), img=img, bg=bg, db=db.astype(np.float32), disp=disp,
                        nk=o.t["nearestkindex"], frac=o.t["fractionalk"], ylin=d["ylin"].astype(np.float32))


def synth_case(name, *, w, h, N, D, A, variant, seed, nB=2, thr=-30.0, clampupper=False, weight_mode=0):
    p = Params(w=w, h=h, numfftpoints=N, numdisplaypoints=D, averages=A, variant=variant, lambdamin=840.5e-9,
               lambdamax=859.5e-9, bscanthreshold=thr, clampupper=clampupper, weight_mode=weight_mode)
    dark = variant == 1
    frames = synth.make_frames(nB * A, w, h, seed=seed, dark=dark)
    o = Oracle(p, strict=True)
    if dark:
        yd = o.calib_mean_of_frames(synth.make_dark_frames(A, w, h, seed=seed + 2))
        yr = o.calib_mean_of_frames(synth.make_background_frames(A, w, h, seed=seed + 1, dark=True))
        ys = o.calib_mean_of_frames(synth.make_dark_frames(A, w, h, seed=seed + 3)) + 0.01 * (yr - yd)
        o.set_dark(yd)
        yb = dark_background(yr, yd, ys)
        extra = dict(yd=yd)
    else:
        yb = o.calib_mean_of_frames(synth.make_background_frames(max(A, 2), w, h, seed=seed + 1))
        extra = {}
    o.set_background(yb)
    out8, outdb = o.process_bscans(frames)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), frames=frames, yb=yb, out8=out8, outdb=outdb.astype(np.float32),
                        params=np.array([w, h, N, D, A, variant, seed, int(clampupper), weight_mode]), thr=thr, **extra)


def consumers_case(name, *, w, h, N, D, A, seed, thr=-30.0):
    p = Params(w=w, h=h, numfftpoints=N, numdisplaypoints=D, averages=A, lambdamin=840.5e-9, lambdamax=859.5e-9, bscanthreshold=thr)
    frames = synth.make_frames(2 * A, w, h, seed=seed)
    jframes = synth.make_frames(A, w, h, seed=seed + 7)  # the scene key 'j' was pressed on
    o = Oracle(p, strict=True)
    yb = o.calib_mean_of_frames(synth.make_background_frames(2, w, h, seed=seed + 1))
    o.set_background(yb)
    _, _, jscan = o.process_bscans(jframes, want_linear=True)
    out8, outdb, lin = o.process_bscans(frames, want_linear=True)
    jsub = np.stack([jlockin_display(b, jscan[0], thr) for b in lin])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), frames=frames, jframes=jframes, yb=yb, out8=out8, lin=lin.astype(np.float32),
                        jscan=jscan[0].astype(np.float32), jsub=jsub, bgr=colormap_jet(out8), jbgr=colormap_jet(jsub),
                        params=np.array([w, h, N, D, A, seed]), thr=thr)


INI_ORDER = {
    "bscanfft": ["camgain", "camtime", "bpp", "w", "h", "offsetx", "offsety", "camspeed", "cambinx", "cambiny", "usbtraffic",
                 "binvalue", "dirdescr", "averages", "numfftpoints", "saveframes", "manualaveraging", "manualaverages",
                 "saveinterferograms", "movavgn", "numdisplaypoints", "lambdamin", "lambdamax", "mediann",
                 "increasefftpointsmultiplier", "rowwisenormalize", "donotnormalize"],
}
INI_ORDER["spinj"] = INI_ORDER["bscanfft"] + ["offlinetoolpath"]
INI_ORDER["spinjnt"] = [x for x in INI_ORDER["bscanfft"] if x != "binvalue"]
_i = INI_ORDER["spinjnt"].index("dirdescr")
INI_ORDER["spinjnt"][_i:_i] = ["binvaluex", "binvaluey", "bscanbinx", "bscanbiny"]
INI_ORDER["spinjnt"] += ["offlinetoolpath"]
_noofs = [x for x in INI_ORDER["bscanfft"] if x not in ("offsetx", "offsety")]
INI_ORDER["dark"] = _noofs + ["bandpassfilter", "lowpassfilter"]
INI_ORDER["peak"] = _noofs + ["peakholdnumframes"]
INI_ORDER["webcam"] = _noofs + ["channelnum"]
INI_ORDER["sim"] = _noofs[: _noofs.index("increasefftpointsmultiplier") + 1]

INI_VALUES = dict(camgain=12, camtime=1000, bpp=16, w=1280, h=960, offsetx=0, offsety=0, camspeed=2, cambinx=1, cambiny=1,
                  usbtraffic=0, binvalue=1, binvaluex=2, binvaluey=1, bscanbinx=1, bscanbiny=1, dirdescr="golden_test", averages=8,
                  numfftpoints=1280, saveframes=0, manualaveraging=0, manualaverages=3, saveinterferograms=0, movavgn=0,
                  numdisplaypoints=640, lambdamin="840.5e-9", lambdamax="859.5e-9", mediann=0, increasefftpointsmultiplier=1,
                  rowwisenormalize=0, donotnormalize=1, offlinetoolpath="/opt/Bscancompute.bin", bandpassfilter=1, lowpassfilter=0,
                  peakholdnumframes=50, channelnum=2)


def write_inis():
    for flav, order in INI_ORDER.items():
        with open(os.path.join(HERE, f"{flav}.ini"), "w") as f:
            f.write("#golden_ini_for_%s\n#positional_layout_one_comment_token_then_one_value_token\n" % flav)
            for k in order:
                f.write(f"#{k}\n{INI_VALUES[k]}\n")


if __name__ == "__main__":
    wang128()
    synth_case("synth_fft_1280x32", w=1280, h=32, N=1280, D=640, A=1, variant=0, seed=1001)
    synth_case("synth_dark_1280x16_a4", w=1280, h=16, N=1280, D=640, A=4, variant=1, seed=1002)
    synth_case("synth_fft_1024x17_n2048_clamp", w=1024, h=17, N=2048, D=700, A=2, variant=0, seed=1005, thr=5.0, clampupper=True)
    consumers_case("consumers_1024x24_a2", w=1024, h=24, N=1024, D=400, A=2, seed=1011)
    write_inis()
    print("golden fixtures written to", HERE)
