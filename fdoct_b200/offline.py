"""Offline volume tool: the slot `Bscancompute.bin dirname manualaverages` occupies in BscanFFTspinj
(BscanFFTspinj.cpp:1138-1139, 2388-2411; its source lives in another repository and is not part of the reference).

Input, as written by the triggered capture (BscanFFTspinj.cpp:258-448, 1683-1719, 1788-1789):
    <dirname>/Trig{NNN}-{iii}.png   Mono16 raw frames, NNN = capture (B-scan) counter, iii = 0 .. manualaverages-1
    <dirname>/spectrum.ocv          data_yb, CV_64F oph x opw (the background captured on key 'b')
    <dirname>/KTrig{NNN}-{iii}.png  optional J0 frames (ignored here)
Output, the files the live program saves on key 's' (BscanFFTspinj.cpp:2040-2045):
    <dirname>/bscan{NNN}.ocv        bscandb, CV_64F D x oph
    <dirname>/bscan{NNN}.png        bscandisp, 8-bit D x oph
    <dirname>/bscanc{NNN}.png       cmagI = applyColorMap(bscandisp, COLORMAP_JET), 8-bit BGR D x oph

    python -m fdoct_b200.offline DIRNAME MANUALAVERAGES [--ini BscanFFTspinj.ini] [--flavour spinj] [--gpus N]

PNG decoding / encoding is file IO (cv2); every bit of arithmetic runs in libabcoct.so on the GPU(s): the batch is
sharded by B-scan over `--gpus` devices inside one context (no collective, results land at the right offsets).
"""
from __future__ import annotations

import argparse
import glob
import os
import re
import sys

import numpy as np

from . import api
from .ocv import read_ocv, write_ocv

FLAVOURS = {"bscanfft": api.INI_BSCANFFT, "spinj": api.INI_SPINJ, "spinjnt": api.INI_SPINJNT, "dark": api.INI_DARK,
            "peak": api.INI_PEAK, "webcam": api.INI_WEBCAM, "sim": api.INI_SIM}
_TRIG = re.compile(r"^Trig(\d+)-(\d+)\.png$")


def list_captures(dirname: str, averages: int) -> dict[int, list[str]]:
    """{capture number: [paths of its `averages` frames in acquisition order]}; incomplete captures are an error."""
    caps: dict[int, dict[int, str]] = {}
    for p in glob.glob(os.path.join(dirname, "Trig*.png")):
        m = _TRIG.match(os.path.basename(p))
        if m:
            caps.setdefault(int(m.group(1)), {})[int(m.group(2))] = p
    out = {}
    for n in sorted(caps):
        have = caps[n]
        missing = [i for i in range(averages) if i not in have]
        if missing:
            raise FileNotFoundError(f"capture {n:03d}: frames {missing} of {averages} are missing in {dirname}")
        out[n] = [have[i] for i in range(averages)]
    return out


def load_frames(paths: list[str], w: int, h: int) -> np.ndarray:
    import cv2

    frames = np.empty((len(paths), h, w), dtype=np.uint16)
    for i, p in enumerate(paths):
        img = cv2.imread(p, cv2.IMREAD_UNCHANGED)
        if img is None or img.ndim != 2 or img.dtype != np.uint16 or img.shape != (h, w):
            raise ValueError(f"{p}: expected a {w}x{h} Mono16 PNG, got {None if img is None else (img.dtype, img.shape)}")
        frames[i] = img
    return frames


def run(dirname: str, averages: int, params: api.Params, ngpu: int = 1, write_png: bool = True, batch_bscans: int = 64) -> list[int]:
    """Reconstruct every capture of `dirname`; returns the capture numbers written."""
    import cv2

    caps = list_captures(dirname, averages)
    if not caps:
        raise FileNotFoundError(f"no Trig*.png captures in {dirname}")
    params.averages = averages
    params.bpp = 16
    yb = read_ocv(os.path.join(dirname, "spectrum.ocv")).astype(np.float64)
    oph, opw = params.h // params.biny, params.w // params.binx
    if yb.shape != (oph, opw):
        raise ValueError(f"spectrum.ocv is {yb.shape}, the ini file implies {(oph, opw)}")
    numbers = list(caps)
    with api.Context(params, ngpu=ngpu) as ctx:
        ctx.set_background(yb)
        for i in range(0, len(numbers), batch_bscans):
            chunk = numbers[i:i + batch_bscans]
            frames = load_frames([p for n in chunk for p in caps[n]], params.w, params.h)
            r = ctx.process_bscans_ex(frames, want=("bscan_u8", "bscan_db") + (("bscan_bgr",) if write_png else ()))
            for j, n in enumerate(chunk):
                write_ocv(os.path.join(dirname, f"bscan{n:03d}.ocv"), r["bscan_db"][j].astype(np.float64))
                if write_png:
                    cv2.imwrite(os.path.join(dirname, f"bscan{n:03d}.png"), r["bscan_u8"][j])
                    cv2.imwrite(os.path.join(dirname, f"bscanc{n:03d}.png"), r["bscan_bgr"][j])
    return numbers


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("dirname")
    ap.add_argument("manualaverages", type=int)
    ap.add_argument("--ini", default=None, help="ini file (default: <dirname>/BscanFFTspinj.ini, else ./BscanFFTspinj.ini)")
    ap.add_argument("--flavour", default="spinj", choices=sorted(FLAVOURS))
    ap.add_argument("--gpus", type=int, default=1)
    a = ap.parse_args(argv)
    ini = a.ini
    if ini is None:
        for cand in (os.path.join(a.dirname, "BscanFFTspinj.ini"), "BscanFFTspinj.ini"):
            if os.path.exists(cand):
                ini = cand
                break
    if ini is None:
        print("no ini file found (use --ini)", file=sys.stderr)
        return 2
    params = api.params_from_ini(ini, FLAVOURS[a.flavour])
    done = run(a.dirname, a.manualaverages, params, ngpu=a.gpus)
    print(f"{len(done)} B-scans written to {a.dirname}")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
