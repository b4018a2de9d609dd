"""Builds oracle/_ref/abcoct_ref*.so: the reference's own processing block - helpers, table precompute, window, frame ingest, key
handler (calibration captures), the block itself with the J0 lock-in display and the JET mapping - cut out of
/root/reference/BscanFFT.cpp and BscanDark.cpp (plus the webcam front end and BscanFFTspinjnt's output re-binning) at build time and
compiled verbatim against oracle/cvshim (OpenCV calls forwarded to cv2).  TEST INFRASTRUCTURE ONLY.

Only runs where /root/reference exists (this container); the built module travels to the GPU box, the fragments are deleted right
after the compile.  `python oracle/build_ref.py` or __graft_entry__.build()."""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = "/root/reference/BscanFFT.cpp"
REF_DARK = "/root/reference/BscanDark.cpp"

# name -> (first line, last line, anchor expected in the first line, anchor expected in the last line, text appended)
FRAGMENTS = {
    "frag_normalizerows": (88, 97, "inline void normalizerows(", "}", ""),
    "frag_helpers": (173, 305, "inline void makeonlypositive(", "", ""),
    "frag_tables": (615, 698, "double deltalambda = (lambdamax - lambdamin) / data_y.cols;", "}", ""),
    "frag_window": (936, 944, "for (uint p = 0; p<(opw); p++)", "}", ""),
    "frag_ingest1": (953, 958, "if (mediann>0)", "resize(m, opm, Size(), 1.0 / binvalue, 1.0 / binvalue, INTER_AREA);", ""),
    "frag_ingest2": (987, 991, "opm.convertTo(data_y, CV_64F);", "data_y = smoothmovavg(data_y, movavgn);", ""),
    "frag_keys": (1000, 1099, "if (bkeypressed == 1)", "}", ""),
    "frag_block": (1125, 1284, "data_y.convertTo(data_y, CV_64F);", "applyColorMap(bscandisp, cmagI, COLORMAP_JET);", ""),
}


FRAGMENTS_DARK = {
    "frag_normalizerows": (82, 91, "inline void normalizerows(", "}", ""),
    "frag_helpers": (111, 314, "inline void makeonlypositive(", "", ""),
    "frag_tables": (614, 697, "double deltalambda = (lambdamax - lambdamin) / data_y.cols;", "}", ""),
    "frag_window": (929, 937, "for (uint p = 0; p<(opw); p++)", "}", ""),
    "frag_ingest1": (946, 951, "if (mediann>0)", "resize(m, opm, Size(), 1.0 / binvalue, 1.0 / binvalue, INTER_AREA);", ""),
    "frag_ingest2": (980, 984, "opm.convertTo(data_y, CV_64F);", "data_y = smoothmovavg(data_y, movavgn);", ""),
    "frag_keys": (993, 1249, "if (bkeypressed == 1)", "}", ""),
    "frag_block": (1268, 1393, "data_y.convertTo(data_y, CV_64F);", "bscandisp.convertTo(bscandisp, CV_8UC1, 255.0);", ""),
}
# one more range, from BscanFFTwebcam.cpp: the channel selection / channel sum in front of the same block (compiled into abcoct_ref)
EXTRA_FRAGMENTS = {
    "abcoct_ref": [("/root/reference/BscanFFTwebcam.cpp", "frag_webcam", 1018, 1038, "split(frame, rgbchannels);", "}"),
                   ("/root/reference/BscanFFTspinjnt.cpp", "frag_spinjnt_rebin", 1856, 1862, "if(bscanbinx > 1 || bscanbiny > 1 || binvaluex > 1", "}")],
}
VARIANTS = {"abcoct_ref": (REF, FRAGMENTS, []), "abcoct_ref_dark": (REF_DARK, FRAGMENTS_DARK, ["-DREF_DARK"])}


def module_path(name: str = "abcoct_ref") -> str:
    return os.path.join(OUT, name + sysconfig.get_config_var("EXT_SUFFIX"))


def _build_one(name: str, force: bool) -> str | None:
    ref, fragments, defs = VARIANTS[name]
    so = module_path(name)
    if not os.path.exists(ref):
        return so if os.path.exists(so) else None
    srcs = [ref, os.path.join(HERE, "ref_harness.cpp"), os.path.join(HERE, "cvshim", "opencv2", "opencv.hpp"), os.path.abspath(__file__)]
    if not force and os.path.exists(so) and all(os.path.getmtime(so) >= os.path.getmtime(s) for s in srcs):
        return so
    os.makedirs(OUT, exist_ok=True)
    lines = open(ref, encoding="utf-8", errors="replace").read().split("\n")
    written = []
    try:
        for frag, (a, b, first, last, tail) in fragments.items():
            if first not in lines[a - 1] or last not in lines[b - 1]:
                raise RuntimeError(f"{ref}:{a}-{b} is not the text this recipe was written for ({frag})")
            path = os.path.join(OUT, frag + ".inc")
            with open(path, "w") as f:
                f.write("\n".join(lines[a - 1:b]) + "\n" + tail)
            written.append(path)
        for src, frag, a, b, first, last in EXTRA_FRAGMENTS.get(name, []):
            xl = open(src, encoding="utf-8", errors="replace").read().split("\n")
            if first not in xl[a - 1] or last not in xl[b - 1]:
                raise RuntimeError(f"{src}:{a}-{b} is not the text this recipe was written for ({frag})")
            path = os.path.join(OUT, frag + ".inc")
            with open(path, "w") as f:
                f.write("\n".join(xl[a - 1:b]) + "\n")
            written.append(path)
        import pybind11

        cmd = ["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-fvisibility=hidden", "-ffp-contract=off", "-w", *defs,
               "-I", os.path.join(HERE, "cvshim"), "-I", HERE, "-I", pybind11.get_include(), "-I", sysconfig.get_paths()["include"],
               os.path.join(HERE, "ref_harness.cpp"), "-o", so]
        subprocess.run(cmd, check=True)
    finally:
        for p in written:
            os.remove(p)
    return so


def build(force: bool = False):
    """Returns the paths of the built modules (None where the reference is not here and nothing was built before)."""
    return [_build_one(name, force) for name in VARIANTS]


def load(name: str = "abcoct_ref"):
    """Imports oracle/_ref/<name> (None if it was never built)."""
    import importlib

    if not os.path.exists(module_path(name)):
        return None
    for p in (OUT, os.path.join(HERE, "cvshim")):
        if p not in sys.path:
            sys.path.insert(0, p)
    return importlib.import_module(name)


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
