"""The reference has no function on this path: every variant's main() carries its own copy of the processing block.  oracle/_ref
compiles the one in BscanFFT.cpp (and BscanDark.cpp's); this test shows what that pin covers - it compares the code lines of the
block (comments and all white space removed) across the variants.  Needs /root/reference (this container); skipped elsewhere."""
import difflib
import os
import re

import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "BscanFFT.cpp")), reason="the reference tree is not on this machine")


def block(name):
    """Code lines from `data_y.convertTo(data_y, CV_64F)` (BscanFFT.cpp:1125) through the 8-bit conversion of bscandisp (:1255)."""
    lines = open(os.path.join(REF, name + ".cpp"), encoding="utf-8", errors="replace").read().split("\n")
    squeezed = [re.sub(r"\s+", "", re.sub(r"//.*", "", ln)) for ln in lines]
    a = squeezed.index("data_y.convertTo(data_y,CV_64F);")
    b = next(i for i in range(a, len(squeezed)) if squeezed[i] == "bscandisp.convertTo(bscandisp,CV_8UC1,255.0);")
    return [s for s in squeezed[a:b + 1] if s]


def changed(a, b):
    return [ln for ln in difflib.unified_diff(a, b, lineterm="", n=0) if ln[:1] in "+-" and ln[:3] not in ("+++", "---")]


def test_live_variants_carry_the_same_block():
    base = block("BscanFFT")
    assert len(base) > 70
    for name in ("BscanFFTspin", "BscanFFTspinj", "BscanFFTwebcam"):  # Spinnaker live, Spinnaker triggered volume (C4), webcam
        assert changed(base, block(name)) == [], name
    # BscanFFTpeak: one more display call, nothing else
    d = changed(base, block("BscanFFTpeak"))
    assert d and all(ln.startswith("+") for ln in d) and d[0].startswith("+printPeakHoldAscan(")
    # BscanFFTspinjnt: the output re-binning block (compiled separately: oracle/_ref spinjnt_rebin) and the 30 dB clamp value
    d = changed(base, block("BscanFFTspinjnt"))
    assert "-bscandisp.at<double>(5,5)=50.0;" in d and "+bscandisp.at<double>(5,5)=30.0;" in d
    rest = [ln for ln in d if "at<double>(5,5)" not in ln]
    assert all(ln.startswith("+") for ln in rest)
    assert "+resize(bscan,bscanbinned,Size(),1.0/bscanbinx,1.0/bscanbiny,INTER_AREA);" in rest
    assert "+resize(multiplyfactor*bscanbinned,bscan,Size(),bscanbinx*binvaluey,bscanbiny,INTER_CUBIC);" in rest
    assert len(rest) == 5  # if (...) { resize; resize; }


def test_dark_variant_differs_where_the_dark_module_is_compiled_for():
    d = changed(block("BscanFFT"), block("BscanDark"))
    assert "+data_y=data_y-data_yd;" in d  # BscanDark.cpp:1269
    assert "+data_y=zeropadrowwise(data_y,increasefftpointsmultiplier,bandpassfilter);" in d
    arithmetic = [ln for ln in d if re.search(r"dft\(|magnitude\(|log\(|normalize\(|2\.303|accumulate\(|transpose\(|fractionalk|slopes", ln)]
    assert arithmetic == [], arithmetic  # resampling, DFT, magnitude, averaging, dB and the display normalisation are the same lines


def test_sim_variant_has_the_same_arithmetic_lines():
    """BscanFFTsim.cpp (BASELINE configs[0]) is an older copy: no normalise switches, no averaging buffer toggling - the arithmetic
    statements of the block are the same ones."""
    sim = set(block("BscanFFTsim"))
    for ln in ("data_y=(data_y-data_yp)/data_yb;", "Scalarmeanval=mean(data_y.row(p));", "data_y.row(p)=data_y.row(p)-meanval(0);",
               "multiply(data_y.row(p),barthannwin,data_y.row(p));", "slopes.at<double>(p,q)=data_y.at<double>(p,q)-data_y.at<double>(p,q-1);",
               "slopes.at<double>(p,0)=slopes.at<double>(p,1);", "dft(complexI,complexI,DFT_ROWS|DFT_INVERSE);",
               "magnitude(planes[0],planes[1],magI);", "log(bscan,bscanlog);", "bscandb=20.0*bscanlog/2.303;",
               "bscandb.row(4).copyTo(bscandb.row(1));", "bscandb.row(4).copyTo(bscandb.row(0));",
               "normalize(bscandisp,bscandisp,0,1,NORM_MINMAX);"):
        assert ln in sim, ln
