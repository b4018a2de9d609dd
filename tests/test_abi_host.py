"""CPU tests of the boundary: the C-ABI library loads and exports every symbol include/abcoct.h declares, the
host-only entry points (tables, .ini parser, validation) behave like the reference's code, and compute entry points
fail loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from util import GOLDEN, ROOT, abi_params, oracle_params

from fdoct_b200 import api
from oracle.abcoct_oracle import barthann_window, build_tables


def _declared():
    src = open(os.path.join(ROOT, "include", "abcoct.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(abcoct_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = api.lib()
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"libabcoct.so lacks {n}"
    assert sorted(api.EXPORTS) == names


def test_struct_layout_matches_header(tmp_path):
    """The header is plain C (gcc -std=c99 compiles it) and the ctypes mirror has the same size and offsets."""
    import subprocess

    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "abcoct.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(abcoct_params),offsetof(abcoct_params,lambdamin),offsetof(abcoct_params,bscanthreshold),'
                   'offsetof(abcoct_params,clamp_db),sizeof(abcoct_info),offsetof(abcoct_info,kernel_launches),'
                   'sizeof(abcoct_outputs),offsetof(abcoct_outputs,jsub_u8),offsetof(abcoct_outputs,reserved));return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert got == [C.sizeof(api.Params), api.Params.lambdamin.offset, api.Params.bscanthreshold.offset, api.Params.clamp_db.offset,
                   C.sizeof(api.Info), api.Info.kernel_launches.offset,
                   C.sizeof(api.Outputs), api.Outputs.jsub_u8.offset, api.Outputs.reserved.offset]
    p = api.default_params()
    assert (p.w, p.h, p.bpp, p.numfftpoints, p.numdisplaypoints, p.mediann) == (640, 480, 8, 1024, 512, 5)
    assert p.bscanthreshold == -30.0 and p.clamp_db == 50.0 and p.donotnormalize == 1


@pytest.mark.parametrize("w,binx,m,N", [(1280, 1, 1, 1280), (1280, 1, 1, 2048), (1280, 2, 4, 2560), (1920, 1, 2, 3840),
                                        (1440, 2, 4, 2880), (2048, 1, 1, 2048), (4096, 1, 1, 4096), (128, 1, 1, 128)])
def test_tables_bit_exact(w, binx, m, N):
    """nearestkindex int32-equal and fractionalk / window f64-equal to the oracle (BscanFFT.cpp:615-698, 936-944)."""
    lmin, lmax = (816e-9, 884e-9) if w == 128 else (840.5e-9, 859.5e-9)
    op = oracle_params(w=w, h=8, binx=binx, biny=binx, numfftpoints=N, fft_multiplier=m, lambdamin=lmin, lambdamax=lmax)
    nk, fr, win = api.build_tables(abi_params(op))
    t = build_tables(op)
    assert nk.dtype == np.int32 and np.array_equal(nk, t["nearestkindex"])
    assert np.array_equal(fr, t["fractionalk"])
    assert np.array_equal(win, barthann_window(op.opw))


INI = [("bscanfft", api.INI_BSCANFFT), ("spinj", api.INI_SPINJ), ("spinjnt", api.INI_SPINJNT), ("dark", api.INI_DARK),
       ("peak", api.INI_PEAK), ("webcam", api.INI_WEBCAM), ("sim", api.INI_SIM)]


@pytest.mark.parametrize("name,flavour", INI)
def test_ini_positional_parser(name, flavour):
    p = api.params_from_ini(os.path.join(GOLDEN, name + ".ini"), flavour)
    assert (p.bpp, p.w, p.h) == (16, 1280, 960)
    assert (p.averages, p.numfftpoints, p.numdisplaypoints, p.movavgn, p.mediann, p.fft_multiplier) == (8, 1280, 640, 0, 0, 1)
    assert p.lambdamin == 840.5e-9 and p.lambdamax == 859.5e-9
    if flavour == api.INI_SPINJNT:
        assert (p.binx, p.biny) == (2, 1) and p.clamp_db == 30.0
    else:
        assert (p.binx, p.biny) == (1, 1) and p.clamp_db == 50.0
    if flavour != api.INI_SIM:
        assert (p.rowwisenormalize, p.donotnormalize) == (0, 1)
    assert p.variant == (1 if flavour == api.INI_DARK else 0)
    if flavour == api.INI_DARK:
        assert p.bandpassfilter == 1 and p.lowpassfilter == 0
    assert p.channelnum == (2 if flavour == api.INI_WEBCAM else 0)


# The reference's own shipped configuration files (build/*.ini, copied verbatim as fixtures): expected values read off the files.
REF_INI = {
    #                  flavour          bpp  w     h    binx biny avg N     D    movavg median mult rown nonorm
    "BscanFFT":        (api.INI_BSCANFFT, 8, 320, 240, 2, 2, 10, 2560, 320, 0, 0, 4, 0, 1),
    "BscanFFTspin":    (api.INI_BSCANFFT, 8, 1280, 960, 2, 2, 10, 2560, 320, 0, 0, 4, 0, 1),
    "BscanFFTspinj":   (api.INI_SPINJ, 16, 720, 480, 1, 1, 10, 2880, 360, 0, 0, 4, 0, 1),
    "BscanFFTspinjnt": (api.INI_SPINJNT, 16, 720, 480, 1, 1, 10, 2880, 360, 0, 0, 4, 0, 1),
    "BscanDark":       (api.INI_DARK, 16, 1280, 960, 2, 2, 10, 2560, 320, 0, 0, 4, 0, 1),
    "BscanFFTpeak":    (api.INI_PEAK, 8, 1280, 960, 2, 2, 10, 2560, 320, 0, 0, 4, 0, 1),
    "BscanFFTwebcam":  (api.INI_WEBCAM, 8, 640, 480, 1, 1, 10, 640, 320, 0, 0, 1, 0, 1),
}


@pytest.mark.parametrize("name", sorted(REF_INI))
def test_ini_parser_on_the_reference_shipped_files(name):
    """abcoct_params_from_ini on the seven .ini files the reference ships (positional layouts of BscanFFT.cpp:395-484 and its
    variants): every field the reconstruction block reads."""
    exp = REF_INI[name]
    p = api.params_from_ini(os.path.join(GOLDEN, "ref_ini", name + ".ini"), exp[0])
    got = (p.bpp, p.w, p.h, p.binx, p.biny, p.averages, p.numfftpoints, p.numdisplaypoints, p.movavgn, p.mediann, p.fft_multiplier,
           p.rowwisenormalize, p.donotnormalize)
    assert got == exp[1:], (name, got)
    assert p.lambdamin == 840.5e-9 and p.lambdamax == 859.5e-9
    assert p.variant == (1 if exp[0] == api.INI_DARK else 0)
    if exp[0] == api.INI_SPINJNT:
        assert (p.bscanbinx, p.bscanbiny, p.output_rebin, p.clamp_db) == (1, 1, 1, 30.0)
    if exp[0] == api.INI_DARK:
        assert p.bandpassfilter == 1  # the shipped file stops there: lowpassfilter keeps its default
        assert p.lowpassfilter == 0
    if exp[0] == api.INI_WEBCAM:
        assert p.channelnum == 3  # BscanFFTwebcam.cpp:508, the last field: the shipped file asks for the SUM of the three channels


def test_ini_wrong_flavour_misparses_like_the_reference():
    """BscanFFTsim reads BscanFFT.ini, whose offsets shift every later field by two."""
    p = api.params_from_ini(os.path.join(GOLDEN, "bscanfft.ini"), api.INI_SIM)
    assert p.numfftpoints != 1280


def test_ini_missing_file_keeps_defaults():
    p = api.Params()
    rc = api.lib().abcoct_params_from_ini(b"/nonexistent/x.ini", api.INI_BSCANFFT, C.byref(p))
    assert rc == api.ERR_IO and p.numfftpoints == 1024 and p.w == 640


def _create_rc(**kw):
    op = oracle_params(**kw)
    h = C.c_void_p()
    rc = api.lib().abcoct_create(C.byref(abi_params(op)), None, 1, C.byref(h))
    msg = (api.lib().abcoct_last_error(None) or b"").decode()
    if rc == 0:
        api.lib().abcoct_destroy(h)
    return rc, msg


def test_create_rejects_reference_undefined_behaviour():
    rc, msg = _create_rc(w=1280, h=8, numfftpoints=1024, numdisplaypoints=256)  # N < M
    assert rc == api.ERR_INVALID and "1170" in msg
    rc, _ = _create_rc(w=1280, h=8, numfftpoints=1280, numdisplaypoints=4)
    assert rc == api.ERR_INVALID
    rc, _ = _create_rc(w=1280, h=8, numfftpoints=1280, numdisplaypoints=1281)  # D > N: colRange throws (BscanFFT.cpp:1193)
    assert rc == api.ERR_INVALID
    rc, _ = _create_rc(w=1280, h=8, numfftpoints=1280, numdisplaypoints=640, lambdamin=860e-9, lambdamax=840e-9)
    assert rc == api.ERR_INVALID


def test_create_accepts_what_cv_dft_accepts():
    """cv::dft takes any length and colRange any D <= N (BscanFFT.cpp:1185, 1193): transform lengths without a fused plan, rows
    that are not a multiple of 8 samples and display rows above N / 2 pass validation (generic kernel); ERR_CUDA = no GPU here."""
    for kw in (dict(w=1280, h=8, numfftpoints=1280, numdisplaypoints=1000),  # D > N/2
               dict(w=1280, h=8, numfftpoints=1280, numdisplaypoints=1280),  # D = N
               dict(w=1000, h=8, numfftpoints=2000, numdisplaypoints=512),  # 2^4 5^3
               dict(w=1284, h=8, numfftpoints=3072, numdisplaypoints=512),  # row width 4 mod 8, 2^10 3
               dict(w=750, h=8, numfftpoints=3000, numdisplaypoints=100, fft_multiplier=2)):
        rc, msg = _create_rc(**kw)
        assert rc in (api.OK, api.ERR_CUDA), (kw, rc, msg)


def test_create_reports_unsupported_not_silently_wrong():
    rc, msg = _create_rc(w=1280, h=8, numfftpoints=1344, numdisplaypoints=512)  # 1344 = 2^6 3 7: neither a plan nor 2^a 3^b 5^c
    assert rc == api.ERR_UNSUPPORTED and "1280" in msg
    rc, msg = _create_rc(w=1280, h=8, numfftpoints=1280, numdisplaypoints=512, mediann=7)  # OpenCV: 3 / 5 only for 16-bit
    assert rc == api.ERR_UNSUPPORTED and "median" in msg
    rc, msg = _create_rc(w=1288, h=8, numfftpoints=2048, numdisplaypoints=512, binx=2, biny=2, fft_multiplier=2)  # 644 = 4 * 7 * 23
    assert rc == api.ERR_UNSUPPORTED
    rc, msg = _create_rc(w=1281, h=8, numfftpoints=1280, numdisplaypoints=512, binx=2, biny=2)  # cv::resize would round the size
    assert rc == api.ERR_INVALID
    # BscanFFTspinjnt re-bins the linear B-scan at the output whenever any binning factor exceeds 1 (BscanFFTspinjnt.cpp:1856-1862).
    # The shipped shape (binvaluex = 2, everything else 1) reduces to bscan *= multiplyfactor and is built; a real resampling step
    # (bscanbinx / bscanbiny / binvaluey > 1) is refused rather than silently reconstructed without it
    p = api.params_from_ini(os.path.join(GOLDEN, "spinjnt.ini"), api.INI_SPINJNT)
    assert (p.output_rebin, p.bscanbinx, p.bscanbiny, p.binx, p.biny) == (1, 1, 1, 2, 1)
    h = C.c_void_p()
    rc = api.lib().abcoct_create(C.byref(p), None, 1, C.byref(h))
    assert rc in (api.OK, api.ERR_CUDA)  # ERR_CUDA: no GPU on the CPU test box - validation passed
    if rc == api.OK:
        api.lib().abcoct_destroy(h)
    p.bscanbinx = 2
    assert api.lib().abcoct_create(C.byref(p), None, 1, C.byref(h)) == api.ERR_UNSUPPORTED
    assert "1856" in api.lib().abcoct_last_error(None).decode()
    p.bscanbinx = 1
    p.biny = 2
    p.h = 2 * (p.h // 2)
    assert api.lib().abcoct_create(C.byref(p), None, 1, C.byref(h)) == api.ERR_UNSUPPORTED
    p.biny = 1
    p.binx = 1  # no binning anywhere: the re-binning block is skipped by the reference too
    rc = api.lib().abcoct_create(C.byref(p), None, 1, C.byref(h))
    assert rc in (api.OK, api.ERR_CUDA)  # ERR_CUDA: no GPU on the CPU test box - validation passed
    if rc == api.OK:
        api.lib().abcoct_destroy(h)


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    rc, msg = _create_rc(w=1280, h=8, numfftpoints=1280, numdisplaypoints=640)
    assert rc == api.ERR_CUDA and "no CPU fallback" in msg


def test_product_package_does_not_import_the_oracle():
    pk = os.path.join(ROOT, "fdoct_b200")
    for dp, _, fs in os.walk(pk):
        for f in fs:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|#include.*oracle|oracle[./]abcoct_oracle", txt, flags=re.M), f
