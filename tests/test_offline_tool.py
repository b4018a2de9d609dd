"""The offline volume tool (the Bscancompute.bin slot of BscanFFTspinj) and the .ocv matrix format."""
import os

import numpy as np
import pytest

from util import GOLDEN, ROOT  # noqa: F401

from fdoct_b200 import offline
from fdoct_b200.ocv import read_ocv, write_ocv


def test_ocv_round_trip_and_header(tmp_path):
    rng = np.random.default_rng(1)
    for arr in (rng.normal(size=(7, 5)), rng.integers(0, 65535, size=(4, 9), dtype=np.uint16), rng.normal(size=(3, 4)).astype(np.float32),
                rng.integers(0, 255, size=(6, 2, 3), dtype=np.uint8)):
        p = str(tmp_path / "m.ocv")
        write_ocv(p, arr)
        assert np.array_equal(read_ocv(p), arr)
    write_ocv(str(tmp_path / "d.ocv"), np.zeros((2, 3)))
    hdr = np.fromfile(str(tmp_path / "d.ocv"), dtype=np.int32, count=4)
    assert hdr.tolist() == [2, 3, 6, 1]  # rows, cols, CV_64F, channels (matwrite, BscanFFTspinj.cpp:677-683)
    assert os.path.getsize(str(tmp_path / "d.ocv")) == 16 + 2 * 3 * 8
    with open(str(tmp_path / "t.ocv"), "wb") as f:
        f.write(b"\x01\x00\x00\x00")
    with pytest.raises(ValueError):
        read_ocv(str(tmp_path / "t.ocv"))


def test_capture_listing(tmp_path):
    for n in (1, 2, 10):
        for i in range(3):
            (tmp_path / f"Trig{n:03d}-{i:03d}.png").write_bytes(b"")
    (tmp_path / "KTrig001-000.png").write_bytes(b"")  # J0 frames are not signal frames
    caps = offline.list_captures(str(tmp_path), 3)
    assert list(caps) == [1, 2, 10] and [os.path.basename(p) for p in caps[2]] == ["Trig002-000.png", "Trig002-001.png", "Trig002-002.png"]
    with pytest.raises(FileNotFoundError):
        offline.list_captures(str(tmp_path), 4)


@pytest.mark.gpu
def test_offline_tool_end_to_end(tmp_path):
    """Synthetic capture directory -> tool -> bscanNNN.ocv / .png, against the oracle on the same frames."""
    import cv2

    from fdoct_b200 import api, synth
    from oracle.abcoct_oracle import Oracle
    from util import assert_display_parity, mag_rel_err, oracle_params

    d = str(tmp_path)
    p = api.params_from_ini(os.path.join(GOLDEN, "spinj.ini"), api.INI_SPINJ)  # 1280 x 960, N = 1280, D = 640
    p.h = 48
    A, nB = 3, 4
    frames = synth.make_frames(nB * A, p.w, p.h, seed=5)
    yb = synth.make_background_frames(2, p.w, p.h, seed=6).mean(axis=0)
    write_ocv(os.path.join(d, "spectrum.ocv"), yb)
    for b in range(nB):
        for i in range(A):
            assert cv2.imwrite(os.path.join(d, f"Trig{b + 1:03d}-{i:03d}.png"), frames[b * A + i])
    done = offline.run(d, A, p)
    assert done == [1, 2, 3, 4]
    op = oracle_params(w=p.w, h=p.h, numfftpoints=p.numfftpoints, numdisplaypoints=p.numdisplaypoints, averages=A,
                       lambdamin=p.lambdamin, lambdamax=p.lambdamax)
    o = Oracle(op)
    o.set_background(yb)
    ref8, refdb = o.process_bscans(frames)
    got_db = np.stack([read_ocv(os.path.join(d, f"bscan{n:03d}.ocv")) for n in done])
    got8 = np.stack([cv2.imread(os.path.join(d, f"bscan{n:03d}.png"), cv2.IMREAD_UNCHANGED) for n in done])
    assert got_db.dtype == np.float64 and got_db.shape == refdb.shape
    assert mag_rel_err(got_db, refdb) <= 1e-4
    assert_display_parity(got8, ref8, "offline tool")
    for n, g8 in zip(done, got8):  # the colour image the live program saves next to it (BscanFFTspinj.cpp:2044-2045)
        col = cv2.imread(os.path.join(d, f"bscanc{n:03d}.png"), cv2.IMREAD_UNCHANGED)
        assert col.shape == g8.shape + (3,) and np.array_equal(col, cv2.applyColorMap(g8, cv2.COLORMAP_JET))
