"""GPU-side probe (test infrastructure, not collected by pytest): wall-clock latency of ONE live call of
abcoct_process_bscans - the way the reference's camera loop would use it - for the camera shapes of BASELINE configs
C1 / C2 / C3, pinned and pageable caller buffers.  Run under gpurun."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root
from fdoct_b200 import api, synth

for name, w, h, N, D, A, extra in [("C1 1280x960 A=1", 1280, 960, 1280, 640, 1, {}), ("C2 1280x960 A=8 DARK", 1280, 960, 1280, 640, 8, dict(variant=1)),
                                   ("C3 1920x1200 m=2", 1920, 1200, 3840, 1024, 1, dict(fft_multiplier=2)), ("C4 1920x1200", 1920, 1200, 1920, 960, 1, {})]:
    p = api.default_params(w=w, h=h, bpp=16, averages=A, numfftpoints=N, numdisplaypoints=D, lambdamin=840.5e-9, lambdamax=859.5e-9, mediann=0, **extra)
    frames = synth.make_frames(A, w, h, seed=1, n_unique=2)
    yb = synth.make_background_frames(2, w, h, seed=2).mean(axis=0)
    with api.Context(p) as ctx:
        ctx.set_background(yb)
        if extra.get("variant"):
            ctx.set_dark(synth.make_dark_frames(2, w, h, seed=3).mean(axis=0))
        pin = api.PinnedArray(frames.shape, np.uint16)
        pin.array[...] = frames
        pout = api.PinnedArray((1, D, h), np.uint8)
        res = {}
        for label, src, dst in (("pinned", pin.array, pout.array), ("pageable", frames, None)):
            for _ in range(5):
                ctx.process_bscans(src, out8=dst)
            ts = []
            for _ in range(50):
                t0 = time.perf_counter()
                ctx.process_bscans(src, out8=dst)
                ts.append(time.perf_counter() - t0)
            res[label] = np.median(ts) * 1e3
        print(f"{name:24s} one B-scan per call: {res['pinned']:.3f} ms pinned, {res['pageable']:.3f} ms pageable  ->  {1e3 / res['pinned']:.0f} B-scans/s, "
              f"{A * h * 1e3 / res['pinned']:.3e} A-scans/s live")
        pin.free(); pout.free()
