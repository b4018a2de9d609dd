"""Tiny end-to-end cases for compute-sanitizer (one per plan family): run under
    compute-sanitizer --tool racecheck|memcheck python tools/sanitize_case.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fdoct_b200 import api, synth

for (w, h, N, D, A, variant, extra) in [(2048, 9, 2048, 1024, 1, 0, {}), (1280, 6, 1280, 640, 2, 1, {}), (1024, 5, 1024, 300, 1, 0, {}),
                                        (640, 4, 1280, 512, 1, 0, dict(fft_multiplier=2))]:
    frames = synth.make_frames(2 * A, w, h, seed=3, dark=bool(variant))
    yb = synth.make_background_frames(2, w, h, seed=4, dark=bool(variant)).mean(axis=0)
    p = api.default_params(w=w, h=h, bpp=16, averages=A, numfftpoints=N, numdisplaypoints=D, lambdamin=840.5e-9, lambdamax=859.5e-9,
                           mediann=0, variant=variant, **extra)
    with api.Context(p) as ctx:
        ctx.set_background(yb)
        if variant:
            ctx.set_dark(synth.make_dark_frames(2, w, h, seed=5).mean(axis=0))
        out8, outdb = ctx.process_bscans(frames, want_db=True)
    print(w, N, A, variant, out8.shape, int(out8.sum()), float(outdb.mean()))
