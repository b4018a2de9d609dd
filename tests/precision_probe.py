"""GPU-side probe (test infrastructure: it uses the oracle, so it lives under tests/): deviation of the CUDA path and of the reference's OpenCV f32 path from an exact f64 evaluation of
the same block, under the parity metric |a-b| / max(|b|, floor * A-scan max).  Run under gpurun."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root
from fdoct_b200 import api, synth
from oracle.abcoct_oracle import Oracle, Params

K = 20.0 / 2.303


def metric(g, r, floor):
    den = np.maximum(np.abs(r), floor * np.abs(r).max(axis=-2, keepdims=True))
    return float((np.abs(g - r) / den).max())


for (w, h, N, D) in [(1024, 256, 1024, 512), (2048, 256, 2048, 1024), (4096, 128, 4096, 2048), (1280, 256, 1280, 640), (1920, 128, 1920, 960)]:
    op = Params(w=w, h=h, numfftpoints=N, numdisplaypoints=D, lambdamin=840.5e-9, lambdamax=859.5e-9)
    fr = synth.make_frames(2, w, h, seed=1005)
    yb = synth.make_background_frames(2, w, h, seed=1006).mean(axis=0)
    o = Oracle(op)
    o.set_background(yb)
    ref8, refdb = o.process_bscans(fr)
    ref32 = np.exp(refdb / K) - 1e-5
    exact = np.stack([np.abs(np.fft.ifft(o.linearised(f), axis=1) * N)[:, :D].T for f in fr])
    exact[:, 0] = exact[:, 4]; exact[:, 1] = exact[:, 4]
    p = api.default_params(w=w, h=h, bpp=16, numfftpoints=N, numdisplaypoints=D, lambdamin=840.5e-9, lambdamax=859.5e-9, mediann=0)
    with api.Context(p) as ctx:
        ctx.set_background(yb)
        out8, outdb = ctx.process_bscans(fr, want_db=True)
    ours = np.exp(outdb.astype(np.float64) / K) - 1e-5
    for fl in (1e-3, 3e-3):
        print(f"N={N} floor={fl:g}: ours-vs-opencv {metric(ours, ref32, fl):.2e}  ours-vs-exact {metric(ours, exact, fl):.2e}  opencv-vs-exact {metric(ref32, exact, fl):.2e}"
              f"  u8 max diff {np.abs(out8.astype(int) - ref8.astype(int)).max()} frac {(out8 != ref8).mean():.2e}")
