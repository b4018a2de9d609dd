"""End-to-end throughput of abcoct_process_bscans with PAGEABLE caller buffers (plain numpy arrays, what the offline tool and a
naive caller pass) against pinned ones.  Run on a B200:  python tools/bench_pageable.py [frames]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fdoct_b200 import api, synth  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    w, h, N, D = 2048, 1024, 2048, 1024
    p = api.default_params(w=w, h=h, bpp=16, binx=1, biny=1, averages=1, numfftpoints=N, numdisplaypoints=D, lambdamin=840.5e-9,
                           lambdamax=859.5e-9, mediann=0, movavgn=0, fft_multiplier=1, donotnormalize=1)
    uniq = synth.make_frames(8, w, h, seed=5)
    frames = np.ascontiguousarray(uniq[np.arange(n) % 8])
    yb = synth.make_background_frames(2, w, h, seed=6).mean(axis=0)
    out = np.empty((n, D, h), np.uint8)
    pin_in, pin_out = api.PinnedArray(frames.shape, np.uint16), api.PinnedArray(out.shape, np.uint8)
    pin_in.array[...] = frames
    with api.Context(p) as ctx:
        ctx.set_background(yb)
        for name, fi, fo in (("pageable", frames, out), ("pinned", pin_in.array, pin_out.array)):
            for threads in (("1", "default") if name == "pageable" else ("default",)):
                if threads == "1":
                    os.environ["ABCOCT_COPY_THREADS"] = "1"
                else:
                    os.environ.pop("ABCOCT_COPY_THREADS", None)
                ctx.process_bscans(fi, out8=fo)
                ts = []
                for _ in range(3):
                    t0 = time.perf_counter()
                    ctx.process_bscans(fi, out8=fo)
                    ts.append(time.perf_counter() - t0)
                t = min(ts)
                print(f"{name:9s} copy threads {threads:8s} {n * h / t:.3e} A-scans/s  {frames.nbytes / t / 1e9:6.1f} GB/s in")
        assert np.array_equal(out, pin_out.array)
    pin_in.free()
    pin_out.free()


if __name__ == "__main__":
    main()
