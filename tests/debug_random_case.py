"""Debug helper (not a test): where does the CUDA path deviate most from the oracle for one seed of the randomised sweep?
    python tests/debug_random_case.py SEED [SEED ...]"""
import sys

import numpy as np

import test_parity_gpu as t
from util import MAG_FLOOR, abi_params, db_to_mag, oracle_params


def main():
    from fdoct_b200 import api, synth
    from oracle.abcoct_oracle import Oracle, dark_background

    for seed in [int(a) for a in sys.argv[1:]]:
        c = t._random_config(4200 + seed)
        w, h, A, variant, N, D = c["w"], c["h"], c["A"], c["variant"], c["N"], c["D"]
        op = oracle_params(w=w, h=h, numfftpoints=N, numdisplaypoints=D, averages=A, variant=variant, lambdamin=840.5e-9,
                           lambdamax=859.5e-9, **c["extra"])
        frames = synth.make_frames(c["nB"] * A, w, h, seed=seed, dark=variant == 1)
        o = Oracle(op)
        yd = None
        if variant == 1:
            yd = o.calib_capture(synth.make_dark_frames(2, w, h, seed=seed + 2))
            yr = o.calib_capture(synth.make_background_frames(2, w, h, seed=seed + 1, dark=True))
            yb = dark_background(yr, yd, yd + 0.02 * (yr - yd))
            o.set_dark(yd)
        else:
            yb = o.calib_capture(synth.make_background_frames(2, w, h, seed=seed + 1))
        o.set_background(yb)
        ref8, refdb = o.process_bscans(frames)
        with api.Context(abi_params(op)) as ctx:
            ctx.set_background(yb)
            if yd is not None:
                ctx.set_dark(yd)
            out8, outdb = ctx.process_bscans(frames, want_db=True)
            ylin_dev = ctx.debug_linearised(frames[0])
        ylin_ref = o.linearised(frames[0])
        g, r = db_to_mag(outdb), db_to_mag(refdb)
        colmax = np.abs(r).max(axis=-2, keepdims=True)
        den = np.maximum(np.abs(r), MAG_FLOOR * colmax)
        e = np.abs(g - r) / den
        b, d, a = np.unravel_index(np.argmax(e), e.shape)
        print(f"seed {seed}: {c}")
        print(f"  worst {e.max():.3g} at bscan {b} bin {d} A-scan {a}: ours {g[b, d, a]:.6g} ref {r[b, d, a]:.6g} colmax {colmax[b, 0, a]:.6g} "
              f"partner colmax {colmax[b, 0, a ^ 1] if (a ^ 1) < colmax.shape[-1] else float('nan'):.6g}")
        print("  error by bin (max over A-scans), first 12:", np.array2string(e[b].max(axis=1)[:12], precision=2))
        print("  bins with error > 1e-4:", np.argwhere(e[b].max(axis=1) > 1e-4).ravel()[:30])
        dy = np.abs(ylin_dev - ylin_ref)
        print(f"  ylin: max |ref| {np.abs(ylin_ref).max():.4g}, median |ref| {np.median(np.abs(ylin_ref)):.4g}, max abs diff {dy.max():.3g} "
              f"(rel to max {dy.max() / np.abs(ylin_ref).max():.3g}), row means of diff {np.abs((ylin_dev - ylin_ref).mean(axis=1)).max():.3g}")
        print(f"  yb min {np.abs(yb).min():.3g} max {np.abs(yb).max():.3g}")


if __name__ == "__main__":
    main()
