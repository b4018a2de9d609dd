"""Cost of the B-scan consumers (post_kernels.cu) on top of the fused kernel: device-resident C1-shaped batch, CUDA events on the
launching stream.  Prints A-scans/s for: display only; + dB; + linear; + JET; + J0 lock-in display (+ its JET).  Run on a B200:
    python tools/bench_consumers.py > gpurun_out/consumers.txt"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fdoct_b200 import api, synth  # noqa: E402


def main():
    w, h, N, D, nB = 1280, 960, 1280, 640, 512
    p = api.default_params(w=w, h=h, bpp=16, binx=1, biny=1, averages=1, numfftpoints=N, numdisplaypoints=D, lambdamin=840.5e-9,
                           lambdamax=859.5e-9, mediann=0, movavgn=0, fft_multiplier=1, donotnormalize=1)
    frames = synth.make_frames(nB, w, h, seed=5, n_unique=8)
    yb = synth.make_background_frames(2, w, h, seed=6).mean(axis=0)
    dev = torch.device("cuda:0")
    d_in = torch.from_numpy(frames.view(np.int16)).to(dev)
    bufs = {k: torch.empty((nB, D, h) + tail, dtype=torch.uint8 if dt is np.uint8 else torch.float32, device=dev)
            for k, (dt, tail) in api.OUTPUT_KINDS.items()}
    st = torch.cuda.Stream(device=dev)
    sets = [("display only", ["bscan_u8"]), ("+ dB", ["bscan_u8", "bscan_db"]), ("+ dB + linear", ["bscan_u8", "bscan_db", "bscan_lin"]),
            ("+ JET", ["bscan_u8", "bscan_bgr"]), ("+ J0 display", ["bscan_u8", "jsub_u8"]),
            ("everything", list(api.OUTPUT_KINDS))]
    with api.Context(p) as ctx:
        ctx.set_background(yb)
        o = {"bscan_u8": bufs["bscan_u8"].data_ptr(), "bscan_lin": bufs["bscan_lin"].data_ptr()}
        ctx.process_bscans_device_ex(d_in.data_ptr(), 1, o)
        torch.cuda.synchronize()
        ctx.set_jscan(bufs["bscan_lin"][0].cpu().numpy())
        base = None
        for name, keys in sets:
            o = {k: bufs[k].data_ptr() for k in keys}
            ms = []
            for it in range(8):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                ctx.process_bscans_device_ex(d_in.data_ptr(), nB, o, stream=st.cuda_stream)
                e1.record(st)
                e1.synchronize()
                if it >= 3:
                    ms.append(e0.elapsed_time(e1))
            t = float(np.median(ms))
            base = base or t
            out_bytes = sum(bufs[k].numel() * bufs[k].element_size() for k in keys)
            print(f"{name:16s} {t:8.3f} ms / {nB} B-scans  {nB * h / t * 1e3:.3e} A-scans/s  +{t - base:6.3f} ms  outputs {out_bytes / 1e6:7.1f} MB"
                  f"  ({out_bytes / max(t - base, 1e-9) / 1e6:8.0f} GB/s of extra output)" if t > base else
                  f"{name:16s} {t:8.3f} ms / {nB} B-scans  {nB * h / t * 1e3:.3e} A-scans/s  outputs {out_bytes / 1e6:7.1f} MB")


if __name__ == "__main__":
    main()
