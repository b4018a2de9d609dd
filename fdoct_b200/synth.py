"""Deterministic synthetic spectrometer-camera frames.

Recipe after the reference's own generator ``Matlab files/wangOCTimg2.m:12-63``:
every camera row is one spectrum ``I(lambda) = S(lambda) |1 + sum_j rho_j exp(i 4 pi n z_j / lambda)|^2``
with a Gaussian source (850 nm, 20 nm FWHM), lambda linear across the row, two reflectors whose depth
varies with the row, scaled to 60 % of full scale, plus Gaussian read noise, rounded to the integer
pixel type.  Nothing here touches the GPU or the oracle; tests, bench.py and smoke() all draw inputs
from this one place so every arm sees the same frames.
"""
from __future__ import annotations

import numpy as np

LAMBDA0 = 850e-9
FWHM = 20e-9
SIGMA_LAMBDA = FWHM / np.sqrt(2 * np.log(2))  # wangOCTimg2.m:22


def _lambdas(w: int, lmin: float, lmax: float) -> np.ndarray:
    return np.linspace(lmin, lmax, w)


def source_spectrum(w: int, lmin: float = 840.5e-9, lmax: float = 859.5e-9) -> np.ndarray:
    lam = _lambdas(w, lmin, lmax)
    return np.exp(-0.5 * (lam - LAMBDA0) ** 2 / SIGMA_LAMBDA**2)


def clean_frame(w: int, h: int, lmin: float = 840.5e-9, lmax: float = 859.5e-9, phase: float = 0.0,
                rho=(0.5, 0.25), depth0: float = 0.4e-3, depth_span: float = 3.0e-3, gap: float = 1.2e-3) -> np.ndarray:
    """Noise-free interferogram in [0, 1] (f64, h x w); `phase` shifts the depths slightly per frame."""
    lam = _lambdas(w, lmin, lmax)[None, :]
    s = source_spectrum(w, lmin, lmax)[None, :]
    r = np.arange(h, dtype=np.float64)[:, None] / max(h - 1, 1)
    z1 = depth0 + depth_span * r + 2e-6 * phase
    z2 = z1 + gap
    e = rho[0] * np.exp(1j * 4 * np.pi * z1 / lam) + rho[1] * np.exp(1j * 4 * np.pi * z2 / lam)
    i = s * np.abs(1.0 + e) ** 2
    return i / i.max()


def make_frames(nframes: int, w: int, h: int, *, seed: int, full_scale: int = 65535, dark: bool = False,
                lmin: float = 840.5e-9, lmax: float = 859.5e-9, n_unique: int | None = None,
                dtype=np.uint16) -> np.ndarray:
    """(nframes, h, w) integer frames. `dark=True` adds the N(64, 4)-count dark pedestal (BscanDark shape).

    `n_unique` bounds the number of distinct noise-free interferograms that are synthesised (the rest reuse
    them with fresh noise) so that large benchmark batches are cheap to build.
    """
    rng = np.random.default_rng(seed)
    nu = nframes if n_unique is None else max(1, min(nframes, n_unique))
    base = [0.6 * full_scale * clean_frame(w, h, lmin, lmax, phase=float(j)) for j in range(nu)]
    out = np.empty((nframes, h, w), dtype=dtype)
    sigma = 0.005 * full_scale
    for f in range(nframes):
        x = base[f % nu] + rng.normal(0.0, sigma, size=(h, w))
        if dark:
            x = x + rng.normal(64.0, 4.0, size=(h, w))
        out[f] = np.clip(np.rint(x), 0, np.iinfo(dtype).max).astype(dtype)
    return out


def make_background_frames(averages: int, w: int, h: int, *, seed: int, full_scale: int = 65535, dark: bool = False,
                           lmin: float = 840.5e-9, lmax: float = 859.5e-9, scale: float = 0.6 / 2.25,
                           dtype=np.uint16) -> np.ndarray:
    """Frames of the source spectrum alone (reference arm only), same noise model, seed = config seed + 1.

    `scale`: the interferogram peaks near |1 + 0.5 + 0.25|^2 = 3.06 x S and is normalised to 0.6 FS, so the
    reference-arm-only level is about 0.6/ (1 + 0.5^2 + 0.25^2 + cross terms); an exact ratio is not needed
    because the block divides by the background and then removes the row mean.
    """
    rng = np.random.default_rng(seed)
    s = source_spectrum(w, lmin, lmax)[None, :] * np.ones((h, 1))
    out = np.empty((averages, h, w), dtype=dtype)
    sigma = 0.005 * full_scale
    for f in range(averages):
        x = scale * full_scale * s + 0.02 * full_scale + rng.normal(0.0, sigma, size=(h, w))
        if dark:
            x = x + rng.normal(64.0, 4.0, size=(h, w))
        out[f] = np.clip(np.rint(x), 1, np.iinfo(dtype).max).astype(dtype)
    return out


def make_dark_frames(averages: int, w: int, h: int, *, seed: int, dtype=np.uint16) -> np.ndarray:
    """Dark frames: N(64, 4) counts, seed = config seed + 2."""
    rng = np.random.default_rng(seed)
    x = rng.normal(64.0, 4.0, size=(averages, h, w))
    return np.clip(np.rint(x), 0, np.iinfo(dtype).max).astype(dtype)
