"""Host-side execution of the kernel's own __host__ __device__ code (no GPU): the in-register DFT templates against a
naive f64 DFT, and a thread-by-thread emulation of one recon_kernel thread group for the compiled plans."""
import subprocess

from util import ROOT  # noqa: F401

from fdoct_b200 import build


def _run(name, *args):
    exe = build.build_native_test(name)
    r = subprocess.run([exe, *args], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


def test_register_dft_radices():
    out = _run("test_dft_host")
    assert "worst=" in out


def test_thread_group_emulation_small_plans():
    out = _run("test_group_host", "quick")
    assert out.count("max|dB err|") >= 7


def test_warp_per_ascan_kernel_lockstep_emulation():
    """The WHOLE warp-per-A-scan kernel body (wrow_kernel.cuh: ticket scheduler, both FFT passes, the lane pairing of the split
    step, DC-row / clampupper special cases, worker -> service-warp mailboxes, completion protocol, normalisation jobs, all
    three load modes) executed by 32 host threads per warp against an
    f64 restatement: magnitude within 1e-4 of max(|ref|, 1e-3 A-scan max), display within 1 LSB."""
    out = _run("test_wrow_host")
    assert out.count("max rel mag err") >= 8 and "worst (in units of the tolerance)" in out
