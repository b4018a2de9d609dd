"""NumPy model of the in-place mixed-radix DIF plan used by fdoct_b200/csrc/recon_kernel.cuh.

Checks, on the CPU, the index algebra the CUDA kernel relies on:
  pass 0 : butterfly b in [0,N1): elements at N1*a + b, twiddle w_N^(b*c), written back in place
  pass 1 : butterfly (c,b'): elements at N1*c + N2*a' + b', twiddle w_N1^(b'*c')
  pass L : butterfly k0 = c + R0*c': elements at N1*c + RL*c' + a'', output bin k0 + S*c''
and the two-for-one split (rows a,b packed as re/im) with the unit pairing (k0, S-k0).
"""
import numpy as np


def dft_mat(R, sgn):
    a = np.arange(R)
    return np.exp(sgn * 2j * np.pi * np.outer(a, a) / R)


def run_plan(x, R0, R1, RL, sgn=+1):
    N = x.size
    three = R1 > 1
    N1 = N // R0
    N2 = N1 // R1 if three else N1
    assert N2 == RL and R0 * (R1 if three else 1) * RL == N
    S = N // RL
    buf = x.astype(np.complex128).copy()
    F0 = dft_mat(R0, sgn)
    for b in range(N1):
        pos = N1 * np.arange(R0) + b
        y = F0 @ buf[pos]
        y *= np.exp(sgn * 2j * np.pi * b * np.arange(R0) / N)
        buf[pos] = y
    if three:
        F1 = dft_mat(R1, sgn)
        for c in range(R0):
            for bp in range(N2):
                pos = N1 * c + N2 * np.arange(R1) + bp
                y = F1 @ buf[pos]
                y *= np.exp(sgn * 2j * np.pi * bp * np.arange(R1) / N1)
                buf[pos] = y
    FL = dft_mat(RL, sgn)
    X = np.zeros(N, dtype=np.complex128)
    for k0 in range(S):
        c, cp = k0 % R0, k0 // R0
        pos = N1 * c + RL * cp + np.arange(RL)
        X[k0 + S * np.arange(RL)] = FL @ buf[pos]
    return X


def split_pairs(Z, R0, R1, RL):
    """Return |A[k]|, |B[k]| for k < N/2 using the kernel's unit/slot enumeration."""
    N = Z.size
    S = N // RL
    magA = np.full(N // 2, np.nan)
    magB = np.full(N // 2, np.nan)

    def emit(P, Q, kk):
        assert 0 <= kk < N // 2 and np.isnan(magA[kk]), kk
        magA[kk] = 0.5 * np.hypot(P.real + Q.real, P.imag - Q.imag)
        magB[kk] = 0.5 * np.hypot(P.imag + Q.imag, P.real - Q.real)

    JH = (RL + 1) // 2
    for u in range(S // 2):
        kA = u
        kB = S // 2 if u == 0 else S - u
        Za = Z[kA + S * np.arange(RL)]
        Zb = Z[kB + S * np.arange(RL)]
        if u == 0:
            JA = (RL + 1) // 2
            for j in range(RL):
                if j < JA:
                    if j == 0 or True:
                        emit(Za[j], Za[(RL - j) % RL], S * j)
                else:
                    jj = j - JA
                    emit(Zb[jj], Zb[RL - 1 - jj], S // 2 + S * jj)
        else:
            for j in range(RL):
                kk = u + S * j if j < JH else (S - u) + S * (RL - 1 - j)
                emit(Za[j], Zb[RL - 1 - j], kk)
    return magA, magB


def stockham(x, radices, sign):
    """Index algebra of block_fft / stockham_pass in fdoct_b200/csrc/prep_kernels.cu (autosort, natural order out):
    pass with radix r, current length nc, stride s: y[q + s (r p + c)] = (sum_j x[q + s (p + (nc/r) j)] w_r^(jc)) w_n^(p c s),
    with p c s < n so that the twiddle table needs no index reduction."""
    n = x.size
    a = x.astype(np.complex128).copy()
    s, ncur = 1, n
    tab = np.exp(sign * 2j * np.pi * np.arange(n) / n)
    for r in radices:
        mq = ncur // r
        y = np.zeros(n, dtype=np.complex128)
        F = dft_mat(r, sign)
        for p in range(mq):
            for q in range(s):
                assert p * s * (r - 1) < n
                y[q + s * (r * p + np.arange(r))] = (F @ a[q + s * (p + mq * np.arange(r))]) * tab[p * np.arange(r) * s]
        a, ncur, s = y, mq, s * r
    return a


def greedy_radices(n):
    out = []
    for f in (16, 15, 12, 10, 9, 8, 6, 5, 4, 3, 2):
        while n % f == 0:
            out.append(f)
            n //= f
    assert n == 1
    return out


if __name__ == "__main__":
    rng0 = np.random.default_rng(1)
    for n in (640, 720, 1920, 2560, 2880, 3840, 96):
        z = rng0.normal(size=n) + 1j * rng0.normal(size=n)
        for sgn in (+1, -1):
            ref = np.fft.ifft(z) * n if sgn > 0 else np.fft.fft(z)
            err = np.abs(stockham(z, greedy_radices(n), sgn) - ref).max() / np.abs(ref).max()
            assert err < 1e-12, (n, sgn, err)
        print("ok stockham", n, greedy_radices(n))
    rng = np.random.default_rng(0)
    plans = [(1024, 16, 8, 8), (2048, 16, 16, 8), (4096, 32, 16, 8), (1280, 20, 8, 8), (1920, 15, 16, 8),
             (3840, 30, 16, 8), (2560, 20, 16, 8), (2880, 30, 12, 8), (1280, 16, 16, 5), (1024, 32, 1, 32),
             (1920, 16, 8, 15), (640, 16, 1, 40), (128, 16, 1, 8), (960, 15, 8, 8)]
    for N, R0, R1, RL in plans:
        a = rng.normal(size=N)
        b = rng.normal(size=N)
        z = a + 1j * b
        for sgn in (+1, -1):
            X = run_plan(z, R0, R1, RL, sgn)
            ref = np.fft.ifft(z) * N if sgn > 0 else np.fft.fft(z)
            err = np.abs(X - ref).max() / np.abs(ref).max()
            assert err < 1e-12, (N, R0, R1, RL, sgn, err)
        X = run_plan(z, R0, R1, RL, +1)
        mA, mB = split_pairs(X, R0, R1, RL)
        rA = np.abs(np.fft.ifft(a) * N)[: N // 2]
        rB = np.abs(np.fft.ifft(b) * N)[: N // 2]
        assert not np.isnan(mA).any(), (N, np.isnan(mA).sum())
        assert np.allclose(mA, rA, rtol=1e-9, atol=1e-9) and np.allclose(mB, rB, rtol=1e-9, atol=1e-9), (N,)
        print("ok", N, R0, R1, RL)
