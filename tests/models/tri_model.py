"""Design model (numpy) of the FFT x recurrence form of the periodic Poisson solve  (csrc/vmk_tri.cuh).

Common.jl:97-125 divides the 2-D spectrum by  aa + bb cos(kx) + cc cos(ky).  For a fixed kx that divisor is the symbol
of the cyclic tridiagonal operator  (cc/2)(psi[j-1] + psi[j+1]) + (aa + bb cos kx) psi[j]  along j, so
    fft_j -> divide -> ifft_j     ==     one cyclic tridiagonal solve along j per kx
and, the coefficients being constant along j, the inverse is the two-sided geometric kernel K r^|m|:
    psi_j = K (y+_j + y-_j - x_j),  y+_j = x_j + r y+_(j-1),  y-_j = x_j + r y-_(j+1)  (cyclic),  r + 1/r = -b/a.
Two first-order recurrences (about 12 FP64 instructions per complex value) replace two length-N FFTs (about 100), and
along j they split into chunks whose only coupling is one carry per chunk and direction -- across GPUs too, so the
slab decomposition no longer needs the all-to-all transposes of the distributed 2-D FFT.

What the reference does differently from the exact operator, and how the model (and the kernels) keep its numbers:
  * ky[1] = eps (Common.jl:112-113): the j-mean of every row is divided by  b + cc cos(eps)  instead of  b + cc.
    -> rank-one correction from the row sum X0:  psi_j += X0/N (1/d_ref - 1/d_tri).
  * the FP64 evaluation of the divisor carries a rounding noise of ~|aa| 1e-16 = 7e-10 (8192^2) which matters only
    where |d| is small: rows kx < K0 (and the packed kx = 0 / N/2 row with its e[1,1] = 0 and near-singular kx = eps
    operator) keep the FFT form with the reference's literal divisor.  For kx >= K0: |d| >= bb (1 - cos(K0 hx)).

usage: python tests/models/tri_model.py [n] [K0] [chunk]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import oracle_np as onp  # noqa: E402


def row_tables(n, dx, dy, eps=1.e-6):
    """per kx = 0 .. n/2: r, K/1 and the eps-quirk coefficient, from the FP64 row constant the reference uses"""
    aa = -2 / dx**2 - 2 / dy**2
    bb = 2 / dx**2
    cc = 2 / dy**2
    ck = np.cos(onp.wavenumbers(n, eps))
    ab = aa + bb * ck[:n // 2 + 1]  # fl(aa + fl(bb cos kx)): exactly what K2 adds cc cos ky to
    L = np.longdouble
    a = L(cc) / 2
    b = ab.astype(L)
    delta = -(b / a) - 2  # = 4 (bb/cc) sin^2(theta/2) up to the rounding of ab
    delta = np.maximum(delta, L(0))
    r = 2 / (2 + delta + np.sqrt(delta * (delta + 4)))
    K = 1 / (a * (r - 1 / r))
    d_tri0 = b + L(cc)
    d_ref0 = (ab + cc * ck[0]).astype(L)  # fl(ab + fl(cc cos eps))
    with np.errstate(divide="ignore", invalid="ignore"):
        q0 = (1 / d_ref0 - 1 / d_tri0) / n
    return ab, r, K, q0


def solve_rows_recurrence(x, r, K, q0, chunk):
    """x: [rows, n] complex (one row per kx); chunked two-sided recurrence in FP64, carries in FP64"""
    rows, n = x.shape
    nch = n // chunk
    r64 = r.astype(np.float64)[:, None]
    xs = x.reshape(rows, nch, chunk)
    # phase 1: chunk totals with zero carry-in (Horner)
    tp = np.zeros((rows, nch), complex)
    tm = np.zeros((rows, nch), complex)
    for m in range(chunk):
        tp = tp * r64 + xs[:, :, m]
        tm = tm * r64 + xs[:, :, chunk - 1 - m]
    x0 = xs.sum(axis=2).sum(axis=1)
    # phase 1.5: carries.  c+[c] = y+ at the last row of chunk c-1 (cyclic), R = r^chunk
    R = (r**chunk).astype(np.float64)[:, None]
    RN = (r**n).astype(np.float64)
    # cyclic closure: run twice around (decays) -- model only; the kernel sums until the weight underflows
    cp = np.zeros((rows, nch), complex)
    acc = np.zeros(rows, complex)
    for sweep in range(2):
        for c in range(nch):
            cp[:, c] = acc
            acc = acc * R[:, 0] + tp[:, c]
    acc_exact = None
    # exact closure instead of a second sweep when R^nch is not negligible
    acc = np.zeros(rows, complex)
    for c in range(nch):
        acc = acc * R[:, 0] + tp[:, c]
    wrap = acc / (1 - RN)  # y+ at the last row of the whole line
    acc = wrap.copy()
    for c in range(nch):
        cp[:, c] = acc
        acc = acc * R[:, 0] + tp[:, c]
    cm = np.zeros((rows, nch), complex)
    acc = np.zeros(rows, complex)
    for c in range(nch - 1, -1, -1):
        acc = acc * R[:, 0] + tm[:, c]
    wrap = acc / (1 - RN)
    acc = wrap.copy()
    for c in range(nch - 1, -1, -1):
        cm[:, c] = acc
        acc = acc * R[:, 0] + tm[:, c]
    # phase 2
    out = np.empty_like(xs)
    p = np.empty_like(xs)
    y = cp.copy()
    for m in range(chunk):
        p[:, :, m] = y * r64
        y = p[:, :, m] + xs[:, :, m]
    y = cm.copy()
    K64 = K.astype(np.float64)[:, None]
    dc = (x0 * q0.astype(np.float64))[:, None]
    for m in range(chunk - 1, -1, -1):
        y = y * r64 + xs[:, :, m]
        out[:, :, m] = (p[:, :, m] + y) * K64 + dc
    return out.reshape(rows, n)


def poisson_tri(n, dx, dy, f, K0=64, chunk=32):
    ab, r, K, q0 = row_tables(n, dx, dy)
    cc = 2 / dy**2
    ck = np.cos(onp.wavenumbers(n))
    X = np.fft.rfft(f, axis=0)  # [kx, j]
    out = np.empty_like(X)
    # low rows and the Nyquist row: the reference's literal form
    low = list(range(K0)) + [n // 2]
    E = np.fft.fft(X[low, :], axis=1)
    d = ab[low, None] + (cc * ck)[None, :]
    E[0, 0] = 0
    out[low, :] = np.fft.ifft(E / d, axis=1)
    hi = np.arange(K0, n // 2)
    out[hi, :] = solve_rows_recurrence(X[hi, :], r[hi], K[hi], q0[hi], chunk)
    return np.fft.irfft(out, n=n, axis=0)


def solve_rows_fused(x, r, q0, cc, units):
    """The fused form (csrc/vmk_kernels.cuh, k1_body / k3_body with FUSED; kt_scanf_body), same blocking and the same
    FP64 operations per value: unit q owns the row pairs [q npairs / U, (q+1) npairs / U).
      K1: u0_m = x_m + r u0_(m-1) from zero at the block's first row; tp = u0 at its last row,
          A = Horner in 1/r of u0 (al = r^(M-1) A), S = sum u0 (sum of x = (1 - r) S + r tp)
      scan: blocks compose as (tp, al, R = r^M, Gam = r (1 - R^2) / (1 - r^2)); cyclic closure W = 1 / (1 - r^N)
      K3: last row first, v_m = u0_m + q_m + r v_(m+1), q_(m-1) = q_m / r from q = r^M cu, psi_m = -(2 r / cc) v_m + dc
    x: [rows, n] complex, one row per kx; r: long double per row."""
    rows, n = x.shape
    npairs = n // 2
    first = [(q * npairs) // units for q in range(units + 1)]
    r64 = r.astype(np.float64)
    rinv = (1 / r).astype(np.float64)
    L = np.longdouble
    rl = r64.astype(L)  # the tables are powers of the rounded r
    u0 = np.empty_like(x)
    tp = np.zeros((rows, units), complex)
    al = np.zeros((rows, units), complex)
    xs = np.zeros(rows, complex)
    Rb = np.ones((rows, units))
    Gb = np.zeros((rows, units))
    for q in range(units):
        j0, j1 = 2 * first[q], 2 * first[q + 1]
        M = j1 - j0
        if M == 0:
            continue
        u = np.zeros(rows, complex)
        A = np.zeros(rows, complex)
        S = np.zeros(rows, complex)
        for j in range(j0, j1):
            u = u * r64 + x[:, j]
            A = A * rinv + u
            S = S + u
            u0[:, j] = u
        R = rl**M
        tp[:, q] = u
        al[:, q] = A * (rl**(M - 1)).astype(np.float64)
        xs += S * (1.0 - r64) + r64 * u
        Rb[:, q] = R.astype(np.float64)
        Gb[:, q] = (rl * (1 - R * R) / (1 - rl * rl)).astype(np.float64)
    # upwards over the blocks, cyclic closure, carries back down
    lp = np.zeros(rows, complex)
    AL = np.zeros(rows, complex)
    rp = np.ones(rows)
    for q in range(units):
        z = lp * Gb[:, q] + al[:, q]
        AL = AL + z * rp
        rp = rp * Rb[:, q]
        lp = lp * Rb[:, q] + tp[:, q]
    RN = r**n
    W = (1 / (1 - RN)).astype(np.float64)
    GJ = (r * (1 - RN * RN) / (1 - r * r)).astype(np.float64)
    cu0 = lp * W
    cvd = (cu0 * GJ + AL) * W
    cu = np.zeros((rows, units), complex)
    acc = cu0
    for q in range(units):
        cu[:, q] = acc
        acc = acc * Rb[:, q] + tp[:, q]
    cv = np.zeros((rows, units), complex)
    acc = cvd
    for q in range(units - 1, -1, -1):
        cv[:, q] = acc
        acc = acc * Rb[:, q] + (cu[:, q] * Gb[:, q] + al[:, q])
    dc = xs * q0.astype(np.float64)
    kv = -2.0 * r64 / cc
    out = np.empty_like(x)
    for q in range(units):
        j0, j1 = 2 * first[q], 2 * first[q + 1]
        if j1 == j0:
            continue
        y = cv[:, q]
        qq = cu[:, q] * Rb[:, q]
        for j in range(j1 - 1, j0 - 1, -1):
            y = y * r64 + (u0[:, j] + qq)
            qq = qq * rinv
            out[:, j] = y * kv + dc
    return out


def poisson_fused(n, dx, dy, f, K0=64, units=148):
    ab, r, K, q0 = row_tables(n, dx, dy)
    cc = 2 / dy**2
    ck = np.cos(onp.wavenumbers(n))
    X = np.fft.rfft(f, axis=0)  # [kx, j]
    out = np.empty_like(X)
    low = list(range(K0)) + [n // 2]
    E = np.fft.fft(X[low, :], axis=1)
    d = ab[low, None] + (cc * ck)[None, :]
    E[0, 0] = 0
    out[low, :] = np.fft.ifft(E / d, axis=1)
    hi = np.arange(K0, n // 2)
    out[hi, :] = solve_rows_fused(X[hi, :], r[hi], q0[hi], cc, units)
    return np.fft.irfft(out, n=n, axis=0)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "fused":  # python tests/models/tri_model.py fused [n] [K0] [units]
        n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
        K0 = int(sys.argv[3]) if len(sys.argv) > 3 else 64
        units = int(sys.argv[4]) if len(sys.argv) > 4 else 148
        dx = dy = 2 * np.pi / n
        f = np.random.default_rng(3).standard_normal((n, n))
        ref = onp.poisson(n, n, dx, dy, f)
        got = poisson_fused(n, dx, dy, f, K0, units)
        print(f"fused form: n={n} K0={K0} units={units}  white noise  rel-L2 = "
              f"{np.linalg.norm(got - ref) / np.linalg.norm(ref):.3e}")
        return
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    K0 = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 32
    dx = dy = 2 * np.pi / n
    x = dx * np.arange(n + 1)
    w = np.zeros((n + 2, n + 2), order="F")
    onp.vm_ic(n, n, x, x, w)
    rng = np.random.default_rng(3)
    for name, f in (("vortex merger", -w[1:n + 1, 1:n + 1]), ("white noise", rng.standard_normal((n, n)))):
        ref = onp.poisson(n, n, dx, dy, f)
        got = poisson_tri(n, dx, dy, f, K0, chunk)
        print(f"n={n} K0={K0} chunk={chunk}  {name:14s} rel-L2 = {np.linalg.norm(got - ref) / np.linalg.norm(ref):.3e}   "
              f"max = {np.abs(got - ref).max() / np.abs(ref).max():.3e}", flush=True)


if __name__ == "__main__":
    main()
