"""Executable spec (numpy) of the in-shared-memory FFT scheme used by the CUDA kernels.

Not part of the product or of the tests' oracle: it is the design notebook that the index
math in cfd_julia_b200/csrc/vmk_fft.cuh was derived from, kept so the derivation can be re-run.

  * mixed-radix in-place DIF forward (natural in -> digit-reversed "position" order out),
    adjoint DIT inverse (position order in -> natural out)
  * K1: two real rows -> one complex FFT -> unpack to two half spectra (no twiddles)
  * K2: column FFT -> divide -> inverse, incl. the packed DC/Nyquist row 0
  * K3: repack -> inverse -> two real rows
  * shared-memory bank-conflict model for the padded layouts
"""
import numpy as np


def pass_bits(M, LE):
    P = -(-M // LE)
    base, extra = divmod(M, P)
    bits = [base + (1 if k >= P - extra else 0) for k in range(P)]   # larger radices last
    assert sum(bits) == M and max(bits) <= LE
    return bits


def geometry(M, bits):
    h = M
    out = []
    for b in bits:
        out.append((h, h - b, b))   # (hi, lo, b)
        h -= b
    return out


def k_of_pos(M, bits):
    """spectral index held at each position after the forward DIF."""
    N = 1 << M
    pos = np.arange(N)
    k = np.zeros(N, dtype=np.int64)
    mul = 1
    for (h, l, b) in geometry(M, bits):
        k += ((pos >> l) & ((1 << b) - 1)) * mul
        mul <<= b
    return k


def dif_fwd(x, M, bits, sign=-1):
    N = 1 << M
    a = np.asarray(x, dtype=np.complex128).copy()
    for (h, l, b) in geometry(M, bits):
        r, B = 1 << b, 1 << h
        a = a.reshape(N >> h, r, 1 << l)
        q = np.arange(r)
        F = np.exp(sign * 2j * np.pi * np.outer(q, q) / r)           # F[p,q]
        y = np.einsum("pq,hql->hpl", F, a)
        low = np.arange(1 << l)
        tw = np.exp(sign * 2j * np.pi * np.outer(q, low) / B)         # tw[p,low]
        a = (y * tw[None]).reshape(N)
    return a


def dit_inv(a, M, bits):
    """adjoint of dif_fwd (unnormalised inverse): position order in, natural order out."""
    N = 1 << M
    a = np.asarray(a, dtype=np.complex128).copy()
    for (h, l, b) in reversed(geometry(M, bits)):
        r, B = 1 << b, 1 << h
        a = a.reshape(N >> h, r, 1 << l)
        q = np.arange(r)
        low = np.arange(1 << l)
        tw = np.exp(+2j * np.pi * np.outer(q, low) / B)
        F = np.exp(+2j * np.pi * np.outer(q, q) / r)
        a = np.einsum("qp,hpl->hql", F, a * tw[None]).reshape(N)
    return a


def poisson_r2c(f, dx, dy, M, bits, eps=1e-6, sign_in=+1):
    """Whole K1->K2->K3 pipeline on an N x N real field f[i, j] (i contiguous)."""
    N = 1 << M
    kp = k_of_pos(M, bits)
    pos_of_k = np.argsort(kp)
    # ---- K1: rows j, j+1 packed; T[kx, j] for kx < N/2 (row 0 = packed DC/Nyquist)
    T = np.zeros((N // 2, N), dtype=np.complex128)
    for j in range(0, N, 2):
        Zp = dif_fwd(f[:, j] + 1j * f[:, j + 1], M, bits)
        Z = Zp[pos_of_k]                                  # natural k (spec only)
        Zc = np.conj(Z[(-np.arange(N)) % N])
        A2 = Z + Zc
        B2 = -1j * (Z - Zc)
        T[1:, j] = A2[1:N // 2]
        T[1:, j + 1] = B2[1:N // 2]
        T[0, j] = A2[0].real + 1j * A2[N // 2].real
        T[0, j + 1] = B2[0].real + 1j * B2[N // 2].real
    # ---- divisor tables (Common.jl:101-113,120)
    aa = -2 / dx**2 - 2 / dy**2
    hx = 2 * np.pi / N
    kx = np.empty(N)
    i = np.arange(1, N // 2 + 1)
    kx[i - 1] = hx * (i - 1)
    kx[i + N // 2 - 1] = hx * (i - N // 2 - 1)
    kx[0] = eps
    bbcos = (2 / dx**2) * np.cos(kx)
    cccos = (2 / dy**2) * np.cos(kx)
    scale = sign_in / (2.0 * N * N)
    # ---- K2
    U = np.zeros_like(T)
    for r in range(N // 2):
        C = dif_fwd(T[r], M, bits)                        # position order, ky = kp
        if r > 0:
            d = (aa + bbcos[r]) + cccos[kp]
            U[r] = dit_inv(C * (scale / d), M, bits)
        else:
            partner = pos_of_k[(-kp) % N]
            Cc = np.conj(C[partner])
            Ah = 0.5 * (C + Cc)
            Bh = -0.5j * (C - Cc)
            d0 = (aa + bbcos[0]) + cccos[kp]
            dn = (aa + bbcos[N // 2]) + cccos[kp]
            Pp = Ah * (scale / d0)
            Pp[pos_of_k[0]] = 0.0                          # e[1,1] = 0, Common.jl:118
            Qp = Bh * (scale / dn)
            U[0] = dit_inv(Pp + 1j * Qp, M, bits)
    # ---- K3
    psi = np.zeros((N, N))
    for j in range(0, N, 2):
        Zin = np.zeros(N, dtype=np.complex128)
        a, b = U[:, j], U[:, j + 1]
        Zin[1:N // 2] = a[1:] + 1j * b[1:]
        Zin[N - np.arange(1, N // 2)] = np.conj(a[1:]) + 1j * np.conj(b[1:])
        Zin[0] = a[0].real + 1j * b[0].real
        Zin[N // 2] = a[0].imag + 1j * b[0].imag
        z = dit_inv(Zin[kp], M, bits)                      # scatter to position order
        psi[:, j] = z.real
        psi[:, j + 1] = z.imag
    return psi


# ----------------------------------------------------------------------------------------
def conflict_degree(addrs16):
    """addrs16: (32,) 16-byte-unit shared-memory addresses of one LDS.128/STS.128 warp op.
    Returns cycles (sum over 4 quarter-warps of the max multiplicity over the 8 bank groups)."""
    cyc = 0
    for qw in range(4):
        a = addrs16[8 * qw:8 * qw + 8]
        uniq = {}
        for x in a:
            uniq.setdefault(x % 8, set()).add(x)
        cyc += max(len(v) for v in uniq.values())
    return cyc


def data_addr(pos):
    return pos + (pos >> 4)


def report_conflicts(M, LE, tw_addr):
    bits = pass_bits(M, LE)
    N, E, T = 1 << M, 1 << LE, 1 << (M - LE)
    print(f"M={M} bits={bits} T={T}")
    lanes = np.arange(32)
    for K, (h, l, b) in enumerate(geometry(M, bits)):
        r = 1 << b
        worst_d, tot_t, n_t = 0, 0, 0
        for warp in range(max(1, T // 32)):
            t = warp * 32 + lanes
            for u in range(E // r):
                bid = t + T * u
                low, high = bid & ((1 << l) - 1), bid >> l
                for q in range(r):
                    pos = (high << h) | (q << l) | low
                    worst_d = max(worst_d, conflict_degree(data_addr(pos)))
                    if q and K < len(bits) - 1:
                        x = (low * q) & ((1 << (h - 2)) - 1) if h >= 2 else low * q
                        tot_t += conflict_degree(tw_addr(x))
                        n_t += 1
        print(f"  pass {K}: radix {r:2d} data worst cyc/op {worst_d} (ideal 4)"
              + (f", twiddle avg cyc/op {tot_t / n_t:.2f}" if n_t else ""))


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for M, LE in [(5, 4), (6, 4), (7, 4), (8, 4), (9, 4), (10, 4), (11, 4), (12, 4), (13, 5)]:
        bits = pass_bits(M, LE)
        N = 1 << M
        x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
        X = np.fft.fft(x)
        a = dif_fwd(x, M, bits)
        e1 = np.abs(a - X[k_of_pos(M, bits)]).max()
        e2 = np.abs(dit_inv(a, M, bits) / N - x).max()
        print(M, bits, "fwd err", e1, "roundtrip err", e2)
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    from oracle import oracle_np as onp
    for M in (5, 6):
        N = 1 << M
        bits = pass_bits(M, 4)
        f = rng.uniform(-1, 1, (N, N))
        dx = dy = 2 * np.pi / N
        ref = onp.poisson(N, N, dx, dy, f)
        got = poisson_r2c(f, dx, dy, M, bits)
        print("poisson r2c", N, np.linalg.norm(got - ref) / np.linalg.norm(ref))
    for name, fn in [("nopad", lambda x: x), ("pad8", lambda x: x + (x >> 3)), ("pad16", lambda x: x + (x >> 4))]:
        print("== twiddle table layout", name)
        report_conflicts(13, 5, fn)
        report_conflicts(12, 4, fn)
