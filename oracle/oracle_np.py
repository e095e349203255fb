"""numpy restatement of the CFD_Julia vortex-merger hot path (second, independent oracle).

TEST INFRASTRUCTURE, NOT THE PRODUCT: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.

It follows the same reference lines as oracle/vm_oracle.c (Common.jl:97-182,208-237,
vm.jl:12-90, tgv.jl:82-90, fft_p.jl:8-42) but uses numpy's pocketfft instead of the C file's
radix-2 FFT, so the two oracles cross-check each other's FFT and index conventions.
Arrays are Fortran-ordered (column-major, like Julia); "ghosted" = shape (nx+2, ny+2).

Parity pins (tests/test_oracle.py): order.jl:13 (five fft_p.jl L2 errors); the outputs of the reference's own
Python twins of script 19 run unmodified (tests/golden/make_ref_fixtures.py -> tests/golden/ref_py_*.npz, agreement
4e-16..4e-15 on vorticity, streamfunction, rhs); the analytic Taylor-Green solution.  The Julia scripts themselves
cannot run here (no julia / FFTW), so the f-row solvers (hybrid, pseudo-spectral, cavity) stay pinned only through
the shared Arakawa / Poisson pieces and closed-form solutions: "parity unpinned" for their script-specific parts.
"""
from __future__ import annotations

import numpy as np

# 13_Poisson_Solver_FFT_Spectral/specrtral_vs_FDM/order.jl:13 -- the only recorded reference outputs
ORDER_JL_FFT_FDM = {
    32: .0015607100315532957,
    64: .0005987381110678801,
    128: .00014313734718665358,
    256: 3.549617203207291e-5,
    512: 8.865373334924762e-6,
}


def wavenumbers(nx: int, eps: float = 1.e-6) -> np.ndarray:
    """Common.jl:106-112."""
    hx = 2 * np.pi / nx
    kx = np.empty(nx)
    i = np.arange(1, nx // 2 + 1)
    kx[i - 1] = hx * (i - 1)
    kx[i + nx // 2 - 1] = hx * (i - nx // 2 - 1)
    kx[0] = eps
    return kx


def divisor(nx, ny, dx, dy, eps=1.e-6):
    """Common.jl:101-103,119-121: aa + bb*cos(kx[i]) + cc*cos(ky[j]) (ky = kx, :113)."""
    assert nx == ny
    aa = -2 / dx**2 - 2 / dy**2
    bb = 2 / dx**2
    cc = 2 / dy**2
    ck = np.cos(wavenumbers(nx, eps))
    return (aa + bb * ck[:, None]) + cc * ck[None, :]


def poisson(nx, ny, dx, dy, f, eps=1.e-6):
    """real(ifft(fft(f)/divisor)) with e[1,1]=0 (Common.jl:115-123); f is nx x ny."""
    e = np.fft.fft2(np.asarray(f, dtype=np.complex128))
    e[0, 0] = 0
    return np.real(np.fft.ifft2(e / divisor(nx, ny, dx, dy, eps)))


def fps(nx, ny, dx, dy, f, s, eps=1.e-6):
    s[1:nx + 1, 1:ny + 1] = poisson(nx, ny, dx, dy, f[:nx, :ny], eps)


def ps_fft(nx, ny, dx, dy, f, eps=1.e-6):
    """fft_p.jl:8-42; f is (nx+1)x(ny+1)."""
    return poisson(nx, ny, dx, dy, f[:nx, :ny], eps)


def ghost_fill(nx, ny, a):
    """Common.jl:138-146 / vm.jl:30-38 order."""
    a[nx + 1, :] = a[1, :]
    a[:, ny + 1] = a[:, 1]
    a[0, :] = a[nx, :]
    a[:, 0] = a[:, ny]


def vm_rhs(nx, ny, dx, dy, re, w, r, s, f):
    """Common.jl:132-182."""
    f[:nx, :ny] = -w[1:nx + 1, 1:ny + 1]
    fps(nx, ny, dx, dy, f, s)
    ghost_fill(nx, ny, s)
    aa = 1 / (re * dx**2)
    bb = 1 / (re * dy**2)
    gg = 1 / (4 * dx * dy)
    hh = 1 / 3
    c = slice(1, nx + 1), slice(1, ny + 1)

    def sh(a, di, dj):
        return a[1 + di:nx + 1 + di, 1 + dj:ny + 1 + dj]

    j1 = ((sh(w, 1, 0) - sh(w, -1, 0)) * (sh(s, 0, 1) - sh(s, 0, -1)) -
          (sh(w, 0, 1) - sh(w, 0, -1)) * (sh(s, 1, 0) - sh(s, -1, 0)))
    j2 = (sh(w, 1, 0) * (sh(s, 1, 1) - sh(s, 1, -1)) -
          sh(w, -1, 0) * (sh(s, -1, 1) - sh(s, -1, -1)) -
          sh(w, 0, 1) * (sh(s, 1, 1) - sh(s, -1, 1)) +
          sh(w, 0, -1) * (sh(s, 1, -1) - sh(s, -1, -1)))
    j3 = (sh(w, 1, 1) * (sh(s, 0, 1) - sh(s, 1, 0)) -
          sh(w, -1, -1) * (sh(s, -1, 0) - sh(s, 0, -1)) -
          sh(w, -1, 1) * (sh(s, 0, 1) - sh(s, -1, 0)) +
          sh(w, 1, -1) * (sh(s, 1, 0) - sh(s, 0, -1)))
    jac = gg * (j1 + j2 + j3) * hh
    r[c] = -jac + (aa * (sh(w, 1, 0) - 2 * sh(w, 0, 0) + sh(w, -1, 0)) +
                   bb * (sh(w, 0, 1) - 2 * sh(w, 0, 0) + sh(w, 0, -1)))


def numerical(nx, ny, nt, dx, dy, dt, re, wn, snap=None, freq=0):
    """vm.jl:12-90 / tgv.jl:13-79 (no file output).  Returns (wn[2:nx+2,2:ny+2], last psi ghosted)."""
    wt = np.zeros_like(wn)
    r = np.zeros_like(wn)
    s = np.zeros_like(wn)
    f = np.zeros((nx, ny), order="F")
    c = slice(1, nx + 1), slice(1, ny + 1)
    for k in range(1, nt + 1):
        vm_rhs(nx, ny, dx, dy, re, wn, r, s, f)
        wt[c] = wn[c] + dt * r[c]
        ghost_fill(nx, ny, wt)
        vm_rhs(nx, ny, dx, dy, re, wt, r, s, f)
        wt[c] = .75 * wn[c] + .25 * wt[c] + (.25 * dt) * r[c]
        ghost_fill(nx, ny, wt)
        vm_rhs(nx, ny, dx, dy, re, wt, r, s, f)
        wn[c] = wn[c] / 3. + (2 / 3) * wt[c] + ((2 / 3) * dt) * r[c]
        ghost_fill(nx, ny, wn)
        if snap is not None and freq and k % freq == 0:
            snap(k, wn)
    return wn[1:nx + 2, 1:ny + 2].copy(order="F"), s


def vm_ic(nx, ny, x, y, w):
    """Common.jl:208-219 followed by main()'s ghost fill vm.jl:121-128."""
    sigma = np.pi
    xc1, yc1 = np.pi - np.pi / 4., np.pi
    xc2, yc2 = np.pi + np.pi / 4., np.pi
    X = x[:, None]
    Y = y[None, :]
    w[1:nx + 2, 1:ny + 2] = (np.exp(-sigma * ((X - xc1)**2 + (Y - yc1)**2)) +
                             np.exp(-sigma * ((X - xc2)**2 + (Y - yc2)**2)))
    w[0, :] = w[nx, :]
    w[:, 0] = w[:, ny]
    w[nx + 1, :] = w[1, :]
    w[:, ny + 1] = w[:, 1]


def exact_tgv(nx, ny, x, y, time, re):
    """tgv.jl:82-90."""
    nq = 4
    return np.asfortranarray(2 * nq * np.cos(nq * x[:, None]) * np.cos(nq * y[None, :]) *
                             np.exp(-2 * nq**2 * time / re))


def compute_l2norm_bnds(nx, ny, r):
    """Common.jl:234-237."""
    return float(np.sqrt(np.sum(r[:nx + 1, :ny + 1]**2) / ((nx + 1) * (ny + 1))))


def fft_p_case(nx):
    """fft_p.jl:44-108 for one grid size: returns (rms_error, max_error)."""
    ny = nx
    dx = 1. / nx
    dy = 1. / ny
    x = dx * np.arange(nx + 1)
    y = dy * np.arange(ny + 1)
    km = 16
    c1 = (1. / km)**2
    c2 = -8 * np.pi**2
    X = x[:, None]
    Y = y[None, :]
    ue = (np.sin(2 * np.pi * X) * np.sin(2 * np.pi * Y) +
          c1 * np.sin(km * 2 * np.pi * X) * np.sin(km * 2 * np.pi * Y))
    f = (c2 * np.sin(2 * np.pi * X) * np.sin(2 * np.pi * Y) +
         c2 * np.sin(km * 2 * np.pi * X) * np.sin(km * 2 * np.pi * Y))
    un = np.zeros_like(f)
    un[:nx, :ny] = ps_fft(nx, ny, dx, dy, f)
    un[nx, :] = un[0, :]
    un[:, ny] = un[:, 0]
    err = un - ue
    return compute_l2norm_bnds(nx, ny, err), float(np.max(np.abs(err)))


def grid(nx, ny, lx=2 * np.pi, ly=2 * np.pi):
    dx = lx / nx
    dy = ly / ny
    return dx, dy, dx * np.arange(nx + 1), dy * np.arange(ny + 1)


# ---- 20_NS2D_Hybrid_Solver/hybrid.jl ------------------------------------------------------------
def wavespace(nx, ny, dx, dy, eps=1.e-6):
    """Common.jl:184-204: k2[i,j] = kx[i]^2 + ky[j]^2, hx = 2 pi/(nx dx), kx[1] = eps, ky = kx."""
    hx = 2 * np.pi / (nx * dx)
    kx = np.empty(nx)
    i = np.arange(1, nx // 2 + 1)
    kx[i - 1] = hx * (i - 1.)
    kx[i + nx // 2 - 1] = hx * (i - nx // 2 - 1)
    kx[0] = eps
    ky = kx
    return kx[:, None]**2 + ky[None, :]**2


def hybrid_jacobian(nx, ny, dx, dy, wf, k2):
    """hybrid.jl:96-152: jf = fft(-J(w, psi)), w = real(ifft(wf)), psi = real(ifft(wf / k2)), Arakawa J."""
    gg = 1. / (4. * dx * dy)
    hh = 1. / 3.

    def ghosted(a):
        g = np.empty((nx + 2, ny + 2))
        g[1:nx + 1, 1:ny + 1] = a
        g[nx + 1, :] = g[1, :]
        g[:, ny + 1] = g[:, 1]
        g[0, :] = g[nx, :]
        g[:, 0] = g[:, ny]
        return g

    w = ghosted(np.real(np.fft.ifft2(wf)))
    s = ghosted(np.real(np.fft.ifft2(wf / k2)))
    c = slice(1, nx + 1)
    p = slice(2, nx + 2)
    m = slice(0, nx)
    j1 = gg * ((w[p, c] - w[m, c]) * (s[c, p] - s[c, m]) - (w[c, p] - w[c, m]) * (s[p, c] - s[m, c]))
    j2 = gg * (w[p, c] * (s[p, p] - s[p, m]) - w[m, c] * (s[m, p] - s[m, m]) -
               w[c, p] * (s[p, p] - s[m, p]) + w[c, m] * (s[p, m] - s[m, m]))
    j3 = gg * (w[p, p] * (s[c, p] - s[p, c]) - w[m, m] * (s[m, c] - s[c, m]) -
               w[m, p] * (s[c, p] - s[m, c]) + w[p, m] * (s[p, c] - s[c, m]))
    return np.fft.fft2(-(j1 + j2 + j3) * hh)


def hybrid_numerical(nx, ny, nt, dx, dy, dt, re, wn, freq=0, snapshot=None):
    """hybrid.jl:14-90: RK3 (explicit Jacobian) / Crank-Nicolson (implicit diffusion) in Fourier space.  Returns
    ut = real(ifft(wnf)) with the periodic duplicate row/column, (nx+1) x (ny+1); wn (ghosted) is only read.
    snapshot(k, ut) every `freq` steps (hybrid.jl:71-86; the text dump stays with the caller)."""
    assert nx == ny
    k2 = wavespace(nx, ny, dx, dy)
    wnf = np.fft.fft2(wn[1:nx + 1, 1:ny + 1].astype(np.complex128))
    wnf[0, 0] = 0.
    a1, a2, a3 = 8. / 15., 2. / 15., 1. / 3.
    g1, g2, g3 = 8. / 15., 5. / 12., 3. / 4.
    r2, r3 = -17. / 60., -5. / 12.
    z = .5 * dt * k2 / re
    d1, d2, d3 = a1 * z, a2 * z, a3 * z

    def field(wf):
        ut = np.empty((nx + 1, ny + 1), order="F")
        ut[:nx, :ny] = np.real(np.fft.ifft2(wf))
        ut[nx, :] = ut[0, :]
        ut[:, ny] = ut[:, 0]
        return ut

    for k in range(1, nt + 1):
        jnf = hybrid_jacobian(nx, ny, dx, dy, wnf, k2)
        w1f = ((1. - d1) / (1. + d1)) * wnf + (g1 * dt * jnf) / (1. + d1)
        w1f[0, 0] = 0.
        j1f = hybrid_jacobian(nx, ny, dx, dy, w1f, k2)
        w2f = ((1. - d2) / (1. + d2)) * w1f + (r2 * dt * jnf + g2 * dt * j1f) / (1. + d2)
        w2f[0, 0] = 0.
        j2f = hybrid_jacobian(nx, ny, dx, dy, w2f, k2)
        wnf = ((1. - d3) / (1. + d3)) * w2f + (r3 * dt * j1f + g3 * dt * j2f) / (1. + d3)
        if snapshot is not None and freq > 0 and k % freq == 0:
            snapshot(k, field(wnf))
    return field(wnf)


# ---- 18_NS2D_Lid_Driven_Cavity/lid_driven_cavity.jl ----------------------------------------------
# Arrays are (nx+1) x (ny+1) node arrays, no ghost cells; index [i, j] here = Julia [i+1, j+1].
def ldc_rhs(nx, ny, dx, dy, re, w, s, r):
    """lid_driven_cavity.jl:123-158: r[2:nx, 2:ny] = -J(w, s) (Arakawa) + (1/re) lap(w)."""
    aa = 1 / (re * dx**2)
    bb = 1 / (re * dy**2)
    gg = 1 / (4 * dx * dy)
    hh = 1 / 3
    c = slice(1, nx)
    p = slice(2, nx + 1)
    m = slice(0, nx - 1)
    j1 = gg * ((w[p, c] - w[m, c]) * (s[c, p] - s[c, m]) - (w[c, p] - w[c, m]) * (s[p, c] - s[m, c]))
    j2 = gg * (w[p, c] * (s[p, p] - s[p, m]) - w[m, c] * (s[m, p] - s[m, m]) -
               w[c, p] * (s[p, p] - s[m, p]) + w[c, m] * (s[p, m] - s[m, m]))
    j3 = gg * (w[p, p] * (s[c, p] - s[p, c]) - w[m, m] * (s[m, c] - s[c, m]) -
               w[m, p] * (s[c, p] - s[m, c]) + w[p, m] * (s[p, c] - s[c, m]))
    jac = (j1 + j2 + j3) * hh
    r[c, c] = -jac + (aa * (w[p, c] - 2 * w[c, c] + w[m, c]) + bb * (w[c, p] - 2 * w[c, c] + w[c, m]))


def ldc_bc2(nx, ny, dx, dy, w, s):
    """lid_driven_cavity.jl:38-52 (Jensen): left/right walls for every j, then bottom/top (lid: -3/dy) for every i."""
    w[0, :] = (-4 * s[1, :] + .5 * s[2, :]) / dx**2
    w[nx, :] = (-4 * s[nx - 1, :] + .5 * s[nx - 2, :]) / dx**2
    w[:, 0] = (-4 * s[:, 1] + .5 * s[:, 2]) / dy**2
    w[:, ny] = (-4 * s[:, ny - 1] + .5 * s[:, ny - 2]) / dy**2 - 3. / dy


def ldc_iden(nx, ny, dx, dy):
    """lid_driven_cavity.jl:66-71: iden[i,j] = 1/((2/dx^2)(cos(pi i/nx) - 1) + (2/dy^2)(cos(pi j/ny) - 1)), i, j from 1."""
    i = np.arange(1, nx + 2)
    j = np.arange(1, ny + 2)
    return 1. / ((2 / dx**2) * (np.cos(np.pi * i / nx) - 1.)[:, None] + (2 / dy**2) * (np.cos(np.pi * j / ny) - 1.)[None, :])


def ldc_fps_sine(nx, ny, f, iden, sn):
    """lid_driven_cavity.jl:11-21: sn[2:nx, 2:ny] = RODFT00(RODFT00(f[2:nx, 2:ny]) * iden) / ((2nx)(2ny)).
    FFTW's RODFT00 is scipy's unnormalised DST-I."""
    from scipy.fft import dstn
    e = dstn(f[1:nx, 1:ny], type=1)
    sn[1:nx, 1:ny] = dstn(e * iden[:nx - 1, :ny - 1], type=1) / ((2 * nx) * (2 * ny))


def ldc_numerical(nx, ny, nt, dx, dy, dt, re, wn, sn, rms):
    """lid_driven_cavity.jl:59-117: RK3 steps; mutates wn, sn (node arrays) and rms[0:nt]."""
    wt = np.zeros_like(wn)
    r = np.zeros_like(wn)
    iden = ldc_iden(nx, ny, dx, dy)
    c = slice(1, nx)
    for k in range(nt):
        sp = sn.copy()
        ldc_rhs(nx, ny, dx, dy, re, wn, sn, r)
        wt[c, c] = wn[c, c] + dt * r[c, c]
        ldc_bc2(nx, ny, dx, dy, wt, sn)
        ldc_fps_sine(nx, ny, -wt, iden, sn)
        ldc_rhs(nx, ny, dx, dy, re, wt, sn, r)
        wt[c, c] = .75 * wn[c, c] + .25 * wt[c, c] + .25 * dt * r[c, c]
        ldc_bc2(nx, ny, dx, dy, wt, sn)
        ldc_fps_sine(nx, ny, -wt, iden, sn)
        ldc_rhs(nx, ny, dx, dy, re, wt, sn, r)
        wn[c, c] = (1 / 3) * wn[c, c] + (2 / 3) * wt[c, c] + (2 / 3) * dt * r[c, c]
        ldc_bc2(nx, ny, dx, dy, wn, sn)
        ldc_fps_sine(nx, ny, -wn, iden, sn)
        rms[k] = np.sqrt(np.sum((sn - sp)**2) / ((nx + 1) * (ny + 1)))


# ---- 21_NS2D_PseudoSpectral_32_Rule, 22_NS2D_PseudoSpectral_23_Rule -------------------------------
def ps_wavenumbers(nx, dx, eps=1.e-6):
    """pseudospectral_23_rule.jl:96-107 (= pseudospectral_32_rule.jl:96-107): hx = 2 pi/(nx dx), kx[1] = eps."""
    hx = 2 * np.pi / (nx * dx)
    kx = np.empty(nx)
    i = np.arange(1, nx // 2 + 1)
    kx[i - 1] = hx * (i - 1.)
    kx[i + nx // 2 - 1] = hx * (i - nx // 2 - 1)
    kx[0] = eps
    return kx


def ps23_jacobian(nx, ny, dx, dy, wf, k2):
    """pseudospectral_23_rule.jl:95-144: the four derivative spectra, truncated with the 2/3 rule (modes i in
    floor(nxe/2)+1 .. nx-floor(nxe/2), 1-based, are zeroed in both directions: the retained band is -K .. K-1 with
    K = floor(floor(2nx/3)/2), NOT symmetric), real(ifft) of each, the product in real space and its fft."""
    kx = ps_wavenumbers(nx, dx)
    ky = kx
    j1f = 1j * wf * kx[:, None] / k2
    j4f = 1j * wf * kx[:, None]
    j2f = 1j * wf * ky[None, :]
    j3f = 1j * wf * ky[None, :] / k2
    nxe = int(np.floor(2 * nx / 3))
    nye = int(np.floor(2 * ny / 3))
    for a in (j1f, j2f, j3f, j4f):
        a[nxe // 2:nx - nxe // 2, :] = 0.   # 1-based floor(nxe/2)+1 : nx-floor(nxe/2)
        a[:, nye // 2:ny - nye // 2] = 0.
    j1, j2, j3, j4 = (np.real(np.fft.ifft2(a)) for a in (j1f, j2f, j3f, j4f))
    return np.fft.fft2(j1 * j2 - j3 * j4)


def ps32_jacobian(nx, ny, dx, dy, wf, k2):
    """pseudospectral_32_rule.jl:95-177: the four derivative spectra zero-padded to 1.5nx x 1.5ny (the nx/2 upper
    indices go to the top of the padded array: retained modes -nx/2 .. nx/2-1), real(ifft) on the padded grid (scaled
    by nxe nye/(nx ny)), product, fft, extraction of the same nx x ny modes, scaled back."""
    kx = ps_wavenumbers(nx, dx)
    ky = kx
    j1f = 1j * wf * kx[:, None] / k2
    j4f = 1j * wf * kx[:, None]
    j2f = 1j * wf * ky[None, :]
    j3f = 1j * wf * ky[None, :] / k2
    nxe, nye = int(1.5 * nx), int(1.5 * ny)
    hx, hy = nx // 2, ny // 2

    def pad(a):
        p = np.zeros((nxe, nye), dtype=np.complex128)
        p[:hx, :hy] = a[:hx, :hy]
        p[nxe - hx:, :hy] = a[hx:, :hy]
        p[:hx, nye - hy:] = a[:hx, hy:]
        p[nxe - hx:, nye - hy:] = a[hx:, hy:]
        return p

    j1, j2, j3, j4 = (np.real(np.fft.ifft2(pad(a) * (nxe * nye) / (nx * ny))) for a in (j1f, j2f, j3f, j4f))
    jacpf = np.fft.fft2(j1 * j2 - j3 * j4)
    jf = np.zeros((nx, ny), dtype=np.complex128)
    jf[:hx, :hy] = jacpf[:hx, :hy]
    jf[hx:, :hy] = jacpf[nxe - hx:, :hy]
    jf[:hx, hy:] = jacpf[:hx, nye - hy:]
    jf[hx:, hy:] = jacpf[nxe - hx:, nye - hy:]
    return jf * (nx * ny) / (nxe * nye)


def ps_numerical(rule, nx, ny, nt, dx, dy, dt, re, wn, freq=0, snapshot=None):
    """pseudospectral_23_rule.jl:13-89 / pseudospectral_32_rule.jl:13-89 (identical time loops): RK3 for the
    pseudo-spectral Jacobian, Crank-Nicolson per mode for the diffusion.  rule = 23 or 32.  Returns real(ifft(wnf)) with
    the periodic duplicates, (nx+1) x (ny+1) (the reference returns the field of its last snapshot, which is the final
    one whenever nt is a multiple of nt ÷ ns, as in its own configuration); wn (ghosted) is only read."""
    assert nx == ny and rule in (23, 32)
    jacobian = ps23_jacobian if rule == 23 else ps32_jacobian
    k2 = wavespace(nx, ny, dx, dy)
    wnf = np.fft.fft2(wn[1:nx + 1, 1:ny + 1].astype(np.complex128))
    wnf[0, 0] = 0.
    a1, a2, a3 = 8. / 15., 2. / 15., 1. / 3.
    g1, g2, g3 = 8. / 15., 5. / 12., 3. / 4.
    r2, r3 = -17. / 60., -5. / 12.
    z = .5 * dt * k2 / re
    d1, d2, d3 = a1 * z, a2 * z, a3 * z

    def field(wf):
        ut = np.empty((nx + 1, ny + 1), order="F")
        ut[:nx, :ny] = np.real(np.fft.ifft2(wf))
        ut[nx, :] = ut[0, :]
        ut[:, ny] = ut[:, 0]
        return ut

    for k in range(1, nt + 1):
        jnf = jacobian(nx, ny, dx, dy, wnf, k2)
        w1f = ((1. - d1) / (1. + d1)) * wnf + (g1 * dt * jnf) / (1. + d1)
        w1f[0, 0] = 0.
        j1f = jacobian(nx, ny, dx, dy, w1f, k2)
        w2f = ((1. - d2) / (1. + d2)) * w1f + (r2 * dt * jnf + g2 * dt * j1f) / (1. + d2)
        w2f[0, 0] = 0.
        j2f = jacobian(nx, ny, dx, dy, w2f, k2)
        wnf = ((1. - d3) / (1. + d3)) * w2f + (r3 * dt * j1f + g3 * dt * j2f) / (1. + d3)
        if snapshot is not None and freq > 0 and k % freq == 0:
            snapshot(k, field(wnf))
    return field(wnf)
