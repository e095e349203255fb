"""numpy model of the 3/2-rule pseudo-spectral stage as the CUDA path decomposes it (csrc/vmk_pseudo32.cuh): the
1.5N-point transforms are split radix-3 into three N/2-point transforms per direction ("sub-grids" rx, ry = 0..2 of the
padded grid, x = 2 pi (3 px + rx) / M), real space is kept sub-grid-major (the pointwise product does not care), and
every fold / unfold / update step is a natural-order elementwise pass.  Development aid: each function below is the
specification of one kernel (E0..E5); tests compare the CUDA path with oracle_np.ps_numerical, not with this file."""
import numpy as np


class Model32:
    def __init__(self, n, dx, dt, re, eps=1e-6):
        self.N, self.L, self.M = n, n // 2, 3 * n // 2
        N, L, M = self.N, self.L, self.M
        self.dt, self.re = dt, re
        hx = 2 * np.pi / (n * dx)
        k = np.arange(-L, L + 1)                      # symmetric mode set, column c = k + L
        kap = hx * k.astype(float)
        kap[L] = eps                                  # kx[1] = eps
        self.ksq = kap**2
        mp = np.ones(2 * L + 1); mp[2 * L] = 0        # 1_I per direction: k in [-L, L-1]
        mm = mp[::-1].copy()                          # mp(-k)
        self.mp, self.mm = mp, mm
        self.cc = kap * mp / 2
        self.dd = -kap[::-1] * mm / 2                 # -kap(-k) mm(k) / 2
        self.tw = np.exp(-2j * np.pi * np.arange(M) / M)   # w^m
        self.k = k

    def w(self, e):                                   # w^e for integer arrays e (any sign)
        return self.tw[np.mod(e, self.M)]

    # E0: N-grid K1 + forward-j output Xn [L rows (row 0 packed)][N] (2 x spectrum) -> S [L+1][2L+1]
    def init_state(self, w0):
        N, L = self.N, self.L
        F = 2 * np.fft.fft(w0, axis=0)                # K1: 2 x DFT along i, rows kx
        T = F[:L, :].copy()
        T[0, :] = F[0, :].real + 1j * F[L, :].real    # packed row (values are real before the j transform)
        Xn = np.fft.fft(T, axis=1)                    # KX_fwd size N along j
        S = np.zeros((L + 1, 2 * L + 1), complex)
        ky = self.k
        idx = np.mod(ky, N)
        X0, X0m = Xn[0, idx], Xn[0, np.mod(-ky, N)]
        A = (X0 + np.conj(X0m)) / 2
        B = (X0 - np.conj(X0m)) / 2j
        S[1:L, :] = Xn[1:L, idx]
        S[0, :] = A
        S[0, 2 * L] = np.conj(A[0])
        Bm = (X0m - np.conj(X0)) / 2j                 # B at -ky
        S[L, :] = np.conj(Bm)
        S[L, 0] = 0
        S[0, L] = 0                                   # wf[1,1] = 0
        return S

    # E1: S -> Yj[q][kx][ry][ky'] (folded along j, scaled)
    def spectra(self, S):
        L, M = self.L, self.M
        kx = np.arange(0, L + 1) + L                  # columns of the tables for kx = 0..L
        k2 = self.ksq[kx][:, None] + self.ksq[None, :]
        Dx = self.cc[kx][:, None] * self.mp[None, :] + self.dd[kx][:, None] * self.mm[None, :]
        Dy = self.mp[kx][:, None] * self.cc[None, :] + self.mm[kx][:, None] * self.dd[None, :]
        sc = 1 / (2 * self.N**2)
        G = [1j * Dx * S * sc / k2, 1j * Dy * S * sc, 1j * Dy * S * sc / k2, 1j * Dx * S * sc]
        Y = np.zeros((4, L + 1, 3, L), complex)
        kyp = np.arange(L)
        for q in range(4):
            for ry in range(3):
                # ky = ky' (column ky'+L), ky = ky'-L (column ky'), and ky = +L (column 2L) for ky' = 0
                Y[q, :, ry, :] = G[q][:, kyp + L] * self.w(-kyp * ry)[None, :] + G[q][:, kyp] * self.w(-(kyp - L) * ry)[None, :]
                Y[q, :, ry, 0] += G[q][:, 2 * L] * self.w(-L * ry)
        return Y

    # E2: Vj[q][kx][ry][py] -> K3_L input VF[q][rx][ry][kx' 0..L/2-1 (row 0 packed)][py]
    def fold_i(self, V):
        L = self.L
        h = L // 2
        VF = np.zeros((4, 3, 3, h, L), complex)
        for rx in range(3):
            kxp = np.arange(1, h)
            Yi = (V[:, kxp, :, :] * self.w(-kxp * rx)[None, :, None, None] +
                  np.conj(V[:, L - kxp, :, :]) * self.w(-(kxp - L) * rx)[None, :, None, None])
            VF[:, rx, :, 1:, :] = np.transpose(Yi, (0, 2, 1, 3))
            y0 = V[:, 0] + V[:, L] * self.w(-L * rx) + np.conj(V[:, L]) * self.w(L * rx)          # kx' = 0 (real)
            yh = V[:, h] * self.w(-h * rx) + np.conj(V[:, h]) * self.w(h * rx)                      # kx' = L/2 (real)
            VF[:, rx, :, 0, :] = y0.real + 1j * yh.real
        return VF

    # E1 + E2 as ONE spectral-space pass (design for the next fusion step, not yet a kernel): the fold along i is linear
    # in kx (with a conjugation on the mirrored term), so it commutes with the transform along j -- conj(V[k][py]) is the
    # inverse transform of conj(Y[k][-ky' mod L]) -- and K3's input can be produced as  VF = ifft_j(YF)  with YF computed
    # from S directly.  The packed row takes real parts in E2; in spectral space Re(v) is (Y[ky'] + conj Y[-ky'])/2.
    def spectra_folded(self, S):
        L = self.L
        h = L // 2
        Y = self.spectra(S)                                   # [q][kx][ry][ky']
        Ym = np.conj(Y[..., np.mod(-np.arange(L), L)])        # conj(Y[..][-ky' mod L]): spectrum of conj(V)
        YF = np.zeros((4, 3, 3, h, L), complex)
        for rx in range(3):
            kxp = np.arange(1, h)
            yi = (Y[:, kxp] * self.w(-kxp * rx)[None, :, None, None] +
                  Ym[:, L - kxp] * self.w(-(kxp - L) * rx)[None, :, None, None])
            YF[:, rx, :, 1:, :] = np.transpose(yi, (0, 2, 1, 3))
            wl, wh = self.w(-L * rx), self.w(-h * rx)
            y0 = (Y[:, 0] + Ym[:, 0]) / 2 + Y[:, L] * wl + Ym[:, L] * np.conj(wl)
            yh = Y[:, h] * wh + Ym[:, h] * np.conj(wh)
            YF[:, rx, :, 0, :] = y0 + 1j * yh
        return YF

    @staticmethod
    def k3(VF):  # [..., L/2 (packed), L(py)] -> real [..., px, py]: out[px] = sum over the Hermitian-completed kx'
        h, L = VF.shape[-2], VF.shape[-1]
        full = np.zeros(VF.shape[:-2] + (L, L), complex)
        full[..., 0, :] = VF[..., 0, :].real
        full[..., h, :] = VF[..., 0, :].imag
        full[..., 1:h, :] = VF[..., 1:, :]
        full[..., h + 1:, :] = np.conj(VF[..., 1:, :][..., ::-1, :])
        return np.real(np.fft.ifft(full, axis=-2) * L)

    @staticmethod
    def k1(f):   # real [..., px, py] -> T [..., L/2 (packed), py] = 2 x DFT along px
        L = f.shape[-2]
        h = L // 2
        F = 2 * np.fft.fft(f, axis=-2)
        T = F[..., :h, :].copy()
        T[..., 0, :] = F[..., 0, :].real + 1j * F[..., h, :].real
        return T

    # E3: T[rx][ry][kx'][py] -> Pi[kx 0..L][ry][py]
    def unfold_i(self, T):
        L = self.L
        h = L // 2
        Tf = np.zeros((3, 3, L, L), complex)
        Tf[:, :, 0, :] = T[:, :, 0, :].real
        Tf[:, :, h, :] = T[:, :, 0, :].imag
        Tf[:, :, 1:h, :] = T[:, :, 1:, :]
        Tf[:, :, h + 1:, :] = np.conj(T[:, :, 1:, :][:, :, ::-1, :])
        kx = np.arange(L + 1)
        Pi = np.zeros((L + 1, 3, L), complex)
        for rx in range(3):
            Pi += (4. / 9.) * self.w(kx * rx)[:, None, None] * np.transpose(Tf[rx][:, np.mod(kx, L), :], (1, 0, 2))
        return Pi

    # E4: Qj[kx][ry][ky'] -> Pf[kx][ky]; update S, J
    def update(self, S, Jp, Q, stage):
        L = self.L
        alpha = [0, 8 / 15, 2 / 15, 1 / 3][stage]
        gam = [0, 8 / 15, 5 / 12, 3 / 4][stage]
        rho = [0, 0, -17 / 60, -5 / 12][stage]
        ky = self.k
        Pf = np.zeros((L + 1, 2 * L + 1), complex)
        for ry in range(3):
            Pf += self.w(ky * ry)[None, :] * Q[:, ry, np.mod(ky, L)]
        kx = np.arange(0, L + 1) + L
        k2 = self.ksq[kx][:, None] + self.ksq[None, :]
        d = alpha * (.5 * self.dt * k2 / self.re)
        Sn = ((1 - d) / (1 + d)) * S + (rho * self.dt * Jp + gam * self.dt * Pf) / (1 + d)
        Sn[L, 0] = 0
        if stage != 3:
            Sn[0, L] = 0
        return Sn, Pf

    def jacobian(self, S):
        Y = self.spectra(S)
        V = np.fft.ifft(Y, axis=-1) * self.L                 # KX_inv size L
        F = self.k3(self.fold_i(V))                          # [q][rx][ry][px][py]
        jac = F[0] * F[1] - F[2] * F[3]
        T = self.k1(jac)
        Pi = self.unfold_i(T)
        return np.fft.fft(Pi, axis=-1)                       # KX_fwd size L -> Q

    def step(self, S):
        J = np.zeros_like(S)
        for stage in (1, 2, 3):
            Q = self.jacobian(S)
            S, J = self.update(S, J, Q, stage)
        return S

    # E5: S -> N-grid packed half spectrum Un[L][N] (scaled 1/N^2), then KX_inv size N, K3_N
    def field(self, S):
        N, L = self.N, self.L
        U = np.zeros((L + 1, N), complex)                    # Herm_N part, rows kx = 0..L (L = Nyquist)
        ky = np.arange(-L + 1, L)
        c, cm = ky + L, -ky + L
        U[1:L, np.mod(ky, N)] = S[1:L, c] / 2
        U[1:L, L] = (S[1:L, 0] + S[1:L, 2 * L]) / 4
        U[0, np.mod(ky, N)] = (S[0, c] + np.conj(S[0, cm])) / 4
        U[0, L] = S[0, 0].real / 2
        U[L, np.mod(ky, N)] = (np.conj(S[L, cm]) + S[L, c]) / 4
        U[L, L] = S[L, 2 * L].real / 2
        Un = U[:L, :].copy()
        Un[0, :] = U[0, :] + 1j * U[L, :]
        Un /= N**2
        V = np.fft.ifft(Un, axis=1) * N                      # KX_inv size N along j
        return Model32.k3(V)                                 # K3_N: [px][py]


if __name__ == "__main__":
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    from oracle import oracle_np as o
    for n, amp in ((16, 1.), (32, 1.), (64, .05)):
        dx = 2 * np.pi / n
        rng = np.random.default_rng(1)
        x = dx * np.arange(n)
        X, Y = np.meshgrid(x, x, indexing="ij")
        w0 = np.exp(-np.pi * ((X - 3 * np.pi / 4)**2 + (Y - np.pi)**2)) + amp * rng.uniform(-1, 1, (n, n))
        wn = np.zeros((n + 2, n + 2), order="F")
        wn[1:n + 1, 1:n + 1] = w0
        nt, dt, re = 3, 1e-3, 1000.
        ref = o.ps_numerical(32, n, n, nt, dx, dx, dt, re, wn)[:n, :n]
        m = Model32(n, dx, dt, re)
        S = m.init_state(w0)
        for _ in range(nt):
            S = m.step(S)
        u = m.field(S)
        print(n, np.linalg.norm(u - ref) / np.linalg.norm(ref))
        # the one-pass form of E1 + E2 equals the two-pass form
        a = m.fold_i(np.fft.ifft(m.spectra(S), axis=-1) * m.L)
        b = np.fft.ifft(m.spectra_folded(S), axis=-1) * m.L
        print("   folded-spectra form vs fold after transform:", np.linalg.norm(a - b) / np.linalg.norm(a))
