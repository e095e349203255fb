// vmk_pseudo32.cuh -- the pseudo-spectral solver with the 3/2 padding rule
// (21_NS2D_PseudoSpectral_32_Rule/pseudospectral_32_rule.jl:13-177).
//
// jacobian() (:95-177) zero-pads the four derivative spectra from nx x ny to M x M, M = 1.5 nx, takes real(ifft) on the
// padded grid, multiplies in real space and transforms back, keeping the nx x ny modes.  M = 3 L with L = nx/2 a power
// of two, so every M-point transform is split radix-3 into three L-point transforms -- and because real space is only
// ever used for a pointwise product, the three (per direction) interleaved sub-grids  x = 2 pi (3 px + rx) / M,
// rx = 0, 1, 2,  never have to be interleaved: real space is kept as 9 independent L x L sub-grids, transformed by the
// finite-difference path's own K1 / K3 (an L x L plan inside the plan), and the radix-3 butterflies become elementwise
// "fold" / "unfold" passes over the spectra with the twiddles w^m = exp(-2 pi i m / M):
//   inverse, one direction:  u[3p + r] = sum_{k' < L} ( sum_{k = k' mod L} U[k] w^{-k r} ) exp(+2 pi i k' p / L)
//   forward, one direction:  X[k]      = sum_r w^{k r} ( sum_p x[3p + r] exp(-2 pi i (k mod L) p / L) )
// The retained modes are k in -L .. L-1 per direction (:137-155) -- not a symmetric set -- and real(ifft(.)) keeps the
// Hermitian part, exactly as in the 2/3 rule (vmk_pseudo.cuh).  The state is therefore held on the symmetric set
//   S[kx][ky],  kx = 0 .. L,  ky = -L .. L  (column c = ky + L),   S = 2 x wf[kx,ky] if (kx,ky) is a retained mode,
//                                                                  else 2 x conj(wf[-kx,-ky])  ((L,-L): neither, 0),
// which the per-mode RK3 / Crank-Nicolson update (:41-66; real, even coefficients) maps onto itself, with the weights
// Dx, Dy of vmk_pseudo.cuh built from the indicator of the retained set.  Unlike the 2/3 rule the Nyquist lines carry
// an O(1) non-Hermitian part here (jf[-L, ky] is taken from the padded spectrum's mode -L alone, :171-174); the final
// field real(ifft(wnf)) on the nx x ny grid (:71) folds +L and -L back together (p32_final_body).
//
// One stage = E1 spectra -> KX (12 (L+1) inverse L-point rows, in place) -> E2 fold along i -> 4 x K3_L (each over the
// 9 sub-grids of one field: 9L rows of L points) -> product -> K1_L (9L rows) -> E3 unfold along i -> KX (3 (L+1)
// forward rows, in place) -> E4 unfold along j + mode update: 12 launches.
// Every pass works in natural index order on global memory: a deliberately plain first version (many passes over the
// spectra, scattered stores in KX) whose arithmetic was fixed first as a numpy model (tests/models/ps32_model.py, 3e-16
// against the literal restatement of the script with white-noise input); fusing E1/E4 into the row transforms the way
// KP does for the 2/3 rule is the obvious next step.
#pragma once
#include "vmk_kernels.cuh"

namespace vmk {

// ---- batched complex FFT along contiguous rows, natural order in and out (in place allowed) ----------------------
struct KXArgs {
  const double2* in;   // [nrows][N]
  double2* out;        // [nrows][N]
  const double2* tw;
  int nrows;
  int inverse;         // 0: sum x exp(-2 pi i k n / N), 1: sum X exp(+2 pi i k n / N) (both unnormalised)
};

template <class C>
VMK_HD void kx_body(const Ctx& c, const KXArgs& a) {
  using F = Fft<C>;
  constexpr int N = C::N, E = C::E, T = C::T, P = C::P;
  constexpr int bl = C::bits(P - 1), rl = 1 << bl;
  static_assert(!C::SPLIT, "kx_body uses the plain exchange buffer");
  double2* tw = F::tables(c.smem);
  F::load_tables(c, tw, a.tw);
  c.sync();
  const int g = c.tid / T, t = c.tid % T;
  double2* sm = F::xbuf(c.smem, g);
  const int nblocks = (a.nrows + C::FPC - 1) / C::FPC;
  for (int rb = c.bid; rb < nblocks; rb += c.nblk) {
    const int row = rb * C::FPC + g;
    const bool active = row < a.nrows;
    // a transform slot past the last row redoes the last row and discards the result: its loads and barriers are then
    // unconditional (no thread-varying branch around the barriers for a compiler to specialise), only the store is not
    const size_t roff = (size_t)(active ? row : a.nrows - 1) * N;
    double2 v[E];
    if (a.inverse) {
      static_for<0, E>([&](auto e_) {
        constexpr int e = decltype(e_)::value, u = e / rl, p = e % rl;
        const int k = F::k_of_pos(((t + T * u) << bl) | p);
        v[e] = a.in[roff + k];
      });
      c.sync();  // the previous row's last exchange has been read everywhere
      F::inverse(c, v, sm, tw, t);
      if (active) {
        static_for<0, E>([&](auto e_) {
          constexpr int e = decltype(e_)::value;
          a.out[roff + F::template own_pos<e>(t)] = v[e];
        });
      }
    } else {
      static_for<0, E>([&](auto e_) {
        constexpr int e = decltype(e_)::value;
        v[e] = a.in[roff + F::template own_pos<e>(t)];
      });
      c.sync();
      F::forward(c, v, sm, tw, t);
      if (active) {
        static_for<0, E>([&](auto e_) {
          constexpr int e = decltype(e_)::value, u = e / rl, p = e % rl;
          a.out[roff + F::k_of_pos(((t + T * u) << bl) | p)] = v[e];
        });
      }
    }
  }
}

// ---- elementwise passes -------------------------------------------------------------------------------------------
struct P32Args {
  int L;                 // nx / 2; M = 3 L, N = 2 L
  const double2* twM;    // [3L] w^m = exp(-2 pi i m / M)
  const double* ksq;     // [2L+1] by column c = k + L: k[k]^2 (k[0] = eps, :107; +-L: (L hx)^2)
  const double* mp;      //        1 if k is retained (-L .. L-1)
  const double* mm;      //        mp(-k)
  const double* cc;      //        k[k] mp(k) / 2
  const double* dd;      //        -k[-k] mp(-k) / 2
  double2* S;            // [L+1][2L+1] state
  double2* J;            // [L+1][2L+1] previous stage's jf
  double2* Y;            // [4][L+1][3][L]: E1 output, transformed in place by KX, E2 input
  double2* VF;           // [4][L/2][9 sub-grids (3 rx + ry)][L]: K3_L input of one launch per field (9L rows of L points)
  const double2* T9;     // [L/2][9 sub-grids][L]: K1_L output of one launch
  double2* Pi;           // [L+1][3][L]: E3 output, transformed in place by KX, E4 input
  const double2* Xn;     // init:  [L][N] forward-j of K1_N(w0) (row 0 packed), natural order
  double2* Un;           // final: [L][N] N-grid half spectrum (row 0 packed), natural order
  double zfac;           // .5 dt / re
  double alpha, gdt, rdt;
  double scale;          // 1 / (2 N^2)
  int stage;
};

VMK_HD size_t p32_begin(const Ctx& c) { return (size_t)c.bid * kK5Threads + c.tid; }
VMK_HD size_t p32_stride(const Ctx& c) { return (size_t)c.nblk * kK5Threads; }

// E0: S <- the nx x ny spectrum of the initial field (pseudospectral_32_rule.jl:24-27), J <- 0
VMK_HD void p32_init_body(const Ctx& c, const P32Args& a) {
  const int L = a.L, N = 2 * L, W = 2 * L + 1;
  const size_t total = (size_t)(L + 1) * W;
  for (size_t q = p32_begin(c); q < total; q += p32_stride(c)) {
    const int kx = (int)(q / W), col = (int)(q % W), ky = col - L;
    const int n = (ky + N) % N, nm = (N - ky) % N;
    double2 s;
    if (kx >= 1 && kx < L) {
      s = a.Xn[(size_t)kx * N + n];
    } else {
      // packed row X = A + i B (A: kx = 0, B: kx = -L): A[ky] = (X[ky] + conj X[-ky]) / 2, B[ky] = (X[ky] - conj X[-ky]) / 2i
      const double2 x = a.Xn[n], xm = a.Xn[nm];
      if (kx == 0) {
        s = mk2(.5 * (x.x + xm.x), .5 * (x.y - xm.y));
        if (col == 2 * L) s = cconj(s);
      } else {
        // S[L][ky] = conj(B[-ky]),  B[-ky] = (X[-ky] - conj X[ky]) / 2i
        const double2 bm = mk2(.5 * (xm.y + x.y), .5 * (x.x - xm.x));
        s = cconj(bm);
        if (col == 0) s = mk2(0.0, 0.0);
      }
    }
    if (kx == 0 && col == L) s = mk2(0.0, 0.0);  // wf[1,1] = 0 (:27)
    a.S[q] = s;
    a.J[q] = mk2(0.0, 0.0);
  }
}

// the four derivative spectra of one mode (:113-122 with the Hermitian-part weights), scaled by 1 / (2 N^2)
VMK_HD void p32_mode(const P32Args& a, int kxc, int col, double2 s, double2* g) {
  const double k2 = ld_ro(a.ksq + kxc) + ld_ro(a.ksq + col);
  const double dx = ld_ro(a.cc + kxc) * ld_ro(a.mp + col) + ld_ro(a.dd + kxc) * ld_ro(a.mm + col);
  const double dy = ld_ro(a.mp + kxc) * ld_ro(a.cc + col) + ld_ro(a.mm + kxc) * ld_ro(a.dd + col);
  const double fx = dx * a.scale, fy = dy * a.scale, r = 1.0 / k2;
  const double2 is = mk2(-s.y, s.x);  // i S
  g[0] = cscale(is, fx * r);          // j1f = i kx wf / k2
  g[1] = cscale(is, fy);              // j2f = i ky wf
  g[2] = cscale(is, fy * r);          // j3f = i ky wf / k2
  g[3] = cscale(is, fx);              // j4f = i kx wf
}

// E1: Y[q][kx][ry][ky'] = sum over ky = ky' mod L of G_q[kx][ky] w^{-ky ry}
VMK_HD void p32_spectra_body(const Ctx& c, const P32Args& a) {
  const int L = a.L, M = 3 * L, W = 2 * L + 1;
  const size_t total = (size_t)(L + 1) * L;
  for (size_t id = p32_begin(c); id < total; id += p32_stride(c)) {
    const int kx = (int)(id / L), kp = (int)(id % L), kxc = kx + L;
    const double2* srow = a.S + (size_t)kx * W;
    double2 ga[4], gb[4], ge[4];
    p32_mode(a, kxc, kp + L, srow[kp + L], ga);  // ky = ky'
    p32_mode(a, kxc, kp, srow[kp], gb);          // ky = ky' - L
    if (kp == 0) p32_mode(a, kxc, 2 * L, srow[2 * L], ge);  // ky = +L
#pragma unroll
    for (int ry = 0; ry < 3; ry++) {
      const double2 wa = ld_ro2(a.twM + (M - kp * ry) % M);      // w^{-ky' ry}
      const double2 wb = ld_ro2(a.twM + (L - kp) * ry);          // w^{-(ky'-L) ry}
      const double2 we = ld_ro2(a.twM + (M - L * ry) % M);       // w^{-L ry}
#pragma unroll
      for (int q = 0; q < 4; q++) {
        double2 y = cadd(cmul(ga[q], wa), cmul(gb[q], wb));
        if (kp == 0) y = cadd(y, cmul(ge[q], we));
        a.Y[(((size_t)q * (L + 1) + kx) * 3 + ry) * L + kp] = y;
      }
    }
  }
}

// E1 fused into the inverse row transform (opt-in, set_option "ps32_fuse"): row = (q, kx, ry); the transform's inputs
// Y[q][kx][ry][ky'] are computed from S in its load stage instead of being written by p32_spectra_body and read back --
// one write and one read of the 12 (L+1) L-point spectra less per stage.  Output layout as kx_body's (E2 unchanged).
struct KXSArgs {
  P32Args a;
  const double2* tw;  // the FFT engine's tables
  int nrows;          // 12 (L + 1)
};

template <class C>
VMK_HD void kxs_body(const Ctx& c, const KXSArgs& ka) {
  using F = Fft<C>;
  constexpr int N = C::N, E = C::E, T = C::T, P = C::P;
  constexpr int bl = C::bits(P - 1), rl = 1 << bl;
  static_assert(!C::SPLIT, "kxs_body uses the plain exchange buffer");
  const P32Args& a = ka.a;
  double2* tw = F::tables(c.smem);
  F::load_tables(c, tw, ka.tw);
  c.sync();
  const int g = c.tid / T, t = c.tid % T;
  double2* sm = F::xbuf(c.smem, g);
  const int L = a.L, M = 3 * L, W = 2 * L + 1;  // L == N
  const int nblocks = (ka.nrows + C::FPC - 1) / C::FPC;
  for (int rb = c.bid; rb < nblocks; rb += c.nblk) {
    const int slot = rb * C::FPC + g;
    const bool active = slot < ka.nrows;
    const int row = active ? slot : ka.nrows - 1;  // spare slots redo the last row, only the store is conditional
    const int ry = row % 3, kx = (row / 3) % (L + 1), q = row / (3 * (L + 1));
    const int kxc = kx + L;
    const double2* srow = a.S + (size_t)kx * W;
    const bool xdir = (q == 0 || q == 3), div = (q == 0 || q == 2);
    const double cxr = ld_ro(a.cc + kxc), dxr = ld_ro(a.dd + kxc), mpr = ld_ro(a.mp + kxc), mmr = ld_ro(a.mm + kxc);
    const double kxx = ld_ro(a.ksq + kxc);
    // i S D(kx, col) [/ k2] / (2 N^2) for one mode column (p32_mode restricted to the row's q)
    auto mode = [&](int col) {
      const double2 s = srow[col];
      double f = xdir ? cxr * ld_ro(a.mp + col) + dxr * ld_ro(a.mm + col) : mpr * ld_ro(a.cc + col) + mmr * ld_ro(a.dd + col);
      f = f * a.scale;
      if (div) f = f * (1.0 / (kxx + ld_ro(a.ksq + col)));
      return mk2(-s.y * f, s.x * f);
    };
    double2 v[E];
    static_for<0, E>([&](auto e_) {
      constexpr int e = decltype(e_)::value, u = e / rl, p = e % rl;
      const int kp = F::k_of_pos(((t + T * u) << bl) | p);
      double2 y = cadd(cmul(mode(kp + L), ld_ro2(a.twM + (M - kp * ry) % M)),   // ky = ky'
                       cmul(mode(kp), ld_ro2(a.twM + (L - kp) * ry)));          // ky = ky' - L
      if (kp == 0) y = cadd(y, cmul(mode(2 * L), ld_ro2(a.twM + (M - L * ry) % M)));  // ky = +L
      v[e] = y;
    });
    c.sync();  // the previous row's last exchange has been read everywhere
    F::inverse(c, v, sm, tw, t);
    if (active) {
      static_for<0, E>([&](auto e_) {
        constexpr int e = decltype(e_)::value;
        a.Y[(size_t)row * N + F::template own_pos<e>(t)] = v[e];
      });
    }
  }
}

// E1 + E2 fused into the inverse row transform (opt-in, set_option "ps32_fuse" = 2).  The fold along i is linear in kx
// (with a conjugation on the mirrored term), so it commutes with the transform along j: conj(V[k][py]) is the inverse
// transform of conj(Y[k][-ky' mod L]).  K3's input is therefore  VF = ifft_j(YF)  with, for kx' >= 1,
//   YF[kx'][ky'] = Y[kx'][ky'] w^{-kx' rx} + conj(Y[L-kx'][-ky']) w^{(L-kx') rx}
// and for the packed row (real parts in E2 = Hermitian parts here)
//   YF[0][ky'] = (Y[0][ky'] + conj Y[0][-ky'])/2 + Y[L][ky'] w^{-L rx} + conj(Y[L][-ky']) w^{L rx}
//                + i ( Y[L/2][ky'] w^{-(L/2) rx} + conj(Y[L/2][-ky']) w^{(L/2) rx} ),
// Y being p32_spectra_body's folded spectra, all computed from S in the load stage (tests/models/ps32_model.py,
// spectra_folded: 2e-16 against fold-after-transform).  Row = (q, rx, ry, kx'): 36 (L/2) transforms instead of
// 12 (L+1), but the spectra are never written or re-read and E2 disappears.
template <class C>
VMK_HD void kxf_body(const Ctx& c, const KXSArgs& ka) {
  using F = Fft<C>;
  constexpr int E = C::E, T = C::T, P = C::P;
  constexpr int bl = C::bits(P - 1), rl = 1 << bl;
  static_assert(!C::SPLIT, "kxf_body uses the plain exchange buffer");
  const P32Args& a = ka.a;
  double2* tw = F::tables(c.smem);
  F::load_tables(c, tw, ka.tw);
  c.sync();
  const int g = c.tid / T, t = c.tid % T;
  double2* sm = F::xbuf(c.smem, g);
  const int L = a.L, M = 3 * L, W = 2 * L + 1, h = L / 2;  // L == N
  const int nblocks = (ka.nrows + C::FPC - 1) / C::FPC;
  for (int rb = c.bid; rb < nblocks; rb += c.nblk) {
    const int slot = rb * C::FPC + g;
    const bool active = slot < ka.nrows;
    const int row = active ? slot : ka.nrows - 1;  // spare slots redo the last row, only the store is conditional
    const int kxp = row % h, ry = (row / h) % 3, rx = (row / (3 * h)) % 3, q = row / (9 * h);
    const bool xdir = (q == 0 || q == 3), div = (q == 0 || q == 2);
    // Y[q][kx][ry][kp] of p32_spectra_body
    auto yval = [&](int kx, int kp) {
      const int kxc = kx + L;
      const double2* srow = a.S + (size_t)kx * W;
      const double cxr = ld_ro(a.cc + kxc), dxr = ld_ro(a.dd + kxc), mpr = ld_ro(a.mp + kxc), mmr = ld_ro(a.mm + kxc);
      const double kxx = ld_ro(a.ksq + kxc);
      auto mode = [&](int col) {
        const double2 s = srow[col];
        double f = xdir ? cxr * ld_ro(a.mp + col) + dxr * ld_ro(a.mm + col) : mpr * ld_ro(a.cc + col) + mmr * ld_ro(a.dd + col);
        f = f * a.scale;
        if (div) f = f * (1.0 / (kxx + ld_ro(a.ksq + col)));
        return mk2(-s.y * f, s.x * f);
      };
      double2 y = cadd(cmul(mode(kp + L), ld_ro2(a.twM + (M - kp * ry) % M)), cmul(mode(kp), ld_ro2(a.twM + (L - kp) * ry)));
      if (kp == 0) y = cadd(y, cmul(mode(2 * L), ld_ro2(a.twM + (M - L * ry) % M)));
      return y;
    };
    double2 v[E];
    static_for<0, E>([&](auto e_) {
      constexpr int e = decltype(e_)::value, u = e / rl, p = e % rl;
      const int kp = F::k_of_pos(((t + T * u) << bl) | p);
      const int km = (L - kp) % L;  // -ky' mod L
      if (kxp >= 1) {
        v[e] = cadd(cmul(yval(kxp, kp), ld_ro2(a.twM + (M - kxp * rx) % M)),
                    cmul(cconj(yval(L - kxp, km)), ld_ro2(a.twM + (L - kxp) * rx)));
      } else {
        const double2 wl = ld_ro2(a.twM + (M - L * rx) % M), wh = ld_ro2(a.twM + (M - h * rx) % M);
        const double2 y0 = cadd(cscale(cadd(yval(0, kp), cconj(yval(0, km))), .5),
                                cadd(cmul(yval(L, kp), wl), cmul(cconj(yval(L, km)), cconj(wl))));
        const double2 yh = cadd(cmul(yval(h, kp), wh), cmul(cconj(yval(h, km)), cconj(wh)));
        v[e] = mk2(y0.x - yh.y, y0.y + yh.x);  // y0 + i yh
      }
    });
    c.sync();  // the previous row's last exchange has been read everywhere
    F::inverse(c, v, sm, tw, t);
    if (active) {
      double2* dst = a.VF + (((size_t)q * h + kxp) * 9 + (rx * 3 + ry)) * L;
      static_for<0, E>([&](auto e_) {
        constexpr int e = decltype(e_)::value;
        dst[F::template own_pos<e>(t)] = v[e];
      });
    }
  }
}

// E2: fold along i.  VF[q][rx][ry][kx'][py] = sum over kx = kx' mod L (Hermitian completion in kx) of V w^{-kx rx};
// row kx' = 0 packs the (real) kx' = 0 and kx' = L/2 lines as K3 expects
VMK_HD void p32_fold_body(const Ctx& c, const P32Args& a) {
  const int L = a.L, M = 3 * L, h = L / 2;
  const size_t total = (size_t)4 * 3 * h * L;
  for (size_t id = p32_begin(c); id < total; id += p32_stride(c)) {
    const int py = (int)(id % L);
    size_t r = id / L;
    const int kp = (int)(r % h);
    r /= h;
    const int ry = (int)(r % 3), q = (int)(r / 3);
    auto V = [&](int kx) { return a.Y[(((size_t)q * (L + 1) + kx) * 3 + ry) * L + py]; };
    for (int rx = 0; rx < 3; rx++) {
      double2 y;
      if (kp == 0) {
        const double2 v0 = V(0), vl = V(L), vh = V(h);
        const double2 tl = cmul(vl, ld_ro2(a.twM + (M - L * rx) % M));  // V[L] w^{-L rx}; its conjugate is the -L term
        const double2 th = cmul(vh, ld_ro2(a.twM + (M - h * rx) % M));  // V[L/2] w^{-(L/2) rx}; conjugate: kx = -L/2
        y = mk2(v0.x + (tl.x + tl.x), th.x + th.x);
      } else {
        const double2 v1 = V(kp), v2 = cconj(V(L - kp));
        y = cadd(cmul(v1, ld_ro2(a.twM + (M - kp * rx) % M)), cmul(v2, ld_ro2(a.twM + (L - kp) * rx)));
      }
      a.VF[(((size_t)q * h + kp) * 9 + (rx * 3 + ry)) * L + py] = y;
    }
  }
}

// E3: unfold along i.  Pi[kx][ry][py] = (4/9) sum_rx w^{kx rx} T_rx[kx mod L][ry][py]   (N^2 / M^2 = 4/9, :176)
VMK_HD void p32_unfold_body(const Ctx& c, const P32Args& a) {
  const int L = a.L, h = L / 2;
  const size_t total = (size_t)(L + 1) * 3 * L;
  for (size_t id = p32_begin(c); id < total; id += p32_stride(c)) {
    const int py = (int)(id % L);
    const size_t r = id / L;
    const int ry = (int)(r % 3), kx = (int)(r / 3), kk = kx % L;
    double2 acc = mk2(0.0, 0.0);
    for (int rx = 0; rx < 3; rx++) {
      const double2* blk = a.T9 + (size_t)(rx * 3 + ry) * L + py;  // [kx'][9 sub-grids][py]
      const size_t pitch = (size_t)9 * L;
      double2 tv;
      if (kk == 0) {
        tv = mk2(blk[0].x, 0.0);
      } else if (kk == h) {
        tv = mk2(blk[0].y, 0.0);
      } else if (kk < h) {
        tv = blk[(size_t)kk * pitch];
      } else {
        tv = cconj(blk[(size_t)(L - kk) * pitch]);
      }
      acc = cadd(acc, cmul(tv, ld_ro2(a.twM + kx * rx)));
    }
    a.Pi[id] = cscale(acc, 4.0 / 9.0);
  }
}

// E4: unfold along j, jf[kx][ky] = sum_ry w^{ky ry} Q[kx][ry][ky mod L], and the RK3 / Crank-Nicolson update (:41-66)
VMK_HD void p32_update_body(const Ctx& c, const P32Args& a) {
  const int L = a.L, M = 3 * L, W = 2 * L + 1;
  const size_t total = (size_t)(L + 1) * W;
  for (size_t id = p32_begin(c); id < total; id += p32_stride(c)) {
    const int kx = (int)(id / W), col = (int)(id % W), ky = col - L;
    const int km = (ky + 2 * L) % L;
    double2 pf = mk2(0.0, 0.0);
    for (int ry = 0; ry < 3; ry++) {
      const int e = ((ky * ry) % M + M) % M;
      pf = cadd(pf, cmul(a.Pi[((size_t)kx * 3 + ry) * L + km], ld_ro2(a.twM + e)));
    }
    const double k2 = ld_ro(a.ksq + kx + L) + ld_ro(a.ksq + col);
    const double d = a.alpha * (a.zfac * k2);
    const double g = 1.0 / (1.0 + d), cf = (1.0 - d) * g;
    double2 y = cscale(pf, a.gdt);
    if (a.stage >= 2) {
      const double2 jp = a.J[id];
      y = mk2(fma_(a.rdt, jp.x, y.x), fma_(a.rdt, jp.y, y.y));
    }
    if (a.stage == 1 || a.stage == 2) a.J[id] = pf;
    const double2 s = a.S[id];
    double2 sn = mk2(fma_(cf, s.x, g * y.x), fma_(cf, s.y, g * y.y));
    if (kx == L && col == 0) sn = mk2(0.0, 0.0);                  // (L, -L): neither it nor its mirror is retained
    if (kx == 0 && col == L && a.stage != 3) sn = mk2(0.0, 0.0);  // w1f[1,1] = w2f[1,1] = 0 (:47,58)
    a.S[id] = sn;
  }
}

// E5: the nx x ny half spectrum of real(ifft(wnf)) (:71): the Hermitian part on the N grid, where +L and -L coincide
VMK_HD void p32_final_body(const Ctx& c, const P32Args& a) {
  const int L = a.L, N = 2 * L, W = 2 * L + 1;
  const size_t total = (size_t)L * N;
  const double inv = 1.0 / ((double)N * (double)N);
  for (size_t id = p32_begin(c); id < total; id += p32_stride(c)) {
    const int kx = (int)(id / N), n = (int)(id % N), ky = n < L ? n : n - N;  // ky in -L .. L-1
    const int col = ky + L, colm = L - ky;
    double2 u;
    if (kx >= 1) {
      const double2* s = a.S + (size_t)kx * W;
      u = ky != -L ? cscale(s[col], .5) : cscale(cadd(s[0], s[2 * L]), .25);
    } else {
      const double2* s0 = a.S;
      const double2* sl = a.S + (size_t)L * W;
      double2 u0, ul;
      if (ky != -L) {
        u0 = cscale(cadd(s0[col], cconj(s0[colm])), .25);
        ul = cscale(cadd(cconj(sl[colm]), sl[col]), .25);
      } else {
        u0 = mk2(.5 * s0[0].x, 0.0);
        ul = mk2(.5 * sl[2 * L].x, 0.0);
      }
      u = mk2(u0.x - ul.y, u0.y + ul.x);  // U[0] + i U[Nyquist]
    }
    a.Un[id] = cscale(u, inv);
  }
}

}  // namespace vmk
