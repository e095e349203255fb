// vmk_hybrid.cuh -- the spectral-space stage of the hybrid RK3 / Crank-Nicolson solver
// (20_NS2D_Hybrid_Solver/hybrid.jl:14-152).
//
// hybrid.jl keeps the vorticity in Fourier space.  Per RK3 stage it evaluates jacobian(wf) (hybrid.jl:96-152):
//   w = real(ifft(wf)), psi = real(ifft(wf / k2)), Arakawa J in real space, jf = fft(-J/3)
// and then updates every mode (hybrid.jl:40-66):
//   wf' = ((1 - d)/(1 + d)) wf + (rho dt jf_prev + gamma dt jf) / (1 + d),   d = alpha (dt/2) k2 / re.
// Here one stage is five launches that reuse the finite-difference path's kernels for everything in real space:
//   K3 (V_w -> w rows), K3 (V_s -> psi rows), K4 in mode 0 with 1/re = 0 (r = -J/3 exactly), K1 (rows of -J/3 -> T),
// and KH below, which works on one spectrum row (fixed kx, all ky) per CTA iteration:
//   forward FFT along j of the row of T (= the last pass of fft(-J/3))  ->  jf in registers
//   the Crank-Nicolson / RK3 update with wf and jf_prev read from global memory (W, J: [kx][e*T + t], the threads' own
//   register order, so no index permutation is ever materialised), wf' and jf written back in place
//   wf'/k2 -> inverse FFT along j -> V_s,   wf' -> inverse FFT along j -> V_w      (the first half of the NEXT jacobian)
// The state is the half spectrum kx < N/2 (w is real, and k2 is even in both indices including the eps quirk of
// Common.jl:196, so Hermitian symmetry is preserved exactly as in the finite-difference path); spectrum row 0 packs
// kx = 0 (real part) and kx = N/2 (imaginary part) like K2's, and is separated with the mirror trick of k2_body.
// Scaling: W and J hold 2 x the reference's unnormalised spectra (the factor of K1's real-pair unpack); the inverse
// side multiplies by 1/(2 N^2).
#pragma once
#include "vmk_kernels.cuh"

namespace vmk {

struct KHArgs {
  const double2* T;      // K1 output [N/2][N]: rows of fft_i(data)
  double2* W;            // vorticity spectrum [N/2][N] in register order (in/out)
  double2* J;            // previous stage's jf [N/2][N] in register order (in/out)
  double2* Vw;           // inverse-j of wf' [N/2][N] (K3 input, row-major)
  double2* Vs;           // inverse-j of wf'/k2
  const double2* tw;
  const double* ksq;     // [N]  kx[i]^2 (kx[1] = eps; Common.jl:189-196)
  const double* ksqperm; // ksq in the threads' register order: ksqperm[e*T + t] = ksq[k_of_pos(...)]
  double zfac;           // .5 dt / re                      hybrid.jl:34
  double alpha;          // alpha_s                         hybrid.jl:29
  double gdt, rdt;       // gamma_s dt, rho_s dt            hybrid.jl:30-31
  double scale;          // 1 / (2 N^2)
  int stage;             // 0: wf = fft(w0) (hybrid.jl:26-27), 1..3: RK3 stages
  int nrows;             // N/2
};

template <class C>
VMK_HD void kh_body(const Ctx& c, const KHArgs& a) {
  using F = Fft<C>;
  constexpr int N = C::N, E = C::E, T = C::T, P = C::P;
  constexpr int bl = C::bits(P - 1), rl = 1 << bl;
  static_assert(!C::SPLIT, "kh_body uses the plain exchange buffer");
  double2* tw = F::tables(c.smem);
  F::load_tables(c, tw, a.tw);
  c.sync();
  const int g = c.tid / T, t = c.tid % T;
  double2* sm = F::xbuf(c.smem, g);
  const int nblocks = (a.nrows + C::FPC - 1) / C::FPC;
  const bool zero_mode = a.stage != 3;  // wf[1,1] = 0 after the transform and after stages 1, 2 (hybrid.jl:27,46,57)
  for (int rb = c.bid; rb < nblocks; rb += c.nblk) {
    const int row = rb * C::FPC + g;
    const bool active = row < a.nrows;
    const int kx = active ? row : 1;
    const bool cta_has_row0 = rb == 0;
    const size_t roff = (size_t)kx * N;
    double2 v[E];
    static_for<0, E>([&](auto e_) {
      constexpr int e = decltype(e_)::value;
      v[e] = active ? ld_stream2(a.T + roff + F::template own_pos<e>(t)) : mk2(0.0, 0.0);
    });
    c.sync();  // the previous row's last exchange has been read everywhere
    F::forward(c, v, sm, tw, t);
    // ---- mode update (registers hold the last-pass layout: element e of thread t is ky = k_of_pos(((t+T*u)<<bl)|p)) ----
    if (cta_has_row0) {
      // packed row: X[ky] = A^[ky] + i B^[ky] (A: kx = 0, B: kx = N/2).  Y = rho dt J + gamma dt jf and W are separated
      // into their A and B parts with the values at the mirrored index, updated with each part's own coefficients
      // and packed again; wf'/k2 follows from the parts directly.
      double2 y[E];
      static_for<0, E>([&](auto e_) {
        constexpr int e = decltype(e_)::value;
        y[e] = cscale(v[e], a.gdt);
        if (a.stage >= 2 && active) {
          const double2 jp = a.J[roff + e * T + t];
          y[e] = mk2(fma_(a.rdt, jp.x, y[e].x), fma_(a.rdt, jp.y, y[e].y));
        }
        if ((a.stage == 1 || a.stage == 2) && active) a.J[roff + e * T + t] = v[e];
      });
      auto mirror_addr = [&](int u, int p) {
        return F::addr(F::pos_of_k((N - F::k_of_pos(((t + T * u) << bl) | p)) & (N - 1)));
      };
      // (all barriers of this branch sit outside the per-transform `kx == 0` test: at small N several transforms share
      // a warp, and a barrier inside a divergent branch would hang)
      double2 ym[E], w[E], wm[E];
      F::template store_smem<P - 1>(y, sm, t);
      c.sync();
      static_for<0, E>([&](auto e_) {
        constexpr int e = decltype(e_)::value;
        ym[e] = sm[mirror_addr(e / rl, e % rl)];
      });
      c.sync();
      static_for<0, E>([&](auto e_) {
        constexpr int e = decltype(e_)::value;
        w[e] = (a.stage >= 1 && active) ? a.W[roff + e * T + t] : mk2(0.0, 0.0);
      });
      F::template store_smem<P - 1>(w, sm, t);
      c.sync();
      static_for<0, E>([&](auto e_) {
        constexpr int e = decltype(e_)::value;
        wm[e] = sm[mirror_addr(e / rl, e % rl)];
      });
      c.sync();
      if (kx == 0) {
        const double ka = ld_ro(a.ksq + 0), kb = ld_ro(a.ksq + N / 2);
        static_for<0, E>([&](auto e_) {
          constexpr int e = decltype(e_)::value, u = e / rl, p = e % rl;
          const int ky = F::k_of_pos(((t + T * u) << bl) | p);
          const double kyy = ld_ro(a.ksq + ky);
          const double k2a = ka + kyy, k2b = kb + kyy;                  // Common.jl:199-201
          const double da = a.alpha * (a.zfac * k2a), db = a.alpha * (a.zfac * k2b);
          const double ga = rcp_rn(1.0 + da), gb = rcp_rn(1.0 + db);
          const double ca = (1.0 - da) * ga, cb = (1.0 - db) * gb;
          // parts: A = (X + conj Xm)/2, B = -i (X - conj Xm)/2
          const double2 wa = mk2(.5 * (w[e].x + wm[e].x), .5 * (w[e].y - wm[e].y));
          const double2 wb = mk2(.5 * (w[e].y + wm[e].y), .5 * (wm[e].x - w[e].x));
          const double2 ya = mk2(.5 * (y[e].x + ym[e].x), .5 * (y[e].y - ym[e].y));
          const double2 yb = mk2(.5 * (y[e].y + ym[e].y), .5 * (ym[e].x - y[e].x));
          double2 pa, pb;
          if (a.stage == 0) {
            pa = ya;  // caller passes gamma dt = 1: wf = fft(w0)
            pb = yb;
          } else {
            pa = mk2(fma_(ca, wa.x, ga * ya.x), fma_(ca, wa.y, ga * ya.y));
            pb = mk2(fma_(cb, wb.x, gb * yb.x), fma_(cb, wb.y, gb * yb.y));
          }
          if (ky == 0 && zero_mode) pa = mk2(0.0, 0.0);
          v[e] = mk2(pa.x - pb.y, pa.y + pb.x);  // A' + i B'
          const double ra = a.scale * rcp_rn(k2a), rb2 = a.scale * rcp_rn(k2b);
          y[e] = mk2(pa.x * ra - pb.y * rb2, pa.y * ra + pb.x * rb2);  // (A'/k2a + i B'/k2b) / (2 N^2)
        });
      } else {
        // an ordinary row that shares the CTA with row 0 (several transforms per CTA at small N)
        const double kxx = ld_ro(a.ksq + kx);
        static_for<0, E>([&](auto e_) {
          constexpr int e = decltype(e_)::value, u = e / rl, p = e % rl;
          const int ky = F::k_of_pos(((t + T * u) << bl) | p);
          const double k2 = kxx + ld_ro(a.ksq + ky);
          const double d = a.alpha * (a.zfac * k2);
          const double gg = rcp_rn(1.0 + d), cc = (1.0 - d) * gg;
          if (a.stage >= 1) {
            v[e] = mk2(fma_(cc, w[e].x, gg * y[e].x), fma_(cc, w[e].y, gg * y[e].y));
          } else {
            v[e] = y[e];
          }
          const double r = a.scale * rcp_rn(k2);
          y[e] = cscale(v[e], r);
        });
      }
      if (active) {
        static_for<0, E>([&](auto e_) {
          constexpr int e = decltype(e_)::value;
          a.W[roff + e * T + t] = v[e];
        });
      }
      static_for<0, E>([&](auto e_) {
        constexpr int e = decltype(e_)::value;
        v[e] = y[e];  // the s-spectrum goes first; wf' is read back from W afterwards
      });
    } else {
      const double kxx = ld_ro(a.ksq + kx);
      static_for<0, E>([&](auto e_) {
        constexpr int e = decltype(e_)::value;
        const size_t off = roff + e * T + t;
        const double k2 = kxx + ld_ro(a.ksqperm + e * T + t);  // kx^2 + ky^2, Common.jl:199-201
        const double d = a.alpha * (a.zfac * k2);               // hybrid.jl:34-37
        const double gg = rcp_fast(1.0 + d), cc = (1.0 - d) * gg;
        double2 yy = cscale(v[e], a.gdt);
        if (a.stage >= 2 && active) {
          const double2 jp = a.J[off];
          yy = mk2(fma_(a.rdt, jp.x, yy.x), fma_(a.rdt, jp.y, yy.y));
        }
        if ((a.stage == 1 || a.stage == 2) && active) a.J[off] = v[e];
        double2 wn = yy;  // stage 0: gamma dt = 1
        if (a.stage >= 1) {
          const double2 w = active ? a.W[off] : mk2(0.0, 0.0);
          wn = mk2(fma_(cc, w.x, gg * yy.x), fma_(cc, w.y, gg * yy.y));  // hybrid.jl:42-45,51-55,61-65
        }
        if (active) a.W[off] = wn;
        v[e] = cscale(wn, a.scale * rcp_fast(k2));  // sf = wf / k2, hybrid.jl:124
      });
    }
    // ---- psi spectrum back along j, then wf' (re-read: the thread's own elements, just written) -------------------
    F::inverse(c, v, sm, tw, t);
    if (active) {
      static_for<0, E>([&](auto e_) {
        constexpr int e = decltype(e_)::value;
        st_stream2(a.Vs + roff + F::template own_pos<e>(t), v[e]);
      });
    }
    static_for<0, E>([&](auto e_) {
      constexpr int e = decltype(e_)::value;
      v[e] = active ? cscale(a.W[roff + e * T + t], a.scale) : mk2(0.0, 0.0);
    });
    c.sync();  // the inverse's last exchange has been read everywhere
    F::inverse(c, v, sm, tw, t);
    if (active) {
      static_for<0, E>([&](auto e_) {
        constexpr int e = decltype(e_)::value;
        st_stream2(a.Vw + roff + F::template own_pos<e>(t), v[e]);
      });
    }
  }
}

}  // namespace vmk
