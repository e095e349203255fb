// vmk_kernels.cuh -- the kernel bodies of the vortex-merger step.
//
//   K1  k1_body  two real rows -> one complex FFT along i -> unpack -> transposed half spectrum
//                replaces Common.jl:134 (f = -w; the sign is folded into K2's scale), :115, :117 (dim 1)
//   K2  k2_body  FFT along j -> divide by aa + bb cos kx + cc cos ky (eps quirk, zero mode) -> inverse
//                FFT along j, one read and one write of the spectrum.  Common.jl:117-123 (dim 2)
//   K3  k3_body  repack -> inverse FFT along i -> psi rows (+ periodic/halo rows).  Common.jl:123,138-146
//   K4  k4_body  Arakawa Jacobian + Laplacian fused with the RK3 stage combine.
//                Common.jl:154-181 + vm.jl:28,43-47,62-66 (+ ghost fills :30-38)
//
// Device layout (one slab per GPU; a single GPU owns the whole grid):
//   real fields  w, psi : (NJ+2) rows of N doubles; row 0 / NJ+1 are the periodic (or neighbour's)
//                         halo rows, interior row jl lives at (jl+1)*N.  i (Julia dim 1) is contiguous.
//   spectrum     T      : N/2 rows (kx) of NJ complex (local j); row 0 packs kx=0 (re) and kx=N/2 (im),
//                         both of which are real sequences in j after the real-pair unpack.
#pragma once
#include "vmk_fft.cuh"

namespace vmk {

// idx in [0, N/2) -> position whose spectral index is < N/2 (last-pass digit < r_last/2); consecutive
// idx give consecutive positions in groups of r_last/2, so the shared-memory reads stay conflict-free
template <class C>
VMK_HD int halfspec_pos(int idx) {
  constexpr int bl = C::bits(C::P - 1);
  return (idx & ((1 << (bl - 1)) - 1)) | ((idx >> (bl - 1)) << bl);
}

// ======================================== K1 ====================================================
struct K1Args {
  const double* w;    // slab with halo rows (or the fps source f in the same layout)
  double2* T;         // local spectrum [N/2][NJ]
  const double2* tw;  // twiddle tables (global)
  int NJ;             // local rows
  int npairs;         // NJ/2
};

template <class C>
VMK_HD void k1_body(const Ctx& c, const K1Args& a) {
  using F = Fft<C>;
  constexpr int N = C::N, E = C::E, T = C::T, P = C::P;
  double2* sm_all = reinterpret_cast<double2*>(c.smem);
  double2* tw = sm_all + (size_t)C::SMN * C::FPC;
  F::load_tables(c, tw, a.tw);
  c.sync();
  const int g = c.tid / T, t = c.tid % T;
  double2* sm = sm_all + (size_t)C::SMN * g;
  const int nblocks = (a.npairs + C::FPC - 1) / C::FPC;
  for (int pb = c.bid; pb < nblocks; pb += c.nblk) {
    const int pair = pb * C::FPC + g;
    const bool active = pair < a.npairs;
    const int jl = 2 * pair;
    double2 v[E];
    {
      constexpr int r = 1 << C::bits(0), l = C::lo(0);
      const double* r0 = a.w + (size_t)(jl + 1) * N;
      const double* r1 = r0 + N;
      static_for<0, E / r>([&](auto u_) {
        constexpr int u = decltype(u_)::value;
        int low;
        const int bp = F::template base_pos<0>(t, u, low);
        static_for<0, r>([&](auto q_) {
          constexpr int q = decltype(q_)::value;
          const int pos = bp | (q << l);
          v[u * r + q] = active ? mk2(ld_stream1(r0 + pos), ld_stream1(r1 + pos)) : mk2(0.0, 0.0);
        });
      });
    }
    F::forward(c, v, sm, tw, t);
    F::template store_smem<P - 1>(v, sm, t);
    c.sync();
    if (active) {
      for (int idx = t; idx < N / 2; idx += T) {
        const int pos = halfspec_pos<C>(idx);
        const int k = F::k_of_pos(pos);
        double2 o0, o1;
        if (k == 0) {
          const double2 z0 = sm[F::addr(0)], zh = sm[F::addr(F::pos_of_k(N / 2))];
          o0 = mk2(2.0 * z0.x, 2.0 * zh.x);
          o1 = mk2(2.0 * z0.y, 2.0 * zh.y);
        } else {
          const double2 zk = sm[F::addr(pos)], zm = sm[F::addr(F::pos_of_k(N - k))];
          o0 = mk2(zk.x + zm.x, zk.y - zm.y);  // 2 X_j[k]   = Z[k] + conj Z[N-k]
          o1 = mk2(zk.y + zm.y, zm.x - zk.x);  // 2 X_j+1[k] = -i (Z[k] - conj Z[N-k])
        }
        st_stream4(a.T + (size_t)k * a.NJ + jl, o0, o1);
      }
    }
    c.sync();
  }
}

// ======================================== K2 ====================================================
struct K2Args {
  PeerPtrs T;           // per-rank spectrum buffers [N/2][NJ], transformed in place
  const double2* tw;    // twiddle tables
  const double* bbcos;  // [N]  bb*cos(kx[i])   Common.jl:120 (kx[1]=eps quirk inside)
  const double* cccos;  // [N]  cc*cos(ky[j])   (ky = kx, Common.jl:113)
  double aa;            // -2/dx^2 - 2/dy^2
  double scale;         // sign / (2 N^2): ifft normalisation, the factor 2 of the unpack, f = -w
  int NJ, log2NJ;
  int row0, nrows;      // kx rows owned by this rank
};

template <class C>
VMK_HD void k2_body(const Ctx& c, const K2Args& a) {
  using F = Fft<C>;
  constexpr int N = C::N, E = C::E, T = C::T, P = C::P, M = C::M;
  constexpr int bl = C::bits(P - 1), rl = 1 << bl;
  double2* sm_all = reinterpret_cast<double2*>(c.smem);
  double2* tw = sm_all + (size_t)C::SMN * C::FPC;
  F::load_tables(c, tw, a.tw);
  c.sync();
  const int g = c.tid / T, t = c.tid % T;
  double2* sm = sm_all + (size_t)C::SMN * g;
  const int nblocks = (a.nrows + C::FPC - 1) / C::FPC;
  for (int rb = c.bid; rb < nblocks; rb += c.nblk) {
    const int row = rb * C::FPC + g;
    const bool active = row < a.nrows;
    const int kx = a.row0 + row;
    const bool cta_has_row0 = (a.row0 + rb * C::FPC) == 0;
    double2 v[E];
    {
      constexpr int r = 1 << C::bits(0), l = C::lo(0);
      static_for<0, E / r>([&](auto u_) {
        constexpr int u = decltype(u_)::value;
        int low;
        const int bp = F::template base_pos<0>(t, u, low);
        static_for<0, r>([&](auto q_) {
          constexpr int q = decltype(q_)::value;
          const int j = bp | (q << l);
          const double2* src =
              reinterpret_cast<const double2*>(a.T.p[j >> a.log2NJ]) + (size_t)kx * a.NJ + (j & (a.NJ - 1));
          v[u * r + q] = active ? ld_stream2(src) : mk2(0.0, 0.0);
        });
      });
    }
    F::forward(c, v, sm, tw, t);
    // ---- divide (registers hold the last-pass layout: butterfly id = t + T*u, digit p) ----------
    if (cta_has_row0) {
      // packed DC/Nyquist row: C[ky] = A^[ky] + i B^[ky]; separate, divide each by its own divisor, repack
      F::template store_smem<P - 1>(v, sm, t);
      c.sync();
      if (kx == 0) {
        const double ab0 = a.aa + ld_ro(a.bbcos + 0), abn = a.aa + ld_ro(a.bbcos + N / 2);
        static_for<0, E / rl>([&](auto u_) {
          constexpr int u = decltype(u_)::value;
          const int id = t + T * u;
          static_for<0, rl>([&](auto p_) {
            constexpr int p = decltype(p_)::value;
            const int k = F::k_of_pos((id << bl) | p);
            const double2 cm = sm[F::addr(F::pos_of_k((N - k) & (N - 1)))];
            const double2 ck = v[u * rl + p];
            const double cc = ld_ro(a.cccos + k);
            const double g0 = 0.5 * a.scale * rcp_rn(ab0 + cc), gn = 0.5 * a.scale * rcp_rn(abn + cc);
            double2 pp = cscale(mk2(ck.x + cm.x, ck.y - cm.y), g0);  // A^' = (C + conj Cm)/2 * g
            const double2 qq = cscale(mk2(ck.y + cm.y, cm.x - ck.x), gn);  // B^' = -i(C - conj Cm)/2 * g
            if (k == 0) pp = mk2(0.0, 0.0);                              // e[1,1] = 0, Common.jl:118
            v[u * rl + p] = mk2(pp.x - qq.y, pp.y + qq.x);               // A^' + i B^'
          });
        });
      } else {
        const double ab = a.aa + ld_ro(a.bbcos + kx);
        static_for<0, E / rl>([&](auto u_) {
          constexpr int u = decltype(u_)::value;
          const int id = t + T * u;
          static_for<0, rl>([&](auto p_) {
            constexpr int p = decltype(p_)::value;
            const int k = F::k_of_pos((id << bl) | p);
            v[u * rl + p] = cscale(v[u * rl + p], a.scale * rcp_rn(ab + ld_ro(a.cccos + k)));
          });
        });
      }
      c.sync();
    } else {
      const double ab = a.aa + ld_ro(a.bbcos + (active ? kx : 1));
      static_for<0, E / rl>([&](auto u_) {
        constexpr int u = decltype(u_)::value;
        const int kb = F::k_of_pos((t + T * u) << bl);
        static_for<0, rl>([&](auto p_) {
          constexpr int p = decltype(p_)::value;
          const int k = kb | (p << (M - bl));
          v[u * rl + p] = cscale(v[u * rl + p], a.scale * rcp_rn(ab + ld_ro(a.cccos + k)));
        });
      });
    }
    F::inverse(c, v, sm, tw, t);
    if (active) {
      constexpr int r = 1 << C::bits(0), l = C::lo(0);
      static_for<0, E / r>([&](auto u_) {
        constexpr int u = decltype(u_)::value;
        int low;
        const int bp = F::template base_pos<0>(t, u, low);
        static_for<0, r>([&](auto q_) {
          constexpr int q = decltype(q_)::value;
          const int j = bp | (q << l);
          double2* dst = reinterpret_cast<double2*>(a.T.p[j >> a.log2NJ]) + (size_t)kx * a.NJ + (j & (a.NJ - 1));
          st_stream2(dst, v[u * r + q]);
        });
      });
    }
  }
}

// ======================================== K3 ====================================================
struct K3Args {
  const double2* T;   // local spectrum after K2: U[kx][jl]
  const double2* tw;
  double* psi;        // slab with halo rows
  double* lo_dst;     // where interior row 0 is mirrored: previous rank's top halo row (row NJ+1 there)
  double* hi_dst;     // where interior row NJ-1 is mirrored: next rank's bottom halo row (row 0 there)
  int NJ, npairs;
};

template <class C>
VMK_HD void k3_body(const Ctx& c, const K3Args& a) {
  using F = Fft<C>;
  constexpr int N = C::N, E = C::E, T = C::T, P = C::P;
  double2* sm_all = reinterpret_cast<double2*>(c.smem);
  double2* tw = sm_all + (size_t)C::SMN * C::FPC;
  F::load_tables(c, tw, a.tw);
  c.sync();
  const int g = c.tid / T, t = c.tid % T;
  double2* sm = sm_all + (size_t)C::SMN * g;
  const int nblocks = (a.npairs + C::FPC - 1) / C::FPC;
  for (int pb = c.bid; pb < nblocks; pb += c.nblk) {
    const int pair = pb * C::FPC + g;
    const bool active = pair < a.npairs;
    const int jl = 2 * pair;
    if (active) {
      for (int idx = t; idx < N / 2; idx += T) {
        const int pos = halfspec_pos<C>(idx);
        const int k = F::k_of_pos(pos);
        double2 ua, ub;
        ld_stream4(a.T + (size_t)k * a.NJ + jl, ua, ub);
        if (k == 0) {
          sm[F::addr(0)] = mk2(ua.x, ub.x);                      // Z[0]   = u0_j + i u0_j+1
          sm[F::addr(F::pos_of_k(N / 2))] = mk2(ua.y, ub.y);     // Z[N/2] = uN2_j + i uN2_j+1
        } else {
          sm[F::addr(pos)] = mk2(ua.x - ub.y, ua.y + ub.x);                  // U_j[k] + i U_j+1[k]
          sm[F::addr(F::pos_of_k(N - k))] = mk2(ua.x + ub.y, ub.x - ua.y);   // conj U_j[k] + i conj U_j+1[k]
        }
      }
    }
    c.sync();
    double2 v[E];
    F::template load_smem<P - 1>(v, sm, t);
    F::inverse(c, v, sm, tw, t);
    if (active) {
      constexpr int r = 1 << C::bits(0), l = C::lo(0);
      double* r0 = a.psi + (size_t)(jl + 1) * N;
      double* r1 = r0 + N;
      const bool first = (jl == 0), last = (jl + 2 == a.NJ);
      static_for<0, E / r>([&](auto u_) {
        constexpr int u = decltype(u_)::value;
        int low;
        const int bp = F::template base_pos<0>(t, u, low);
        static_for<0, r>([&](auto q_) {
          constexpr int q = decltype(q_)::value;
          const int pos = bp | (q << l);
          st_stream1(r0 + pos, v[u * r + q].x);
          st_stream1(r1 + pos, v[u * r + q].y);
          if (first) st_stream1(a.lo_dst + pos, v[u * r + q].x);
          if (last) st_stream1(a.hi_dst + pos, v[u * r + q].y);
        });
      });
    }
    c.sync();
  }
}

// ======================================== K4 ====================================================
struct K4Args {
  const double* w;    // stencil input (slab with halo rows): wn (stage 1) or wt (stages 2,3)
  const double* psi;  // streamfunction (slab with halo rows)
  const double* wn;   // pointwise input of stages 2,3 (may alias out in stage 3)
  double* out;        // slab with halo rows
  double* lo_dst;     // mirror of interior row 0   (previous rank's top halo row of `out`)
  double* hi_dst;     // mirror of interior row NJ-1 (next rank's bottom halo row of `out`)
  int N, log2N, NJ;
  int rows_per_cta;   // rows marched by one thread column
  double aa, bb;      // 1/(re dx^2), 1/(re dy^2)     Common.jl:149-150
  double gg, hh;      // 1/(4 dx dy), 1/3             Common.jl:151-152
  double dt;
};
constexpr int kK4Threads = 128;

// r at one point from the 3x3 neighbourhoods (index [dj+1][di+1]); expression shapes of Common.jl:155-180
VMK_HD double rhs_point(const double (&w)[3][3], const double (&s)[3][3], const K4Args& a) {
  const double j1 = (w[1][2] - w[1][0]) * (s[2][1] - s[0][1]) - (w[2][1] - w[0][1]) * (s[1][2] - s[1][0]);
  const double j2 = w[1][2] * (s[2][2] - s[0][2]) - w[1][0] * (s[2][0] - s[0][0]) - w[2][1] * (s[2][2] - s[2][0]) +
                    w[0][1] * (s[0][2] - s[0][0]);
  const double j3 = w[2][2] * (s[2][1] - s[1][2]) - w[0][0] * (s[1][0] - s[0][1]) - w[2][0] * (s[2][1] - s[1][0]) +
                    w[0][2] * (s[1][2] - s[0][1]);
  const double jac = a.gg * (j1 + j2 + j3) * a.hh;
  return -jac + (a.aa * (w[1][2] - 2.0 * w[1][1] + w[1][0]) + a.bb * (w[2][1] - 2.0 * w[1][1] + w[0][1]));
}

// MODE 0: out = r (vm_rhs);  1: wn + dt r;  2: .75 wn + .25 wt + (.25 dt) r;  3: wn/3 + (2/3) wt + ((2/3) dt) r
template <int MODE>
VMK_HD double rk_combine(double wn, double wt, double r, double dt) {
  if constexpr (MODE == 0) return r;
  if constexpr (MODE == 1) return wt + dt * r;  // stage 1: the stencil input IS wn (vm.jl:28)
  if constexpr (MODE == 2) return .75 * wn + .25 * wt + (.25 * dt) * r;             // vm.jl:43-47
  return wn / 3. + (2. / 3.) * wt + ((2. / 3.) * dt) * r;                           // vm.jl:62-66
}

// Each thread owns two adjacent columns (i, i+1) and marches rows_per_cta rows, keeping a rolling
// 3-row x 4-column window of w and psi in registers; i wraps periodically, j uses the halo rows.
template <int MODE>
VMK_HD void k4_body(const Ctx& c, const K4Args& a) {
  const int N = a.N;
  const int cols = N / 2;                                        // column pairs
  const int tw = cols < kK4Threads ? cols : kK4Threads;          // threads across i
  const int groups = kK4Threads / tw;                            // row groups per CTA
  const int ctas_x = cols / tw;
  const int bx = c.bid % ctas_x, by = c.bid / ctas_x;
  const int i0 = 2 * (bx * tw + c.tid % tw);
  const int grp = c.tid / tw;
  const int jbeg = (by * groups + grp) * a.rows_per_cta;
  if (jbeg >= a.NJ) return;
  const int jend = (jbeg + a.rows_per_cta < a.NJ) ? jbeg + a.rows_per_cta : a.NJ;
  const int im = (i0 - 1) & (N - 1), ip = (i0 + 2) & (N - 1);

  double W[3][4], S[3][4];  // [row: j-1, j, j+1][col: i-1, i, i+1, i+2]
  auto load_row = [&](int jl_halo /*slab row index incl. halo offset*/, double (&wr)[4], double (&sr)[4]) {
    const double* pw = a.w + (size_t)jl_halo * N;
    const double* ps = a.psi + (size_t)jl_halo * N;
    const double2 wc = *reinterpret_cast<const double2*>(pw + i0);
    const double2 sc = *reinterpret_cast<const double2*>(ps + i0);
    wr[0] = pw[im]; wr[1] = wc.x; wr[2] = wc.y; wr[3] = pw[ip];
    sr[0] = ps[im]; sr[1] = sc.x; sr[2] = sc.y; sr[3] = ps[ip];
  };
  load_row(jbeg, W[0], S[0]);      // j-1 of the first row (slab row jbeg = interior jbeg-1 + 1)
  load_row(jbeg + 1, W[1], S[1]);
  for (int jl = jbeg; jl < jend; jl++) {
    load_row(jl + 2, W[2], S[2]);
    double wn0 = 0.0, wn1 = 0.0;
    if constexpr (MODE >= 2) {
      const double2 t2 = *reinterpret_cast<const double2*>(a.wn + (size_t)(jl + 1) * N + i0);
      wn0 = t2.x;
      wn1 = t2.y;
    }
    double o[2];
#pragma unroll
    for (int e = 0; e < 2; e++) {
      double w9[3][3], s9[3][3];
#pragma unroll
      for (int dj = 0; dj < 3; dj++)
#pragma unroll
        for (int di = 0; di < 3; di++) {
          w9[dj][di] = W[dj][e + di];
          s9[dj][di] = S[dj][e + di];
        }
      const double r = rhs_point(w9, s9, a);
      o[e] = rk_combine<MODE>(e ? wn1 : wn0, W[1][e + 1], r, a.dt);
    }
    const double2 ov = mk2(o[0], o[1]);
    *reinterpret_cast<double2*>(a.out + (size_t)(jl + 1) * N + i0) = ov;
    if (jl == 0) *reinterpret_cast<double2*>(a.lo_dst + i0) = ov;
    if (jl == a.NJ - 1) *reinterpret_cast<double2*>(a.hi_dst + i0) = ov;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      W[0][q] = W[1][q]; W[1][q] = W[2][q];
      S[0][q] = S[1][q]; S[1][q] = S[2][q];
    }
  }
}

// ======================================== K5 ====================================================
// Conversions between the caller's ghosted column-major layout and the device slab (vm.jl:30-38,89,
// Common.jl:138-146: the ghost fills exist only on the host side of the boundary).
//   staging : (NJ+2) rows of N+2 doubles = ghosted rows j0 .. j0+NJ+1 of the caller's array
//   slab    : (NJ+2) rows of N doubles (row 0 / NJ+1 halo)
struct K5Args {
  const double* src;
  double* dst;
  int N, NJ;
};
constexpr int kK5Threads = 256;

// staging -> slab: drops the two i-ghost columns (periodicity in i is by index on the device)
VMK_HD void k5_unpack_body(const Ctx& c, const K5Args& a) {
  const size_t total = (size_t)(a.NJ + 2) * a.N;
  for (size_t q = (size_t)c.bid * kK5Threads + c.tid; q < total; q += (size_t)c.nblk * kK5Threads) {
    const size_t row = q / a.N, i = q % a.N;
    a.dst[q] = a.src[row * (a.N + 2) + i + 1];
  }
}
// slab -> staging: adds the i-ghost columns a[1,:] = a[nx+1,:], a[nx+2,:] = a[2,:]
VMK_HD void k5_pack_body(const Ctx& c, const K5Args& a) {
  const size_t ld = (size_t)a.N + 2, total = (size_t)(a.NJ + 2) * ld;
  for (size_t q = (size_t)c.bid * kK5Threads + c.tid; q < total; q += (size_t)c.nblk * kK5Threads) {
    const size_t row = q / ld, g = q % ld;
    a.dst[q] = a.src[row * a.N + ((g + a.N - 1) & (size_t)(a.N - 1))];
  }
}
// f = -w interior (Common.jl:134): slab rows 1..NJ -> NJ rows of N
VMK_HD void k5_negate_body(const Ctx& c, const K5Args& a) {
  const size_t total = (size_t)a.NJ * a.N;
  for (size_t q = (size_t)c.bid * kK5Threads + c.tid; q < total; q += (size_t)c.nblk * kK5Threads)
    a.dst[q] = -a.src[q + a.N];
}

}  // namespace vmk
