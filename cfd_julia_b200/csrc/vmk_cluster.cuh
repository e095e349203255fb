// vmk_cluster.cuh -- rows longer than one SM can hold (N = 16384, 32768): one transform per thread-block CLUSTER.
//
// A complex row of N = Q*N' points is 256 KB (N = 16384) or 512 KB (N = 32768); the register file of one SM holds
// 8192 points.  A cluster of Q = 2 or 4 CTAs (one per SM, same GPC) owns the row instead, and the CTAs talk through
// distributed shared memory (SM-to-SM network), so the row still crosses HBM once per kernel:
//
//   forward  = one radix-Q decimation-in-frequency pass ACROSS the cluster, then Q independent N'-point transforms
//              (the register-resident engine of vmk_fft.cuh, unchanged):
//                y_r[n] = (sum_q x[n + q N'] w_Q^{q r}) * W_N^{n r},   X[Q k' + r] = FFT_{N'}(y_r)[k']
//              CTA c loads x[n + q N'] for ITS quarter of the n (all q), does the butterflies and the twiddles, and
//              stores y_r[n] into the exchange buffer of CTA r.  After a cluster barrier CTA r holds y_r.
//   inverse  = the adjoint: Q independent N'-point inverse transforms, then CTA r sends its quarter c of yhat_r to
//              CTA c, which applies conj(W_N^{n r}) and the inverse radix-Q butterfly and owns x[n + q N'] for its n.
//
// CTA r of the cluster therefore holds the spectral indices k = Q k' + r.  Everything else is K1/K2/K3 of
// vmk_kernels.cuh with that index map: the half spectrum k < N/2 is k' < N'/2 in every CTA, and the mirror index
// N - k lives in CTA (Q - r) mod Q (at N' - k' for r = 0, N' - 1 - k' otherwise), so the real-pair unpack (K1), the
// repack (K3) and the packed DC/Nyquist row of K2 read or write the PARTNER CTA's shared memory.
//
// Same reference lines as vmk_kernels.cuh: K1 Common.jl:134,115,117; K2 :117-123; K3 :123,138-146.
#pragma once
#include "vmk_kernels.cuh"

namespace vmk {

template <class C, int Q>
struct Cl {
  using F = Fft<C>;
  static constexpr int NP = C::N;       // points per CTA (N')
  static constexpr int N = Q * C::N;    // points per row
  static constexpr int E = C::E, T = C::T, P = C::P, EQ = C::E / Q, LQ = ilog2c(Q);
  static constexpr int bl = C::bits(C::P - 1), rl = 1 << bl, hl = rl / 2;
  static constexpr int NI = NP / 2 / T;  // half-spectrum values per thread
  static_assert(!C::SPLIT, "the cluster kernels use the plain exchange buffer");
  static_assert(Q == 2 || Q == 4, "cluster size");
  static_assert(E % Q == 0, "values per thread must split over the cluster");

  // pass-0 position of the thread's e-th value, e not a compile-time constant (Fft::own_pos<e>)
  VMK_HD static int own_pos_rt(int t, int e) {
    constexpr int r = 1 << C::bits(0), l = C::lo(0), h = C::hi(0);
    const int id = t + T * (e / r);
    return (((id >> l) << h) | (id & ((1 << l) - 1))) | ((e % r) << l);
  }
  // the CTA that holds the mirror index N - k of this CTA's k, and the N'-point index of that mirror
  VMK_HD static int partner(int r) { return (Q - r) % Q; }
  VMK_HD static int mirror_kp(int r, int kp) { return r == 0 ? ((NP - kp) & (NP - 1)) : NP - 1 - kp; }

  // ---- forward pass across the cluster -------------------------------------------------------------------------
  // x[i][q] = the row's element n_i + q N' (n_i = the thread's i-th position of this CTA's share); stores y_r[n_i]
  // into CTA r's landing slots.  Callers bracket it with cluster barriers.
  // bases of the same buffer in every CTA of the cluster
  VMK_HD static void remote_bases(const Ctx& c, double2* buf, double2* (&rb)[Q]) {
    static_for<0, Q>([&](auto r_) {
      constexpr int r = decltype(r_)::value;
      rb[r] = c.remote(buf, r);
    });
  }
  VMK_HD static void scatter_forward(const Ctx& c, double2 (&x)[EQ][Q], double2* const (&land)[Q], const double2* ctw,
                                     int t) {
    static_for<0, EQ>([&](auto i_) {
      constexpr int i = decltype(i_)::value;
      const int n = own_pos_rt(t, c.crank * EQ + i);
      Net<Q, -1, 0>::run(x[i]);
      const double2 w1 = ld_ro2(ctw + n);  // W_N^n
      double2 w = w1;
      static_for<0, Q>([&](auto r_) {
        constexpr int r = decltype(r_)::value;
        double2 y = x[i][brev(r, LQ)];
        if constexpr (r == 1) y = cmul(y, w1);
        if constexpr (r == 2) {
          w = csqr(w1);
          y = cmul(y, w);
        }
        if constexpr (r == 3) {
          w = cmul(w, w1);
          y = cmul(y, w);
        }
        land[r][F::land_addr(n)] = y;
      });
    });
  }
  // after the barrier that follows scatter_forward: the CTA's y_r in the engine's pass-0 register layout
  VMK_HD static void gather_forward(double2 (&v)[E], const double2* land, int t, bool active) {
    static_for<0, E>([&](auto e_) {
      constexpr int e = decltype(e_)::value;
      v[e] = active ? land[F::land_addr(F::template own_pos<e>(t))] : mk2(0.0, 0.0);
    });
  }

  // ---- inverse pass across the cluster -------------------------------------------------------------------------
  // v = yhat_r in the pass-0 layout.  Slot (r*EQ + i)*T + t of CTA c's buffer receives yhat_r[n] for the i-th
  // position of thread t's share in CTA c.
  VMK_HD static void scatter_inverse(const Ctx& c, const double2 (&v)[E], double2* const (&buf)[Q], int t) {
    static_for<0, E>([&](auto e_) {
      constexpr int e = decltype(e_)::value;
      buf[e / EQ][(c.crank * EQ + e % EQ) * T + t] = v[e];
    });
  }
  // after the barrier: x[i][q] = the row's element n_i + q N'
  VMK_HD static void gather_inverse(const Ctx& c, double2 (&x)[EQ][Q], const double2* buf, const double2* ctw, int t) {
    static_for<0, EQ>([&](auto i_) {
      constexpr int i = decltype(i_)::value;
      const int n = own_pos_rt(t, c.crank * EQ + i);
      const double2 w1 = ld_ro2(ctw + n);
      double2 a[Q];
      double2 w = w1;
      static_for<0, Q>([&](auto r_) {
        constexpr int r = decltype(r_)::value;
        a[r] = buf[(r * EQ + i) * T + t];
        if constexpr (r == 1) a[r] = cmulc(a[r], w1);
        if constexpr (r == 2) {
          w = csqr(w1);
          a[r] = cmulc(a[r], w);
        }
        if constexpr (r == 3) {
          w = cmul(w, w1);
          a[r] = cmulc(a[r], w);
        }
      });
      Net<Q, +1, 0>::run(a);
      static_for<0, Q>([&](auto q_) {
        constexpr int q = decltype(q_)::value;
        x[i][q] = a[brev(q, LQ)];
      });
    });
  }
  // work item of K2 (PIECES layout) -> spectrum row, and the row's slot in a row pair's block of V:
  // item = r*(N'/2) + idx, idx = the order in which K3's threads consume the rows held by CTA r
  VMK_HD static int kx_of_item(int item) {
    const int r = item / (NP / 2), idx = item % (NP / 2);
    return Q * F::k_of_pos(halfspec_pos<C>(idx)) + r;
  }
};

// ======================================== K1 (cluster) ==========================================
// NAT (vmk_tri.cuh): CTA r stores its half spectrum to the slots [r N'/2, (r+1) N'/2) of the natural rows X[jl][N/2],
// slot r N'/2 + t + T i  <->  kx = Q own_half_k(t, i) + r; the rows kx < k0 also go to every rank's L
template <class C, int Q, bool NAT = false>
VMK_HD void k1c_body(const Ctx& c, const K1Args& a) {
  using L = Cl<C, Q>;
  using F = Fft<C>;
  constexpr int N = L::N, NP = L::NP, E = C::E, T = C::T, NI = L::NI, EQ = L::EQ;
  constexpr int bl = L::bl, rl = L::rl, hl = L::hl;
  double2* tw = F::tables(c.smem);
  F::load_tables(c, tw, a.tw);
  c.sync();
  const double2* ctw = a.tw + C::TWN;  // W_N^n, n < N' (global)
  const int g = c.tid / T, t = c.tid % T;
  double2* sm = F::xbuf(c.smem, g);
  double2* rsm[Q];
  L::remote_bases(c, sm, rsm);
  const int ncl = c.nblk / Q, cid = c.bid / Q;
  const int nblocks = (a.npairs + C::FPC - 1) / C::FPC;
  double2* const psm = c.remote(sm, L::partner(c.crank));  // the partner CTA's buffer (mirror indices)
  for (int pb = cid; pb < nblocks; pb += ncl) {
    const int pair = pb * C::FPC + g;
    const bool active = pair < a.npairs;
    const int jl = 2 * pair;
    // the CTA's share of the NEXT pair's rows (EQ*Q runs of T doubles per row) is pulled into L2 while this pair is
    // transformed: the loads at the top of the next iteration are synchronous (registers and shared memory are full)
    if (a.prefetch && pair + ncl * C::FPC < a.npairs) {
      const double* nx0 = a.w + (size_t)(jl + 2 * ncl * C::FPC + 1) * N;
      for (int b = t; b < EQ * Q; b += T) {
        const int off = L::own_pos_rt(0, c.crank * EQ + b / Q) + (b % Q) * NP;
        prefetch_l2_bulk(nx0 + off, (unsigned)(T * sizeof(double)));
        prefetch_l2_bulk(nx0 + N + off, (unsigned)(T * sizeof(double)));
      }
    }
    double2 v[E];
    {
      double2 x[EQ][Q];
      const double* r0 = a.w + (size_t)(jl + 1) * N;
      const double* r1 = r0 + N;
      static_for<0, EQ>([&](auto i_) {
        constexpr int i = decltype(i_)::value;
        const int n = L::own_pos_rt(t, c.crank * EQ + i);
        static_for<0, Q>([&](auto q_) {
          constexpr int q = decltype(q_)::value;
          x[i][q] = active ? mk2(ld_stream1(r0 + n + q * NP), ld_stream1(r1 + n + q * NP)) : mk2(0.0, 0.0);
        });
      });
      // every CTA is done with its buffer: the partner's mirror reads of the previous pair were consumed by that pair's
      // stores.  Nothing to publish, so no release fence -- the global loads just issued stay in flight across it
      c.cluster_sync_relaxed();
      L::scatter_forward(c, x, rsm, ctw, t);
      c.cluster_sync();
      L::gather_forward(v, sm, t, true);
    }
    F::forward(c, v, sm, tw, t);
    // unpack: the thread holds Z[k] for its k' < N'/2; the upper halves are published and the mirror values
    // Z[N-k] fetched from the partner CTA's shared memory
    double2 zm[NI];  // (a thread overwrites exactly the slots it read in the last exchange: no barrier needed)
    static_for<0, NI>([&](auto i_) {
      constexpr int i = decltype(i_)::value, u = i / hl, p = hl + i % hl;
      sm[F::addr(((t + T * u) << bl) | p)] = v[u * rl + p];
    });
    c.cluster_sync();
    static_for<0, NI>([&](auto i_) {
      constexpr int i = decltype(i_)::value;
      const int kp = own_half_k<C>(t, i);
      const int km = (c.crank == 0 && kp == 0) ? NP / 2 : L::mirror_kp(c.crank, kp);  // k == 0: Z[N/2]
      zm[i] = psm[F::addr(F::pos_of_k(km))];
    });
    if (active) {
      static_for<0, NI>([&](auto i_) {
        constexpr int i = decltype(i_)::value, u = i / hl, p = i % hl;
        const int k = Q * own_half_k<C>(t, i) + c.crank;
        const double2 zk = v[u * rl + p];
        double2 o0, o1;
        if (k == 0) {
          o0 = mk2(2.0 * zk.x, 2.0 * zm[i].x);
          o1 = mk2(2.0 * zk.y, 2.0 * zm[i].y);
        } else {
          o0 = mk2(zk.x + zm[i].x, zk.y - zm[i].y);  // 2 X_j[k]   = Z[k] + conj Z[N-k]
          o1 = mk2(zk.y + zm[i].y, zm[i].x - zk.x);  // 2 X_j+1[k] = -i (Z[k] - conj Z[N-k])
        }
        if constexpr (NAT) {
          double2* xr = a.X + (size_t)jl * (N / 2) + (size_t)c.crank * (NP / 2) + t + T * i;
          st_stream2(xr, o0);
          st_stream2(xr + N / 2, o1);
          if (k < a.k0) {
            for (int q = 0; q < a.nranks; q++)
              st_stream4(reinterpret_cast<double2*>(a.Lpeer.p[q]) + (size_t)k * N + a.jbase + jl, o0, o1);
          }
        } else {
          double2* dst = ((k >= a.k_own0 && k < a.k_own1) ? a.Tloc + (size_t)(k - a.k_own0) * a.NJ
                                                          : a.S + (size_t)k * a.NJ) + jl;
          st_stream4(dst, o0, o1);
        }
      });
    }
  }
  c.cluster_sync();  // no CTA leaves while a partner may still read its shared memory
}

// ======================================== K2 (cluster) ==========================================
template <class C, int Q, bool PIECES>
VMK_HD void k2c_body(const Ctx& c, const K2Args& a) {
  using L = Cl<C, Q>;
  using F = Fft<C>;
  constexpr int N = L::N, NP = L::NP, E = C::E, T = C::T, P = C::P, EQ = L::EQ;
  constexpr int bl = L::bl, rl = L::rl;
  double2* tw = F::tables(c.smem);
  F::load_tables(c, tw, a.tw);
  c.sync();
  const double2* ctw = a.tw + C::TWN;
  const int g = c.tid / T, t = c.tid % T;
  double2* sm = F::xbuf(c.smem, g);
  double2* rsm[Q];
  L::remote_bases(c, sm, rsm);
  const int ncl = c.nblk / Q, cid = c.bid / Q;
  const int nblocks = (a.nrows + C::FPC - 1) / C::FPC;
  double2* const psm = c.remote(sm, L::partner(c.crank));  // the partner CTA's buffer (mirror indices)
  for (int rb = cid; rb < nblocks; rb += ncl) {
    const int row = rb * C::FPC + g;
    const bool active = row < a.nrows;
    const int kx = PIECES ? L::kx_of_item(active ? row : 0) : a.row0 + row;
    const int trow = PIECES ? kx : a.rloc0 + row;
    bool cluster_has_row0 = false;  // uniform over the cluster: one of its transforms is the packed DC/Nyquist row
    if constexpr (PIECES) {
      cluster_has_row0 = rb == 0;  // item 0 is kx = 0
    } else {
      cluster_has_row0 = (a.row0 + rb * C::FPC) == 0;
    }
    if (a.prefetch && row + ncl * C::FPC < a.nrows && a.NJ >= T) {  // next row of this transform slot -> L2 (see k1c_body)
      const int nrow = row + ncl * C::FPC;
      const int ntrow = PIECES ? L::kx_of_item(nrow) : a.rloc0 + nrow;
      for (int b = t; b < EQ * Q; b += T) {
        const int j = L::own_pos_rt(0, c.crank * EQ + b / Q) + (b % Q) * NP;
        prefetch_l2_bulk(a.T + ((size_t)(j >> a.log2NJ) * a.R + ntrow) * a.NJ + (j & (a.NJ - 1)),
                         (unsigned)(T * sizeof(double2)));
      }
    }
    double2 v[E];
    {
      double2 x[EQ][Q];
      static_for<0, EQ>([&](auto i_) {
        constexpr int i = decltype(i_)::value;
        const int n = L::own_pos_rt(t, c.crank * EQ + i);
        static_for<0, Q>([&](auto q_) {
          constexpr int q = decltype(q_)::value;
          const int j = n + q * NP;
          x[i][q] = active ? ld_stream2(a.T + ((size_t)(j >> a.log2NJ) * a.R + trow) * a.NJ + (j & (a.NJ - 1)))
                           : mk2(0.0, 0.0);
        });
      });
      c.cluster_sync_relaxed();  // the previous row's gather_inverse has read every buffer (values consumed)
      L::scatter_forward(c, x, rsm, ctw, t);
      c.cluster_sync();
      L::gather_forward(v, sm, t, true);
    }
    F::forward(c, v, sm, tw, t);
    // ---- divide: CTA r holds ky = Q k' + r -----------------------------------------------------
    if (cluster_has_row0) {
      // packed DC/Nyquist row (see k2_body); the mirrored index lives in the partner CTA
      F::template store_smem<P - 1>(v, sm, t);
      c.cluster_sync();
      auto mirror = [&](int u, int p) {
        return psm[F::addr(F::pos_of_k(L::mirror_kp(c.crank, F::k_of_pos(((t + T * u) << bl) | p))))];
      };
      if (kx == 0) {
        const double ab0 = a.aa + ld_ro(a.bbcos + 0), abn = a.aa + ld_ro(a.bbcos + N / 2);
        static_for<0, E>([&](auto e_) {
          constexpr int e = decltype(e_)::value, u = e / rl, p = e % rl;
          const int k = Q * F::k_of_pos(((t + T * u) << bl) | p) + c.crank;
          const double2 cmv = mirror(u, p), ck = v[e];
          const double cc = ld_ro(a.cccos + k);
          const double g0 = 0.5 * a.scale * rcp_rn(ab0 + cc), gn = 0.5 * a.scale * rcp_rn(abn + cc);
          double2 pp = cscale(mk2(ck.x + cmv.x, ck.y - cmv.y), g0);        // A^' = (C + conj Cm)/2 * g
          const double2 qq = cscale(mk2(ck.y + cmv.y, cmv.x - ck.x), gn);  // B^' = -i(C - conj Cm)/2 * g
          if (k == 0) pp = mk2(0.0, 0.0);                                  // e[1,1] = 0, Common.jl:118
          v[e] = mk2(pp.x - qq.y, pp.y + qq.x);                            // A^' + i B^'
        });
      } else {
        const double ab = a.aa + ld_ro(a.bbcos + (active ? kx : 1));
        static_for<0, E>([&](auto e_) {
          constexpr int e = decltype(e_)::value, u = e / rl, p = e % rl;
          const int k = Q * F::k_of_pos(((t + T * u) << bl) | p) + c.crank;
          v[e] = cscale(v[e], a.scale * rcp_rn(ab + ld_ro(a.cccos + k)));
        });
      }
      c.cluster_sync();  // every mirror value has been read: the inverse's exchanges may overwrite the buffers
    } else {
      const double ab = a.aa + ld_ro(a.bbcos + (active ? kx : 1));
      double dd[E];
      static_for<0, E>([&](auto e_) {
        constexpr int e = decltype(e_)::value;
        dd[e] = ab + ld_ro(a.ccperm + (size_t)c.crank * NP + e * T + t);  // (aa + bb cos kx) + cc cos ky
      });
      static_for<0, E>([&](auto e_) {
        constexpr int e = decltype(e_)::value;
        v[e] = cscale(v[e], a.scale * rcp_fast(dd[e]));
      });
    }
    F::inverse(c, v, sm, tw, t);
    double2 x[EQ][Q];
    c.cluster_sync_relaxed();  // every thread of the cluster has read its last exchange (and used the values)
    L::scatter_inverse(c, v, rsm, t);
    c.cluster_sync();
    L::gather_inverse(c, x, sm, ctw, t);
    if constexpr (PIECES) {
      // lanes 2m, 2m+1 hold columns j, j+1: one shuffle per value pair gives each lane a whole 32-byte piece
      const bool odd = (t & 1) != 0;
      const size_t piece = (size_t)(active ? row : 0) * 2;
      static_for<0, EQ * Q / 2>([&](auto h_) {
        constexpr int f0 = 2 * decltype(h_)::value, f1 = f0 + 1;  // flat indices i*Q + q
        double2 send = odd ? x[f0 / Q][f0 % Q] : x[f1 / Q][f1 % Q];
        c.shfl_xor2(send.x, send.y, 1);
        const double2 keep = odd ? x[f1 / Q][f1 % Q] : x[f0 / Q][f0 % Q];
        const int fi = odd ? f1 : f0;
        const int j = L::own_pos_rt(t, c.crank * EQ + fi / Q) + (fi % Q) * NP;
        if (active) st_stream4(a.V + (size_t)(j >> 1) * N + piece, odd ? send : keep, odd ? keep : send);
      });
    } else if (active) {
      static_for<0, EQ>([&](auto i_) {
        constexpr int i = decltype(i_)::value;
        const int n = L::own_pos_rt(t, c.crank * EQ + i);
        static_for<0, Q>([&](auto q_) {
          constexpr int q = decltype(q_)::value;
          const int j = n + q * NP;
          const int h = j >> a.log2NJ;
          double2* dst = (h == a.rank ? a.V + (size_t)kx * a.NJ
                          : a.push    ? reinterpret_cast<double2*>(a.Vpeer.p[h]) + (size_t)kx * a.NJ
                                      : a.S + ((size_t)h * a.R + a.rloc0 + row) * a.NJ) +
                         (j & (a.NJ - 1));
          st_stream2(dst, x[i][q]);
        });
      });
    }
  }
  c.cluster_sync();
}

// ======================================== K3 (cluster) ==========================================
// LAYOUT as in k3_body: 0 rows [kx][NJ], 1 PIECES, 2 natural rows [jl][N/2] in k1c_body's slot order (vmk_tri.cuh)
template <class C, int Q, int LAYOUT>
VMK_HD void k3c_body(const Ctx& c, const K3Args& a) {
  using L = Cl<C, Q>;
  using F = Fft<C>;
  constexpr bool PIECES = LAYOUT == 1, NAT = LAYOUT == 2;
  constexpr int N = L::N, NP = L::NP, E = C::E, T = C::T, P = C::P, NI = L::NI, EQ = L::EQ;
  constexpr int bl = L::bl, hl = L::hl;
  double2* tw = F::tables(c.smem);
  F::load_tables(c, tw, a.tw);
  c.sync();
  const double2* ctw = a.tw + C::TWN;
  const int g = c.tid / T, t = c.tid % T;
  // position (last-pass layout) of the value a thread loads for its i-th half-spectrum register
  auto load_pos = [&](int i) { return NAT ? (((t + T * (i / hl)) << bl) | (i % hl)) : halfspec_pos<C>(t + T * i); };
  double2* sm = F::xbuf(c.smem, g);
  double2* rsm[Q];
  L::remote_bases(c, sm, rsm);
  const int ncl = c.nblk / Q, cid = c.bid / Q;
  const int nblocks = (a.npairs + C::FPC - 1) / C::FPC;
  double2* const psm = c.remote(sm, L::partner(c.crank));  // the partner CTA's buffer (mirror indices)
  for (int pb = cid; pb < nblocks; pb += ncl) {
    const int pair = pb * C::FPC + g;
    const bool active = pair < a.npairs;
    const int jl = 2 * pair;
    if constexpr (PIECES) {  // the CTA's pieces of the next pair are one contiguous 16 N' byte block
      if (a.prefetch && t == 0 && pair + ncl * C::FPC < a.npairs)
        prefetch_l2_bulk(a.T + (size_t)(pair + ncl * C::FPC) * N + (size_t)c.crank * NP, (unsigned)(NP * sizeof(double2)));
    }
    if constexpr (NAT) {  // ... two contiguous 8 N' byte runs (rows jl and jl + 1)
      if (a.prefetch && t == 0 && pair + ncl * C::FPC < a.npairs) {
        const double2* nx = a.T + (size_t)(pair + ncl * C::FPC) * N + (size_t)c.crank * (NP / 2);
        prefetch_l2_bulk(nx, (unsigned)(NP / 2 * sizeof(double2)));
        prefetch_l2_bulk(nx + N / 2, (unsigned)(NP / 2 * sizeof(double2)));
      }
    }
    double2 v[E];
    {
      // the pieces (U[k][j], U[k][j+1]) of this CTA's k = Q k' + r, k' < N'/2
      double2 ua[NI], ub[NI];
      static_for<0, NI>([&](auto i_) {
        constexpr int i = decltype(i_)::value;
        const int idx = t + T * i;
        if constexpr (NAT) {
          const double2* src = a.T + (size_t)jl * (N / 2) + (size_t)c.crank * (NP / 2) + idx;
          ua[i] = active ? ld_stream2(src) : mk2(0.0, 0.0);
          ub[i] = active ? ld_stream2(src + N / 2) : mk2(0.0, 0.0);
        } else {
          const double2* src = PIECES ? a.T + (size_t)pair * N + 2 * ((size_t)c.crank * (NP / 2) + idx)
                                      : a.T + (size_t)(Q * F::k_of_pos(halfspec_pos<C>(idx)) + c.crank) * a.NJ + jl;
          if (active) {
            ld_stream4(src, ua[i], ub[i]);
          } else {
            ua[i] = ub[i] = mk2(0.0, 0.0);
          }
        }
      });
      c.cluster_sync_relaxed();  // the previous pair's gather_inverse has read every buffer (values consumed)
      // Z = U_j + i U_j+1: Z[k] into the own buffer, Z[N-k] (from the conjugates) into the partner's
      static_for<0, NI>([&](auto i_) {
        constexpr int i = decltype(i_)::value;
        const int pos = load_pos(i);
        const int kp = F::k_of_pos(pos);
        if (c.crank == 0 && kp == 0) {
          sm[F::addr(0)] = mk2(ua[i].x, ub[i].x);                    // Z[0]   = u0_j + i u0_j+1
          sm[F::addr(F::pos_of_k(NP / 2))] = mk2(ua[i].y, ub[i].y);  // Z[N/2] = uN2_j + i uN2_j+1 (k' = N'/2, CTA 0)
        } else {
          sm[F::addr(pos)] = mk2(ua[i].x - ub[i].y, ua[i].y + ub[i].x);  // U_j[k] + i U_j+1[k]
          psm[F::addr(F::pos_of_k(L::mirror_kp(c.crank, kp)))] =
              mk2(ua[i].x + ub[i].y, ub[i].x - ua[i].y);  // conj(U_j[k]) + i conj(U_j+1[k]) = Z[N-k]
        }
      });
      c.cluster_sync();
      F::template load_smem<P - 1>(v, sm, t);
    }
    F::inverse(c, v, sm, tw, t);
    double2 x[EQ][Q];
    c.cluster_sync_relaxed();  // every thread of the cluster has read its last exchange (and used the values)
    L::scatter_inverse(c, v, rsm, t);
    c.cluster_sync();
    L::gather_inverse(c, x, sm, ctw, t);
    if (active) {
      double* r0 = a.psi + (size_t)(jl + 1) * N;
      double* r1 = r0 + N;
      const bool first = (jl == 0), last = (jl + 2 == a.NJ);
      static_for<0, EQ>([&](auto i_) {
        constexpr int i = decltype(i_)::value;
        const int n = L::own_pos_rt(t, c.crank * EQ + i);
        static_for<0, Q>([&](auto q_) {
          constexpr int q = decltype(q_)::value;
          const int pos = n + q * NP;
          st_stream1(r0 + pos, x[i][q].x);
          st_stream1(r1 + pos, x[i][q].y);
          if (first) st_stream1(a.lo_dst + pos, x[i][q].x);
          if (last) st_stream1(a.hi_dst + pos, x[i][q].y);
        });
      });
    }
  }
  c.cluster_sync();
}

}  // namespace vmk
