// vmk_fft.cuh -- CTA-level power-of-two complex FP64 FFT, data in registers, exchanged through
// padded shared memory.  Replaces the reference's FFTW calls (Common.jl:117 fft, :123 ifft).
//
// Scheme (derivation and numpy spec: tests/models/fft_proto.py):
//   * N = 2^M points, E = 2^LE values per thread, T = N/E threads per transform.
//   * forward = in-place mixed-radix DIF: pass k does radix-2^b_k butterflies on position bits
//     [lo_k, hi_k) entirely in registers, multiplies by W_{2^hi_k}^{low*p}, and the values go back
//     to the SAME positions.  Natural order in, digit-reversed ("position") order out.
//   * inverse = the adjoint DIT (conjugate twiddle first, then the conjugate butterfly), passes in
//     reverse order: position order in, natural order out, unnormalised.
//   * the spectral side never needs natural order: the divide, the real-pair unpack and the
//     transposed store are all index-agnostic, so no reordering pass exists anywhere.
//   * shared memory address of position pos is pos + (pos >> b_last): conflict-free 128-bit
//     accesses for every pass (checked with the bank model in tests/models/fft_proto.py).
#pragma once
#include "vmk_common.cuh"

namespace vmk {

// cos(pi*j/16), j = 0..8, correctly rounded (mpmath)
VMK_HD constexpr double cos16th(int j) {
  return j == 0   ? 1.0
         : j == 1 ? 0.9807852804032304
         : j == 2 ? 0.9238795325112867
         : j == 3 ? 0.8314696123025452
         : j == 4 ? 0.7071067811865476
         : j == 5 ? 0.5555702330196022
         : j == 6 ? 0.3826834323650898
         : j == 7 ? 0.19509032201612828
                  : 0.0;
}
// cos / sin of 2*pi*j/32 for j in [0,16]
VMK_HD constexpr double cos32(int j) { return j <= 8 ? cos16th(j) : -cos16th(16 - j); }
VMK_HD constexpr double sin32(int j) { return j <= 8 ? cos16th(8 - j) : cos16th(j - 8); }

VMK_HD constexpr int brev(int x, int bits) {
  int r = 0;
  for (int b = 0; b < bits; b++)
    if (x & (1 << b)) r |= 1 << (bits - 1 - b);
  return r;
}
VMK_HD constexpr int ilog2c(int n) {
  int m = 0;
  while ((1 << m) < n) m++;
  return m;
}

// d * W_R^I  (SIGN=-1: W = exp(-2 pi i/R), SIGN=+1: its conjugate), I in [0, R/2)
template <int R, int I, int SIGN>
VMK_HD double2 mul_const(double2 d) {
  if constexpr (I == 0) {
    return d;
  } else if constexpr (4 * I == R) {
    return SIGN < 0 ? mk2(d.y, -d.x) : mk2(-d.y, d.x);
  } else {
    constexpr int j = I * (32 / R);
    constexpr double c = cos32(j), s = sin32(j);
    if constexpr (j == 4) {  // c == s: 2 adds + 2 multiplies
      if constexpr (SIGN < 0)
        return mk2((d.x + d.y) * c, (d.y - d.x) * c);
      else
        return mk2((d.x - d.y) * c, (d.y + d.x) * c);
    } else if constexpr (j == 12) {  // c == -s
      if constexpr (SIGN < 0)
        return mk2((d.y - d.x) * s, -((d.x + d.y) * s));
      else
        return mk2(-((d.x + d.y) * s), (d.x - d.y) * s);
    } else if constexpr (SIGN < 0) {
      return mk2(fma_(d.x, c, d.y * s), fma_(d.y, c, -(d.x * s)));
    } else {
      return mk2(fma_(d.x, c, -(d.y * s)), fma_(d.y, c, d.x * s));
    }
  }
}

// radix-R DIF network on a[OFF .. OFF+R): natural order in, bit-reversed order out.
template <int R, int SIGN, int OFF>
struct Net {
  template <class A>
  VMK_HD static void run(A& a) {
    constexpr int H = R / 2;
    static_for<0, H>([&](auto i_) {
      constexpr int i = decltype(i_)::value;
      const double2 x = a[OFF + i], y = a[OFF + i + H];
      a[OFF + i] = cadd(x, y);
      a[OFF + i + H] = mul_const<R, i, SIGN>(csub(x, y));
    });
    Net<H, SIGN, OFF>::run(a);
    Net<H, SIGN, OFF + H>::run(a);
  }
};
template <int SIGN, int OFF>
struct Net<1, SIGN, OFF> {
  template <class A>
  VMK_HD static void run(A&) {}
};

// ---- configuration ---------------------------------------------------------------------------
// SPLIT_: the exchange buffer holds one double per position (real parts, then imaginary parts, three barriers per
// exchange instead of one) and the table of the first pass is stored as one octant.  Both shrink shared memory so that
// a separate full-row LANDING buffer fits next to them: the next row's asynchronous copies are then issued a whole
// row ahead instead of only during the last butterflies (used for N = 8192, where one row fills an SM).
template <int M_, int LE_, int P_, int B0_, int B1_, int B2_, int CT_, int MINB_, int B3_ = 0, bool SPLIT_ = false>
struct FftCfg {
  static constexpr bool SPLIT = SPLIT_;
  static constexpr int M = M_, N = 1 << M_, LE = LE_, E = 1 << LE_, T = N >> LE_, P = P_;
  static constexpr int CT = CT_ > T ? CT_ : T;  // CTA threads
  static constexpr int FPC = CT / T;            // transforms per CTA
  static constexpr int MINB = MINB_;            // min CTAs per SM (launch bounds)
  VMK_HD static constexpr int bits(int k) { return k == 0 ? B0_ : (k == 1 ? B1_ : (k == 2 ? B2_ : B3_)); }
  VMK_HD static constexpr int hi(int k) {
    int h = M_;
    for (int m = 0; m < k; m++) h -= bits(m);
    return h;
  }
  VMK_HD static constexpr int lo(int k) { return hi(k) - bits(k); }
  static constexpr int PADSH = bits(P_ - 1);
  static constexpr int SMN = N + (N >> PADSH);  // padded complex slots per transform
  // twiddle tables: pass k < P-1 uses W_{Bk}^x, Bk = 2^hi(k); full wave if small, else half wave
  VMK_HD static constexpr bool tw_full(int k) { return (1 << hi(k)) <= 1024; }
  VMK_HD static constexpr bool tw_oct(int k) { return SPLIT_ && !tw_full(k); }  // one octant + 1 entries (+1 pad)
  VMK_HD static constexpr int tw_len(int k) {
    return k >= P_ - 1 ? 0 : (tw_full(k) ? (1 << hi(k)) : (tw_oct(k) ? (1 << (hi(k) - 3)) + 2 : (1 << (hi(k) - 1))));
  }
  VMK_HD static constexpr int tw_off(int k) {
    int o = 0;
    for (int m = 0; m < k; m++) o += tw_len(m);
    return o;
  }
  static constexpr int TWN = tw_off(P_);  // total table entries (complex)
  static constexpr size_t XBYTES = (SPLIT_ ? sizeof(double) : sizeof(double2)) * (size_t)SMN;  // exchange buffer
  static constexpr size_t LANDBYTES = SPLIT_ ? sizeof(double2) * (size_t)N : 0;                // landing buffer
  static constexpr size_t XSTRIDE = XBYTES + LANDBYTES;                                         // per transform
  static constexpr size_t SMEM_DATA = XSTRIDE * FPC;
  static constexpr size_t SMEM_BYTES = SMEM_DATA + sizeof(double2) * (size_t)TWN;
  static_assert(XBYTES % 16 == 0, "alignment");
  static_assert(B0_ + B1_ + (P_ > 2 ? B2_ : 0) + (P_ > 3 ? B3_ : 0) == M_, "pass bits must sum to M");
  static_assert(B0_ <= LE_ && B1_ <= LE_ && B2_ <= LE_ && B3_ <= LE_, "radix exceeds per-thread elements");
};

// one configuration per supported size (32 .. 8192); radices grow towards the last
// (twiddle-free) pass.  M=13 keeps 32 values per thread so a row needs only 2 exchanges.
template <int M>
struct CfgFor;
template <> struct CfgFor<5>  { using type = FftCfg<5, 4, 2, 2, 3, 0, 128, 4>; };
template <> struct CfgFor<6>  { using type = FftCfg<6, 4, 2, 3, 3, 0, 128, 4>; };
template <> struct CfgFor<7>  { using type = FftCfg<7, 4, 2, 3, 4, 0, 128, 4>; };
template <> struct CfgFor<8>  { using type = FftCfg<8, 4, 2, 4, 4, 0, 128, 4>; };
#ifdef VMK_SPLIT_TEST  // test builds only: run the split-exchange / landing-buffer / octant-table path at small sizes too
template <> struct CfgFor<9>  { using type = FftCfg<9, 4, 3, 3, 3, 3, 128, 1, 0, true>; };
template <> struct CfgFor<10> { using type = FftCfg<10, 4, 3, 3, 3, 4, 128, 1, 0, true>; };
template <> struct CfgFor<11> { using type = FftCfg<11, 4, 3, 3, 4, 4, 128, 1, 0, true>; };
template <> struct CfgFor<12> { using type = FftCfg<12, 4, 3, 4, 4, 4, 256, 1, 0, true>; };
#else
template <> struct CfgFor<9>  { using type = FftCfg<9, 4, 3, 3, 3, 3, 128, 4>; };
template <> struct CfgFor<10> { using type = FftCfg<10, 4, 3, 3, 3, 4, 128, 4>; };
template <> struct CfgFor<11> { using type = FftCfg<11, 4, 3, 3, 4, 4, 128, 3>; };
template <> struct CfgFor<12> { using type = FftCfg<12, 4, 3, 4, 4, 4, 256, 2>; };
#endif
#if defined(VMK_M13_T512)
template <> struct CfgFor<13> { using type = FftCfg<13, 4, 4, 3, 3, 3, 512, 1, 4>; };
#elif defined(VMK_M13_SPLIT)  // measured alternative (profiles/r01_notes.md): K1 -2 %, K2 +-0, K3 +12 %
template <> struct CfgFor<13> { using type = FftCfg<13, 5, 3, 4, 4, 5, 256, 1, 0, true>; };
#else
template <> struct CfgFor<13> { using type = FftCfg<13, 5, 3, 4, 4, 5, 256, 1>; };
#endif

// ---- the engine --------------------------------------------------------------------------------
template <class C>
struct Fft {
  static constexpr int E = C::E, T = C::T, P = C::P, M = C::M, N = C::N;

  VMK_HD static int addr(int pos) { return pos + (pos >> C::PADSH); }

  // shared-memory carving: [transform g: exchange buffer | landing buffer] ... [twiddle tables]
  VMK_HD static double2* xbuf(unsigned char* smem, int g) {
    return reinterpret_cast<double2*>(smem + C::XSTRIDE * (size_t)g);
  }
  // where asynchronous copies of the next row land: a buffer of its own (SPLIT) or the exchange buffer itself
  VMK_HD static double2* landing(unsigned char* smem, int g) {
    return reinterpret_cast<double2*>(smem + C::XSTRIDE * (size_t)g + (C::SPLIT ? C::XBYTES : 0));
  }
  VMK_HD static int land_addr(int pos) { return C::SPLIT ? pos : addr(pos); }
  VMK_HD static double2* tables(unsigned char* smem) { return reinterpret_cast<double2*>(smem + C::SMEM_DATA); }

  // position of element q of butterfly u of thread t in pass K
  template <int K>
  VMK_HD static int base_pos(int t, int u, int& low) {
    constexpr int l = C::lo(K), h = C::hi(K);
    const int id = t + T * u;
    low = id & ((1 << l) - 1);
    return ((id >> l) << h) | low;
  }

  template <int K>
  VMK_HD static void load_smem(double2 (&v)[E], const double2* sm, int t) {
    constexpr int b = C::bits(K), r = 1 << b, l = C::lo(K);
    static_for<0, E / r>([&](auto u_) {
      constexpr int u = decltype(u_)::value;
      int low;
      const int bp = base_pos<K>(t, u, low);
      static_for<0, r>([&](auto q_) {
        constexpr int q = decltype(q_)::value;
        v[u * r + q] = sm[addr(bp | (q << l))];
      });
    });
  }
  template <int K>
  VMK_HD static void store_smem(const double2 (&v)[E], double2* sm, int t) {
    constexpr int b = C::bits(K), r = 1 << b, l = C::lo(K);
    static_for<0, E / r>([&](auto u_) {
      constexpr int u = decltype(u_)::value;
      int low;
      const int bp = base_pos<K>(t, u, low);
      static_for<0, r>([&](auto q_) {
        constexpr int q = decltype(q_)::value;
        sm[addr(bp | (q << l))] = v[u * r + q];
      });
    });
  }

  // one component (PART 0: real, 1: imaginary) of the registers to / from the split exchange buffer
  template <int K, int PART>
  VMK_HD static void store_part(const double2 (&v)[E], double* sd, int t) {
    constexpr int b = C::bits(K), r = 1 << b, l = C::lo(K);
    static_for<0, E / r>([&](auto u_) {
      constexpr int u = decltype(u_)::value;
      int low;
      const int bp = base_pos<K>(t, u, low);
      static_for<0, r>([&](auto q_) {
        constexpr int q = decltype(q_)::value;
        sd[addr(bp | (q << l))] = PART ? v[u * r + q].y : v[u * r + q].x;
      });
    });
  }
  template <int K, int PART>
  VMK_HD static void load_part(double2 (&v)[E], const double* sd, int t) {
    constexpr int b = C::bits(K), r = 1 << b, l = C::lo(K);
    static_for<0, E / r>([&](auto u_) {
      constexpr int u = decltype(u_)::value;
      int low;
      const int bp = base_pos<K>(t, u, low);
      static_for<0, r>([&](auto q_) {
        constexpr int q = decltype(q_)::value;
        if constexpr (PART)
          v[u * r + q].y = sd[addr(bp | (q << l))];
        else
          v[u * r + q].x = sd[addr(bp | (q << l))];
      });
    });
  }
  // registers in pass-KS layout -> registers in pass-KL layout.  A thread writes exactly the slots it read in the
  // previous exchange, so no barrier is needed before the first store.
  template <int KS, int KL>
  VMK_HD static void exchange(const Ctx& c, double2 (&v)[E], double2* sm, int t) {
    if constexpr (C::SPLIT) {
      double* sd = reinterpret_cast<double*>(sm);
      store_part<KS, 0>(v, sd, t);
      c.sync();
      load_part<KL, 0>(v, sd, t);
      c.sync();
      store_part<KS, 1>(v, sd, t);
      c.sync();
      load_part<KL, 1>(v, sd, t);
    } else {
      store_smem<KS>(v, sm, t);
      c.sync();
      load_smem<KL>(v, sm, t);
    }
  }

  // W_{2^hi(K)}^x, x < 2^hi(K); tw points at the start of all tables
  template <int K>
  VMK_HD static double2 twiddle(const double2* tw, int x) {
    const double2* tk = tw + C::tw_off(K);
    if constexpr (C::tw_full(K)) {
      return tk[x];
    } else if constexpr (C::tw_oct(K)) {
      // table: (cos, sin)(2 pi r / B), r = 0 .. B/8; the other seven octants by reflection
      constexpr int ob = C::hi(K) - 3;
      const int o = x >> ob;
      int r = x & ((1 << ob) - 1);
      if (o & 1) r = (1 << ob) - r;
      const double2 cs = tk[r];
      const bool swap = ((o ^ (o >> 1)) & 1) != 0;
      double co = swap ? cs.y : cs.x, si = swap ? cs.x : cs.y;
      if (((o + 2) >> 2) & 1) co = -co;
      if (!((o >> 2) & 1)) si = -si;  // W = cos - i sin
      return mk2(co, si);
    } else {
      constexpr int half = 1 << (C::hi(K) - 1);
      double2 w = tk[x & (half - 1)];
      if (x & half) {
        w.x = -w.x;
        w.y = -w.y;
      }
      return w;
    }
  }

  // forward pass K on registers: radix network, then twiddle.  v[u*r+p] <- y[p].
  template <int K>
  VMK_HD static void fwd_compute(double2 (&v)[E], const double2* tw, int t) {
    constexpr int b = C::bits(K), r = 1 << b;
    static_for<0, E / r>([&](auto u_) {
      constexpr int u = decltype(u_)::value;
      double2 a[r];
      static_for<0, r>([&](auto q_) {
        constexpr int q = decltype(q_)::value;
        a[q] = v[u * r + q];
      });
      Net<r, -1, 0>::run(a);
      int low;
      (void)base_pos<K>(t, u, low);
      static_for<0, r>([&](auto p_) {
        constexpr int p = decltype(p_)::value;
        v[u * r + p] = a[brev(p, b)];
      });
      if constexpr (K < P - 1) {
        // W^(low p): only the odd powers are read from the table (lanes hold consecutive `low`, so an odd stride is
        // bank-conflict free while p = 2, 4, 8 were 2-, 4-, 8-way conflicts); even powers follow by squaring
        static_for<0, r / 2>([&](auto h_) {
          constexpr int q = 2 * decltype(h_)::value + 1;
          double2 w = twiddle<K>(tw, low * q);
          v[u * r + q] = cmul(v[u * r + q], w);
          static_for<1, ilog2c(r)>([&](auto s_) {
            constexpr int m = q << decltype(s_)::value;
            if constexpr (m < r) {
              w = csqr(w);
              v[u * r + m] = cmul(v[u * r + m], w);
            }
          });
        });
      }
    });
  }
  // inverse pass K: conjugate twiddle, then conjugate network.  v[u*r+q] <- x[q].
  template <int K>
  VMK_HD static void inv_compute(double2 (&v)[E], const double2* tw, int t) {
    constexpr int b = C::bits(K), r = 1 << b;
    static_for<0, E / r>([&](auto u_) {
      constexpr int u = decltype(u_)::value;
      double2 a[r];
      int low;
      (void)base_pos<K>(t, u, low);
      static_for<0, r>([&](auto p_) {
        constexpr int p = decltype(p_)::value;
        a[p] = v[u * r + p];
      });
      if constexpr (K < P - 1) {
        static_for<0, r / 2>([&](auto h_) {
          constexpr int q = 2 * decltype(h_)::value + 1;
          double2 w = twiddle<K>(tw, low * q);
          a[q] = cmulc(a[q], w);
          static_for<1, ilog2c(r)>([&](auto s_) {
            constexpr int m = q << decltype(s_)::value;
            if constexpr (m < r) {
              w = csqr(w);
              a[m] = cmulc(a[m], w);
            }
          });
        });
      }
      Net<r, +1, 0>::run(a);
      static_for<0, r>([&](auto q_) {
        constexpr int q = decltype(q_)::value;
        v[u * r + q] = a[brev(q, b)];
      });
    });
  }

  // registers hold pass-0 layout (natural positions) on entry, last-pass layout on exit
  VMK_HD static void forward(const Ctx& c, double2 (&v)[E], double2* sm, const double2* tw, int t) {
    static_for<0, P>([&](auto k_) {
      constexpr int K = decltype(k_)::value;
      fwd_compute<K>(v, tw, t);
      if constexpr (K < P - 1) exchange<K, K + 1>(c, v, sm, t);
    });
  }
  // registers hold last-pass layout on entry, pass-0 layout (natural positions) on exit
  // `after_last_exchange` runs once the thread has read its pass-0 values back: from then on the thread's own
  // pass-0 slots of the exchange buffer are free (the kernels start the next row's asynchronous loads there)
  template <class Hook>
  VMK_HD static void inverse(const Ctx& c, double2 (&v)[E], double2* sm, const double2* tw, int t,
                             Hook&& after_last_exchange) {
    static_for<0, P>([&](auto k_) {
      constexpr int K = P - 1 - decltype(k_)::value;
      inv_compute<K>(v, tw, t);
      if constexpr (K > 0) {
        exchange<K, K - 1>(c, v, sm, t);
        if constexpr (K == 1) after_last_exchange();
      }
    });
  }
  VMK_HD static void inverse(const Ctx& c, double2 (&v)[E], double2* sm, const double2* tw, int t) {
    inverse(c, v, sm, tw, t, [] {});
  }
  // position of the thread's e-th register in the pass-0 layout (the slots a thread owns across iterations)
  template <int e>
  VMK_HD static int own_pos(int t) {
    constexpr int r = 1 << C::bits(0), l = C::lo(0);
    int low;
    return base_pos<0>(t, e / r, low) | ((e % r) << l);
  }

  // spectral index held at position pos after forward(), and its inverse map
  VMK_HD static int k_of_pos(int pos) {
    int k = 0;
    static_for<0, P>([&](auto k_) {
      constexpr int K = decltype(k_)::value;
      constexpr int sh = M - C::hi(K);  // sum of bits of earlier passes
      k |= ((pos >> C::lo(K)) & ((1 << C::bits(K)) - 1)) << sh;
    });
    return k;
  }
  VMK_HD static int pos_of_k(int k) {
    int pos = 0;
    static_for<0, P>([&](auto k_) {
      constexpr int K = decltype(k_)::value;
      constexpr int sh = M - C::hi(K);
      pos |= ((k >> sh) & ((1 << C::bits(K)) - 1)) << C::lo(K);
    });
    return pos;
  }

  // cooperative copy of the twiddle tables into shared memory (all CT threads)
  VMK_HD static void load_tables(const Ctx& c, double2* tw_sm, const double2* tw_g) {
    if (c.tables_resident) return;
    for (int i = c.tid; i < C::TWN; i += C::CT) tw_sm[i] = tw_g[i];
  }
};

}  // namespace vmk
