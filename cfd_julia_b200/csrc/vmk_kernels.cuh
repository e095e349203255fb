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
//   K1 output    S      : N/2 rows (all kx) of NJ complex (the rank's own j), written by K1 in 32-byte pieces (two
//                         adjacent j).  Global row 0 packs kx=0 (re) and kx=N/2 (im), both of which are real
//                         sequences in j after the real-pair unpack.
//   spectrum     T      : the rank's R = N/(2P) spectrum rows (kx = rank*R ..), ALL j, stored as P blocks [g][R][NJ]
//                         (block g = the columns that came from rank g), read by the local K2 row by row as P
//                         contiguous NJ-element segments.  On one GPU T is S.  On P GPUs rank g's S rows
//                         [h*R, (h+1)*R) -- one contiguous R*NJ*16-byte block -- are copied by a copy engine over
//                         NVLink into block g of T_h: the forward transpose of the distributed 2-D FFT.  (Pushing
//                         K1's 32-byte pieces straight into the peers' T ran NVLink at ~400 GB/s.)
//   solution     V      : N/2 rows (all kx) of NJ complex (the rank's own j), read by the local K3.  K2 stores the
//                         columns it owns itself straight into V and the others into S (block h = columns of rank h);
//                         block h -- contiguous -- is then copied over NVLink into rows [g*R, (g+1)*R) of V_h: the
//                         backward transpose.  K2 is launched in row chunks so that the copy of one chunk overlaps
//                         the transforms of the next.
//   Data crosses NVLink only as copy-engine transfers of large contiguous blocks (and the 64 KB halo rows that K3/K4
//   store directly); no kernel ever waits on a remote load.  Measured alternatives are in profiles/r01_notes.md.
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

// Half-spectrum index held by thread t in its i-th lower-half register of the last-pass layout (i = u*(r_last/2) + p,
// p < r_last/2: the last-pass digit is the top digit of k, so these are exactly the k < N/2): what a thread needs
// from the half spectrum is what it already holds in registers, and only the mirror values Z[N-k] go through shared
// memory.  (Used by K1's unpack; the same scheme for K3's repack was measured slower -- its hoisted address registers
// made ptxas spill 360 bytes per thread.)
template <class C>
VMK_HD int own_half_k(int t, int i) {
  constexpr int bl = C::bits(C::P - 1), hl = 1 << (bl - 1);
  return Fft<C>::k_of_pos(((t + C::T * (i / hl)) << bl) | (i % hl));
}

// ======================================== K1 ====================================================
struct K1Args {
  const double* w;    // slab with halo rows (or the fps source f in the same layout)
  double2* S;         // K1 output [N/2][NJ] for the spectrum rows owned by OTHER ranks (nullptr on a single GPU)
  double2* Tloc;      // block `rank` of the own spectrum buffer T ([R][NJ]): rows this rank owns go here
  const double2* tw;  // twiddle tables (global)
  int NJ;             // local rows (row pitch of S)
  int npairs;         // row pairs handled by this launch (w, S and Tloc are offset to its first pair)
  int k_own0, k_own1; // spectrum rows [k_own0, k_own1) are owned by this rank
  int prefetch;       // 0 off, 1: bulk L2 prefetch of the rows of the pair after next
  double2* X = nullptr;  // NAT: spectrum rows [jl][N/2] in slot order (vmk_tri.cuh), offset to the launch's first row
  PeerPtrs Lpeer = {};   // NAT: every rank's L[kx][j] (the rows kx < k0 keep the FFT form along j: K2 solves them there)
  int k0 = 0, jbase = 0, nranks = 1;  // NAT: rows kx < k0 also go to L; global j of the launch's first row; ranks
  int rev = 0;  // 1: the row pairs are taken in descending order: consecutive streaming kernels alternate direction, so
                // each starts on the ~100 MB its predecessor touched last, which are still in the 126 MB L2
  // FUSED (vmk_tri.cuh, "fused form"): the forward recurrence along j runs in K1's epilogue
  const double2* rr = nullptr;  // [N/2]: (r, 1/r) per slot
  double2* tot = nullptr;       // [3][units][N/2]: tp, A, S of every unit's block of rows
};


// Per row pair: rows (already in the exchange buffer, put there asynchronously during the previous pair's store
// phase) -> registers -> forward FFT -> spectrum to shared memory -> Z[k], Z[N-k] to registers -> start the
// asynchronous copy of the next pair's rows -> unpack and transposed 32-byte stores.
// NAT: the half spectrum of row jl goes to X[jl][s], slot s = t + T i = the order in which the threads hold it
// (coalesced 16-byte stores, nothing is transposed); the solve along j is then vmk_tri.cuh's.
template <class C, bool NAT = false, bool FUSED = false>
VMK_HD void k1_body(const Ctx& c, const K1Args& a) {
  using F = Fft<C>;
  static_assert(!(NAT && C::SPLIT), "the slot order of the natural layout is the own-half order");
  static_assert(!FUSED || (NAT && C::T >= 32), "the fused form stores natural rows; its state is per warp");
  constexpr int N = C::N, E = C::E, T = C::T, P = C::P, NI = N / 2 / T;
  constexpr int TMCOLS = tm_cols(6 * NI, C::CT);
  TmState<6 * NI> tm;  // FUSED: per slot u (last row's recurrence value), A, S
  if constexpr (FUSED) tm.template open<TMCOLS>(c);
  double2* tw = F::tables(c.smem);
  F::load_tables(c, tw, a.tw);
  c.sync();
  const int g = c.tid / T, t = c.tid % T;
  double2* sm = F::xbuf(c.smem, g);
  double* rows = reinterpret_cast<double*>(F::landing(c.smem, g));  // the pair's two rows as [2][N] doubles
  // FUSED: the loop counter pb is the position inside the unit's own block of pairs; otherwise the block of FPC pairs
  const int units = c.nblk * C::FPC, unit = c.bid * C::FPC + g;
  const int u_first = FUSED ? fz_first(unit, units, a.npairs) : 0;
  const int u_len = FUSED ? fz_first(unit + 1, units, a.npairs) - u_first : 0;
  const int nblocks = FUSED ? (a.npairs + units - 1) / units : (a.npairs + C::FPC - 1) / C::FPC;
  const int pb_begin = FUSED ? 0 : c.bid, pb_step = FUSED ? 1 : c.nblk;
  auto blk = [&](int pb) { return a.rev ? nblocks - 1 - pb : pb; };  // iteration index -> block of row pairs
  auto pair_of = [&](int pb) { return FUSED ? (pb < u_len ? u_first + pb : a.npairs) : blk(pb) * C::FPC + g; };
  // the T threads of a transform copy its two rows (contiguous 2N doubles) as N 16-byte chunks
  auto issue_rows = [&](int pb) {
    const int pair = pair_of(pb);
    if (pb < nblocks && pair < a.npairs) {
      const char* src = reinterpret_cast<const char*>(a.w + (size_t)(2 * pair + 1) * N);
      char* dst = reinterpret_cast<char*>(rows);
      static_for<0, E>([&](auto i_) {
        constexpr int i = decltype(i_)::value;
        cp_async16(dst + 16 * (t + T * i), src + 16 * (t + T * i));
      });
    }
    cp_async_commit();
  };
  issue_rows(pb_begin);
  if constexpr (FUSED) {
    static_for<0, NI>([&](auto i_) {
      constexpr int i = decltype(i_)::value;
      tm.template st4<6 * i>(0.0, 0.0, 0.0, 0.0);
      tm.template st2<6 * i + 4>(0.0, 0.0);
    });
    tm.fence_st();
  }
  for (int pb = pb_begin; pb < nblocks; pb += pb_step) {
    if constexpr (FUSED) {
      if (a.prefetch && t == 0 && pb + 2 < u_len)
        prefetch_l2_bulk(a.w + (size_t)(2 * (u_first + pb + 2) + 1) * N, (unsigned)(2 * N * sizeof(double)));
    } else if (a.prefetch && c.tid == 0 && pb + 2 * c.nblk < nblocks) {
      const int p0 = blk(pb + 2 * c.nblk) * C::FPC;
      const int np = (a.npairs - p0) < C::FPC ? (a.npairs - p0) : C::FPC;
      prefetch_l2_bulk(a.w + (size_t)(2 * p0 + 1) * N, (unsigned)(np * 2 * N * sizeof(double)));
    }
    const int pair = pair_of(pb);
    const bool active = pair < a.npairs;
    const int jl = 2 * pair;
    cp_async_wait_all();
    c.sync();
    double2 v[E];
    {
      const double* r0 = rows;
      const double* r1 = r0 + N;
      static_for<0, E>([&](auto e_) {
        constexpr int e = decltype(e_)::value;
        const int pos = F::template own_pos<e>(t);
        v[e] = active ? mk2(r0[pos], r1[pos]) : mk2(0.0, 0.0);
      });
    }
    c.sync();  // all rows are in registers
    // SPLIT: the landing buffer is a buffer of its own, so the next pair's rows start streaming in right away;
    // otherwise they share the exchange buffer and have to wait until the spectrum has been read back (below)
    if constexpr (C::SPLIT) issue_rows(pb + pb_step);
    F::forward(c, v, sm, tw, t);
    double2 zk[NI], zm[NI];
    if constexpr (C::SPLIT) {
      double* sd = reinterpret_cast<double*>(sm);
      F::template store_part<P - 1, 0>(v, sd, t);
      c.sync();
      static_for<0, NI>([&](auto i_) {
        constexpr int i = decltype(i_)::value;
        const int pos = halfspec_pos<C>(t + T * i);
        const int k = F::k_of_pos(pos);
        zk[i].x = sd[F::addr(pos)];
        zm[i].x = sd[F::addr(F::pos_of_k(k == 0 ? N / 2 : N - k))];
      });
      c.sync();
      F::template store_part<P - 1, 1>(v, sd, t);
      c.sync();
      static_for<0, NI>([&](auto i_) {
        constexpr int i = decltype(i_)::value;
        const int pos = halfspec_pos<C>(t + T * i);
        const int k = F::k_of_pos(pos);
        zk[i].y = sd[F::addr(pos)];
        zm[i].y = sd[F::addr(F::pos_of_k(k == 0 ? N / 2 : N - k))];
      });
      c.sync();  // the next pair's first exchange may overwrite the buffer
    } else {
      // own layout: the thread already holds Z[k] for its k < N/2 (lower-half registers); only the upper halves are
      // published, and each thread fetches the mirror values Z[N-k] of its own k (half the shared-memory traffic)
      constexpr int bl = C::bits(P - 1), rl = 1 << bl, hl = rl / 2;
      static_for<0, NI>([&](auto i_) {
        constexpr int i = decltype(i_)::value, u = i / hl, p = hl + i % hl;
        sm[F::addr(((t + T * u) << bl) | p)] = v[u * rl + p];
      });
      c.sync();
      static_for<0, NI>([&](auto i_) {
        constexpr int i = decltype(i_)::value, u = i / hl, p = i % hl;
        const int k = own_half_k<C>(t, i);
        zk[i] = v[u * rl + p];                                     // k == 0: Z[0]
        zm[i] = sm[F::addr(F::pos_of_k(k == 0 ? N / 2 : N - k))];  // k == 0: Z[N/2]
      });
      c.sync();  // the spectrum is in registers: the buffer is free for the next pair's rows
      issue_rows(pb + pb_step);
    }
    if constexpr (FUSED) {
      if (active) {
        // the half spectra of the two rows, in place (zk <- 2 X_j, zm <- 2 X_j+1); the rows kx < k0 also go to L
        static_for<0, NI>([&](auto i_) {
          constexpr int i = decltype(i_)::value;
          const int k = own_half_k<C>(t, i);
          double2 o0, o1;
          if (k == 0) {
            o0 = mk2(2.0 * zk[i].x, 2.0 * zm[i].x);
            o1 = mk2(2.0 * zk[i].y, 2.0 * zm[i].y);
          } else {
            o0 = mk2(zk[i].x + zm[i].x, zk[i].y - zm[i].y);
            o1 = mk2(zk[i].y + zm[i].y, zm[i].x - zk[i].x);
          }
          zk[i] = o0;
          zm[i] = o1;
          if (k < a.k0) st_stream4(reinterpret_cast<double2*>(a.Lpeer.p[0]) + (size_t)k * N + a.jbase + jl, o0, o1);
        });
        // forward recurrence u_j = x_j + r u_(j-1) from zero at the unit's first row, with the block's totals on the
        // way: tp = u at its last row, A = sum_m (1/r)^(M-1-m) u_m (Horner; r^(M-1) A = v at its first row when
        // nothing enters from the right), S = sum_m u_m.  State per slot in tensor memory.
        static_for<0, NI>([&](auto i_) {
          constexpr int i = decltype(i_)::value;
          const double2 rq = ld_ro2(a.rr + t + T * i);
          double2 u, A, S;
          tm.template ld6<6 * i>(u.x, u.y, A.x, A.y, S.x, S.y);
          const double2 u1 = cfma(u, rq.x, zk[i]);
          const double2 u2 = cfma(u1, rq.x, zm[i]);
          A = cfma(cfma(A, rq.y, u1), rq.y, u2);
          S = cadd(S, cadd(u1, u2));
          double2* xr = a.X + (size_t)jl * (N / 2) + t + T * i;
          st_stream2_nc(xr, u1);
          st_stream2_nc(xr + N / 2, u2);
          tm.template st4<6 * i>(u2.x, u2.y, A.x, A.y);
          tm.template st2<6 * i + 4>(S.x, S.y);
        });
        tm.fence_st();
      }
    } else if (active) {
      static_for<0, NI>([&](auto i_) {
        constexpr int i = decltype(i_)::value;
        const int k = C::SPLIT ? F::k_of_pos(halfspec_pos<C>(t + T * i)) : own_half_k<C>(t, i);
        double2 o0, o1;
        if (k == 0) {
          o0 = mk2(2.0 * zk[i].x, 2.0 * zm[i].x);
          o1 = mk2(2.0 * zk[i].y, 2.0 * zm[i].y);
        } else {
          o0 = mk2(zk[i].x + zm[i].x, zk[i].y - zm[i].y);  // 2 X_j[k]   = Z[k] + conj Z[N-k]
          o1 = mk2(zk[i].y + zm[i].y, zm[i].x - zk[i].x);  // 2 X_j+1[k] = -i (Z[k] - conj Z[N-k])
        }
        if constexpr (NAT) {
          double2* xr = a.X + (size_t)jl * (N / 2) + t + T * i;
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
  cp_async_wait_all();
  if constexpr (FUSED) {
    if (u_len > 0) {  // the block's totals
      const size_t plane = (size_t)units * (N / 2);
      static_for<0, NI>([&](auto i_) {
        constexpr int i = decltype(i_)::value;
        double2 u, A, S;
        tm.template ld6<6 * i>(u.x, u.y, A.x, A.y, S.x, S.y);
        double2* tp = a.tot + (size_t)unit * (N / 2) + t + T * i;
        tp[0] = u;
        tp[plane] = A;
        tp[2 * plane] = S;
      });
    }
    tm.template close<TMCOLS>(c);
  }
}

// ======================================== K2 ====================================================
struct K2Args {
  const double2* T;     // local spectrum rows as P blocks [g][R][NJ] (input)
  double2* V;           // own solution buffer [N/2][NJ]: the columns j this rank owns are stored here directly
  double2* S;           // staging [P blocks h][R][NJ] for the columns owned by rank h != rank (copied to V_h afterwards)
  PeerPtrs Vpeer;       // push mode: every rank's V; foreign columns are stored straight into V_h over NVLink
  int push;             // 1: push mode, 0: staged in S and moved by the copy engines
  int pieces;           // single GPU: 1 = store V as [row pair][idx][2], idx = the order in which K3's threads consume the
                        // spectrum rows: K3 then reads one contiguous, coalesced 16N-byte block per pair (bulk-
                        // prefetched into L2) instead of gathering 32-byte pieces from N/2 rows; the transposing
                        // access is K2's fire-and-forget store (full 32-byte sectors via a lane-pair shuffle)
  const double2* tw;    // twiddle tables
  const double* bbcos;  // [N]  bb*cos(kx[i])   Common.jl:120 (kx[1]=eps quirk inside)
  const double* cccos;  // [N]  cc*cos(ky[j])   (ky = kx, Common.jl:113)
  const double* ccperm; // cccos in the order the threads hold the spectrum after the forward pass:
                        // ccperm[(u*r_last + p)*T + t] = cccos[k_of_pos(((t + T*u) << b_last) | p)]  (coalesced)
  double aa;            // -2/dx^2 - 2/dy^2
  double scale;         // sign / (2 N^2): ifft normalisation, the factor 2 of the unpack, f = -w
  int NJ, log2NJ;
  int row0, nrows;      // global kx of this launch's first row, rows in this launch
  int R, rloc0, rank;   // spectrum rows per rank, local index of the launch's first row, this rank
  int prefetch;         // 0 off, 1: bulk L2 prefetch of the next row (single rank only)
  double2* Xnat = nullptr;       // non-null: store row kx as column lowslot[kx] of the natural-layout rows X[jl][N/2]
  const int* lowslot = nullptr;  // instead of V (vmk_tri.cuh: the rows kx < K0 solved here go straight back to their slots);
  int xnat_j0 = 0, xnat_nj = 0;  // only the columns j0 <= j < j0 + nj (the rank's own rows of X)
};

template <class C, bool PIECES>
VMK_HD void k2_body(const Ctx& c, const K2Args& a) {
  using F = Fft<C>;
  constexpr int N = C::N, E = C::E, T = C::T, P = C::P, M = C::M;
  constexpr int bl = C::bits(P - 1), rl = 1 << bl;
  double2* tw = F::tables(c.smem);
  F::load_tables(c, tw, a.tw);
  c.sync();
  const int g = c.tid / T, t = c.tid % T;
  double2* sm = F::xbuf(c.smem, g);
  double2* land = F::landing(c.smem, g);
  const int nblocks = (a.nrows + C::FPC - 1) / C::FPC;
  // A thread owns the same E landing slots (its pass-0 positions) for every row, so the NEXT row is copied there
  // asynchronously, element by element, without any barrier: right after the current row has been read out when
  // the landing buffer is a buffer of its own (SPLIT), else during the current row's last butterflies and stores.
  auto issue_row = [&](int rb) {
    const int row = rb * C::FPC + g;
    if (rb < nblocks && row < a.nrows) {
      // PIECES (single GPU): work item `row` is the piece index; its spectrum row is kx = k(idx = row)
      const int trow = PIECES ? F::k_of_pos(halfspec_pos<C>(row)) : a.rloc0 + row;
      static_for<0, E>([&](auto e_) {
        constexpr int e = decltype(e_)::value;
        const int j = F::template own_pos<e>(t);
        cp_async16(land + F::land_addr(j),
                   a.T + ((size_t)(j >> a.log2NJ) * a.R + trow) * a.NJ + (j & (a.NJ - 1)));
      });
    }
    cp_async_commit();
  };
  issue_row(c.bid);
  for (int rb = c.bid; rb < nblocks; rb += c.nblk) {
    const int row = rb * C::FPC + g;
    const bool active = row < a.nrows;
    // PIECES: consecutive work items (= concurrently running CTAs) take the spectrum rows in the order in which K3
    // consumes them, so their 32-byte stores fill the same 128-byte lines of V at about the same time (the L2 merges
    // them; with far-apart writers the partial lines were evicted and K2 cost +0.12 ms)
    const int kx = (PIECES && active) ? F::k_of_pos(halfspec_pos<C>(row)) : a.row0 + row;
    const bool cta_has_row0 = (a.row0 + rb * C::FPC) == 0;
    if (a.prefetch && c.tid == 0 && rb + 2 * c.nblk < nblocks && a.log2NJ == M) {  // single block: rows contiguous
      const int r0n = (rb + 2 * c.nblk) * C::FPC;
      const int nr = (a.nrows - r0n) < C::FPC ? (a.nrows - r0n) : C::FPC;
      prefetch_l2_bulk(a.T + (size_t)(a.rloc0 + r0n) * N, (unsigned)(nr * N * sizeof(double2)));
    }
    cp_async_wait_all();
    double2 v[E];
    static_for<0, E>([&](auto e_) {
      constexpr int e = decltype(e_)::value;
      v[e] = active ? land[F::land_addr(F::template own_pos<e>(t))] : mk2(0.0, 0.0);
    });
    if constexpr (C::SPLIT) issue_row(rb + c.nblk);
    F::forward(c, v, sm, tw, t);
    // ---- divide (registers hold the last-pass layout: butterfly id = t + T*u, digit p) ----------
    if (cta_has_row0) {
      // packed DC/Nyquist row: C[ky] = A^[ky] + i B^[ky]; separate, divide each by its own divisor, repack.
      // cm[] <- the value at the mirrored index (N - k) & (N - 1) of every register slot, via the exchange buffer
      double2 cm[E];
      auto mirror = [&](int e_rt, int u, int p) {
        (void)e_rt;
        return F::addr(F::pos_of_k((N - F::k_of_pos(((t + T * u) << bl) | p)) & (N - 1)));
      };
      if constexpr (C::SPLIT) {
        double* sd = reinterpret_cast<double*>(sm);
        F::template store_part<P - 1, 0>(v, sd, t);
        c.sync();
        static_for<0, E>([&](auto e_) {
          constexpr int e = decltype(e_)::value;
          cm[e].x = sd[mirror(e, e / rl, e % rl)];
        });
        c.sync();
        F::template store_part<P - 1, 1>(v, sd, t);
        c.sync();
        static_for<0, E>([&](auto e_) {
          constexpr int e = decltype(e_)::value;
          cm[e].y = sd[mirror(e, e / rl, e % rl)];
        });
      } else {
        F::template store_smem<P - 1>(v, sm, t);
        c.sync();
        static_for<0, E>([&](auto e_) {
          constexpr int e = decltype(e_)::value;
          cm[e] = sm[mirror(e, e / rl, e % rl)];
        });
      }
      if (kx == 0) {
        const double ab0 = a.aa + ld_ro(a.bbcos + 0), abn = a.aa + ld_ro(a.bbcos + N / 2);
        static_for<0, E / rl>([&](auto u_) {
          constexpr int u = decltype(u_)::value;
          const int id = t + T * u;
          // the divisor's cc cos(ky) in batches of independent, coalesced loads (ccperm holds cccos in register order):
          // one load -> two IEEE reciprocals -> next load serialised E L2 latencies on the one CTA that every solve of the
          // recurrence form waits for (the rows kx < K0)
          constexpr int CB = rl < 8 ? rl : 8;
          static_for<0, rl / CB>([&](auto b_) {
            constexpr int pb0 = decltype(b_)::value * CB;
            double ccv[CB];
            static_for<0, CB>([&](auto q_) {
              constexpr int q = decltype(q_)::value;
              ccv[q] = ld_ro(a.ccperm + (u * rl + pb0 + q) * T + t);
            });
            static_for<0, CB>([&](auto q_) {
              constexpr int p = pb0 + decltype(q_)::value;
              const int k = F::k_of_pos((id << bl) | p);
              const double2 cmv = cm[u * rl + p];
              const double2 ck = v[u * rl + p];
              const double cc = ccv[p - pb0];
              const double g0 = 0.5 * a.scale * rcp_rn(ab0 + cc), gn = 0.5 * a.scale * rcp_rn(abn + cc);
              double2 pp = cscale(mk2(ck.x + cmv.x, ck.y - cmv.y), g0);  // A^' = (C + conj Cm)/2 * g
              const double2 qq = cscale(mk2(ck.y + cmv.y, cmv.x - ck.x), gn);  // B^' = -i(C - conj Cm)/2 * g
              if (k == 0) pp = mk2(0.0, 0.0);                              // e[1,1] = 0, Common.jl:118
              v[u * rl + p] = mk2(pp.x - qq.y, pp.y + qq.x);               // A^' + i B^'
            });
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
      // all divisor loads first (independent, L1/L2 hits), then branch-free reciprocals: the literal
      // load -> add -> IEEE-reciprocal chain per element serialised 32 L2 latencies per thread
      const double ab = a.aa + ld_ro(a.bbcos + (active ? kx : 1));
      double dd[E];
      static_for<0, E / rl>([&](auto u_) {
        constexpr int u = decltype(u_)::value;
        const int kb = F::k_of_pos((t + T * u) << bl);
        static_for<0, rl>([&](auto p_) {
          constexpr int p = decltype(p_)::value;
          (void)kb;
          dd[u * rl + p] = ab + ld_ro(a.ccperm + (u * rl + p) * T + t);  // (aa + bb cos kx) + cc cos ky
        });
      });
      static_for<0, E>([&](auto i_) {
        constexpr int i = decltype(i_)::value;
        v[i] = cscale(v[i], a.scale * rcp_fast(dd[i]));
      });
    }
    F::inverse(c, v, sm, tw, t, [&] {
      if constexpr (!C::SPLIT) issue_row(rb + c.nblk);
    });
    if constexpr (PIECES) {
      // lanes 2m, 2m+1 hold columns j, j+1 of every register: one shuffle per register pair gives each lane both
      // halves of one 32-byte piece (even lane: register e, odd lane: register e+1)
      const bool odd = (t & 1) != 0;
      const size_t piece = (size_t)(active ? row : 0) * 2;
      static_for<0, E / 2>([&](auto h_) {
        constexpr int e = 2 * decltype(h_)::value;
        double2 send = odd ? v[e] : v[e + 1];
        c.shfl_xor2(send.x, send.y, 1);
        const double2 keep = odd ? v[e + 1] : v[e];
        const int j = odd ? F::template own_pos<e + 1>(t) : F::template own_pos<e>(t);
        if (active)
          st_stream4(a.V + (size_t)(j >> 1) * N + piece, odd ? send : keep, odd ? keep : send);
      });
    } else if (active && a.Xnat) {
      // recurrence form: the solved row goes straight to its slot of the natural-layout rows (this rank's columns of it)
      double2* col = a.Xnat + ld_roi(a.lowslot + kx);
      static_for<0, E>([&](auto e_) {
        constexpr int e = decltype(e_)::value;
        const int jl = F::template own_pos<e>(t) - a.xnat_j0;
        if ((unsigned)jl < (unsigned)a.xnat_nj) st_stream2(col + (size_t)jl * (N / 2), v[e]);
      });
    } else if (active) {
      static_for<0, E>([&](auto e_) {
        constexpr int e = decltype(e_)::value;
        const int j = F::template own_pos<e>(t);
        const int h = j >> a.log2NJ;
        double2* dst = (h == a.rank ? a.V + (size_t)kx * a.NJ
                        : a.push    ? reinterpret_cast<double2*>(a.Vpeer.p[h]) + (size_t)kx * a.NJ
                                    : a.S + ((size_t)h * a.R + a.rloc0 + row) * a.NJ) +
                       (j & (a.NJ - 1));
        st_stream2(dst, v[e]);
      });
    }
  }
  cp_async_wait_all();
}

// ======================================== K3 ====================================================
struct K3Args {
  const double2* T;   // local solution buffer V after K2: U[kx][jl]
  const double2* tw;
  double* psi;        // slab with halo rows
  double* lo_dst;     // where interior row 0 is mirrored: previous rank's top halo row (row NJ+1 there)
  double* hi_dst;     // where interior row NJ-1 is mirrored: next rank's bottom halo row (row 0 there)
  int NJ, npairs;
  int pieces;         // 1: T is laid out [pair][idx][2] (see K2Args::pieces): contiguous, coalesced reads
  int prefetch;       // cluster kernels: bulk L2 prefetch of the next pair's pieces
  int rev = 0;        // 1: the row pairs are taken in descending order (see K1Args::rev)
  // FUSED (vmk_tri.cuh, "fused form"): T holds the forward recurrence values; the backward recurrence, the scaling
  // and the eps correction run in K3's load stage
  const double2* rr = nullptr;   // [N/2]: (r, 1/r) per slot
  const double2* cin = nullptr;  // [2][units][N/2]: q at the block's last row, carry from the right; then [N/2]: dc
  double kc = 0.0;               // scale factor per slot = r kc (slots that keep the FFT form, r = 0, pass through)
};

// LAYOUT of the solution spectrum: 0 = rows [kx][NJ] (32-byte pieces gathered from N/2 rows), 1 = PIECES,
// 2 = natural rows [jl][N/2] in K1's slot order (vmk_tri.cuh): two contiguous 8N-byte rows per pair
template <class C, int LAYOUT, bool FUSED = false>
VMK_HD void k3_body(const Ctx& c, const K3Args& a) {
  using F = Fft<C>;
  constexpr bool PIECES = LAYOUT == 1, NAT = LAYOUT == 2;
  static_assert(!(NAT && C::SPLIT), "the slot order of the natural layout is the own-half order");
  static_assert(!FUSED || (NAT && C::T >= 32), "the fused form reads natural rows; its state is per warp");
  constexpr int N = C::N, E = C::E, T = C::T, P = C::P, NI = N / 2 / T;
  constexpr int TMCOLS = tm_cols(4 * NI, C::CT);
  TmState<4 * NI> tm;  // FUSED: per slot y (the backward recurrence value of the row above), q (what the carry from the
                       // left still contributes at the current row)
  if constexpr (FUSED) tm.template open<TMCOLS>(c);
  double2* tw = F::tables(c.smem);
  F::load_tables(c, tw, a.tw);
  c.sync();
  const int g = c.tid / T, t = c.tid % T;
  double2* sm = F::xbuf(c.smem, g);
  double2* land = F::landing(c.smem, g);
  // FUSED: the loop counter pb counts DOWN the unit's own block of pairs (see fz_first)
  const int units = c.nblk * C::FPC, unit = c.bid * C::FPC + g;
  const int u_first = FUSED ? fz_first(unit, units, a.npairs) : 0;
  const int u_len = FUSED ? fz_first(unit + 1, units, a.npairs) - u_first : 0;
  const int nblocks = FUSED ? (a.npairs + units - 1) / units : (a.npairs + C::FPC - 1) / C::FPC;
  const int pb_begin = FUSED ? 0 : c.bid, pb_step = FUSED ? 1 : c.nblk;
  auto blk = [&](int pb) { return a.rev ? nblocks - 1 - pb : pb; };  // iteration index -> block of row pairs
  auto pair_of = [&](int pb) {
    return FUSED ? (pb < u_len ? u_first + u_len - 1 - pb : a.npairs) : blk(pb) * C::FPC + g;
  };
  // SPLIT: the 32-byte pieces (U[k][j], U[k][j+1]) of the NEXT pair are gathered asynchronously into the landing
  // buffer (first halves at [idx], second halves at [N/2 + idx]; a thread only touches its own idx = t + T*i) while
  // the current pair is transformed; otherwise the gather is synchronous (the exchange buffer is all there is).
  auto issue_gather = [&](int pb) {
    if constexpr (C::SPLIT) {
      const int pair = blk(pb) * C::FPC + g;
      if (pb < nblocks && pair < a.npairs) {
        static_for<0, NI>([&](auto i_) {
          constexpr int i = decltype(i_)::value;
          const int idx = t + T * i;
          const double2* src = PIECES ? a.T + (size_t)pair * N + 2 * idx
                                      : a.T + (size_t)F::k_of_pos(halfspec_pos<C>(idx)) * a.NJ + 2 * pair;
          cp_async16(land + idx, src);
          cp_async16(land + N / 2 + idx, src + 1);
        });
      }
      cp_async_commit();
    }
  };
  issue_gather(c.bid);
  if constexpr (FUSED) {
    // state at the block's last row, from the scan: y = the carry from the right, q = r^M cu
    if (u_len > 0) {
      static_for<0, NI>([&](auto i_) {
        constexpr int i = decltype(i_)::value;
        const size_t o = (size_t)unit * (N / 2) + t + T * i;
        const double2 q = a.cin[o], y = a.cin[(size_t)units * (N / 2) + o];
        tm.template st4<4 * i>(y.x, y.y, q.x, q.y);
      });
      tm.fence_st();
    }
  }
  for (int pb = pb_begin; pb < nblocks; pb += pb_step) {
    const int pair = pair_of(pb);
    const bool active = pair < a.npairs;
    const int jl = 2 * pair;
    double2 v[E];
    if constexpr (C::SPLIT) {
      double2 ua[NI], ub[NI];
      cp_async_wait_all();
      static_for<0, NI>([&](auto i_) {
        constexpr int i = decltype(i_)::value;
        ua[i] = land[t + T * i];
        ub[i] = land[N / 2 + t + T * i];
      });
      issue_gather(pb + c.nblk);
      c.sync();  // the previous pair's last exchange has been read everywhere: the buffer may be overwritten
      // Z = U_j + i U_j+1 in position order (Z[N-k] from the conjugates), straight into the last-pass layout
      auto zval = [&](int i, bool mirror) {
        const int k = F::k_of_pos(halfspec_pos<C>(t + T * i));
        if (k == 0) return mirror ? mk2(ua[i].y, ub[i].y) : mk2(ua[i].x, ub[i].x);  // Z[N/2], Z[0] (packed row)
        return mirror ? mk2(ua[i].x + ub[i].y, ub[i].x - ua[i].y) : mk2(ua[i].x - ub[i].y, ua[i].y + ub[i].x);
      };
      auto zpos = [&](int i, bool mirror) {
        const int pos = halfspec_pos<C>(t + T * i);
        const int k = F::k_of_pos(pos);
        return F::addr(mirror ? F::pos_of_k(k == 0 ? N / 2 : N - k) : pos);
      };
      double* sd = reinterpret_cast<double*>(sm);
      if (active) {
        static_for<0, NI>([&](auto i_) {
          constexpr int i = decltype(i_)::value;
          sd[zpos(i, false)] = zval(i, false).x;
          sd[zpos(i, true)] = zval(i, true).x;
        });
      }
      c.sync();
      F::template load_part<P - 1, 0>(v, sd, t);
      c.sync();
      if (active) {
        static_for<0, NI>([&](auto i_) {
          constexpr int i = decltype(i_)::value;
          sd[zpos(i, false)] = zval(i, false).y;
          sd[zpos(i, true)] = zval(i, true).y;
        });
      }
      c.sync();
      F::template load_part<P - 1, 1>(v, sd, t);
    } else {
      if constexpr (PIECES || NAT) {
        // the pair's N/2 pieces (NAT: its two rows) are one contiguous 16N-byte block in consumption order; the next
        // pair's block is pulled into L2 by a single bulk prefetch while this pair is transformed
        if constexpr (FUSED) {
          if (t == 0 && pb + 1 < u_len)
            prefetch_l2_bulk(a.T + (size_t)(pair - 1) * N, (unsigned)(N * sizeof(double2)));
        } else if (c.tid == 0 && pb + c.nblk < nblocks) {
          const int p0 = blk(pb + c.nblk) * C::FPC;
          const int np = (a.npairs - p0) < C::FPC ? (a.npairs - p0) : C::FPC;
          prefetch_l2_bulk(a.T + (size_t)p0 * N, (unsigned)(np * N * sizeof(double2)));
        }
      }
      // two half batches (loads of a half in flight together, then its repack) keep the register peak below the cap
      constexpr int NB = NI >= 8 ? 2 : 1, NH = NI / NB;
      constexpr int bl_ = C::bits(P - 1), hl_ = 1 << (bl_ - 1);
      static_for<0, NB>([&](auto b_) {
        constexpr int b = decltype(b_)::value;
        double2 ua[NH], ub[NH];
        if (active) {
          static_for<0, NH>([&](auto i_) {
            constexpr int i = b * NH + decltype(i_)::value;
            const int idx = t + T * i;
            if constexpr (NAT) {
              const double2* src = a.T + (size_t)jl * (N / 2) + idx;
              ua[i - b * NH] = ld_stream2(src);
              ub[i - b * NH] = ld_stream2(src + N / 2);
            } else {
              const double2* src = PIECES ? a.T + (size_t)pair * N + 2 * idx
                                          : a.T + (size_t)F::k_of_pos(halfspec_pos<C>(idx)) * a.NJ + jl;
              ld_stream4(src, ua[i - b * NH], ub[i - b * NH]);
            }
          });
          if constexpr (FUSED) {
            // backward recurrence over the unit's block, last row first: with u0 the forward values K1 left (zero
            // carry into the block) and q_m = r^(m+1) cu what the true carry cu adds at row m,
            //   v_m = u0_m + q_m + r v_(m+1),   q_(m-1) = q_m / r,   psi_m = kv v_m + dc;
            // q and v at the block's last row come from the scan (state_init).  Slots that keep the FFT form along j
            // (r = 0) hold final values already: they pass through.  The constants are loaded together with the rows.
            double2 rq[NH], dc[NH];
            static_for<0, NH>([&](auto i_) {
              constexpr int ii = decltype(i_)::value, i = b * NH + ii;
              rq[ii] = ld_ro2(a.rr + t + T * i);
              dc[ii] = ld_stream2(a.cin + 2 * (size_t)units * (N / 2) + t + T * i);
            });
            static_for<0, NH>([&](auto i_) {
              constexpr int ii = decltype(i_)::value, i = b * NH + ii;
              const double kv = rq[ii].x == 0.0 ? 1.0 : rq[ii].x * a.kc;  // -r sign / (cc N)
              double2 y, q;
              tm.template ld4<4 * i>(y.x, y.y, q.x, q.y);
              const double2 v1 = cfma(y, rq[ii].x, cadd(ub[ii], q));
              q = cscale(q, rq[ii].y);
              const double2 v0 = cfma(v1, rq[ii].x, cadd(ua[ii], q));
              q = cscale(q, rq[ii].y);
              tm.template st4<4 * i>(v0.x, v0.y, q.x, q.y);
              ub[ii] = mk2(fma_(v1.x, kv, dc[ii].x), fma_(v1.y, kv, dc[ii].y));
              ua[ii] = mk2(fma_(v0.x, kv, dc[ii].x), fma_(v0.y, kv, dc[ii].y));
            });
            tm.fence_st();
          }
        }
        if constexpr (b == 0) c.sync();  // the previous pair's last exchange has been read everywhere
        // Z = U_j + i U_j+1 in position order (Z[N-k] from the conjugates), straight into the last-pass layout
        if (active) {
          static_for<0, NH>([&](auto i_) {
            constexpr int ii = decltype(i_)::value, i = b * NH + ii;
            // NAT: slot t + T i holds kx = own_half_k(t, i) (the position K1's thread t held it at)
            const int pos = NAT ? (((t + T * (i / hl_)) << bl_) | (i % hl_)) : halfspec_pos<C>(t + T * i);
            const int k = F::k_of_pos(pos);
            if (k == 0) {
              sm[F::addr(0)] = mk2(ua[ii].x, ub[ii].x);                   // Z[0]   = u0_j + i u0_j+1
              sm[F::addr(F::pos_of_k(N / 2))] = mk2(ua[ii].y, ub[ii].y);  // Z[N/2] = uN2_j + i uN2_j+1
            } else {
              sm[F::addr(pos)] = mk2(ua[ii].x - ub[ii].y, ua[ii].y + ub[ii].x);                 // U_j[k] + i U_j+1[k]
              sm[F::addr(F::pos_of_k(N - k))] = mk2(ua[ii].x + ub[ii].y, ub[ii].x - ua[ii].y);  // conj(..) + i conj(..)
            }
          });
        }
      });
      c.sync();
      F::template load_smem<P - 1>(v, sm, t);
    }
    F::inverse(c, v, sm, tw, t);
    if (active) {
      double* r0 = a.psi + (size_t)(jl + 1) * N;
      double* r1 = r0 + N;
      const bool first = (jl == 0), last = (jl + 2 == a.NJ);
      static_for<0, E>([&](auto e_) {
        constexpr int e = decltype(e_)::value;
        const int pos = F::template own_pos<e>(t);
        st_stream1(r0 + pos, v[e].x);
        st_stream1(r1 + pos, v[e].y);
        if (first) st_stream1(a.lo_dst + pos, v[e].x);
        if (last) st_stream1(a.hi_dst + pos, v[e].y);
      });
    }
  }
  cp_async_wait_all();
  if constexpr (FUSED) tm.template close<TMCOLS>(c);
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
  int rev = 0;        // 1: row blocks in descending order (see K1Args::rev)
  int rows_per_cta;   // rows marched by one thread column
  int ahead;          // rows ahead of the march that are prefetched into L2 (0 = off)
  double aa, bb;      // 1/(re dx^2), 1/(re dy^2)     Common.jl:149-150
  double gg, hh;      // 1/(4 dx dy), 1/3             Common.jl:151-152
  double dt;
};
constexpr int kK4Threads = 128;
constexpr int kK4Cols = 4;  // grid columns per thread

// x / 3 with three FP64 instructions instead of a division sequence: q = x*(1/3), one exact-residual
// correction step (Markstein); equal to the correctly rounded quotient of vm.jl:63
VMK_HD double div3(double x) {
  const double third = 1.0 / 3.0;
  const double q = x * third;
  return fma_(fma_(-3.0, q, x), third, q);
}

// MODE 0: out = r (vm_rhs);  1: wn + dt r;  2: .75 wn + .25 wt + (.25 dt) r;  3: wn/3 + (2/3) wt + ((2/3) dt) r
template <int MODE>
VMK_HD double rk_combine(double wn, double wt, double r, double dt) {
  if constexpr (MODE == 0) return r;
  if constexpr (MODE == 1) return fma_(dt, r, wt);  // stage 1: the stencil input IS wn (vm.jl:28)
  if constexpr (MODE == 2) return fma_(.25 * dt, r, fma_(.25, wt, .75 * wn));   // vm.jl:43-47
  return fma_((2. / 3.) * dt, r, fma_(2. / 3., wt, div3(wn)));                  // vm.jl:62-66
}

VMK_HD void ld4(const double* p, double& a, double& b, double& c, double& d) {
#ifdef __CUDA_ARCH__
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
#else
  a = p[0];
  b = p[1];
  c = p[2];
  d = p[3];
#endif
}
VMK_HD void st4(double* p, double a, double b, double c, double d) {
#ifdef __CUDA_ARCH__
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
#else
  p[0] = a;
  p[1] = b;
  p[2] = c;
  p[3] = d;
#endif
}

// Each thread owns four adjacent columns (one 32-byte load per row and field) and marches rows_per_cta rows with a
// rolling 3-row x 6-column window of w and psi in registers; the row after next is already in flight while the
// current one is evaluated.  i wraps periodically by index, j uses the slab's halo rows.
//
// Arithmetic: the Arakawa terms j1+j2+j3 of Common.jl:155-172 are accumulated as one chain of 10 products with
// FMAs, sharing the vertical psi differences between neighbouring columns (27 FP64 instructions per point instead
// of 45: the FP64 pipe, not HBM, was the limiter of the literal form).  The reference evaluates this loop under
// @fastmath, i.e. it leaves association and contraction to the compiler as well; results differ from the unfused
// oracle at the 1e-16 level.
template <int MODE>
VMK_HD void k4_body(const Ctx& c, const K4Args& a) {
  const int N = a.N;
  const int cols = N / kK4Cols;
  const int tx = cols < kK4Threads ? cols : kK4Threads;  // threads across i
  const int groups = kK4Threads / tx;                    // row groups per CTA
  const int ctas_x = cols / tx;
  const int bx = c.bid % ctas_x, by = a.rev ? c.nblk / ctas_x - 1 - c.bid / ctas_x : c.bid / ctas_x;
  const int i0 = kK4Cols * (bx * tx + c.tid % tx);
  const int grp = c.tid / tx;
  const int jbeg = (by * groups + grp) * a.rows_per_cta;
  if (jbeg >= a.NJ) return;
  const int jend = (jbeg + a.rows_per_cta < a.NJ) ? jbeg + a.rows_per_cta : a.NJ;
  const int im = (i0 - 1) & (N - 1), ip = (i0 + kK4Cols) & (N - 1);
  const double gh = a.gg * a.hh;

  double W[3][6], S[3][6], Wp[6], Sp[6];  // [row: j-1, j, j+1][col: i-1 .. i+4]
  auto load_row = [&](int row /*slab row incl. halo offset*/, double (&wr)[6], double (&sr)[6]) {
    const double* pw = a.w + (size_t)row * N;
    const double* ps = a.psi + (size_t)row * N;
    ld4(pw + i0, wr[1], wr[2], wr[3], wr[4]);
    ld4(ps + i0, sr[1], sr[2], sr[3], sr[4]);
    wr[0] = pw[im];
    wr[5] = pw[ip];
    sr[0] = ps[im];
    sr[5] = ps[ip];
  };
  load_row(jbeg, W[0], S[0]);  // slab row jbeg = interior row jbeg-1
  load_row(jbeg + 1, W[1], S[1]);
  load_row(jbeg + 2, W[2], S[2]);
  const bool pf_lane = a.ahead > 0 && (c.tid & 3) == 0;  // one lane per 128-byte line
  for (int jl = jbeg; jl < jend; jl++) {
    if (jl + 1 < jend) load_row(jl + 3, Wp, Sp);
    if (pf_lane && jl + a.ahead < jend) {  // turn the DRAM latency of the rows further down into an L2 hit
      const size_t off = (size_t)(jl + 2 + a.ahead) * N + i0;
      prefetch_l2(a.w + off);
      prefetch_l2(a.psi + off);
      if constexpr (MODE >= 2) prefetch_l2(a.wn + off - N);
    }
    double n0 = 0.0, n1 = 0.0, n2 = 0.0, n3 = 0.0;
    if constexpr (MODE >= 2) ld4(a.wn + (size_t)(jl + 1) * N + i0, n0, n1, n2, n3);
    const double wnv[4] = {n0, n1, n2, n3};
    double dv[6];
#pragma unroll
    for (int q = 0; q < 6; q++) dv[q] = S[2][q] - S[0][q];
    double o[4];
#pragma unroll
    for (int e = 0; e < 4; e++) {
      const double wc = W[1][e + 1], we = W[1][e + 2], ww = W[1][e], wno = W[2][e + 1], wso = W[0][e + 1];
      const double sn = S[2][e + 1], ss = S[0][e + 1], se = S[1][e + 2], sw = S[1][e];
      double acc = (we - ww) * dv[e + 1];                      // j1, Common.jl:155-158
      acc = fma_(-(wno - wso), se - sw, acc);
      acc = fma_(we, dv[e + 2], acc);                          // j2, :160-165
      acc = fma_(-ww, dv[e], acc);
      acc = fma_(-wno, S[2][e + 2] - S[2][e], acc);
      acc = fma_(wso, S[0][e + 2] - S[0][e], acc);
      acc = fma_(W[2][e + 2], sn - se, acc);                   // j3, :167-172
      acc = fma_(-W[0][e], sw - ss, acc);
      acc = fma_(-W[2][e], sn - sw, acc);
      acc = fma_(W[0][e + 2], se - ss, acc);
      const double lapx = fma_(-2.0, wc, we + ww), lapy = fma_(-2.0, wc, wno + wso);
      const double r = fma_(-gh, acc, fma_(a.aa, lapx, a.bb * lapy));  // :174-180
      o[e] = rk_combine<MODE>(wnv[e], wc, r, a.dt);
    }
    double* po = a.out + (size_t)(jl + 1) * N + i0;
    st4(po, o[0], o[1], o[2], o[3]);
    if (jl == 0) st4(a.lo_dst + i0, o[0], o[1], o[2], o[3]);
    if (jl == a.NJ - 1) st4(a.hi_dst + i0, o[0], o[1], o[2], o[3]);
#pragma unroll
    for (int q = 0; q < 6; q++) {
      W[0][q] = W[1][q];
      W[1][q] = W[2][q];
      W[2][q] = Wp[q];
      S[0][q] = S[1][q];
      S[1][q] = S[2][q];
      S[2][q] = Sp[q];
    }
  }
}

// ======================================== KS ====================================================
// Small grids (N <= 256): the whole device-resident loop of `numerical` (vm.jl:24-76) as ONE launch of one thread-block
// cluster.  At vm.jl's own size (128^2) a step is 12 dependent kernels of ~4.5 us launch latency each and less than a
// microsecond of work (profiles/r01_notes.md, "Small grids"); here the same kernel bodies run back to back, separated
// by a hardware cluster barrier (release/acquire at cluster scope orders the global-memory traffic between the CTAs;
// everything stays in L2), and the twiddle tables are loaded once.  Results are bit-identical to the 12-launch path
// (same bodies, same work decomposition per body).
struct KSArgs {
  K1Args k1[3];  // per RK3 stage: the field that is transformed (wn, wtA, wtB)
  K2Args k2;
  K3Args k3;
  K4Args k4[3];
  int k4_grid[3];  // CTAs the stand-alone K4 launch would use (its work decomposition is by CTA index)
  long long nsteps;
};

template <class C>
VMK_HD void ks_body(const Ctx& c0, const KSArgs& a) {
  using F = Fft<C>;
  Ctx c = c0;
  F::load_tables(c, F::tables(c.smem), a.k1[0].tw);
  c.sync();
  c.tables_resident = true;
  auto k4_all = [&](auto mode_, const K4Args& k, int grid) {
    constexpr int MODE = decltype(mode_)::value;
    for (int vb = c.bid; vb < grid; vb += c.nblk) {
      Ctx cv = c;
      cv.bid = vb;
      cv.nblk = grid;
      k4_body<MODE>(cv, k);
    }
  };
  for (long long step = 0; step < a.nsteps; step++) {
    static_for<0, 3>([&](auto s_) {
      constexpr int s = decltype(s_)::value;
      k1_body<C>(c, a.k1[s]);
      c.sync();  // the CTA's shared memory changes hands between the bodies
      c.cluster_sync();
      k2_body<C, true>(c, a.k2);
      c.sync();
      c.cluster_sync();
      k3_body<C, 1>(c, a.k3);
      c.sync();
      c.cluster_sync();
      k4_all(std::integral_constant<int, s + 1>{}, a.k4[s], a.k4_grid[s]);
      c.cluster_sync();
    });
  }
}

// ======================================== K6 ====================================================
// The transposes of the distributed FFT as an SM copy from a local staging buffer into the peers' buffers (NVLink
// stores, coalesced 16-byte lanes, runs of ncols*16 bytes).  Forward (K1 -> K2): rows of S owned by rank h -> block
// `rank` of T_h, columns of the K1 launch that just finished.  Backward (K2 -> K3, k2_push = 2): block h of the staged
// result rows -> rows of V_h.  Runs on a second stream beside the launch that produces the next chunk, so a stall on
// the NVLink store queue never blocks a transform (K2's own push epilogue does: its warps wait at the stores).
struct K6Args {
  const double2* src;    // local staging buffer
  PeerPtrs dst;          // every rank's destination buffer
  long long src_hstride; // elements between the source blocks of consecutive destination ranks
  long long src_off;     // first element of the launch's rows / columns inside a source block
  long long dst_off;     // ... and inside a destination buffer (the same in every peer)
  int pitch;             // row pitch of both, elements
  int rows, ncols;       // rows per destination, elements per row
  int rank, nranks;
  int order;             // 0: work items destination-major (all CTAs feed one peer at a time), 1: destinations interleaved
};
constexpr int kK6Threads = 256;
constexpr int kK6Rows = 8;  // rows per work item = independent 16-byte loads in flight per thread

VMK_HD void k6_push_body(const Ctx& c, const K6Args& a) {
  const int rgroups = (a.rows + kK6Rows - 1) / kK6Rows, ctiles = (a.ncols + kK6Threads - 1) / kK6Threads;
  const int per_dst = rgroups * ctiles, items = (a.nranks - 1) * per_dst;
  for (int it = c.bid; it < items; it += c.nblk) {
    const int q = a.order ? it % (a.nranks - 1) : it / per_dst;
    const int w = a.order ? it / (a.nranks - 1) : it % per_dst;
    const int h = (a.rank + 1 + q) % a.nranks;  // neighbour first: the ranks do not all hit one peer at a time
    const int row0 = (w / ctiles) * kK6Rows, col = (w % ctiles) * kK6Threads + c.tid;
    const double2* src = a.src + (size_t)h * a.src_hstride + a.src_off + (size_t)row0 * a.pitch + col;
    double2* dst = reinterpret_cast<double2*>(a.dst.p[h]) + a.dst_off + (size_t)row0 * a.pitch + col;
    if (col < a.ncols) {
      double2 v[kK6Rows];
#pragma unroll
      for (int u = 0; u < kK6Rows; u++)
        if (row0 + u < a.rows) v[u] = ld_stream2(src + (size_t)u * a.pitch);
#pragma unroll
      for (int u = 0; u < kK6Rows; u++)
        if (row0 + u < a.rows) st_stream2(dst + (size_t)u * a.pitch, v[u]);
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
