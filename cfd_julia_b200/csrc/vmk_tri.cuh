// vmk_tri.cuh -- the j direction of the periodic Poisson solve as a cyclic constant-coefficient tridiagonal solve.
//
// Common.jl:117-123 transforms along both directions, divides by  aa + bb cos(kx) + cc cos(ky)  and transforms back.
// For a fixed kx that divisor is the symbol of the cyclic operator  (cc/2)(psi[j-1] + psi[j+1]) + b psi[j],
// b = aa + bb cos(kx), so "fft along j -> divide -> ifft along j" IS the solution of one cyclic tridiagonal system
// per kx, and with constant coefficients its inverse is the two-sided geometric kernel: with r + 1/r = -2b/cc, |r| < 1,
//     u_j = x_j + r u_(j-1),   v_j = u_j + r v_(j+1)   (both cyclic),   psi_j = -(2 r / cc) v_j
// (derivation, the handling of the reference's quirks and the accuracy model: tests/models/tri_model.py).
// Two first-order recurrences -- 12 FP64 instructions per complex value -- replace two length-N FFTs (~100), and
// recurrences split into chunks that are coupled only through ONE carry per chunk and direction.  That holds across
// GPUs too: the slab decomposition exchanges three complex numbers per kx and rank instead of transposing the spectrum
// twice (58.7 MB per rank and transpose at 8192^2 on 8 GPUs), and the spectrum never leaves the [j][kx] layout in which
// K1 produces and K3 consumes it (no 32-byte transposing stores, no PIECES layout).
//
// What the reference does differently from the exact operator, and how its numbers are kept (to ~1e-15):
//   * ky[1] = eps (Common.jl:112-113): the j-mean of a row is divided by b + cc cos(eps), not b + cc
//       -> rank-one correction from the row sum X0:  psi_j += X0 (1/d_ref - 1/d_tri) / N      (table Qs)
//   * the FP64 evaluation of the divisor carries ~|aa| 1e-16 of rounding noise, visible where |d| is small; b is taken
//     as the reference's own FP64 row constant, and the rows kx < K0 (with the packed kx = 0 / N/2 row: e[1,1] = 0,
//     near-singular kx = eps operator) keep the FFT form with the literal divisor: the slots of those rows are copied
//     into L[kx][j] by k1_body (every rank's L), solved there by k2_body beside the kernels of this file, and copied back by kt_low_body (k3_body with a
//     per-load choice of the source was 0.05 ms slower at 8192^2).
//
// Layout: X[jl][s], jl = local row, s = slot in [0, H = N/2): the order in which K1's threads hold the half spectrum
// (s = t + T i  <->  kx = own_half_k(t, i)); consecutive threads <-> consecutive 16-byte slots in every kernel here.
//
//   kt_totals_body   per (chunk of 32 rows, slot): u0 (zero carry-in) in registers -> tp = u0 at the chunk's last row,
//                    al = sum r^m u0_m, xs = sum x_m
//   kt_scan_body<0>  per slot: the rank's totals (TP, AL, X0) from its chunks' -> G of every rank
//   [cross-rank barrier]
//   kt_scan_body<1>  per slot: carries into the rank from all ranks' totals (cyclic closure), then into every chunk
//   (one rank: kt_scan_body<2> does both in one launch)
//   kt_solve_body    per (chunk, slot): u, v in registers (in place), scaled, + the eps correction
//   kt_low_body      X[.][slots of kx < K0] <- L
#pragma once
#include "vmk_common.cuh"

namespace vmk {

constexpr int kTriCH = 32;        // rows per register-resident chunk
constexpr int kTriThreads = 128;  // slots per CTA
constexpr int kTriFTab = 7;       // fused form, doubles per slot: g0 = r / (1 - r^2), then (R, Gam(R), r^(M-1)) of a block of
                                  // M = 2 Lmin rows and of a block of M = 2 Lmin + 2 rows, Lmin = npairs / units
constexpr int kTriTab = 10;       // doubles per slot: r, R = r^CH, RJ = r^NJ, W = 1/(1 - r^N), Gam(R), Gam(RJ), Kv, Qs,
                                  // Rg = R^cpg, Gam(Rg) -- Gam(x) = r (1 - x^2) / (1 - r^2)

struct KTArgs {
  double2* X;          // [NJ][H], solved in place
  const double* tab;   // [kTriTab][H]
  const int* lowrow;   // [H]: row of L for the slots of kx < K0, -1 otherwise
  double2* tot;        // [3][nch][H]: tp, al, xs per local chunk
  double2* cin;        // [2][nch][H]: carry into the chunk from the left (u) and from the right (v); then [H]: dc
  double2* G;          // [P][3][H]: TP, alpha, X0 of every rank (each rank writes its own block into every rank's G)
  PeerPtrs Gpeer;
  double2* L;          // [K0][N]: low rows, all j (filled by K1, solved in place by K2)
  int H, NJ, nch, N, j0, rank, nranks;
  double sign;         // +1: solve for f, -1: for -f (Common.jl:134)
  int rev = 0;         // 1: totals / solve take the chunks in descending order (see K1Args::rev)
  const int* lowslot = nullptr;  // [k0]: slot of kx
  int k0 = 0;
  // fused form (kt_scanf_body): blocks of rows = the units of the fused K1 / K3 (fz_first), two possible lengths
  const double* ftab = nullptr;  // [kTriFTab][H]
  int units = 0, npairs = 0;
};


VMK_HD void kt_totals_body(const Ctx& c, const KTArgs& a) {
  const int tiles = (a.H + kTriThreads - 1) / kTriThreads;
  const int items = a.nch * tiles;
  for (int it0 = c.bid; it0 < items; it0 += c.nblk) {
    const int it = a.rev ? items - 1 - it0 : it0;
    const int ch = it / tiles, s = (it % tiles) * kTriThreads + c.tid;
    if (s >= a.H) continue;
    const double2* xp = a.X + (size_t)ch * kTriCH * a.H + s;
    double2 x[kTriCH];
#pragma unroll
    for (int m = 0; m < kTriCH; m++) x[m] = ld_stream2(xp + (size_t)m * a.H);
    const double r = ld_ro(a.tab + s);
    double2 xs = mk2(0.0, 0.0), u = mk2(0.0, 0.0);
#pragma unroll
    for (int m = 0; m < kTriCH; m++) {
      xs = cadd(xs, x[m]);
      u = cfma(u, r, x[m]);
      x[m] = u;
    }
    double2 al = mk2(0.0, 0.0);
#pragma unroll
    for (int m = kTriCH - 1; m >= 0; m--) al = cfma(al, r, x[m]);
    double2* t = a.tot + (size_t)ch * a.H + s;
    const size_t plane = (size_t)a.nch * a.H;
    t[0] = u;
    t[plane] = al;
    t[2 * plane] = xs;
  }
}

// ---- the scan over a rank's chunks -----------------------------------------------------------------------------------
// A block of consecutive rows is summarised, for one slot, by (tp, al, xs): tp = u at its last row and al = v at its
// first row when nothing enters the block from outside, xs = the sum of its x; with R = r^rows and
// Gam = r (1 - R^2) / (1 - r^2) the block turns carries (cu from the left, cv from the right) into
//     u at its last row = tp + R cu,      v at its first row = al + Gam cu + R cv.
// Blocks compose (Horner in R), so the same two-line recursion runs chunk -> group of chunks -> rank -> all ranks
// upwards, and the carries run back down.  One CTA handles kTriScanSlots slots x up to kTriScanGroups groups of chunks:
// the groups of a slot are scanned through shared memory by the group-0 thread, everything else is parallel, and the
// loads of a group's chunk totals are issued in batches of 8 (a per-slot sequential loop over 256 chunks with one
// dependent L2 round trip per chunk took 0.2 ms at 8192^2).
constexpr int kTriScanSlots = 32, kTriScanGroups = 16, kTriScanThreads = kTriScanSlots * kTriScanGroups;
constexpr int kTriScanSmem = kTriScanThreads * 5 * (int)sizeof(double2);
constexpr int kTriScanFMaxUnits = 3072;  // blocks of row pairs kt_scanf_body keeps the lengths of
constexpr int kTriScanFSmem = kTriScanThreads * 4 * (int)sizeof(double2) + kTriScanFMaxUnits * (int)sizeof(int);

VMK_HD int tri_scan_groups(int nch) { return nch >= 8 * kTriScanGroups ? kTriScanGroups : (nch >= 8 ? nch / 8 : 1); }

// PHASE 0: rank totals -> G of every rank (before the cross-rank barrier); 1: carries from G (after it);
// 2: both in one launch, for a single rank (no G)
template <int PHASE>
VMK_HD void kt_scan_body(const Ctx& c, const KTArgs& a) {
  const int H = a.H, P = a.nranks;
  const int ngr = tri_scan_groups(a.nch), cpg = a.nch / ngr;
  const int lx = c.tid % kTriScanSlots, g = c.tid / kTriScanSlots;
  const int s = c.bid * kTriScanSlots + lx;
  const bool on = s < H && g < ngr;
  double2* sm = reinterpret_cast<double2*>(c.smem);  // [5][groups][slots]: tp, al, xs, then cu, cv
  auto at = [&](int q, int gg) -> double2& { return sm[(q * kTriScanGroups + gg) * kTriScanSlots + lx]; };
  const size_t plane = (size_t)a.nch * H;
  const double R = on ? ld_ro(a.tab + H + s) : 0.0;
  const double gam = on ? ld_ro(a.tab + 4 * H + s) : 0.0;
  // R and Gam of a group of cpg chunks and of the whole rank (tables: rounded once from long double)
  const double Rg = on ? ld_ro(a.tab + 8 * H + s) : 0.0, Gg = on ? ld_ro(a.tab + 9 * H + s) : 0.0;
  const int ch0 = g * cpg;
  // upwards: the group's (tp, al, xs)
  if (on) {
    double2 lp = mk2(0.0, 0.0), al = mk2(0.0, 0.0), xs = mk2(0.0, 0.0);
    double rp = 1.0;
    for (int b0 = 0; b0 < cpg; b0 += 8) {
      double2 t0[8], t1[8], t2[8];
#pragma unroll
      for (int i = 0; i < 8; i++) {
        if (b0 + i < cpg) {
          const double2* t = a.tot + (size_t)(ch0 + b0 + i) * H + s;
          t0[i] = t[0];
          t1[i] = t[plane];
          t2[i] = t[2 * plane];
        }
      }
#pragma unroll
      for (int i = 0; i < 8; i++) {
        if (b0 + i < cpg) {
          const double2 z = cfma(lp, gam, t1[i]);
          al = mk2(fma_(z.x, rp, al.x), fma_(z.y, rp, al.y));
          rp *= R;
          lp = cfma(lp, R, t0[i]);
          xs = cadd(xs, t2[i]);
        }
      }
    }
    at(0, g) = lp;
    at(1, g) = al;
    at(2, g) = xs;
  }
  c.sync();
  if (on && g == 0) {
    const double RJ = ld_ro(a.tab + 2 * H + s), GJ = ld_ro(a.tab + 5 * H + s);
    // the rank's own totals
    double2 lp = mk2(0.0, 0.0), al = mk2(0.0, 0.0), xs = mk2(0.0, 0.0);
    double rp = 1.0;
    for (int gg = 0; gg < ngr; gg++) {
      const double2 z = cfma(lp, Gg, at(1, gg));
      al = mk2(fma_(z.x, rp, al.x), fma_(z.y, rp, al.y));
      rp *= Rg;
      lp = cfma(lp, Rg, at(0, gg));
      xs = cadd(xs, at(2, gg));
    }
    if constexpr (PHASE == 0) {
      for (int q = 0; q < P; q++) {
        double2* gp = reinterpret_cast<double2*>(a.Gpeer.p[q]) + (size_t)a.rank * 3 * H + s;
        gp[0] = lp;
        gp[H] = al;
        gp[2 * H] = xs;
      }
    } else {
      // carries into the rank: cu = W sum_{i=1..P} RJ^(i-1) TP[rank-i]; the v carry needs every rank's own cu:
      // TM[q] = AL[q] + GJ cu[q];  cv = W sum_{i=1..P} RJ^(i-1) TM[rank+i]
      const double W = ld_ro(a.tab + 3 * H + s), qs = ld_ro(a.tab + 7 * H + s) * a.sign;
      double2 cu_own, cv = mk2(0.0, 0.0), x0 = mk2(0.0, 0.0);
      if constexpr (PHASE == 2) {
        cu_own = cscale(lp, W);
        cv = cscale(cfma(cu_own, GJ, al), W);
        x0 = xs;
      } else {
        cu_own = mk2(0.0, 0.0);
        for (int i = P; i >= 1; i--) {
          const int q = (a.rank + i) % P;
          double2 cuq = mk2(0.0, 0.0);
          for (int k = P; k >= 1; k--) cuq = cfma(cuq, RJ, a.G[(size_t)((q - k + 2 * P) % P) * 3 * H + s]);
          cuq = cscale(cuq, W);
          if (i == P) cu_own = cuq;  // q == rank
          cv = cfma(cv, RJ, cfma(cuq, GJ, a.G[(size_t)q * 3 * H + H + s]));
          x0 = cadd(x0, a.G[(size_t)q * 3 * H + 2 * H + s]);
        }
        cv = cscale(cv, W);
      }
      a.cin[2 * plane + s] = mk2(x0.x * qs, x0.y * qs);  // the eps correction, already scaled and signed
      // downwards: carries into the groups
      double2 acc = cu_own;
      for (int gg = 0; gg < ngr; gg++) {
        at(3, gg) = acc;
        acc = cfma(acc, Rg, at(0, gg));
      }
      acc = cv;
      for (int gg = ngr - 1; gg >= 0; gg--) {
        at(4, gg) = acc;
        acc = cfma(acc, Rg, cfma(at(3, gg), Gg, at(1, gg)));
      }
    }
  }
  if constexpr (PHASE == 0) return;
  c.sync();
  if (on) {
    // carries into the group's chunks: u carries ascending (kept in registers), then v carries descending
    double2 cu = at(3, g), cv = at(4, g);
    for (int b0 = 0; b0 < cpg; b0 += 8) {
      double2 t0[8];
#pragma unroll
      for (int i = 0; i < 8; i++)
        if (b0 + i < cpg) t0[i] = a.tot[(size_t)(ch0 + b0 + i) * H + s];
#pragma unroll
      for (int i = 0; i < 8; i++) {
        if (b0 + i < cpg) {
          a.cin[(size_t)(ch0 + b0 + i) * H + s] = cu;
          cu = cfma(cu, R, t0[i]);
        }
      }
    }
    for (int b0 = cpg; b0 > 0; b0 -= 8) {
      double2 t1[8], cuv[8];
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const int cl = b0 - 1 - i;
        if (cl >= 0) {
          t1[i] = a.tot[plane + (size_t)(ch0 + cl) * H + s];
          cuv[i] = a.cin[(size_t)(ch0 + cl) * H + s];
        }
      }
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const int cl = b0 - 1 - i;
        if (cl >= 0) {
          a.cin[plane + (size_t)(ch0 + cl) * H + s] = cv;
          cv = cfma(cv, R, cfma(cuv[i], gam, t1[i]));
        }
      }
    }
  }
}

// ---- the fused form: K1 runs the forward recurrence, K3 the backward one (vmk_kernels.cuh, FUSED) ---------------------
// The spectrum then crosses HBM twice per solve (K1's store, K3's load) instead of five times.  K1's unit q (a CTA's
// transform group) owns a contiguous block of row pairs, starts the forward recurrence from zero at its first row and
// leaves, per slot, tp = u at the block's last row, A = sum_m (1/r)^(M-1-m) u_m (so al = r^(M-1) A) and S = sum_m u_m
// (the block's sum of x is (1 - r) S + r tp).  This kernel is kt_scan_body<2> for those blocks: the same two-line
// composition, with R and Gam per block (two possible lengths) from tables and per group of blocks on the fly
// (Gam(R) = g0 (1 - R^2)).  It leaves what K3's unit starts from at the block's LAST row:
//   cin[0][q] = r^M cu_q (what the carry from the left still adds there),  cin[1][q] = cv_q,  cin[2] = dc.
template <int DUMMY = 0>
VMK_HD void kt_scanf_body(const Ctx& c, const KTArgs& a) {
  const int H = a.H, U = a.units;
  const int cpg = (U + kTriScanGroups - 1) / kTriScanGroups;
  const int lx = c.tid % kTriScanSlots, g = c.tid / kTriScanSlots;
  const int s = c.bid * kTriScanSlots + lx;
  const bool on = s < H;
  // [4][groups][slots]: tp, al, xs -> cu, (Rg, -), with cv taking tp's place on the way down; then the blocks' lengths
  double2* sm = reinterpret_cast<double2*>(c.smem);
  int* lens = reinterpret_cast<int*>(sm + 4 * kTriScanThreads);
  auto at = [&](int q, int gg) -> double2& { return sm[(q * kTriScanGroups + gg) * kTriScanSlots + lx]; };
  const size_t plane = (size_t)U * H;
  const int lmin = a.npairs / U;
  for (int q = c.tid; q < U; q += kTriScanThreads) lens[q] = fz_first(q + 1, U, a.npairs) - fz_first(q, U, a.npairs);
  c.sync();
  const double r = on ? ld_ro(a.tab + s) : 0.0, g0 = on ? ld_ro(a.ftab + s) : 0.0;
  const double Ra = on ? ld_ro(a.ftab + H + s) : 0.0, Ga = on ? ld_ro(a.ftab + 2 * H + s) : 0.0;
  const double Pa = on ? ld_ro(a.ftab + 3 * H + s) : 0.0;
  const double Rb = on ? ld_ro(a.ftab + 4 * H + s) : 0.0, Gb = on ? ld_ro(a.ftab + 5 * H + s) : 0.0;
  const double Pb = on ? ld_ro(a.ftab + 6 * H + s) : 0.0;
  const int ch0 = g * cpg, ch1 = ch0 + cpg < U ? ch0 + cpg : U;
  // upwards: the group's (tp, al, xs) and R
  {
    double2 lp = mk2(0.0, 0.0), al = mk2(0.0, 0.0), xs = mk2(0.0, 0.0);
    double rp = 1.0;
    if (on) {
      for (int b0 = ch0; b0 < ch1; b0 += 8) {
        double2 t0[8], t1[8], t2[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
          if (b0 + i < ch1 && lens[b0 + i] > 0) {
            const double2* t = a.tot + (size_t)(b0 + i) * H + s;
            t0[i] = t[0];
            t1[i] = t[plane];
            t2[i] = t[2 * plane];
          }
        }
#pragma unroll
        for (int i = 0; i < 8; i++) {
          if (b0 + i < ch1 && lens[b0 + i] > 0) {
            const bool lg = lens[b0 + i] > lmin;
            const double R = lg ? Rb : Ra, gam = lg ? Gb : Ga, pw = lg ? Pb : Pa;
            const double2 z = cfma(lp, gam, cscale(t1[i], pw));
            al = mk2(fma_(z.x, rp, al.x), fma_(z.y, rp, al.y));
            rp *= R;
            lp = cfma(lp, R, t0[i]);
            // the block's sum of x = (1 - r) S + r tp
            xs = mk2(xs.x + fma_(t2[i].x, 1.0 - r, r * t0[i].x), xs.y + fma_(t2[i].y, 1.0 - r, r * t0[i].y));
          }
        }
      }
    }
    at(0, g) = lp;
    at(1, g) = al;
    at(2, g) = xs;
    at(3, g) = mk2(rp, 0.0);
  }
  c.sync();
  if (on && g == 0) {
    double2 lp = mk2(0.0, 0.0), al = mk2(0.0, 0.0), xs = mk2(0.0, 0.0);
    double rp = 1.0;
    for (int gg = 0; gg < kTriScanGroups; gg++) {
      const double Rg = at(3, gg).x, Gg = g0 * (1.0 - Rg * Rg);
      const double2 z = cfma(lp, Gg, at(1, gg));
      al = mk2(fma_(z.x, rp, al.x), fma_(z.y, rp, al.y));
      rp *= Rg;
      lp = cfma(lp, Rg, at(0, gg));
      xs = cadd(xs, at(2, gg));
    }
    // cyclic closure over the whole domain (one rank: NJ = N): W = 1 / (1 - r^N), GJ = Gam(r^N)
    const double W = ld_ro(a.tab + 3 * H + s), GJ = ld_ro(a.tab + 5 * H + s), qs = ld_ro(a.tab + 7 * H + s) * a.sign;
    const double2 cu_own = cscale(lp, W);
    const double2 cv = cscale(cfma(cu_own, GJ, al), W);
    a.cin[2 * plane + s] = mk2(xs.x * qs, xs.y * qs);  // the eps correction, already scaled and signed
    double2 acc = cu_own;
    for (int gg = 0; gg < kTriScanGroups; gg++) {  // carry from the left into every group (takes xs's place)
      const double Rg = at(3, gg).x;
      at(2, gg) = acc;
      acc = cfma(acc, Rg, at(0, gg));
    }
    acc = cv;
    for (int gg = kTriScanGroups - 1; gg >= 0; gg--) {  // carry from the right (takes tp's place)
      const double Rg = at(3, gg).x, Gg = g0 * (1.0 - Rg * Rg);
      const double2 cug = at(2, gg);
      at(0, gg) = acc;
      acc = cfma(acc, Rg, cfma(cug, Gg, at(1, gg)));
    }
  }
  c.sync();
  if (on) {
    // carries into the group's blocks: u carries ascending, then v carries descending
    double2 cu = at(2, g), cv = at(0, g);
    for (int b0 = ch0; b0 < ch1; b0 += 8) {
      double2 t0[8];
#pragma unroll
      for (int i = 0; i < 8; i++)
        if (b0 + i < ch1 && lens[b0 + i] > 0) t0[i] = a.tot[(size_t)(b0 + i) * H + s];
#pragma unroll
      for (int i = 0; i < 8; i++) {
        if (b0 + i < ch1) {
          a.cin[(size_t)(b0 + i) * H + s] = cu;
          if (lens[b0 + i] > 0) cu = cfma(cu, lens[b0 + i] > lmin ? Rb : Ra, t0[i]);
        }
      }
    }
    for (int b0 = ch1; b0 > ch0; b0 -= 8) {
      double2 t1[8], cuv[8];
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const int q = b0 - 1 - i;
        if (q >= ch0) {
          cuv[i] = a.cin[(size_t)q * H + s];
          if (lens[q] > 0) t1[i] = a.tot[plane + (size_t)q * H + s];
        }
      }
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const int q = b0 - 1 - i;
        if (q >= ch0) {
          const bool lg = lens[q] > lmin;
          const double R = lg ? Rb : Ra, gam = lg ? Gb : Ga, pw = lg ? Pb : Pa;
          a.cin[plane + (size_t)q * H + s] = cv;
          a.cin[(size_t)q * H + s] = cscale(cuv[i], R);  // q at the block's last row: r^M cu
          if (lens[q] > 0) cv = cfma(cv, R, cfma(cuv[i], gam, cscale(t1[i], pw)));
        }
      }
    }
  }
}

VMK_HD void kt_solve_body(const Ctx& c, const KTArgs& a) {
  const int tiles = (a.H + kTriThreads - 1) / kTriThreads;
  const int items = a.nch * tiles;
  const size_t plane = (size_t)a.nch * a.H;
  for (int it0 = c.bid; it0 < items; it0 += c.nblk) {
    const int it = a.rev ? items - 1 - it0 : it0;
    const int ch = it / tiles, s = (it % tiles) * kTriThreads + c.tid;
    if (s >= a.H) continue;
    double2* xp = a.X + (size_t)ch * kTriCH * a.H + s;
    if (ld_roi(a.lowrow + s) >= 0) continue;  // solved by k2_body in L and copied back by kt_low_body
    double2 x[kTriCH];
#pragma unroll
    for (int m = 0; m < kTriCH; m++) x[m] = ld_stream2(xp + (size_t)m * a.H);
    const double r = ld_ro(a.tab + s), kv = ld_ro(a.tab + 6 * a.H + s) * a.sign;
    double2 y = a.cin[(size_t)ch * a.H + s];
    const double2 cm = a.cin[plane + (size_t)ch * a.H + s], dc = a.cin[2 * plane + s];
#pragma unroll
    for (int m = 0; m < kTriCH; m++) {
      y = cfma(y, r, x[m]);
      x[m] = y;
    }
    y = cm;
#pragma unroll
    for (int m = kTriCH - 1; m >= 0; m--) {
      y = cfma(y, r, x[m]);
      st_stream2(xp + (size_t)m * a.H, mk2(fma_(y.x, kv, dc.x), fma_(y.y, kv, dc.y)));
    }
  }
}

// the rows kx < k0, solved in L by K2: X[jl][slot(kx)] = L[kx][j0 + jl]   (k0 <= 64 slots x NJ rows: megabytes)
VMK_HD void kt_low_body(const Ctx& c, const KTArgs& a) {
  const int total = a.k0 * a.NJ;
  for (int q = c.bid * kTriThreads + c.tid; q < total; q += c.nblk * kTriThreads) {
    const int k = q / a.NJ, jl = q % a.NJ;  // consecutive threads read consecutive j of one row of L
    a.X[(size_t)jl * a.H + ld_roi(a.lowslot + k)] = ld_stream2(a.L + (size_t)k * a.N + a.j0 + jl);
  }
}

}  // namespace vmk
