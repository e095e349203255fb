// vmk_cavity.cuh -- kernels of the lid-driven cavity solver (18_NS2D_Lid_Driven_Cavity/lid_driven_cavity.jl;
// SURVEY 8f row f2) that are not shared with the periodic path.
//
// The cavity fields are (nx+1) x (ny+1) node arrays without ghost cells (walls at index 1 and nx+1), i contiguous,
// element (i, j) (0-based) at i + (nx+1) j -- the caller's layout, kept on the device.
//
// Poisson solve: fps_sine (lid_driven_cavity.jl:11-21) is two DST-I passes (FFTW RODFT00) with a division by
// (2/dx^2)(cos(pi i/nx) - 1) + (2/dy^2)(cos(pi j/ny) - 1) in between.  The DST-I of a sequence is the DFT of its odd
// extension to length 2 nx, so the whole solve is the PERIODIC solve of the finite-difference path (K1 -> K2 -> K3,
// unchanged kernels, plan of size 2nx x 2ny) applied to the odd extension of the source in both directions, with
// the divisor tables filled with the cavity's expression (no eps quirk; the mean mode of an odd field is zero and is
// dropped by K2 exactly like e[1,1] = 0).  kc_extend builds the extension, kc_extract takes the interior back.
// That is 2x the transform work of a dedicated DST kernel -- a deliberate first version: every FFT kernel it runs is
// already validated on the GPU, including the cluster kernels for nx = 8192 (2 nx = 16384).
#pragma once
#include "vmk_common.cuh"

namespace vmk {

constexpr int kKCThreads = 256;

struct KCArgs {
  const double* w;    // stencil input, node array                (kc_stage: wn of stage 1 / wt of stages 2, 3)
  const double* s;    // streamfunction, node array
  const double* wn;   // point-wise input of stages 2, 3 (may alias out in stage 3)
  double* out;        // node array written (interior nodes only by kc_stage, wall nodes only by kc_bc2)
  double* slab;       // extended periodic field: (2n + 2) rows of 2n doubles, interior row j at (j + 1) * 2n
  double* part;       // per-CTA partial sums of kc_rms_partial
  double* rms;        // rms[k] written by kc_rms_final
  int n;              // nx = ny
  int nparts;
  double aa, bb;      // 1/(re dx^2), 1/(re dy^2)      lid_driven_cavity.jl:125-126
  double gg, hh;      // 1/(4 dx dy), 1/3              :127-128
  double dt;
  double dx2, dy2;    // dx^2, dy^2                    (bc2)
  double lid;         // 3/dy                          (bc2, :50)
  double count;       // (nx + 1)(ny + 1)              (rms, :112)
};

// r = -J(w, s)/3 + (1/re) lap(w) at an interior node, in the source's order (lid_driven_cavity.jl:130-157)
VMK_HD double kc_rhs_at(const KCArgs& a, const double* w, const double* s, size_t c, size_t ld) {
  const size_t e = c + 1, wst = c - 1, no = c + ld, so = c - ld;
  const double j1 = a.gg * ((w[e] - w[wst]) * (s[no] - s[so]) - (w[no] - w[so]) * (s[e] - s[wst]));
  const double j2 = a.gg * (w[e] * (s[no + 1] - s[so + 1]) - w[wst] * (s[no - 1] - s[so - 1]) -
                            w[no] * (s[no + 1] - s[no - 1]) + w[so] * (s[so + 1] - s[so - 1]));
  const double j3 = a.gg * (w[no + 1] * (s[no] - s[e]) - w[so - 1] * (s[wst] - s[so]) -
                            w[no - 1] * (s[no] - s[wst]) + w[so + 1] * (s[e] - s[so]));
  const double jac = (j1 + j2 + j3) * a.hh;
  return -jac + (a.aa * (w[e] - 2.0 * w[c] + w[wst]) + a.bb * (w[no] - 2.0 * w[c] + w[so]));
}

// MODE 1: wt = wn + dt r(wn)  (:84);  2: wt' = .75 wn + .25 wt + .25 dt r(wt)  (:93-97);
//      3: wn' = (1/3) wn + (2/3) wt + (2/3) dt r(wt)  (:106-110).  Interior nodes 2:nx, 2:ny only.
template <int MODE>
VMK_HD void kc_stage_body(const Ctx& c, const KCArgs& a) {
  const size_t ld = (size_t)a.n + 1, m = (size_t)a.n - 1, total = m * m;
  for (size_t q = (size_t)c.bid * kKCThreads + c.tid; q < total; q += (size_t)c.nblk * kKCThreads) {
    const size_t i = 1 + q % m, j = 1 + q / m, p = i + ld * j;
    const double r = kc_rhs_at(a, a.w, a.s, p, ld);
    double o = 0.0;
    if constexpr (MODE == 1) o = a.w[p] + a.dt * r;
    if constexpr (MODE == 2) o = .75 * a.wn[p] + .25 * a.w[p] + .25 * a.dt * r;
    if constexpr (MODE == 3) o = (1. / 3.) * a.wn[p] + (2. / 3.) * a.w[p] + (2. / 3.) * a.dt * r;
    a.out[p] = o;
  }
}

// Jensen wall vorticity, lid_driven_cavity.jl:38-52: left/right walls for j = 1:ny+1, then bottom/top for i = 1:nx+1
// (the corners therefore end up with the bottom/top formula; the lid adds -3/dy)
VMK_HD void kc_bc2_body(const Ctx& c, const KCArgs& a) {
  const size_t n = (size_t)a.n, ld = n + 1;
  for (size_t q = (size_t)c.bid * kKCThreads + c.tid; q < 4 * ld; q += (size_t)c.nblk * kKCThreads) {
    const size_t side = q / ld, t = q % ld;
    if (side == 0) {  // bottom row j = 0
      a.out[t] = (-4.0 * a.s[t + ld] + .5 * a.s[t + 2 * ld]) / a.dy2;
    } else if (side == 1) {  // top row j = n (the lid)
      a.out[t + ld * n] = (-4.0 * a.s[t + ld * (n - 1)] + .5 * a.s[t + ld * (n - 2)]) / a.dy2 - a.lid;
    } else if (t >= 1 && t < n) {  // left / right walls, corners excluded
      if (side == 2)
        a.out[ld * t] = (-4.0 * a.s[1 + ld * t] + .5 * a.s[2 + ld * t]) / a.dx2;
      else
        a.out[n + ld * t] = (-4.0 * a.s[n - 1 + ld * t] + .5 * a.s[n - 2 + ld * t]) / a.dx2;
    }
  }
}

// odd extension of the interior of w (walls count as zero: the DST only sees 2:nx, 2:ny) to the 2n x 2n periodic slab
VMK_HD void kc_extend_body(const Ctx& c, const KCArgs& a) {
  const size_t n = (size_t)a.n, N = 2 * n, ld = n + 1, total = N * N;
  for (size_t q = (size_t)c.bid * kKCThreads + c.tid; q < total; q += (size_t)c.nblk * kKCThreads) {
    const size_t ie = q % N, je = q / N;
    double v = 0.0;
    if (ie != 0 && ie != n && je != 0 && je != n) {
      const size_t i = ie < n ? ie : N - ie, j = je < n ? je : N - je;
      v = a.w[i + ld * j];
      if ((ie > n) != (je > n)) v = -v;
    }
    a.slab[(je + 1) * N + ie] = v;
  }
}

// sn[2:nx, 2:ny] = the periodic solution restricted to the interior (lid_driven_cavity.jl:19)
VMK_HD void kc_extract_body(const Ctx& c, const KCArgs& a) {
  const size_t n = (size_t)a.n, N = 2 * n, ld = n + 1, m = n - 1, total = m * m;
  for (size_t q = (size_t)c.bid * kKCThreads + c.tid; q < total; q += (size_t)c.nblk * kKCThreads) {
    const size_t i = 1 + q % m, j = 1 + q / m;
    a.out[i + ld * j] = a.slab[(j + 1) * N + i];
  }
}

// rms[k] = sqrt(sum((sn - sp)^2) / ((nx+1)(ny+1)))  (:111-113): fixed-order partial sums (s = sn, w = sp), then one
// thread adds the partials -- deterministic, and a different summation order than Julia's sequential loop (1e-16 level)
VMK_HD void kc_rms_partial_body(const Ctx& c, const KCArgs& a) {
  double* red = reinterpret_cast<double*>(c.smem);
  const size_t ld = (size_t)a.n + 1, total = ld * ld;
  double acc = 0.0;
  for (size_t q = (size_t)c.bid * kKCThreads + c.tid; q < total; q += (size_t)c.nblk * kKCThreads) {
    const double d = a.s[q] - a.w[q];
    acc += d * d;
  }
  red[c.tid] = acc;
  c.sync();
  for (int s = kKCThreads / 2; s > 0; s >>= 1) {
    if (c.tid < s) red[c.tid] += red[c.tid + s];
    c.sync();
  }
  if (c.tid == 0) a.part[c.bid] = red[0];
}
VMK_HD void kc_rms_final_body(const Ctx& c, const KCArgs& a) {
  if (c.bid == 0 && c.tid == 0) {
    double acc = 0.0;
    for (int i = 0; i < a.nparts; i++) acc += a.part[i];
    *a.rms = sqrt(acc / a.count);
  }
}

}  // namespace vmk
