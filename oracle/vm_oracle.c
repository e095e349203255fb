/*
 * vm_oracle.c -- CPU restatement of the CFD_Julia vortex-merger hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT THE PRODUCT.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (cfd_julia_b200/libvmk.so) never links, loads or calls anything in oracle/.
 *
 * What it restates (file:line relative to the reference checkout):
 *   orc_fps            Common.jl:97-125   periodic FFT Poisson solve (eps quirk :112, ky=kx :113,
 *                                         zero mode :118, cos divisor :119-121, ifft+real :123)
 *   orc_vm_rhs         Common.jl:132-182  f=-w, fps, psi ghost fill (order :138-146), Arakawa + Laplacian
 *   orc_numerical      19_NS2D_Vortex_Merger/vm.jl:12-90 (twin tgv.jl:13-79) SSP-RK3 loop
 *   orc_vm_ic          Common.jl:208-219 + vm.jl:121-128 (ghost fill order of main)
 *   orc_exact_tgv      19_NS2D_Vortex_Merger/tgv.jl:82-90
 *   orc_ps_fft         12_Poisson_Solver_FFT/fft_p.jl:8-42
 *   orc_l2norm_bnds    Common.jl:234-237
 *
 * Canonical arithmetic: source order, left to right, IEEE double, no FMA contraction
 * (build with -ffp-contract=off; see oracle/Makefile).  The reference's @fastmath leaves
 * the exact association to LLVM; this file fixes it to the written order.
 *
 * Third-party arithmetic: the reference's FFT is FFTW.jl -> libfftw3 (un-vendored, version
 * unpinned: the reference has no Project.toml/Manifest.toml).  It is restated here as the
 * unnormalised DFT  X[k] = sum_n x[n] exp(-2 pi i n k / N)  (forward) and its conjugate
 * scaled by 1/N (inverse), which is FFTW's published definition, evaluated by an iterative
 * radix-2 Cooley-Tukey with twiddles rounded from long double.
 *
 * Parity pins (tests/test_oracle.py):
 *   (1) the five L2 errors of fft_p.jl hard-coded at 13_Poisson_Solver_FFT_Spectral/specrtral_vs_FDM/order.jl:13
 *       (orc_ps_fft reproduces all five);
 *   (2) outputs of the REFERENCE'S OWN CODE for the whole step: its two Python twins of script 19
 *       (19_NS2D_Vortex_Merger/Python_Vectorized/fdm_vortex_merge_vectorized.py:31-129,221-256 and
 *       Python/fdm_vortex_merger.py), executed unmodified in the build container with a pyfftw->numpy.fft shim by
 *       tests/golden/make_ref_fixtures.py; committed as tests/golden/ref_py_*.npz (vorticity + streamfunction after
 *       10-50 RK3 steps at 32^2..128^2, one rhs and one Poisson solve on white noise).  This file agrees with them
 *       to 4e-16..4e-15 (gate 1e-12);
 *   (3) the analytic Taylor-Green solution (tgv.jl:87) and the independent numpy restatement oracle/oracle_np.py.
 * The Julia scripts themselves cannot run here (no julia, no libfftw3, un-vendored Unroll/Utils packages), so there
 * is no oracle/_ref build; the Julia-vs-Python-twin differences are listed in make_ref_fixtures.py (rounding level).
 *
 * Layout: all arrays are column-major like Julia.  "ghosted" = (nx+2) x (ny+2) doubles,
 * element (i,j) 1-based at [(i-1) + (nx+2)*(j-1)].
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { double re, im; } cplx;

static int ilog2_exact(int64_t n) {
  int m = 0;
  if (n <= 0) return -1;
  while ((((int64_t)1) << m) < n) m++;
  return ((((int64_t)1) << m) == n) ? m : -1;
}

/* ---- 1-D radix-2 FFT on a contiguous vector, in place, unnormalised ---------------- */
typedef struct {
  int n, m;
  cplx *tw;     /* tw[k] = exp(-2 pi i k / n), k < n/2 */
  int32_t *rev; /* bit reversal */
} fft_plan;

static int fft_plan_init(fft_plan *p, int n) {
  p->n = n;
  p->m = ilog2_exact(n);
  if (p->m < 0) return 1;
  p->tw = (cplx *)malloc(sizeof(cplx) * (size_t)(n / 2 + 1));
  p->rev = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
  if (!p->tw || !p->rev) return 2;
  const long double tau = 6.283185307179586476925286766559005768L;
  for (int k = 0; k < n / 2; k++) {
    long double a = tau * (long double)k / (long double)n;
    p->tw[k].re = (double)cosl(a);
    p->tw[k].im = (double)(-sinl(a));
  }
  for (int i = 0; i < n; i++) {
    int r = 0;
    for (int b = 0; b < p->m; b++)
      if (i & (1 << b)) r |= 1 << (p->m - 1 - b);
    p->rev[i] = r;
  }
  return 0;
}

static void fft_plan_free(fft_plan *p) {
  free(p->tw);
  free(p->rev);
  p->tw = NULL;
  p->rev = NULL;
}

/* sign = -1 forward, +1 backward (unnormalised) */
static void fft1d(const fft_plan *p, cplx *a, int sign) {
  const int n = p->n;
  for (int i = 0; i < n; i++) {
    int r = p->rev[i];
    if (r > i) {
      cplx t = a[i];
      a[i] = a[r];
      a[r] = t;
    }
  }
  for (int len = 2; len <= n; len <<= 1) {
    const int half = len >> 1, step = n / len;
    for (int base = 0; base < n; base += len) {
      for (int k = 0; k < half; k++) {
        const cplx w = p->tw[k * step];
        const double wr = w.re, wi = (sign < 0) ? w.im : -w.im;
        cplx *x = &a[base + k], *y = &a[base + k + half];
        const double tr = wr * y->re - wi * y->im;
        const double ti = wr * y->im + wi * y->re;
        y->re = x->re - tr;
        y->im = x->im - ti;
        x->re = x->re + tr;
        x->im = x->im + ti;
      }
    }
  }
}

/* 2-D in-place transform of an nx x ny column-major complex array (dim 1 contiguous). */
static int fft2d(int nx, int ny, cplx *a, int sign) {
  fft_plan px, py;
  if (fft_plan_init(&px, nx)) return 1;
  if (fft_plan_init(&py, ny)) { fft_plan_free(&px); return 1; }
#pragma omp parallel for schedule(static)
  for (int j = 0; j < ny; j++) fft1d(&px, a + (size_t)j * nx, sign);
  /* dim 2: gather blocks of columns-of-the-transposed view for cache friendliness */
  enum { BLK = 8 };
#pragma omp parallel
  {
    cplx *buf = (cplx *)malloc(sizeof(cplx) * (size_t)ny * BLK);
#pragma omp for schedule(static)
    for (int i0 = 0; i0 < nx; i0 += BLK) {
      const int nb = (nx - i0 < BLK) ? nx - i0 : BLK;
      for (int j = 0; j < ny; j++)
        for (int b = 0; b < nb; b++) buf[(size_t)b * ny + j] = a[(size_t)j * nx + i0 + b];
      for (int b = 0; b < nb; b++) fft1d(&py, buf + (size_t)b * ny, sign);
      for (int j = 0; j < ny; j++)
        for (int b = 0; b < nb; b++) a[(size_t)j * nx + i0 + b] = buf[(size_t)b * ny + j];
    }
    free(buf);
  }
  fft_plan_free(&px);
  fft_plan_free(&py);
  return 0;
}

/* ---- wavenumber table, Common.jl:106-113 ------------------------------------------- */
static void wavenumbers(int nx, double eps, double *kx) {
  const double hx = 2.0 * M_PI / (double)nx; /* Common.jl:106 */
  for (int i = 1; i <= nx / 2; i++) {        /* :108-111, 1-based i */
    kx[i - 1] = hx * (double)(i - 1);
    kx[i + nx / 2 - 1] = hx * (double)(i - nx / 2 - 1);
  }
  kx[0] = eps; /* :112 */
}

/* exported for the tests: the divisor tables the product must reproduce bit for bit */
int orc_divisor_tables(int nx, int ny, double dx, double dy, double eps,
                       double *aa_out, double *bbcos, double *cccos) {
  if (nx != ny) return 1; /* Common.jl:113 ky = kx */
  double *kx = (double *)malloc(sizeof(double) * (size_t)nx);
  wavenumbers(nx, eps, kx);
  const double aa = -2.0 / (dx * dx) - 2.0 / (dy * dy);
  const double bb = 2.0 / (dx * dx);
  const double cc = 2.0 / (dy * dy);
  *aa_out = aa;
  for (int i = 0; i < nx; i++) bbcos[i] = bb * cos(kx[i]);
  for (int j = 0; j < ny; j++) cccos[j] = cc * cos(kx[j]);
  free(kx);
  return 0;
}

/* core of fps / ps_fft: f is nx x ny (leading dimension ldf), out is real(ifft(..)) nx x ny */
static int poisson_core(int nx, int ny, double dx, double dy, const double *f, int ldf,
                        double *out, int ldo, double eps) {
  if (nx != ny) return 1;
  if (ilog2_exact(nx) < 1) return 1;
  const size_t n2 = (size_t)nx * ny;
  cplx *data = (cplx *)malloc(sizeof(cplx) * n2);
  double *kx = (double *)malloc(sizeof(double) * (size_t)nx);
  double *ck = (double *)malloc(sizeof(double) * (size_t)nx);
  if (!data || !kx || !ck) return 2;
  const double aa = -2.0 / (dx * dx) - 2.0 / (dy * dy); /* :101 */
  const double bb = 2.0 / (dx * dx);                     /* :102 */
  const double cc = 2.0 / (dy * dy);                     /* :103 */
  wavenumbers(nx, eps, kx);
  for (int i = 0; i < nx; i++) ck[i] = cos(kx[i]);
#pragma omp parallel for schedule(static)
  for (int j = 0; j < ny; j++)
    for (int i = 0; i < nx; i++) { /* :115 */
      data[(size_t)j * nx + i].re = f[(size_t)j * ldf + i];
      data[(size_t)j * nx + i].im = 0.0;
    }
  if (fft2d(nx, ny, data, -1)) return 3; /* :117 */
  data[0].re = 0.0;                      /* :118 */
  data[0].im = 0.0;
#pragma omp parallel for schedule(static)
  for (int j = 0; j < ny; j++)
    for (int i = 0; i < nx; i++) { /* :119-121 */
      const double d = aa + bb * ck[i] + cc * ck[j];
      cplx *e = &data[(size_t)j * nx + i];
      e->re = e->re / d;
      e->im = e->im / d;
    }
  if (fft2d(nx, ny, data, +1)) return 3; /* :123 ifft = backward / (nx*ny) */
  const double sc = 1.0 / ((double)nx * (double)ny);
#pragma omp parallel for schedule(static)
  for (int j = 0; j < ny; j++)
    for (int i = 0; i < nx; i++) out[(size_t)j * ldo + i] = data[(size_t)j * nx + i].re * sc;
  free(data);
  free(kx);
  free(ck);
  return 0;
}

/* fps(nx,ny,dx,dy,u,e,data,data1,f,s,eps): writes s[2:nx+1,2:ny+1] only.  Common.jl:97-125 */
int orc_fps(int nx, int ny, double dx, double dy, const double *f, double *s, double eps) {
  const int ld = nx + 2;
  return poisson_core(nx, ny, dx, dy, f, nx, s + ld + 1, ld, eps);
}

/* ps_fft(nx,ny,dx,dy,f,eps): f is (nx+1)x(ny+1), reads [1:nx,1:ny]; returns nx x ny.  fft_p.jl:8-42 */
int orc_ps_fft(int nx, int ny, double dx, double dy, const double *f, double *u, double eps) {
  return poisson_core(nx, ny, dx, dy, f, nx + 1, u, nx, eps);
}

/* ghost fill in the order of Common.jl:138-146 / vm.jl:30-38:
 *   a[nx+2,:]=a[2,:]; a[:,ny+2]=a[:,2]; a[1,:]=a[nx+1,:]; a[:,1]=a[:,ny+1]            */
void orc_ghost_fill(int nx, int ny, double *a) {
  const int ld = nx + 2;
  for (int j = 0; j < ny + 2; j++) a[(size_t)j * ld + nx + 1] = a[(size_t)j * ld + 1];
  for (int i = 0; i < nx + 2; i++) a[(size_t)(ny + 1) * ld + i] = a[(size_t)1 * ld + i];
  for (int j = 0; j < ny + 2; j++) a[(size_t)j * ld + 0] = a[(size_t)j * ld + nx];
  for (int i = 0; i < nx + 2; i++) a[(size_t)0 * ld + i] = a[(size_t)ny * ld + i];
}

/* vm_rhs(nx,ny,dx,dy,re,w,u,e,data,data1,r,s,f).  Common.jl:132-182.
 * w ghosted (read), r ghosted (interior written), s ghosted (all written), f nx x ny (written). */
int orc_vm_rhs(int nx, int ny, double dx, double dy, double re, const double *w, double *r,
               double *s, double *f) {
  const int ld = nx + 2;
#define W(i, j) w[(size_t)(j) * ld + (i)]
#define S(i, j) s[(size_t)(j) * ld + (i)]
#pragma omp parallel for schedule(static)
  for (int j = 0; j < ny; j++)
    for (int i = 0; i < nx; i++) f[(size_t)j * nx + i] = -W(i + 1, j + 1); /* :134 */
  int rc = orc_fps(nx, ny, dx, dy, f, s, 1.e-6);                            /* :136 */
  if (rc) return rc;
  orc_ghost_fill(nx, ny, s); /* :138-146 */
  const double aa = 1.0 / (re * (dx * dx)); /* :149 */
  const double bb = 1.0 / (re * (dy * dy)); /* :150 */
  const double gg = 1.0 / (4.0 * dx * dy);  /* :151 */
  const double hh = 1.0 / 3.0;              /* :152 */
#pragma omp parallel for schedule(static)
  for (int j = 1; j <= ny; j++)
    for (int i = 1; i <= nx; i++) { /* 0-based ghosted indices: interior 1..n */
      const double j1 = (W(i + 1, j) - W(i - 1, j)) * (S(i, j + 1) - S(i, j - 1)) -
                        (W(i, j + 1) - W(i, j - 1)) * (S(i + 1, j) - S(i - 1, j));
      const double j2 = W(i + 1, j) * (S(i + 1, j + 1) - S(i + 1, j - 1)) -
                        W(i - 1, j) * (S(i - 1, j + 1) - S(i - 1, j - 1)) -
                        W(i, j + 1) * (S(i + 1, j + 1) - S(i - 1, j + 1)) +
                        W(i, j - 1) * (S(i + 1, j - 1) - S(i - 1, j - 1));
      const double j3 = W(i + 1, j + 1) * (S(i, j + 1) - S(i + 1, j)) -
                        W(i - 1, j - 1) * (S(i - 1, j) - S(i, j - 1)) -
                        W(i - 1, j + 1) * (S(i, j + 1) - S(i - 1, j)) +
                        W(i + 1, j - 1) * (S(i + 1, j) - S(i, j - 1));
      const double jac = gg * (j1 + j2 + j3) * hh; /* :174 */
      r[(size_t)j * ld + i] = -jac + (aa * (W(i + 1, j) - 2.0 * W(i, j) + W(i - 1, j)) +
                                      bb * (W(i, j + 1) - 2.0 * W(i, j) + W(i, j - 1)));
    }
#undef W
#undef S
  return 0;
}

/* numerical(...) of vm.jl:12-90 / tgv.jl:13-79 without file output.
 * wn ghosted, mutated in place (all ghosts valid on return).
 * out: (nx+1)x(ny+1) = wn[2:nx+2,2:ny+2], may be NULL.  s_out: ghosted psi of the LAST rhs call, may be NULL.
 * snap: optional callback after every step k with mod(k,freq)==0 (vm.jl:78). */
typedef void (*orc_snap_fn)(int k, const double *wn_ghosted, void *user);

int orc_numerical(int nx, int ny, int64_t nt, double dx, double dy, double dt, double re, double *wn,
                  double *out, double *s_out, int64_t freq, orc_snap_fn snap, void *user) {
  const int ld = nx + 2;
  const size_t ng = (size_t)ld * (ny + 2);
  double *wt = (double *)calloc(ng, sizeof(double));
  double *r = (double *)calloc(ng, sizeof(double));
  double *s = (double *)calloc(ng, sizeof(double));
  double *f = (double *)calloc((size_t)nx * ny, sizeof(double));
  if (!wt || !r || !s || !f) return 2;
  int rc = 0;
  for (int64_t k = 1; k <= nt && !rc; k++) {
    rc = orc_vm_rhs(nx, ny, dx, dy, re, wn, r, s, f); /* vm.jl:26 */
    if (rc) break;
#pragma omp parallel for schedule(static)
    for (int j = 1; j <= ny; j++)
      for (int i = 1; i <= nx; i++) {
        const size_t q = (size_t)j * ld + i;
        wt[q] = wn[q] + dt * r[q]; /* :28 */
      }
    orc_ghost_fill(nx, ny, wt);                        /* :30-38 */
    rc = orc_vm_rhs(nx, ny, dx, dy, re, wt, r, s, f); /* :41 */
    if (rc) break;
#pragma omp parallel for schedule(static)
    for (int j = 1; j <= ny; j++)
      for (int i = 1; i <= nx; i++) {
        const size_t q = (size_t)j * ld + i;
        wt[q] = .75 * wn[q] + .25 * wt[q] + (.25 * dt) * r[q]; /* :43-47 */
      }
    orc_ghost_fill(nx, ny, wt);
    rc = orc_vm_rhs(nx, ny, dx, dy, re, wt, r, s, f); /* :60 */
    if (rc) break;
#pragma omp parallel for schedule(static)
    for (int j = 1; j <= ny; j++)
      for (int i = 1; i <= nx; i++) {
        const size_t q = (size_t)j * ld + i;
        wn[q] = wn[q] / 3. + (2. / 3.) * wt[q] + ((2. / 3.) * dt) * r[q]; /* :62-66 */
      }
    orc_ghost_fill(nx, ny, wn);
    if (snap && freq > 0 && (k % freq) == 0) snap((int)k, wn, user);
  }
  if (!rc && out)
    for (int j = 0; j <= ny; j++)
      for (int i = 0; i <= nx; i++) out[(size_t)j * (nx + 1) + i] = wn[(size_t)(j + 1) * ld + i + 1];
  if (!rc && s_out) memcpy(s_out, s, ng * sizeof(double));
  free(wt);
  free(r);
  free(s);
  free(f);
  return rc;
}

/* vm_ic + the ghost fill of main().  Common.jl:208-219, vm.jl:107-128. x,y have nx+1 / ny+1 entries. */
void orc_vm_ic(int nx, int ny, const double *x, const double *y, double *w) {
  const int ld = nx + 2;
  const double sigma = M_PI;
  const double xc1 = M_PI - M_PI / 4., yc1 = M_PI;
  const double xc2 = M_PI + M_PI / 4., yc2 = M_PI;
  for (int j = 2; j <= ny + 2; j++)
    for (int i = 2; i <= nx + 2; i++) {
      const double xx = x[i - 2], yy = y[j - 2];
      w[(size_t)(j - 1) * ld + (i - 1)] =
          exp(-sigma * ((xx - xc1) * (xx - xc1) + (yy - yc1) * (yy - yc1))) +
          exp(-sigma * ((xx - xc2) * (xx - xc2) + (yy - yc2) * (yy - yc2)));
    }
  /* vm.jl:121-128 order: [1,:],[:,1],[nx+2,:],[:,ny+2] */
  for (int j = 0; j < ny + 2; j++) w[(size_t)j * ld + 0] = w[(size_t)j * ld + nx];
  for (int i = 0; i < nx + 2; i++) w[(size_t)0 * ld + i] = w[(size_t)ny * ld + i];
  for (int j = 0; j < ny + 2; j++) w[(size_t)j * ld + nx + 1] = w[(size_t)j * ld + 1];
  for (int i = 0; i < nx + 2; i++) w[(size_t)(ny + 1) * ld + i] = w[(size_t)1 * ld + i];
}

/* exact_tgv, tgv.jl:82-90: ue is (nx+1)x(ny+1) column-major */
void orc_exact_tgv(int nx, int ny, const double *x, const double *y, double time, double re, double *ue) {
  const int nq = 4;
  for (int i = 0; i <= nx; i++)
    for (int j = 0; j <= ny; j++)
      ue[(size_t)j * (nx + 1) + i] =
          2 * nq * cos(nq * x[i]) * cos(nq * y[j]) * exp(-2 * nq * nq * time / re);
}

/* compute_l2norm_bnds, Common.jl:234-237: r is (nx+1)x(ny+1) */
double orc_l2norm_bnds(int nx, int ny, const double *r) {
  double rms = 0.0;
  for (size_t q = 0; q < (size_t)(nx + 1) * (ny + 1); q++) rms += r[q] * r[q];
  return sqrt(rms / ((double)(nx + 1) * (double)(ny + 1)));
}

/* plain 2-D DFT for the FFT-kernel unit tests: a is nx x ny interleaved complex, in place */
int orc_fft2(int nx, int ny, double *a, int sign) { return fft2d(nx, ny, (cplx *)a, sign); }

/* bench.py sets the thread count explicitly (torchrun exports OMP_NUM_THREADS=1 to its workers) */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
