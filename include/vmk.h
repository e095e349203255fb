/*
 * vmk.h -- C ABI of libvmk.so, the B200 (sm_100a) implementation of the CFD_Julia vortex-merger step.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  The reference is a set of
 * Julia scripts with no FFI of its own; the entry points below are what a `ccall` wrapper keeping the
 * reference's Julia signatures binds (julia/CommonB200.jl).  File:line citations are
 * relative to the reference checkout (t-bltg/CFD_Julia).
 *
 * Conventions
 *   - All host arrays are column-major Float64 exactly as Julia lays them out.  "ghosted" means
 *     (nx+2) x (ny+2) with interior [2:nx+1, 2:ny+1] (1-based), element (i,j) at [(i-1) + (nx+2)*(j-1)].
 *   - nx == ny (the reference aliases ky = kx, Common.jl:113) and a power of two in [32, 32768] (16384 and 32768: one row per
 *     thread-block cluster of 2 / 4 SMs, csrc/vmk_cluster.cuh).
 *   - Every function returns 0 on success; otherwise a VMK_E* code, and vmk_last_error() (thread-local)
 *     describes it.  The library never falls back to a CPU path: without a usable CUDA device every
 *     compute entry point fails with VMK_ECUDA.
 *   - Entry points taking host pointers are synchronous (they return after the device->host copy), because
 *     Julia only roots ccall arguments for the duration of the call.
 *   - Threads: the reference is single-threaded.  Every entry point that takes a plan holds that plan's (recursive)
 *     mutex for the duration of the call, so calls on ONE plan from several host threads are serialised, and different
 *     plans run concurrently.  vmk_step is asynchronous: the lock covers the enqueue, the stream orders the work.
 *     vmk_plan_destroy must not race with other calls on the same plan.  The library never calls back into the host
 *     language except through the snapshot callback, on the calling thread.
 */
#ifndef VMK_H_
#define VMK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VMK_OK 0
#define VMK_ESIZE 1  /* unsupported grid size (not square, not a power of two, out of range) */
#define VMK_ECUDA 2  /* CUDA runtime error (including "no device") */
#define VMK_EARG 3   /* bad argument (NULL pointer, bad rank, ...) */
#define VMK_ESTATE 4 /* call sequence error (e.g. step before upload) */
#define VMK_EIO 5    /* snapshot file could not be opened, written or parsed */

typedef struct vmk_plan vmk_plan;

int vmk_version(void);
const char* vmk_last_error(void);

/* Owns device buffers, twiddle/divisor tables and the stream for one grid size on the current device.
 * Replaces the per-call allocations of numerical() (vm.jl:13-19) and the per-call FFTW plans and kx table
 * of fps (Common.jl:98-113,117,123). */
int vmk_plan_create(int64_t nx, int64_t ny, vmk_plan** plan);
int vmk_plan_destroy(vmk_plan* plan);

/* ---- slab decomposition over several GPUs (no counterpart in the reference, which is single-threaded) --
 * Rank g of nranks (1, 2, 4 or 8) owns grid columns j in [g*ny/nranks, (g+1)*ny/nranks) (Julia dim 2) on the
 * CUDA device that is current when the plan is created.  The kernels exchange data by loading from /
 * storing to the other ranks' buffers directly over NVLink (halo rows of w and psi; the j-direction FFT
 * reads the whole spectrum row from all ranks), so every rank needs the others' buffer addresses:
 *   one process per GPU : vmk_peer_export -> all-gather the blobs -> vmk_peer_import (CUDA IPC)
 *   one process, N GPUs : vmk_peer_attach_local (enables peer access between the plans' devices itself; every
 *                         entry point makes its plan's device current, so one host thread can drive all ranks:
 *                         call vmk_step for every rank -- it is asynchronous -- before any vmk_download)
 * and a cross-rank barrier that the plan enqueues on its stream between dependent kernels
 * (vmk_barrier_hook; e.g. a 1-element NCCL all-reduce on that stream).
 * Host-array entry points of a slab plan read/write only the rank's own columns of the caller's arrays. */
int vmk_plan_create_slab(int64_t nx, int64_t ny, int rank, int nranks, vmk_plan** plan);
/* same, on CUDA device `device` (made current): for a single host process that owns all ranks */
int vmk_plan_create_on(int device, int64_t nx, int64_t ny, int rank, int nranks, vmk_plan** plan);
size_t vmk_peer_blob_bytes(void);
int vmk_peer_export(vmk_plan* plan, void* blob);
int vmk_peer_import(vmk_plan* plan, const void* blobs /* nranks blobs in rank order */);
int vmk_peer_attach_local(vmk_plan* plan, vmk_plan* const* plans /* nranks plans in rank order */);
int vmk_barrier_hook(vmk_plan* plan, void (*enqueue_barrier)(void* user), void* user);

/* ---- reference-signature entry points on HOST arrays ------------------------------------------------ */

/* fps(nx,ny,dx,dy,u,e,data,data1,f,s,eps)  Common.jl:97-125.
 * f: nx x ny source.  s: ghosted; only s[2:nx+1,2:ny+1] is written (the reference leaves ghosts alone).
 * The dead scratch arguments u,e,data,data1 of the reference are not part of the ABI. */
int vmk_fps(vmk_plan* plan, double dx, double dy, const double* f, double* s, double eps);

/* ps_fft(nx,ny,dx,dy,f,eps)  12_Poisson_Solver_FFT/fft_p.jl:8-42.
 * f: (nx+1) x (ny+1), only [1:nx,1:ny] is read.  u: nx x ny result. */
int vmk_ps_fft(vmk_plan* plan, double dx, double dy, const double* f, double* u, double eps);

/* vm_rhs(nx,ny,dx,dy,re,w,u,e,data,data1,r,s,f)  Common.jl:132-182.
 * w: ghosted vorticity with periodic ghosts (read).  r: ghosted, interior written.  s: ghosted, all cells
 * written (ghost fill of Common.jl:138-146).  f: nx x ny, receives -w interior (Common.jl:134); may be NULL. */
int vmk_rhs(vmk_plan* plan, double dx, double dy, double re, const double* w, double* r, double* s, double* f);

/* numerical(nx,ny,nt,dx,dy,dt,re,x,y,wn,ns)  19_NS2D_Vortex_Merger/vm.jl:12-90 and
 * numerical(nx,ny,nt,dx,dy,dt,re,wn)        19_NS2D_Vortex_Merger/tgv.jl:13-79.
 * wn: ghosted, mutated in place, all ghosts valid on return.  out: (nx+1) x (ny+1) = wn[2:nx+2,2:ny+2], or NULL.
 * If snap != NULL and freq > 0, after every step k with k % freq == 0 (vm.jl:78) wn is brought to the host
 * (into the caller's wn) and snap(k, wn, user) is called; file output stays on the caller's side. */
typedef void (*vmk_snapshot_fn)(int64_t k, const double* wn_ghosted, void* user);
int vmk_numerical(vmk_plan* plan, int64_t nt, double dx, double dy, double dt, double re, double* wn, double* out,
                  int64_t freq, vmk_snapshot_fn snap, void* user);

/* numerical(nx,ny,nt,dx,dy,dt,re,x,y,wn,ns)  20_NS2D_Hybrid_Solver/hybrid.jl:14-90 -- the hybrid solver: the vorticity
 * lives in Fourier space, the Arakawa Jacobian is evaluated in real space (jacobian(), hybrid.jl:96-152) and advanced
 * with RK3, the diffusion with Crank-Nicolson per mode (k2 from wavespace, Common.jl:184-204, kx[1] = eps).
 * wn: ghosted initial vorticity (read only, as in the reference).  ut: (nx+1) x (ny+1), receives real(ifft(wf)) with
 * the periodic duplicates (what the reference returns).  snap(k, ut, user) after every step k with k % freq == 0
 * (here the array passed is ut, (nx+1) x (ny+1)).  Single-GPU plans, nx == ny <= 8192, dx == dy. */
int vmk_hybrid_numerical(vmk_plan* plan, int64_t nt, double dx, double dy, double dt, double re, const double* wn,
                         double* ut, int64_t freq, vmk_snapshot_fn snap, void* user);

/* numerical(nx,ny,nt,dx,dy,dt,re,x,y,wn,ns)  22_NS2D_PseudoSpectral_23_Rule/pseudospectral_23_rule.jl:13-89 -- the pseudo-
 * spectral solver with the 2/3 truncation rule: the vorticity lives in Fourier space, jacobian() (:95-144) multiplies
 * it by i kx [/ k2], i ky [/ k2] (kx[1] = eps, :107), zeroes the band floor(nxe/2)+1 .. nx-floor(nxe/2), takes
 * real(ifft) of the four spectra, forms j1 j2 - j3 j4 in real space and transforms it back; RK3 for the Jacobian and
 * Crank-Nicolson per mode for the diffusion exactly as in hybrid.jl.  Arguments as vmk_hybrid_numerical: wn ghosted, read
 * only; ut (nx+1) x (ny+1) receives real(ifft(wf)) with the periodic duplicates; snap(k, ut, user) after every step k
 * with k % freq == 0.  (The reference returns the field of its LAST snapshot; that is the final field whenever nt is a
 * multiple of freq, as in its own configuration -- ut here is always the final field.)
 * Single-GPU plans, nx == ny <= 8192, dx == dy. */
int vmk_ps23_numerical(vmk_plan* plan, int64_t nt, double dx, double dy, double dt, double re, const double* wn,
                       double* ut, int64_t freq, vmk_snapshot_fn snap, void* user);

/* numerical(nx,ny,nt,dx,dy,dt,re,x,y,wn,ns)  21_NS2D_PseudoSpectral_32_Rule/pseudospectral_32_rule.jl:13-89 -- the pseudo-
 * spectral solver with the 3/2 padding rule: as vmk_ps23_numerical, but jacobian() (:95-177) zero-pads the four spectra
 * to 1.5nx x 1.5ny, multiplies on that grid and keeps the nx x ny modes of the product's transform.  The 1.5nx-point
 * transforms run as radix-3 splits into nx/2-point transforms (csrc/vmk_pseudo32.cuh).  Same arguments and contract as
 * vmk_ps23_numerical.  Single-GPU plans, nx == ny in [64, 8192], dx == dy. */
int vmk_ps32_numerical(vmk_plan* plan, int64_t nt, double dx, double dy, double dt, double re, const double* wn,
                       double* ut, int64_t freq, vmk_snapshot_fn snap, void* user);

/* numerical(nx,ny,nt,dx,dy,dt,re,wn,sn,rms)  18_NS2D_Lid_Driven_Cavity/lid_driven_cavity.jl:59-117 -- lid-driven cavity:
 * RK3, the Arakawa/Laplacian rhs without periodic wrap (:123-158), Jensen wall vorticity bc2 (:38-52) and the sine-
 * transform Poisson solve fps_sine (:11-21, FFTW RODFT00).  wn, sn: (nx+1) x (ny+1) node arrays, both mutated in place
 * (sn's wall nodes are never written, as in the reference); rms[0..nt-1] receives the per-step change of sn (:111-113).
 * `plan` must have size 2nx x 2ny: the DST is evaluated as the periodic transform of the odd extension with the
 * kernels of the vortex-merger path.  nx == ny, a power of two in [16, 16384]; single-GPU plans. */
int vmk_ldc_numerical(vmk_plan* plan, int64_t nx, int64_t ny, int64_t nt, double dx, double dy, double dt, double re,
                      double* wn, double* sn, double* rms);

/* ---- the callers' snapshot files (host code) ------------------------------------------- */
/* Julia's print(::Float64) -- "$(x)" in the scripts' write calls (vm.jl:83): shortest round-trip digits, positional
 * for 1e-4 <= |v| < 1e6, otherwise d.ddde[-]X without exponent padding ("1.0e-5", not C's "1e-05").  buf: at least 32
 * bytes, no terminator is written; returns the number of characters. */
int vmk_print_float64(double v, char* buf);
/* for j in 1:ny1, i in 1:nx1: "x[i] y[j] ut[i,j]\n" (vm.jl:81-85,132-136,142-146; hybrid.jl:79-83); ut column-major
 * nx1 x ny1.  Byte-identical to what the Julia scripts write for the same values. */
int vmk_write_field(const char* path, const double* x, const double* y, const double* ut, int64_t nx1, int64_t ny1);
/* plotting.jl:14-28 (readdlm): three whitespace-separated columns per line into x, y, w (each may be NULL; at most cap
 * rows are stored), *nrows = rows in the file. */
int vmk_read_field(const char* path, double* x, double* y, double* w, int64_t cap, int64_t* nrows);

/* ---- device-resident path (what numerical() is built from) ------------------------------------------- */
int vmk_upload(vmk_plan* plan, const double* wn_ghosted);
int vmk_step(vmk_plan* plan, double dx, double dy, double dt, double re, int64_t nsteps); /* asynchronous */
int vmk_download(vmk_plan* plan, double* wn_ghosted, double* psi_ghosted);                /* either may be NULL */
int vmk_sync(vmk_plan* plan);

/* ---- measurement ------------------------------------------------------------------------------------ */
/* cudaStream_t the plan launches on (cast to void*), so callers can bracket calls with their own events */
void* vmk_stream(vmk_plan* plan);
/* device time of the last vmk_step call, from CUDA events recorded on the plan's stream */
int vmk_step_elapsed_ms(vmk_plan* plan, double* ms);
/* runs `nsteps` steps with events around every kernel; ms[0..3] receive the summed device time of
 * K1 (row forward), K2 (column forward+divide+inverse), K3 (row inverse), K4 (stencil+RK3); launches[0..3]
 * the launch counts */
int vmk_profile_steps(vmk_plan* plan, double dx, double dy, double dt, double re, int64_t nsteps, double* ms,
                      int64_t* launches);
/* after vmk_profile_steps on a plan that solves along j by recurrences (option "fps_mode", csrc/vmk_tri.cuh): the part of
 * ms[1] / launches[1] spent in ms[0] chunk totals, ms[1] scan, ms[2] in-place solve (the rest of ms[1] is K2 on the rows
 * kx < K0).  Fused form (fps_mode 2): totals and solve run inside K1 / K3 (classes 0 and 2 of vmk_profile_steps), so only
 * the scan entry is non-zero */
int vmk_profile_tri(vmk_plan* plan, double* ms, int64_t* launches);
/* With the option "profile" set to 1, every kernel of the following calls on this plan (any solver) is bracketed by
 * CUDA events; vmk_profile_read sums them per class since the last read -- ms[0] row-forward transforms (K1), ms[1]
 * spectrum-row kernels (K2, KH, KP, the batched row FFT of the 3/2 rule), ms[2] row-inverse transforms (K3), ms[3]
 * pointwise kernels (K4, products, the fold / unfold / update passes) -- and resets the sums.  (CUDA graphs are not
 * used while profiling.) */
int vmk_profile_read(vmk_plan* plan, double* ms, int64_t* launches);
/* kernels launched by this plan since creation */
int64_t vmk_launch_count(vmk_plan* plan);
/* tuning knobs (integers; defaults are the measured best, see profiles/r01_notes.md):
 *   "ps32_fuse"   0      3/2 rule: 1 = the four derivative spectra are computed in the load stage of the inverse row
 *                        transform instead of being written and read back; 2 = and folded along i there, so that the
 *                        transform writes K3's input directly (both validated on the host emulator only so far)
 *   "fps_mode"    -1=auto how fps / vm_rhs / the RK3 step solve along j (Common.jl:117-123): 0 = forward FFT, divide, inverse
 *                        FFT (K2, with the two all-to-all transposes on several GPUs); 1 = the cyclic tridiagonal solve
 *                        the divisor is the symbol of, by two-sided recurrences (csrc/vmk_tri.cuh: no transposes; needs
 *                        32 | rows per rank, N in [64, 8192]); 2 = the same recurrences INSIDE K1 / K3 (fused form, one
 *                        GPU, N in [512, 8192]: the spectrum crosses HBM twice per solve instead of five times, the
 *                        per-slot running values live in tensor memory); auto = 1 from 2048^2 up, 2 at 8192^2 on one
 *                        GPU.  Same results to ~1e-15
 *   "fz_grid"            fps_mode 2: CTAs of K1 / K3 (each owns contiguous blocks of row pairs); tuning / tests
 *   "tri_k0"      0=auto fps_mode 1: the rows kx < K0 keep the FFT form (0: N/16, at most 64)
 *   "fuse_small"  0      N <= 256, one GPU: the whole step loop as ONE cluster launch (measured slower than the graph)
 *   "profile"     0      1: bracket every kernel with events (see vmk_profile_read)
 *   "graph"       1      replay the step (kernels, copies, barriers) from a CUDA graph
 *   "k4_rows"     32     rows marched by one K4 thread column (shortened automatically on small slabs)
 *   "k4_ahead"    4      rows ahead of the march that K4 prefetches into L2 (0 = off)
 *   "v_pieces"    1      single GPU: K2 stores the solution spectrum as per-row-pair blocks in K3's read order
 *   "k1_prefetch", "k2_prefetch"  0   extra L2 bulk prefetch two rows ahead (no gain once cp.async existed)
 *   "cl_prefetch" -1=auto 16384 / 32768 (cluster kernels): bit mask of the kernels (1 K1, 2 K2, 4 K3) that bulk-prefetch
 *                        their CTA's share of the next row into L2
 *   "a2a_chunks"  0=auto launches K1 (and a staged K2) is split into so that the transpose overlaps with it
 *   "a2a_engine"  -1=auto forward transpose by copy engines (1; best at 2 GPUs) or the SM push kernel (0; best at 4, 8)
 *   "a2a_ctas"    128    CTAs of the push kernel
 *   "k2_push"     -1=auto backward transpose by direct NVLink stores from K2 (1) or staged + copy engines (0) */
int vmk_set_option(vmk_plan* plan, const char* key, int64_t value);
/* bytes of device memory held by the plan */
int64_t vmk_device_bytes(vmk_plan* plan);

#ifdef __cplusplus
}
#endif
#endif /* VMK_H_ */
