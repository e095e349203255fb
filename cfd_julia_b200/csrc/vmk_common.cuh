// vmk_common.cuh -- shared helpers for the vmk kernels.
//
// Every kernel body in this directory is written as  body(const Ctx&, Args)  and is
// __host__ __device__: the same source runs as a CUDA kernel (product) and, for CI on machines
// without a GPU, as pthreads on the CPU (tests/emul, test infrastructure only -- the product
// library never executes kernel bodies on the host).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <type_traits>

#define VMK_HD __host__ __device__ __forceinline__

namespace vmk {

#ifndef __CUDA_ARCH__
extern "C" void vmk_host_barrier_wait(void* bar);          // defined in the emulator only
extern "C" void vmk_host_cluster_barrier_wait(void* bar);  // ditto: barrier over all CTAs of a thread-block cluster
#endif

struct Ctx {
  int tid;              // thread index in the CTA
  int bid;              // CTA index
  int nblk;             // number of CTAs in the grid
  unsigned char* smem;  // dynamic shared memory base (16-byte aligned)
  void* hbar;           // host emulation barrier (unused on device)
  double* hscratch;     // host emulation: two doubles per thread for warp-shuffle emulation (unused on device)
  bool tables_resident = false;  // the FFT twiddle tables are already in shared memory (fused small-grid step, ks_body)
  int crank;            // rank of the CTA in its thread-block cluster (0 for plain launches)
  int csize;            // CTAs per cluster (1 for plain launches)
  unsigned char* const* hcsmem;  // host emulation: shared-memory bases of the cluster's CTAs (unused on device)
  // barrier over every thread of the cluster; release/acquire: shared-memory writes made before it -- to the own
  // CTA's memory or to a peer CTA's through remote() -- are visible to every thread of the cluster after it
  VMK_HD void cluster_sync() const {
#ifdef __CUDA_ARCH__
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
#else
    vmk_host_cluster_barrier_wait(hbar);
#endif
  }
  // same barrier without the release fence on the arriving side: for "this buffer may now be overwritten" hand-offs,
  // where the arriving thread has nothing to publish (its reads are complete: their values have been consumed).  The
  // releasing form makes every thread wait for all its outstanding global stores first (ERRBAR in SASS).
  VMK_HD void cluster_sync_relaxed() const {
#ifdef __CUDA_ARCH__
    asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
#else
    vmk_host_cluster_barrier_wait(hbar);
#endif
  }
  // address of the same shared-memory location in CTA `rank` of the cluster (distributed shared memory); the result
  // is a generic pointer, ordinary loads and stores through it travel over the SM-to-SM network
  template <class T>
  VMK_HD T* remote(T* p, int rank) const {
#ifdef __CUDA_ARCH__
    unsigned long long out;
    asm volatile("mapa.u64 %0, %1, %2;" : "=l"(out) : "l"((unsigned long long)p), "r"(rank));
    return reinterpret_cast<T*>(out);
#else
    return reinterpret_cast<T*>(hcsmem[rank] + (reinterpret_cast<unsigned char*>(p) - smem));
#endif
  }
  VMK_HD void sync() const {
#ifdef __CUDA_ARCH__
    __syncthreads();
#else
    vmk_host_barrier_wait(hbar);
#endif
  }
  // exchange a complex value with lane ^ mask (all threads of the CTA call it together)
  VMK_HD void shfl_xor2(double& x, double& y, int mask) const {
#ifdef __CUDA_ARCH__
    x = __shfl_xor_sync(0xffffffffu, x, mask);
    y = __shfl_xor_sync(0xffffffffu, y, mask);
#else
    hscratch[2 * tid] = x;
    hscratch[2 * tid + 1] = y;
    vmk_host_barrier_wait(hbar);
    x = hscratch[2 * (tid ^ mask)];
    y = hscratch[2 * (tid ^ mask) + 1];
    vmk_host_barrier_wait(hbar);
#endif
  }
};

// ---- compile-time loop ---------------------------------------------------------------------
template <int I, int N, class F>
VMK_HD void static_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, N>(f);
  }
}

// ---- complex arithmetic on double2 ---------------------------------------------------------
VMK_HD double2 mk2(double x, double y) {
  double2 r;
  r.x = x;
  r.y = y;
  return r;
}
VMK_HD double2 cadd(double2 a, double2 b) { return mk2(a.x + b.x, a.y + b.y); }
VMK_HD double2 csub(double2 a, double2 b) { return mk2(a.x - b.x, a.y - b.y); }
// Explicit fused multiply-add.  The library is compiled with --fmad=false so that the stencil and the divisor
// keep the reference's unfused source-order arithmetic; the FFT butterflies ask for FMAs by name (2 DMUL + 2 DFMA
// per complex multiply instead of 4 DMUL + 2 DADD; the FP64 pipe is a co-bottleneck of the FFT kernels).
VMK_HD double fma_(double a, double b, double c) {
#ifdef __CUDA_ARCH__
  return __fma_rn(a, b, c);
#else
  return __builtin_fma(a, b, c);
#endif
}
VMK_HD double2 cmul(double2 a, double2 b) {
  return mk2(fma_(a.x, b.x, -(a.y * b.y)), fma_(a.x, b.y, a.y * b.x));
}
VMK_HD double2 cmulc(double2 a, double2 b) {  // a * conj(b)
  return mk2(fma_(a.x, b.x, a.y * b.y), fma_(a.y, b.x, -(a.x * b.y)));
}
VMK_HD double2 csqr(double2 a) {  // a^2: 4 FP64 instructions
  const double t = a.x * a.y;
  return mk2(fma_(a.x, a.x, -(a.y * a.y)), t + t);
}
VMK_HD double2 cconj(double2 a) { return mk2(a.x, -a.y); }
VMK_HD double2 cfma(double2 y, double r, double2 x) { return mk2(fma_(y.x, r, x.x), fma_(y.y, r, x.y)); }  // y r + x
VMK_HD double2 cscale(double2 a, double s) { return mk2(a.x * s, a.y * s); }

// ---- global memory access with cache hints -------------------------------------------------
// Streaming data is touched once per kernel: keep it out of L1 (it is staged in registers /
// shared memory by the kernels themselves).
VMK_HD double2 ld_stream2(const double2* p) {
#ifdef __CUDA_ARCH__
  double2 r;
  asm volatile("ld.global.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
#else
  return *p;
#endif
}
VMK_HD double ld_stream1(const double* p) {
#ifdef __CUDA_ARCH__
  double r;
  asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
  return r;
#else
  return *p;
#endif
}
VMK_HD void st_stream2(double2* p, double2 v) {
#ifdef __CUDA_ARCH__
  asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
#else
  *p = v;
#endif
}
// the same store without the "memory" clobber: for buffers the kernel never reads (later loads may be hoisted above it)
VMK_HD void st_stream2_nc(double2* p, double2 v) {
#ifdef __CUDA_ARCH__
  asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v.x), "d"(v.y));
#else
  *p = v;
#endif
}
VMK_HD void st_stream1(double* p, double v) {
#ifdef __CUDA_ARCH__
  asm volatile("st.global.L1::no_allocate.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
#else
  *p = v;
#endif
}
// one 32-byte store (two adjacent complex values); p must be 32-byte aligned.
// sm_100 has 256-bit global stores (st.global.v4.f64).
VMK_HD void st_stream4(double2* p, double2 a, double2 b) {
#ifdef __CUDA_ARCH__
  asm volatile("st.global.L1::no_allocate.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a.x), "d"(a.y), "d"(b.x),
               "d"(b.y)
               : "memory");
#else
  p[0] = a;
  p[1] = b;
#endif
}
#ifndef VMK_GATHER_L2
#define VMK_GATHER_L2 ""
#endif
VMK_HD void ld_stream4(const double2* p, double2& a, double2& b) {
#ifdef __CUDA_ARCH__
  asm volatile("ld.global.L1::no_allocate" VMK_GATHER_L2 ".v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(a.x), "=d"(a.y), "=d"(b.x), "=d"(b.y)
               : "l"(p));
#else
  a = p[0];
  b = p[1];
#endif
}
// Asynchronous 16-byte copy global -> shared (LDGSTS): no destination registers and no scoreboard wait, so a thread
// can put the next row's loads in flight while its registers still hold the current row.  The data is visible to
// the issuing thread after cp_async_wait_all(), to the rest of the CTA after a barrier on top of that.
VMK_HD void cp_async16(void* smem_dst, const void* gsrc) {
#ifdef __CUDA_ARCH__
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
#else
  memcpy(smem_dst, gsrc, 16);
#endif
}
VMK_HD void cp_async_commit() {
#ifdef __CUDA_ARCH__
  asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
VMK_HD void cp_async_wait_all() {
#ifdef __CUDA_ARCH__
  asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
}
VMK_HD double ld_ro(const double* p) {
#ifdef __CUDA_ARCH__
  return __ldg(p);
#else
  return *p;
#endif
}
VMK_HD int ld_roi(const int* p) {
#ifdef __CUDA_ARCH__
  return __ldg(p);
#else
  return *p;
#endif
}
VMK_HD double2 ld_ro2(const double2* p) {
#ifdef __CUDA_ARCH__
  return __ldg(p);
#else
  return *p;
#endif
}
VMK_HD void prefetch_l2(const void* p) {
#ifdef __CUDA_ARCH__
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
  (void)p;
#endif
}
// one instruction asks the L2 to fetch a whole contiguous range (16-byte aligned, multiple of 16 bytes):
// issued by one thread for the NEXT row while the CTA transforms the current one, so that HBM streams
// during the compute phase and the register loads of the next iteration are L2 hits
VMK_HD void prefetch_l2_bulk(const void* p, unsigned bytes) {
#ifdef __CUDA_ARCH__
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
#else
  (void)p;
  (void)bytes;
#endif
}
VMK_HD double rcp_rn(double d) {
#ifdef __CUDA_ARCH__
  return __drcp_rn(d);
#else
  return 1.0 / d;
#endif
}

// 1/d to <= 1 ulp without the branches of the IEEE-rounded reciprocal: hardware seed (20 mantissa bits) and two
// Newton steps.  d is a Poisson divisor: finite, normal, never zero (the eps quirk keeps mode (0,0) off zero).
VMK_HD double rcp_fast(double d) {
#ifdef __CUDA_ARCH__
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
  double e = __fma_rn(-d, x, 1.0);
  x = __fma_rn(x, e, x);
  e = __fma_rn(-d, x, 1.0);
  x = __fma_rn(x, e, x);
  e = __fma_rn(-d, x, 1.0);
  return __fma_rn(x, e, x);
#else
  return 1.0 / d;
#endif
}

// ---- tensor memory as per-thread state ---------------------------------------------------------------------------------
// The FP64 kernels have no use for the tensor cores, but the 256 KB of tensor memory next to them (128 lanes x 512
// 32-bit columns per SM) is storage a thread can reach without the LSU: lane 32 (w % 4) + l belongs to lane l of warp w,
// so a CTA of 8 warps has 128 doubles per thread of state that survives across loop iterations when the register
// file is full (the fused forms of K1 / K3 keep the running values of the recurrences along j there, vmk_tri.cuh).
// tcgen05.ld / .st are warp-wide (.sync.aligned): every call site is reached by whole warps.
// The state is private to the thread and the instructions are volatile (kept in program order among themselves), so
// they carry no "memory" clobber: the compiler stays free to move ordinary loads across them.
// Emulator: a plain per-thread array.
// columns a CTA of `ct` threads allocates for `nd` doubles per thread (power of two >= 32)
VMK_HD constexpr int tm_cols(int nd, int ct) {
  const int need = 2 * nd * ((ct + 127) / 128);
  int c = 32;
  while (c < need) c *= 2;
  return c;
}
template <int ND>  // doubles per thread
struct TmState {
#ifdef __CUDA_ARCH__
  unsigned base;  // tensor-memory address of this thread's first column: (lane base << 16) | column
#else
  double buf[ND];
#endif
  // COLS = columns allocated by the CTA (power of two >= 32, >= 2 ND x warps/4); all threads of the CTA call it
  template <int COLS>
  VMK_HD void open(const Ctx& c) {
#ifdef __CUDA_ARCH__
    __shared__ unsigned tm_base_slot;
    if (c.tid < 32) {
      const unsigned dst = (unsigned)__cvta_generic_to_shared(&tm_base_slot);
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst), "r"(COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned w = (unsigned)c.tid >> 5;
    base = tm_base_slot + (((w & 3u) * 32u) << 16) + (w >> 2) * (unsigned)(2 * ND);
#else
    (void)c;
    for (int i = 0; i < ND; i++) buf[i] = 0.0;
#endif
  }
  template <int COLS>
  VMK_HD void close(const Ctx& c) {
#ifdef __CUDA_ARCH__
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (c.tid < 32) {
      const unsigned w0 = base & 0xffffu;  // warp 0: lane base 0, column offset 0
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(w0), "r"(COLS) : "memory");
    }
#else
    (void)c;
#endif
  }
  // two / four / six consecutive doubles at offset OFF (in doubles); load and wait are one asm statement, so the
  // registers are valid when it returns
  template <int OFF>
  VMK_HD void ld2(double& a, double& b) const {
#ifdef __CUDA_ARCH__
    unsigned r0, r1, r2, r3;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
        : "r"(base + 2u * OFF));
    a = __hiloint2double((int)r1, (int)r0);
    b = __hiloint2double((int)r3, (int)r2);
#else
    a = buf[OFF];
    b = buf[OFF + 1];
#endif
  }
  template <int OFF>
  VMK_HD void ld4(double& a, double& b, double& d, double& e) const {
#ifdef __CUDA_ARCH__
    unsigned r0, r1, r2, r3, r4, r5, r6, r7;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7)
        : "r"(base + 2u * OFF));
    a = __hiloint2double((int)r1, (int)r0);
    b = __hiloint2double((int)r3, (int)r2);
    d = __hiloint2double((int)r5, (int)r4);
    e = __hiloint2double((int)r7, (int)r6);
#else
    a = buf[OFF];
    b = buf[OFF + 1];
    d = buf[OFF + 2];
    e = buf[OFF + 3];
#endif
  }
  template <int OFF>
  VMK_HD void ld6(double& a, double& b, double& d, double& e, double& f, double& g) const {
#ifdef __CUDA_ARCH__
    unsigned r0, r1, r2, r3, r4, r5, r6, r7, r8, r9, r10, r11;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%12];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%8, %9, %10, %11}, [%13];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7), "=r"(r8), "=r"(r9),
          "=r"(r10), "=r"(r11)
        : "r"(base + 2u * OFF), "r"(base + 2u * OFF + 8u));
    a = __hiloint2double((int)r1, (int)r0);
    b = __hiloint2double((int)r3, (int)r2);
    d = __hiloint2double((int)r5, (int)r4);
    e = __hiloint2double((int)r7, (int)r6);
    f = __hiloint2double((int)r9, (int)r8);
    g = __hiloint2double((int)r11, (int)r10);
#else
    a = buf[OFF];
    b = buf[OFF + 1];
    d = buf[OFF + 2];
    e = buf[OFF + 3];
    f = buf[OFF + 4];
    g = buf[OFF + 5];
#endif
  }
  template <int OFF>
  VMK_HD void st2(double a, double b) {
#ifdef __CUDA_ARCH__
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + 2u * OFF),
                 "r"(__double2loint(a)), "r"(__double2hiint(a)), "r"(__double2loint(b)), "r"(__double2hiint(b)));
#else
    buf[OFF] = a;
    buf[OFF + 1] = b;
#endif
  }
  template <int OFF>
  VMK_HD void st4(double a, double b, double d, double e) {
#ifdef __CUDA_ARCH__
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(base + 2u * OFF),
                 "r"(__double2loint(a)), "r"(__double2hiint(a)), "r"(__double2loint(b)), "r"(__double2hiint(b)),
                 "r"(__double2loint(d)), "r"(__double2hiint(d)), "r"(__double2loint(e)), "r"(__double2hiint(e)));
#else
    buf[OFF] = a;
    buf[OFF + 1] = b;
    buf[OFF + 2] = d;
    buf[OFF + 3] = e;
#endif
  }
  // stores issued so far are complete (before the same columns are loaded again)
  VMK_HD void fence_st() const {
#ifdef __CUDA_ARCH__
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
#endif
  }
};

// FUSED kernels: unit q = CTA * FPC + transform group owns the contiguous pairs [fz_first(q), fz_first(q + 1)) and
// takes them in order (K1: ascending, K3: descending), so that the recurrences along j run inside the unit
VMK_HD int fz_first(int q, int units, int npairs) { return (int)(((long long)q * npairs) / units); }

constexpr int kMaxPeers = 8;
struct PeerPtrs {
  void* p[kMaxPeers];
};

}  // namespace vmk
