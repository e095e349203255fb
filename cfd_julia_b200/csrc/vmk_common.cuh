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

constexpr int kMaxPeers = 8;
struct PeerPtrs {
  void* p[kMaxPeers];
};

}  // namespace vmk
