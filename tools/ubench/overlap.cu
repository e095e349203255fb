// overlap.cu -- microbenchmark behind the round-2 redesign of the FFT kernels (profiles/r02_notes.md).
//
// Question: K1/K2/K3 spend  t ~ (FP64 issue time) + (shared-memory wavefront time)  per row because all eight warps of
// the CTA are in the same phase after every barrier.  Can the two phases overlap on a B200 SM when half of the warps
// exchange while the other half computes?  (Round 1 saw no gain from two phase-shifted CTAs per SM at 4096^2.)
//
// One CTA of 256 threads per SM, 32 complex values per thread (the N = 8192 configuration), REPS iterations of
//   compute : two radix-16 butterfly networks + 30 complex twiddle multiplies on the thread's registers
//             (the real pass 0 / pass 1 arithmetic: 512 FP64 instructions per thread)
//   exchange: 32 x STS.128 -> barrier -> 32 x LDS.128 through the padded exchange buffer (conflict-free)
// Variants:
//   0  compute only            1  exchange only           2  compute + exchange, CTA-wide barrier (today's kernels)
//   3  two independent half-CTAs (128 threads, named barriers), same order in both halves
//   4  two independent half-CTAs, opposite order (half 0: compute, exchange; half 1: exchange, compute)
//   5  as 2 but the twiddles are read from shared memory (8 LDS.128 per butterfly, as the kernels do)
//   6  as 4 with shared-memory twiddles
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --fmad=false -I cfd_julia_b200/csrc -o overlap overlap.cu
#include <cstdio>
#include <cstdlib>

#include "vmk_fft.cuh"

using namespace vmk;

constexpr int E = 32, T = 256, N = 8192;

__device__ __forceinline__ int paddr(int pos) { return pos + (pos >> 5); }

template <bool SMEM_TW>
__device__ __forceinline__ void compute(double2 (&v)[E], const double2* tw, int low) {
  static_for<0, 2>([&](auto u_) {
    constexpr int u = decltype(u_)::value;
    double2 a[16];
    static_for<0, 16>([&](auto q_) { constexpr int q = decltype(q_)::value; a[q] = v[u * 16 + q]; });
    Net<16, -1, 0>::run(a);
    static_for<0, 16>([&](auto p_) { constexpr int p = decltype(p_)::value; v[u * 16 + p] = a[brev(p, 4)]; });
    static_for<0, 8>([&](auto h_) {
      constexpr int q = 2 * decltype(h_)::value + 1;
      double2 w;
      if constexpr (SMEM_TW)
        w = tw[(low * q) & 4095];
      else
        w = mk2(1.0 - 1e-9 * q, 1e-5 * q);
      v[u * 16 + q] = cmul(v[u * 16 + q], w);
      static_for<1, 4>([&](auto s_) {
        constexpr int m = q << decltype(s_)::value;
        if constexpr (m < 16) {
          w = csqr(w);
          v[u * 16 + m] = cmul(v[u * 16 + m], w);
        }
      });
    });
  });
}

// store in the pass-0 layout, load in the pass-1 layout (T threads of the group, group-local buffer)
template <int TG>
__device__ __forceinline__ void exch_store(const double2 (&v)[E], double2* sm, int t) {
  static_for<0, E>([&](auto e_) {
    constexpr int e = decltype(e_)::value;
    sm[paddr(t + TG * e)] = v[e];
  });
}
template <int TG>
__device__ __forceinline__ void exch_load(double2 (&v)[E], const double2* sm, int t) {
  // 16 consecutive-by-32 positions per butterfly: ((id >> 5) << 9) | (id & 31) | (q << 5), wrapped into the group's buffer
  static_for<0, 2>([&](auto u_) {
    constexpr int u = decltype(u_)::value;
    const int id = t + TG * u;
    const int bp = (((id >> 5) << 9) | (id & 31)) & (TG * E - 1);
    static_for<0, 16>([&](auto q_) {
      constexpr int q = decltype(q_)::value;
      v[u * 16 + q] = sm[paddr((bp | (q << 5)) & (TG * E - 1))];
    });
  });
}

// explicit shared-memory accesses: volatile asm keeps their order relative to each other, so the source decides where
// the stores of butterfly 0 sit relative to the loads / stores of butterfly 1
__device__ __forceinline__ double2 lds128(const double2* p) {
  double2 r;
  const unsigned a = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "r"(a));
  return r;
}
__device__ __forceinline__ void sts128(double2* p, double2 v) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(a), "d"(v.x), "d"(v.y) : "memory");
}
// one radix-16 butterfly + twiddles on a[16]
template <bool SMEM_TW>
__device__ __forceinline__ void bfly16(double2 (&a)[16], const double2* tw, int low) {
  Net<16, -1, 0>::run(a);
  double2 b[16];
  static_for<0, 16>([&](auto p_) { constexpr int p = decltype(p_)::value; b[p] = a[brev(p, 4)]; });
  static_for<0, 8>([&](auto h_) {
    constexpr int q = 2 * decltype(h_)::value + 1;
    double2 w;
    if constexpr (SMEM_TW)
      w = tw[(low * q) & 4095];
    else
      w = mk2(1.0 - 1e-9 * q, 1e-5 * q);
    b[q] = cmul(b[q], w);
    static_for<1, 4>([&](auto s_) {
      constexpr int m = q << decltype(s_)::value;
      if constexpr (m < 16) {
        w = csqr(w);
        b[m] = cmul(b[m], w);
      }
    });
  });
  static_for<0, 16>([&](auto p_) { constexpr int p = decltype(p_)::value; a[p] = b[p]; });
}

__device__ __forceinline__ void bar_named(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

template <int VAR>
__global__ void __launch_bounds__(256, 1) kern(double* out, long long* cyc, int reps) {
  extern __shared__ __align__(16) unsigned char smem[];
  double2* xb = reinterpret_cast<double2*>(smem);
  double2* tw = xb + (N + N / 32) + 64;
  const int tid = threadIdx.x;
  for (int i = tid; i < 4096; i += 256) tw[i] = mk2(cos(1e-3 * i), -sin(1e-3 * i));
  double2 v[E];
  static_for<0, E>([&](auto e_) {
    constexpr int e = decltype(e_)::value;
    v[e] = mk2(1e-3 * (tid + e), 1e-3 * (tid - e));
  });
  __syncthreads();
  const long long t0 = clock64();
  if constexpr (VAR == 0) {
    for (int r = 0; r < reps; r++) compute<false>(v, tw, tid);
  } else if constexpr (VAR == 1) {
    for (int r = 0; r < reps; r++) {
      exch_store<256>(v, xb, tid);
      __syncthreads();
      exch_load<256>(v, xb, tid);
    }
  } else if constexpr (VAR == 2 || VAR == 5) {
    for (int r = 0; r < reps; r++) {
      compute<VAR == 5>(v, tw, tid);
      exch_store<256>(v, xb, tid);
      __syncthreads();
      exch_load<256>(v, xb, tid);
    }
  } else if constexpr (VAR == 10 || VAR == 11) {
    // software-pipelined pass: [LDS u0][LDS u1][F u0][STS u0][F u1][STS u1] barrier -- half of the loads and half of the
    // stores overlap with the butterflies of the other half inside every warp, nothing else changes
    constexpr bool STW = VAR == 11;
    for (int r = 0; r < reps; r++) {
      double2 a0[16], a1[16];
      const int id0 = tid, id1 = tid + 256;
      const int bp0 = ((id0 >> 5) << 9) | (id0 & 31), bp1 = ((id1 >> 5) << 9) | (id1 & 31);
      static_for<0, 16>([&](auto q_) { constexpr int q = decltype(q_)::value; a0[q] = lds128(xb + paddr(bp0 | (q << 5))); });
      static_for<0, 16>([&](auto q_) { constexpr int q = decltype(q_)::value; a1[q] = lds128(xb + paddr(bp1 | (q << 5))); });
      bfly16<STW>(a0, tw, tid);
      static_for<0, 16>([&](auto q_) { constexpr int q = decltype(q_)::value; sts128(xb + paddr(bp0 | (q << 5)), a0[q]); });
      bfly16<STW>(a1, tw, tid);
      static_for<0, 16>([&](auto q_) { constexpr int q = decltype(q_)::value; sts128(xb + paddr(bp1 | (q << 5)), a1[q]); });
      __syncthreads();
    }
    static_for<0, 16>([&](auto q_) { constexpr int q = decltype(q_)::value; v[q] = xb[paddr(tid + 256 * q)]; });
  } else if constexpr (VAR >= 7 && VAR <= 9) {
    // K2's sequence per row: F0 X F1 X F2 (divide) I2 X I1 X I0.  7: four CTA-wide exchanges (today);
    // 8: the inner two exchanges warp-local (a warp owns its slot set, __syncwarp only); 9: as 8, twiddles from smem
    constexpr bool STW = VAR == 9;
    const int w = tid >> 5, l = tid & 31;
    double2* wb = xb + w * (1024 + 32);
    auto xcta = [&] {
      exch_store<256>(v, xb, tid);
      __syncthreads();
      exch_load<256>(v, xb, tid);
    };
    auto xwarp = [&] {
      static_for<0, E>([&](auto e_) { constexpr int e = decltype(e_)::value; wb[paddr(l + 32 * e)] = v[e]; });
      __syncwarp();
      static_for<0, E>([&](auto e_) { constexpr int e = decltype(e_)::value; v[e] = wb[paddr(l * 32 + e)]; });
    };
    for (int r = 0; r < reps; r++) {
      compute<STW>(v, tw, tid);
      xcta();
      compute<STW>(v, tw, tid);
      if constexpr (VAR == 7) xcta(); else xwarp();
      compute<false>(v, tw, tid);
      compute<false>(v, tw, tid);
      if constexpr (VAR == 7) xcta(); else xwarp();
      compute<STW>(v, tw, tid);
      xcta();
      compute<STW>(v, tw, tid);
    }
  } else {
    const int g = tid >> 7, t = tid & 127;
    double2* sb = xb + g * (N / 2 + N / 64 + 16);
    constexpr bool STW = VAR == 6;
    if ((VAR == 4 || VAR == 6) && g == 1) {
      exch_store<128>(v, sb, t);
      bar_named(1 + g, 128);
      exch_load<128>(v, sb, t);
    }
    for (int r = 0; r < reps; r++) {
      compute<STW>(v, tw, t);
      exch_store<128>(v, sb, t);
      bar_named(1 + g, 128);
      exch_load<128>(v, sb, t);
    }
  }
  const long long t1 = clock64();
  double s = 0;
  static_for<0, E>([&](auto e_) { constexpr int e = decltype(e_)::value; s += v[e].x + v[e].y; });
  out[blockIdx.x * 256 + tid] = s;
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int VAR>
void run(const char* name, double* out, long long* cyc, int reps) {
  const size_t smem = sizeof(double2) * (N + N / 32 + 64 + 4096);
  cudaFuncSetAttribute(kern<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  kern<VAR><<<148, 256, smem>>>(out, cyc, 4);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  cudaEventRecord(a);
  kern<VAR><<<148, 256, smem>>>(out, cyc, reps);
  cudaEventRecord(b);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int i = 0; i < 148; i++) mean += (double)h[i] / 148;
  printf("variant %d  %-58s %8.0f cycles/iteration  (%.3f ms, %s)\n", VAR, name, mean / reps, ms, cudaGetErrorString(e));
}

int main() {
  double* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 256 * sizeof(double));
  cudaMalloc(&cyc, 148 * sizeof(long long));
  const int reps = 2000;
  run<0>("compute only (512 FP64 instr/thread)", out, cyc, reps);
  run<1>("exchange only (32 STS.128 + bar + 32 LDS.128)", out, cyc, reps);
  run<2>("compute + exchange, CTA-wide barrier", out, cyc, reps);
  run<3>("two half-CTAs, named barriers, same order", out, cyc, reps);
  run<4>("two half-CTAs, named barriers, opposite order", out, cyc, reps);
  run<5>("as 2 + twiddles from shared memory", out, cyc, reps);
  run<6>("as 4 + twiddles from shared memory", out, cyc, reps);
  run<10>("software-pipelined pass (LDS u0,u1 | F u0 | STS u0 | F u1 | STS u1)", out, cyc, reps);
  run<11>("as 10 + twiddles from shared memory", out, cyc, reps);
  run<7>("K2 sequence per row, 4 CTA-wide exchanges", out, cyc, reps / 4);
  run<8>("K2 sequence per row, inner 2 exchanges warp-local", out, cyc, reps / 4);
  run<9>("as 8 + twiddles from shared memory", out, cyc, reps / 4);
  return 0;
}
