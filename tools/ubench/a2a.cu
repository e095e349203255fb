// a2a.cu -- how fast can one process per... no: ONE process drive the all-to-all of the distributed FFT's transposes over
// NVLink / NVSwitch?  Every device pushes one contiguous block (default 8 MiB = the 8192^2, P = 8 transpose block) to each
// of its P-1 peers at the same time; the time is the slowest device's (CUDA events), the figure is bytes SENT per device
// per second.  Variants:
//   0  cudaMemcpyPeerAsync, one stream per destination
//   1  SM stores: coalesced 16-byte lanes, 8 independent loads in flight per thread (vmk's k6_push_body)
//   2  SM stores, 32 bytes per lane (st.global.v4.f64)
//   3  bulk copies (TMA): global -> shared (cp.async.bulk + mbarrier), shared -> peer global (cp.async.bulk store),
//      a ring of STAGES x CHUNK-byte buffers per CTA: nothing passes through registers and the in-flight volume per SM
//      is the shared memory, not the LSU's store queue
// usage: a2a [block MiB] [ctas] [reps]          (run under gpurun --gpus N; also the source of profiles/r02_a2a_*.txt)
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

constexpr int kMaxDev = 8;
struct Ptrs {
  char* p[kMaxDev];
};

// ---- variant 1 / 2: plain loads and stores --------------------------------------------------------------------------
template <int VEC>  // 16 or 32 bytes per lane
__global__ void __launch_bounds__(256) push_lsu(const char* src, Ptrs dst, size_t block, int rank, int n, int order) {
  constexpr int U = 8;
  const size_t tile = (size_t)256 * VEC * U;  // bytes per work item
  const size_t per_dst = block / tile, items = per_dst * (n - 1);
  for (size_t it = blockIdx.x; it < items; it += gridDim.x) {
    const int q = order ? (int)(it % (n - 1)) : (int)(it / per_dst);
    const size_t w = order ? it / (n - 1) : it % per_dst;
    const int h = (rank + 1 + q) % n;
    const char* s = src + (size_t)h * block + w * tile + (size_t)threadIdx.x * VEC;
    char* d = dst.p[h] + (size_t)rank * block + w * tile + (size_t)threadIdx.x * VEC;
    if constexpr (VEC == 16) {
      double2 v[U];
#pragma unroll
      for (int u = 0; u < U; u++)
        asm volatile("ld.global.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v[u].x), "=d"(v[u].y) : "l"(s + (size_t)u * 256 * VEC));
#pragma unroll
      for (int u = 0; u < U; u++)
        asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1,%2};" ::"l"(d + (size_t)u * 256 * VEC), "d"(v[u].x), "d"(v[u].y) : "memory");
    } else {
      double4 v[U];
#pragma unroll
      for (int u = 0; u < U; u++)
        asm volatile("ld.global.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                     : "=d"(v[u].x), "=d"(v[u].y), "=d"(v[u].z), "=d"(v[u].w)
                     : "l"(s + (size_t)u * 256 * VEC));
#pragma unroll
      for (int u = 0; u < U; u++)
        asm volatile("st.global.L1::no_allocate.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(d + (size_t)u * 256 * VEC), "d"(v[u].x),
                     "d"(v[u].y), "d"(v[u].z), "d"(v[u].w)
                     : "memory");
    }
  }
}

// ---- variant 3: bulk copies through a shared-memory ring ---------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, unsigned parity) {
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(b)), "r"(parity)
        : "memory");
    if (ok) return;
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    if (now - t0 > 2000000000ull) __trap();  // 2 s: never hang the GPU on a protocol error
  }
}
__device__ __forceinline__ void bulk_g2s(void* sdst, const void* gsrc, unsigned bytes, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sdst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(b))
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* ssrc, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes)
               : "memory");
}

template <int STAGES, int CHUNK>
__global__ void __launch_bounds__(32) push_bulk(const char* src, Ptrs dst, size_t block, int rank, int n, int order) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * CHUNK);
  if (threadIdx.x != 0) return;  // one thread drives the copy unit
  for (int s = 0; s < STAGES; s++) mbar_init(full + s, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  const size_t per_dst = block / CHUNK, items = per_dst * (n - 1);
  // this CTA's items: it = blockIdx.x + k * gridDim.x
  const size_t mine = items > blockIdx.x ? (items - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  auto addr = [&](size_t k, const char*& s, char*& d) {
    const size_t it = blockIdx.x + k * gridDim.x;
    const int q = order ? (int)(it % (n - 1)) : (int)(it / per_dst);
    const size_t w = order ? it / (n - 1) : it % per_dst;
    const int h = (rank + 1 + q) % n;
    s = src + (size_t)h * block + w * CHUNK;
    d = dst.p[h] + (size_t)rank * block + w * CHUNK;
  };
  for (size_t k = 0; k < mine + STAGES - 1; k++) {
    if (k < mine) {  // load item k into stage k % STAGES (its previous store, item k - STAGES, has been read out below)
      const int st = (int)(k % STAGES);
      const char* s;
      char* d;
      addr(k, s, d);
      mbar_expect_tx(full + st, CHUNK);
      bulk_g2s(smem + (size_t)st * CHUNK, s, CHUNK, full + st);
    }
    if (k >= STAGES - 1) {  // store item j = k - (STAGES - 1)
      const size_t j = k - (STAGES - 1);
      const int st = (int)(j % STAGES);
      const char* s;
      char* d;
      addr(j, s, d);
      mbar_wait(full + st, (unsigned)((j / STAGES) & 1));
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      bulk_s2g(d, smem + (size_t)st * CHUNK, CHUNK);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      // before stage st is loaded again (item j + STAGES, at iteration k + 1) its store must have read the buffer:
      // allow STAGES - 2 younger groups to be pending
      asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(STAGES > 1 ? STAGES - 2 : 0) : "memory");
    }
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main(int argc, char** argv) {
  const size_t block = (size_t)(argc > 1 ? atof(argv[1]) : 8.0) * (1 << 20);
  const int ctas = argc > 2 ? atoi(argv[2]) : 128;
  const int reps = argc > 3 ? atoi(argv[3]) : 10;
  int n = 0;
  CK(cudaGetDeviceCount(&n));
  if (n > kMaxDev) n = kMaxDev;
  if (n < 2) {
    printf("needs >= 2 devices\n");
    return 0;
  }
  std::vector<char*> src(n), dst(n);
  std::vector<std::vector<cudaStream_t>> st(n);
  std::vector<cudaEvent_t> e0(n), e1(n);
  for (int d = 0; d < n; d++) {
    CK(cudaSetDevice(d));
    for (int p = 0; p < n; p++)
      if (p != d) CK(cudaDeviceEnablePeerAccess(p, 0));
    CK(cudaMalloc(&src[d], block * n));
    CK(cudaMalloc(&dst[d], block * n));
    CK(cudaMemset(src[d], d + 1, block * n));
    CK(cudaMemset(dst[d], 0, block * n));
    st[d].resize(n);
    for (int p = 0; p < n; p++) CK(cudaStreamCreateWithFlags(&st[d][p], cudaStreamNonBlocking));
    CK(cudaEventCreate(&e0[d]));
    CK(cudaEventCreate(&e1[d]));
  }
  Ptrs P;
  for (int d = 0; d < kMaxDev; d++) P.p[d] = d < n ? dst[d] : nullptr;
  constexpr int STAGES = 4, CHUNK = 32 * 1024;
  const size_t bulk_smem = (size_t)STAGES * CHUNK + 64;
  for (int d = 0; d < n; d++) {
    CK(cudaSetDevice(d));
    CK(cudaFuncSetAttribute(push_bulk<STAGES, CHUNK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bulk_smem));
  }
  auto sync_all = [&] {
    for (int d = 0; d < n; d++) {
      CK(cudaSetDevice(d));
      CK(cudaDeviceSynchronize());
    }
  };
  struct Var {
    int id, order, grid;
    const char* name;
  };
  const Var vars[] = {
      {0, 0, 0, "cudaMemcpyPeerAsync, one stream per destination"},
      {1, 0, ctas, "SM stores 16 B/lane, destination-major"},
      {1, 1, ctas, "SM stores 16 B/lane, destinations interleaved"},
      {1, 1, 2 * ctas, "SM stores 16 B/lane, interleaved, 2x CTAs"},
      {2, 1, ctas, "SM stores 32 B/lane, interleaved"},
      {3, 0, ctas, "bulk copies (4 x 32 KB ring per CTA), destination-major"},
      {3, 1, ctas, "bulk copies (4 x 32 KB ring per CTA), interleaved"},
      {3, 1, 32, "bulk copies, interleaved, 32 CTAs"},
      {3, 1, 16, "bulk copies, interleaved, 16 CTAs"},
  };
  printf("%d devices, %.1f MiB to each of %d peers = %.1f MB sent per device\n", n, block / 1048576.0, n - 1,
         block * (n - 1) / 1e6);
  for (const Var& v : vars) {
    double best = 1e30;
    for (int r = 0; r < reps + 2; r++) {
      sync_all();
      for (int d = 0; d < n; d++) {
        CK(cudaSetDevice(d));
        CK(cudaEventRecord(e0[d], st[d][d]));
        if (v.id == 0) {
          for (int q = 0; q + 1 < n; q++) {
            const int h = (d + 1 + q) % n;
            CK(cudaStreamWaitEvent(st[d][h], e0[d], 0));
            CK(cudaMemcpyPeerAsync(dst[h] + (size_t)d * block, h, src[d] + (size_t)h * block, d, block, st[d][h]));
          }
        } else if (v.id == 1) {
          push_lsu<16><<<v.grid, 256, 0, st[d][d]>>>(src[d], P, block, d, n, v.order);
        } else if (v.id == 2) {
          push_lsu<32><<<v.grid, 256, 0, st[d][d]>>>(src[d], P, block, d, n, v.order);
        } else {
          push_bulk<STAGES, CHUNK><<<v.grid, 32, bulk_smem, st[d][d]>>>(src[d], P, block, d, n, v.order);
        }
      }
      if (v.id == 0) {
        for (int d = 0; d < n; d++) {
          CK(cudaSetDevice(d));
          for (int q = 0; q + 1 < n; q++) {
            const int h = (d + 1 + q) % n;
            cudaEvent_t e;
            CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            CK(cudaEventRecord(e, st[d][h]));
            CK(cudaStreamWaitEvent(st[d][d], e, 0));
            CK(cudaEventDestroy(e));
          }
        }
      }
      for (int d = 0; d < n; d++) {
        CK(cudaSetDevice(d));
        CK(cudaEventRecord(e1[d], st[d][d]));
      }
      sync_all();
      double worst = 0;
      for (int d = 0; d < n; d++) {
        float ms = 0;
        CK(cudaSetDevice(d));
        CK(cudaEventElapsedTime(&ms, e0[d], e1[d]));
        if (ms > worst) worst = ms;
      }
      if (r >= 2 && worst < best) best = worst;
    }
    // verify one byte per block
    bool ok = true;
    for (int d = 0; d < n && ok; d++) {
      CK(cudaSetDevice(d));
      for (int s = 0; s < n; s++) {
        if (s == d) continue;
        char c = 0;
        CK(cudaMemcpy(&c, dst[d] + (size_t)s * block + block - 1, 1, cudaMemcpyDeviceToHost));
        if (c != (char)(s + 1)) ok = false;
      }
      CK(cudaMemset(dst[d], 0, block * n));
    }
    printf("variant %d order %d grid %4d  %-58s %8.3f ms  %7.1f GB/s per device %s\n", v.id, v.order, v.grid, v.name, best,
           block * (n - 1) / best / 1e6, ok ? "" : "DATA MISMATCH");
    fflush(stdout);
  }
  return 0;
}
