// vmk_backend.cuh -- where kernel bodies execute.
//
// Product build (default): CUDA.  Every body becomes a __global__ kernel, memory is device memory,
// errors are CUDA errors.  There is no CPU fallback in this build.
//
// -DVMK_EMUL (tests/emul only, builds libvmk_emul.so with vmke_* symbols): the same bodies and the same
// plan logic run on the host with one pthread per CUDA thread, so the index math can be tested in CI
// without a GPU.  Test infrastructure; never loaded by the package.
#pragma once
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <mutex>
#include <string>

#include "vmk_common.cuh"

namespace vmk {

inline std::string& err_slot() {
  static thread_local std::string e;
  return e;
}
inline int fail(int code, const std::string& msg) {
  err_slot() = msg;
  return code;
}

#ifndef VMK_EMUL
// ================================ CUDA backend =====================================================
#define VMK_CUDA_TRY(expr)                                                                     \
  do {                                                                                         \
    cudaError_t e__ = (expr);                                                                  \
    if (e__ != cudaSuccess)                                                                    \
      return ::vmk::fail(2, std::string(#expr) + ": " + cudaGetErrorName(e__) + ": " + cudaGetErrorString(e__)); \
  } while (0)

struct Stream {
  cudaStream_t s = nullptr;
};
struct Event {
  cudaEvent_t e = nullptr;
};

template <class Body, class Args, int CT, int MINB>
__global__ void __launch_bounds__(CT, MINB) body_kernel(const __grid_constant__ Args a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Ctx c;
  c.tid = (int)threadIdx.x;
  c.bid = (int)blockIdx.x;
  c.nblk = (int)gridDim.x;
  c.smem = smem_raw;
  c.hbar = nullptr;
  c.hscratch = nullptr;
  c.hcsmem = nullptr;
  unsigned cr, cs;  // 0 / 1 when the kernel is launched without a cluster dimension
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cr));
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(cs));
  c.crank = (int)cr;
  c.csize = (int)cs;
  Body::run(c, a);
}

inline int be_malloc(void** p, size_t bytes) {
  VMK_CUDA_TRY(cudaMalloc(p, bytes));
  return 0;
}
inline void be_free(void* p) {
  if (p) cudaFree(p);
}
inline int be_stream_create(Stream& s) {
  VMK_CUDA_TRY(cudaStreamCreateWithFlags(&s.s, cudaStreamNonBlocking));
  return 0;
}
inline void be_stream_destroy(Stream& s) {
  if (s.s) cudaStreamDestroy(s.s);
  s.s = nullptr;
}
inline bool be_stream_valid(const Stream& s) { return s.s != nullptr; }
inline int be_event_create(Event& e) {
  VMK_CUDA_TRY(cudaEventCreate(&e.e));
  return 0;
}
inline void be_event_destroy(Event& e) {
  if (e.e) cudaEventDestroy(e.e);
  e.e = nullptr;
}
inline int be_event_record(Event& e, Stream& s) {
  VMK_CUDA_TRY(cudaEventRecord(e.e, s.s));
  return 0;
}
inline int be_stream_wait(Stream& s, Event& e) {
  VMK_CUDA_TRY(cudaStreamWaitEvent(s.s, e.e, 0));
  return 0;
}
inline int be_event_elapsed(Event& a, Event& b, double* ms) {
  VMK_CUDA_TRY(cudaEventSynchronize(b.e));
  float f = 0.f;
  VMK_CUDA_TRY(cudaEventElapsedTime(&f, a.e, b.e));
  *ms = (double)f;
  return 0;
}
inline int be_sync(Stream& s) {
  VMK_CUDA_TRY(cudaStreamSynchronize(s.s));
  return 0;
}
inline int be_h2d(void* dst, const void* src, size_t bytes, Stream& s) {
  VMK_CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s.s));
  return 0;
}
inline int be_d2h(void* dst, const void* src, size_t bytes, Stream& s) {
  VMK_CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s.s));
  return 0;
}
inline int be_d2d(void* dst, const void* src, size_t bytes, Stream& s) {
  VMK_CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, s.s));
  return 0;
}
inline int be_h2d_2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height, Stream& s) {
  VMK_CUDA_TRY(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, cudaMemcpyHostToDevice, s.s));
  return 0;
}
// device (possibly a peer's) <- device, strided; runs on a copy engine
inline int be_d2d_2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height, Stream& s) {
  VMK_CUDA_TRY(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, cudaMemcpyDefault, s.s));
  return 0;
}
inline int be_d2h_2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height, Stream& s) {
  VMK_CUDA_TRY(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, cudaMemcpyDeviceToHost, s.s));
  return 0;
}
// inter-process handles of device allocations (one process per GPU)
typedef cudaIpcMemHandle_t IpcHandle;
inline int be_ipc_export(void* p, IpcHandle* h) {
  VMK_CUDA_TRY(cudaIpcGetMemHandle(h, p));
  return 0;
}
inline int be_ipc_open(const IpcHandle& h, void** p) {
  VMK_CUDA_TRY(cudaIpcOpenMemHandle(p, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}
inline void be_ipc_close(void* p) { cudaIpcCloseMemHandle(p); }
inline int be_get_device(int* dev) {
  VMK_CUDA_TRY(cudaGetDevice(dev));
  return 0;
}
inline int be_set_device(int dev) {
  VMK_CUDA_TRY(cudaSetDevice(dev));
  return 0;
}
// one host thread driving several devices: let `dev` (current) load from / store to `peer`
inline int be_enable_peer(int dev, int peer) {
  if (dev == peer) return 0;
  int can = 0;
  VMK_CUDA_TRY(cudaDeviceCanAccessPeer(&can, dev, peer));
  if (!can) return fail(2, "devices cannot access each other's memory (no NVLink/PCIe peer path)");
  cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    cudaGetLastError();
    return 0;
  }
  VMK_CUDA_TRY(e);
  return 0;
}
inline int be_num_sms(int* n) {
  int dev = 0;
  VMK_CUDA_TRY(cudaGetDevice(&dev));
  VMK_CUDA_TRY(cudaDeviceGetAttribute(n, cudaDevAttrMultiProcessorCount, dev));
  return 0;
}

// occupancy query + shared-memory opt-in for one kernel on the CURRENT device; resident = SMs x CTAs/SM
template <class Body, class Args, int CT, int MINB>
inline int be_configure(size_t smem, int* resident) {
  auto kfn = body_kernel<Body, Args, CT, MINB>;
  if (smem > 48 * 1024)
    VMK_CUDA_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0, sms = 0;
  VMK_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kfn, CT, smem));
  if (be_num_sms(&sms)) return 2;
  if (occ < 1) return fail(2, "kernel does not fit on an SM (shared memory / registers)");
  *resident = occ * sms;
  return 0;
}

template <class Body, class Args, int CT, int MINB>
inline int be_launch(int grid, size_t smem, const Args& a, Stream& s) {
  body_kernel<Body, Args, CT, MINB><<<grid, CT, smem, s.s>>>(a);
  VMK_CUDA_TRY(cudaGetLastError());
  return 0;
}

// ---- thread-block clusters: `q` consecutive CTAs of the grid form a cluster (same GPC, distributed shared memory) ----
template <class Body, class Args, int CT, int MINB>
inline void cluster_config(cudaLaunchConfig_t& cfg, cudaLaunchAttribute& at, int grid, int q, size_t smem,
                           cudaStream_t st) {
  cfg = cudaLaunchConfig_t{};
  cfg.gridDim = dim3((unsigned)grid, 1, 1);
  cfg.blockDim = dim3(CT, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  at.id = cudaLaunchAttributeClusterDimension;
  at.val.clusterDim.x = (unsigned)q;
  at.val.clusterDim.y = 1;
  at.val.clusterDim.z = 1;
  cfg.attrs = &at;
  cfg.numAttrs = 1;
}
// resident = CTAs of the clusters that can be co-resident on the device (a multiple of q)
template <class Body, class Args, int CT, int MINB>
inline int be_configure_cluster(int q, size_t smem, int* resident) {
  auto kfn = body_kernel<Body, Args, CT, MINB>;
  if (smem > 48 * 1024)
    VMK_CUDA_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int sms = 0;
  if (be_num_sms(&sms)) return 2;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute at;
  cluster_config<Body, Args, CT, MINB>(cfg, at, (sms / q) * q, q, smem, nullptr);
  int ncl = 0;
  VMK_CUDA_TRY(cudaOccupancyMaxActiveClusters(&ncl, kfn, &cfg));
  if (ncl < 1) return fail(2, "cluster kernel does not fit on the device (shared memory / registers / GPC size)");
  *resident = ncl * q;
  return 0;
}
template <class Body, class Args, int CT, int MINB>
inline int be_launch_cluster(int grid, int q, size_t smem, const Args& a, Stream& s) {
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute at;
  cluster_config<Body, Args, CT, MINB>(cfg, at, grid, q, smem, s.s);
  VMK_CUDA_TRY(cudaLaunchKernelEx(&cfg, body_kernel<Body, Args, CT, MINB>, a));
  return 0;
}

#else
// ================================ host emulation backend (tests only) ================================
struct Stream {
  int dummy = 0;
};
struct Event {
  double t = 0;
};
// "Device memory" of the emulator lives in POSIX shared-memory segments, so that the slab decomposition can be
// run by several PROCESSES (one per rank, gloo for the barrier) exactly like the one-process-per-GPU CUDA path:
// be_ipc_export/open are the stand-ins for cudaIpcGetMemHandle/OpenMemHandle.
struct IpcHandle {
  char name[56];
  size_t bytes;
};
struct EmulAlloc {
  std::string name;
  size_t bytes;
  bool owner;
};
std::map<void*, EmulAlloc>& emul_allocs();
std::mutex& emul_alloc_mutex();
int emul_shm_create(void** p, size_t bytes, std::string* name);
int emul_shm_open(const char* name, size_t bytes, void** p);
void emul_shm_release(void* p, size_t bytes, const char* name, bool owner);

inline int be_malloc(void** p, size_t bytes) {
  std::string name;
  if (emul_shm_create(p, bytes, &name)) return fail(2, "emul: shared-memory allocation failed");
  memset(*p, 0xff, bytes);  // NaN-fill so reads of never-written memory show up
  std::lock_guard<std::mutex> lk(emul_alloc_mutex());
  emul_allocs()[*p] = EmulAlloc{name, bytes, true};
  return 0;
}
inline void be_free(void* p) {
  if (!p) return;
  std::lock_guard<std::mutex> lk(emul_alloc_mutex());
  auto it = emul_allocs().find(p);
  if (it == emul_allocs().end()) return;
  emul_shm_release(p, it->second.bytes, it->second.name.c_str(), it->second.owner);
  emul_allocs().erase(it);
}
inline int be_ipc_export(void* p, IpcHandle* h) {
  std::lock_guard<std::mutex> lk(emul_alloc_mutex());
  auto it = emul_allocs().find(p);
  if (it == emul_allocs().end()) return fail(3, "emul: unknown allocation");
  memset(h, 0, sizeof(*h));
  strncpy(h->name, it->second.name.c_str(), sizeof(h->name) - 1);
  h->bytes = it->second.bytes;
  return 0;
}
inline int be_ipc_open(const IpcHandle& h, void** p) {
  if (emul_shm_open(h.name, h.bytes, p)) return fail(2, "emul: cannot map a peer's shared-memory segment");
  std::lock_guard<std::mutex> lk(emul_alloc_mutex());
  emul_allocs()[*p] = EmulAlloc{h.name, h.bytes, false};
  return 0;
}
inline void be_ipc_close(void* p) { be_free(p); }
inline int be_stream_create(Stream&) { return 0; }
inline void be_stream_destroy(Stream&) {}
inline bool be_stream_valid(const Stream&) { return true; }
inline int be_event_create(Event&) { return 0; }
inline void be_event_destroy(Event&) {}
double emul_now_ms();
inline int be_event_record(Event& e, Stream&) {
  e.t = emul_now_ms();
  return 0;
}
inline int be_stream_wait(Stream&, Event&) { return 0; }
inline int be_event_elapsed(Event& a, Event& b, double* ms) {
  *ms = b.t - a.t;
  return 0;
}
inline int be_sync(Stream&) { return 0; }
inline int be_h2d(void* dst, const void* src, size_t bytes, Stream&) {
  memcpy(dst, src, bytes);
  return 0;
}
inline int be_d2h(void* dst, const void* src, size_t bytes, Stream&) {
  memcpy(dst, src, bytes);
  return 0;
}
inline int be_d2d(void* dst, const void* src, size_t bytes, Stream&) {
  memmove(dst, src, bytes);
  return 0;
}
inline int be_h2d_2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height, Stream&) {
  for (size_t r = 0; r < height; r++) memcpy((char*)dst + r * dpitch, (const char*)src + r * spitch, width);
  return 0;
}
inline int be_d2h_2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height, Stream& s) {
  return be_h2d_2d(dst, dpitch, src, spitch, width, height, s);
}
inline int be_d2d_2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height, Stream& s) {
  return be_h2d_2d(dst, dpitch, src, spitch, width, height, s);
}
inline int be_get_device(int* dev) {
  *dev = 0;
  return 0;
}
inline int be_set_device(int) { return 0; }
inline int be_enable_peer(int, int) { return 0; }
inline int be_num_sms(int* n) {
  *n = 4;  // a small "GPU" so that the persistent loops are exercised
  return 0;
}
typedef void (*emul_body_fn)(const Ctx&, const void*);
int emul_run(int grid, int block, int cluster, size_t smem, emul_body_fn fn, const void* args);

template <class Body, class Args, int CT, int MINB>
inline int be_configure(size_t, int* resident) {
  int sms = 0;
  be_num_sms(&sms);
  *resident = sms * MINB;
  return 0;
}
template <class Body, class Args, int CT, int MINB>
inline int be_launch(int grid, size_t smem, const Args& a, Stream&) {
  return emul_run(
      grid, CT, 1, smem, [](const Ctx& c, const void* p) { Body::run(c, *static_cast<const Args*>(p)); }, &a);
}
template <class Body, class Args, int CT, int MINB>
inline int be_configure_cluster(int q, size_t, int* resident) {
  *resident = 2 * q;  // two co-resident clusters, so that the persistent loops are exercised
  return 0;
}
template <class Body, class Args, int CT, int MINB>
inline int be_launch_cluster(int grid, int q, size_t smem, const Args& a, Stream&) {
  return emul_run(
      grid, CT, q, smem, [](const Ctx& c, const void* p) { Body::run(c, *static_cast<const Args*>(p)); }, &a);
}
#endif

}  // namespace vmk
