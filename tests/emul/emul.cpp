// emul.cpp -- host execution of the vmk kernel bodies for CI machines without a GPU.
//
// TEST INFRASTRUCTURE ONLY.  Built (with -DVMK_EMUL) into tests/emul/libvmk_emul.so, which exports the
// C ABI under vmke_* names; the package never loads it.  It exists so that the index arithmetic of the
// kernels (FFT digit permutations, transposed spectrum layout, packed DC/Nyquist row, halo rows, slab
// decomposition) can be checked against the oracle in the "-m 'not gpu'" suite.  It proves nothing about
// the CUDA build's performance, and the parity claims are made by the "-m gpu" tests on the real library.
//
// Execution model: one CTA (or one thread-block cluster) at a time per worker; the threads are ucontext fibers run
// round-robin, Ctx::sync() yields to the scheduler, which resumes the next fiber -- after a full round every fiber
// of the CTA has reached the barrier, exactly __syncthreads() semantics (threads that returned are skipped, as on
// the device).  In a cluster each CTA runs AHEAD as far as it can -- round after round, until all its threads wait
// at a cluster barrier -- before the next CTA of the cluster gets to run, in alternating CTA order from one cluster
// barrier to the next: a missing cluster barrier around a distributed-shared-memory access shows up as a read of
// poisoned or stale data instead of being hidden by lock-step execution.  CTAs / clusters of a grid are
// distributed over a few OS threads.
#include <fcntl.h>
#include <pthread.h>
#include <stdio.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <ucontext.h>

#include <atomic>
#include <thread>
#include <vector>

#include "../../cfd_julia_b200/csrc/vmk_backend.cuh"

namespace {

constexpr size_t kStackBytes = 96 * 1024;

struct Fiber {
  ucontext_t ctx;
  void* stack = nullptr;
  bool done = false;
};

struct Worker {
  ucontext_t sched;
  std::vector<Fiber> fibers;
  int current = -1;
  // launch description
  vmk::emul_body_fn fn = nullptr;
  const void* args = nullptr;
  unsigned char* smem = nullptr;  // shared memory of all CTAs of the cluster, `smem_stride` apart
  size_t smem_stride = 0;
  std::vector<unsigned char*> csmem;  // per-CTA bases
  std::vector<double> scratch;  // warp-shuffle emulation
  int bid = 0, nblk = 0;        // first CTA of the cluster, CTAs in the grid
  int block = 0, cluster = 1;
  int yield_kind = 0;           // what the fiber that just yielded waits for: 0 CTA barrier, 1 cluster barrier
  void* yield_site = nullptr;   // return address of that barrier call = the call site in the (inlined) kernel body
};

void fiber_entry(unsigned lo, unsigned hi) {
  Worker* w = reinterpret_cast<Worker*>(((uintptr_t)hi << 32) | (uintptr_t)lo);
  const int fid = w->current, cr = fid / w->block;
  vmk::Ctx c;
  c.tid = fid % w->block;
  c.bid = w->bid + cr;
  c.nblk = w->nblk;
  c.smem = w->csmem[cr];
  c.hbar = w;
  c.hscratch = w->scratch.data() + 2 * (size_t)cr * w->block;
  c.crank = cr;
  c.csize = w->cluster;
  c.hcsmem = w->csmem.data();
  w->fn(c, w->args);
  w->fibers[fid].done = true;
  swapcontext(&w->fibers[fid].ctx, &w->sched);
}

void run_cluster(Worker* w, int block, int cluster) {
  for (int t = 0; t < block * cluster; t++) {
    Fiber& f = w->fibers[t];
    f.done = false;
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = f.stack;
    f.ctx.uc_stack.ss_size = kStackBytes;
    f.ctx.uc_link = nullptr;
    const uintptr_t p = (uintptr_t)w;
    makecontext(&f.ctx, (void (*)())fiber_entry, 2, (unsigned)(p & 0xffffffffu), (unsigned)(p >> 32));
  }
  std::vector<int> alive(cluster, block);
  int total = block * cluster;
  bool flip = false;
  while (total > 0) {
    // every CTA runs until it is blocked at a cluster barrier (or has finished); then the barrier opens
    for (int k = 0; k < cluster; k++) {
      const int cr = flip ? cluster - 1 - k : k;
      bool blocked = false;
      while (alive[cr] > 0 && !blocked) {
        int kinds[2] = {0, 0};
        void* warp_site = nullptr;
        for (int t = 0; t < block; t++) {
          Fiber& f = w->fibers[cr * block + t];
          if (t % 32 == 0) warp_site = nullptr;
          if (f.done) continue;
          w->current = cr * block + t;
          swapcontext(&w->sched, &f.ctx);
          if (f.done) {
            alive[cr]--;
            total--;
          } else {
            kinds[w->yield_kind]++;
            // On the device the 32 threads of a warp must reach the SAME barrier instruction together: a barrier inside
            // a branch that diverges within a warp hangs (or is undefined).  Fibers cannot hang that way, so check it.
            if (!warp_site) warp_site = w->yield_site;
            if (warp_site != w->yield_site) {
              fprintf(stderr, "emul: threads of one warp (CTA %d, warp %d) wait at different barrier call sites: "
                              "a barrier sits inside a branch that diverges within the warp\n", w->bid + cr, t / 32);
              abort();
            }
          }
        }
        if (kinds[0] && kinds[1]) {
          fprintf(stderr, "emul: threads of one CTA wait at a CTA barrier and at a cluster barrier at once\n");
          abort();
        }
        blocked = kinds[1] > 0;
      }
    }
    flip = !flip;
  }
}

}  // namespace

extern "C" void vmk_host_barrier_wait(void* bar) {
  Worker* w = static_cast<Worker*>(bar);
  w->yield_kind = 0;
  w->yield_site = __builtin_return_address(0);
  swapcontext(&w->fibers[w->current].ctx, &w->sched);
}
extern "C" void vmk_host_cluster_barrier_wait(void* bar) {
  Worker* w = static_cast<Worker*>(bar);
  w->yield_kind = 1;
  w->yield_site = __builtin_return_address(0);
  swapcontext(&w->fibers[w->current].ctx, &w->sched);
}

namespace vmk {

std::map<void*, EmulAlloc>& emul_allocs() {
  static std::map<void*, EmulAlloc> m;
  return m;
}

std::mutex& emul_alloc_mutex() {
  static std::mutex m;
  return m;
}

int emul_shm_create(void** p, size_t bytes, std::string* name) {
  static std::atomic<unsigned> counter{0};
  char buf[56];
  snprintf(buf, sizeof(buf), "/vmke_%d_%u", (int)getpid(), counter.fetch_add(1));
  const int fd = shm_open(buf, O_CREAT | O_EXCL | O_RDWR, 0600);
  if (fd < 0) return 1;
  const size_t len = bytes ? bytes : 1;
  if (ftruncate(fd, (off_t)len)) {
    close(fd);
    shm_unlink(buf);
    return 1;
  }
  void* q = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (q == MAP_FAILED) {
    shm_unlink(buf);
    return 1;
  }
  *p = q;
  *name = buf;
  return 0;
}

int emul_shm_open(const char* name, size_t bytes, void** p) {
  const int fd = shm_open(name, O_RDWR, 0600);
  if (fd < 0) return 1;
  void* q = mmap(nullptr, bytes ? bytes : 1, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (q == MAP_FAILED) return 1;
  *p = q;
  return 0;
}

void emul_shm_release(void* p, size_t bytes, const char* name, bool owner) {
  munmap(p, bytes ? bytes : 1);
  if (owner) shm_unlink(name);
}

double emul_now_ms() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return 1e3 * (double)ts.tv_sec + 1e-6 * (double)ts.tv_nsec;
}

int emul_run(int grid, int block, int cluster, size_t smem, emul_body_fn fn, const void* args) {
  if (grid < 1 || block < 1 || cluster < 1 || grid % cluster) return fail(2, "emul: bad launch dimensions");
  const int nclusters = grid / cluster;
  const size_t stride = ((smem ? smem : 256) + 255) / 256 * 256;
  unsigned hw = std::thread::hardware_concurrency();
  int nworkers = (int)(hw ? hw : 1);
  if (nworkers > nclusters) nworkers = nclusters;
  if (nworkers > 8) nworkers = 8;
  std::atomic<int> next{0};
  std::atomic<int> failed{0};
  auto work = [&]() {
    Worker w;
    w.fn = fn;
    w.args = args;
    w.nblk = grid;
    w.block = block;
    w.cluster = cluster;
    w.fibers.resize((size_t)block * cluster);
    w.scratch.assign(2 * (size_t)block * cluster, 0.0);
    void* sm = nullptr;
    if (posix_memalign(&sm, 256, stride * cluster)) {
      failed = 1;
      return;
    }
    w.smem = static_cast<unsigned char*>(sm);
    w.smem_stride = stride;
    for (int r = 0; r < cluster; r++) w.csmem.push_back(w.smem + stride * r);
    for (auto& f : w.fibers) {
      f.stack = malloc(kStackBytes);
      if (!f.stack) failed = 1;
    }
    if (!failed) {
      for (;;) {
        const int b = next.fetch_add(1);
        if (b >= nclusters) break;
        w.bid = b * cluster;
        memset(w.smem, 0xff, stride * cluster);  // shared memory is uninitialised on the device: poison it
        run_cluster(&w, block, cluster);
      }
    }
    for (auto& f : w.fibers) free(f.stack);
    free(sm);
  };
  if (nworkers == 1) {
    work();
  } else {
    std::vector<std::thread> th;
    for (int i = 0; i < nworkers; i++) th.emplace_back(work);
    for (auto& t : th) t.join();
  }
  return failed ? fail(2, "emul: out of memory") : 0;
}

}  // namespace vmk
