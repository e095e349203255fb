// vmk.cu -- plan object and C ABI (include/vmk.h) of the vortex-merger step on B200.
//
// Host-side orchestration only: which kernel runs on which buffer in which order.  All arithmetic on
// field data happens in the kernel bodies of vmk_kernels.cuh, on the device.  There is no CPU
// compute path in this file; without a CUDA device every entry point returns VMK_ECUDA.
//
// Reference lines replaced (relative to the CFD_Julia checkout):
//   vmk_plan_create   vm.jl:13-19 (per-call allocations), Common.jl:98-113 (k table), FFTW plans :117,:123
//   enqueue_poisson   Common.jl:115-123   (K1, K2, K3)
//   enqueue_stage     Common.jl:132-182 + vm.jl:28-38 / 43-57 / 62-76 (K1..K4)
//   vmk_numerical     vm.jl:24-89
#ifdef VMK_EMUL
// the host-emulation build (tests/emul) exports the same ABI under vmke_* so both can be loaded at once
#define vmk_version vmke_version
#define vmk_last_error vmke_last_error
#define vmk_plan_create vmke_plan_create
#define vmk_plan_create_slab vmke_plan_create_slab
#define vmk_plan_create_on vmke_plan_create_on
#define vmk_plan_destroy vmke_plan_destroy
#define vmk_fps vmke_fps
#define vmk_ps_fft vmke_ps_fft
#define vmk_rhs vmke_rhs
#define vmk_numerical vmke_numerical
#define vmk_hybrid_numerical vmke_hybrid_numerical
#define vmk_ps23_numerical vmke_ps23_numerical
#define vmk_ps32_numerical vmke_ps32_numerical
#define vmk_print_float64 vmke_print_float64
#define vmk_write_field vmke_write_field
#define vmk_read_field vmke_read_field
#define vmk_ldc_numerical vmke_ldc_numerical
#define vmk_upload vmke_upload
#define vmk_step vmke_step
#define vmk_download vmke_download
#define vmk_sync vmke_sync
#define vmk_stream vmke_stream
#define vmk_step_elapsed_ms vmke_step_elapsed_ms
#define vmk_profile_steps vmke_profile_steps
#define vmk_profile_read vmke_profile_read
#define vmk_profile_tri vmke_profile_tri
#define vmk_launch_count vmke_launch_count
#define vmk_set_option vmke_set_option
#define vmk_device_bytes vmke_device_bytes
#define vmk_peer_blob_bytes vmke_peer_blob_bytes
#define vmk_peer_export vmke_peer_export
#define vmk_peer_import vmke_peer_import
#define vmk_peer_attach_local vmke_peer_attach_local
#define vmk_barrier_hook vmke_barrier_hook
#endif
#include "../../include/vmk.h"

#include <math.h>

#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "vmk_backend.cuh"
#include "vmk_cavity.cuh"
#include "vmk_cluster.cuh"
#include "vmk_hybrid.cuh"
#include "vmk_io.hpp"
#include "vmk_kernels.cuh"
#include "vmk_pseudo.cuh"
#include "vmk_pseudo32.cuh"
#include "vmk_tri.cuh"

using namespace vmk;

// serialises the entry points of one plan; recursive because vmk_numerical & co. call vmk_upload / vmk_step themselves
#ifdef VMK_NO_GUARD  // (test builds only: shows that tests/test_emul.py's two-thread case fails without the lock)
#define VMK_GUARD(p) (void)0
#else
#define VMK_GUARD(p)                                   \
  std::unique_lock<std::recursive_mutex> guard__;      \
  if (p) guard__ = std::unique_lock<std::recursive_mutex>((p)->mu)
#endif

#define VMK_TRY(expr)            \
  do {                           \
    int rc__ = (expr);           \
    if (rc__) return rc__;       \
  } while (0)

namespace {

struct StepParams {
  double dx, dy, dt, re;
  bool operator<(const StepParams& o) const {
    return std::tie(dx, dy, dt, re) < std::tie(o.dx, o.dy, o.dt, o.re);
  }
};

struct SizeOps {
  size_t twn;       // twiddle table entries (cluster sizes: the CTA-level tables, then W_N^n for n < N/cluster)
  size_t smem;      // dynamic shared memory of K1/K2/K3
  int fpc;          // transforms per CTA (per cluster for the cluster sizes)
  int cluster;      // CTAs that share one transform (1: the row fits one SM; 2, 4: vmk_cluster.cuh)
  void (*fill_tw)(double2*);
  void (*fill_ccperm)(const double* cccos, double* ccperm);
  int (*configure)(int* res_k1, int* res_k2, int* res_k3);
  int (*k1)(int grid, const K1Args&, Stream&);
  int (*k2)(int grid, const K2Args&, Stream&);
  int (*k3)(int grid, const K3Args&, Stream&);
  // natural-layout variants for the recurrence form of the solve along j (vmk_tri.cuh); null where it does not apply
  int (*k1n)(int grid, const K1Args&, Stream&);
  int (*k3n)(int grid, const K3Args&, Stream&);
  int (*slot_k)(int s);  // kx held in slot s of a natural-layout spectrum row
  // fused form (vmk_tri.cuh): the recurrences along j inside K1 / K3; null where it does not apply
  int (*kf_configure)(int* res);
  int (*k1f)(int grid, const K1Args&, Stream&);
  int (*k3f)(int grid, const K3Args&, Stream&);
  // fused device-resident loop for small grids (ks_body): one cluster of `q` CTAs; null where it does not apply
  int (*ks_configure)(int q);
  int (*ks)(int q, const KSArgs&, Stream&);
  // hybrid RK3/CN solver (vmk_hybrid.cuh); null for the cluster sizes
  int (*kh_configure)(int* res);
  int (*kh)(int grid, const KHArgs&, Stream&);
  void (*fill_ksqperm)(const double* ksq, double* out);
  // pseudo-spectral solver, 2/3 rule (vmk_pseudo.cuh); null where kh is
  int (*kp_configure)(int* res);
  int (*kp)(int grid, const KPArgs&, Stream&);
  // batched row FFT of the 3/2-rule solver (vmk_pseudo32.cuh); null where kh is
  int (*kx_configure)(int* res);
  int (*kx)(int grid, const KXArgs&, Stream&);
  int (*kxs_configure)(int* res);
  int (*kxs)(int grid, const KXSArgs&, Stream&);
  int (*kxf_configure)(int* res);
  int (*kxf)(int grid, const KXSArgs&, Stream&);
};

template <class C, bool NAT = false>
struct K1Body {
  VMK_HD static void run(const Ctx& c, const K1Args& a) { k1_body<C, NAT>(c, a); }
};
template <class C, bool PIECES = false>
struct K2Body {
  VMK_HD static void run(const Ctx& c, const K2Args& a) { k2_body<C, PIECES>(c, a); }
};
template <class C, bool PIECES = false>
struct K3Body {
  VMK_HD static void run(const Ctx& c, const K3Args& a) { k3_body<C, PIECES ? 1 : 0>(c, a); }
};
template <class C>
struct K3NBody {
  VMK_HD static void run(const Ctx& c, const K3Args& a) { k3_body<C, 2>(c, a); }
};
template <class C>
struct K1FBody {
  VMK_HD static void run(const Ctx& c, const K1Args& a) { k1_body<C, true, true>(c, a); }
};
template <class C>
struct K3FBody {
  VMK_HD static void run(const Ctx& c, const K3Args& a) { k3_body<C, 2, true>(c, a); }
};
struct KTScanF {
  VMK_HD static void run(const Ctx& c, const KTArgs& a) { kt_scanf_body<0>(c, a); }
};
struct KTTotals {
  VMK_HD static void run(const Ctx& c, const KTArgs& a) { kt_totals_body(c, a); }
};
template <int PHASE>
struct KTScan {
  VMK_HD static void run(const Ctx& c, const KTArgs& a) { kt_scan_body<PHASE>(c, a); }
};
struct KTSolve {
  VMK_HD static void run(const Ctx& c, const KTArgs& a) { kt_solve_body(c, a); }
};
struct KTLow {
  VMK_HD static void run(const Ctx& c, const KTArgs& a) { kt_low_body(c, a); }
};
template <class C>
struct KSBody {
  VMK_HD static void run(const Ctx& c, const KSArgs& a) { ks_body<C>(c, a); }
};
template <class C>
struct KHBody {
  VMK_HD static void run(const Ctx& c, const KHArgs& a) { kh_body<C>(c, a); }
};
template <class C>
struct KPBody {
  VMK_HD static void run(const Ctx& c, const KPArgs& a) { kp_body<C>(c, a); }
};
template <class C>
struct KXBody {
  VMK_HD static void run(const Ctx& c, const KXArgs& a) { kx_body<C>(c, a); }
};
template <class C>
struct KXSBody {
  VMK_HD static void run(const Ctx& c, const KXSArgs& a) { kxs_body<C>(c, a); }
};
template <class C>
struct KXFBody {
  VMK_HD static void run(const Ctx& c, const KXSArgs& a) { kxf_body<C>(c, a); }
};
struct P32Init {
  VMK_HD static void run(const Ctx& c, const P32Args& a) { p32_init_body(c, a); }
};
struct P32Spectra {
  VMK_HD static void run(const Ctx& c, const P32Args& a) { p32_spectra_body(c, a); }
};
struct P32Fold {
  VMK_HD static void run(const Ctx& c, const P32Args& a) { p32_fold_body(c, a); }
};
struct P32Unfold {
  VMK_HD static void run(const Ctx& c, const P32Args& a) { p32_unfold_body(c, a); }
};
struct P32Update {
  VMK_HD static void run(const Ctx& c, const P32Args& a) { p32_update_body(c, a); }
};
struct P32Final {
  VMK_HD static void run(const Ctx& c, const P32Args& a) { p32_final_body(c, a); }
};
struct KPProduct {
  VMK_HD static void run(const Ctx& c, const KPProdArgs& a) { kp_product_body(c, a); }
};
template <int MODE>
struct K4Body {
  VMK_HD static void run(const Ctx& c, const K4Args& a) { k4_body<MODE>(c, a); }
};
template <int MODE>
struct KCStage {
  VMK_HD static void run(const Ctx& c, const KCArgs& a) { kc_stage_body<MODE>(c, a); }
};
struct KCBc2 {
  VMK_HD static void run(const Ctx& c, const KCArgs& a) { kc_bc2_body(c, a); }
};
struct KCExtend {
  VMK_HD static void run(const Ctx& c, const KCArgs& a) { kc_extend_body(c, a); }
};
struct KCExtract {
  VMK_HD static void run(const Ctx& c, const KCArgs& a) { kc_extract_body(c, a); }
};
struct KCRmsPartial {
  VMK_HD static void run(const Ctx& c, const KCArgs& a) { kc_rms_partial_body(c, a); }
};
struct KCRmsFinal {
  VMK_HD static void run(const Ctx& c, const KCArgs& a) { kc_rms_final_body(c, a); }
};
struct K6Push {
  VMK_HD static void run(const Ctx& c, const K6Args& a) { k6_push_body(c, a); }
};
struct K5Unpack {
  VMK_HD static void run(const Ctx& c, const K5Args& a) { k5_unpack_body(c, a); }
};
struct K5Pack {
  VMK_HD static void run(const Ctx& c, const K5Args& a) { k5_pack_body(c, a); }
};
struct K5Negate {
  VMK_HD static void run(const Ctx& c, const K5Args& a) { k5_negate_body(c, a); }
};

// twiddles W_B^x = exp(-2 pi i x / B), rounded from long double (same recipe as the oracle's tables)
template <class C>
void fill_twiddles(double2* tw) {
  const long double tau = 6.283185307179586476925286766559005768L;
  for (int k = 0; k < C::P - 1; k++) {
    const int B = 1 << C::hi(k), len = C::tw_len(k), off = C::tw_off(k);
    for (int x = 0; x < len; x++) {
      const long double ang = tau * (long double)x / (long double)B;
      tw[off + x].x = (double)cosl(ang);
      // octant tables hold (cos, +sin) of the first octant (the lookup reflects and signs them), the others W = cos - i sin
      tw[off + x].y = C::tw_oct(k) ? (double)sinl(ang) : (double)(-sinl(ang));
    }
  }
}

// cccos permuted into the register order of K2's divide (see K2Args::ccperm)
template <class C>
void fill_ccperm(const double* cccos, double* out) {
  using F = Fft<C>;
  constexpr int bl = C::bits(C::P - 1), rl = 1 << bl;
  for (int u = 0; u < C::E / rl; u++)
    for (int p = 0; p < rl; p++)
      for (int t = 0; t < C::T; t++) out[(u * rl + p) * C::T + t] = cccos[F::k_of_pos(((t + C::T * u) << bl) | p)];
}

// ---- rows spanning a thread-block cluster (vmk_cluster.cuh) ----------------------------------------------------
template <class C, int Q, bool NAT = false>
struct K1CBody {
  VMK_HD static void run(const Ctx& c, const K1Args& a) { k1c_body<C, Q, NAT>(c, a); }
};
template <class C, int Q>
struct K3CNBody {
  VMK_HD static void run(const Ctx& c, const K3Args& a) { k3c_body<C, Q, 2>(c, a); }
};
template <class C, int Q, bool PIECES = false>
struct K2CBody {
  VMK_HD static void run(const Ctx& c, const K2Args& a) { k2c_body<C, Q, PIECES>(c, a); }
};
template <class C, int Q, bool PIECES = false>
struct K3CBody {
  VMK_HD static void run(const Ctx& c, const K3Args& a) { k3c_body<C, Q, PIECES ? 1 : 0>(c, a); }
};

template <class C, int Q>
void fill_twiddles_cluster(double2* tw) {
  fill_twiddles<C>(tw);
  const long double tau = 6.283185307179586476925286766559005768L;
  const int N = Q * C::N;
  for (int n = 0; n < C::N; n++) {  // W_N^n: the twiddles of the radix-Q pass across the cluster
    const long double ang = tau * (long double)n / (long double)N;
    tw[C::TWN + n].x = (double)cosl(ang);
    tw[C::TWN + n].y = (double)(-sinl(ang));
  }
}

// CTA r of the cluster holds ky = Q k' + r: one permuted table per r
template <class C, int Q>
void fill_ccperm_cluster(const double* cccos, double* out) {
  using F = Fft<C>;
  constexpr int bl = C::bits(C::P - 1), rl = 1 << bl;
  for (int r = 0; r < Q; r++)
    for (int u = 0; u < C::E / rl; u++)
      for (int p = 0; p < rl; p++)
        for (int t = 0; t < C::T; t++)
          out[(size_t)r * C::N + (u * rl + p) * C::T + t] = cccos[Q * F::k_of_pos(((t + C::T * u) << bl) | p) + r];
}

template <int MSUB, int Q>
SizeOps make_cluster_ops() {
  using C = typename CfgFor<MSUB>::type;
  SizeOps o;
  o.twn = C::TWN + C::N;
  o.smem = C::SMEM_BYTES;
  o.fpc = C::FPC;
  o.cluster = Q;
  o.fill_tw = &fill_twiddles_cluster<C, Q>;
  o.fill_ccperm = &fill_ccperm_cluster<C, Q>;
  o.configure = [](int* r1, int* r2, int* r3) -> int {
    VMK_TRY((be_configure_cluster<K1CBody<C, Q>, K1Args, C::CT, 1>(Q, C::SMEM_BYTES, r1)));
    VMK_TRY((be_configure_cluster<K2CBody<C, Q>, K2Args, C::CT, 1>(Q, C::SMEM_BYTES, r2)));
    VMK_TRY((be_configure_cluster<K3CBody<C, Q>, K3Args, C::CT, 1>(Q, C::SMEM_BYTES, r3)));
    int d2 = 0, d3 = 0;
    VMK_TRY((be_configure_cluster<K2CBody<C, Q, true>, K2Args, C::CT, 1>(Q, C::SMEM_BYTES, &d2)));
    VMK_TRY((be_configure_cluster<K3CBody<C, Q, true>, K3Args, C::CT, 1>(Q, C::SMEM_BYTES, &d3)));
    if (d2 < *r2) *r2 = d2;
    if (d3 < *r3) *r3 = d3;
    int d1 = 0;
    VMK_TRY((be_configure_cluster<K1CBody<C, Q, true>, K1Args, C::CT, 1>(Q, C::SMEM_BYTES, &d1)));
    VMK_TRY((be_configure_cluster<K3CNBody<C, Q>, K3Args, C::CT, 1>(Q, C::SMEM_BYTES, &d3)));
    if (d1 < *r1) *r1 = d1;
    if (d3 < *r3) *r3 = d3;
    return 0;
  };
  o.k1 = [](int grid, const K1Args& a, Stream& s) -> int {
    return be_launch_cluster<K1CBody<C, Q>, K1Args, C::CT, 1>(grid, Q, C::SMEM_BYTES, a, s);
  };
  o.k2 = [](int grid, const K2Args& a, Stream& s) -> int {
    return a.pieces ? be_launch_cluster<K2CBody<C, Q, true>, K2Args, C::CT, 1>(grid, Q, C::SMEM_BYTES, a, s)
                    : be_launch_cluster<K2CBody<C, Q>, K2Args, C::CT, 1>(grid, Q, C::SMEM_BYTES, a, s);
  };
  o.k3 = [](int grid, const K3Args& a, Stream& s) -> int {
    return a.pieces ? be_launch_cluster<K3CBody<C, Q, true>, K3Args, C::CT, 1>(grid, Q, C::SMEM_BYTES, a, s)
                    : be_launch_cluster<K3CBody<C, Q>, K3Args, C::CT, 1>(grid, Q, C::SMEM_BYTES, a, s);
  };
  o.k1n = [](int grid, const K1Args& a, Stream& s) -> int {
    return be_launch_cluster<K1CBody<C, Q, true>, K1Args, C::CT, 1>(grid, Q, C::SMEM_BYTES, a, s);
  };
  o.k3n = [](int grid, const K3Args& a, Stream& s) -> int {
    return be_launch_cluster<K3CNBody<C, Q>, K3Args, C::CT, 1>(grid, Q, C::SMEM_BYTES, a, s);
  };
  o.slot_k = [](int s) -> int {
    constexpr int HP = C::N / 2;  // slots per CTA of the cluster
    return Q * own_half_k<C>((s % HP) % C::T, (s % HP) / C::T) + s / HP;
  };
  o.kf_configure = nullptr;
  o.k1f = nullptr;
  o.k3f = nullptr;
  o.ks_configure = nullptr;
  o.ks = nullptr;
  o.kh_configure = nullptr;
  o.kh = nullptr;
  o.fill_ksqperm = nullptr;
  o.kp_configure = nullptr;
  o.kp = nullptr;
  o.kx_configure = nullptr;
  o.kx = nullptr;
  o.kxs_configure = nullptr;
  o.kxs = nullptr;
  o.kxf_configure = nullptr;
  o.kxf = nullptr;
  return o;
}

// sizes with a natural-layout K1 / K3 (the recurrence form of the solve along j, vmk_tri.cuh)
template <class C>
constexpr bool kTriSize = !C::SPLIT && C::M >= 6;
// ... and with the fused form (whole warps per transform: the tensor-memory state is addressed per warp)
template <class C>
constexpr bool kFusedSize = kTriSize<C> && C::T >= 32;

template <int M>
SizeOps make_ops() {
  using C = typename CfgFor<M>::type;
  SizeOps o;
  o.twn = C::TWN;
  o.smem = C::SMEM_BYTES;
  o.fpc = C::FPC;
  o.cluster = 1;
  o.fill_tw = &fill_twiddles<C>;
  o.fill_ccperm = &fill_ccperm<C>;
  o.configure = [](int* r1, int* r2, int* r3) -> int {
    VMK_TRY((be_configure<K1Body<C>, K1Args, C::CT, C::MINB>(C::SMEM_BYTES, r1)));
    VMK_TRY((be_configure<K2Body<C>, K2Args, C::CT, C::MINB>(C::SMEM_BYTES, r2)));
    VMK_TRY((be_configure<K3Body<C>, K3Args, C::CT, C::MINB>(C::SMEM_BYTES, r3)));
    int dummy = 0;
    VMK_TRY((be_configure<K2Body<C, true>, K2Args, C::CT, C::MINB>(C::SMEM_BYTES, &dummy)));
    VMK_TRY((be_configure<K3Body<C, true>, K3Args, C::CT, C::MINB>(C::SMEM_BYTES, &dummy)));
    if constexpr (kTriSize<C>) {
      VMK_TRY((be_configure<K1Body<C, true>, K1Args, C::CT, C::MINB>(C::SMEM_BYTES, &dummy)));
      VMK_TRY((be_configure<K3NBody<C>, K3Args, C::CT, C::MINB>(C::SMEM_BYTES, &dummy)));
    }
    return 0;
  };
  o.k1n = nullptr;
  o.k3n = nullptr;
  o.slot_k = nullptr;
  o.kf_configure = nullptr;
  o.k1f = nullptr;
  o.k3f = nullptr;
  if constexpr (kFusedSize<C>) {
    o.kf_configure = [](int* r) -> int {
      int r1 = 0, r3 = 0;
      VMK_TRY((be_configure<K1FBody<C>, K1Args, C::CT, C::MINB>(C::SMEM_BYTES, &r1)));
      VMK_TRY((be_configure<K3FBody<C>, K3Args, C::CT, C::MINB>(C::SMEM_BYTES, &r3)));
      *r = r1 < r3 ? r1 : r3;
      return 0;
    };
    o.k1f = [](int grid, const K1Args& a, Stream& s) -> int {
      return be_launch<K1FBody<C>, K1Args, C::CT, C::MINB>(grid, C::SMEM_BYTES, a, s);
    };
    o.k3f = [](int grid, const K3Args& a, Stream& s) -> int {
      return be_launch<K3FBody<C>, K3Args, C::CT, C::MINB>(grid, C::SMEM_BYTES, a, s);
    };
  }
  if constexpr (kTriSize<C>) {
    o.k1n = [](int grid, const K1Args& a, Stream& s) -> int {
      return be_launch<K1Body<C, true>, K1Args, C::CT, C::MINB>(grid, C::SMEM_BYTES, a, s);
    };
    o.k3n = [](int grid, const K3Args& a, Stream& s) -> int {
      return be_launch<K3NBody<C>, K3Args, C::CT, C::MINB>(grid, C::SMEM_BYTES, a, s);
    };
    o.slot_k = [](int s) -> int { return own_half_k<C>(s % C::T, s / C::T); };
  }
  o.k1 = [](int grid, const K1Args& a, Stream& s) -> int {
    return be_launch<K1Body<C>, K1Args, C::CT, C::MINB>(grid, C::SMEM_BYTES, a, s);
  };
  o.k2 = [](int grid, const K2Args& a, Stream& s) -> int {
    return a.pieces ? be_launch<K2Body<C, true>, K2Args, C::CT, C::MINB>(grid, C::SMEM_BYTES, a, s)
                    : be_launch<K2Body<C>, K2Args, C::CT, C::MINB>(grid, C::SMEM_BYTES, a, s);
  };
  o.k3 = [](int grid, const K3Args& a, Stream& s) -> int {
    return a.pieces ? be_launch<K3Body<C, true>, K3Args, C::CT, C::MINB>(grid, C::SMEM_BYTES, a, s)
                    : be_launch<K3Body<C>, K3Args, C::CT, C::MINB>(grid, C::SMEM_BYTES, a, s);
  };
  o.ks_configure = nullptr;
  o.ks = nullptr;
  if constexpr (!C::SPLIT && M <= 8) {
    o.ks_configure = [](int q) -> int {
      int r = 0;
      return be_configure_cluster<KSBody<C>, KSArgs, C::CT, 1>(q, C::SMEM_BYTES, &r);
    };
    o.ks = [](int q, const KSArgs& a, Stream& s) -> int {
      return be_launch_cluster<KSBody<C>, KSArgs, C::CT, 1>(q, q, C::SMEM_BYTES, a, s);
    };
  }
  if constexpr (C::SPLIT) {  // (measurement / test configurations with the split exchange buffer)
    o.kh_configure = nullptr;
    o.kh = nullptr;
    o.kp_configure = nullptr;
    o.kp = nullptr;
    o.kx_configure = nullptr;
    o.kx = nullptr;
    o.kxs_configure = nullptr;
    o.kxs = nullptr;
    o.kxf_configure = nullptr;
    o.kxf = nullptr;
  } else {
    o.kh_configure = [](int* r) -> int { return be_configure<KHBody<C>, KHArgs, C::CT, C::MINB>(C::SMEM_BYTES, r); };
    o.kh = [](int grid, const KHArgs& a, Stream& s) -> int {
      return be_launch<KHBody<C>, KHArgs, C::CT, C::MINB>(grid, C::SMEM_BYTES, a, s);
    };
    o.kp_configure = [](int* r) -> int { return be_configure<KPBody<C>, KPArgs, C::CT, C::MINB>(C::SMEM_BYTES, r); };
    o.kp = [](int grid, const KPArgs& a, Stream& s) -> int {
      return be_launch<KPBody<C>, KPArgs, C::CT, C::MINB>(grid, C::SMEM_BYTES, a, s);
    };
    o.kx_configure = [](int* r) -> int { return be_configure<KXBody<C>, KXArgs, C::CT, C::MINB>(C::SMEM_BYTES, r); };
    o.kx = [](int grid, const KXArgs& a, Stream& s) -> int {
      return be_launch<KXBody<C>, KXArgs, C::CT, C::MINB>(grid, C::SMEM_BYTES, a, s);
    };
    o.kxs_configure = [](int* r) -> int { return be_configure<KXSBody<C>, KXSArgs, C::CT, C::MINB>(C::SMEM_BYTES, r); };
    o.kxs = [](int grid, const KXSArgs& a, Stream& s) -> int {
      return be_launch<KXSBody<C>, KXSArgs, C::CT, C::MINB>(grid, C::SMEM_BYTES, a, s);
    };
    o.kxf_configure = [](int* r) -> int { return be_configure<KXFBody<C>, KXSArgs, C::CT, C::MINB>(C::SMEM_BYTES, r); };
    o.kxf = [](int grid, const KXSArgs& a, Stream& s) -> int {
      return be_launch<KXFBody<C>, KXSArgs, C::CT, C::MINB>(grid, C::SMEM_BYTES, a, s);
    };
  }
  o.fill_ksqperm = &fill_ccperm<C>;  // the same register-order permutation as the divisor table
  return o;
}

bool ops_for(int M, SizeOps* o) {
  switch (M) {
#ifdef VMK_DEV_SIZES  // kernel-tuning builds only (tools/devbuild.sh): two sizes instead of thirteen, ~2 min of ptxas
    case 7: *o = make_ops<7>(); return true;
    case 8: *o = make_ops<8>(); return true;
    case 10: *o = make_ops<10>(); return true;
    case 13: *o = make_ops<13>(); return true;
    default: return false;
#else
    case 5: *o = make_ops<5>(); return true;
#ifdef VMK_CLUSTER_TEST  // test builds only: the cluster kernels at sizes the oracle (and the emulator) can afford
    case 6: *o = make_cluster_ops<5, 2>(); return true;
    case 7: *o = make_cluster_ops<5, 4>(); return true;
    case 8: *o = make_cluster_ops<7, 2>(); return true;
    case 9: *o = make_cluster_ops<7, 4>(); return true;
    case 10: *o = make_cluster_ops<9, 2>(); return true;
    case 11: *o = make_cluster_ops<9, 4>(); return true;
    case 12: *o = make_cluster_ops<11, 2>(); return true;
    case 13: *o = make_cluster_ops<11, 4>(); return true;
#else
    case 6: *o = make_ops<6>(); return true;
    case 7: *o = make_ops<7>(); return true;
    case 8: *o = make_ops<8>(); return true;
    case 9: *o = make_ops<9>(); return true;
    case 10: *o = make_ops<10>(); return true;
    case 11: *o = make_ops<11>(); return true;
    case 12: *o = make_ops<12>(); return true;
    case 13: *o = make_ops<13>(); return true;
    case 14: *o = make_cluster_ops<13, 2>(); return true;  // 16384 = 2 x 8192
    case 15: *o = make_cluster_ops<13, 4>(); return true;  // 32768 = 4 x 8192
#endif
    default: return false;
#endif
  }
}

int ilog2_exact(int64_t n) {
  int m = 0;
  while ((((int64_t)1) << m) < n) m++;
  return ((((int64_t)1) << m) == n) ? m : -1;
}

// the first four are the classes of vmk_profile_steps / vmk_profile_read; KT1..KT3 (chunk totals, scan, solve of
// vmk_tri.cuh) are folded into K2's class there and listed separately by vmk_profile_tri
enum { KI_K1 = 0, KI_K2, KI_K3, KI_K4, KI_KT1, KI_KT2, KI_KT3, KI_COUNT };
constexpr int KI_ABI = 4;
constexpr int kTriMaxK0 = 64;    // rows kx < K0 keep the FFT form along j (vmk_tri.cuh, tests/models/tri_model.py)
constexpr int kFusedAutoN = 8192;  // (4096^2 measured 0.879 against 0.895 ms/step: kept on the separate kernels) smallest grid for which the fused form is the default on one GPU (measured on a B200 at
                                   // 8192^2: 3.15 against 3.52 ms/step; its units need long blocks of row pairs:
                                   // 4096 pairs over 148 CTAs; 1024^2: 0.23 against 0.10 ms/step)
constexpr int kTriAutoN = 2048;  // smallest grid for which the recurrence form is the default (measured on one B200:
                                 // 1024^2 0.104 against 0.099 ms/step, 2048^2 0.279 / 0.289, 4096^2 0.893 / 0.930, 8192^2 3.50 / 3.87)

}  // namespace

struct vmk_plan {
  // one host thread at a time per plan (the reference is single-threaded; a second thread blocks here)
  std::recursive_mutex mu;
  int N = 0, M = 0, rank = 0, nranks = 1, NJ = 0, log2NJ = 0, j0 = 0;
  int sms = 0;
  int device = 0;  // CUDA device the plan lives on (the one current at creation)
  SizeOps ops{};
  int res_k1 = 0, res_k2 = 0, res_k3 = 0;
  // device buffers
  double* w[3] = {nullptr, nullptr, nullptr};  // wn, wtA, wtB: slabs with halo rows
  double* psi = nullptr;                       // slab with halo rows
  double2* T = nullptr;                        // spectrum rows owned by this rank [N/(2P)][N]   (K1 -> K2)
  double2* V = nullptr;                        // solution spectrum for this rank's j [N/2][NJ]   (K2 -> K3)
  double2* S = nullptr;                        // K1 output [N/2][NJ] before the forward transpose (P > 1 only)
  double2* tw = nullptr;
  double* bbcos = nullptr;
  double* cccos = nullptr;
  double* ccperm = nullptr;
  double* staging = nullptr;  // (NJ+2) x (N+2), allocated on first host-array call
  // hybrid solver (vmk_hybrid_numerical), allocated on first use
  double2* hW = nullptr;      // vorticity spectrum [N/2][N], register order
  double2* hJ = nullptr;      // previous stage's Jacobian spectrum
  double2* hVs = nullptr;     // inverse-j of wf/k2 (V holds the one of wf)
  double* hksq = nullptr;     // kx^2 table and its register-order permutation
  double* hksqperm = nullptr;
  double hyb_dx = 0;
  int res_kh = 0;
  // pseudo-spectral solver (vmk_ps23_numerical), allocated on first use; shares hW, hJ, hksq, hksqperm, V, hVs
  double2* pV[2] = {nullptr, nullptr};  // inverse-j of j3f, j4f (V and hVs hold j1f, j2f)
  double2* pA0 = nullptr;               // kx = 0 part of the packed spectrum row, [N]
  double* ptab = nullptr;               // 8 x [N]: mp, mm, cc, dd (natural order), then the same in register order
  double ps_dx = 0;
  int res_kp = 0;
  // pseudo-spectral solver, 3/2 rule (vmk_ps32_numerical), allocated on first use: an (N/2)^2 plan for the 9 sub-grids of
  // the padded grid (its K1 / K3 / product kernels and its stream) and the natural-order spectra of vmk_pseudo32.cuh
  vmk_plan* child = nullptr;
  double2* qS = nullptr;    // state [L+1][2L+1]
  double2* qJ = nullptr;
  double2* qY = nullptr;    // [4][L+1][3][L]
  double2* qVF = nullptr;   // [4][L/2][9][L]
  double2* qT9 = nullptr;   // [L/2][9][L]
  double2* qPi = nullptr;   // [L+1][3][L]
  double* qF = nullptr;     // [4] real-space batches: (9L+2) rows x L (the 9 sub-grids of one field + two halo rows)
  double2* qtw = nullptr;   // [3L]
  double* qtab = nullptr;   // 5 x [2L+1]
  double q_dx = 0;
  int res_kx = 0, res_kxs = 0, res_kxf = 0;
  int ps32_fuse = 0;  // 1: spectra computed in the load stage of the inverse row transform (kxs_body); 2: and folded along i there (kxf_body)
  // lid-driven cavity (vmk_ldc_numerical), allocated on first use: node arrays (n+1)^2 with n = N/2
  double* cw[3] = {nullptr, nullptr, nullptr};  // wn, wtA, wtB
  double* cs = nullptr;                         // sn
  double* csp = nullptr;                        // sp (previous step's sn)
  double* cpart = nullptr;                      // partial sums of the rms reduction
  int div_kind = 0;                             // 0: fps divisor (Common.jl:101-121), 1: cavity (lid_driven_cavity.jl:66-71)
  int64_t dev_bytes = 0;
  // peers (slab decomposition): pointers to every rank's buffers, own entries included
  double* peer_w[3][kMaxPeers];
  double* peer_psi[kMaxPeers];
  double2* peer_T[kMaxPeers];
  double2* peer_V[kMaxPeers];
  unsigned long long* flags = nullptr;  // [kMaxPeers] epochs published by the peers, [kMaxPeers] own epoch, then error
  unsigned long long* peer_flags[kMaxPeers];
  bool peers_ready = false;
  std::vector<void*> ipc_opened;
  void (*barrier_fn)(void*) = nullptr;  // enqueues a cross-rank barrier on the plan's stream
  void* barrier_user = nullptr;
  // divisor cache (Common.jl:101-113)
  bool div_valid = false;
  double div_dx = 0, div_dy = 0, div_eps = 0, div_aa = 0;
  Stream st, st_copy, st_copy2;
  Event ev0, ev1, ev_join, ev_join2, ev_chunk[8];
  int a2a_chunks = 0 /* auto */, a2a_engine = -1 /* auto */, a2a_ctas = 128, k2_push = -1 /* auto */;
  int a2a_order = 1, k2_chunks = 0 /* auto */;
  bool ev_valid = false;
  bool uploaded = false;
  int64_t launches = 0, graph_launches = 12;
  int k4_rows = 32, k4_ahead = 4, k4_waves = 3;  // k4_waves: measured on 8 GPUs (profiles/r02_notes.md)
  int k1_prefetch = 0, k2_prefetch = 0, v_pieces = 1, cl_prefetch = -1 /* auto */;
  int use_graph = 1;
  // small grids: the step loop as one cluster launch (ks_body); ks_q = CTAs of the cluster.  Opt-in: measured SLOWER than
  // the CUDA graph of 12 launches (128^2: 67.5 against 51.2 us per step, profiles/r02_notes.md section 5)
  int fuse_small = 0, ks_q = 0;
  // solve along j as a cyclic tridiagonal solve by two-sided recurrences (vmk_tri.cuh) instead of K2's FFT pair
  int fps_mode = -1 /* auto: recurrences where the buffers exist and N >= kTriAutoN */, tri_k0 = 0 /* auto */;
  int tri_nch = 0, tri_k0_tab = -1;
  int zigzag = 0, zz = 0;  // consecutive streaming kernels alternate their row direction (K1Args::rev)
  // fused form (fps_mode 2): the recurrences inside K1 / K3; fz_grid CTAs = fz_grid * fpc units of row pairs
  int fz_grid = 0;
  double fz_kc = 0.0;          // psi = (r fz_kc sign) v + dc: the per-slot scale factor of kt_solve_body, tab[6] = r fz_kc
  size_t fz_units_cap = 0;     // blocks the totals / carry buffers were sized for
  bool fz_ok = false;          // tables valid and within the range of the (1/r)^M Horner form
  double* tri_ftab = nullptr;  // [kTriFTab][N/2]
  double2* tri_rr = nullptr;   // [N/2]: (r, 1/r)
  double* tri_tab = nullptr;   // [kTriTab][N/2]
  int* tri_low = nullptr;      // [N/2]
  double2* tri_tot = nullptr;  // [3][nch][N/2]
  double2* tri_cin = nullptr;  // [2 nch + 1][N/2]
  double2* tri_G = nullptr;    // [P][3][N/2]   (peers write their own block)
  double2* tri_L = nullptr;    // [kTriMaxK0][N] (peers write their own columns)
  double2* peer_G[kMaxPeers];
  double2* peer_L[kMaxPeers];
#ifndef VMK_EMUL
  std::map<StepParams, cudaGraphExec_t> graphs;
#endif
  // per-kernel timing (vmk_profile_steps)
  bool profiling = false;
  double prof_ms[KI_COUNT] = {0, 0, 0, 0, 0, 0, 0};
  int64_t prof_n[KI_COUNT] = {0, 0, 0, 0, 0, 0, 0};
  double tri_ms[3] = {0, 0, 0};  // last vmk_profile_steps: totals, scan, solve
  int64_t tri_n[3] = {0, 0, 0};
  std::vector<std::pair<int, std::pair<Event, Event>>> prof_events;
};

namespace {

size_t slab_elems(const vmk_plan* p) { return (size_t)(p->NJ + 2) * p->N; }

int dev_alloc(vmk_plan* p, void** ptr, size_t bytes) {
  VMK_TRY(be_malloc(ptr, bytes));
  p->dev_bytes += (int64_t)bytes;
  return 0;
}

void drop_graphs(vmk_plan* p) {
#ifndef VMK_EMUL
  for (auto& g : p->graphs) cudaGraphExecDestroy(g.second);
  p->graphs.clear();
#else
  (void)p;
#endif
}

int ensure_staging(vmk_plan* p) {
  if (p->staging) return 0;
  return dev_alloc(p, (void**)&p->staging, sizeof(double) * (size_t)(p->NJ + 2) * (p->N + 2));
}

bool tri_on(const vmk_plan* p) {
  // (the cavity solver's divisor tables, div_kind 1, have no slot tables: it keeps the FFT form)
  return p->tri_tab && p->div_kind == 0 && (p->fps_mode >= 1 || (p->fps_mode < 0 && p->N >= kTriAutoN));
}
// the fused form of the recurrences (one GPU): asked for, or by default where it was measured faster
bool fz_on(const vmk_plan* p) {
  return tri_on(p) && p->tri_ftab && p->fz_ok && p->nranks == 1 &&
         (p->fps_mode == 2 || (p->fps_mode < 0 && p->N >= kFusedAutoN));
}
int tri_k0(const vmk_plan* p) {
  const int cap = p->N / 4 < kTriMaxK0 ? p->N / 4 : kTriMaxK0;
  int k0 = p->tri_k0 > 0 ? p->tri_k0 : (p->N / 16 < kTriMaxK0 ? p->N / 16 : kTriMaxK0);
  if (k0 > cap) k0 = cap;
  return k0 < 1 ? 1 : k0;
}

// Per-slot constants of the recurrence form (vmk_tri.cuh; derivation in tests/models/tri_model.py).  The row constant
// is the reference's own FP64 value  ab = fl(aa + fl(bb cos kx))  (Common.jl:120, what K2 adds cc cos(ky) to); the
// rest is evaluated in long double and rounded once.
int fill_tri_tables(vmk_plan* p, const double* bbcos, const double* cccos, double cc) {
  typedef long double L;
  const int H = p->N / 2, k0 = tri_k0(p);
  const double aa = p->div_aa;
  std::vector<double> tab((size_t)kTriTab * H, 0.0);
  std::vector<int> low((size_t)H + kTriMaxK0, -1);  // lowrow[H], then lowslot[K0]
  const L a = (L)cc / 2;
  for (int s = 0; s < H; s++) {
    const int k = p->ops.slot_k(s);
    if (k < k0) {  // keeps the FFT form; neutral constants (its totals are computed but never used)
      low[s] = k;
      low[(size_t)H + k] = s;
      tab[(size_t)3 * H + s] = 1.0;
      continue;
    }
    const double ab = aa + bbcos[k];
    L delta = -((L)ab / a) - 2;
    if (delta < 0) delta = 0;
    const L r = 2 / (2 + delta + sqrtl(delta * (delta + 4)));
    const L R = powl(r, (L)kTriCH), RJ = powl(r, (L)p->NJ), RN = powl(r, (L)p->N);
    const L d_tri0 = (L)ab + (L)cc;
    const L d_ref0 = (L)(double)(ab + cccos[0]);  // fl(ab + fl(cc cos eps)): ky[1] = eps, Common.jl:112-113
    tab[s] = (double)r;
    tab[(size_t)H + s] = (double)R;
    tab[(size_t)2 * H + s] = (double)RJ;
    tab[(size_t)3 * H + s] = (double)(1 / (1 - RN));
    tab[(size_t)4 * H + s] = (double)(r * (1 - R * R) / (1 - r * r));
    tab[(size_t)5 * H + s] = (double)(r * (1 - RJ * RJ) / (1 - r * r));
    tab[(size_t)6 * H + s] = (double)(-r / ((L)cc * (L)p->N));
    tab[(size_t)7 * H + s] = (double)((1 / d_ref0 - 1 / d_tri0) / (L)p->N / (2 * (L)p->N));
    const L Rg = powl(r, (L)(kTriCH * (p->tri_nch / tri_scan_groups(p->tri_nch))));
    tab[(size_t)8 * H + s] = (double)Rg;
    tab[(size_t)9 * H + s] = (double)(r * (1 - Rg * Rg) / (1 - r * r));
  }
  VMK_TRY(be_h2d(p->tri_tab, tab.data(), sizeof(double) * tab.size(), p->st));
  VMK_TRY(be_h2d(p->tri_low, low.data(), sizeof(int) * low.size(), p->st));
  p->fz_ok = false;
  if (p->tri_ftab) {
    // fused form: blocks of 2 Lmin or 2 Lmin + 2 rows (the units of K1 / K3, fz_first)
    const int units = p->fz_grid * p->ops.fpc, npairs = p->NJ / 2;
    const int lmin = npairs / units;
    std::vector<double> ft((size_t)kTriFTab * H, 0.0);
    std::vector<double2> rr(H);
    bool ok = true;
    for (int s = 0; s < H; s++) {
      rr[s].x = rr[s].y = 0.0;
      if (low[s] >= 0) continue;  // r = 0: the slot passes through K3's recurrence unchanged
      const L r = (L)tab[s];      // the rounded r the kernels multiply by
      rr[s].x = (double)r;
      rr[s].y = (double)(1 / r);
      // (1/r)^M and r^M must stay far inside the FP64 range (K1's Horner sum A, K3's q)
      if ((L)(2 * lmin + 2) * log10l(1 / r) > 200) ok = false;
      ft[s] = (double)(r / (1 - r * r));
      for (int cls = 0; cls < 2; cls++) {
        const int M = 2 * (lmin + cls);
        if (M == 0) continue;
        const L R = powl(r, (L)M);
        ft[(size_t)(1 + 3 * cls) * H + s] = (double)R;
        ft[(size_t)(2 + 3 * cls) * H + s] = (double)(r * (1 - R * R) / (1 - r * r));
        ft[(size_t)(3 + 3 * cls) * H + s] = (double)powl(r, (L)(M - 1));
      }
    }
    p->fz_kc = (double)(-1 / ((L)cc * (L)p->N));
    VMK_TRY(be_h2d(p->tri_ftab, ft.data(), sizeof(double) * ft.size(), p->st));
    VMK_TRY(be_h2d(p->tri_rr, rr.data(), sizeof(double2) * rr.size(), p->st));
    p->fz_ok = ok;
  }
  VMK_TRY(be_sync(p->st));
  p->tri_k0_tab = k0;
  return 0;
}

// bb*cos(kx[i]) and cc*cos(ky[j]) exactly as Common.jl:101-113,120 evaluates them (kx[1] = eps, ky = kx).
// 2N cos() evaluations on the host per (dx,dy,eps); the 2 N^2 per call of the reference disappear.
int ensure_divisor(vmk_plan* p, double dx, double dy, double eps) {
  if (p->div_valid && p->div_kind == 0 && p->div_dx == dx && p->div_dy == dy && p->div_eps == eps) return 0;
  p->div_kind = 0;
  const int n = p->N;
  std::vector<double> kx(n), b(n), c(n);
  const double hx = 2.0 * M_PI / (double)n;  // Common.jl:106
  for (int i = 1; i <= n / 2; i++) {         // :108-111
    kx[i - 1] = hx * (double)(i - 1);
    kx[i + n / 2 - 1] = hx * (double)(i - n / 2 - 1);
  }
  kx[0] = eps;  // :112
  const double bb = 2.0 / (dx * dx), cc = 2.0 / (dy * dy);
  for (int i = 0; i < n; i++) {
    const double ck = cos(kx[i]);
    b[i] = bb * ck;
    c[i] = cc * ck;
  }
  VMK_TRY(be_sync(p->st));  // the tables may still be in use by queued kernels
  VMK_TRY(be_h2d(p->bbcos, b.data(), sizeof(double) * n, p->st));
  VMK_TRY(be_h2d(p->cccos, c.data(), sizeof(double) * n, p->st));
  std::vector<double> cp(n);
  p->ops.fill_ccperm(c.data(), cp.data());
  VMK_TRY(be_h2d(p->ccperm, cp.data(), sizeof(double) * n, p->st));
  VMK_TRY(be_sync(p->st));
  p->div_aa = -2.0 / (dx * dx) - 2.0 / (dy * dy);  // :101
  if (p->tri_tab) VMK_TRY(fill_tri_tables(p, b.data(), c.data(), cc));
  p->div_dx = dx;
  p->div_dy = dy;
  p->div_eps = eps;
  p->div_valid = true;
  drop_graphs(p);
  return 0;
}

struct Timed {
  vmk_plan* p;
  int which;
  Event a, b;
  bool on;
  Timed(vmk_plan* p_, int w) : p(p_), which(w), on(p_->profiling) {
    if (on) {
      be_event_create(a);
      be_event_create(b);
      be_event_record(a, p->st);
    }
  }
  void done() {
    if (on) {
      be_event_record(b, p->st);
      p->prof_events.push_back({which, {a, b}});
    }
  }
};

#ifndef VMK_EMUL
// Cross-GPU barrier as a one-warp kernel on the plan's stream: every rank bumps its epoch, publishes it in each
// peer's flag array with a system-scope release store over NVLink and spins (acquire loads on its own array) until
// every peer has published the same epoch.  Release/acquire are cumulative, so everything the earlier kernels of
// the publishing rank wrote -- in its own memory or in a peer's -- is visible to the kernels that follow the
// barrier on the acquiring rank.  The epoch lives in device memory, so the kernel is CUDA-graph replayable.
// One rank per GPU only: two spinning ranks on one device could wait for each other forever.
struct BarrierArgs {
  unsigned long long* peer_flags[kMaxPeers];  // flag arrays of all ranks ([rank] is the local one)
  unsigned long long* epoch;                  // local epoch counter
  int* error;                                 // set to 1 on timeout
  int rank, nranks;
};
__global__ void __launch_bounds__(32, 1) barrier_kernel(const __grid_constant__ BarrierArgs a) {
  const int t = (int)threadIdx.x;
  const unsigned long long e = *a.epoch + 1;
  if (t < a.nranks && t != a.rank) {
    unsigned long long* remote = a.peer_flags[t] + a.rank;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(remote), "l"(e) : "memory");
    const unsigned long long* mine = a.peer_flags[a.rank] + t;
    unsigned long long seen = 0, t0 = 0, now = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (unsigned spin = 0;; spin++) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(mine) : "memory");
      if (seen >= e) break;
      if ((spin & 1023u) == 1023u) {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (now - t0 > 10000000000ull) {  // 10 s: a peer died; do not hang the GPU
          *a.error = 1;
          break;
        }
      }
    }
  }
  __syncwarp();
  if (t == 0) *a.epoch = e;
}
#endif

int cross_rank_barrier(vmk_plan* p) {
  if (p->nranks == 1) return 0;
  if (p->barrier_fn) {
    p->barrier_fn(p->barrier_user);
    return 0;
  }
#ifndef VMK_EMUL
  BarrierArgs a;
  for (int r = 0; r < kMaxPeers; r++) a.peer_flags[r] = r < p->nranks ? p->peer_flags[r] : nullptr;
  a.epoch = p->flags + kMaxPeers;
  a.error = reinterpret_cast<int*>(p->flags + kMaxPeers + 1);
  a.rank = p->rank;
  a.nranks = p->nranks;
  barrier_kernel<<<1, 32, 0, p->st.s>>>(a);
  VMK_CUDA_TRY(cudaGetLastError());
  p->launches++;
  return 0;
#else
  return fail(VMK_ESTATE, "emulated slab plans need a barrier hook (vmk_barrier_hook)");
#endif
}

// cluster kernels: which of K1 (1), K2 (2), K3 (4) bulk-prefetch their share of the next row into L2.  Measured
// (profiles/r01_notes.md): K3's single contiguous block always pays, K1's runs only with 4 CTAs per row, K2's never.
int cl_prefetch_mask(const vmk_plan* p) { return p->cl_prefetch >= 0 ? p->cl_prefetch : (p->ops.cluster == 4 ? 5 : 4); }

// K1/K3 work units: groups of row pairs (see k1_body / k3_body)
int rowpair_units(const vmk_plan* p, int npairs, int g) {
  const int fpc = p->ops.fpc;
  const int nblocks = (npairs + fpc - 1) / fpc;
  return (nblocks + g - 1) / g;
}

int launch_k6(vmk_plan* p, const K6Args& k) {
  const int items = (k.nranks - 1) * ((k.rows + kK6Rows - 1) / kK6Rows) * ((k.ncols + kK6Threads - 1) / kK6Threads);
  const int grid = items < p->a2a_ctas ? items : p->a2a_ctas;
  VMK_TRY((be_launch<K6Push, K6Args, kK6Threads, 2>(grid < 1 ? 1 : grid, 0, k, p->st_copy)));
  p->launches++;
  return 0;
}

// K1 + the forward transpose of the distributed FFT.  On P > 1 GPUs the launch is split into `a2a_chunks` ranges
// of row pairs; as soon as a range is done, its columns of the rows owned by the other ranks are copied into those
// ranks' T buffers by the copy engines on a second stream (NJ/chunks*16-byte contiguous pieces over NVLink), while
// the next range is being transformed.
int launch_k1(vmk_plan* p, const double* src) {
  const int N = p->N, P = p->nranks, R = (N / 2) / P;
  const int npairs = p->NJ / 2;
  int chunks = P > 1 ? (p->a2a_chunks > 0 ? p->a2a_chunks : (P == 2 ? 2 : 1)) : 1;
  while (chunks > 1 && (npairs % chunks || npairs / chunks < 1)) chunks--;
  Timed t(p, KI_K1);
  for (int c = 0; c < chunks; c++) {
    const int pair0 = c * (npairs / chunks), np = npairs / chunks;
    K1Args a;
    a.w = src + (size_t)2 * pair0 * N;
    a.S = P > 1 ? p->S + 2 * pair0 : nullptr;
    a.Tloc = p->T + (size_t)p->rank * R * p->NJ + 2 * pair0;
    a.tw = p->tw;
    a.NJ = p->NJ;
    a.npairs = np;
    a.k_own0 = p->rank * R;
    a.k_own1 = a.k_own0 + R;
    a.prefetch = p->ops.cluster > 1 ? (cl_prefetch_mask(p) & 1) : p->k1_prefetch;
    const int work = rowpair_units(p, np, 1) * p->ops.cluster;
    VMK_TRY(p->ops.k1(work < p->res_k1 ? work : p->res_k1, a, p->st));
    p->launches++;
    if (P > 1) {
      VMK_TRY(be_event_record(p->ev_chunk[c], p->st));
      VMK_TRY(be_stream_wait(p->st_copy, p->ev_chunk[c]));
      VMK_TRY(be_stream_wait(p->st_copy2, p->ev_chunk[c]));
      // measured (profiles/r01_notes.md): copy engines win with one peer (64 KB pieces), the SM push kernel with 3 or 7
      const int engine = p->a2a_engine >= 0 ? p->a2a_engine : (P == 2 ? 1 : 0);
      if (engine) {  // copy engines: one peer copy per destination (contiguous when the launch is not split)
        const size_t rowb = sizeof(double2) * (size_t)p->NJ, width = sizeof(double2) * (size_t)(2 * np);
        for (int q = 0; q + 1 < P; q++) {
          const int h = (p->rank + 1 + q) % P;  // start with the neighbour so the ranks do not all hit one peer
          double2* dst = p->peer_T[h] + (size_t)p->rank * R * p->NJ + 2 * pair0;
          const double2* srcb = p->S + (size_t)h * R * p->NJ + 2 * pair0;
          Stream& cs = (q & 1) ? p->st_copy2 : p->st_copy;
          if (chunks == 1)
            VMK_TRY(be_d2d(dst, srcb, rowb * R, cs));
          else
            VMK_TRY(be_d2d_2d(dst, rowb, srcb, rowb, width, (size_t)R, cs));
        }
      } else {  // SM push kernel
        K6Args k;
        k.src = p->S;
        for (int r = 0; r < kMaxPeers; r++) k.dst.p[r] = r < P ? (void*)p->peer_T[r] : nullptr;
        k.src_hstride = (long long)R * p->NJ;
        k.src_off = 2 * pair0;
        k.dst_off = (long long)p->rank * R * p->NJ + 2 * pair0;
        k.pitch = p->NJ;
        k.rows = R;
        k.ncols = 2 * np;
        k.rank = p->rank;
        k.nranks = P;
        k.order = p->a2a_order;
        VMK_TRY(launch_k6(p, k));
      }
    }
  }
  if (P > 1) {
    VMK_TRY(be_event_record(p->ev_join, p->st_copy));
      VMK_TRY(be_event_record(p->ev_join2, p->st_copy2));
    VMK_TRY(be_stream_wait(p->st, p->ev_join2));
  }
  t.done();
  return 0;
}

// K2 + the backward transpose.  On P > 1 GPUs the rows are processed in `chunks` launches; after each one the copy
// engines move the finished rows' foreign columns (contiguous blocks of S) into the owners' V over NVLink while the
// next launch transforms the following rows.
int launch_k2(vmk_plan* p, double sign) {
  const int P = p->nranks, R = (p->N / 2) / P;
  // k2_push 1: K2 stores foreign columns straight into the peers' V (its warps stall on the NVLink store queue);
  // 0: staged in S and moved by the copy engines; 2: staged in S and moved by the SM push kernel on a second stream
  // while K2 transforms the next row chunk
  const int mode = p->k2_push >= 0 ? p->k2_push : 1;
  const int push = mode == 1;
  int chunks = (P > 1 && !push) ? (p->k2_chunks > 0 ? p->k2_chunks : 4) : 1;
  while (chunks > 1 && (R % chunks || R / chunks < 1)) chunks--;
  Timed t(p, KI_K2);
  for (int c = 0; c < chunks; c++) {
    const int nr = R / chunks, rloc0 = c * nr;
    K2Args a;
    a.T = p->T;
    a.V = p->V;
    a.S = p->S;
    for (int r = 0; r < kMaxPeers; r++) a.Vpeer.p[r] = r < P ? (void*)p->peer_V[r] : nullptr;
    a.push = push;
    a.pieces = (P == 1 && p->v_pieces) ? 1 : 0;
    a.tw = p->tw;
    a.bbcos = p->bbcos;
    a.cccos = p->cccos;
    a.ccperm = p->ccperm;
    a.aa = p->div_aa;
    a.scale = sign / (2.0 * (double)p->N * (double)p->N);
    a.NJ = p->NJ;
    a.log2NJ = p->log2NJ;
    a.nrows = nr;
    a.row0 = p->rank * R + rloc0;
    a.R = R;
    a.rloc0 = rloc0;
    a.rank = p->rank;
    a.prefetch = p->ops.cluster > 1 ? (cl_prefetch_mask(p) & 2) : p->k2_prefetch;
    const int work = (nr + p->ops.fpc - 1) / p->ops.fpc * p->ops.cluster;
    VMK_TRY(p->ops.k2(work < p->res_k2 ? work : p->res_k2, a, p->st));
    p->launches++;
    if (P > 1 && !push) {
      VMK_TRY(be_event_record(p->ev_chunk[c], p->st));
      VMK_TRY(be_stream_wait(p->st_copy, p->ev_chunk[c]));
      VMK_TRY(be_stream_wait(p->st_copy2, p->ev_chunk[c]));
      if (mode == 2) {
        K6Args k;
        k.src = p->S;
        for (int r = 0; r < kMaxPeers; r++) k.dst.p[r] = r < P ? (void*)p->peer_V[r] : nullptr;
        k.src_hstride = (long long)R * p->NJ;
        k.src_off = (long long)rloc0 * p->NJ;
        k.dst_off = (long long)(p->rank * R + rloc0) * p->NJ;
        k.pitch = p->NJ;
        k.rows = nr;
        k.ncols = p->NJ;
        k.rank = p->rank;
        k.nranks = P;
        k.order = p->a2a_order;
        VMK_TRY(launch_k6(p, k));
      } else {
        const size_t bytes = sizeof(double2) * (size_t)nr * p->NJ;
        for (int q = 0; q + 1 < P; q++) {
          const int h = (p->rank + 1 + q) % P;
          VMK_TRY(be_d2d(p->peer_V[h] + (size_t)(p->rank * R + rloc0) * p->NJ, p->S + ((size_t)h * R + rloc0) * p->NJ,
                         bytes, (q & 1) ? p->st_copy2 : p->st_copy));
        }
      }
    }
  }
  if (P > 1 && !push) {
    VMK_TRY(be_event_record(p->ev_join, p->st_copy));
      VMK_TRY(be_event_record(p->ev_join2, p->st_copy2));
    VMK_TRY(be_stream_wait(p->st, p->ev_join2));
  }
  t.done();
  return 0;
}

// hybrid solver: inverse transform along i of a row-major half spectrum `V` into the slab `out` (single GPU)
int launch_k3_rows(vmk_plan* p, const double2* V, double* out) {
  K3Args a;
  a.T = V;
  a.pieces = 0;
  a.prefetch = 0;
  a.tw = p->tw;
  a.psi = out;
  a.lo_dst = out + (size_t)(p->NJ + 1) * p->N;
  a.hi_dst = out;
  a.NJ = p->NJ;
  a.npairs = p->NJ / 2;
  const int work = rowpair_units(p, a.npairs, 1);
  Timed t(p, KI_K3);
  VMK_TRY(p->ops.k3(work < p->res_k3 ? work : p->res_k3, a, p->st));
  t.done();
  p->launches++;
  return 0;
}

int launch_k3(vmk_plan* p) {
  K3Args a;
  a.T = p->V;
  a.pieces = (p->nranks == 1 && p->v_pieces) ? 1 : 0;
  a.prefetch = p->ops.cluster > 1 ? (cl_prefetch_mask(p) & 4) : 0;
  a.tw = p->tw;
  a.psi = p->psi;
  const int prev = (p->rank + p->nranks - 1) % p->nranks, next = (p->rank + 1) % p->nranks;
  a.lo_dst = p->peer_psi[prev] + (size_t)(p->NJ + 1) * p->N;
  a.hi_dst = p->peer_psi[next];
  a.NJ = p->NJ;
  a.npairs = p->NJ / 2;
  const int work = rowpair_units(p, a.npairs, 1) * p->ops.cluster;
  Timed t(p, KI_K3);
  VMK_TRY(p->ops.k3(work < p->res_k3 ? work : p->res_k3, a, p->st));
  t.done();
  p->launches++;
  return 0;
}

// mode 0: out = r; 1..3: RK3 stages.  win/wn/out index p->w[]
int make_k4(vmk_plan* p, int win, int wn, int out, const StepParams& sp, K4Args& a) {
  a.w = p->w[win];
  a.psi = p->psi;
  a.wn = p->w[wn];
  a.out = p->w[out];
  const int prev = (p->rank + p->nranks - 1) % p->nranks, next = (p->rank + 1) % p->nranks;
  a.lo_dst = p->peer_w[out][prev] + (size_t)(p->NJ + 1) * p->N;
  a.hi_dst = p->peer_w[out][next];
  a.N = p->N;
  a.log2N = p->M;
  a.NJ = p->NJ;
  a.rows_per_cta = p->k4_rows;
  a.ahead = p->k4_ahead;
  a.rev = (p->zigzag && tri_on(p)) ? (p->zz++ & 1) : 0;
  a.aa = 1.0 / (sp.re * (sp.dx * sp.dx));  // Common.jl:149
  a.bb = 1.0 / (sp.re * (sp.dy * sp.dy));  // :150
  a.gg = 1.0 / (4.0 * sp.dx * sp.dy);      // :151
  a.hh = 1.0 / 3.0;                        // :152
  a.dt = sp.dt;
  const int cols = p->N / kK4Cols, tw = cols < kK4Threads ? cols : kK4Threads, groups = kK4Threads / tw;
  const int ctas_x = cols / tw;
  // small slabs: shorter row marches so that the grid is several waves of resident CTAs (3 per SM)
  while (a.rows_per_cta > 4 &&
         ctas_x * ((p->NJ + groups * a.rows_per_cta - 1) / (groups * a.rows_per_cta)) < p->k4_waves * 3 * p->sms)
    a.rows_per_cta /= 2;
  const int rows_per = groups * a.rows_per_cta;
  return ctas_x * ((p->NJ + rows_per - 1) / rows_per);
}

int launch_k4(vmk_plan* p, int mode, int win, int wn, int out, const StepParams& sp) {
  K4Args a;
  const int grid = make_k4(p, win, wn, out, sp, a);
  Timed t(p, KI_K4);
  int rc = 0;
  switch (mode) {
    case 0: rc = be_launch<K4Body<0>, K4Args, kK4Threads, 3>(grid, 0, a, p->st); break;
    case 1: rc = be_launch<K4Body<1>, K4Args, kK4Threads, 3>(grid, 0, a, p->st); break;
    case 2: rc = be_launch<K4Body<2>, K4Args, kK4Threads, 3>(grid, 0, a, p->st); break;
    default: rc = be_launch<K4Body<3>, K4Args, kK4Threads, 3>(grid, 0, a, p->st); break;
  }
  VMK_TRY(rc);
  t.done();
  p->launches++;
  return 0;
}

template <class Body>
int launch_k5(vmk_plan* p, const double* src, double* dst, size_t total) {
  K5Args a;
  a.src = src;
  a.dst = dst;
  a.N = p->N;
  a.NJ = p->NJ;
  size_t want = (total + kK5Threads - 1) / kK5Threads;
  const size_t cap = (size_t)p->sms * 16;
  const int grid = (int)(want < cap ? want : cap);
  VMK_TRY((be_launch<Body, K5Args, kK5Threads, 4>(grid < 1 ? 1 : grid, 0, a, p->st)));
  p->launches++;
  return 0;
}

template <class Body>
int launch_kt(vmk_plan* p, const KTArgs& a) {
  const int tiles = (a.H + kTriThreads - 1) / kTriThreads;
  const int items = a.nch * tiles, cap = p->sms * 3 * 4;
  VMK_TRY((be_launch<Body, KTArgs, kTriThreads, 3>(items < cap ? items : cap, 0, a, p->st)));
  p->launches++;
  return 0;
}
template <int PHASE>
int launch_kt_scan(vmk_plan* p, const KTArgs& a) {
  const int grid = (a.H + kTriScanSlots - 1) / kTriScanSlots;
  VMK_TRY((be_launch<KTScan<PHASE>, KTArgs, kTriScanThreads, 1>(grid, kTriScanSmem, a, p->st)));
  p->launches++;
  return 0;
}

// The same solve with the j direction in its tridiagonal form (vmk_tri.cuh): K1 (natural rows) -> chunk totals ->
// rank totals -> [one cross-rank barrier] -> carries, the rows kx < K0 by K2 -> solve in place -> K3 (natural rows).
// No transpose, on any number of GPUs.
int enqueue_poisson_tri(vmk_plan* p, const double* src, double sign) {
  const int N = p->N, H = N / 2, P = p->nranks, k0 = p->tri_k0_tab;
  {
    K1Args a;
    a.w = src;
    a.S = nullptr;
    a.Tloc = nullptr;
    a.X = p->T;
    a.tw = p->tw;
    a.NJ = p->NJ;
    a.npairs = p->NJ / 2;
    a.k_own0 = 0;
    a.k_own1 = 0;
    a.prefetch = p->k1_prefetch;
    for (int r = 0; r < kMaxPeers; r++) a.Lpeer.p[r] = r < P ? (void*)p->peer_L[r] : nullptr;
    a.k0 = k0;
    a.jbase = p->j0;
    a.nranks = P;
    a.rev = p->zigzag ? (p->zz++ & 1) : 0;
    a.prefetch = p->ops.cluster > 1 ? (cl_prefetch_mask(p) & 1) : p->k1_prefetch;
    const int work = rowpair_units(p, a.npairs, 1) * p->ops.cluster;
    Timed t(p, KI_K1);
    VMK_TRY(p->ops.k1n(work < p->res_k1 ? work : p->res_k1, a, p->st));
    t.done();
    p->launches++;
  }
  KTArgs k;
  k.X = p->T;
  k.tab = p->tri_tab;
  k.lowrow = p->tri_low;
  k.tot = p->tri_tot;
  k.cin = p->tri_cin;
  k.G = p->tri_G;
  k.L = p->tri_L;
  for (int r = 0; r < kMaxPeers; r++) k.Gpeer.p[r] = r < P ? (void*)p->peer_G[r] : nullptr;
  k.H = H;
  k.NJ = p->NJ;
  k.nch = p->tri_nch;
  k.N = N;
  k.j0 = p->j0;
  k.rank = p->rank;
  k.nranks = P;
  k.sign = sign;
  k.lowslot = p->tri_low + H;
  k.k0 = k0;
  auto next_rev = [&]() { return p->zigzag ? (p->zz++ & 1) : 0; };
  // the rows kx < K0 in L by K2's FFT pair, on the second stream beside the totals / the scan (on the plan's own
  // stream while profiling, so that its events bracket it): L is complete after K1 on one rank, after the barrier on
  // several
  auto low_rows = [&]() -> int {
    K2Args a;
    a.T = p->tri_L;
    a.V = p->tri_L;
    a.S = nullptr;
    for (int r = 0; r < kMaxPeers; r++) a.Vpeer.p[r] = nullptr;
    a.push = 0;
    a.pieces = 0;
    a.tw = p->tw;
    a.bbcos = p->bbcos;
    a.cccos = p->cccos;
    a.ccperm = p->ccperm;
    a.aa = p->div_aa;
    a.scale = sign / (2.0 * (double)N * (double)N);
    a.NJ = N;
    a.log2NJ = p->M;
    a.nrows = k0;
    a.row0 = 0;
    a.R = k0;
    a.rloc0 = 0;
    a.rank = 0;
    a.prefetch = 0;
    if (p->ops.cluster == 1) {  // (the cluster K2 keeps V = L; kt_low_body copies)
      a.Xnat = p->T;
      a.lowslot = p->tri_low + H;
      a.xnat_j0 = p->j0;
      a.xnat_nj = p->NJ;
    }
    const int work = (k0 + p->ops.fpc - 1) / p->ops.fpc * p->ops.cluster;
    const int grid = work < p->res_k2 ? work : p->res_k2;
    p->launches++;
    if (p->profiling) {
      Timed t(p, KI_K2);
      VMK_TRY(p->ops.k2(grid, a, p->st));
      t.done();
      return be_event_record(p->ev_join, p->st);
    }
    VMK_TRY(be_event_record(p->ev_chunk[0], p->st));
    VMK_TRY(be_stream_wait(p->st_copy, p->ev_chunk[0]));
    VMK_TRY(p->ops.k2(grid, a, p->st_copy));
    return be_event_record(p->ev_join, p->st_copy);
  };
  auto timed = [&](int which, auto&& launch) -> int {
    Timed t(p, which);
    VMK_TRY(launch());
    t.done();
    return 0;
  };
  if (P == 1) {
    VMK_TRY(low_rows());
    k.rev = next_rev();
    VMK_TRY(timed(KI_KT1, [&] { return launch_kt<KTTotals>(p, k); }));
    VMK_TRY(timed(KI_KT2, [&] { return launch_kt_scan<2>(p, k); }));
  } else {
    k.rev = next_rev();
    VMK_TRY(timed(KI_KT1, [&] { return launch_kt<KTTotals>(p, k); }));
    VMK_TRY(timed(KI_KT2, [&] { return launch_kt_scan<0>(p, k); }));
    VMK_TRY(cross_rank_barrier(p));  // every rank's totals and low-row columns have landed
    VMK_TRY(low_rows());
    VMK_TRY(timed(KI_KT2, [&] { return launch_kt_scan<1>(p, k); }));
  }
  k.rev = next_rev();
  VMK_TRY(timed(KI_KT3, [&] { return launch_kt<KTSolve>(p, k); }));
  VMK_TRY(be_stream_wait(p->st, p->ev_join));  // the rows kx < K0 are solved and (one SM per row: K2 stored them
                                               // there itself) back in their slots; cluster sizes: copy them
  if (p->ops.cluster > 1) {
    VMK_TRY(timed(KI_K2, [&] {
      const int want = (k0 * p->NJ + kTriThreads - 1) / kTriThreads, cap = p->sms * 8;
      p->launches++;
      return be_launch<KTLow, KTArgs, kTriThreads, 3>(want < cap ? want : cap, 0, k, p->st);
    }));
  }
  {
    K3Args a;
    a.rev = next_rev();
    a.T = p->T;
    a.pieces = 0;
    a.prefetch = 0;
    a.tw = p->tw;
    a.psi = p->psi;
    const int prev = (p->rank + P - 1) % P, next = (p->rank + 1) % P;
    a.lo_dst = p->peer_psi[prev] + (size_t)(p->NJ + 1) * N;
    a.hi_dst = p->peer_psi[next];
    a.NJ = p->NJ;
    a.npairs = p->NJ / 2;
    a.prefetch = p->ops.cluster > 1 ? (cl_prefetch_mask(p) & 4) : 0;
    const int work = rowpair_units(p, a.npairs, 1) * p->ops.cluster;
    Timed t3(p, KI_K3);
    VMK_TRY(p->ops.k3n(work < p->res_k3 ? work : p->res_k3, a, p->st));
    t3.done();
    p->launches++;
  }
  return 0;
}

// The recurrence form with the recurrences INSIDE K1 and K3 (one GPU; vmk_tri.cuh "fused form"): K1 (forward
// recurrence per unit, block totals) -> scan over the blocks (beside it: the rows kx < K0 by K2, stored straight into
// their slots) -> K3 (backward recurrence, scaling, eps correction).  The spectrum crosses HBM twice instead of five times.
int enqueue_poisson_fused(vmk_plan* p, const double* src, double sign) {
  const int N = p->N, H = N / 2, k0 = p->tri_k0_tab;
  const int units = p->fz_grid * p->ops.fpc;
  {
    K1Args a;
    a.w = src;
    a.S = nullptr;
    a.Tloc = nullptr;
    a.X = p->T;
    a.tw = p->tw;
    a.NJ = p->NJ;
    a.npairs = p->NJ / 2;
    a.k_own0 = 0;
    a.k_own1 = 0;
    a.prefetch = p->k1_prefetch;
    for (int r = 0; r < kMaxPeers; r++) a.Lpeer.p[r] = r < 1 ? (void*)p->peer_L[r] : nullptr;
    a.k0 = k0;
    a.jbase = 0;
    a.nranks = 1;
    a.rr = p->tri_rr;
    a.tot = p->tri_tot;
    Timed t(p, KI_K1);
    VMK_TRY(p->ops.k1f(p->fz_grid, a, p->st));
    t.done();
    p->launches++;
  }
  KTArgs k;
  k.X = p->T;
  k.tab = p->tri_tab;
  k.lowrow = p->tri_low;
  k.tot = p->tri_tot;
  k.cin = p->tri_cin;
  k.G = p->tri_G;
  k.L = p->tri_L;
  for (int r = 0; r < kMaxPeers; r++) k.Gpeer.p[r] = nullptr;
  k.H = H;
  k.NJ = p->NJ;
  k.nch = p->tri_nch;
  k.N = N;
  k.j0 = 0;
  k.rank = 0;
  k.nranks = 1;
  k.sign = sign;
  k.lowslot = p->tri_low + H;
  k.k0 = k0;
  k.ftab = p->tri_ftab;
  k.units = units;
  k.npairs = p->NJ / 2;
  {  // the rows kx < K0 in L by K2's FFT pair, on the second stream beside the scan
    K2Args a;
    a.T = p->tri_L;
    a.V = p->tri_L;
    a.S = nullptr;
    for (int r = 0; r < kMaxPeers; r++) a.Vpeer.p[r] = nullptr;
    a.push = 0;
    a.pieces = 0;
    a.tw = p->tw;
    a.bbcos = p->bbcos;
    a.cccos = p->cccos;
    a.ccperm = p->ccperm;
    a.aa = p->div_aa;
    a.scale = sign / (2.0 * (double)N * (double)N);
    a.NJ = N;
    a.log2NJ = p->M;
    a.nrows = k0;
    a.row0 = 0;
    a.R = k0;
    a.rloc0 = 0;
    a.rank = 0;
    a.prefetch = 0;
    a.Xnat = p->T;
    a.lowslot = p->tri_low + H;
    a.xnat_j0 = 0;
    a.xnat_nj = N;
    const int work = (k0 + p->ops.fpc - 1) / p->ops.fpc * p->ops.cluster;
    const int grid = work < p->res_k2 ? work : p->res_k2;
    p->launches++;
    if (p->profiling) {
      Timed t(p, KI_K2);
      VMK_TRY(p->ops.k2(grid, a, p->st));
      t.done();
      VMK_TRY(be_event_record(p->ev_join, p->st));
    } else {
      VMK_TRY(be_event_record(p->ev_chunk[0], p->st));
      VMK_TRY(be_stream_wait(p->st_copy, p->ev_chunk[0]));
      VMK_TRY(p->ops.k2(grid, a, p->st_copy));
      VMK_TRY(be_event_record(p->ev_join, p->st_copy));
    }
  }
  {
    Timed t(p, KI_KT2);
    const int grid = (H + kTriScanSlots - 1) / kTriScanSlots;
    VMK_TRY((be_launch<KTScanF, KTArgs, kTriScanThreads, 1>(grid, kTriScanFSmem, k, p->st)));
    t.done();
    p->launches++;
  }
  VMK_TRY(be_stream_wait(p->st, p->ev_join));  // the rows kx < K0 are solved and back in their slots
  {
    K3Args a;
    a.T = p->T;
    a.pieces = 0;
    a.prefetch = 0;
    a.tw = p->tw;
    a.psi = p->psi;
    a.lo_dst = p->peer_psi[0] + (size_t)(p->NJ + 1) * N;
    a.hi_dst = p->peer_psi[0];
    a.NJ = p->NJ;
    a.npairs = p->NJ / 2;
    a.rr = p->tri_rr;
    a.cin = p->tri_cin;
    a.kc = p->fz_kc * sign;
    Timed t3(p, KI_K3);
    VMK_TRY(p->ops.k3f(p->fz_grid, a, p->st));
    t3.done();
    p->launches++;
  }
  return 0;
}

// psi = solve(sign * src): K1 -> K2 -> K3.  Common.jl:115-123
int enqueue_poisson(vmk_plan* p, const double* src, double sign) {
  if (fz_on(p)) return enqueue_poisson_fused(p, src, sign);
  if (tri_on(p)) return enqueue_poisson_tri(p, src, sign);
  VMK_TRY(launch_k1(p, src));
  VMK_TRY(cross_rank_barrier(p));  // every rank's spectrum is written before any rank transforms along j
  VMK_TRY(launch_k2(p, sign));
  VMK_TRY(cross_rank_barrier(p));  // ... and transformed back before the owners read it
  VMK_TRY(launch_k3(p));
  return 0;
}

// One SSP-RK3 step, vm.jl:26-76.  Ping-pong: S1 wn -> wtA; S2 (wn, wtA) -> wtB; S3 (wn, wtB) -> wn
// (in place is safe in S3: wn is only read point-wise there).
int enqueue_step(vmk_plan* p, const StepParams& sp) {
  VMK_TRY(enqueue_poisson(p, p->w[0], -1.0));
  VMK_TRY(cross_rank_barrier(p));  // neighbours' psi halo rows have landed
  VMK_TRY(launch_k4(p, 1, 0, 0, 1, sp));
  VMK_TRY(enqueue_poisson(p, p->w[1], -1.0));
  VMK_TRY(cross_rank_barrier(p));
  VMK_TRY(launch_k4(p, 2, 1, 0, 2, sp));
  VMK_TRY(enqueue_poisson(p, p->w[2], -1.0));
  VMK_TRY(cross_rank_barrier(p));
  VMK_TRY(launch_k4(p, 3, 2, 0, 0, sp));
  return 0;
}

// Small grids on one GPU: `nsteps` steps as one launch of one cluster (ks_body).  Same buffers, tables and per-body
// work decomposition as enqueue_step.
bool small_fused_on(const vmk_plan* p) {
  return p->nranks == 1 && p->ops.ks && p->fuse_small != 0 && p->v_pieces && !p->profiling && !p->barrier_fn &&
         p->ks_q > 0;
}

int enqueue_small(vmk_plan* p, const StepParams& sp, int64_t nsteps) {
  KSArgs a;
  const int N = p->N;
  const int src_of[3] = {0, 1, 2}, out_of[3] = {1, 2, 0};
  for (int s = 0; s < 3; s++) {
    K1Args& k = a.k1[s];
    k.w = p->w[src_of[s]];
    k.S = nullptr;
    k.Tloc = p->T;
    k.tw = p->tw;
    k.NJ = p->NJ;
    k.npairs = p->NJ / 2;
    k.k_own0 = 0;
    k.k_own1 = N / 2;
    k.prefetch = 0;
    a.k4_grid[s] = make_k4(p, src_of[s], 0, out_of[s], sp, a.k4[s]);
  }
  K2Args& k2 = a.k2;
  k2.T = p->T;
  k2.V = p->V;
  k2.S = nullptr;
  for (int r = 0; r < kMaxPeers; r++) k2.Vpeer.p[r] = nullptr;
  k2.push = 0;
  k2.pieces = 1;
  k2.tw = p->tw;
  k2.bbcos = p->bbcos;
  k2.cccos = p->cccos;
  k2.ccperm = p->ccperm;
  k2.aa = p->div_aa;
  k2.scale = -1.0 / (2.0 * (double)N * (double)N);
  k2.NJ = p->NJ;
  k2.log2NJ = p->log2NJ;
  k2.nrows = N / 2;
  k2.row0 = 0;
  k2.R = N / 2;
  k2.rloc0 = 0;
  k2.rank = 0;
  k2.prefetch = 0;
  K3Args& k3 = a.k3;
  k3.T = p->V;
  k3.pieces = 1;
  k3.prefetch = 0;
  k3.tw = p->tw;
  k3.psi = p->psi;
  k3.lo_dst = p->psi + (size_t)(p->NJ + 1) * N;
  k3.hi_dst = p->psi;
  k3.NJ = p->NJ;
  k3.npairs = p->NJ / 2;
  a.nsteps = nsteps;
  VMK_TRY(p->ops.ks(p->ks_q, a, p->st));
  p->launches++;
  return 0;
}

// ---- lid-driven cavity (18_NS2D_Lid_Driven_Cavity/lid_driven_cavity.jl) ----------------------------------------------
// Divisor of fps_sine on the odd extension (plan size N = 2 nx): mode k of the 2nx-periodic grid is the DST frequency
// pi k / nx, so d(k, l) = (2/dx^2)(cos(pi k/nx) - 1) + (2/dy^2)(cos(pi l/ny) - 1) exactly as lid_driven_cavity.jl:66-71
// writes it (the two bracketed terms are tabulated and added: no aa + bb cos cancellation for the low modes).
int ensure_divisor_cavity(vmk_plan* p, double dx, double dy) {
  if (p->div_valid && p->div_kind == 1 && p->div_dx == dx && p->div_dy == dy) return 0;
  const int N = p->N, n = N / 2;
  std::vector<double> b(N), c(N), cp(N);
  for (int k = 0; k < N; k++) {
    const int km = k <= n ? k : N - k;  // cos is even about k = n: evaluate the mirror index, bitwise symmetric tables
    const double ck = cos(M_PI * (double)km / (double)n) - 1.;
    b[k] = (2. / (dx * dx)) * ck;
    c[k] = (2. / (dy * dy)) * ck;
  }
  p->ops.fill_ccperm(c.data(), cp.data());
  VMK_TRY(be_sync(p->st));
  VMK_TRY(be_h2d(p->bbcos, b.data(), sizeof(double) * N, p->st));
  VMK_TRY(be_h2d(p->cccos, c.data(), sizeof(double) * N, p->st));
  VMK_TRY(be_h2d(p->ccperm, cp.data(), sizeof(double) * N, p->st));
  VMK_TRY(be_sync(p->st));
  p->div_aa = 0.0;
  p->div_dx = dx;
  p->div_dy = dy;
  p->div_eps = 0.0;
  p->div_kind = 1;
  p->div_valid = true;
  drop_graphs(p);
  return 0;
}

template <class Body>
int launch_kc(vmk_plan* p, const KCArgs& a, size_t items, size_t smem = 0, int grid_override = 0) {
  size_t want = (items + kKCThreads - 1) / kKCThreads;
  const size_t cap = (size_t)p->sms * 16;
  int grid = (int)(want < cap ? want : cap);
  if (grid < 1) grid = 1;
  if (grid_override) grid = grid_override;
  VMK_TRY((be_launch<Body, KCArgs, kKCThreads, 4>(grid, smem, a, p->st)));
  p->launches++;
  return 0;
}

// sn[2:nx, 2:ny] = fps_sine(-w)  (lid_driven_cavity.jl:87,100,113): odd extension -> periodic solve -> interior
int cavity_poisson(vmk_plan* p, KCArgs a, const double* w) {
  const size_t N = (size_t)p->N, m = N / 2 - 1;
  a.w = w;
  a.slab = p->w[1];
  VMK_TRY(launch_kc<KCExtend>(p, a, N * N));
  VMK_TRY(enqueue_poisson(p, p->w[1], -1.0));
  a.slab = p->psi;
  a.out = p->cs;
  VMK_TRY(launch_kc<KCExtract>(p, a, m * m));
  return 0;
}

// ---- hybrid RK3 / Crank-Nicolson solver (20_NS2D_Hybrid_Solver/hybrid.jl) -------------------------------------------
int ensure_hybrid(vmk_plan* p, double dx) {
  if (p->nranks != 1) return fail(VMK_EARG, "the hybrid solver runs on single-GPU plans");
  if (!p->ops.kh) return fail(VMK_ESIZE, "the hybrid solver supports grids up to 8192^2");
  const size_t spec = sizeof(double2) * (size_t)(p->N / 2) * p->N;
  if (!p->hW) {
    VMK_TRY(p->ops.kh_configure(&p->res_kh));
    VMK_TRY(dev_alloc(p, (void**)&p->hW, spec));
    VMK_TRY(dev_alloc(p, (void**)&p->hJ, spec));
    VMK_TRY(dev_alloc(p, (void**)&p->hVs, spec));
    VMK_TRY(dev_alloc(p, (void**)&p->hksq, sizeof(double) * p->N));
    VMK_TRY(dev_alloc(p, (void**)&p->hksqperm, sizeof(double) * p->N));
    p->hyb_dx = 0;
  }
  if (p->hyb_dx != dx) {
    // wavespace, Common.jl:184-204: hx = 2 pi/(nx dx), kx[i] = hx (i-1), kx[i+nx/2] = hx (i-nx/2-1), kx[1] = eps, ky = kx
    const int n = p->N;
    std::vector<double> k(n), kp(n);
    const double hx = 2.0 * M_PI / ((double)n * dx);
    for (int i = 1; i <= n / 2; i++) {
      k[i - 1] = hx * ((double)i - 1.0);
      k[i + n / 2 - 1] = hx * (double)(i - n / 2 - 1);
    }
    k[0] = 1.e-6;
    for (int i = 0; i < n; i++) k[i] = k[i] * k[i];
    p->ops.fill_ksqperm(k.data(), kp.data());
    VMK_TRY(be_sync(p->st));
    VMK_TRY(be_h2d(p->hksq, k.data(), sizeof(double) * n, p->st));
    VMK_TRY(be_h2d(p->hksqperm, kp.data(), sizeof(double) * n, p->st));
    VMK_TRY(be_sync(p->st));
    p->hyb_dx = dx;
  }
  return 0;
}

// stage 0: wf = fft(w0); 1..3: the RK3/CN stages (alpha, gamma, rho of hybrid.jl:29-31)
int launch_kh(vmk_plan* p, int stage, double dt, double re) {
  static const double alpha[4] = {0.0, 8. / 15., 2. / 15., 1. / 3.};
  static const double gamma[4] = {0.0, 8. / 15., 5. / 12., 3. / 4.};
  static const double rho[4] = {0.0, 0.0, -17. / 60., -5. / 12.};
  KHArgs a;
  a.T = p->T;
  a.W = p->hW;
  a.J = p->hJ;
  a.Vw = p->V;
  a.Vs = p->hVs;
  a.tw = p->tw;
  a.ksq = p->hksq;
  a.ksqperm = p->hksqperm;
  a.zfac = .5 * dt / re;
  a.alpha = alpha[stage];
  a.gdt = stage ? gamma[stage] * dt : 1.0;
  a.rdt = rho[stage] * dt;
  a.scale = 1.0 / (2.0 * (double)p->N * (double)p->N);
  a.stage = stage;
  a.nrows = p->N / 2;
  const int work = (a.nrows + p->ops.fpc - 1) / p->ops.fpc;
  Timed t(p, KI_K2);
  VMK_TRY(p->ops.kh(work < p->res_kh ? work : p->res_kh, a, p->st));
  t.done();
  p->launches++;
  return 0;
}

// ut = real(ifft(wf)) with the periodic duplicates (hybrid.jl:73-78): (N+1) x (N+1) on the host
int hybrid_field(vmk_plan* p, double* ut) {
  const size_t N = (size_t)p->N, n1 = N + 1;
  VMK_TRY(launch_k3_rows(p, p->V, p->w[1]));
  VMK_TRY(be_d2h_2d(ut, sizeof(double) * n1, p->w[1] + N, sizeof(double) * N, sizeof(double) * N, N, p->st));
  VMK_TRY(be_sync(p->st));
  for (size_t j = 0; j < N; j++) ut[N + j * n1] = ut[j * n1];
  memcpy(ut + N * n1, ut, sizeof(double) * n1);
  return 0;
}

// ---- pseudo-spectral solver, 2/3 rule (22_NS2D_PseudoSpectral_23_Rule/pseudospectral_23_rule.jl) -------------------
int ensure_ps23(vmk_plan* p, double dx) {
  VMK_TRY(ensure_hybrid(p, dx));  // wf, jf, k2 tables: the same state as the hybrid solver's
  if (!p->ops.kp) return fail(VMK_ESIZE, "the pseudo-spectral solver supports grids up to 8192^2");
  const int n = p->N;
  const size_t spec = sizeof(double2) * (size_t)(n / 2) * n;
  if (!p->ptab) {  // (ptab is allocated last: a failed attempt is resumed, not repeated)
    VMK_TRY(p->ops.kp_configure(&p->res_kp));
    auto once = [&](void** ptr, size_t bytes) -> int { return *ptr ? 0 : dev_alloc(p, ptr, bytes); };
    VMK_TRY(once((void**)&p->pV[0], spec));
    VMK_TRY(once((void**)&p->pV[1], spec));
    VMK_TRY(once((void**)&p->pA0, sizeof(double2) * n));
    VMK_TRY(once((void**)&p->ptab, sizeof(double) * 8 * n));
    p->ps_dx = 0;
  }
  if (p->ps_dx != dx) {
    // jacobian(), pseudospectral_23_rule.jl:96-108: hx = 2 pi/(nx dx), kx[1] = eps, ky = kx; :124-133: the zeroed band
    std::vector<double> k(n), m(n), tab(8 * (size_t)n);
    const double hx = 2.0 * M_PI / ((double)n * dx);
    for (int i = 1; i <= n / 2; i++) {
      k[i - 1] = hx * ((double)i - 1.0);
      k[i + n / 2 - 1] = hx * (double)(i - n / 2 - 1);
    }
    k[0] = 1.e-6;
    const int nxe = (int)floor(2.0 * (double)n / 3.0), half = nxe / 2;
    for (int i = 0; i < n; i++) m[i] = (i >= half && i < n - half) ? 0.0 : 1.0;  // 1-based half+1 .. n-half zeroed
    if (m[n / 2] != 0.0) return fail(VMK_ESIZE, "2/3 rule: the Nyquist mode is expected inside the truncated band");
    double *mp = tab.data(), *mm = mp + n, *cc = mm + n, *dd = cc + n;
    for (int i = 0; i < n; i++) {
      const int mi = (n - i) % n;
      mp[i] = m[i];
      mm[i] = m[mi];
      cc[i] = k[i] * m[i] * .5;
      dd[i] = -k[mi] * m[mi] * .5;
    }
    for (int q = 0; q < 4; q++) p->ops.fill_ksqperm(tab.data() + (size_t)q * n, tab.data() + (size_t)(4 + q) * n);
    VMK_TRY(be_sync(p->st));
    VMK_TRY(be_h2d(p->ptab, tab.data(), sizeof(double) * 8 * n, p->st));
    VMK_TRY(be_sync(p->st));
    p->ps_dx = dx;
  }
  return 0;
}

// stage 0: wf = fft(w0); 1..3: the RK3/CN stages (alpha, gamma, rho of pseudospectral_23_rule.jl:30-32); 4: V = ifft_j(wf)
int launch_kp(vmk_plan* p, int stage, double dt, double re) {
  static const double alpha[5] = {0.0, 8. / 15., 2. / 15., 1. / 3., 0.0};
  static const double gamma[5] = {0.0, 8. / 15., 5. / 12., 3. / 4., 0.0};
  static const double rho[5] = {0.0, 0.0, -17. / 60., -5. / 12., 0.0};
  const size_t n = (size_t)p->N;
  KPArgs a;
  a.T = p->T;
  a.W = p->hW;
  a.J = p->hJ;
  a.A0 = p->pA0;
  a.V[0] = stage == 4 ? p->T : p->V;
  a.V[1] = p->hVs;
  a.V[2] = p->pV[0];
  a.V[3] = p->pV[1];
  a.tw = p->tw;
  a.ksq = p->hksq;
  a.ksqperm = p->hksqperm;
  a.mp = p->ptab;
  a.mm = p->ptab + n;
  a.cc = p->ptab + 2 * n;
  a.dd = p->ptab + 3 * n;
  a.mpperm = p->ptab + 4 * n;
  a.mmperm = p->ptab + 5 * n;
  a.ccperm = p->ptab + 6 * n;
  a.ddperm = p->ptab + 7 * n;
  a.zfac = .5 * dt / re;
  a.alpha = alpha[stage];
  a.gdt = (stage >= 1 && stage <= 3) ? gamma[stage] * dt : 1.0;
  a.rdt = rho[stage] * dt;
  a.scale = 1.0 / (2.0 * (double)p->N * (double)p->N);
  a.stage = stage;
  a.nrows = p->N / 2;
  const int work = (a.nrows + p->ops.fpc - 1) / p->ops.fpc;
  Timed t(p, KI_K2);
  VMK_TRY(p->ops.kp(work < p->res_kp ? work : p->res_kp, a, p->st));
  t.done();
  p->launches++;
  return 0;
}

// jacp = j1 j2 - j3 j4 over the interior rows of four slabs, in place over the first (pseudospectral_23_rule.jl:138-141)
int launch_kp_product(vmk_plan* p, double* q1, const double* q2, const double* q3, const double* q4) {
  KPProdArgs a;
  const size_t N = (size_t)p->N;
  a.q1 = q1 + N;
  a.q2 = q2 + N;
  a.q3 = q3 + N;
  a.q4 = q4 + N;
  a.n = N * (size_t)p->NJ;
  size_t want = (a.n / 2 + kK5Threads - 1) / kK5Threads;
  const size_t cap = (size_t)p->sms * 16;
  const int grid = (int)(want < cap ? want : cap);
  Timed t(p, KI_K4);
  VMK_TRY((be_launch<KPProduct, KPProdArgs, kK5Threads, 4>(grid < 1 ? 1 : grid, 0, a, p->st)));
  t.done();
  p->launches++;
  return 0;
}

// ut = real(ifft(wf)) with the periodic duplicates (pseudospectral_23_rule.jl:71-77): (N+1) x (N+1) on the host.
// Scratch: T (free once the last stage's KP has consumed it) and w[1]; the four derivative spectra V[0..3] that the next
// step's first jacobian reads stay untouched.
int ps23_field(vmk_plan* p, double* ut, double dt, double re) {
  const size_t N = (size_t)p->N, n1 = N + 1;
  VMK_TRY(launch_kp(p, 4, dt, re));
  VMK_TRY(launch_k3_rows(p, p->T, p->w[1]));
  VMK_TRY(be_d2h_2d(ut, sizeof(double) * n1, p->w[1] + N, sizeof(double) * N, sizeof(double) * N, N, p->st));
  VMK_TRY(be_sync(p->st));
  for (size_t j = 0; j < N; j++) ut[N + j * n1] = ut[j * n1];
  memcpy(ut + N * n1, ut, sizeof(double) * n1);
  return 0;
}

// ---- pseudo-spectral solver, 3/2 rule (21_NS2D_PseudoSpectral_32_Rule/pseudospectral_32_rule.jl; vmk_pseudo32.cuh) ----
int ensure_kx(vmk_plan* p) {
  if (!p->ops.kx) return fail(VMK_ESIZE, "the 3/2-rule solver supports grids of 64^2 .. 8192^2");
  if (!p->res_kx) VMK_TRY(p->ops.kx_configure(&p->res_kx));
  return 0;
}

int ensure_ps32(vmk_plan* p, double dx) {
  if (p->nranks != 1) return fail(VMK_EARG, "the pseudo-spectral solver runs on single-GPU plans");
  if (p->N < 64) return fail(VMK_ESIZE, "the 3/2-rule solver supports grids of 64^2 .. 8192^2");
  VMK_TRY(ensure_kx(p));
  const size_t L = (size_t)p->N / 2, W = 2 * L + 1, c2 = sizeof(double2);
  if (!p->qtab) {  // (qtab is allocated last: a failed attempt is resumed, not repeated)
    if (!p->child) VMK_TRY(vmk_plan_create((int64_t)L, (int64_t)L, &p->child));
    VMK_TRY(ensure_kx(p->child));
    auto once = [&](void** ptr, size_t bytes) -> int { return *ptr ? 0 : dev_alloc(p, ptr, bytes); };
    VMK_TRY(once((void**)&p->qS, c2 * (L + 1) * W));
    VMK_TRY(once((void**)&p->qJ, c2 * (L + 1) * W));
    VMK_TRY(once((void**)&p->qY, c2 * 4 * (L + 1) * 3 * L));
    VMK_TRY(once((void**)&p->qVF, c2 * 36 * (L / 2) * L));
    VMK_TRY(once((void**)&p->qT9, c2 * 9 * (L / 2) * L));
    VMK_TRY(once((void**)&p->qPi, c2 * (L + 1) * 3 * L));
    VMK_TRY(once((void**)&p->qF, sizeof(double) * 36 * (L + 2) * L));
    VMK_TRY(once((void**)&p->qtw, c2 * 3 * L));
    VMK_TRY(once((void**)&p->qtab, sizeof(double) * 5 * W));
    p->q_dx = 0;
  }
  if (p->q_dx != dx) {
    // jacobian(), pseudospectral_32_rule.jl:96-108: hx = 2 pi/(nx dx), kx[1] = eps, ky = kx; :137-155: the retained modes
    const int l = (int)L, n = p->N, w = (int)W, m3 = 3 * l;
    std::vector<double> kap(W), tab(5 * W);
    const double hx = 2.0 * M_PI / ((double)n * dx);
    for (int c = 0; c < w; c++) kap[c] = hx * (double)(c - l);
    kap[l] = 1.e-6;
    double *ksq = tab.data(), *mp = ksq + W, *mm = mp + W, *cc = mm + W, *dd = cc + W;
    for (int c = 0; c < w; c++) mp[c] = c < 2 * l ? 1.0 : 0.0;
    for (int c = 0; c < w; c++) {
      ksq[c] = kap[c] * kap[c];
      mm[c] = mp[2 * l - c];
      cc[c] = kap[c] * mp[c] * .5;
      dd[c] = -kap[2 * l - c] * mm[c] * .5;
    }
    std::vector<double2> tw(m3);
    const long double tau = 6.283185307179586476925286766559005768L;
    for (int e = 0; e < m3; e++) {
      const long double ang = tau * (long double)e / (long double)m3;
      tw[e].x = (double)cosl(ang);
      tw[e].y = (double)(-sinl(ang));
    }
    VMK_TRY(be_sync(p->st));
    VMK_TRY(be_sync(p->child->st));
    VMK_TRY(be_h2d(p->qtab, tab.data(), sizeof(double) * 5 * W, p->st));
    VMK_TRY(be_h2d(p->qtw, tw.data(), c2 * m3, p->st));
    VMK_TRY(be_sync(p->st));
    p->q_dx = dx;
  }
  return 0;
}

// batched row FFT on plan q's stream (natural order, in place allowed)
int launch_kx(vmk_plan* q, const double2* in, double2* out, int nrows, int inverse) {
  KXArgs a;
  a.in = in;
  a.out = out;
  a.tw = q->tw;
  a.nrows = nrows;
  a.inverse = inverse;
  const int work = (nrows + q->ops.fpc - 1) / q->ops.fpc;
  Timed t(q, KI_K2);
  VMK_TRY(q->ops.kx(work < q->res_kx ? work : q->res_kx, a, q->st));
  t.done();
  q->launches++;
  return 0;
}

template <class Body>
int launch_p32(vmk_plan* q, const P32Args& a, size_t items) {
  size_t want = (items + kK5Threads - 1) / kK5Threads;
  const size_t cap = (size_t)q->sms * 16;
  const int grid = (int)(want < cap ? want : cap);
  Timed t(q, KI_K4);
  VMK_TRY((be_launch<Body, P32Args, kK5Threads, 4>(grid < 1 ? 1 : grid, 0, a, q->st)));
  t.done();
  q->launches++;
  return 0;
}

P32Args p32_args(vmk_plan* p, int stage, double dt, double re) {
  static const double alpha[4] = {0.0, 8. / 15., 2. / 15., 1. / 3.};
  static const double gamma[4] = {0.0, 8. / 15., 5. / 12., 3. / 4.};
  static const double rho[4] = {0.0, 0.0, -17. / 60., -5. / 12.};
  const size_t W = (size_t)p->N + 1;
  P32Args a;
  a.L = p->N / 2;
  a.twM = p->qtw;
  a.ksq = p->qtab;
  a.mp = p->qtab + W;
  a.mm = p->qtab + 2 * W;
  a.cc = p->qtab + 3 * W;
  a.dd = p->qtab + 4 * W;
  a.S = p->qS;
  a.J = p->qJ;
  a.Y = p->qY;
  a.VF = p->qVF;
  a.T9 = p->qT9;
  a.Pi = p->qPi;
  a.Xn = p->V;
  a.Un = p->T;
  a.zfac = .5 * dt / re;
  a.alpha = alpha[stage];
  a.gdt = gamma[stage] * dt;
  a.rdt = rho[stage] * dt;
  a.scale = 1.0 / (2.0 * (double)p->N * (double)p->N);
  a.stage = stage;
  return a;
}

// one RK3 stage on the child plan's stream.  The 9 sub-grids of a field are one batch for K3 / K1: 9L rows of L points
// (the kernels take the row count and pitch as arguments; only the plan-level launchers assume a square slab)
int ps32_stage(vmk_plan* p, int stage, double dt, double re) {
  vmk_plan* ch = p->child;
  const size_t L = (size_t)p->N / 2, rows = 9 * L, slab = (rows + 2) * L, blk = (L / 2) * rows;
  const P32Args a = p32_args(p, stage, dt, re);
  if (p->ps32_fuse == 2) {                                                // spectra, fold along i and ifft along j in one kernel
    if (!ch->res_kxf) VMK_TRY(ch->ops.kxf_configure(&ch->res_kxf));
    KXSArgs k;
    k.a = a;
    k.tw = ch->tw;
    k.nrows = (int)(36 * (L / 2));
    const int units = (k.nrows + ch->ops.fpc - 1) / ch->ops.fpc;
    Timed t(ch, KI_K2);
    VMK_TRY(ch->ops.kxf(units < ch->res_kxf ? units : ch->res_kxf, k, ch->st));
    t.done();
    ch->launches++;
  } else if (p->ps32_fuse) {                                              // both of the following in one kernel
    if (!ch->res_kxs) VMK_TRY(ch->ops.kxs_configure(&ch->res_kxs));
    KXSArgs k;
    k.a = a;
    k.tw = ch->tw;
    k.nrows = (int)(12 * (L + 1));
    const int units = (k.nrows + ch->ops.fpc - 1) / ch->ops.fpc;
    Timed t(ch, KI_K2);
    VMK_TRY(ch->ops.kxs(units < ch->res_kxs ? units : ch->res_kxs, k, ch->st));
    t.done();
    ch->launches++;
  } else {
    VMK_TRY(launch_p32<P32Spectra>(ch, a, (L + 1) * L));                  // i k wf [/ k2], folded along j   :113-155
    VMK_TRY(launch_kx(ch, p->qY, p->qY, (int)(12 * (L + 1)), 1));         // ifft along j (3 sub-rows each)  :157-160
  }
  if (p->ps32_fuse != 2) VMK_TRY(launch_p32<P32Fold>(ch, a, 4 * 3 * (L / 2) * L));  // folded along i
  const int npairs = (int)(rows / 2);
  const int work = rowpair_units(ch, npairs, 1);
  for (int q = 0; q < 4; q++) {                                            // ifft along i: 9 sub-grids per launch
    K3Args k;
    double* out = p->qF + (size_t)q * slab;
    k.T = p->qVF + (size_t)q * blk;
    k.pieces = 0;
    k.prefetch = 0;
    k.tw = ch->tw;
    k.psi = out;
    k.lo_dst = out + (rows + 1) * L;  // the halo rows of the batch are written and never read
    k.hi_dst = out;
    k.NJ = (int)rows;
    k.npairs = npairs;
    Timed t(ch, KI_K3);
    VMK_TRY(ch->ops.k3(work < ch->res_k3 ? work : ch->res_k3, k, ch->st));
    t.done();
    ch->launches++;
  }
  {
    KPProdArgs k;                                                          // j1 j2 - j3 j4                   :163-166
    k.q1 = p->qF + L;
    k.q2 = p->qF + slab + L;
    k.q3 = p->qF + 2 * slab + L;
    k.q4 = p->qF + 3 * slab + L;
    k.n = rows * L;
    size_t want = (k.n / 2 + kK5Threads - 1) / kK5Threads;
    const size_t cap = (size_t)ch->sms * 16;
    const int grid = (int)(want < cap ? want : cap);
    Timed t(ch, KI_K4);
    VMK_TRY((be_launch<KPProduct, KPProdArgs, kK5Threads, 4>(grid < 1 ? 1 : grid, 0, k, ch->st)));
    t.done();
    ch->launches++;
  }
  {
    K1Args k;                                                              // fft along i                     :168
    k.w = p->qF;
    k.S = nullptr;
    k.Tloc = p->qT9;
    k.tw = ch->tw;
    k.NJ = (int)rows;
    k.npairs = npairs;
    k.k_own0 = 0;
    k.k_own1 = (int)(L / 2);
    k.prefetch = 0;
    Timed t(ch, KI_K1);
    VMK_TRY(ch->ops.k1(work < ch->res_k1 ? work : ch->res_k1, k, ch->st));
    t.done();
    ch->launches++;
  }
  VMK_TRY(launch_p32<P32Unfold>(ch, a, (L + 1) * 3 * L));                 // unfolded along i, x nx ny/(nxe nye)  :176
  VMK_TRY(launch_kx(ch, p->qPi, p->qPi, (int)(3 * (L + 1)), 0));          // fft along j
  VMK_TRY(launch_p32<P32Update>(ch, a, (L + 1) * (2 * L + 1)));           // unfolded along j + the RK3/CN update  :41-66
  return 0;
}

// ut = real(ifft(wnf)) on the nx x ny grid with the periodic duplicates (:71-76), on the parent plan's stream
int ps32_field(vmk_plan* p, double* ut, double dt, double re) {
  const size_t N = (size_t)p->N, n1 = N + 1;
  VMK_TRY(be_sync(p->child->st));
  const P32Args a = p32_args(p, 0, dt, re);
  VMK_TRY(launch_p32<P32Final>(p, a, (N / 2) * N));
  VMK_TRY(launch_kx(p, p->T, p->V, (int)(N / 2), 1));
  VMK_TRY(launch_k3_rows(p, p->V, p->w[1]));
  VMK_TRY(be_d2h_2d(ut, sizeof(double) * n1, p->w[1] + N, sizeof(double) * N, sizeof(double) * N, N, p->st));
  VMK_TRY(be_sync(p->st));
  for (size_t j = 0; j < N; j++) ut[N + j * n1] = ut[j * n1];
  memcpy(ut + N * n1, ut, sizeof(double) * n1);
  return 0;
}

int check_plan(vmk_plan* p) {
  if (!p) return fail(VMK_EARG, "plan is NULL");
  VMK_TRY(be_set_device(p->device));  // a host thread may drive several plans on different devices
  if (p->nranks > 1 && !p->peers_ready) return fail(VMK_ESTATE, "slab plan: peers not attached yet");
  return 0;
}

// host ghosted array -> slab (rows j0..j0+NJ-1 plus periodic halo rows, read from the interior)
int upload_ghosted(vmk_plan* p, const double* host, double* slab) {
  VMK_TRY(ensure_staging(p));
  const size_t ld = (size_t)p->N + 2;
  const int N = p->N, NJ = p->NJ;
  VMK_TRY(be_h2d(p->staging + ld, host + ld * (size_t)(p->j0 + 1), sizeof(double) * ld * NJ, p->st));
  const int jlo = (p->j0 + N - 1) % N, jhi = (p->j0 + NJ) % N;  // interior rows that wrap into the halos
  VMK_TRY(be_h2d(p->staging, host + ld * (size_t)(jlo + 1), sizeof(double) * ld, p->st));
  VMK_TRY(be_h2d(p->staging + ld * (size_t)(NJ + 1), host + ld * (size_t)(jhi + 1), sizeof(double) * ld, p->st));
  VMK_TRY(launch_k5<K5Unpack>(p, p->staging, slab, slab_elems(p)));
  return 0;
}

// slab -> host ghosted rows j0 .. j0+NJ+1 (all of the array on a single GPU), i-ghosts included
int download_ghosted(vmk_plan* p, const double* slab, double* host) {
  VMK_TRY(ensure_staging(p));
  const size_t ld = (size_t)p->N + 2;
  VMK_TRY(launch_k5<K5Pack>(p, slab, p->staging, (size_t)(p->NJ + 2) * ld));
  VMK_TRY(be_d2h(host + ld * (size_t)p->j0, p->staging, sizeof(double) * ld * (p->NJ + 2), p->st));
  return 0;
}

// slab interior -> interior of a host ghosted array (ghost cells untouched)
int download_interior(vmk_plan* p, const double* slab, double* host) {
  const size_t ld = (size_t)p->N + 2;
  return be_d2h_2d(host + ld * (size_t)(p->j0 + 1) + 1, sizeof(double) * ld, slab + p->N, sizeof(double) * p->N,
                   sizeof(double) * p->N, (size_t)p->NJ, p->st);
}

// stream sync + the cross-GPU barrier's timeout flag (a peer that died must surface as an error, not as a hang)
int sync_and_check(vmk_plan* p) {
  VMK_TRY(be_sync(p->st));
#ifndef VMK_EMUL
  if (p->nranks > 1 && !p->barrier_fn) {
    int err = 0;
    VMK_CUDA_TRY(cudaMemcpy(&err, p->flags + kMaxPeers + 1, sizeof(int), cudaMemcpyDeviceToHost));
    if (err) return fail(VMK_ECUDA, "cross-GPU barrier timed out: a peer rank did not arrive");
  }
#endif
  return 0;
}

int collect_profile(vmk_plan* p) {
  VMK_TRY(be_sync(p->st));
  for (auto& e : p->prof_events) {
    double ms = 0;
    VMK_TRY(be_event_elapsed(e.second.first, e.second.second, &ms));
    p->prof_ms[e.first] += ms;
    p->prof_n[e.first]++;
    be_event_destroy(e.second.first);
    be_event_destroy(e.second.second);
  }
  p->prof_events.clear();
  return 0;
}

}  // namespace

// ============================================ C ABI =================================================
extern "C" {

int vmk_version(void) { return 100; }

const char* vmk_last_error(void) { return err_slot().c_str(); }

int vmk_plan_create_slab(int64_t nx, int64_t ny, int rank, int nranks, vmk_plan** out) {
  if (!out) return fail(VMK_EARG, "plan output pointer is NULL");
  *out = nullptr;
  if (nx != ny) return fail(VMK_ESIZE, "nx != ny: the reference aliases ky = kx (Common.jl:113)");
  const int M = ilog2_exact(nx);
  SizeOps ops;
  if (M < 0 || !ops_for(M, &ops))
    return fail(VMK_ESIZE, "grid size must be a power of two in [32, 32768]");
  if (nranks < 1 || nranks > kMaxPeers || rank < 0 || rank >= nranks || ilog2_exact(nranks) < 0)
    return fail(VMK_EARG, "nranks must be 1, 2, 4 or 8 and 0 <= rank < nranks");
  if ((nx / nranks) < 2 || ((nx / 2) % nranks) != 0)
    return fail(VMK_ESIZE, "too many ranks for this grid");
  vmk_plan* p = new vmk_plan();
  p->N = (int)nx;
  p->M = M;
  p->rank = rank;
  p->nranks = nranks;
  p->NJ = (int)(nx / nranks);
  p->log2NJ = ilog2_exact(p->NJ);
  p->j0 = rank * p->NJ;
  p->ops = ops;
  int rc = 0;
  do {
    if ((rc = be_get_device(&p->device))) break;
    if ((rc = be_num_sms(&p->sms))) break;
    if ((rc = be_stream_create(p->st))) break;
    if ((rc = be_event_create(p->ev0)) || (rc = be_event_create(p->ev1))) break;
    // second stream: the copies of the transposes (P > 1) / the FFT-form rows beside the recurrences (vmk_tri.cuh)
    if ((rc = be_stream_create(p->st_copy)) || (rc = be_event_create(p->ev_join))) break;
    for (int c = 0; c < 8 && !rc; c++) rc = be_event_create(p->ev_chunk[c]);
    if (rc) break;
    if (nranks > 1) {
      if ((rc = be_stream_create(p->st_copy2)) || (rc = be_event_create(p->ev_join2))) break;
    }
    if ((rc = ops.configure(&p->res_k1, &p->res_k2, &p->res_k3))) break;
    if (ops.ks && nranks == 1) {
      // one CTA per block of FPC row pairs, at most 8 (the portable cluster size); 0 = the cluster does not fit
      const int blocks = ((p->N / 2) + ops.fpc - 1) / ops.fpc;
      int q = blocks < 8 ? blocks : 8;
      while (q > 1 && ops.ks_configure(q)) q--;
      p->ks_q = (q >= 1 && ops.ks_configure(q) == 0) ? q : 0;
    }
    const size_t sb = sizeof(double) * slab_elems(p);
    for (int b = 0; b < 3 && !rc; b++) rc = dev_alloc(p, (void**)&p->w[b], sb);
    if (rc) break;
    if ((rc = dev_alloc(p, (void**)&p->psi, sb))) break;
    if ((rc = dev_alloc(p, (void**)&p->T, sizeof(double2) * (size_t)(p->N / 2) * p->NJ))) break;
    if ((rc = dev_alloc(p, (void**)&p->V, sizeof(double2) * (size_t)(p->N / 2) * p->NJ))) break;
    if (nranks > 1 && (rc = dev_alloc(p, (void**)&p->S, sizeof(double2) * (size_t)(p->N / 2) * p->NJ))) break;
    {
      // buffers of the recurrence form (vmk_tri.cuh); G and L always exist (tiny when unused) because peers map them
      const bool tri = ops.k1n && (p->NJ % kTriCH) == 0;
      const size_t H = (size_t)p->N / 2;
      p->tri_nch = tri ? p->NJ / kTriCH : 0;
      if (tri) {
        // fused form (one GPU): K1 and K3 on the same grid, so that their units own the same blocks of row pairs
        size_t blocks = (size_t)p->tri_nch;
        if (ops.kf_configure && nranks == 1) {
          int res = 0;
          if ((rc = ops.kf_configure(&res))) break;
          const int work = rowpair_units(p, p->NJ / 2, 1);
          p->fz_grid = work < res ? work : res;
          if ((size_t)p->fz_grid * ops.fpc > (size_t)kTriScanFMaxUnits) p->fz_grid = kTriScanFMaxUnits / ops.fpc;
          const size_t units = (size_t)p->fz_grid * ops.fpc;
          if (units > blocks) blocks = units;
          p->fz_units_cap = blocks;
          if ((rc = dev_alloc(p, (void**)&p->tri_ftab, sizeof(double) * kTriFTab * H))) break;
          if ((rc = dev_alloc(p, (void**)&p->tri_rr, sizeof(double2) * H))) break;
        }
        if ((rc = dev_alloc(p, (void**)&p->tri_tab, sizeof(double) * kTriTab * H))) break;
        if ((rc = dev_alloc(p, (void**)&p->tri_low, sizeof(int) * (H + kTriMaxK0)))) break;  // lowrow[H], lowslot[K0]
        if ((rc = dev_alloc(p, (void**)&p->tri_tot, sizeof(double2) * 3 * blocks * H))) break;
        if ((rc = dev_alloc(p, (void**)&p->tri_cin, sizeof(double2) * (2 * blocks + 1) * H))) break;
      }
      if ((rc = dev_alloc(p, (void**)&p->tri_G, tri ? sizeof(double2) * 3 * H * nranks : 16))) break;
      const size_t lrows = H < (size_t)kTriMaxK0 ? H : (size_t)kTriMaxK0;
      if ((rc = dev_alloc(p, (void**)&p->tri_L, tri ? sizeof(double2) * lrows * p->N : 16))) break;
    }
    if ((rc = dev_alloc(p, (void**)&p->tw, sizeof(double2) * (ops.twn ? ops.twn : 1)))) break;
    if ((rc = dev_alloc(p, (void**)&p->flags, sizeof(unsigned long long) * (kMaxPeers + 2)))) break;
#ifndef VMK_EMUL
    if ((rc = cudaMemset(p->flags, 0, sizeof(unsigned long long) * (kMaxPeers + 2)) != cudaSuccess ? fail(VMK_ECUDA, "cudaMemset(flags)") : 0)) break;
#else
    memset(p->flags, 0, sizeof(unsigned long long) * (kMaxPeers + 2));
#endif
    if ((rc = dev_alloc(p, (void**)&p->bbcos, sizeof(double) * p->N))) break;
    if ((rc = dev_alloc(p, (void**)&p->cccos, sizeof(double) * p->N))) break;
    if ((rc = dev_alloc(p, (void**)&p->ccperm, sizeof(double) * p->N))) break;
    std::vector<double2> tw(ops.twn ? ops.twn : 1);
    ops.fill_tw(tw.data());
    if ((rc = be_h2d(p->tw, tw.data(), sizeof(double2) * tw.size(), p->st))) break;
    if ((rc = be_sync(p->st))) break;
  } while (0);
  if (rc) {
    vmk_plan_destroy(p);
    return rc;
  }
  for (int r = 0; r < kMaxPeers; r++) {
    for (int b = 0; b < 3; b++) p->peer_w[b][r] = nullptr;
    p->peer_psi[r] = nullptr;
    p->peer_T[r] = nullptr;
    p->peer_V[r] = nullptr;
    p->peer_flags[r] = nullptr;
    p->peer_G[r] = nullptr;
    p->peer_L[r] = nullptr;
  }
  p->peer_G[rank] = p->tri_G;
  p->peer_L[rank] = p->tri_L;
  for (int b = 0; b < 3; b++) p->peer_w[b][rank] = p->w[b];
  p->peer_psi[rank] = p->psi;
  p->peer_T[rank] = p->T;
  p->peer_V[rank] = p->V;
  p->peer_flags[rank] = p->flags;
  p->peers_ready = (nranks == 1);
  *out = p;
  return VMK_OK;
}

int vmk_plan_create(int64_t nx, int64_t ny, vmk_plan** out) { return vmk_plan_create_slab(nx, ny, 0, 1, out); }

int vmk_plan_create_on(int device, int64_t nx, int64_t ny, int rank, int nranks, vmk_plan** out) {
  VMK_TRY(be_set_device(device));
  return vmk_plan_create_slab(nx, ny, rank, nranks, out);
}

int vmk_plan_destroy(vmk_plan* p) {
  if (!p) return VMK_OK;
  be_set_device(p->device);
  if (be_stream_valid(p->st)) be_sync(p->st);
  drop_graphs(p);
  for (void* q : p->ipc_opened) be_ipc_close(q);
  for (int b = 0; b < 3; b++) be_free(p->w[b]);
  be_free(p->psi);
  be_free(p->T);
  be_free(p->V);
  be_free(p->S);
  be_free(p->flags);
  be_free(p->tw);
  be_free(p->bbcos);
  be_free(p->cccos);
  be_free(p->ccperm);
  be_free(p->tri_tab);
  be_free(p->tri_low);
  be_free(p->tri_tot);
  be_free(p->tri_cin);
  be_free(p->tri_ftab);
  be_free(p->tri_rr);
  be_free(p->tri_G);
  be_free(p->tri_L);
  be_free(p->staging);
  be_free(p->hW);
  be_free(p->hJ);
  be_free(p->hVs);
  be_free(p->hksq);
  be_free(p->hksqperm);
  if (p->child) vmk_plan_destroy(p->child);
  be_free(p->qS);
  be_free(p->qJ);
  be_free(p->qY);
  be_free(p->qVF);
  be_free(p->qT9);
  be_free(p->qPi);
  be_free(p->qF);
  be_free(p->qtw);
  be_free(p->qtab);
  be_free(p->pV[0]);
  be_free(p->pV[1]);
  be_free(p->pA0);
  be_free(p->ptab);
  for (int b = 0; b < 3; b++) be_free(p->cw[b]);
  be_free(p->cs);
  be_free(p->csp);
  be_free(p->cpart);
  be_event_destroy(p->ev0);
  be_event_destroy(p->ev1);
  be_event_destroy(p->ev_join);
  be_event_destroy(p->ev_join2);
  for (int c = 0; c < 8; c++) be_event_destroy(p->ev_chunk[c]);
  be_stream_destroy(p->st_copy);
  be_stream_destroy(p->st_copy2);
  be_stream_destroy(p->st);
  delete p;
  return VMK_OK;
}

// ---- slab decomposition: peer buffer exchange ------------------------------------------------------
// A blob carries the handles of one rank's seven exchange buffers (w[0..2], psi, T, V, barrier flags).
constexpr int kPeerBufs = 9;
struct PeerBlob {
  IpcHandle h[kPeerBufs];
};

size_t vmk_peer_blob_bytes(void) { return sizeof(PeerBlob); }

int vmk_peer_export(vmk_plan* p, void* blob) {
  VMK_GUARD(p);
  if (!p || !blob) return fail(VMK_EARG, "NULL argument");
  PeerBlob* b = static_cast<PeerBlob*>(blob);
  void* ptrs[kPeerBufs] = {p->w[0], p->w[1], p->w[2], p->psi, p->T, p->V, p->flags, p->tri_G, p->tri_L};
  for (int i = 0; i < kPeerBufs; i++) VMK_TRY(be_ipc_export(ptrs[i], &b->h[i]));
  return VMK_OK;
}

// blobs: nranks blobs in rank order (own entry ignored): one process per GPU, handles via CUDA IPC
int vmk_peer_import(vmk_plan* p, const void* blobs) {
  VMK_GUARD(p);
  if (!p || !blobs) return fail(VMK_EARG, "NULL argument");
  const PeerBlob* b = static_cast<const PeerBlob*>(blobs);
  for (int r = 0; r < p->nranks; r++) {
    if (r == p->rank) continue;
    void* ptrs[kPeerBufs];
    for (int i = 0; i < kPeerBufs; i++) {
      VMK_TRY(be_ipc_open(b[r].h[i], &ptrs[i]));
      p->ipc_opened.push_back(ptrs[i]);
    }
    for (int q = 0; q < 3; q++) p->peer_w[q][r] = (double*)ptrs[q];
    p->peer_psi[r] = (double*)ptrs[3];
    p->peer_T[r] = (double2*)ptrs[4];
    p->peer_V[r] = (double2*)ptrs[5];
    p->peer_flags[r] = (unsigned long long*)ptrs[6];
    p->peer_G[r] = (double2*)ptrs[7];
    p->peer_L[r] = (double2*)ptrs[8];
  }
  p->peers_ready = true;
  return VMK_OK;
}

// all ranks live in this process (one host thread driving several devices, the Julia model):
// plans[r] is rank r's plan; peer access must already be enabled between the devices
int vmk_peer_attach_local(vmk_plan* p, vmk_plan* const* plans) {
  VMK_GUARD(p);
  if (!p || !plans) return fail(VMK_EARG, "NULL argument");
  VMK_TRY(be_set_device(p->device));
  for (int r = 0; r < p->nranks; r++) {
    if (!plans[r] || plans[r]->N != p->N || plans[r]->nranks != p->nranks || plans[r]->rank != r)
      return fail(VMK_EARG, "plans[] does not hold one matching plan per rank");
    VMK_TRY(be_enable_peer(p->device, plans[r]->device));
    for (int q = 0; q < 3; q++) p->peer_w[q][r] = plans[r]->w[q];
    p->peer_psi[r] = plans[r]->psi;
    p->peer_T[r] = plans[r]->T;
    p->peer_V[r] = plans[r]->V;
    p->peer_flags[r] = plans[r]->flags;
    p->peer_G[r] = plans[r]->tri_G;
    p->peer_L[r] = plans[r]->tri_L;
  }
  p->peers_ready = true;
  return VMK_OK;
}

int vmk_barrier_hook(vmk_plan* p, void (*fn)(void*), void* user) {
  VMK_GUARD(p);
  if (!p) return fail(VMK_EARG, "plan is NULL");
  p->barrier_fn = fn;
  p->barrier_user = user;
  return VMK_OK;
}

// ---- reference-signature entry points on host arrays -----------------------------------------------
int vmk_fps(vmk_plan* p, double dx, double dy, const double* f, double* s, double eps) {
  VMK_GUARD(p);
  VMK_TRY(check_plan(p));
  if (!f || !s) return fail(VMK_EARG, "f or s is NULL");
  VMK_TRY(ensure_divisor(p, dx, dy, eps));
  // wtA is scratch between steps; the source needs no halo rows
  VMK_TRY(be_h2d(p->w[1] + p->N, f + (size_t)p->j0 * p->N, sizeof(double) * (size_t)p->N * p->NJ, p->st));
  VMK_TRY(enqueue_poisson(p, p->w[1], +1.0));
  // recurrence form: a rank may not start its next solve (whose K1 stores into the peers' L and G) before every
  // rank has consumed this one's
  if (tri_on(p)) VMK_TRY(cross_rank_barrier(p));
  VMK_TRY(download_interior(p, p->psi, s));
  return sync_and_check(p);  // slab plans: a barrier that timed out is an error, not VMK_OK with stale peer data
}

int vmk_ps_fft(vmk_plan* p, double dx, double dy, const double* f, double* u, double eps) {
  VMK_GUARD(p);
  VMK_TRY(check_plan(p));
  if (!f || !u) return fail(VMK_EARG, "f or u is NULL");
  VMK_TRY(ensure_divisor(p, dx, dy, eps));
  const size_t N = (size_t)p->N;
  // f is (nx+1) x (ny+1); [1:nx, 1:ny] is read (fft_p.jl:23-27)
  VMK_TRY(be_h2d_2d(p->w[1] + N, sizeof(double) * N, f + (size_t)p->j0 * (N + 1), sizeof(double) * (N + 1),
                    sizeof(double) * N, (size_t)p->NJ, p->st));
  VMK_TRY(enqueue_poisson(p, p->w[1], +1.0));
  if (tri_on(p)) VMK_TRY(cross_rank_barrier(p));
  VMK_TRY(be_d2h(u + (size_t)p->j0 * N, p->psi + N, sizeof(double) * N * p->NJ, p->st));
  return sync_and_check(p);  // slab plans: a barrier that timed out is an error, not VMK_OK with stale peer data
}

int vmk_rhs(vmk_plan* p, double dx, double dy, double re, const double* w, double* r, double* s, double* f) {
  VMK_GUARD(p);
  VMK_TRY(check_plan(p));
  if (!w || !r || !s) return fail(VMK_EARG, "w, r or s is NULL");
  VMK_TRY(ensure_divisor(p, dx, dy, 1.e-6));  // vm_rhs calls fps with the default eps (Common.jl:136)
  VMK_TRY(upload_ghosted(p, w, p->w[1]));
  if (f) {
    // V is not written before the first cross-rank barrier of this call: borrow it for f = -w (Common.jl:134)
    double* fd = reinterpret_cast<double*>(p->V);
    VMK_TRY(launch_k5<K5Negate>(p, p->w[1], fd, (size_t)p->NJ * p->N));
    VMK_TRY(be_d2h(f + (size_t)p->j0 * p->N, fd, sizeof(double) * (size_t)p->N * p->NJ, p->st));
  }
  VMK_TRY(enqueue_poisson(p, p->w[1], -1.0));
  VMK_TRY(cross_rank_barrier(p));
  const StepParams sp{dx, dy, 0.0, re};
  VMK_TRY(launch_k4(p, 0, 1, 1, 2, sp));
  VMK_TRY(download_interior(p, p->w[2], r));
  VMK_TRY(download_ghosted(p, p->psi, s));
  return sync_and_check(p);  // slab plans: a barrier that timed out is an error, not VMK_OK with stale peer data
}

int vmk_upload(vmk_plan* p, const double* wn) {
  VMK_GUARD(p);
  VMK_TRY(check_plan(p));
  if (!wn) return fail(VMK_EARG, "wn is NULL");
  VMK_TRY(upload_ghosted(p, wn, p->w[0]));
  VMK_TRY(be_sync(p->st));
  p->uploaded = true;
  return VMK_OK;
}

int vmk_step(vmk_plan* p, double dx, double dy, double dt, double re, int64_t nsteps) {
  VMK_GUARD(p);
  VMK_TRY(check_plan(p));
  if (!p->uploaded) return fail(VMK_ESTATE, "vmk_step before vmk_upload");
  if (nsteps < 0) return fail(VMK_EARG, "nsteps < 0");
  VMK_TRY(ensure_divisor(p, dx, dy, 1.e-6));
  const StepParams sp{dx, dy, dt, re};
  VMK_TRY(be_event_record(p->ev0, p->st));
  if (small_fused_on(p) && nsteps > 0) {
    VMK_TRY(enqueue_small(p, sp, nsteps));
    VMK_TRY(be_event_record(p->ev1, p->st));
    p->ev_valid = true;
    return VMK_OK;
  }
#ifndef VMK_EMUL
  if (p->use_graph && !p->barrier_fn && !p->profiling && nsteps > 0) {
    auto it = p->graphs.find(sp);
    if (it == p->graphs.end()) {
      cudaGraph_t g = nullptr;
      cudaGraphExec_t ge = nullptr;
      VMK_CUDA_TRY(cudaStreamBeginCapture(p->st.s, cudaStreamCaptureModeThreadLocal));
      const int64_t before = p->launches;
      int rc = enqueue_step(p, sp);
      p->graph_launches = p->launches - before;
      p->launches = before;
      cudaError_t e = cudaStreamEndCapture(p->st.s, &g);
      if (rc) return rc;
      VMK_CUDA_TRY(e);
      VMK_CUDA_TRY(cudaGraphInstantiate(&ge, g, 0));
      cudaGraphDestroy(g);
      it = p->graphs.emplace(sp, ge).first;
    }
    for (int64_t k = 0; k < nsteps; k++) {
      VMK_CUDA_TRY(cudaGraphLaunch(it->second, p->st.s));
      p->launches += p->graph_launches;
    }
    VMK_TRY(cross_rank_barrier(p));  // the neighbours' last halo rows have landed (vmk_download reads them)
    VMK_TRY(be_event_record(p->ev1, p->st));
    p->ev_valid = true;
    return VMK_OK;
  }
#endif
  for (int64_t k = 0; k < nsteps; k++) VMK_TRY(enqueue_step(p, sp));
  VMK_TRY(cross_rank_barrier(p));
  VMK_TRY(be_event_record(p->ev1, p->st));
  p->ev_valid = true;
  return VMK_OK;
}

int vmk_download(vmk_plan* p, double* wn, double* psi) {
  VMK_GUARD(p);
  VMK_TRY(check_plan(p));
  if (!p->uploaded) return fail(VMK_ESTATE, "vmk_download before vmk_upload");
  if (wn) VMK_TRY(download_ghosted(p, p->w[0], wn));
  if (psi) VMK_TRY(download_ghosted(p, p->psi, psi));
  return sync_and_check(p);
}

int vmk_sync(vmk_plan* p) {
  VMK_GUARD(p);
  VMK_TRY(check_plan(p));
  return sync_and_check(p);
}

int vmk_numerical(vmk_plan* p, int64_t nt, double dx, double dy, double dt, double re, double* wn, double* out,
                  int64_t freq, vmk_snapshot_fn snap, void* user) {
  VMK_GUARD(p);
  VMK_TRY(check_plan(p));
  if (!wn) return fail(VMK_EARG, "wn is NULL");
  if (nt < 0) return fail(VMK_EARG, "nt < 0");
  if (out && p->nranks != 1) return fail(VMK_EARG, "out is only produced by single-GPU plans");
  VMK_TRY(vmk_upload(p, wn));
  if (snap && freq > 0) {
    for (int64_t k = 0; k < nt;) {
      const int64_t chunk = (freq - (k % freq)) < (nt - k) ? (freq - (k % freq)) : (nt - k);
      VMK_TRY(vmk_step(p, dx, dy, dt, re, chunk));
      k += chunk;
      if (k % freq == 0) {  // vm.jl:78
        VMK_TRY(vmk_download(p, wn, nullptr));
        snap(k, wn, user);
      }
    }
  } else {
    VMK_TRY(vmk_step(p, dx, dy, dt, re, nt));
  }
  VMK_TRY(vmk_download(p, wn, nullptr));
  if (out) {
    // wn[2:nx+2, 2:ny+2] (vm.jl:89): a strided view of the array that was just downloaded
    const size_t ld = (size_t)p->N + 2, n1 = (size_t)p->N + 1;
    for (size_t j = 0; j < n1; j++) memcpy(out + j * n1, wn + (j + 1) * ld + 1, sizeof(double) * n1);
  }
  return VMK_OK;
}

int vmk_hybrid_numerical(vmk_plan* p, int64_t nt, double dx, double dy, double dt, double re, const double* wn,
                         double* ut, int64_t freq, vmk_snapshot_fn snap, void* user) {
  VMK_GUARD(p);
  VMK_TRY(check_plan(p));
  if (!wn || !ut) return fail(VMK_EARG, "wn or ut is NULL");
  if (nt < 0) return fail(VMK_EARG, "nt < 0");
  if (dx != dy) return fail(VMK_EARG, "the hybrid solver needs dx == dy (wavespace aliases ky = kx, Common.jl:197)");
  VMK_TRY(ensure_hybrid(p, dx));
  VMK_TRY(upload_ghosted(p, wn, p->w[0]));
  VMK_TRY(be_event_record(p->ev0, p->st));
  VMK_TRY(launch_k1(p, p->w[0]));
  VMK_TRY(launch_kh(p, 0, dt, re));  // wf = fft(w0), wf[1,1] = 0 (hybrid.jl:26-27) and the first half of jacobian(wf)
  const StepParams sp{dx, dy, 0.0, INFINITY};  // 1/re = 0: K4 in mode 0 returns -J/3 exactly (hybrid.jl:130-149)
  for (int64_t k = 1; k <= nt; k++) {
    for (int stage = 1; stage <= 3; stage++) {
      VMK_TRY(launch_k3_rows(p, p->V, p->w[1]));   // w   = real(ifft(wf))        hybrid.jl:109
      VMK_TRY(launch_k3_rows(p, p->hVs, p->psi));  // psi = real(ifft(wf / k2))   hybrid.jl:124-126
      VMK_TRY(launch_k4(p, 0, 1, 1, 2, sp));       // -J(w, psi)/3                hybrid.jl:130-149
      VMK_TRY(launch_k1(p, p->w[2]));              // fft along i                 hybrid.jl:151
      VMK_TRY(launch_kh(p, stage, dt, re));        // fft along j, update, ifft along j of wf' and wf'/k2
    }
    if (snap && freq > 0 && k % freq == 0 && k != nt) {  // hybrid.jl:71-86 (the final field is produced below)
      VMK_TRY(hybrid_field(p, ut));
      snap(k, ut, user);
    }
  }
  VMK_TRY(be_event_record(p->ev1, p->st));
  p->ev_valid = true;
  VMK_TRY(hybrid_field(p, ut));
  if (snap && freq > 0 && nt > 0 && nt % freq == 0) snap(nt, ut, user);
  p->uploaded = false;  // w[0..2] were used as scratch
  return VMK_OK;
}

int vmk_ps23_numerical(vmk_plan* p, int64_t nt, double dx, double dy, double dt, double re, const double* wn,
                       double* ut, int64_t freq, vmk_snapshot_fn snap, void* user) {
  VMK_GUARD(p);
  VMK_TRY(check_plan(p));
  if (!wn || !ut) return fail(VMK_EARG, "wn or ut is NULL");
  if (nt < 0) return fail(VMK_EARG, "nt < 0");
  if (dx != dy) return fail(VMK_EARG, "the pseudo-spectral solver needs dx == dy (wavespace aliases ky = kx, Common.jl:197)");
  VMK_TRY(ensure_ps23(p, dx));
  VMK_TRY(upload_ghosted(p, wn, p->w[0]));
  VMK_TRY(be_event_record(p->ev0, p->st));
  VMK_TRY(launch_k1(p, p->w[0]));
  VMK_TRY(launch_kp(p, 0, dt, re));  // wf = fft(w0), wf[1,1] = 0 (:24-27) and the spectral half of jacobian(wf)
  for (int64_t k = 1; k <= nt; k++) {
    for (int stage = 1; stage <= 3; stage++) {
      VMK_TRY(launch_k3_rows(p, p->V, p->w[0]));      // j1 = real(ifft(i kx wf / k2))    pseudospectral_23_rule.jl:135
      VMK_TRY(launch_k3_rows(p, p->hVs, p->w[1]));    // j2 = real(ifft(i ky wf))
      VMK_TRY(launch_k3_rows(p, p->pV[0], p->w[2]));  // j3 = real(ifft(i ky wf / k2))
      VMK_TRY(launch_k3_rows(p, p->pV[1], p->psi));   // j4 = real(ifft(i kx wf))
      VMK_TRY(launch_kp_product(p, p->w[0], p->w[1], p->w[2], p->psi));  // j1 j2 - j3 j4       :138-141
      VMK_TRY(launch_k1(p, p->w[0]));                 // fft along i                      :143
      VMK_TRY(launch_kp(p, stage, dt, re));           // fft along j, update, the four ifft along j for the next stage
    }
    if (snap && freq > 0 && k % freq == 0 && k != nt) {  // :69-86 (the final field is produced below)
      VMK_TRY(ps23_field(p, ut, dt, re));
      snap(k, ut, user);
    }
  }
  VMK_TRY(be_event_record(p->ev1, p->st));
  p->ev_valid = true;
  VMK_TRY(ps23_field(p, ut, dt, re));
  if (snap && freq > 0 && nt > 0 && nt % freq == 0) snap(nt, ut, user);
  p->uploaded = false;  // w[0..2], psi were used as scratch
  return VMK_OK;
}

int vmk_ps32_numerical(vmk_plan* p, int64_t nt, double dx, double dy, double dt, double re, const double* wn,
                       double* ut, int64_t freq, vmk_snapshot_fn snap, void* user) {
  VMK_GUARD(p);
  VMK_TRY(check_plan(p));
  if (!wn || !ut) return fail(VMK_EARG, "wn or ut is NULL");
  if (nt < 0) return fail(VMK_EARG, "nt < 0");
  if (dx != dy) return fail(VMK_EARG, "the pseudo-spectral solver needs dx == dy (wavespace aliases ky = kx, Common.jl:197)");
  VMK_TRY(ensure_ps32(p, dx));
  const size_t L = (size_t)p->N / 2;
  const int64_t child_launches0 = p->child->launches;
  p->child->profiling = p->profiling;
  VMK_TRY(upload_ghosted(p, wn, p->w[0]));
  VMK_TRY(be_event_record(p->ev0, p->st));
  VMK_TRY(launch_k1(p, p->w[0]));                               // wnf = fft(data), nx x ny   :24
  VMK_TRY(launch_kx(p, p->T, p->V, (int)L, 0));
  VMK_TRY(launch_p32<P32Init>(p, p32_args(p, 0, dt, re), (L + 1) * (2 * L + 1)));  // -> S, wnf[1,1] = 0  :27
  VMK_TRY(be_sync(p->st));
  for (int64_t k = 1; k <= nt; k++) {
    for (int stage = 1; stage <= 3; stage++) VMK_TRY(ps32_stage(p, stage, dt, re));
    if (snap && freq > 0 && k % freq == 0 && k != nt) {  // :69-86 (the final field is produced below)
      VMK_TRY(ps32_field(p, ut, dt, re));
      snap(k, ut, user);
    }
  }
  VMK_TRY(be_sync(p->child->st));
  VMK_TRY(be_event_record(p->ev1, p->st));
  p->ev_valid = true;
  VMK_TRY(ps32_field(p, ut, dt, re));
  if (snap && freq > 0 && nt > 0 && nt % freq == 0) snap(nt, ut, user);
  p->launches += p->child->launches - child_launches0;
  p->uploaded = false;  // w[0..1], T, V were used as scratch
  return VMK_OK;
}

// ---- the callers' snapshot files (vmk_io.hpp; host code, no device involved) ------------------------------------------
int vmk_print_float64(double v, char* buf) { return buf ? julia_print_f64(v, buf) : 0; }

int vmk_write_field(const char* path, const double* x, const double* y, const double* ut, int64_t nx1, int64_t ny1) {
  if (!path || !x || !y || !ut) return fail(VMK_EARG, "path, x, y or ut is NULL");
  if (nx1 < 0 || ny1 < 0) return fail(VMK_EARG, "negative extent");
  std::string err;
  if (write_field_text(path, x, y, ut, nx1, ny1, &err)) return fail(VMK_EIO, err);
  return VMK_OK;
}

int vmk_read_field(const char* path, double* x, double* y, double* w, int64_t cap, int64_t* nrows) {
  if (!path) return fail(VMK_EARG, "path is NULL");
  std::string err;
  if (read_field_text(path, x, y, w, cap < 0 ? 0 : cap, nrows, &err)) return fail(VMK_EIO, err);
  return VMK_OK;
}

int vmk_ldc_numerical(vmk_plan* p, int64_t nx, int64_t ny, int64_t nt, double dx, double dy, double dt, double re,
                      double* wn, double* sn, double* rms) {
  VMK_GUARD(p);
  VMK_TRY(check_plan(p));
  if (!wn || !sn || (nt > 0 && !rms)) return fail(VMK_EARG, "wn, sn or rms is NULL");
  if (nt < 0) return fail(VMK_EARG, "nt < 0");
  if (p->nranks != 1) return fail(VMK_EARG, "the cavity solver runs on single-GPU plans");
  if (nx != ny || 2 * nx != p->N || nx < 16)
    return fail(VMK_ESIZE, "cavity of nx x ny cells needs nx == ny >= 16 and a plan of size 2nx x 2ny (odd extension)");
  const size_t n = (size_t)nx, nodes = (n + 1) * (n + 1), nb = sizeof(double) * nodes;
  const int nparts = p->sms * 4;
  if (!p->cs) {
    for (int b = 0; b < 3; b++) VMK_TRY(dev_alloc(p, (void**)&p->cw[b], nb));
    VMK_TRY(dev_alloc(p, (void**)&p->cs, nb));
    VMK_TRY(dev_alloc(p, (void**)&p->csp, nb));
    VMK_TRY(dev_alloc(p, (void**)&p->cpart, sizeof(double) * (size_t)nparts));
  }
  double* rms_dev = nullptr;
  if (nt > 0) VMK_TRY(be_malloc((void**)&rms_dev, sizeof(double) * (size_t)nt));
  int rc = 0;
  do {
    if ((rc = ensure_divisor_cavity(p, dx, dy))) break;
    if ((rc = be_h2d(p->cw[0], wn, nb, p->st)) || (rc = be_h2d(p->cs, sn, nb, p->st))) break;
    // the wall values of the stage arrays are written by bc2 before they are read; start them from wn anyway
    if ((rc = be_d2d(p->cw[1], p->cw[0], nb, p->st)) || (rc = be_d2d(p->cw[2], p->cw[0], nb, p->st))) break;
    KCArgs a{};
    a.n = (int)nx;
    a.s = p->cs;
    a.aa = 1.0 / (re * (dx * dx));  // lid_driven_cavity.jl:125-128
    a.bb = 1.0 / (re * (dy * dy));
    a.gg = 1.0 / (4.0 * dx * dy);
    a.hh = 1.0 / 3.0;
    a.dt = dt;
    a.dx2 = dx * dx;
    a.dy2 = dy * dy;
    a.lid = 3.0 / dy;
    a.count = (double)nodes;
    a.part = p->cpart;
    a.nparts = nparts;
    const size_t m2 = (n - 1) * (n - 1);
    VMK_TRY(be_event_record(p->ev0, p->st));
    for (int64_t k = 0; k < nt && !rc; k++) {
      if ((rc = be_d2d(p->csp, p->cs, nb, p->st))) break;  // sp = sn, :77
      // stage 1: wtA = wn + dt r(wn, sn); bc2; sn = fps_sine(-wtA)          :80-87
      a.w = p->cw[0], a.wn = p->cw[0], a.out = p->cw[1];
      if ((rc = launch_kc<KCStage<1>>(p, a, m2))) break;
      if ((rc = launch_kc<KCBc2>(p, a, 4 * (n + 1)))) break;
      if ((rc = cavity_poisson(p, a, p->cw[1]))) break;
      // stage 2: wtB = .75 wn + .25 wtA + .25 dt r(wtA, sn)                   :90-100
      a.w = p->cw[1], a.wn = p->cw[0], a.out = p->cw[2];
      if ((rc = launch_kc<KCStage<2>>(p, a, m2))) break;
      if ((rc = launch_kc<KCBc2>(p, a, 4 * (n + 1)))) break;
      if ((rc = cavity_poisson(p, a, p->cw[2]))) break;
      // stage 3: wn = (1/3) wn + (2/3) wtB + (2/3) dt r(wtB, sn)              :103-113
      a.w = p->cw[2], a.wn = p->cw[0], a.out = p->cw[0];
      if ((rc = launch_kc<KCStage<3>>(p, a, m2))) break;
      if ((rc = launch_kc<KCBc2>(p, a, 4 * (n + 1)))) break;
      if ((rc = cavity_poisson(p, a, p->cw[0]))) break;
      // rms[k] = sqrt(sum((sn - sp)^2) / ((nx+1)(ny+1)))                      :111-113
      a.w = p->csp, a.rms = rms_dev + k;
      if ((rc = launch_kc<KCRmsPartial>(p, a, nodes, sizeof(double) * kKCThreads, nparts))) break;
      if ((rc = launch_kc<KCRmsFinal>(p, a, 1, 0, 1))) break;
    }
    if (rc) break;
    if ((rc = be_event_record(p->ev1, p->st))) break;
    p->ev_valid = true;
    if ((rc = be_d2h(wn, p->cw[0], nb, p->st)) || (rc = be_d2h(sn, p->cs, nb, p->st))) break;
    if (nt > 0 && (rc = be_d2h(rms, rms_dev, sizeof(double) * (size_t)nt, p->st))) break;
    rc = be_sync(p->st);
  } while (0);
  if (rms_dev) {
    be_sync(p->st);
    be_free(rms_dev);
  }
  p->uploaded = false;  // w[1], psi were used as scratch
  return rc;
}

// ---- measurement -----------------------------------------------------------------------------------
void* vmk_stream(vmk_plan* p) {
#ifndef VMK_EMUL
  return p ? (void*)p->st.s : nullptr;
#else
  (void)p;
  return nullptr;
#endif
}

int vmk_step_elapsed_ms(vmk_plan* p, double* ms) {
  VMK_GUARD(p);
  VMK_TRY(check_plan(p));
  if (!ms) return fail(VMK_EARG, "ms is NULL");
  if (!p->ev_valid) return fail(VMK_ESTATE, "no vmk_step call to time");
  return be_event_elapsed(p->ev0, p->ev1, ms);
}

int vmk_profile_steps(vmk_plan* p, double dx, double dy, double dt, double re, int64_t nsteps, double* ms,
                      int64_t* launches) {
  VMK_GUARD(p);
  VMK_TRY(check_plan(p));
  if (!p->uploaded) return fail(VMK_ESTATE, "vmk_profile_steps before vmk_upload");
  for (int k = 0; k < KI_COUNT; k++) {
    p->prof_ms[k] = 0;
    p->prof_n[k] = 0;
  }
  const bool was = p->profiling;
  p->profiling = true;
  int rc = vmk_step(p, dx, dy, dt, re, nsteps);
  if (!rc) rc = collect_profile(p);
  p->profiling = was;
  VMK_TRY(rc);
  for (int k = 0; k < 3; k++) {
    p->tri_ms[k] = p->prof_ms[KI_KT1 + k];
    p->tri_n[k] = p->prof_n[KI_KT1 + k];
    p->prof_ms[KI_K2] += p->prof_ms[KI_KT1 + k];
    p->prof_n[KI_K2] += p->prof_n[KI_KT1 + k];
  }
  for (int k = 0; k < KI_ABI; k++) {
    if (ms) ms[k] = p->prof_ms[k];
    if (launches) launches[k] = p->prof_n[k];
  }
  return VMK_OK;
}

int vmk_profile_tri(vmk_plan* p, double* ms, int64_t* launches) {
  VMK_GUARD(p);
  VMK_TRY(check_plan(p));
  for (int k = 0; k < 3; k++) {
    if (ms) ms[k] = p->tri_ms[k];
    if (launches) launches[k] = p->tri_n[k];
  }
  return VMK_OK;
}

int vmk_profile_read(vmk_plan* p, double* ms, int64_t* launches) {
  VMK_GUARD(p);
  VMK_TRY(check_plan(p));
  VMK_TRY(collect_profile(p));
  if (p->child) VMK_TRY(collect_profile(p->child));
  for (vmk_plan* q : {p, p->child}) {
    if (!q) continue;
    for (int k = KI_KT1; k < KI_COUNT; k++) {
      q->prof_ms[KI_K2] += q->prof_ms[k];
      q->prof_n[KI_K2] += q->prof_n[k];
      q->prof_ms[k] = 0;
      q->prof_n[k] = 0;
    }
  }
  for (int k = 0; k < KI_ABI; k++) {
    if (ms) ms[k] = p->prof_ms[k] + (p->child ? p->child->prof_ms[k] : 0.0);
    if (launches) launches[k] = p->prof_n[k] + (p->child ? p->child->prof_n[k] : 0);
    p->prof_ms[k] = 0;
    p->prof_n[k] = 0;
    if (p->child) {
      p->child->prof_ms[k] = 0;
      p->child->prof_n[k] = 0;
    }
  }
  return VMK_OK;
}

int64_t vmk_launch_count(vmk_plan* p) { return p ? p->launches : 0; }

int vmk_set_option(vmk_plan* p, const char* key, int64_t value) {
  VMK_GUARD(p);
  if (!p || !key) return fail(VMK_EARG, "NULL argument");
  const std::string k(key);
  int* knob = k == "v_pieces" ? &p->v_pieces : k == "k1_prefetch" ? &p->k1_prefetch : k == "k2_prefetch" ? &p->k2_prefetch
              : k == "cl_prefetch" ? &p->cl_prefetch
              : k == "k4_ahead" ? &p->k4_ahead : k == "a2a_chunks" ? &p->a2a_chunks
              : k == "a2a_engine" ? &p->a2a_engine : k == "a2a_ctas" ? &p->a2a_ctas
              : k == "k2_push" ? &p->k2_push : k == "a2a_order" ? &p->a2a_order
              : k == "k2_chunks" ? &p->k2_chunks : k == "k4_waves" ? &p->k4_waves : nullptr;
  if (knob) {
    const int64_t hi = k == "a2a_ctas" ? 4096 : (k == "a2a_chunks" || k == "k2_chunks") ? 8 : 64;
    const int64_t lo = (k == "a2a_ctas" || k.find("group") != std::string::npos) ? 1
                       : (k == "a2a_engine" || k == "k2_push" || k == "cl_prefetch") ? -1 : 0;
    if (value < lo || value > hi) return fail(VMK_EARG, "option value out of range");
    *knob = (int)value;
    drop_graphs(p);
  } else if (k == "ps32_fuse") {
    if (value < 0 || value > 2) return fail(VMK_EARG, "ps32_fuse: 0, 1 or 2");
    p->ps32_fuse = (int)value;
  } else if (k == "profile") {
    p->profiling = value != 0;  // events around every kernel of the following calls; read with vmk_profile_read
    if (p->child) p->child->profiling = p->profiling;
  } else if (k == "graph") {
    p->use_graph = value != 0;
  } else if (k == "fuse_small") {
    p->fuse_small = value != 0;
  } else if (k == "fps_mode") {  // 0: FFT along j (K2); 1: recurrences along j (vmk_tri.cuh); 2: ... inside K1 / K3
                                 // (fused form, one GPU); -1: by grid size
    if (value > 0 && !p->tri_tab)
      return fail(VMK_ESIZE, "the recurrence form needs 32 | rows per rank and N >= 64");
    if (value == 2 && (!p->tri_ftab || p->nranks != 1))
      return fail(VMK_ESIZE, "the fused form needs one GPU and N in [512, 8192]");
    p->fps_mode = value < 0 ? -1 : (value >= 2 ? 2 : (int)value);
    drop_graphs(p);
  } else if (k == "fz_grid") {  // fused form: CTAs of K1 / K3 (their units own blocks of row pairs); tuning / tests
    if (!p->tri_ftab) return fail(VMK_ESIZE, "the fused form needs one GPU and N in [512, 8192]");
    const int work = rowpair_units(p, p->NJ / 2, 1);
    if (value < 1 || value > work || (size_t)value * p->ops.fpc > p->fz_units_cap)
      return fail(VMK_EARG, "fz_grid out of range");
    p->fz_grid = (int)value;
    p->div_valid = false;  // the block-length tables follow the grid
    drop_graphs(p);
  } else if (k == "zigzag") {  // fps_mode 1: consecutive streaming kernels sweep the rows in alternating directions
    p->zigzag = value != 0;
    drop_graphs(p);
  } else if (k == "tri_k0") {  // rows kx < K0 keep the FFT form (0: N/16, at most 64)
    if (value < 0 || value > kTriMaxK0) return fail(VMK_EARG, "tri_k0 out of range");
    p->tri_k0 = (int)value;
    p->div_valid = false;  // the slot tables are rebuilt by the next call
    drop_graphs(p);
  } else if (k == "k4_rows") {
    if (value < 1 || value > 8192) return fail(VMK_EARG, "k4_rows out of range");
    p->k4_rows = (int)value;
    drop_graphs(p);
  } else {
    return fail(VMK_EARG, "unknown option: " + k);
  }
  return VMK_OK;
}

int64_t vmk_device_bytes(vmk_plan* p) { return p ? p->dev_bytes : 0; }

}  // extern "C"
