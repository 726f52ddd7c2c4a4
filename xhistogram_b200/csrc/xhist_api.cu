// xhist_api.cu — host side of the C-ABI declared in include/xhist_b200.h.
//
// Responsibilities: per-device context (stream, workspaces), exact preparation of the bin
// edges for the device compare (SURVEY.md §8a rules R1/R2), the launch plan (shared-memory
// histogram mode, partition, grid), the pinned/pageable host->device pipeline for host
// inputs, the single-process multi-GPU fan-out and the NCCL reduction of partial histograms.
// The O(samples) work itself is only ever done by the kernels in xhist_kernels.cu — there is
// no CPU fallback: if no sm_100 device is usable every entry point fails with XH_ERR_NO_DEVICE.
#include "../../include/xhist_b200.h"
#include "xhist_kernels.cuh"

#include <dlfcn.h>
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

thread_local std::string g_err;
// host-side phase clock of the calling thread's last xh_hist (microseconds since entry): [0] edge tables ready,
// [1] everything enqueued, [2] stream synchronised, [3] return.  Read with xh_last_call_phases().
thread_local double g_phase[4] = {0, 0, 0, 0};
thread_local std::chrono::steady_clock::time_point g_t0;
inline void phase_mark(int i) { g_phase[i] = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - g_t0).count(); }

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  g_err = buf;
  return code;
}

#define CU(call)                                                                                     \
  do {                                                                                               \
    cudaError_t e__ = (call);                                                                        \
    if (e__ != cudaSuccess)                                                                          \
      return fail(e__ == cudaErrorMemoryAllocation ? XH_ERR_NOMEM : XH_ERR_CUDA, "%s failed: %s (%s:%d)", #call, \
                  cudaGetErrorString(e__), __FILE__, __LINE__);                                      \
  } while (0)

size_t dsize(int dt) { return dt == XH_F32 ? 4 : (dt == XH_F64 || dt == XH_I64) ? 8 : 0; }

// ------------------------------------------------------------------------------------------ NCCL (dlopen)
struct Id128 { char b[XH_NCCL_UNIQUE_ID_BYTES]; };
struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, Id128, int) = nullptr;
  int (*CommInitAll)(void**, int, const int*) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
std::mutex g_nccl_mu;

int nccl_load() {
  std::lock_guard<std::mutex> lk(g_nccl_mu);
  if (g_nccl.handle) return XH_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) { h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
  if (!h) return fail(XH_ERR_NCCL, "NCCL not found (dlopen libnccl.so.2): %s", dlerror());
#define SYM(field, name)                                                             \
  *reinterpret_cast<void**>(&g_nccl.field) = dlsym(h, name);                         \
  if (!g_nccl.field) return fail(XH_ERR_NCCL, "NCCL symbol %s missing", name);
  SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommInitAll, "ncclCommInitAll")
  SYM(CommDestroy, "ncclCommDestroy") SYM(AllReduce, "ncclAllReduce") SYM(AllGather, "ncclAllGather") SYM(GroupStart, "ncclGroupStart")
  SYM(GroupEnd, "ncclGroupEnd") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  g_nccl.handle = h;
  return XH_OK;
}
#define NC(call)                                                                                  \
  do {                                                                                            \
    int r__ = (call);                                                                             \
    if (r__ != 0) return fail(XH_ERR_NCCL, "%s failed: %s", #call, g_nccl.GetErrorString(r__));    \
  } while (0)
constexpr int kNcclInt8 = 0, kNcclInt32 = 2, kNcclInt64 = 4, kNcclFloat64 = 8, kNcclSum = 0, kNcclMin = 3;

// ------------------------------------------------------------------------------------------ context
struct Ctx {
  int device = -1;
  std::mutex mu;
  cudaStream_t stream = nullptr, copy_stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, tev0 = nullptr, tev1 = nullptr;
  cudaEvent_t copied[2] = {nullptr, nullptr}, consumed[2] = {nullptr, nullptr};
  int sm_count = 0, smem_optin = 0, smem_per_sm = 0;
  XhkWindow* window = nullptr;       // device
  void* edges = nullptr;             // device edge table
  size_t edges_cap = 0;
  double* minmax = nullptr;          // device [296*3]
  void* stage[2] = {nullptr, nullptr};  // device staging slots for host inputs
  size_t stage_cap = 0;
  void* bcast = nullptr;             // device copies of broadcast (stride 0) rows
  size_t bcast_cap = 0;
  void* flush = nullptr;             // L2 flush scratch
  size_t flush_bytes = 0;
  double* widths = nullptr;          // device bin widths (density)
  size_t widths_cap = 0;
  void* rowsums = nullptr;           // device row sums (density), 8 bytes per row
  size_t rowsums_cap = 0;
  void* outbuf = nullptr;            // device histogram when the caller's out is host memory
  size_t outbuf_cap = 0;
  void* outpin = nullptr;            // pinned landing buffer for small results (a D2H straight into pageable memory costs
  size_t outpin_cap = 0;             // tens of microseconds of driver staging; DMA into pinned memory + memcpy does not)
  void* comm = nullptr;              // NCCL communicator (multi-process mode)
  int comm_ranks = 0, comm_rank = 0;
  // peer-memory reduction of small partial histograms (one rank per GPU, same node): every rank's symmetric buffer
  // [flags | slot 0 | slot 1] is mapped into every other rank through CUDA IPC; see peer_allreduce()
  struct Peer {
    int state = 0;                   // 0 not tried, 1 usable, -1 unavailable (NCCL is used instead)
    size_t cap = 0;                  // bytes per slot
    unsigned char* local = nullptr;
    unsigned char* mapped[XHK_MAX_PEERS] = {};
    unsigned long long seq = 0;
  } peer;
  // caches (guarded by mu): prepared edge tables by edge content, probe verdicts by (tables, buffers, shape)
  std::vector<struct PrepEntry*> preps;
  struct Verdict* verdicts = nullptr;     // kVerdictSlots entries; entry i owns slot i of the two slabs below
  unsigned char* vslab_dev = nullptr;     // device: per slot [XhkWindow][u64 slow-path counter at +kVerdictStatsOff]
  unsigned char* vslab_host = nullptr;    // pinned mirror of the same
  cudaEvent_t vev[32];
  unsigned long long stamp = 0, next_prep_id = 1;
  unsigned long long* dummy_stats = nullptr;   // device counter for launches that bypass the verdict cache
};
constexpr int kVerdictSlots = 32, kVerdictSlotBytes = 128, kVerdictStatsOff = 96;
static_assert(sizeof(XhkWindow) <= kVerdictStatsOff, "verdict slot layout");

std::mutex g_ctx_mu;
std::map<int, Ctx*> g_ctx;
struct Verdict* alloc_verdicts();

int get_ctx(int device, Ctx** out) {
  std::lock_guard<std::mutex> lk(g_ctx_mu);
  auto it = g_ctx.find(device);
  if (it != g_ctx.end()) { *out = it->second; return XH_OK; }
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) return fail(XH_ERR_NO_DEVICE, "no CUDA device: %s", cudaGetErrorString(e));
  if (device < 0 || device >= n) return fail(XH_ERR_INVALID, "device %d out of range (have %d)", device, n);
  cudaDeviceProp pr;
  CU(cudaGetDeviceProperties(&pr, device));
  if (pr.major != 10) return fail(XH_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device, pr.major, pr.minor);
  CU(cudaSetDevice(device));
  Ctx* c = new Ctx();
  c->device = device; c->sm_count = pr.multiProcessorCount;
  c->smem_optin = static_cast<int>(pr.sharedMemPerBlockOptin);
  c->smem_per_sm = static_cast<int>(pr.sharedMemPerMultiprocessor);
  CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  CU(cudaEventCreate(&c->ev0)); CU(cudaEventCreate(&c->ev1)); CU(cudaEventCreate(&c->tev0)); CU(cudaEventCreate(&c->tev1));
  for (int i = 0; i < 2; ++i) { CU(cudaEventCreateWithFlags(&c->copied[i], cudaEventDisableTiming)); CU(cudaEventCreateWithFlags(&c->consumed[i], cudaEventDisableTiming)); }
  CU(cudaMalloc(&c->window, sizeof(XhkWindow)));
  CU(cudaMalloc(&c->minmax, 296 * 3 * sizeof(double)));
  CU(cudaMalloc(&c->vslab_dev, kVerdictSlots * kVerdictSlotBytes + 64));
  CU(cudaMemset(c->vslab_dev, 0, kVerdictSlots * kVerdictSlotBytes + 64));
  c->dummy_stats = reinterpret_cast<unsigned long long*>(c->vslab_dev + kVerdictSlots * kVerdictSlotBytes);
  CU(cudaHostAlloc(reinterpret_cast<void**>(&c->vslab_host), kVerdictSlots * kVerdictSlotBytes, cudaHostAllocPortable));
  std::memset(c->vslab_host, 0, kVerdictSlots * kVerdictSlotBytes);
  for (int i = 0; i < kVerdictSlots; ++i) CU(cudaEventCreateWithFlags(&c->vev[i], cudaEventDisableTiming));
  c->verdicts = alloc_verdicts();
  CU(xhk_set_smem_limits(c->smem_optin - XHK_STATIC_SMEM));  // static shared memory of k_hist
  g_ctx[device] = c;
  *out = c;
  return XH_OK;
}

// ------------------------------------------------------------------------------------------ edges
// Effective edges for the device compare.  numpy compares data and edges after promotion
// (fp32 data vs float64 edges -> float64).  For fp32 data x and a float64 edge e:
//   e <= x  <=>  up32(e) <= x   with up32(e) the smallest fp32 >= e,
//   x <= e  <=>  x <= dn32(e)   with dn32(e) the largest  fp32 <= e,
// so an fp32 compare against the rounded-up edges is exact.
float up32(double e) { float f = static_cast<float>(e); if (static_cast<double>(f) < e) f = std::nextafterf(f, INFINITY); return f; }
float dn32(double e) { float f = static_cast<float>(e); if (static_cast<double>(f) > e) f = std::nextafterf(f, -INFINITY); return f; }

template <typename T> struct CSel;
template <> struct CSel<float> { static float* row(XhkParams& p, int k) { return p.cf[k]; } };
template <> struct CSel<double> { static double* row(XhkParams& p, int k) { return p.cd[k]; } };

// Fill the classification constants of variable k; append its effective edges to `table`.
template <typename T> void set_lut_inv(XhkParams& p, int k, long double v);
template <> void set_lut_inv<float>(XhkParams& p, int k, long double v) { p.lut_invf[k] = static_cast<float>(v); }
template <> void set_lut_inv<double>(XhkParams& p, int k, long double v) { p.lut_invd[k] = static_cast<double>(v); }

// Lookup table of a non-uniform variable: G equal cells over [lo, hi]; entry c = (#effective edges <= left
// boundary of cell c) - 1.  The device brackets its search with the entries of cells c-1 and c+2, so a
// rounding error of up to one cell in the device's cell index cannot exclude the true bin.
template <typename T>
void build_lut(int k, XhkParams& p, const std::vector<T>& table, std::vector<unsigned short>& lut) {
  const int nb = p.nb[k];
  p.lut_n[k] = 0; p.lut_steps[k] = 0; p.lut_off[k] = static_cast<int>(lut.size());
  if (p.uniform[k] || nb < 4 || nb > 60000) return;
  const T* e = table.data() + p.eoff[k];
  const long double lo = e[0], hi = e[nb];
  if (!std::isfinite(static_cast<double>(lo)) || !std::isfinite(static_cast<double>(hi)) || !(hi > lo)) return;
  auto count_le = [&](long double b) {   // #{effective edges <= b}
    const T* it = std::upper_bound(e, e + nb + 1, b, [](long double v, T edge) { return v < static_cast<long double>(edge); });
    return static_cast<int>(it - e);
  };
  // grow the table until no 3 consecutive cells hold more than 4 edges (then the device needs no search loop)
  int G = 64; while (G < 16 * nb && G < 8192) G <<= 1;
  std::vector<int> cnt;
  int steps = 0;
  // how fine the table gets: until at most `target` edges lie within any 3 cells (or 8192 cells).  Measured on config 5
  // (XH_LUT_STEPS = 4 / 3 / 2 / 1: 3.33 / 3.33 / 3.25 / 3.20 ms): fewer compare-and-advance steps per sample are worth more
  // than the shared memory the larger table takes from the window
  static const int target = [] { const char* e = std::getenv("XH_LUT_STEPS"); const int t = e ? std::atoi(e) : 1; return t >= 1 && t <= 4 ? t : 1; }();
  for (;; G <<= 1) {
    // keep the device's cell index within one cell of the exact one: G * (few ulp) must stay far below 1
    if (static_cast<long double>(G) * (sizeof(T) == 4 ? 6e-7L : 1e-15L) > 0.05L) { if (G > 64) G >>= 1; break; }
    cnt.assign(G + 1, 0);
    for (int c = 0; c <= G; ++c) cnt[c] = std::max(1, count_le(c == G ? hi : lo + c * (hi - lo) / G));
    steps = 0;
    for (int c = 0; c < G; ++c) steps = std::max(steps, cnt[std::min(c + 2, G)] - cnt[std::max(c - 1, 0)]);
    if (steps <= target || G >= 8192) break;
  }
  if (cnt.size() != static_cast<size_t>(G) + 1) {
    cnt.assign(G + 1, 0);
    for (int c = 0; c <= G; ++c) cnt[c] = std::max(1, count_le(c == G ? hi : lo + c * (hi - lo) / G));
    steps = 0;
    for (int c = 0; c < G; ++c) steps = std::max(steps, cnt[std::min(c + 2, G)] - cnt[std::max(c - 1, 0)]);
  }
  const long double inv = G / (hi - lo);
  if (!std::isfinite(static_cast<double>(inv)) || !(inv > 0)) return;
  set_lut_inv<T>(p, k, inv);
  for (int c = 0; c < G; ++c) lut.push_back(static_cast<unsigned short>(std::min(cnt[c] - 1, nb)));
  p.lut_n[k] = G;
  p.lut_steps[k] = (steps >= 1 && steps <= 4) ? steps : (steps == 0 ? 1 : 0);
}

template <typename T>
int prep_var(const double* e, int E, int k, bool force_search, XhkParams& p, std::vector<T>& table) {
  for (int j = 0; j < E; ++j) if (std::isnan(e[j])) return fail(XH_ERR_INVALID, "edges of variable %d contain NaN", k);
  for (int j = 1; j < E; ++j) if (e[j] < e[j - 1]) return fail(XH_ERR_INVALID, "edges of variable %d must increase monotonically", k);
  p.nb[k] = E - 1;
  p.eoff[k] = static_cast<int>(table.size());
  const bool f32 = sizeof(T) == 4;
  for (int j = 0; j < E; ++j) table.push_back(f32 ? static_cast<T>(up32(e[j])) : static_cast<T>(e[j]));
  const T lo = table[p.eoff[k]];
  const T hi = f32 ? static_cast<T>(dn32(e[E - 1])) : static_cast<T>(e[E - 1]);
  T* c = CSel<T>::row(p, k);
  c[XHK_C_LO] = lo; c[XHK_C_HI] = hi;
  p.uniform[k] = 0;
  c[XHK_C_E0] = T(0); c[XHK_C_INV] = T(0); c[XHK_C_DELTA] = T(2); c[XHK_C_OMD] = T(-1); c[XHK_C_CHALF] = T(-1);
  if (force_search) return XH_OK;
  // uniform fast path: usable when the edges are an arithmetic progression up to a small, bounded deviation
  const long double e0 = e[0], eN = e[E - 1];
  if (!std::isfinite(static_cast<double>(e0)) || !std::isfinite(static_cast<double>(eN)) || !(eN > e0)) return XH_OK;
  const long double w = (eN - e0) / (E - 1);
  if (!(w > 0) || !std::isfinite(static_cast<double>(w))) return XH_OK;
  long double dev = 0;  // max deviation of the effective edges from e0 + j*w, in bin units
  for (int j = 0; j < E; ++j) {
    const long double eff = static_cast<long double>(table[p.eoff[k] + j]);
    dev = std::max(dev, std::fabs(eff - (e0 + j * w)) / w);
  }
  const long double u = f32 ? std::ldexp(1.0L, -24) : std::ldexp(1.0L, -53);
  const T e0T = static_cast<T>(static_cast<double>(e0));
  const T invT = static_cast<T>(static_cast<double>(1.0L / w));
  if (!std::isfinite(static_cast<double>(invT)) || invT <= T(0)) return XH_OK;
  const long double d0 = std::fabs(static_cast<long double>(e0T) - e0) / w;
  const long double dinv = std::fabs(static_cast<long double>(invT) * w - 1.0L);  // relative error of inv
  // |t_hat - tau| <= d0 + (E + d0) * (2u + dinv + slack)   (one rounding in the subtraction, one in the product)
  const long double eps_t = d0 + (E + d0) * (2 * u + dinv + u) + 1e-30L;
  const long double delta = 2 * (eps_t + dev) + 4 * u;
  if (!(delta < 0.125L)) return XH_OK;
  // round delta up and 1-delta down in T
  T dT = static_cast<T>(static_cast<double>(delta)); if (static_cast<long double>(dT) < delta) dT = std::nextafter(dT, T(1));
  T oT = static_cast<T>(static_cast<double>(1.0L - delta)); if (static_cast<long double>(oT) > 1.0L - delta) oT = std::nextafter(oT, T(0));
  // one-compare form |frac - 0.5| <= chalf: frac - 0.5 is computed as t - (floor(t) + 0.5), which can round by up to
  // half an ulp of 0.5 when t < 0.25, so chalf gives up 8u on top of delta (rounded down): never certain by mistake
  T hT = static_cast<T>(static_cast<double>(0.5L - delta - 8 * u));
  if (static_cast<long double>(hT) > 0.5L - delta - 8 * u) hT = std::nextafter(hT, T(0));
  c[XHK_C_E0] = e0T; c[XHK_C_INV] = invT; c[XHK_C_DELTA] = dT; c[XHK_C_OMD] = oT; c[XHK_C_CHALF] = hT;
  p.uniform[k] = 1;
  return XH_OK;
}

// int64 data (integers, datetime64/timedelta64 ticks): exact integer compares, binary search only
int prep_var_i64(const int64_t* e, int E, int k, XhkParams& p, std::vector<long long>& table) {
  for (int j = 1; j < E; ++j) if (e[j] < e[j - 1]) return fail(XH_ERR_INVALID, "edges of variable %d must increase monotonically", k);
  p.nb[k] = E - 1;
  p.eoff[k] = static_cast<int>(table.size());
  for (int j = 0; j < E; ++j) table.push_back(e[j]);
  p.ci[k][XHK_C_LO] = e[0]; p.ci[k][XHK_C_HI] = e[E - 1];
  p.uniform[k] = 0; p.lut_n[k] = 0; p.lut_steps[k] = 0; p.lut_off[k] = 0;
  return XH_OK;
}

// L2 prefetch in the vector loops: on, unless XH_PREFETCH=0 (A/B runs)
int prefetch_default() {
  static const int v = [] { const char* e = std::getenv("XH_PREFETCH"); return e ? (std::atoi(e) ? 1 : 0) : 1; }();
  return v;
}

// ------------------------------------------------------------------------------------------ plan + launch
struct Plan {
  XhkParams p;
  XhkLaunch l;
  bool need_window = false;
  int window_budget = 0;
  enum Zero { ZERO_NONE, ZERO_ALL, ZERO_SHARED } zero = ZERO_NONE;
};

// Per-call preparation shared by every block of the call: classification constants and the
// effective-edge table (uploaded once), bin-space geometry.
struct Prep {
  XhkParams base;
  std::vector<unsigned char> edge_host;   // effective edges followed (16-byte aligned) by the lookup tables
  size_t edges_al = 0;                     // shared-memory bytes of edges + tables
  size_t lut_dev_off = 0;                  // byte offset of the tables inside edge_host / the device buffer
  void* dev_edges = nullptr;               // device copy of edge_host (owned by the cache entry)
};

// A prepared set of bin edges, cached per device by edge CONTENT (dtype, counts, search flag, raw edge bytes): a repeat
// call with the same edges skips the host preparation (effective edges, certainty margins, lookup tables) and the
// upload.  The density widths of the last call with these edges are kept next to it.
struct PrepEntry {
  std::vector<unsigned char> key;
  Prep pr;
  unsigned long long id = 0, stamp = 0;
  std::vector<double> widths;              // host copy of what dev_widths holds
  double* dev_widths = nullptr;
  size_t dev_widths_cap = 0;
};

// The verdict of the probe kernel (shared-memory window + fixed-point form of the weights) for one block, cached by
// (edge tables, buffer addresses, shape).  Results never depend on it — a sample outside the window or a weight outside
// the fixed-point form takes a slower exact path — so a stale verdict can only cost time.  That cost is watched:
// k_hist counts the samples that took a slow path, and a verdict whose slow fraction grows well beyond what the
// probing call saw is dropped (the next call probes again).
struct Verdict {
  bool used = false;
  unsigned long long prep_id = 0, stamp = 0;
  const void* data[XH_MAX_VARS]; const void* w = nullptr;
  long long stride[XH_MAX_VARS]; long long wstride = 0, M = 0, N = 0, tile_n = 0;
  int K = 0, dtype = 0, w_dtype = 0, tile_rows = 1, budget = 0, budget32 = 0;
  unsigned flags = 0;
  int state = 0;                 // 1: probe enqueued, host mirror not yet known to be valid; 2: host mirror valid
  unsigned long long last_slow = 0, calls = 0;
  long long samples_since = 0;
  double base_frac = -1.0;       // slow fraction of the first checked interval
  bool stats_pending = false;    // a copy of the counter to the host mirror has been enqueued since the last check
};

int prep_call(const xh_desc* d, Prep& pr) {
  XhkParams& p = pr.base;
  std::memset(&p, 0, sizeof p);
  const int K = d->n_vars;
  const size_t tsz = dsize(d->dtype);
  p.n_vars = K;
  const bool force_search = (d->flags & XH_FLAG_FORCE_SEARCH) != 0;
  std::vector<float> tf; std::vector<double> td; std::vector<long long> ti;
  for (int k = 0; k < K; ++k) {
    int rc = (d->dtype == XH_I64) ? prep_var_i64(d->iedges[k], d->n_edges[k], k, p, ti)
             : (d->dtype == XH_F32) ? prep_var<float>(d->edges[k], d->n_edges[k], k, force_search, p, tf)
                                    : prep_var<double>(d->edges[k], d->n_edges[k], k, force_search, p, td);
    if (rc) return rc;
  }
  long long B = 1;
  for (int k = 0; k < K; ++k) {
    if (B > (1ll << 40) / std::max(1, p.nb[k])) return fail(XH_ERR_INVALID, "bin space too large");
    B *= p.nb[k];
  }
  p.B = B;
  p.all_uniform = (B < 2147483646ll) ? 1 : 0;     // the fast kernel encodes global bins in an int
  // (and takes floor(t) from the mantissa of t + 1.5 * 2^23, which needs bin numbers below 2^21)
  for (int k = 0; k < K; ++k) p.all_uniform = p.all_uniform && p.uniform[k] && p.nb[k] <= (1 << 21);
  long long mul = 1;
  for (int k = K - 1; k >= 0; --k) { p.gmul[k] = mul; mul *= p.nb[k]; }
  p.n_edges_total = static_cast<int>(d->dtype == XH_F32 ? tf.size() : d->dtype == XH_F64 ? td.size() : ti.size());
  const size_t edge_bytes = p.n_edges_total * tsz;
  std::vector<unsigned short> lut;
  for (int k = 0; k < K; ++k) {
    if ((d->flags & XH_FLAG_FORCE_SEARCH) || d->dtype == XH_I64) { p.lut_n[k] = 0; p.lut_off[k] = 0; p.lut_steps[k] = 0; continue; }
    if (d->dtype == XH_F32) build_lut<float>(k, p, tf, lut); else build_lut<double>(k, p, td, lut);
  }
  p.n_lut_total = static_cast<int>(lut.size());
  p.all_branch_free = (p.B < 2147483646ll) ? 1 : 0;
  for (int k = 0; k < K; ++k) p.all_branch_free = p.all_branch_free && (p.uniform[k] || p.lut_steps[k] > 0);
  pr.lut_dev_off = (edge_bytes + 15) & ~static_cast<size_t>(15);
  pr.edge_host.assign(pr.lut_dev_off + lut.size() * 2, 0);
  std::memcpy(pr.edge_host.data(), d->dtype == XH_F32 ? static_cast<const void*>(tf.data())
                                   : d->dtype == XH_F64 ? static_cast<const void*>(td.data()) : static_cast<const void*>(ti.data()), edge_bytes);
  if (!lut.empty()) std::memcpy(pr.edge_host.data() + pr.lut_dev_off, lut.data(), lut.size() * 2);
  pr.edges_al = pr.lut_dev_off + ((lut.size() * 2 + 15) & ~static_cast<size_t>(15));
  return XH_OK;
}

// ---- caches ---------------------------------------------------------------------------------------------------
constexpr size_t kMaxPreps = 16;

std::vector<unsigned char> prep_key(const xh_desc* d) {
  std::vector<unsigned char> k;
  auto put = [&](const void* q, size_t n) { const unsigned char* b = static_cast<const unsigned char*>(q); k.insert(k.end(), b, b + n); };
  const int32_t head[3] = {d->dtype, d->n_vars, (d->flags & XH_FLAG_FORCE_SEARCH) ? 1 : 0};
  put(head, sizeof head);
  put(d->n_edges, sizeof(int32_t) * d->n_vars);
  for (int i = 0; i < d->n_vars; ++i)
    put(d->dtype == XH_I64 ? static_cast<const void*>(d->iedges[i]) : static_cast<const void*>(d->edges[i]), static_cast<size_t>(d->n_edges[i]) * 8);
  return k;
}

void free_prep(PrepEntry* e) {
  if (e->pr.dev_edges) cudaFree(e->pr.dev_edges);
  if (e->dev_widths) cudaFree(e->dev_widths);
  delete e;
}

// Prepared edges for this request: from the cache, or prepared now and uploaded (blocking: first use only).
int lookup_prep(Ctx* c, const xh_desc* d, PrepEntry** out) {
  std::vector<unsigned char> key = prep_key(d);
  for (PrepEntry* e : c->preps)
    if (e->key.size() == key.size() && std::memcmp(e->key.data(), key.data(), key.size()) == 0) { e->stamp = ++c->stamp; *out = e; return XH_OK; }
  PrepEntry* e = new PrepEntry();
  int rc = prep_call(d, e->pr);
  if (rc) { delete e; return rc; }
  const size_t bytes = std::max<size_t>(e->pr.edge_host.size(), 16);
  cudaError_t ce = cudaMalloc(&e->pr.dev_edges, bytes);
  if (ce == cudaSuccess && !e->pr.edge_host.empty()) ce = cudaMemcpy(e->pr.dev_edges, e->pr.edge_host.data(), e->pr.edge_host.size(), cudaMemcpyHostToDevice);
  if (ce != cudaSuccess) { free_prep(e); return fail(ce == cudaErrorMemoryAllocation ? XH_ERR_NOMEM : XH_ERR_CUDA, "edge table upload: %s", cudaGetErrorString(ce)); }
  e->key.swap(key);
  e->id = c->next_prep_id++;
  e->stamp = ++c->stamp;
  if (c->preps.size() >= kMaxPreps) {
    // evict the least recently used entry; kernels of earlier (asynchronous) calls may still read its tables
    size_t lru = 0;
    for (size_t i = 1; i < c->preps.size(); ++i) if (c->preps[i]->stamp < c->preps[lru]->stamp) lru = i;
    cudaDeviceSynchronize();
    for (int i = 0; i < kVerdictSlots; ++i) if (c->verdicts[i].used && c->verdicts[i].prep_id == c->preps[lru]->id) c->verdicts[i].used = false;
    free_prep(c->preps[lru]);
    c->preps.erase(c->preps.begin() + lru);
  }
  c->preps.push_back(e);
  *out = e;
  return XH_OK;
}

Verdict* alloc_verdicts() { return new Verdict[kVerdictSlots]; }

XhkWindow* verdict_dev(Ctx* c, int slot) { return reinterpret_cast<XhkWindow*>(c->vslab_dev + static_cast<size_t>(slot) * kVerdictSlotBytes); }
const XhkWindow* verdict_host(Ctx* c, int slot) { return reinterpret_cast<const XhkWindow*>(c->vslab_host + static_cast<size_t>(slot) * kVerdictSlotBytes); }
unsigned long long* verdict_stats_dev(Ctx* c, int slot) { return reinterpret_cast<unsigned long long*>(c->vslab_dev + static_cast<size_t>(slot) * kVerdictSlotBytes + kVerdictStatsOff); }
unsigned long long verdict_stats_host(Ctx* c, int slot) { return *reinterpret_cast<volatile unsigned long long*>(c->vslab_host + static_cast<size_t>(slot) * kVerdictSlotBytes + kVerdictStatsOff); }

// Launch plan of one block whose data/weights/out pointers are DEVICE pointers.
// fx32 = true plans the k_hist<W = 3> sibling of an fp32-weighted block (4 bytes per shared bin instead of 8).
int plan_block(Ctx* c, const Prep& pr, const xh_desc* d, cudaStream_t stream, Plan& pl, int tile_rows = 1, long long tile_n = 0,
               bool fx32 = false, bool packed = false) {
  XhkParams& p = pl.p;
  p = pr.base;
  p.tile_rows = tile_rows; p.tile_n = static_cast<int>(tile_n); p.tile_magic = 1; p.tile_shift = 0;
  if (tile_rows > 1) {
    // exact n / tile_n for n < 2^31: q = (umulhi(n, magic) + n) >> shift
    int sh = 0; while ((1ll << sh) < tile_n) ++sh;
    p.tile_shift = sh;
    p.tile_magic = static_cast<unsigned>(((1ull << 32) * ((1ull << sh) - static_cast<unsigned long long>(tile_n))) / static_cast<unsigned long long>(tile_n) + 1ull);
    p.B = pr.base.B * tile_rows;    // the launch addresses whole tiles: tile_rows rows of bins each
  }
  const int K = d->n_vars;
  p.M = d->n_rows; p.N = d->n_cols;
  for (int k = 0; k < K; ++k) { p.data[k] = d->data[k]; p.stride[k] = d->row_stride[k]; }
  p.w = d->weights; p.wstride = d->w_row_stride;
  p.out = d->out;
  p.window = c->window;
  p.edges = pr.dev_edges;
  p.lut = reinterpret_cast<const unsigned short*>(static_cast<const unsigned char*>(pr.dev_edges) + pr.lut_dev_off);
  p.stats = c->dummy_stats;
  const long long B = p.B;

  // shared-memory budget
  const size_t edges_al = pr.edges_al;
  if (packed) {
    // counts packed two to a shared word (k_hist<W = 4>): the whole bin space, half the bytes; only planned when it fits
    const long long budget = static_cast<long long>(c->smem_optin) - XHK_STATIC_SMEM - static_cast<long long>(pr.edges_al);
    const size_t smem_pk = pr.edges_al + (static_cast<size_t>((B + 1) / 2) + 32) * 4;
    if (static_cast<long long>(smem_pk) > budget + static_cast<long long>(pr.edges_al)) return fail(XH_ERR_UNSUPPORTED, "packed counts do not fit");
    p.hist_mode = XHK_FULL; p.hist_capacity = static_cast<int>(B);
    const long long total = p.M * p.N;
    int grid = c->sm_count;
    const long long min_per_cta = 4096;
    if (total < static_cast<long long>(grid) * min_per_cta) grid = static_cast<int>(std::max<long long>(1, (total + min_per_cta - 1) / min_per_cta));
    if (p.M >= 16ll * grid) p.partition = XHK_PART_ROWS;
    else {
      p.partition = XHK_PART_SAMPLES;
      long long per = (total + grid - 1) / grid;
      per = (per + 1023) / 1024 * 1024;
      p.per_cta = per;
      grid = static_cast<int>((total + per - 1) / per);
    }
    p.fx_vbits = 24; p.w_dtype = XH_NONE; p.store_owned_rows = 0; p.prefetch = prefetch_default();
    pl.need_window = false;
    pl.zero = (d->flags & XH_FLAG_NO_ZERO) ? Plan::ZERO_NONE : Plan::ZERO_ALL;
    pl.l.dtype = d->dtype; pl.l.w_dtype = 4; pl.l.grid = grid; pl.l.threads = XHK_THREADS; pl.l.smem_bytes = smem_pk; pl.l.stream = stream;
    return XH_OK;
  }
  const size_t item = (d->w_dtype == XH_NONE || fx32) ? 4 : 8;
  const long long budget1 = static_cast<long long>(c->smem_optin) - XHK_STATIC_SMEM - static_cast<long long>(edges_al);  // 1 CTA / SM
  if (budget1 < 0) return fail(XH_ERR_UNSUPPORTED, "bin edges (%zu bytes) do not fit in shared memory", pr.edge_host.size());
  const long long cap1 = budget1 / static_cast<long long>(item) - 32;   // 32 trash slots (see k_hist)
  int mode;
  if (d->flags & XH_FLAG_FORCE_GLOBAL) mode = XHK_GLOBAL;
  else if (B <= cap1 && !(d->flags & XH_FLAG_FORCE_WINDOW)) mode = XHK_FULL;
  else mode = XHK_WINDOW;
  long long marg_bytes = 0; for (int k = 0; k < K; ++k) marg_bytes += 4ll * p.nb[k];
  if (mode == XHK_WINDOW && (static_cast<long long>(edges_al) + marg_bytes > c->smem_optin - 256 || cap1 < 1)) mode = XHK_GLOBAL;
  p.hist_mode = mode;

  int ctas_per_sm = 1, threads = XHK_THREADS;   // k_hist is compiled for <= XHK_THREADS threads per CTA
  size_t smem = edges_al + 32 * item;           // XHK_GLOBAL: only the trash slots
  if (mode == XHK_FULL) {
    smem = edges_al + static_cast<size_t>(B + 32) * item;
    // two 512-thread CTAs per SM when both histograms fit (a flush of one overlaps the stream of the other)
    if (2 * (smem + 1024 + XHK_STATIC_SMEM) <= static_cast<size_t>(c->smem_per_sm)) { ctas_per_sm = 2; threads = XHK_THREADS / 2; }
    p.hist_capacity = static_cast<int>(B);
  } else if (mode == XHK_WINDOW) {
    long long cap = cap1;
    if (d->flags & XH_FLAG_FORCE_WINDOW) cap = std::max<long long>(1, std::min<long long>(cap1, B / 3));
    pl.need_window = true; pl.window_budget = static_cast<int>(std::min<long long>(cap, 1ll << 30));
    smem = edges_al + static_cast<size_t>(pl.window_budget + 32) * item;
    p.hist_capacity = pl.window_budget;
  } else {
    ctas_per_sm = 2; threads = XHK_THREADS / 2;
  }
  p.prefetch = prefetch_default();
  const long long total = p.M * p.N;
  int grid = c->sm_count * ctas_per_sm;
  const long long min_per_cta = 4096;
  if (total < static_cast<long long>(grid) * min_per_cta) grid = static_cast<int>(std::max<long long>(1, (total + min_per_cta - 1) / min_per_cta));
  if (p.M >= 16ll * grid) p.partition = XHK_PART_ROWS;
  else {
    p.partition = XHK_PART_SAMPLES;
    long long per = (total + grid - 1) / grid;
    per = (per + 1023) / 1024 * 1024;
    p.per_cta = per;
    grid = static_cast<int>((total + per - 1) / per);
  }
  // fixed-point headroom: at most A adds reach one bin between two flushes of a CTA
  {
    long long A = std::min<long long>(p.N, 1ll << 30);
    if (p.partition == XHK_PART_SAMPLES) A = std::min<long long>(A, p.per_cta);
    int lg = 0; while ((1ll << lg) < A) ++lg;
    p.fx_vbits = std::max(24, std::min(50, 62 - lg));
    p.w_dtype = d->w_dtype;
    if (d->w_dtype != XH_NONE && mode != XHK_GLOBAL && !pl.need_window) {
      pl.need_window = true;                                   // FULL mode still needs the weight probe
      pl.window_budget = static_cast<int>(std::min<long long>(B, 1ll << 30));
    }
  }
  const bool no_zero = (d->flags & XH_FLAG_NO_ZERO) != 0;
  p.store_owned_rows = (mode == XHK_FULL && !no_zero && !(p.M > 1 && p.N > (1ll << 30))) ? 1 : 0;
  if (no_zero) pl.zero = Plan::ZERO_NONE;
  else if (p.store_owned_rows) pl.zero = (p.partition == XHK_PART_ROWS) ? Plan::ZERO_NONE : Plan::ZERO_SHARED;
  else pl.zero = Plan::ZERO_ALL;
  pl.l.dtype = d->dtype; pl.l.w_dtype = fx32 ? 3 : d->w_dtype; pl.l.grid = grid; pl.l.threads = threads; pl.l.smem_bytes = smem; pl.l.stream = stream;
  return XH_OK;
}

// ---- probe verdicts ---------------------------------------------------------------------------------------------
bool verdict_matches(const Verdict& v, const PrepEntry& pe, const xh_desc* d, int tile_rows, long long tile_n, int budget, int budget32) {
  if (!v.used || v.prep_id != pe.id || v.K != d->n_vars || v.dtype != d->dtype || v.w_dtype != d->w_dtype || v.M != d->n_rows || v.N != d->n_cols ||
      v.w != d->weights || v.wstride != d->w_row_stride || v.tile_rows != tile_rows || v.tile_n != tile_n || v.budget != budget || v.budget32 != budget32 ||
      v.flags != (d->flags & (XH_FLAG_FORCE_GLOBAL | XH_FLAG_FORCE_WINDOW | XH_FLAG_NO_FX32 | XH_FLAG_NO_ZERO)))
    return false;
  for (int k = 0; k < d->n_vars; ++k) if (v.data[k] != d->data[k] || v.stride[k] != d->row_stride[k]) return false;
  return true;
}

// Slot of the verdict for this block (existing or newly claimed).  *ready: the host knows the verdict (no probe needed).
int find_verdict(Ctx* c, const PrepEntry& pe, const xh_desc* d, int tile_rows, long long tile_n, int budget, int budget32, bool* ready) {
  *ready = false;
  for (int i = 0; i < kVerdictSlots; ++i) {
    Verdict& v = c->verdicts[i];
    if (!verdict_matches(v, pe, d, tile_rows, tile_n, budget, budget32)) continue;
    v.stamp = ++c->stamp;
    if (v.state == 1 && cudaEventQuery(c->vev[i]) == cudaSuccess) v.state = 2;
    if (v.state == 2 && v.stats_pending && cudaEventQuery(c->vev[i]) == cudaSuccess) {
      // self-check: fraction of the samples since the last check that left the fast path
      v.stats_pending = false;
      const unsigned long long cur = verdict_stats_host(c, i);
      if (v.samples_since > 0) {
        const double frac = static_cast<double>(cur - v.last_slow) / static_cast<double>(v.samples_since);
        static const bool dbg = std::getenv("XH_DEBUG_VERDICT") != nullptr;
        if (dbg) {
          const XhkWindow* w = verdict_host(c, i);
          std::fprintf(stderr, "[xh verdict %d] window lo=(%d,%d,%d) len=(%d,%d,%d) fx_mode=%d slow fraction %.4f (base %.4f) over %lld samples\n", i,
                       w->lo[0], w->lo[1], w->lo[2], w->len[0], w->len[1], w->len[2], w->fx_mode, frac, v.base_frac, v.samples_since);
        }
        if (v.base_frac < 0.0) v.base_frac = frac;
        else if (frac > 2.0 * v.base_frac + 0.02) { v.state = 0; v.base_frac = -1.0; }      // the data changed under the verdict: probe again
        v.last_slow = cur; v.samples_since = 0;
      }
    }
    *ready = v.state == 2;
    return i;
  }
  int pick = -1;
  for (int i = 0; i < kVerdictSlots; ++i) if (!c->verdicts[i].used) { pick = i; break; }
  if (pick < 0) {
    for (int i = 0; i < kVerdictSlots; ++i)
      if (c->verdicts[i].state != 1 && (pick < 0 || c->verdicts[i].stamp < c->verdicts[pick].stamp)) pick = i;
    if (pick < 0) pick = 0;
  }
  Verdict& v = c->verdicts[pick];
  v = Verdict();
  v.used = true; v.prep_id = pe.id; v.stamp = ++c->stamp;
  v.K = d->n_vars; v.dtype = d->dtype; v.w_dtype = d->w_dtype; v.M = d->n_rows; v.N = d->n_cols; v.w = d->weights; v.wstride = d->w_row_stride;
  v.tile_rows = tile_rows; v.tile_n = tile_n; v.budget = budget; v.budget32 = budget32;
  v.flags = d->flags & (XH_FLAG_FORCE_GLOBAL | XH_FLAG_FORCE_WINDOW | XH_FLAG_NO_FX32 | XH_FLAG_NO_ZERO);
  for (int k = 0; k < d->n_vars; ++k) { v.data[k] = d->data[k]; v.stride[k] = d->row_stride[k]; }
  return pick;
}

// Enqueue zero-fill, window selection and the histogram kernel(s) of one planned block.  `sib` is the k_hist<W = 3>
// plan of an fp32-weighted block (4 bytes per shared bin).  On a probing call both kernels are launched and the
// probe's verdict (XhkWindow::fx_mode, read on the device) makes exactly one of them do the work; with a cached
// verdict the host launches only that one.  `cache` = false (host pipeline: every staged chunk is new data in the same
// staging buffers) probes every block.
int enqueue(Ctx* c, const PrepEntry& pe, const xh_desc* d, Plan& pl, Plan* sib, bool cache, int tile_rows, long long tile_n) {
  cudaStream_t s = pl.l.stream;
  const size_t osz = 8;
  const long long total = pl.p.M * pl.p.N;
  bool ready = false;
  int slot = -1;
  if (pl.need_window) {
    if (cache) {
      slot = find_verdict(c, pe, d, tile_rows, tile_n, pl.window_budget, sib ? sib->window_budget : 0, &ready);
      pl.p.window = verdict_dev(c, slot); pl.p.stats = verdict_stats_dev(c, slot);
      if (sib) { sib->p.window = pl.p.window; sib->p.stats = pl.p.stats; }
    } else {
      pl.p.window = c->window;
      if (sib) sib->p.window = c->window;
    }
  }
  if (ready && slot >= 0 && c->verdicts[slot].base_frac > 0.02) {        // data that spills a lot: see kernel_mode()
    pl.p.spilly = 1;
    if (sib) sib->p.spilly = 1;
  }
  Plan* run_main = &pl; Plan* run_sib = sib;
  if (ready && sib) {                       // the host knows which form the probe chose
    if (verdict_host(c, slot)->fx_mode == 32) { run_main = nullptr; }
    else { run_sib = nullptr; }
  }
  if (run_main && run_sib) {
    pl.p.fx32_sibling = 1;
    const bool same = sib->zero == pl.zero &&
                      (pl.zero != Plan::ZERO_SHARED || (sib->l.grid == pl.l.grid && sib->p.per_cta == pl.p.per_cta));
    if (!same) pl.zero = Plan::ZERO_ALL;
  } else if (run_main) {
    pl.p.fx32_sibling = 0;
  }
  Plan& z = run_main ? pl : *sib;
  if (z.zero == Plan::ZERO_ALL) CU(cudaMemsetAsync(z.p.out, 0, static_cast<size_t>(z.p.M) * z.p.B * osz, s));
  else if (z.zero == Plan::ZERO_SHARED) CU(xhk_launch_zero_shared_rows(z.p, z.l));
  if (pl.need_window && !ready) {
    if (slot >= 0) {     // a fresh (or dropped) verdict: its slow-path counter starts from zero
      CU(cudaMemsetAsync(verdict_stats_dev(c, slot), 0, 8, s));
      *reinterpret_cast<volatile unsigned long long*>(c->vslab_host + static_cast<size_t>(slot) * kVerdictSlotBytes + kVerdictStatsOff) = 0ull;
      Verdict& v = c->verdicts[slot];
      v.last_slow = 0; v.samples_since = 0; v.calls = 0; v.stats_pending = false; v.base_frac = -1.0;
    }
    const int n_probe = static_cast<int>(std::min<long long>(total, 1 << 13));
    CU(xhk_launch_window(pl.p, pl.l, const_cast<XhkWindow*>(pl.p.window), pl.window_budget, sib ? sib->window_budget : 0, n_probe));
    if (slot >= 0) {
      CU(cudaMemcpyAsync(c->vslab_host + static_cast<size_t>(slot) * kVerdictSlotBytes, pl.p.window, sizeof(XhkWindow), cudaMemcpyDeviceToHost, s));
      CU(cudaEventRecord(c->vev[slot], s));
      c->verdicts[slot].state = 1;
    }
  }
  if (run_sib) CU(xhk_launch_hist(run_sib->p, run_sib->l));
  if (run_main) CU(xhk_launch_hist(pl.p, pl.l));
  if (slot >= 0) {
    Verdict& v = c->verdicts[slot];
    v.samples_since += total;
    ++v.calls;
    // look at the slow-path counter after calls 1, 2, 4, 8, 16 and every 16th from then on
    if (ready && !v.stats_pending && ((v.calls & (v.calls - 1)) == 0 || (v.calls & 15) == 0)) {
      CU(cudaMemcpyAsync(c->vslab_host + static_cast<size_t>(slot) * kVerdictSlotBytes + kVerdictStatsOff, verdict_stats_dev(c, slot), 8, cudaMemcpyDeviceToHost, s));
      CU(cudaEventRecord(c->vev[slot], s));
      v.stats_pending = true;
    }
  }
  return XH_OK;
}

// Slow-path fraction recorded for the cached verdict of this block, or -1 when there is none yet.
double verdict_slow_fraction(Ctx* c, const PrepEntry& pe, const xh_desc* d, int tile_rows, long long tile_n, int budget, int budget32) {
  for (int i = 0; i < kVerdictSlots; ++i) {
    const Verdict& v = c->verdicts[i];
    if (verdict_matches(v, pe, d, tile_rows, tile_n, budget, budget32) && v.state == 2) return v.base_frac;
  }
  return -1.0;
}

// Plan a block and, for fp32 weights, its fx32 sibling; enqueue.
int plan_and_enqueue(Ctx* c, const PrepEntry& pe, const xh_desc* d, cudaStream_t stream, bool cache, int tile_rows = 1, long long tile_n = 0) {
  const Prep& pr = pe.pr;
  Plan pl;
  int rc = plan_block(c, pr, d, stream, pl, tile_rows, tile_n);
  if (rc) return rc;
  // Counts whose bin space does not fit as 4-byte bins but fits as packed 16-bit fields: when the windowed launch was seen to
  // spill (the data is spread over more bins than the window holds), later calls use the packed form — everything in shared
  // memory, no spills.  For data the window holds (spill fraction below 0.5 %) the windowed form stays: its non-returning
  // shared adds are cheaper than the packed form's returning ones (1.25 vs 1.77 ms per 1e9 samples on config 3), while every
  // per cent of spilled samples costs the windowed form about 0.7 ms.
  const bool forced = (d->flags & XH_FLAG_FORCE_PACKED) != 0;
  static const bool no_packed = std::getenv("XH_NO_PACKED") != nullptr;      // (profiling the windowed form on data that spills)
  if (!no_packed && d->w_dtype == XH_NONE && d->dtype != XH_I64 && pr.base.all_uniform && tile_rows == 1 && (pl.p.hist_mode == XHK_WINDOW || forced) &&
      !(d->flags & (XH_FLAG_FORCE_GLOBAL | XH_FLAG_FORCE_WINDOW | XH_FLAG_FORCE_SEARCH))) {
    if (forced || (cache && verdict_slow_fraction(c, pe, d, tile_rows, tile_n, pl.window_budget, 0) > 0.005)) {
      Plan pk;
      if (plan_block(c, pr, d, stream, pk, tile_rows, tile_n, false, true) == XH_OK) return enqueue(c, pe, d, pk, nullptr, false, tile_rows, tile_n);
    }
  }
  Plan sib;
  // The one-limb form pays off through its larger window, i.e. when the 8-byte bins do not all fit; when they do,
  // the two-limb form has no spills and no wraps.  Wraps reach the output through global adds, so the sibling
  // never writes rows with plain stores (a windowed main plan zero-fills the whole output anyway).
  bool have_sib = d->w_dtype == XH_F32 && d->dtype != XH_I64 && pl.p.hist_mode == XHK_WINDOW && !(d->flags & XH_FLAG_NO_FX32);
  if (have_sib) {
    rc = plan_block(c, pr, d, stream, sib, tile_rows, tile_n, true);
    if (rc) return rc;
    have_sib = sib.p.hist_mode != XHK_GLOBAL && sib.need_window;
    sib.p.store_owned_rows = 0;
  }
  return enqueue(c, pe, d, pl, have_sib ? &sib : nullptr, cache, tile_rows, tile_n);
}

int validate(const xh_desc* d) {
  if (!d) return fail(XH_ERR_INVALID, "null descriptor");
  if (d->n_vars < 1 || d->n_vars > XH_MAX_VARS) return fail(XH_ERR_INVALID, "n_vars must be 1..%d", XH_MAX_VARS);
  if (d->dtype != XH_F32 && d->dtype != XH_F64 && d->dtype != XH_I64) return fail(XH_ERR_INVALID, "dtype must be XH_F32, XH_F64 or XH_I64");
  if (d->w_dtype != XH_NONE && d->w_dtype != XH_F32 && d->w_dtype != XH_F64) return fail(XH_ERR_INVALID, "bad w_dtype");
  if ((d->w_dtype != XH_NONE) != (d->weights != nullptr)) return fail(XH_ERR_INVALID, "weights pointer and w_dtype disagree");
  if (d->n_rows < 0 || d->n_cols < 0) return fail(XH_ERR_INVALID, "negative shape");
  if (!d->out) return fail(XH_ERR_INVALID, "out is null");
  for (int k = 0; k < d->n_vars; ++k) {
    if (!(d->dtype == XH_I64 ? static_cast<const void*>(d->iedges[k]) : static_cast<const void*>(d->edges[k])) || d->n_edges[k] < 2)
      return fail(XH_ERR_INVALID, "variable %d needs at least 2 edges", k);
    if (d->n_rows > 0 && d->n_cols > 0 && !d->data[k]) return fail(XH_ERR_INVALID, "data[%d] is null", k);
    if (d->row_stride[k] < 0) return fail(XH_ERR_INVALID, "negative row stride");
    if (reinterpret_cast<uintptr_t>(d->data[k]) % dsize(d->dtype)) return fail(XH_ERR_INVALID, "data[%d] is not element-aligned", k);
  }
  if (d->weights && reinterpret_cast<uintptr_t>(d->weights) % dsize(d->w_dtype)) return fail(XH_ERR_INVALID, "weights not element-aligned");
  if ((d->flags & XH_FLAG_NO_ZERO) && d->out_mem != XH_DEVICE) return fail(XH_ERR_INVALID, "XH_FLAG_NO_ZERO needs a device out");
  if ((d->flags & XH_FLAG_OUT_PINNED) && d->out_mem != XH_HOST) return fail(XH_ERR_INVALID, "XH_FLAG_OUT_PINNED is for a host out");
  if ((d->flags & XH_FLAG_ASYNC) && (d->mem != XH_DEVICE || d->out_mem != XH_DEVICE || d->kernel_ms))
    return fail(XH_ERR_INVALID, "XH_FLAG_ASYNC needs device data, a device out and no kernel_ms");
  if (d->n_inner > 1 && (d->n_rows % d->n_inner) != 0) return fail(XH_ERR_INVALID, "column layout: n_rows must be a multiple of n_inner");
  if (d->n_inner < 0) return fail(XH_ERR_INVALID, "negative n_inner");
  if (d->n_weights < 0 || d->n_weights > XH_MAX_WEIGHTS) return fail(XH_ERR_INVALID, "n_weights must be 0..%d", XH_MAX_WEIGHTS);
  if (d->n_weights > 1) {
    if (!d->weights || d->w_dtype == XH_NONE) return fail(XH_ERR_INVALID, "n_weights > 1 needs weights and w_dtype");
    for (int q = 0; q + 1 < d->n_weights; ++q)
      if (!d->weights_more[q] || reinterpret_cast<uintptr_t>(d->weights_more[q]) % dsize(d->w_dtype)) return fail(XH_ERR_INVALID, "weights_more[%d] is null or misaligned", q);
    if (d->flags & (XH_FLAG_DENSITY | XH_FLAG_NO_ZERO)) return fail(XH_ERR_UNSUPPORTED, "several weight arrays: the density is taken per plane on the caller side, and the planes are zero-filled by the library");
    if (d->n_inner > 1) return fail(XH_ERR_UNSUPPORTED, "several weight arrays are not available in the column layout");
  }
  if (d->flags & XH_FLAG_DENSITY) {
    if (d->flags & XH_FLAG_NO_ZERO) return fail(XH_ERR_INVALID, "XH_FLAG_DENSITY cannot be combined with XH_FLAG_NO_ZERO");
    for (int k = 0; k < d->n_vars; ++k) if (!d->widths[k]) return fail(XH_ERR_INVALID, "XH_FLAG_DENSITY needs widths[%d]", k);
  }
  return XH_OK;
}

long long bins_per_row(const xh_desc* d) { long long B = 1; for (int k = 0; k < d->n_vars; ++k) B *= (d->n_edges[k] - 1); return B; }

int ensure_stage(Ctx* c, size_t bytes_per_slot) {
  if (bytes_per_slot <= c->stage_cap) return XH_OK;
  for (int i = 0; i < 2; ++i) { if (c->stage[i]) cudaFree(c->stage[i]); c->stage[i] = nullptr; }
  c->stage_cap = 0;
  for (int i = 0; i < 2; ++i) CU(cudaMalloc(&c->stage[i], bytes_per_slot));
  c->stage_cap = bytes_per_slot;
  return XH_OK;
}

// device-resident block: everything on `stream`
// Rows per tile for blocks of many short rows (1 = keep one row per segment).  A tile is handled like one long
// row with tile_rows * B bins, so a CTA pays its barriers and its flush once per ~16K samples instead of once per
// row, and the flush is one coalesced run of plain stores.
int choose_tile_rows(Ctx* c, const Prep& pr, const xh_desc* d) {
  const long long M = d->n_rows, N = d->n_cols, B = pr.base.B;
  if (M < 2 || N < 1 || N >= 8192 || (d->flags & (XH_FLAG_FORCE_GLOBAL | XH_FLAG_FORCE_WINDOW | XH_FLAG_NO_ZERO))) return 1;
  for (int k = 0; k < d->n_vars; ++k) if (d->row_stride[k] != N) return 1;   // contiguous rows only
  if (d->weights && d->w_row_stride != N) return 1;
  const long long item = d->w_dtype == XH_NONE ? 4 : 8;
  const long long per_cta = (static_cast<long long>(c->smem_per_sm) / 2 - 1024 - XHK_STATIC_SMEM - static_cast<long long>(pr.edges_al)) / item - 32;
  const long long r_max = std::min<long long>(per_cta / std::max<long long>(B, 1), M);
  const long long target = 16384;
  long long R = std::min<long long>(r_max, (target + N - 1) / N);
  R = std::min<long long>(R, ((1ll << 30) - 1) / N);
  return R >= 2 ? static_cast<int>(R) : 1;
}

// several weight arrays in one pass: k_hist_mw; `planes` = elements between two weight planes of the output
int run_mw_device(Ctx* c, const Prep& pr, const xh_desc* d, cudaStream_t stream, long long planes, bool zero) {
  XhkParams p = pr.base;
  p.M = d->n_rows; p.N = d->n_cols;
  for (int k = 0; k < d->n_vars; ++k) { p.data[k] = d->data[k]; p.stride[k] = d->row_stride[k]; }
  p.w = d->weights; p.wstride = d->w_row_stride; p.out = d->out; p.edges = pr.dev_edges; p.w_dtype = d->w_dtype;
  p.lut = reinterpret_cast<const unsigned short*>(static_cast<const unsigned char*>(pr.dev_edges) + pr.lut_dev_off);
  p.stats = c->dummy_stats; p.prefetch = prefetch_default();
  XhkMultiWeights m = {};
  m.nw = d->n_weights; m.w[0] = d->weights;
  for (int q = 1; q < m.nw; ++q) m.w[q] = d->weights_more[q - 1];
  const size_t hist_bytes = static_cast<size_t>(m.nw) * pr.base.B * 8;
  const long long budget = static_cast<long long>(c->smem_optin) - 64 - static_cast<long long>(pr.edges_al);
  m.use_smem = (static_cast<long long>(hist_bytes) <= budget && !(d->flags & XH_FLAG_FORCE_GLOBAL)) ? 1 : 0;
  m.chunk = 1 << 14;
  m.plane = planes;          // elements between the planes of the output (a row block of a larger output passes the whole M * B)
  if (zero) CU(cudaMemsetAsync(d->out, 0, static_cast<size_t>(m.nw) * p.M * p.B * 8, stream));
  const long long items = p.M * ((p.N + m.chunk - 1) / m.chunk);
  XhkLaunch l; l.dtype = d->dtype; l.w_dtype = d->w_dtype; l.threads = XHK_THREADS; l.stream = stream;
  l.smem_bytes = pr.edges_al + (m.use_smem ? hist_bytes : 0) + 16;
  l.grid = static_cast<int>(std::max<long long>(1, std::min<long long>(items, c->sm_count)));
  CU(xhk_launch_hist_mw(p, m, l));
  return XH_OK;
}

int run_device_block(Ctx* c, const PrepEntry& pe, const xh_desc* d, cudaStream_t stream, bool cache, long long mw_planes = 0) {
  const Prep& pr = pe.pr;
  if (d->n_weights > 1) {
    const long long planes = mw_planes ? mw_planes : d->n_rows * pr.base.B;
    // Staged host chunks (cache == false) take the one-pass kernel: PCIe is the bound there and the samples cross it once.
    // Device-resident inputs take one fused pass PER WEIGHT ARRAY: those kernels add in exact fixed point at ~0.9 of the HBM
    // peak, while the one-pass kernel pays a float64 shared add per weight array and sample and is bound by them, not by the
    // reads it saves (1e9 samples, two fp32 weight arrays, 100 x 100 bins: 5.2 ms against 7.6 ms; XH_FLAG_ONE_PASS for A/B).
    if (!cache || (d->flags & XH_FLAG_ONE_PASS)) return run_mw_device(c, pr, d, stream, planes, !(d->flags & XH_FLAG_NO_ZERO));
    for (int q = 0; q < d->n_weights; ++q) {
      xh_desc b = *d;
      b.n_weights = 0;
      b.weights = q ? d->weights_more[q - 1] : d->weights;
      b.out = static_cast<unsigned char*>(d->out) + static_cast<size_t>(q) * planes * 8;
      int rc = run_device_block(c, pe, &b, stream, cache);
      if (rc) return rc;
    }
    return XH_OK;
  }
  const int R = choose_tile_rows(c, pr, d);
  if (R > 1) {
    const long long M = d->n_rows, N = d->n_cols, tiles = M / R, rem = M % R;
    const size_t tsz = dsize(d->dtype), wsz = dsize(d->w_dtype);
    xh_desc t = *d;
    t.n_rows = tiles; t.n_cols = static_cast<long long>(R) * N;
    for (int k = 0; k < d->n_vars; ++k) t.row_stride[k] = t.n_cols;
    if (d->weights) t.w_row_stride = t.n_cols;
    int rc = plan_and_enqueue(c, pe, &t, stream, cache, R, N);
    if (rc || rem == 0) return rc;
    xh_desc r = *d;                                  // the last M % R rows
    r.n_rows = rem;
    for (int k = 0; k < d->n_vars; ++k) r.data[k] = static_cast<const unsigned char*>(d->data[k]) + static_cast<size_t>(tiles) * R * N * tsz;
    if (d->weights) r.weights = static_cast<const unsigned char*>(d->weights) + static_cast<size_t>(tiles) * R * N * wsz;
    r.out = static_cast<unsigned char*>(d->out) + static_cast<size_t>(tiles) * R * pr.base.B * 8;
    return run_device_block(c, pe, &r, stream, cache);
  }
  return plan_and_enqueue(c, pe, d, stream, cache);
}

// host inputs: double-buffered H2D pipeline feeding device blocks that accumulate into dev_out
int run_host_pipeline(Ctx* c, const PrepEntry& pe, const xh_desc* d, void* dev_out) {
  const int K = d->n_vars;
  const size_t tsz = dsize(d->dtype), wsz = dsize(d->w_dtype);
  const long long M = d->n_rows, N = d->n_cols, B = bins_per_row(d);
  const int nwt = wsz ? (d->n_weights > 1 ? d->n_weights : 1) : 0;       // weight arrays
  const bool mw = nwt > 1;
  const int narr = K + nwt;
  auto arr_ptr = [&](int a) -> const unsigned char* {
    return static_cast<const unsigned char*>(a < K ? d->data[a] : a == K ? d->weights : d->weights_more[a - K - 1]);
  };
  auto arr_stride = [&](int a) -> long long { return a < K ? d->row_stride[a] : d->w_row_stride; };
  auto arr_size = [&](int a) -> size_t { return a < K ? tsz : wsz; };
  const long long chunk = 1ll << 23;  // samples per staged block (32 MiB per fp32 array)
  // broadcast rows (stride 0, M > 1) are copied once
  size_t bc_bytes = 0;
  std::vector<size_t> bc_off(narr, 0);
  for (int a = 0; a < narr; ++a) if (arr_stride(a) == 0 && M > 1) { bc_off[a] = bc_bytes; bc_bytes += (static_cast<size_t>(N) * arr_size(a) + 255) & ~static_cast<size_t>(255); }
  const bool long_rows = (N >= chunk) || M == 1;
  if (bc_bytes && long_rows) bc_bytes = 0;  // long rows: broadcast arrays are re-staged per column chunk
  if (bc_bytes > c->bcast_cap) {
    if (c->bcast) cudaFree(c->bcast);
    c->bcast = nullptr; c->bcast_cap = 0;
    CU(cudaMalloc(&c->bcast, bc_bytes)); c->bcast_cap = bc_bytes;
  }
  const long long blk_samples = long_rows ? std::min(chunk, std::max<long long>(N, 1)) : (std::max<long long>(1, chunk / N) * N);
  size_t slot_bytes = 0;
  std::vector<size_t> slot_off(narr, 0);
  for (int a = 0; a < narr; ++a) { slot_off[a] = slot_bytes; slot_bytes += (static_cast<size_t>(blk_samples) * arr_size(a) + 255) & ~static_cast<size_t>(255); }
  int rc = ensure_stage(c, slot_bytes);
  if (rc) return rc;
  if (long_rows || mw) CU(cudaMemsetAsync(dev_out, 0, static_cast<size_t>(mw ? nwt : 1) * M * B * 8, c->stream));  // column chunks accumulate
  if (bc_bytes) {
    for (int a = 0; a < narr; ++a)
      if (arr_stride(a) == 0 && M > 1)
        CU(cudaMemcpyAsync(static_cast<unsigned char*>(c->bcast) + bc_off[a], arr_ptr(a), static_cast<size_t>(N) * arr_size(a), cudaMemcpyHostToDevice, c->stream));
  }
  int it = 0;
  auto submit = [&](long long r0, long long nrows, long long c0, long long ncols) -> int {
    const int slot = it & 1; ++it;
    unsigned char* base = static_cast<unsigned char*>(c->stage[slot]);
    CU(cudaStreamWaitEvent(c->copy_stream, c->consumed[slot], 0));
    xh_desc b = *d;
    b.mem = XH_DEVICE; b.out_mem = XH_DEVICE; b.kernel_ms = nullptr;
    if (long_rows || mw) b.flags |= XH_FLAG_NO_ZERO;  // (otherwise whole-row blocks own their slice of out and zero/store it themselves)
    b.n_rows = nrows; b.n_cols = ncols;
    b.out = static_cast<unsigned char*>(dev_out) + static_cast<size_t>(r0) * B * 8;
    for (int a = 0; a < narr; ++a) {
      const size_t es = arr_size(a);
      const long long st = arr_stride(a);
      const void* dev_ptr; long long dev_stride;
      if (st == 0 && M > 1 && bc_bytes) {
        dev_ptr = static_cast<unsigned char*>(c->bcast) + bc_off[a] + static_cast<size_t>(c0) * es; dev_stride = 0;
      } else {
        unsigned char* dst = base + slot_off[a];
        const unsigned char* src = arr_ptr(a) + (static_cast<size_t>(r0) * st + c0) * es;
        if (st == 0) { CU(cudaMemcpyAsync(dst, src, static_cast<size_t>(ncols) * es, cudaMemcpyHostToDevice, c->copy_stream)); dev_stride = 0; }
        else if (nrows == 1 || st == ncols) { CU(cudaMemcpyAsync(dst, src, static_cast<size_t>(nrows) * ncols * es, cudaMemcpyHostToDevice, c->copy_stream)); dev_stride = ncols; }
        else { CU(cudaMemcpy2DAsync(dst, static_cast<size_t>(ncols) * es, src, static_cast<size_t>(st) * es, static_cast<size_t>(ncols) * es, nrows, cudaMemcpyHostToDevice, c->copy_stream)); dev_stride = ncols; }
        dev_ptr = dst;
      }
      if (a < K) { b.data[a] = dev_ptr; b.row_stride[a] = dev_stride; }
      else if (a == K) { b.weights = dev_ptr; b.w_row_stride = dev_stride; }
      else { b.weights_more[a - K - 1] = dev_ptr; }        // (all weight arrays share one addressing, so the same dev_stride)
    }
    CU(cudaEventRecord(c->copied[slot], c->copy_stream));
    CU(cudaStreamWaitEvent(c->stream, c->copied[slot], 0));
    int rc2 = run_device_block(c, pe, &b, c->stream, false, mw ? M * B : 0);   // new data in the same staging slots: probe every chunk
    if (rc2) return rc2;
    CU(cudaEventRecord(c->consumed[slot], c->stream));
    return XH_OK;
  };
  if (long_rows) {
    for (long long r = 0; r < M; ++r)
      for (long long c0 = 0; c0 < N; c0 += blk_samples) { rc = submit(r, 1, c0, std::min(blk_samples, N - c0)); if (rc) return rc; }
  } else {
    const long long rows_per = blk_samples / N;
    for (long long r0 = 0; r0 < M; r0 += rows_per) { rc = submit(r0, std::min(rows_per, M - r0), 0, N); if (rc) return rc; }
  }
  return XH_OK;
}

// ---- column layout (leading axes reduced): k_hist_cols ----------------------------------------------------
// Threads per CTA = columns per CTA (tm): each thread keeps a private histogram of B bins in shared memory.
int cols_tile(Ctx* c, const Prep& pr, const xh_desc* d, int* tm_out, size_t* smem_out) {
  const long long B = pr.base.B;
  const size_t item = d->w_dtype == XH_NONE ? 4 : 8;
  const long long budget = static_cast<long long>(c->smem_optin) - XHK_STATIC_SMEM - static_cast<long long>(pr.edges_al);
  long long tm = budget / static_cast<long long>(B * item);
  tm = std::min<long long>(tm / 32 * 32, XHK_THREADS);
  if (tm < 32) return fail(XH_ERR_UNSUPPORTED, "column layout: %lld bins per column do not fit a per-thread shared histogram", B);
  // small problems: do not spread a few columns over a 1024-thread CTA
  while (tm > 128 && d->n_inner <= tm / 2) tm /= 2;
  *tm_out = static_cast<int>(tm);
  *smem_out = pr.edges_al + static_cast<size_t>(B) * tm * item;
  return XH_OK;
}

int run_cols_device(Ctx* c, const Prep& pr, const xh_desc* d, cudaStream_t stream, bool accumulate) {
  int tm = 0; size_t smem = 0;
  int rc = cols_tile(c, pr, d, &tm, &smem);
  if (rc) return rc;
  XhkParams p = pr.base;
  p.M = d->n_rows; p.N = d->n_cols;
  for (int k = 0; k < d->n_vars; ++k) p.data[k] = d->data[k];
  p.w = d->weights; p.out = d->out; p.edges = pr.dev_edges; p.w_dtype = d->w_dtype;
  p.lut = reinterpret_cast<const unsigned short*>(static_cast<const unsigned char*>(pr.dev_edges) + pr.lut_dev_off);
  p.stats = c->dummy_stats; p.prefetch = prefetch_default();
  const long long inner = d->n_inner, outer = d->n_rows / inner;
  const long long tiles = ((inner + tm - 1) / tm) * outer;
  int nsplit = 1;
  if (tiles < 2ll * c->sm_count) nsplit = static_cast<int>(std::min<long long>((2ll * c->sm_count + tiles - 1) / tiles, std::max<long long>(1, d->n_cols / 256)));
  const bool acc = accumulate || nsplit > 1;
  if (acc && !accumulate) CU(cudaMemsetAsync(d->out, 0, static_cast<size_t>(d->n_rows) * pr.base.B * 8, stream));
  XhkLaunch l; l.dtype = d->dtype; l.w_dtype = d->w_dtype; l.grid = 0; l.threads = tm; l.smem_bytes = smem; l.stream = stream;
  CU(xhk_launch_hist_cols(p, l, inner, tm, nsplit, acc ? 1 : 0));
  return XH_OK;
}

// host inputs in column layout: stage slabs of the reduced axis (all inner columns of n0..n1) and accumulate
int run_cols_host_pipeline(Ctx* c, const Prep& pr, const xh_desc* d, void* dev_out) {
  const int K = d->n_vars;
  const size_t tsz = dsize(d->dtype), wsz = dsize(d->w_dtype);
  const long long inner = d->n_inner, outer = d->n_rows / inner, N = d->n_cols, B = pr.base.B;
  const int narr = K + (wsz ? 1 : 0);
  const long long chunk = 1ll << 23;
  const long long rows_per = std::max<long long>(1, std::min<long long>(N, chunk / inner));
  size_t slot_bytes = 0; std::vector<size_t> off(narr, 0);
  for (int a = 0; a < narr; ++a) { off[a] = slot_bytes; slot_bytes += (static_cast<size_t>(rows_per) * inner * (a < K ? tsz : wsz) + 255) & ~static_cast<size_t>(255); }
  int rc = ensure_stage(c, slot_bytes);
  if (rc) return rc;
  CU(cudaMemsetAsync(dev_out, 0, static_cast<size_t>(d->n_rows) * B * 8, c->stream));
  int it = 0;
  for (long long a = 0; a < outer; ++a) {
    for (long long n0 = 0; n0 < N; n0 += rows_per) {
      const long long nc = std::min(rows_per, N - n0);
      const int slot = it & 1; ++it;
      unsigned char* base = static_cast<unsigned char*>(c->stage[slot]);
      CU(cudaStreamWaitEvent(c->copy_stream, c->consumed[slot], 0));
      xh_desc b = *d;
      b.mem = XH_DEVICE; b.out_mem = XH_DEVICE; b.kernel_ms = nullptr;
      b.n_rows = inner; b.n_cols = nc; b.n_inner = inner;
      b.out = static_cast<unsigned char*>(dev_out) + static_cast<size_t>(a) * inner * B * 8;
      for (int q = 0; q < narr; ++q) {
        const size_t es = q < K ? tsz : wsz;
        const unsigned char* src = static_cast<const unsigned char*>(q < K ? d->data[q] : d->weights) + (static_cast<size_t>(a) * N + n0) * inner * es;
        CU(cudaMemcpyAsync(base + off[q], src, static_cast<size_t>(nc) * inner * es, cudaMemcpyHostToDevice, c->copy_stream));
        if (q < K) b.data[q] = base + off[q]; else b.weights = base + off[q];
      }
      CU(cudaEventRecord(c->copied[slot], c->copy_stream));
      CU(cudaStreamWaitEvent(c->stream, c->copied[slot], 0));
      rc = run_cols_device(c, pr, &b, c->stream, true);
      if (rc) return rc;
      CU(cudaEventRecord(c->consumed[slot], c->stream));
    }
  }
  return XH_OK;
}

constexpr size_t kPinnedResultMax = 8u << 20;

// pinned, device-mapped landing buffer for small host results
bool ensure_outpin(Ctx* c, size_t bytes) {
  if (bytes <= c->outpin_cap) return true;
  if (c->outpin) { cudaDeviceSynchronize(); cudaFreeHost(c->outpin); }
  c->outpin = nullptr; c->outpin_cap = 0;
  const size_t cap = std::max<size_t>(bytes, 1u << 20);
  if (cudaHostAlloc(&c->outpin, cap, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) { c->outpin = nullptr; cudaGetLastError(); return false; }
  c->outpin_cap = cap;
  return true;
}

int allreduce_in_place(Ctx* c, void* buf, size_t count, bool f64, cudaStream_t s) {
  NC(g_nccl.AllReduce(buf, buf, count, f64 ? kNcclFloat64 : kNcclInt64, kNcclSum, c->comm, s));
  return XH_OK;
}

// ---- peer-memory reduction -------------------------------------------------------------------------------------
constexpr size_t kPeerMaxBytes = 16u << 20;     // larger partials are bandwidth-bound: NCCL's ring / NVLS does those

void peer_release(Ctx* c) {
  Ctx::Peer& pc = c->peer;
  for (int r = 0; r < XHK_MAX_PEERS; ++r) { if (pc.mapped[r] && pc.mapped[r] != pc.local) cudaIpcCloseMemHandle(pc.mapped[r]); pc.mapped[r] = nullptr; }
  if (pc.local) cudaFree(pc.local);
  pc.local = nullptr; pc.cap = 0; pc.seq = 0;
}

// Collective over the ranks of c->comm: (re)create the symmetric buffers with `bytes` per slot and map every peer's.
// Every rank takes the same decisions (same `bytes`, agreement on success through an all-reduce), so either all ranks
// end up with state 1 or all with state -1.
int peer_setup(Ctx* c, size_t bytes, cudaStream_t s) {
  Ctx::Peer& pc = c->peer;
  const int n = c->comm_ranks;
  int ok = (n >= 2 && n <= XHK_MAX_PEERS && !std::getenv("XH_NO_P2P")) ? 1 : 0;
  CU(cudaStreamSynchronize(s));
  int* dflag = nullptr; unsigned char* dh = nullptr;
  CU(cudaMalloc(&dflag, sizeof(int)));
  CU(cudaMalloc(&dh, static_cast<size_t>(XHK_MAX_PEERS + 1) * sizeof(cudaIpcMemHandle_t)));
  auto agree = [&](int mine, int* all) -> int {       // min over ranks (also a barrier: nobody still reads the old buffers)
    CU(cudaMemcpyAsync(dflag, &mine, sizeof(int), cudaMemcpyHostToDevice, s));
    NC(g_nccl.AllReduce(dflag, dflag, 1, kNcclInt32, kNcclMin, c->comm, s));
    CU(cudaMemcpyAsync(all, dflag, sizeof(int), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return XH_OK;
  };
  int all = 0;
  int rc = agree(ok, &all);
  if (rc == XH_OK && all) {
    peer_release(c);
    const size_t cap = std::max<size_t>((bytes * 2 + 4095) & ~static_cast<size_t>(4095), 1u << 20);
    const size_t total = XHK_PEER_FLAG_BYTES + 2 * cap;
    cudaIpcMemHandle_t mine_h; std::memset(&mine_h, 0, sizeof mine_h);
    ok = cudaMalloc(&pc.local, total) == cudaSuccess && cudaMemset(pc.local, 0, total) == cudaSuccess &&
         cudaIpcGetMemHandle(&mine_h, pc.local) == cudaSuccess;
    if (!ok) cudaGetLastError();
    std::vector<cudaIpcMemHandle_t> hs(n);
    CU(cudaMemcpyAsync(dh, &mine_h, sizeof mine_h, cudaMemcpyHostToDevice, s));
    NC(g_nccl.AllGather(dh, dh + sizeof(cudaIpcMemHandle_t), sizeof(cudaIpcMemHandle_t), kNcclInt8, c->comm, s));
    CU(cudaMemcpyAsync(hs.data(), dh + sizeof(cudaIpcMemHandle_t), n * sizeof(cudaIpcMemHandle_t), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    rc = agree(ok, &all);                               // every rank has a buffer and a handle?
    if (rc == XH_OK && all) {
      for (int r = 0; r < n && ok; ++r) {
        if (r == c->comm_rank) { pc.mapped[r] = pc.local; continue; }
        void* q = nullptr;
        if (cudaIpcOpenMemHandle(&q, hs[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); }
        pc.mapped[r] = static_cast<unsigned char*>(q);
      }
      rc = agree(ok, &all);
      if (rc == XH_OK && all) pc.cap = cap;
    }
  }
  cudaFree(dflag); cudaFree(dh);
  if (rc) return rc;
  if (!all) { peer_release(c); pc.state = -1; } else pc.state = 1;
  return XH_OK;
}

// Where the histogram kernels of an XH_FLAG_ALLREDUCE call should write their partial: this rank's slot of the
// symmetric buffer when the peer path serves the call, else nullptr (NCCL reduces dev_out in place).
int peer_slot_for_call(Ctx* c, size_t bytes, cudaStream_t s, void** slot) {
  *slot = nullptr;
  Ctx::Peer& pc = c->peer;
  if (c->comm_ranks < 2 || bytes > kPeerMaxBytes || pc.state < 0) return XH_OK;
  if (pc.state == 0 || bytes > pc.cap) { int rc = peer_setup(c, bytes, s); if (rc) return rc; }
  if (pc.state != 1) return XH_OK;
  *slot = pc.local + XHK_PEER_FLAG_BYTES + ((pc.seq + 1) & 1) * pc.cap;
  return XH_OK;
}

int peer_allreduce(Ctx* c, size_t count, bool f64, void* out, cudaStream_t s) {
  Ctx::Peer& pc = c->peer;
  ++pc.seq;
  XhkPeerArgs a = {};
  a.n = c->comm_ranks; a.rank = c->comm_rank; a.seq = pc.seq; a.count = static_cast<long long>(count); a.out = out;
  for (int r = 0; r < a.n; ++r) {
    a.flags[r] = reinterpret_cast<unsigned long long*>(pc.mapped[r]);
    a.slot[r] = pc.mapped[r] + XHK_PEER_FLAG_BYTES + (pc.seq & 1) * pc.cap;
  }
  CU(xhk_launch_peer_allreduce(a, f64 ? 1 : 0, s));
  return XH_OK;
}

int hist_locked(Ctx* c, const xh_desc* d) {
  CU(cudaSetDevice(c->device));
  const long long M = d->n_rows, N = d->n_cols, B = bins_per_row(d);
  const int nw = d->n_weights > 1 ? d->n_weights : 1;
  const size_t out_bytes = static_cast<size_t>(nw) * M * B * 8;
  void* dev_out = d->out;
  if (d->out_mem == XH_HOST && out_bytes) {
    if (out_bytes > c->outbuf_cap) {
      if (c->outbuf) cudaFree(c->outbuf);
      c->outbuf = nullptr; c->outbuf_cap = 0;
      CU(cudaMalloc(&c->outbuf, out_bytes));
      c->outbuf_cap = out_bytes;
    }
    dev_out = c->outbuf;
  }
  PrepEntry* pe = nullptr;
  int rc = lookup_prep(c, d, &pe);
  if (rc) return rc;
  const Prep& pr = pe->pr;
  phase_mark(0);
  cudaStream_t s = (d->mem == XH_DEVICE && d->stream) ? static_cast<cudaStream_t>(d->stream) : c->stream;
  const bool async = (d->flags & XH_FLAG_ASYNC) != 0;
  bool density_landed = false;       // the density kernel wrote the finished result into the pinned landing buffer
  void* final_out = dev_out;         // where the (reduced) histogram ends up; dev_out is where the kernels of this rank write
  bool via_peers = false;
  if ((d->flags & XH_FLAG_ALLREDUCE) && out_bytes) {
    if (!c->comm) return fail(XH_ERR_NCCL, "XH_FLAG_ALLREDUCE: no communicator on device %d (call xh_comm_init_rank first)", c->device);
    void* slot = nullptr;
    rc = peer_slot_for_call(c, out_bytes, s, &slot);
    if (rc) return rc;
    if (slot) { dev_out = slot; via_peers = true; }
  }
  if (d->kernel_ms) cudaEventRecord(c->ev0, s);
  if (M == 0 || N == 0 || out_bytes == 0) {
    if (out_bytes && !(d->flags & XH_FLAG_NO_ZERO)) { cudaError_t e = cudaMemsetAsync(dev_out, 0, out_bytes, s); if (e != cudaSuccess) rc = fail(XH_ERR_CUDA, "memset: %s", cudaGetErrorString(e)); }
  } else if (d->n_inner > 1) {
    if (d->mem == XH_DEVICE) {
      xh_desc b = *d; b.out = dev_out; b.out_mem = XH_DEVICE;
      rc = run_cols_device(c, pr, &b, s, (d->flags & XH_FLAG_NO_ZERO) != 0);
    } else {
      rc = run_cols_host_pipeline(c, pr, d, dev_out);
    }
  } else if (d->mem == XH_DEVICE) {
    xh_desc b = *d; b.out = dev_out; b.out_mem = XH_DEVICE;
    rc = run_device_block(c, *pe, &b, s, true);
  } else {
    rc = run_host_pipeline(c, *pe, d, dev_out);
  }
  if (rc == XH_OK && (d->flags & XH_FLAG_ALLREDUCE) && out_bytes) {
    // partial histograms of the ranks -> global histogram on the same stream (no host round trip): one peer-memory
    // kernel for small histograms, ncclAllReduce in place otherwise
    if (via_peers) rc = peer_allreduce(c, static_cast<size_t>(nw * M * B), d->w_dtype != XH_NONE, final_out, s);
    else rc = allreduce_in_place(c, dev_out, static_cast<size_t>(nw * M * B), d->w_dtype != XH_NONE, s);
    dev_out = final_out;
  }
  if (rc == XH_OK && (d->flags & XH_FLAG_DENSITY) && out_bytes) {
    // core.py:444-462 on the device, in place: counts / bin areas / row sums
    int nb[XH_MAX_VARS], f32[XH_MAX_VARS];
    size_t nw = 0;
    for (int k = 0; k < d->n_vars; ++k) { nb[k] = d->n_edges[k] - 1; f32[k] = d->widths_f32[k]; nw += nb[k]; }
    bool same = pe->widths.size() == nw && pe->dev_widths;
    size_t o = 0;
    for (int k = 0; k < d->n_vars && same; ++k) { same = std::memcmp(pe->widths.data() + o, d->widths[k], sizeof(double) * nb[k]) == 0; o += nb[k]; }
    if (!same) {       // first density call with these edges (or other widths): upload, blocking
      pe->widths.clear();
      for (int k = 0; k < d->n_vars; ++k) pe->widths.insert(pe->widths.end(), d->widths[k], d->widths[k] + nb[k]);
      if (nw > pe->dev_widths_cap) {
        if (pe->dev_widths) { cudaDeviceSynchronize(); cudaFree(pe->dev_widths); }
        pe->dev_widths = nullptr; pe->dev_widths_cap = 0;
        CU(cudaMalloc(&pe->dev_widths, std::max<size_t>(nw, 64) * sizeof(double)));
        pe->dev_widths_cap = std::max<size_t>(nw, 64);
      }
      CU(cudaStreamSynchronize(s));    // an earlier asynchronous call may still read the old widths
      CU(cudaMemcpy(pe->dev_widths, pe->widths.data(), nw * sizeof(double), cudaMemcpyHostToDevice));
    }
    // a small result for the host: one launch that writes the finished density straight into the pinned landing buffer
    void* landing = (d->flags & XH_FLAG_OUT_PINNED) ? d->out : nullptr;      // the caller's own pinned result memory, or ours
    if (d->out_mem == XH_HOST && out_bytes <= kPinnedResultMax && M <= (1 << 20) && (landing || ensure_outpin(c, out_bytes))) {
      if (!landing) landing = c->outpin;
      CU(xhk_launch_density_small(dev_out, static_cast<double*>(landing), M, B, d->w_dtype == XH_NONE ? 1 : 0, pe->dev_widths, nb, f32, d->n_vars, s));
      density_landed = true;
    } else {
      if (B > 1024 && static_cast<size_t>(M) > c->rowsums_cap) {
        if (c->rowsums) { cudaDeviceSynchronize(); cudaFree(c->rowsums); }
        c->rowsums = nullptr; c->rowsums_cap = 0;
        const size_t cap = std::max<size_t>(static_cast<size_t>(M), 4096);
        CU(cudaMalloc(&c->rowsums, cap * 8));
        c->rowsums_cap = cap;
      }
      CU(xhk_launch_density(dev_out, M, B, d->w_dtype == XH_NONE ? 1 : 0, pe->dev_widths, nb, f32, d->n_vars, c->rowsums, s));
    }
  }
  if (rc == XH_OK && d->kernel_ms) cudaEventRecord(c->ev1, s);
  bool via_pin = density_landed && !(d->flags & XH_FLAG_OUT_PINNED);
  if (rc == XH_OK && d->out_mem == XH_HOST && out_bytes && !density_landed) {
    void* dst = d->out;
    if (!(d->flags & XH_FLAG_OUT_PINNED) && out_bytes <= kPinnedResultMax && ensure_outpin(c, out_bytes)) { dst = c->outpin; via_pin = true; }
    cudaError_t e = cudaMemcpyAsync(dst, dev_out, out_bytes, cudaMemcpyDeviceToHost, s);
    if (e != cudaSuccess) rc = fail(XH_ERR_CUDA, "D2H of the histogram failed: %s", cudaGetErrorString(e));
  }
  phase_mark(1);
  if (async && rc == XH_OK) return XH_OK;       // device in, device out: the result is valid in stream order
  cudaError_t e = cudaStreamSynchronize(s);
  phase_mark(2);
  if (via_pin && e == cudaSuccess && rc == XH_OK) std::memcpy(d->out, c->outpin, out_bytes);
  if (rc == XH_OK && e != cudaSuccess) rc = fail(XH_ERR_CUDA, "histogram kernel failed: %s", cudaGetErrorString(e));
  if (d->mem == XH_HOST) { e = cudaStreamSynchronize(c->copy_stream); if (rc == XH_OK && e != cudaSuccess) rc = fail(XH_ERR_CUDA, "copy stream: %s", cudaGetErrorString(e)); }
  // the stream is idle: every pending probe verdict of this context has reached its host mirror
  for (int i = 0; i < kVerdictSlots; ++i) if (c->verdicts[i].used && c->verdicts[i].state == 1 && cudaEventQuery(c->vev[i]) == cudaSuccess) c->verdicts[i].state = 2;
  if (rc == XH_OK && d->kernel_ms) { float ms = 0; cudaEventElapsedTime(&ms, c->ev0, c->ev1); *d->kernel_ms = ms; }
  return rc;
}

}  // namespace

// ============================================================================================ C-ABI
extern "C" {

int xh_version(void) { return XH_VERSION_MAJOR * 1000 + XH_VERSION_MINOR; }
int xh_desc_size(void) { return static_cast<int>(sizeof(xh_desc)); }

int xh_last_error(char* buf, size_t len) {
  if (!buf || !len) return XH_ERR_INVALID;
  std::snprintf(buf, len, "%s", g_err.c_str());
  return XH_OK;
}

int xh_device_count(int* count) {
  if (!count) return fail(XH_ERR_INVALID, "null");
  int n = 0; cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) { *count = 0; return fail(XH_ERR_NO_DEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e)); }
  *count = n; return XH_OK;
}

int xh_init(int device) { Ctx* c; return get_ctx(device, &c); }

int xh_shutdown(void) {
  std::lock_guard<std::mutex> lk(g_ctx_mu);
  for (auto& kv : g_ctx) {
    Ctx* c = kv.second;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    peer_release(c);
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    if (c->window) cudaFree(c->window);
    if (c->edges) cudaFree(c->edges);
    for (PrepEntry* e : c->preps) free_prep(e);
    delete[] c->verdicts;
    if (c->vslab_dev) cudaFree(c->vslab_dev);
    if (c->vslab_host) cudaFreeHost(c->vslab_host);
    for (int i = 0; i < kVerdictSlots; ++i) cudaEventDestroy(c->vev[i]);
    if (c->minmax) cudaFree(c->minmax);
    for (int i = 0; i < 2; ++i) if (c->stage[i]) cudaFree(c->stage[i]);
    if (c->bcast) cudaFree(c->bcast);
    if (c->flush) cudaFree(c->flush);
    if (c->outbuf) cudaFree(c->outbuf);
    if (c->outpin) cudaFreeHost(c->outpin);
    if (c->widths) cudaFree(c->widths);
    if (c->rowsums) cudaFree(c->rowsums);
    for (int i = 0; i < 2; ++i) { cudaEventDestroy(c->copied[i]); cudaEventDestroy(c->consumed[i]); }
    cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1); cudaEventDestroy(c->tev0); cudaEventDestroy(c->tev1);
    cudaStreamDestroy(c->stream); cudaStreamDestroy(c->copy_stream);
    delete c;
  }
  g_ctx.clear();
  return XH_OK;
}

int xh_device_info(int device, int* sm_count, int* smem_optin_bytes, int64_t* total_mem, int* cc_major, int* cc_minor) {
  int n = 0; cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) return fail(XH_ERR_NO_DEVICE, "no CUDA device: %s", cudaGetErrorString(e));
  cudaDeviceProp pr; CU(cudaGetDeviceProperties(&pr, device));
  if (sm_count) *sm_count = pr.multiProcessorCount;
  if (smem_optin_bytes) *smem_optin_bytes = static_cast<int>(pr.sharedMemPerBlockOptin);
  if (total_mem) *total_mem = static_cast<int64_t>(pr.totalGlobalMem);
  if (cc_major) *cc_major = pr.major;
  if (cc_minor) *cc_minor = pr.minor;
  return XH_OK;
}

int xh_hist(const xh_desc* d) {
  g_t0 = std::chrono::steady_clock::now();
  int rc = validate(d);
  if (rc) return rc;
  Ctx* c; rc = get_ctx(d->device, &c);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(c->mu);
  rc = hist_locked(c, d);
  phase_mark(3);
  return rc;
}

int xh_last_call_phases(double* us4) {
  if (!us4) return fail(XH_ERR_INVALID, "null");
  for (int i = 0; i < 4; ++i) us4[i] = g_phase[i];
  return XH_OK;
}

int xh_hist_multi(const xh_desc* d, const int32_t* devices, int32_t n_dev) {
  int rc = validate(d);
  if (rc) return rc;
  if (!devices || n_dev < 1) return fail(XH_ERR_INVALID, "need at least one device");
  if (d->mem != XH_HOST || d->out_mem != XH_HOST) return fail(XH_ERR_INVALID, "xh_hist_multi takes host data and a host out");
  if (n_dev == 1) { xh_desc b = *d; b.device = devices[0]; return xh_hist(&b); }
  if (d->flags & (XH_FLAG_DENSITY | XH_FLAG_ALLREDUCE)) return fail(XH_ERR_UNSUPPORTED, "xh_hist_multi reduces on its own; the density is taken on the caller side");
  if (d->n_inner > 1) return fail(XH_ERR_UNSUPPORTED, "xh_hist_multi does not take the column layout; shard the kept axis on the caller side");
  if (d->n_weights > 1) return fail(XH_ERR_UNSUPPORTED, "xh_hist_multi does not take several weight arrays");
  const long long M = d->n_rows, N = d->n_cols, B = bins_per_row(d);
  const size_t tsz = dsize(d->dtype), wsz = dsize(d->w_dtype);
  std::vector<Ctx*> ctx(n_dev);
  for (int i = 0; i < n_dev; ++i) { rc = get_ctx(devices[i], &ctx[i]); if (rc) return rc; }
  const bool by_rows = M >= n_dev;
  std::vector<int> rcs(n_dev, XH_OK);
  std::vector<std::string> errs(n_dev);
  std::vector<void*> partial(n_dev, nullptr);
  std::vector<float> times(n_dev, 0.f);
  if (!by_rows) { rc = nccl_load(); if (rc) return rc; }
  // per-device shard, run concurrently (one host thread per GPU; each has its own PCIe link)
  std::vector<std::thread> th;
  for (int i = 0; i < n_dev; ++i) {
    th.emplace_back([&, i]() {
      Ctx* c = ctx[i];
      std::lock_guard<std::mutex> lk(c->mu);
      xh_desc b = *d;
      b.device = devices[i]; b.kernel_ms = d->kernel_ms ? &times[i] : nullptr;
      if (by_rows) {
        const long long r0 = M * i / n_dev, r1 = M * (i + 1) / n_dev;
        b.n_rows = r1 - r0;
        for (int k = 0; k < d->n_vars; ++k) b.data[k] = static_cast<const unsigned char*>(d->data[k]) + static_cast<size_t>(r0) * d->row_stride[k] * tsz;
        if (d->weights) b.weights = static_cast<const unsigned char*>(d->weights) + static_cast<size_t>(r0) * d->w_row_stride * wsz;
        b.out = static_cast<unsigned char*>(d->out) + static_cast<size_t>(r0) * B * 8;  // disjoint host slices: no reduction
        rcs[i] = b.n_rows ? hist_locked(c, &b) : XH_OK;
      } else {
        const long long c0 = N * i / n_dev, c1 = N * (i + 1) / n_dev;
        b.n_cols = c1 - c0;
        for (int k = 0; k < d->n_vars; ++k) b.data[k] = static_cast<const unsigned char*>(d->data[k]) + static_cast<size_t>(c0) * tsz;
        if (d->weights) b.weights = static_cast<const unsigned char*>(d->weights) + static_cast<size_t>(c0) * wsz;
        cudaSetDevice(c->device);
        if (cudaMalloc(&partial[i], static_cast<size_t>(M) * B * 8) != cudaSuccess) { rcs[i] = XH_ERR_NOMEM; errs[i] = "partial histogram allocation failed"; return; }
        b.out = partial[i]; b.out_mem = XH_DEVICE;
        rcs[i] = hist_locked(c, &b);
      }
      if (rcs[i]) errs[i] = g_err;
    });
  }
  for (auto& t : th) t.join();
  for (int i = 0; i < n_dev; ++i) if (rcs[i]) { for (void* q : partial) if (q) cudaFree(q); return fail(rcs[i], "device %d: %s", devices[i], errs[i].c_str()); }
  if (d->kernel_ms) *d->kernel_ms = *std::max_element(times.begin(), times.end());
  if (by_rows) return XH_OK;
  // columns were sharded: sum the partial histograms with one grouped ncclAllReduce over NVLink
  static std::mutex comm_mu;
  static std::map<std::vector<int>, std::vector<void*>> comm_cache;
  std::lock_guard<std::mutex> lk(comm_mu);
  std::vector<int> key(devices, devices + n_dev);
  auto it = comm_cache.find(key);
  if (it == comm_cache.end()) {
    std::vector<void*> comms(n_dev, nullptr);
    NC(g_nccl.CommInitAll(comms.data(), n_dev, key.data()));
    it = comm_cache.emplace(key, comms).first;
  }
  const size_t count = static_cast<size_t>(M) * B;
  const int ty = d->w_dtype == XH_NONE ? kNcclInt64 : kNcclFloat64;
  NC(g_nccl.GroupStart());
  for (int i = 0; i < n_dev; ++i) {
    int r = g_nccl.AllReduce(partial[i], partial[i], count, ty, kNcclSum, it->second[i], ctx[i]->stream);
    if (r != 0) { g_nccl.GroupEnd(); return fail(XH_ERR_NCCL, "ncclAllReduce: %s", g_nccl.GetErrorString(r)); }
  }
  NC(g_nccl.GroupEnd());
  CU(cudaSetDevice(ctx[0]->device));
  CU(cudaMemcpyAsync(d->out, partial[0], count * 8, cudaMemcpyDeviceToHost, ctx[0]->stream));
  for (int i = 0; i < n_dev; ++i) { CU(cudaSetDevice(ctx[i]->device)); CU(cudaStreamSynchronize(ctx[i]->stream)); }
  for (int i = 0; i < n_dev; ++i) { cudaSetDevice(ctx[i]->device); cudaFree(partial[i]); }
  return XH_OK;
}

int xh_minmax(int device, const void* data, int dtype, int mem, int64_t n, double* mn, double* mx) {
  if (!mn || !mx || (dtype != XH_F32 && dtype != XH_F64) || n < 0) return fail(XH_ERR_INVALID, "bad arguments");
  Ctx* c; int rc = get_ctx(device, &c); if (rc) return rc;
  std::lock_guard<std::mutex> lk(c->mu);
  CU(cudaSetDevice(c->device));
  if (n == 0) return fail(XH_ERR_INVALID, "zero-size array has no min/max");
  const size_t es = dsize(dtype);
  double hmn = INFINITY, hmx = -INFINITY; bool nan = false;
  std::vector<double> part(296 * 3);
  auto reduce_dev = [&](const void* dptr, long long cnt) -> int {
    CU(xhk_launch_minmax(dptr, dtype, cnt, c->minmax, c->stream));
    CU(cudaMemcpyAsync(part.data(), c->minmax, part.size() * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < 296; ++i) { hmn = std::min(hmn, part[3 * i]); hmx = std::max(hmx, part[3 * i + 1]); nan = nan || part[3 * i + 2] != 0.0; }
    return XH_OK;
  };
  if (mem == XH_DEVICE) { rc = reduce_dev(data, n); if (rc) return rc; }
  else {
    const long long chunk = 1ll << 24;
    rc = ensure_stage(c, static_cast<size_t>(chunk) * 8); if (rc) return rc;
    for (long long o = 0; o < n; o += chunk) {
      const long long cnt = std::min<long long>(chunk, n - o);
      CU(cudaMemcpyAsync(c->stage[0], static_cast<const unsigned char*>(data) + o * es, cnt * es, cudaMemcpyHostToDevice, c->stream));
      rc = reduce_dev(c->stage[0], cnt); if (rc) return rc;
    }
  }
  if (nan) { *mn = NAN; *mx = NAN; } else { *mn = hmn; *mx = hmx; }
  return XH_OK;
}

int xh_malloc(int device, size_t bytes, void** ptr) {
  if (!ptr) return fail(XH_ERR_INVALID, "null");
  Ctx* c; int rc = get_ctx(device, &c); if (rc) return rc;
  CU(cudaSetDevice(c->device));
  CU(cudaMalloc(ptr, bytes ? bytes : 1));
  return XH_OK;
}
int xh_free(int device, void* ptr) {
  Ctx* c; int rc = get_ctx(device, &c); if (rc) return rc;
  CU(cudaSetDevice(c->device));
  CU(cudaFree(ptr));
  return XH_OK;
}
int xh_host_alloc(size_t bytes, void** ptr) {
  if (!ptr) return fail(XH_ERR_INVALID, "null");
  CU(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocPortable | cudaHostAllocMapped));
  return XH_OK;
}
int xh_host_free(void* ptr) { CU(cudaFreeHost(ptr)); return XH_OK; }

int xh_memcpy(int device, void* dst, const void* src, size_t bytes, int dst_mem, int src_mem) {
  Ctx* c; int rc = get_ctx(device, &c); if (rc) return rc;
  std::lock_guard<std::mutex> lk(c->mu);
  CU(cudaSetDevice(c->device));
  cudaMemcpyKind kind = dst_mem == XH_DEVICE ? (src_mem == XH_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice)
                                             : (src_mem == XH_DEVICE ? cudaMemcpyDeviceToHost : cudaMemcpyHostToHost);
  CU(cudaMemcpyAsync(dst, src, bytes, kind, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return XH_OK;
}
int xh_memset(int device, void* dst, int value, size_t bytes) {
  Ctx* c; int rc = get_ctx(device, &c); if (rc) return rc;
  std::lock_guard<std::mutex> lk(c->mu);
  CU(cudaSetDevice(c->device));
  CU(cudaMemsetAsync(dst, value, bytes, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return XH_OK;
}
int xh_sync(int device) {
  Ctx* c; int rc = get_ctx(device, &c); if (rc) return rc;
  CU(cudaSetDevice(c->device));
  CU(cudaDeviceSynchronize());
  return XH_OK;
}

int xh_stream_wait(int device, void* producer_stream) {
  Ctx* c; int rc = get_ctx(device, &c); if (rc) return rc;
  std::lock_guard<std::mutex> lk(c->mu);
  CU(cudaSetDevice(c->device));
  // (cudaStream_t)1 and (cudaStream_t)2 are the legacy and the per-thread default stream handles
  CU(cudaEventRecord(c->copied[0], static_cast<cudaStream_t>(producer_stream)));
  CU(cudaStreamWaitEvent(c->stream, c->copied[0], 0));
  return XH_OK;
}

int xh_permute(int device, const void* src, void* dst, int elem_size, int ndim, const int64_t* shape, const int32_t* perm) {
  if (!shape || !perm || ndim < 1 || ndim > XH_MAX_VARS || (elem_size != 4 && elem_size != 8)) return fail(XH_ERR_INVALID, "xh_permute: 1..%d axes of 4- or 8-byte elements", XH_MAX_VARS);
  long long shp[XH_MAX_VARS]; int prm[XH_MAX_VARS]; unsigned seen = 0; long long total = 1;
  for (int i = 0; i < ndim; ++i) {
    if (shape[i] < 0 || perm[i] < 0 || perm[i] >= ndim || (seen >> perm[i] & 1u)) return fail(XH_ERR_INVALID, "xh_permute: bad shape or permutation");
    seen |= 1u << perm[i]; shp[i] = shape[i]; prm[i] = perm[i]; total *= shape[i];
  }
  if (total && (!src || !dst)) return fail(XH_ERR_INVALID, "xh_permute: null buffer");
  Ctx* c; int rc = get_ctx(device, &c); if (rc) return rc;
  std::lock_guard<std::mutex> lk(c->mu);
  CU(cudaSetDevice(c->device));
  CU(xhk_launch_permute(src, dst, elem_size, ndim, shp, prm, c->stream));   // stream-ordered before any later xh_hist on the library stream
  return XH_OK;
}

static int fill_impl(int device, void* ptr, int dtype, int64_t n, uint64_t seed, int64_t offset, int normal) {
  if (!ptr || (dtype != XH_F32 && dtype != XH_F64) || n < 0) return fail(XH_ERR_INVALID, "bad arguments");
  Ctx* c; int rc = get_ctx(device, &c); if (rc) return rc;
  std::lock_guard<std::mutex> lk(c->mu);
  CU(cudaSetDevice(c->device));
  CU(xhk_launch_fill(ptr, dtype, n, seed, offset, normal, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return XH_OK;
}
int xh_fill_normal(int device, void* ptr, int dtype, int64_t n, uint64_t seed, int64_t offset) { return fill_impl(device, ptr, dtype, n, seed, offset, 1); }
int xh_fill_uniform(int device, void* ptr, int dtype, int64_t n, uint64_t seed, int64_t offset) { return fill_impl(device, ptr, dtype, n, seed, offset, 0); }

int xh_timer_start(int device) {
  Ctx* c; int rc = get_ctx(device, &c); if (rc) return rc;
  CU(cudaSetDevice(c->device));
  CU(cudaEventRecord(c->tev0, c->stream));
  return XH_OK;
}
int xh_timer_stop(int device, float* ms) {
  Ctx* c; int rc = get_ctx(device, &c); if (rc) return rc;
  CU(cudaSetDevice(c->device));
  CU(cudaEventRecord(c->tev1, c->stream));
  CU(cudaEventSynchronize(c->tev1));
  float t = 0; CU(cudaEventElapsedTime(&t, c->tev0, c->tev1));
  if (ms) *ms = t;
  return XH_OK;
}
int xh_flush_l2(int device) {
  Ctx* c; int rc = get_ctx(device, &c); if (rc) return rc;
  std::lock_guard<std::mutex> lk(c->mu);
  CU(cudaSetDevice(c->device));
  if (!c->flush) { c->flush_bytes = static_cast<size_t>(256) << 20; CU(cudaMalloc(&c->flush, c->flush_bytes)); }
  CU(xhk_launch_flush(c->flush, c->flush_bytes, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return XH_OK;
}

int xh_comm_unique_id(void* id128) {
  if (!id128) return fail(XH_ERR_INVALID, "null");
  int rc = nccl_load(); if (rc) return rc;
  NC(g_nccl.GetUniqueId(id128));
  return XH_OK;
}
int xh_comm_init_rank(int device, const void* id128, int n_ranks, int rank) {
  if (!id128 || n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(XH_ERR_INVALID, "bad arguments");
  int rc = nccl_load(); if (rc) return rc;
  Ctx* c; rc = get_ctx(device, &c); if (rc) return rc;
  std::lock_guard<std::mutex> lk(c->mu);
  CU(cudaSetDevice(c->device));
  if (c->comm) { peer_release(c); c->peer.state = 0; g_nccl.CommDestroy(c->comm); c->comm = nullptr; }
  Id128 id; std::memcpy(id.b, id128, sizeof id.b);
  NC(g_nccl.CommInitRank(&c->comm, n_ranks, id, rank));
  c->comm_ranks = n_ranks; c->comm_rank = rank; c->peer.state = 0;
  return XH_OK;
}
int xh_comm_allreduce(int device, void* dev_buf, int64_t count, int dtype_is_f64) {
  Ctx* c; int rc = get_ctx(device, &c); if (rc) return rc;
  std::lock_guard<std::mutex> lk(c->mu);
  if (!c->comm) return fail(XH_ERR_NCCL, "no communicator: call xh_comm_init_rank first");
  CU(cudaSetDevice(c->device));
  NC(g_nccl.AllReduce(dev_buf, dev_buf, static_cast<size_t>(count), dtype_is_f64 ? kNcclFloat64 : kNcclInt64, kNcclSum, c->comm, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return XH_OK;
}
int xh_comm_destroy(int device) {
  Ctx* c; int rc = get_ctx(device, &c); if (rc) return rc;
  std::lock_guard<std::mutex> lk(c->mu);
  if (c->comm) {
    CU(cudaSetDevice(c->device));
    cudaDeviceSynchronize();
    if (c->peer.state == 1) {
      // peers map this rank's symmetric buffer: release it only when every rank has finished its own kernels and arrived here
      // (a collective, like ncclCommDestroy itself)
      int* flag = nullptr;
      if (cudaMalloc(&flag, sizeof(int)) == cudaSuccess) {
        cudaMemsetAsync(flag, 0, sizeof(int), c->stream);
        g_nccl.AllReduce(flag, flag, 1, kNcclInt32, kNcclSum, c->comm, c->stream);
        cudaStreamSynchronize(c->stream);
        cudaFree(flag);
      }
    }
    peer_release(c); c->peer.state = 0;
    NC(g_nccl.CommDestroy(c->comm)); c->comm = nullptr; c->comm_ranks = 0;
  }
  return XH_OK;
}

}  // extern "C"
