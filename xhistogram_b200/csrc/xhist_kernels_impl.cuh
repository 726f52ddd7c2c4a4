// xhist_kernels_impl.cuh — device templates of the histogram kernels (k_hist, k_window, k_hist_cols).
// Included by one translation unit per data type (xhist_k_f32.cu, xhist_k_f64.cu, xhist_k_i64.cu) so that the
// ~200 template instantiations compile in parallel; see xhist_kernels.cu for the design notes and the launchers.
#pragma once
#include "xhist_kernels.cuh"
#include <type_traits>

#ifndef XH_CHEAP_SIDE_W3
#define XH_CHEAP_SIDE_W3 0
#endif
#ifndef XH_PREFETCH_DIST
#define XH_PREFETCH_DIST 2  // prefetch.global.L2 of the groups this many iterations ahead in the vector loops (0 = off).  The kernels
                            // are latency-bound at 32 warps per SM (one 227 KB CTA) and every register is taken, so the extra
                            // memory-level parallelism has to come without registers: measured on 1e9 samples, 0 / 2 / 4 iterations
                            // ahead: weighted 2.254 / 1.992 / 2.032 ms, counts 1.369 / 1.250 / 1.254 ms
#endif



namespace {

constexpr int kMaxThreads = XHK_THREADS;
constexpr long long kSegCap = 1ll << 30;  // u32 shared counters are flushed at least this often



template <typename T> struct Consts;
template <> struct Consts<float> {
  static __device__ __forceinline__ float get(const XhkParams& p, int k, int i) { return p.cf[k][i]; }
};
template <> struct Consts<double> {
  static __device__ __forceinline__ double get(const XhkParams& p, int k, int i) { return p.cd[k][i]; }
};
template <> struct Consts<long long> {   // int64 data (integers, datetime64 ticks): range limits only
  static __device__ __forceinline__ long long get(const XhkParams& p, int k, int i) { return i < 2 ? p.ci[k][i] : 0ll; }
};

__device__ __forceinline__ int floor_to_int(float t) { return __float2int_rd(t); }
__device__ __forceinline__ int floor_to_int(double t) { return __double2int_rd(t); }
__device__ __forceinline__ int floor_to_int(long long t) { return static_cast<int>(t); }   // never used (no uniform / table path for int64)

// shared-memory atomics on explicit 32-bit shared addresses (no generic->shared conversion per use)
__device__ __forceinline__ void reds_add_u32(unsigned addr, unsigned v) {
  asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned atoms_add_u32(unsigned addr, unsigned v) {
  unsigned old;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ void reds_add_f64(unsigned addr, double v) {
  asm volatile("red.shared.add.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ long long to_ll_rn(float v) { return __float2ll_rn(v); }
__device__ __forceinline__ long long to_ll_rn(double v) { return __double2ll_rn(v); }

// ---------------------------------------------------------------------------------------------
// classification: bin of x for variable k, or -1 when x is dropped.
// Rule R1 of SURVEY.md §8a == core.py:157-174: in range iff e[0] <= x <= e[E-1]; bin =
// #{j: e[j] <= x} - 1 with x == e[E-1] falling in the last bin.  `e` holds the EFFECTIVE edges:
// for fp32 data each float64 edge is replaced by the smallest fp32 >= edge, which makes an fp32
// compare decide exactly like numpy's promoted float64 compare (rule R2).
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ int search_bin(const T* __restrict__ e, int nb, T x) {
  int lo = 0, hi = nb + 1;  // #{e[j] <= x} is in [lo, hi]
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (e[mid] <= x) lo = mid + 1; else hi = mid;
  }
  int b = lo - 1;
  return b > nb - 1 ? nb - 1 : b;  // right-inclusive last bin (core.py:171-173)
}

// Uniform-edge arithmetic.  t = (x - e0) * inv is within delta/2 of the exact position of x in
// bin units (bound computed on the host from the rounding errors and from the deviation of the
// edges from an arithmetic progression), so j = floor(t) is THE bin whenever
// delta <= frac(t) <= 1 - delta and 0 <= j < nb; only the other samples (~1e-4) need the search.
template <typename T>
__device__ __forceinline__ bool uniform_guess(const XhkParams& p, int k, T x, int& j) {
  const T t = (x - Consts<T>::get(p, k, XHK_C_E0)) * Consts<T>::get(p, k, XHK_C_INV);
  j = floor_to_int(t);                // NaN -> 0, +-inf / huge -> saturated: both fail the tests below
  const T f = t - static_cast<T>(j);  // NaN stays NaN: compares false
  return (f >= Consts<T>::get(p, k, XHK_C_DELTA)) & (f <= Consts<T>::get(p, k, XHK_C_OMD));
}

// Fast path: r = t - 0.5 is split into rint(r) (as a T, and as an int with a bias) so that d = r - rint(r) is
// frac(t) - 0.5 — the certainty test is |d| <= chalf — and rint(r) == floor(t) whenever the sample is certain.  The
// integer sits in the mantissa of r + 1.5 * 2^23 (fp32) / 1.5 * 2^52 (fp64): no F2I / I2F conversion.  Outside the
// valid range (|r| >= 2^22 resp. 2^31, NaN, inf) ok() is false or the integer lands outside [0, 2^22), where the
// window test of the caller rejects it (bin counts on this path are <= 2^21).
template <typename T> struct RoundSplit;
template <> struct RoundSplit<float> {
  static constexpr int kBias = 0x4B400000;
  static __device__ __forceinline__ void run(float r, float& jf, int& jraw) {
    const float s = r + 12582912.0f;
    jf = s - 12582912.0f;
    jraw = __float_as_int(s);
  }
  static __device__ __forceinline__ bool ok(float) { return true; }
};
template <> struct RoundSplit<double> {
  static constexpr int kBias = 0;
  static __device__ __forceinline__ void run(double r, double& jf, int& jraw) {
    const double s = r + 6755399441055744.0;
    jf = s - 6755399441055744.0;
    jraw = __double2loint(s);
  }
  // 0 <= rint(r) < 2^32  <=>  the high word of s is that of 1.5 * 2^52 (checked on the recomputed s: same value)
  static __device__ __forceinline__ bool ok(double r) { return __double2hiint(r + 6755399441055744.0) == 0x43380000; }
};
template <> struct RoundSplit<long long> {   // never used (no uniform path for int64)
  static constexpr int kBias = 0;
  static __device__ __forceinline__ void run(long long r, long long& jf, int& jraw) { jraw = static_cast<int>(r); jf = r; }
  static __device__ __forceinline__ bool ok(long long) { return true; }
};
__device__ __forceinline__ float fma_t(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ double fma_t(double a, double b, double c) { return __fma_rn(a, b, c); }
__device__ __forceinline__ long long fma_t(long long a, long long b, long long c) { return a * b + c; }
// keep a loop-invariant value in its register (under the 64-register cap ptxas otherwise rematerialises whole address
// chains inside the hot loop)
__device__ __forceinline__ unsigned pin(unsigned v) { asm volatile("" : "+r"(v)); return v; }
__device__ __forceinline__ float pin(float v) { asm volatile("" : "+f"(v)); return v; }

// exact bin of any sample (slow but general): range test, uniform guess when usable, else search
template <typename T> __device__ __forceinline__ T lut_inv(const XhkParams& p, int k);
template <> __device__ __forceinline__ float lut_inv<float>(const XhkParams& p, int k) { return p.lut_invf[k]; }
template <> __device__ __forceinline__ double lut_inv<double>(const XhkParams& p, int k) { return p.lut_invd[k]; }
template <> __device__ __forceinline__ long long lut_inv<long long>(const XhkParams&, int) { return 0ll; }

// Non-uniform edges: the cell of x in a uniform partition of [lo, hi] brackets #{e_j <= x} between the table
// entries of cell c-1 and cell c+2 (one cell of slack on either side absorbs the rounding of the cell index),
// so the binary search runs over a handful of edges instead of all of them.
template <typename T>
__device__ __forceinline__ int lut_bin(const XhkParams& p, int k, const T* __restrict__ e, const unsigned short* __restrict__ lut, T x) {
  const int nb = p.nb[k], G = p.lut_n[k];
  int c = floor_to_int((x - Consts<T>::get(p, k, XHK_C_LO)) * lut_inv<T>(p, k));
  c = max(0, min(c, G - 1));
  int lo = static_cast<int>(lut[max(c - 1, 0)]) + 1;                       // e[lo-1] <= x is known
  int hi = (c + 2 < G) ? static_cast<int>(lut[c + 2]) + 1 : nb + 1;         // #{e_j <= x} <= hi
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (e[mid - 1] <= x) lo = mid; else hi = mid - 1;
  }
  const int b = lo - 1;
  return b > nb - 1 ? nb - 1 : b;
}

template <typename T>
__device__ __forceinline__ int exact_bin_inline(const XhkParams& p, int k, const T* __restrict__ sedges,
                                                const unsigned short* __restrict__ slut, T x) {
  if (!(x >= Consts<T>::get(p, k, XHK_C_LO) && x <= Consts<T>::get(p, k, XHK_C_HI))) return -1;  // NaN: dropped (rule R3)
  const int nb = p.nb[k];
  if (p.uniform[k]) {
    int j;
    if (uniform_guess<T>(p, k, x, j) && static_cast<unsigned>(j) < static_cast<unsigned>(nb)) return j;
  }
  if (p.lut_n[k]) return lut_bin<T>(p, k, sedges + p.eoff[k], slut + p.lut_off[k], x);
  return search_bin<T>(sedges + p.eoff[k], nb, x);
}
// out-of-line copy for the rare exact path of the fast kernel (keeps its hot loop small)
template <typename T>
__device__ __noinline__ int exact_bin(const XhkParams& p, int k, const T* __restrict__ sedges,
                                      const unsigned short* __restrict__ slut, T x) {
  return exact_bin_inline<T>(p, k, sedges, slut, x);
}

// ---------------------------------------------------------------------------------------------
// streaming loads (one-touch data: evict-first)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void load4(const float* p, long long g, float (&v)[4]) {
  float4 q = __ldcs(reinterpret_cast<const float4*>(p) + g);
  v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
}
__device__ __forceinline__ void load4(const long long* p, long long g, long long (&v)[4]) {
  longlong2 a = __ldcs(reinterpret_cast<const longlong2*>(p) + 2 * g);
  longlong2 b = __ldcs(reinterpret_cast<const longlong2*>(p) + 2 * g + 1);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
__device__ __forceinline__ void load4(const double* p, long long g, double (&v)[4]) {
  double2 a = __ldcs(reinterpret_cast<const double2*>(p) + 2 * g);
  double2 b = __ldcs(reinterpret_cast<const double2*>(p) + 2 * g + 1);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}

__device__ __forceinline__ void prefetch_l2(const void* q) { asm volatile("prefetch.global.L2 [%0];" ::"l"(q)); }

// two samples (half a group) per load: big records are classified two samples at a time (see the MODE 2 half-group path)
__device__ __forceinline__ void load2(const float* p, long long h, float (&v)[2]) {
  float2 q = __ldcs(reinterpret_cast<const float2*>(p) + h);
  v[0] = q.x; v[1] = q.y;
}
__device__ __forceinline__ void load2(const double* p, long long h, double (&v)[2]) {
  double2 q = __ldcs(reinterpret_cast<const double2*>(p) + h);
  v[0] = q.x; v[1] = q.y;
}
__device__ __forceinline__ void load2(const long long* p, long long h, long long (&v)[2]) {
  longlong2 q = __ldcs(reinterpret_cast<const longlong2*>(p) + h);
  v[0] = q.x; v[1] = q.y;
}

template <int W> struct WType { using type = float; };   // W = 1 and W = 3: fp32 weights
template <> struct WType<2> { using type = double; };

// ---------------------------------------------------------------------------------------------
// the histogram kernel
//   T    : data type (float / double)        W : 0 no weights, 1 fp32 weights, 2 fp64 weights,
//                                                3 fp32 weights accumulated as ONE u32 limb per bin (fx32, see below)
//   KT   : number of variables at compile time (1..4), or 0 = runtime p.n_vars (scalar loads only)
//   MODE : 0 general (per-sample exact classification: range test, uniform guess / table-bracketed search)
//          1 every variable has evenly spaced edges -> branch-free arithmetic classification of 8 samples at a
//            time; samples that are uncertain or fall outside the shared window take a side path
//          2 every variable is uniform or has a bounded-step lookup table -> branch-free classification of
//            4 samples at a time, window spills as inline global REDs
// ---------------------------------------------------------------------------------------------
template <typename T, int W, int KT, int MODE>
__global__ void __launch_bounds__(kMaxThreads, 1) k_hist(const __grid_constant__ XhkParams p) {
  constexpr bool FAST = (MODE == 1 || MODE == 3 || MODE == 5);   // 3: the same with row tiling compiled in (1: compiled out);
                                                                  // 5: mode 1 with dynamic dealing of the groups to the warps (W = 3 only)
  constexpr bool CNT = (W == 0 || W == 4);    // counts (no weights)
  constexpr bool PK = (W == 4);               // counts packed two to a shared word (16-bit fields with a guard bit, see shared_add1)
  using HT = typename std::conditional<CNT || W == 3, unsigned int, double>::type;   // shared accumulator
  using OT = typename std::conditional<CNT, unsigned long long, double>::type;       // global accumulator
  using WT = typename WType<W>::type;
  constexpr int KMAX = KT ? KT : XHK_MAX_VARS;
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ int s_wlo[XHK_MAX_VARS], s_wlen[XHK_MAX_VARS];
  __shared__ int s_redo;   // fixed point, row owned by this CTA: a weight did not fit -> redo the segment with float64 adds
  __shared__ unsigned s_next;   // dynamic dealing of the groups of a segment to the warps (see `draw` below)
  __shared__ unsigned s_slow;   // samples of this CTA that left the fast path (window spills, weights outside the fixed-point form)
  // fp32 weights are served by two sibling launches; the probe kernel's verdict (XhkWindow::fx_mode) decides which of
  // them does the work, the other one returns at once (no host round trip between probe and histogram):
  //   W == 3 (fx32): 4 bytes per bin — twice the shared window of the 8-byte forms — when (nearly) all weights are
  //                  non-negative multiples of a common power of two spanning <= 25 bits (e.g. k * 2^-24 in [0, 1))
  //   W == 1       : everything else (two u32 limbs, or float64 shared adds)
  if constexpr (W == 3) { if (p.window->fx_mode != 32) return; }
  if constexpr (W == 1) { if (p.fx32_sibling && p.window->fx_mode == 32) return; }

  const int K = KT ? KT : p.n_vars;
  const int tid = threadIdx.x, nthr = blockDim.x;
  T* sedges = reinterpret_cast<T*>(smem);
  const size_t edges_al = (static_cast<size_t>(p.n_edges_total) * sizeof(T) + 15) & ~static_cast<size_t>(15);
  unsigned short* slut = reinterpret_cast<unsigned short*>(smem + edges_al);
  // histogram region: [32 trash slots (u32)][bins ...]; `shist` points at the bins.  A lane with nothing to add puts a
  // zero into its own trash slot (index lane - 32 relative to the bins), which keeps the shared adds free of branches.
  unsigned int* const hregion = reinterpret_cast<unsigned int*>(smem + edges_al + ((static_cast<size_t>(p.n_lut_total) * 2 + 15) & ~static_cast<size_t>(15)));
  HT* const shist = reinterpret_cast<HT*>(hregion + 32);

  for (int i = tid; i < p.n_edges_total; i += nthr) sedges[i] = static_cast<const T*>(p.edges)[i];
  for (int i = tid; i < p.n_lut_total; i += nthr) slut[i] = p.lut[i];
  if (tid < XHK_MAX_VARS) {
    int lo = 0, len = 0;
    if (tid < K) {
      if (p.hist_mode == XHK_WINDOW) { lo = p.window->lo[tid]; len = p.window->len[tid]; }
      else if (p.hist_mode == XHK_FULL) { lo = 0; len = p.nb[tid]; }
    }
    s_wlo[tid] = lo; s_wlen[tid] = len;
  }
  if (tid == 0) { s_redo = 0; s_slow = 0u; s_next = 0u; }
  __syncthreads();
  int wlo[KMAX], wlen[KMAX];
  int wtot = (p.hist_mode == XHK_GLOBAL) ? 0 : 1;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    if (k < K) { wlo[k] = s_wlo[k]; wlen[k] = s_wlen[k]; wtot *= wlen[k]; } else { wlo[k] = 0; wlen[k] = 0; }
  }
  // row tiling: the shared histogram holds tile_rows rows; a sample's bin is offset by (local row) * (bins per row),
  // obtained for free by seeding the Horner evaluation of the joint bin with the local row
  const bool tiled = (MODE == 3) || (MODE != 1 && MODE != 5 && p.tile_rows > 1);
  wtot *= p.tile_rows;
  const int hwords = PK ? ((wtot + 1) / 2 + 32) : (wtot + 32) * static_cast<int>(sizeof(HT) / 4);   // (+ the trash slots)
  for (int i = tid; i < hwords; i += nthr) hregion[i] = 0u;
  // weighted accumulation mode (uniform for the launch): exact fixed point in two u32 limbs, or float64 adds
  // `fx` can fall back to float64 adds for one row segment (see s_redo), hence not const
  bool fx = false, fx_launch = false; WT fx_mul = WT(0), fx_limit = WT(0); double fx_unmul = 0.0, fx_carry = 0.0;
  bool owned = false;      // the current row segment is a whole row that only this CTA touches: flushed with plain stores
  if constexpr (!CNT) {
    if (p.hist_mode != XHK_GLOBAL && p.window->fx_ok) {
      fx = fx_launch = true; fx_mul = static_cast<WT>(p.window->fx_mul); fx_limit = static_cast<WT>(p.window->fx_limit); fx_unmul = p.window->fx_unmul;
      fx_carry = 4294967296.0 * fx_unmul;   // fx32: what one wrap of the u32 limb is worth
    }
  }
  // two-limb fixed point: [32 trash][wtot low limbs][32 trash][wtot high limbs]
  const int wcap = wtot + 32;
  const unsigned sh_lo = pin(static_cast<unsigned>(__cvta_generic_to_shared(shist)));   // counts / lo limbs / doubles
  const unsigned sh_hi = sh_lo + 4u * static_cast<unsigned>(wcap);                  // hi limbs (fixed point)
  const unsigned trash = pin(static_cast<unsigned>((tid & 31) - 32));   // negative index: the lane's trash slot
  __syncthreads();

  OT* const out = static_cast<OT*>(p.out);
  const long long total = p.M * p.N;
  long long s0, s1;
  if (p.partition == XHK_PART_ROWS) {
    s0 = (p.M * blockIdx.x / gridDim.x) * p.N;
    s1 = (p.M * (blockIdx.x + 1ll) / gridDim.x) * p.N;
  } else {
    s0 = blockIdx.x * p.per_cta; if (s0 > total) s0 = total;
    s1 = s0 + p.per_cta; if (s1 > total) s1 = total;
  }

  // ---- accumulation primitives ------------------------------------------------------------
  auto global_add = [&](OT* out_row, long long gbin, double wv) {
    if constexpr (CNT) atomicAdd(out_row + gbin, 1ull); else atomicAdd(out_row + gbin, wv);
  };
  // an in-range sample that the shared histogram could not take (outside the window, or a weight outside the
  // fixed-point form): global add + one tick of the CTA's slow-path counter.  The host watches the counter to
  // notice a cached probe verdict that no longer fits the data (xhist_api.cu, struct Verdict).
  unsigned nslow = 0;      // per thread; summed into s_slow at the end (a shared counter bumped per spill serialises the
                           // lanes of a warp on one address: data that spills a lot — uniform over the bins — ran 4x slower)
  auto note_slow = [&]() { ++nslow; };
  auto spill_add = [&](OT* out_row, long long gbin, double wv) { global_add(out_row, gbin, wv); note_slow(); };
  // general path of one sample: exact bins, then shared window / global spill / drop.
  // Returns the window bin when the caller should do the shared add itself, else -1.
  auto general_sample = [&](const T (&x)[KMAX], double wv, OT* out_row, int rowl) -> int {
    int j[KMAX]; int wbin = rowl; bool ok = true, inwin = true;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      if (k < K) {
        j[k] = FAST ? exact_bin<T>(p, k, sedges, slut, x[k]) : exact_bin_inline<T>(p, k, sedges, slut, x[k]);
        ok = ok && (j[k] >= 0);
        const unsigned jw = static_cast<unsigned>(j[k] - wlo[k]);
        inwin = inwin && (jw < static_cast<unsigned>(wlen[k]));   // also false for j == -1
        wbin = wbin * wlen[k] + static_cast<int>(jw);
      }
    }
    if (inwin) return wbin;
    if (ok) {
      long long gbin = 0;
#pragma unroll
      for (int k = 0; k < KMAX; ++k) if (k < K) gbin = gbin * p.nb[k] + j[k];
      spill_add(out_row, gbin, wv);
    }
    return -1;
  };
  // window bin -> global bin (rare paths only)
  auto window_to_global = [&](int wbin) -> long long {
    if (p.hist_mode == XHK_FULL) return wbin;     // the window is the whole (possibly tiled) bin space
    int rem = wbin; long long gbin = 0;
#pragma unroll
    for (int k = KMAX - 1; k >= 0; --k) {
      if (k < K) { const int q = rem / wlen[k]; const int c = rem - q * wlen[k]; rem = q; gbin += (wlo[k] + c) * p.gmul[k]; }
    }
    return gbin;
  };
  // shared add of one sample.
  //  counts   : native RED.ADD.U32.
  //  weighted : fixed point (fx) — v = w * 2^-s as a 64-bit integer split over two u32 limbs: one native
  //             ATOMS.ADD on the low limb (its return value gives the carry) and a RED on the high limb when
  //             it is non-zero.  Integer adds are exact and order independent; a weight that is not an exact
  //             multiple of 2^s below the limit (also NaN/inf) goes to a float64 global RED instead.
  //             Otherwise float64 adds in shared memory (red.shared.add.f64 = LDS + DADD + ATOMS.CAST.SPIN loop;
  //             an explicit atomicCAS loop compiles to plain ATOMS.CAS.64 and measured >4x slower on B200).
  auto shared_add1 = [&](int wbin, WT w, OT* out_row) {
    if constexpr (PK) {
      // Packed counts: bin b lives in the 16-bit field (b & 1) of word b >> 1.  A field never passes 2^15 + (adds in flight):
      // the add that finds the field at 2^15 - 1 (ATOMS returns the old word) takes 2^15 out of the field again and
      // credits it to the output, so no field ever carries into its neighbour.  65 536 bins cost 128 KB instead of 256 KB
      // — the whole bin space of config 3 fits, no window, no spills, whatever the distribution of the data.
      const unsigned sh = (static_cast<unsigned>(wbin) & 1u) * 16u;
      const unsigned addr = sh_lo + 4u * (static_cast<unsigned>(wbin) >> 1);
      const unsigned old = atoms_add_u32(addr, 1u << sh);
      if (((old >> sh) & 0xFFFFu) == 0x7FFFu) {
        reds_add_u32(addr, 0u - (0x8000u << sh));
        atomicAdd(out_row + window_to_global(wbin), 0x8000ull);
      }
    } else if constexpr (CNT) {
      reds_add_u32(sh_lo + 4u * static_cast<unsigned>(wbin), 1u);
    } else if constexpr (W == 3) {
      // fx32: v = w * 2^s as ONE u32 limb.  The returned old value tells when the limb wraps; a wrap is worth
      // 2^32 * 2^-s and goes straight to the float64 output (about one add in 512 for weights in [0, 1)).
      const float vs = w * fx_mul;
      const unsigned v = __float2uint_rn(vs);
      if ((static_cast<float>(v) == vs) & (vs < fx_limit)) {
        const unsigned old = atoms_add_u32(sh_lo + 4u * static_cast<unsigned>(wbin), v);
        if (old + v < old) atomicAdd(out_row + window_to_global(wbin), fx_carry);
      } else {
        spill_add(out_row, window_to_global(wbin), static_cast<double>(w));
      }
    } else {
      if (fx) {
        const WT vs = w * fx_mul;
        const long long v = to_ll_rn(vs);
        if ((static_cast<WT>(v) == vs) & (fabs(vs) < fx_limit)) {
          const unsigned lo = static_cast<unsigned>(v);
          unsigned hi = static_cast<unsigned>(static_cast<unsigned long long>(v) >> 32);
          const unsigned old = atoms_add_u32(sh_lo + 4u * static_cast<unsigned>(wbin), lo);
          hi += (old + lo < old) ? 1u : 0u;
          if (hi) reds_add_u32(sh_hi + 4u * static_cast<unsigned>(wbin), hi);
        } else if (owned) {
          s_redo = 1;      // no global add may precede the plain stores of an owned row
        } else {
          spill_add(out_row, window_to_global(wbin), static_cast<double>(w));
        }
      } else {
        reds_add_f64(sh_lo + 8u * static_cast<unsigned>(wbin), static_cast<double>(w));
      }
    }
  };
  // Side path of the fused fast kernels (every variable has evenly spaced edges): the exact bin of ONE sample,
  // straight-line.  floor(t) is within one bin of the truth (|t_hat - t| < delta / 2 < 1/16), so comparing x with the
  // two staged edges around the candidate settles it — what numpy's own uniform-bin path does
  // (_histograms_impl.py:851-863).  In range and inside the window -> shared add, in range outside -> global RED.
  auto side_uniform = [&](const T (&x)[KMAX], WT wsel, int rowl, OT* out_row) {
    bool ok = true, inwin = true; int wbin = rowl; long long gbin = 0;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      if (k < K) {
        const T xx = x[k];
        ok = ok & (xx >= Consts<T>::get(p, k, XHK_C_LO)) & (xx <= Consts<T>::get(p, k, XHK_C_HI));   // NaN: false (rule R3)
        const int nb = p.nb[k];
        int b = floor_to_int((xx - Consts<T>::get(p, k, XHK_C_E0)) * Consts<T>::get(p, k, XHK_C_INV));
        b = max(0, min(b, nb - 1));
        const T* ed = sedges + p.eoff[k];
        if (xx < ed[b]) b = max(b - 1, 0);
        else if (b + 1 < nb && xx >= ed[b + 1]) b += 1;      // (the last bin is right-inclusive: never past nb - 1)
        const unsigned jw = static_cast<unsigned>(b - wlo[k]);
        inwin = inwin & (jw < static_cast<unsigned>(wlen[k]));
        wbin = wbin * wlen[k] + static_cast<int>(jw);
        gbin = gbin * nb + b;
      }
    }
    if (!ok) return;
    if (inwin) shared_add1(wbin, wsel, out_row);
    else spill_add(out_row, gbin, static_cast<double>(wsel));
  };
  auto shared_add4 = [&](const int (&wb)[4], const WT (&wv)[4], OT* out_row) {
#pragma unroll
    for (int e = 0; e < 4; ++e) if (wb[e] >= 0) shared_add1(wb[e], wv[e], out_row);
  };

  long long s = s0;
  while (s < s1) {
    const long long r = s / p.N;
    const long long c0 = s - r * p.N;
    long long len = p.N - c0;
    if (len > s1 - s) len = s1 - s;
    if (len > kSegCap) len = kSegCap;
    OT* out_row = out + r * p.B;
    owned = p.hist_mode == XHK_FULL && p.store_owned_rows && c0 == 0 && len == p.N;

    const T* px[KMAX];
    bool vec_ok = (KT != 0);
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
      px[k] = (k < K) ? static_cast<const T*>(p.data[k]) + r * p.stride[k] + c0 : nullptr;
    const WT* pw = (!CNT) ? static_cast<const WT*>(p.w) + r * p.wstride + c0 : nullptr;
    long long head = ((16 - (reinterpret_cast<uintptr_t>(px[0]) & 15)) & 15) / sizeof(T);
    if (head > len) head = len;
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
      if (k < K) vec_ok = vec_ok && ((reinterpret_cast<uintptr_t>(px[k] + head) & 15) == 0);
    if (!CNT) vec_ok = vec_ok && ((reinterpret_cast<uintptr_t>(pw + head) & 15) == 0);
    if (!vec_ok) head = len;
    const long long nvec = (len - head) >> 2;  // groups of 4 samples
    const long long tail0 = head + (nvec << 2);

    // local row of sample i of this segment inside its tile (0 without tiling)
    auto local_row = [&](long long i) -> int {
      if (!tiled) return 0;
      const unsigned n = static_cast<unsigned>(c0 + i);
      return static_cast<int>((__umulhi(n, p.tile_magic) + n) >> p.tile_shift);
    };
    // scalar head and tail (and everything when the arrays are not mutually 16-byte alignable)
    auto scalar_at = [&](long long i) {
      T x[KMAX];
#pragma unroll
      for (int k = 0; k < KMAX; ++k) x[k] = (k < K) ? px[k][i] : T(0);
      WT wv[4] = {WT(1), WT(1), WT(1), WT(1)};
      if constexpr (!CNT) wv[0] = pw[i];
      const int wbin = general_sample(x, static_cast<double>(wv[0]), out_row, local_row(i));
      if (wbin >= 0) shared_add1(wbin, wv[0], out_row);
    };
    for (long long i = tid; i < head; i += nthr) scalar_at(i);
    for (long long i = tail0 + tid; i < len; i += nthr) scalar_at(i);

    if constexpr (KT != 0) {
      // vector body: 4 samples per 16-byte load, U loads in flight per array and thread
      constexpr int REC = static_cast<int>(sizeof(T)) * KMAX + (CNT ? 0 : static_cast<int>(sizeof(WT)));   // bytes per sample
      constexpr int U = (FAST && CNT && REC <= 8) ? 3 : (REC <= 12) ? 2 : 1;
      const float fx_mulp = pin(static_cast<float>(fx_mul) * 2.98023223876953125e-8f);   // fx32: fx_mul * 2^-25
      int jbias[KMAX];          // fused fast path: window offset of the bin number + RoundSplit::kBias
#pragma unroll
      for (int k = 0; k < KMAX; ++k) jbias[k] = wlo[k] + RoundSplit<T>::kBias;
      // How the groups of a segment are dealt to the threads.  Static: thread t takes groups t, t + U * nthr, ... .
      // Dynamic (MODE 5: one-limb weighted fast path on SHORT per-CTA ranges): every warp draws the next 32 * U groups from
      // a shared counter, so the warps of a CTA reach the end of the segment together instead of a few iterations apart
      // (the warps drift: side loops, bank conflicts) and the CTA does not wait at the barrier before the flush for its
      // slowest warp.  Measured on config 3: 6 % of the kernel at 1.25e8 samples (a 1/8 shard), but the extra live state
      // costs spills — 2 % slower at 1e9 — so the host picks it only for short ranges (kernel_mode()).
      constexpr bool DYN = (MODE == 5) && W == 3;
      const long long ustride = DYN ? 32 : nthr;
      auto draw = [&]() -> long long {
        unsigned b = 0;
        if ((tid & 31) == 0) b = atomicAdd(&s_next, static_cast<unsigned>(U * 32));
        return static_cast<long long>(__shfl_sync(0xffffffffu, b, 0)) + (tid & 31);
      };
      auto first_group = [&]() -> long long { if constexpr (DYN) return draw(); else return tid; };
      auto next_group = [&](long long g) -> long long { if constexpr (DYN) return draw(); else return g + static_cast<long long>(U) * nthr; };
      auto group_ok = [&](long long g) -> bool { if constexpr (DYN) return g - (tid & 31) < nvec; else return g < nvec; };   // (warp-uniform when dynamic)
      if constexpr (FAST && CNT) {
        // ---- counts on the fast path: fused classify + RED, software-pipelined over U register slots of 4 samples
        // per array — a slot is refilled with the group U steps ahead as soon as it has been consumed, so a warp keeps
        // loads in flight while it computes (measured +3-4 % on configs 2 / 3-counts / 4; the same order measured
        // 2.5 % SLOWER for the one-limb weighted form, which keeps the plain loop below).  Classification as in the
        // weighted fused path below: mantissa floor, one-compare certainty, `live` seeds the predicate chain; a sample
        // that is not (certain and inside the window) adds 1 to a trash slot and is redone by the side loop.
        T xv[U][KMAX][4];
        auto load_slot = [&](int u, long long gu) {
          if (gu < nvec) {
#pragma unroll
            for (int k = 0; k < KMAX; ++k) load4(px[k] + head, gu, xv[u][k]);
          }
#if XH_PREFETCH_DIST > 0
          const long long gp = gu + static_cast<long long>(XH_PREFETCH_DIST * U) * nthr;
          if (p.prefetch && gp < nvec) {
#pragma unroll
            for (int k = 0; k < KMAX; ++k) prefetch_l2(px[k] + head + 4 * gp);
          }
#endif
        };
#pragma unroll
        for (int u = 0; u < U; ++u) load_slot(u, tid + static_cast<long long>(u) * nthr);
        for (long long g = tid; g < nvec; g += static_cast<long long>(U) * nthr) {
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const long long gu = g + static_cast<long long>(u) * nthr;
            const bool live = (u == 0) || (gu < nvec);
            unsigned idx[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              bool good = live; int wbin = local_row(head + 4 * gu + e);
#pragma unroll
              for (int k = 0; k < KMAX; ++k) {
                const T r = fma_t(xv[u][k][e] - Consts<T>::get(p, k, XHK_C_E0), Consts<T>::get(p, k, XHK_C_INV), T(-0.5));
                T jf; int jraw;
                RoundSplit<T>::run(r, jf, jraw);
                const T d = r - jf;                                  // frac(t) - 0.5, exact
                const unsigned jw = static_cast<unsigned>(jraw - jbias[k]);
                good = good & (fabs(d) <= Consts<T>::get(p, k, XHK_C_CHALF)) & (jw < static_cast<unsigned>(wlen[k])) & RoundSplit<T>::ok(r);
                wbin = wbin * wlen[k] + static_cast<int>(jw);
              }
              idx[e] = good ? static_cast<unsigned>(wbin) : trash;
            }
            if constexpr (PK) {
              unsigned old[4], shf[4], wad[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const bool good = static_cast<int>(idx[e]) >= 0;
                shf[e] = (idx[e] & 1u) * 16u;
                wad[e] = sh_lo + 4u * (good ? (idx[e] >> 1) : idx[e]);       // (a trash slot keeps its own word)
              }
#pragma unroll
              for (int e = 0; e < 4; ++e) old[e] = atoms_add_u32(wad[e], 1u << shf[e]);
              unsigned full15 = 0;
#pragma unroll
              for (int e = 0; e < 4; ++e) full15 |= ((((old[e] >> shf[e]) & 0xFFFFu) == 0x7FFFu) && static_cast<int>(idx[e]) >= 0) ? (1u << e) : 0u;
              if (full15) {           // rare: once per 2^15 adds to one bin
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  if (full15 & (1u << e)) {
                    reds_add_u32(wad[e], 0u - (0x8000u << shf[e]));
                    atomicAdd(out_row + window_to_global(static_cast<int>(idx[e])), 0x8000ull);
                  }
              }
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e) reds_add_u32(sh_lo + 4u * idx[e], 1u);
            }
            if (live && max(max(idx[0], idx[1]), max(idx[2], idx[3])) >= 0x80000000u) {
              unsigned side = 0;
#pragma unroll
              for (int e = 0; e < 4; ++e) side |= (idx[e] >> 31) << e;
              while (side) {     // one sample per trip (a lane rarely has more than one)
                const int se = __ffs(side) - 1;
                side &= side - 1;
                T x[KMAX];
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  if (se == e) {
#pragma unroll
                    for (int k = 0; k < KMAX; ++k) x[k] = xv[u][k][e];
                  }
                side_uniform(x, WT(1), local_row(head + 4 * gu + se), out_row);
              }
            }
            load_slot(u, gu + static_cast<long long>(U) * nthr);
          }
        }
      } else if constexpr (MODE == 2 && (sizeof(T) * KMAX > 16)) {
        // ---- big records (e.g. three float64 variables + float64 weights = 32 bytes per sample; config 5): the same
        // branch-free classification as the MODE 2 path below, but over HALF a group (2 samples) at a time, which is
        // what fits the 64-register budget without spills — the general kernel this replaces for such records spent
        // its time in divergent search loops (267 instructions per sample, 75 % issue-bound: profiles/r2_ncu_cfg5.md).
        // Non-uniform variable: table entry of cell c-1 = a bin at or below the sample's, then lut_steps
        // compare-and-advance steps; window spills leave as inline predicated global REDs.
        // software pipeline: the loads of the next half group are issued before the current one is classified (one
        // 16-byte load per array and thread in flight was latency-bound: 30 % of the stall samples at the loads; measured
        // 3.19 -> 3.13 ms on config 5)
        T xn[KMAX][2]; WT wn[2] = {WT(1), WT(1)};
        auto fetch = [&](long long gg) {
          if (gg < 2 * nvec) {
#pragma unroll
            for (int k = 0; k < KMAX; ++k) load2(px[k] + head, gg, xn[k]);
            if constexpr (!CNT) load2(pw + head, gg, wn);
          }
        };
        fetch(tid);
        for (long long g = tid; g < 2 * nvec; g += nthr) {          // g counts half groups
          T xh[KMAX][2]; WT wh[2];
#pragma unroll
          for (int k = 0; k < KMAX; ++k) { xh[k][0] = xn[k][0]; xh[k][1] = xn[k][1]; }
          wh[0] = wn[0]; wh[1] = wn[1];
          fetch(g + nthr);
#if XH_PREFETCH_DIST > 0
          {
            const long long gp = g + static_cast<long long>(XH_PREFETCH_DIST + 1) * nthr;      // half groups ahead of the register prefetch
            if (p.prefetch && gp < 2 * nvec) {
#pragma unroll
              for (int k = 0; k < KMAX; ++k) prefetch_l2(px[k] + head + 2 * gp);
              if constexpr (!CNT) prefetch_l2(pw + head + 2 * gp);
            }
          }
#endif
          int jb[KMAX][2]; bool okr[2] = {true, true}, cert[2] = {true, true};
#pragma unroll
          for (int k = 0; k < KMAX; ++k) {
            if (p.uniform[k]) {
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                cert[e] = cert[e] & uniform_guess<T>(p, k, xh[k][e], jb[k][e]);
                okr[e] = okr[e] & (static_cast<unsigned>(jb[k][e]) < static_cast<unsigned>(p.nb[k]));
              }
            } else {
              const T lo = Consts<T>::get(p, k, XHK_C_LO), hi = Consts<T>::get(p, k, XHK_C_HI), inv = lut_inv<T>(p, k);
              const int G = p.lut_n[k], nb = p.nb[k], steps = p.lut_steps[k];
              const unsigned short* lut = slut + p.lut_off[k];
              const T* ed = sedges + p.eoff[k];
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const T xx = xh[k][e];
                okr[e] = okr[e] & (xx >= lo) & (xx <= hi);          // NaN: false
                int c = floor_to_int((xx - lo) * inv);
                c = max(0, min(c, G - 1));
                jb[k][e] = static_cast<int>(lut[max(c - 1, 0)]);
              }
              for (int st = 0; st < steps; ++st) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                  const int nx = min(jb[k][e] + 1, nb);
                  jb[k][e] = (ed[nx] <= xh[k][e]) ? nx : jb[k][e];
                }
              }
#pragma unroll
              for (int e = 0; e < 2; ++e) jb[k][e] = min(jb[k][e], nb - 1);   // right-inclusive last bin
            }
          }
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            bool inwin = true; int wbin = local_row(head + 2 * g + e); long long gbin = 0;
#pragma unroll
            for (int k = 0; k < KMAX; ++k) {
              const unsigned jw = static_cast<unsigned>(jb[k][e] - wlo[k]);
              inwin = inwin & (jw < static_cast<unsigned>(wlen[k]));
              wbin = wbin * wlen[k] + static_cast<int>(jw);
              gbin = gbin * p.nb[k] + jb[k][e];
            }
            if (cert[e]) {
              if (okr[e]) {
                if (inwin) shared_add1(wbin, wh[e], out_row);
                else { global_add(out_row, gbin, static_cast<double>(wh[e])); ++nslow; }   // in range, outside the shared window
              }
            } else {                                   // rare: uncertain sample of a uniform variable -> exact path
              T x1[KMAX];
#pragma unroll
              for (int k = 0; k < KMAX; ++k) x1[k] = xh[k][e];
              const int wb1 = general_sample(x1, static_cast<double>(wh[e]), out_row, local_row(head + 2 * g + e));
              if (wb1 >= 0) shared_add1(wb1, wh[e], out_row);
            }
          }
        }
      } else
      for (long long g = first_group(); group_ok(g); g = next_group(g)) {
        T xv[U][KMAX][4];
        WT wv[U][4];
#if XH_PREFETCH_DIST > 0
        {        // pull the lines of a later iteration into L2 (no registers held, unlike a register prefetch); with dynamic
                 // dealing some warp of the CTA will draw these groups about that many iterations from now
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const long long gp = g + static_cast<long long>(XH_PREFETCH_DIST * U) * nthr + static_cast<long long>(u) * ustride;
            if (p.prefetch && gp < nvec) {
#pragma unroll
              for (int k = 0; k < KMAX; ++k) prefetch_l2(px[k] + head + 4 * gp);
              if constexpr (!CNT) prefetch_l2(pw + head + 4 * gp);
            }
          }
        }
#endif
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const long long gu = g + static_cast<long long>(u) * ustride;
          if (gu < nvec) {
#pragma unroll
            for (int k = 0; k < KMAX; ++k) load4(px[k] + head, gu, xv[u][k]);
            if constexpr (!CNT) load4(pw + head, gu, wv[u]);
          }
          if constexpr (CNT) { wv[u][0] = wv[u][1] = wv[u][2] = wv[u][3] = WT(1); }
        }
        int wb[U][4];
        if constexpr (MODE == 2) {
          // ---- branch-free classification for mixed uniform / non-uniform variables.
          // Non-uniform: the table entry of cell c-1 is a bin at or below the sample's; at most lut_steps edges
          // lie between that cell's left boundary and the sample, so that many compare-and-advance steps give
          // the exact bin.  The steps run over the 4 samples of a group together (independent LDS chains).
          unsigned unsure = 0;
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const bool live = g + static_cast<long long>(u) * nthr < nvec;
            int jb[KMAX][4]; bool okr[4], cert[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) { okr[e] = live; cert[e] = live; }
#pragma unroll
            for (int k = 0; k < KMAX; ++k) {
              if (p.uniform[k]) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  cert[e] = cert[e] & uniform_guess<T>(p, k, xv[u][k][e], jb[k][e]);
                  okr[e] = okr[e] & (static_cast<unsigned>(jb[k][e]) < static_cast<unsigned>(p.nb[k]));
                }
              } else {
                const T lo = Consts<T>::get(p, k, XHK_C_LO), hi = Consts<T>::get(p, k, XHK_C_HI), inv = lut_inv<T>(p, k);
                const int G = p.lut_n[k], nb = p.nb[k], steps = p.lut_steps[k];
                const unsigned short* lut = slut + p.lut_off[k];
                const T* ed = sedges + p.eoff[k];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const T x = xv[u][k][e];
                  okr[e] = okr[e] & (x >= lo) & (x <= hi);          // NaN: false
                  int c = floor_to_int((x - lo) * inv);
                  c = max(0, min(c, G - 1));
                  jb[k][e] = static_cast<int>(lut[max(c - 1, 0)]);
                }
                for (int st = 0; st < steps; ++st) {
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const int nx = min(jb[k][e] + 1, nb);
                    jb[k][e] = (ed[nx] <= xv[u][k][e]) ? nx : jb[k][e];
                  }
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) jb[k][e] = min(jb[k][e], nb - 1);   // right-inclusive last bin
              }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              bool inwin = true; int wbin = local_row(head + 4 * (g + static_cast<long long>(u) * nthr) + e), gbin = 0;
#pragma unroll
              for (int k = 0; k < KMAX; ++k) {
                const unsigned jw = static_cast<unsigned>(jb[k][e] - wlo[k]);
                inwin = inwin & (jw < static_cast<unsigned>(wlen[k]));
                wbin = wbin * wlen[k] + static_cast<int>(jw);
                gbin = gbin * p.nb[k] + jb[k][e];
              }
              const bool good = cert[e] & okr[e];
              wb[u][e] = (good & inwin) ? wbin : -1;
              if (good & !inwin) {                      // in range, outside the shared window: one global RED
                if constexpr (CNT) atomicAdd(out_row + gbin, 1ull);
                else atomicAdd(out_row + gbin, static_cast<double>(wv[u][e]));
                ++nslow;
              }
              unsure |= (live & !cert[e]) ? (1u << (4 * u + e)) : 0u;
            }
          }
          while (unsure) {   // rare: uncertain samples of uniform variables -> exact path
            const int idx = __ffs(unsure) - 1;
            unsure &= unsure - 1;
            T x[KMAX]; WT wsel = WT(1);
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (idx == 4 * u + e) {
#pragma unroll
                  for (int k = 0; k < KMAX; ++k) x[k] = xv[u][k][e];
                  wsel = wv[u][e];
                }
            const int wbin = general_sample(x, static_cast<double>(wsel), out_row,
                                            local_row(head + 4 * (g + static_cast<long long>(idx >> 2) * nthr) + (idx & 3)));
            if (wbin >= 0) shared_add1(wbin, wsel, out_row);
          }
        } else if constexpr (FAST && W == 3) {
          // ---- one-limb weights on the fast path: fused classify + accumulate (the count path above explains the
          // classification).  Both slots are loaded at the top of the iteration, classified and added (4 ATOMS back to
          // back per slot), then ONE side loop serves the U * 4 samples.  A sample that is not (certain and inside
          // the window) adds its value to a trash slot and is picked up by the side loop; idx < 0 marks it.
          unsigned idx[U][4];
          unsigned worst = 0;
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const bool live = g + static_cast<long long>(u) * ustride < nvec;
            unsigned vv[4], old[4];
            bool any_rare = false;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              bool good = live; int wbin = local_row(head + 4 * (g + static_cast<long long>(u) * ustride) + e);
#pragma unroll
              for (int k = 0; k < KMAX; ++k) {
                const T r = fma_t(xv[u][k][e] - Consts<T>::get(p, k, XHK_C_E0), Consts<T>::get(p, k, XHK_C_INV), T(-0.5));
                T jf; int jraw;
                RoundSplit<T>::run(r, jf, jraw);
                const T d = r - jf;                                  // frac(t) - 0.5, exact
                const unsigned jw = static_cast<unsigned>(jraw - jbias[k]);
                good = good & (fabs(d) <= Consts<T>::get(p, k, XHK_C_CHALF)) & (jw < static_cast<unsigned>(wlen[k])) & RoundSplit<T>::ok(r);
                wbin = wbin * wlen[k] + static_cast<int>(jw);
              }
              idx[u][e] = good ? static_cast<unsigned>(wbin) : trash;
              // v = round(clamp(w * 2^s, 0, 2^25)); the clamp rides on the multiplier (FMUL.SAT).  A sample that
              // is not `good` adds its v to the trash slot (never read; checks below skip it)
              const float we = wv[u][e];
              vv[e] = __float2uint_rn(__saturatef(we * fx_mulp) * 33554432.0f);
              any_rare = any_rare | (static_cast<float>(vv[e]) != we * static_cast<float>(fx_mul));   // (NaN: true)
            }
            {
#pragma unroll
              for (int e = 0; e < 4; ++e) old[e] = atoms_add_u32(sh_lo + 4u * idx[u][e], vv[e]);
              // v < 2^26: the limb wrapped iff its top bit went from 1 to 0.  A wrap is worth 2^32 * 2^-s and goes
              // straight to the float64 output (about one add in 512 for weights in [0, 1)): one trip per wrap
              unsigned wrapped = 0;
#pragma unroll
              for (int e = 0; e < 4; ++e) wrapped |= old[e] & ~(old[e] + vv[e]);
              if (static_cast<int>(wrapped) < 0) {
                unsigned cm = 0;
#pragma unroll
                for (int e = 0; e < 4; ++e) cm |= ((old[e] & ~(old[e] + vv[e])) >> 31) << e;
                while (cm) {
                  const int ce = __ffs(cm) - 1;
                  cm &= cm - 1;
                  const unsigned sel = ce == 0 ? idx[u][0] : ce == 1 ? idx[u][1] : ce == 2 ? idx[u][2] : idx[u][3];
                  if (static_cast<int>(sel) >= 0) global_add(out_row, window_to_global(static_cast<int>(sel)), fx_carry);   // (not a trash slot)
                }
              }
              if (any_rare) {
                // rarer: v is not the weight (negative, too large, finer than the scale, NaN / inf): the float64
                // output gets the difference, which is exact in float64
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  if (static_cast<int>(idx[u][e]) >= 0) {
                    const double diff = static_cast<double>(wv[u][e]) - static_cast<double>(vv[e]) * fx_unmul;
                    if (diff != 0.0) spill_add(out_row, window_to_global(static_cast<int>(idx[u][e])), diff);
                  }
                }
              }
            }
            worst = max(worst, max(max(idx[u][0], idx[u][1]), max(idx[u][2], idx[u][3])));
          }
          if (worst >= 0x80000000u) {
            // side loop, one sample per trip (a lane rarely has more than one): certain and in range but outside
            // the window -> global RED; everything else (uncertain, out of range, NaN) -> exact path
            unsigned side = 0;
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
              for (int e = 0; e < 4; ++e)
                side |= (static_cast<int>(idx[u][e]) < 0 && g + static_cast<long long>(u) * ustride < nvec) ? (1u << (4 * u + e)) : 0u;
            while (side) {
              const int sidx = __ffs(side) - 1;
              side &= side - 1;
              T x[KMAX]; WT wsel = WT(1);
#pragma unroll
              for (int u = 0; u < U; ++u)
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  if (sidx == 4 * u + e) {
#pragma unroll
                    for (int k = 0; k < KMAX; ++k) x[k] = xv[u][k][e];
                    wsel = wv[u][e];
                  }
#if XH_CHEAP_SIDE_W3
              // experiment switch (make EXTRA=-DXH_CHEAP_SIDE_W3=1, tools/erratic_probe.sh): fast at best (2.10 ms, 0.87 of
              // the HBM peak on config 3) but erratic, identical launches take 2.1 to 4.3 ms; see DESIGN.md section 8
              side_uniform(x, wsel, local_row(head + 4 * (g + static_cast<long long>(sidx >> 2) * ustride) + (sidx & 3)), out_row);
#else
              // (the straight-line side_uniform() of the count path measured UNSTABLE here: identical launches took
              //  2.1 to 4.0 ms; with this call-based exact path they take 2.21 ms every time)
              bool sure = true; long long gbin = 0;
#pragma unroll
              for (int k = 0; k < KMAX; ++k) {
                int jx; const bool certain = uniform_guess<T>(p, k, x[k], jx);
                sure = sure & certain & (static_cast<unsigned>(jx) < static_cast<unsigned>(p.nb[k]));
                gbin = gbin * p.nb[k] + jx;
              }
              if (sure && !tiled && p.hist_mode != XHK_FULL) spill_add(out_row, gbin, static_cast<double>(wsel));
              else {
                const int wbin = general_sample(x, static_cast<double>(wsel), out_row,
                                                local_row(head + 4 * (g + static_cast<long long>(sidx >> 2) * ustride) + (sidx & 3)));
                if (wbin >= 0) shared_add1(wbin, wsel, out_row);
              }
#endif
            }
          }
          continue;   // (the generic accumulation below serves the other forms)
        } else if constexpr (FAST) {
          // phase A (branch-free, U*4*K independent chains): guess, certainty, window test
          unsigned side = 0;   // bit (4u+e) set: that sample needs the side path
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const bool live = g + static_cast<long long>(u) * nthr < nvec;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              bool good = true; int wbin = local_row(head + 4 * (g + static_cast<long long>(u) * nthr) + e);
#pragma unroll
              for (int k = 0; k < KMAX; ++k) {
                int j;
                const bool certain = uniform_guess<T>(p, k, xv[u][k][e], j);
                const unsigned jw = static_cast<unsigned>(j - wlo[k]);
                good = good & certain & (jw < static_cast<unsigned>(wlen[k]));
                wbin = wbin * wlen[k] + static_cast<int>(jw);
              }
              wb[u][e] = (good & live) ? wbin : -1;
              side |= (!good & live) ? (1u << (4 * u + e)) : 0u;
            }
          }
          // side path, one sample per trip (a lane rarely has more than one): certain and in range but
          // outside the window -> global RED; everything else (uncertain, out of range, NaN) -> exact path
          while (side) {
            const int idx = __ffs(side) - 1;
            side &= side - 1;
            T x[KMAX]; WT wsel = WT(1);
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (idx == 4 * u + e) {
#pragma unroll
                  for (int k = 0; k < KMAX; ++k) x[k] = xv[u][k][e];
                  wsel = wv[u][e];
                }
            bool sure = true; long long gbin = 0;
#pragma unroll
            for (int k = 0; k < KMAX; ++k) {
              int jx; const bool certain = uniform_guess<T>(p, k, x[k], jx);
              sure = sure & certain & (static_cast<unsigned>(jx) < static_cast<unsigned>(p.nb[k]));
              gbin = gbin * p.nb[k] + jx;
            }
            if (sure && !tiled) spill_add(out_row, gbin, static_cast<double>(wsel));   // (tiling implies a full window)
            else {
              const int wbin = general_sample(x, static_cast<double>(wsel), out_row,
                                              local_row(head + 4 * (g + static_cast<long long>(idx >> 2) * nthr) + (idx & 3)));
              if (wbin >= 0) shared_add1(wbin, wsel, out_row);
            }
          }
        } else {
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const bool live = g + static_cast<long long>(u) * nthr < nvec;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              T x[KMAX];
#pragma unroll
              for (int k = 0; k < KMAX; ++k) x[k] = xv[u][k][e];
              wb[u][e] = live ? general_sample(x, static_cast<double>(wv[u][e]), out_row, local_row(head + 4 * (g + static_cast<long long>(u) * nthr) + e)) : -1;
            }
          }
        }
        if constexpr (W == 3) {
          // fx32 shared adds, straight-line: 4 ATOMS back to back; wraps of the limb and weights that are not exact
          // at the scale (negative, too large, finer than 2^-s, NaN/inf) are rare and leave through global REDs
#pragma unroll
          for (int u = 0; u < U; ++u) {
            unsigned idx[4], lo[4], old[4], rare = 0;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float vs = wv[u][e] * fx_mul;
              const unsigned v = __float2uint_rn(vs);
              const bool exact = (static_cast<float>(v) == vs) & (vs < fx_limit);
              const bool valid = wb[u][e] >= 0;
              const bool ok = valid & exact;
              rare |= (valid & !exact) ? (1u << e) : 0u;
              idx[e] = ok ? static_cast<unsigned>(wb[u][e]) : trash;
              lo[e] = ok ? v : 0u;
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) old[e] = atoms_add_u32(sh_lo + 4u * idx[e], lo[e]);
#pragma unroll
            for (int e = 0; e < 4; ++e) rare |= (old[e] + lo[e] < old[e]) ? (16u << e) : 0u;
            if (rare) {
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                if (rare & (1u << e)) spill_add(out_row, window_to_global(wb[u][e]), static_cast<double>(wv[u][e]));
                if (rare & (16u << e)) global_add(out_row, window_to_global(wb[u][e]), fx_carry);
              }
            }
          }
        } else if (!CNT && fx) {
          // fixed-point shared adds, straight-line: 4 low-limb ATOMS back to back, then the 4 high-limb REDs
          // (each takes the carry from the value its ATOMS returned)
          unsigned inexact = 0;
#pragma unroll
          for (int u = 0; u < U; ++u) {
            unsigned idx[4], lo[4], hi[4], old[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const WT vs = wv[u][e] * fx_mul;
              const long long v = to_ll_rn(vs);
              const bool exact = (static_cast<WT>(v) == vs) & (fabs(vs) < fx_limit);
              const bool valid = wb[u][e] >= 0;
              const bool ok = valid & exact;
              inexact |= (valid & !exact) ? (1u << (4 * u + e)) : 0u;
              idx[e] = ok ? static_cast<unsigned>(wb[u][e]) : trash;
              lo[e] = ok ? static_cast<unsigned>(v) : 0u;
              hi[e] = ok ? static_cast<unsigned>(static_cast<unsigned long long>(v) >> 32) : 0u;
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) old[e] = atoms_add_u32(sh_lo + 4u * idx[e], lo[e]);
#pragma unroll
            for (int e = 0; e < 4; ++e) reds_add_u32(sh_hi + 4u * idx[e], hi[e] + ((old[e] + lo[e] < old[e]) ? 1u : 0u));
          }
          if (inexact && owned) s_redo = 1;
          else if (inexact) {   // rare: weights that are not exact multiples of the scale -> float64 global RED
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (inexact & (1u << (4 * u + e))) spill_add(out_row, window_to_global(wb[u][e]), static_cast<double>(wv[u][e]));
          }
        } else {
#pragma unroll
          for (int u = 0; u < U; ++u) shared_add4(wb[u], wv[u], out_row);
        }
      }
    }
    // ---- flush the shared histogram of this row segment and clear it
    __syncthreads();
    if constexpr (W == 1 || W == 2) {
      // An owned row is written with plain stores, so nothing may reach it through global adds first.  If a weight
      // of this segment did not fit the fixed-point form, drop the partial sums and redo the segment with float64
      // shared adds (same shared-memory footprint).
      if (s_redo) {     // uniform: written before the barrier above
        __syncthreads();
        for (int i = tid; i < hwords; i += nthr) hregion[i] = 0u;
        if (tid == 0) s_redo = 0;
        fx = false;
        __syncthreads();
        continue;       // s has not advanced
      }
    }
    s += len;
    if (p.hist_mode != XHK_GLOBAL) {
      const bool full = p.hist_mode == XHK_FULL;
      unsigned int* lo32 = reinterpret_cast<unsigned int*>(shist);
      unsigned int* hi32 = lo32 + wcap;
      // value of shared bin b (cleared on the way out)
      auto take = [&](int b, bool& nz) -> OT {
        OT v;
        if constexpr (CNT) { v = static_cast<OT>(shist[b]); nz = v != 0; shist[b] = 0u; }
        else if constexpr (W == 3) { const unsigned q = lo32[b]; nz = q != 0u; v = static_cast<double>(q) * fx_unmul; lo32[b] = 0u; }
        else if (fx) {
          const long long iv = static_cast<long long>((static_cast<unsigned long long>(hi32[b]) << 32) | lo32[b]);
          nz = iv != 0; v = static_cast<double>(iv) * fx_unmul;
          lo32[b] = 0u; hi32[b] = 0u;
        } else { v = shist[b]; nz = v != 0.0; shist[b] = 0.0; }
        return v;
      };
      const int L = (K > 0) ? wlen[K - 1] : 1;          // bins of the last variable inside the window: contiguous in `out`
      if constexpr (PK) {
        // packed counts: one thread takes a whole word (two bins), so clearing needs no atomics
        for (int i = tid; i < (wtot + 1) / 2; i += nthr) {
          const unsigned q = lo32[i];
          lo32[i] = 0u;
#pragma unroll
          for (int hlf = 0; hlf < 2; ++hlf) {
            const int b = 2 * i + hlf;
            const unsigned long long v = (q >> (16 * hlf)) & 0xFFFFu;
            if (b < wtot && v) atomicAdd(out_row + (full ? static_cast<long long>(b) : window_to_global(b)), v);
          }
        }
      } else
      if (owned) {
        for (int b = tid; b < wtot; b += nthr) { bool nz; const OT v = take(b, nz); out_row[b] = v; }
      } else if (!full && L >= 32) {
        // windowed flush, one window row (L contiguous output bins) per warp trip: the window -> global index needs its
        // divisions once per row instead of once per bin, and every CTA starts at a different row so that the CTAs
        // of a launch (which all flush at about the same time) do not walk the same output addresses in step
        const int rows = wtot / L, lane = tid & 31, wid = tid >> 5, nw = nthr >> 5;
        const int start = static_cast<int>((static_cast<long long>(blockIdx.x) * rows) / gridDim.x);
        for (int i = wid; i < rows; i += nw) {
          int row = i + start; if (row >= rows) row -= rows;
          const long long gbase = window_to_global(row * L);
          for (int c = lane; c < L; c += 32) {
            bool nz; const OT v = take(row * L + c, nz);
            if (nz) atomicAdd(out_row + gbase + c, v);
          }
        }
      } else {
        const int start = static_cast<int>((static_cast<long long>(blockIdx.x) * wtot) / gridDim.x) & ~31;
        for (int i = tid; i < wtot; i += nthr) {
          int b = i + start; if (b >= wtot) b -= wtot;
          bool nz; const OT v = take(b, nz);
          if (nz) atomicAdd(out_row + (full ? static_cast<long long>(b) : window_to_global(b)), v);
        }
      }
    }
    fx = fx_launch;     // (a redone segment ran with float64 adds)
    if (tid == 0) s_next = 0u;
    __syncthreads();
  }
  if (nslow) atomicAdd(&s_slow, nslow);
  __syncthreads();
  if (tid == 0 && s_slow) atomicAdd(p.stats, static_cast<unsigned long long>(s_slow));
}

// ---------------------------------------------------------------------------------------------
// probe kernel: (1) marginal histograms of a strided probe of the block -> the densest hyper-rectangle
// of at most `budget` bins (XHK_WINDOW); (2) the fixed-point scale of the weights.  One CTA, ~10-20 us;
// run once per call when the bin space exceeds shared memory or weights are present.
// ---------------------------------------------------------------------------------------------
template <typename T, int KT>
__global__ void __launch_bounds__(kMaxThreads, 1) k_window(const __grid_constant__ XhkParams p, XhkWindow* wout, int budget,
                                                           int budget32, int n_probe) {
  constexpr int KMAX = KT ? KT : XHK_MAX_VARS;
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ int s_moff[XHK_MAX_VARS + 1];
  __shared__ unsigned long long s_wmax;
  __shared__ int s_fail, s_seen, s_fail32, s_budget;
  __shared__ double s_mul, s_limit, s_mul32;
  const int K = KT ? KT : p.n_vars, tid = threadIdx.x, nthr = blockDim.x;
  T* sedges = reinterpret_cast<T*>(smem);
  const size_t edges_al = (static_cast<size_t>(p.n_edges_total) * sizeof(T) + 15) & ~static_cast<size_t>(15);
  unsigned short* slut = reinterpret_cast<unsigned short*>(smem + edges_al);
  unsigned int* marg = reinterpret_cast<unsigned int*>(smem + edges_al + ((static_cast<size_t>(p.n_lut_total) * 2 + 15) & ~static_cast<size_t>(15)));
  for (int i = tid; i < p.n_edges_total; i += nthr) sedges[i] = static_cast<const T*>(p.edges)[i];
  for (int i = tid; i < p.n_lut_total; i += nthr) slut[i] = p.lut[i];
  if (tid == 0) { int o = 0; for (int k = 0; k < K; ++k) { s_moff[k] = o; o += p.nb[k]; } s_moff[K] = o; s_wmax = 0ull; s_fail = 0; s_seen = 0; s_fail32 = 0; s_budget = budget; }
  __syncthreads();
  const int mtot = s_moff[K];
  for (int i = tid; i < mtot; i += nthr) marg[i] = 0u;
  __syncthreads();
  const long long total = p.M * p.N;
  // probe positions: 256-sample contiguous runs spread evenly over the block (few pages touched: far-apart
  // single-element probes cost ~100 us of TLB misses on multi-GB inputs)
  const long long n_runs = (n_probe + 255) / 256;
  const long long run_stride = total / n_runs;
  auto probe_pos = [&](long long i) -> long long { return (i >> 8) * run_stride + (i & 255); };
  const bool flat = p.M == 1;
  auto locate = [&](long long pos, long long& r, long long& c) {
    if (flat) { r = 0; c = pos; } else { r = pos / p.N; c = pos - r * p.N; }
  };
  auto wload = [&](long long r, long long c) -> double {
    return p.w_dtype == 1 ? static_cast<double>((static_cast<const float*>(p.w) + r * p.wstride)[c])
                          : (static_cast<const double*>(p.w) + r * p.wstride)[c];
  };
  constexpr int PB = 8;  // probe positions in flight per thread
  unsigned long long wmx = 0ull;
  double wkeep[PB]; int nkeep = 0;        // this thread's probe weights (first batch; enough for the statistic)
  for (long long i0 = tid; i0 < n_probe; i0 += static_cast<long long>(PB) * nthr) {
    T xs[PB][KMAX]; double ws[PB]; bool have[PB];
#pragma unroll
    for (int b = 0; b < PB; ++b) {
      const long long i = i0 + static_cast<long long>(b) * nthr;
      const long long pos = probe_pos(i);
      have[b] = i < n_probe && pos < total;
      ws[b] = 0.0;
      if (have[b]) {
        long long r, c; locate(pos, r, c);
#pragma unroll
        for (int k = 0; k < KMAX; ++k) if (k < K) xs[b][k] = (static_cast<const T*>(p.data[k]) + r * p.stride[k])[c];
        if (p.w_dtype != 0) ws[b] = wload(r, c);
      }
    }
#pragma unroll
    for (int b = 0; b < PB; ++b) {
      if (have[b]) {
        int j[KMAX]; bool ok = true;
#pragma unroll
        for (int k = 0; k < KMAX; ++k) if (k < K) { j[k] = exact_bin_inline<T>(p, k, sedges, slut, xs[b][k]); ok = ok && j[k] >= 0; }
        if (ok) {
#pragma unroll
          for (int k = 0; k < KMAX; ++k) if (k < K) atomicAdd(&marg[s_moff[k] + j[k]], 1u);
        }
        if (p.w_dtype != 0) {
          const double a = fabs(ws[b]);
          if (a < INFINITY) { const unsigned long long bits = static_cast<unsigned long long>(__double_as_longlong(a)); if (bits > wmx) wmx = bits; }
          if (i0 == tid) { wkeep[b] = ws[b]; nkeep = b + 1; }
        }
      }
    }
  }
  if (p.w_dtype != 0) {   // non-negative doubles order like their bit patterns; one shared atomic per warp
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long t = __shfl_xor_sync(0xffffffffu, wmx, o); if (t > wmx) wmx = t; }
    if ((tid & 31) == 0) atomicMax(&s_wmax, wmx);
  }
  __syncthreads();
  // ---- weights: scale from the largest finite probe weight; count the probe weights that would not be
  //      exact at that scale; too many (> 1/64) -> float64 shared adds instead of fixed point
  if (p.w_dtype != 0) {
    if (tid == 0) {
      const double m = __longlong_as_double(static_cast<long long>(s_wmax));
      const int e = m > 0.0 ? ilogb(m) : 0;
      const int sh = p.fx_vbits - 3 - e;          // v = w * 2^sh ; the largest probe weight maps below 2^(vbits-2)
      s_mul = ldexp(1.0, sh); s_limit = ldexp(1.0, p.fx_vbits);
      s_seen = (p.w_dtype == 1 && (sh > 100 || sh < -100)) ? -(1 << 30) : 0;   // outside fp32's exact power-of-two range
      // fx32 candidate: the largest probe weight maps below 2^24 (one u32 limb per bin; k_hist<W = 3>)
      const int sh32 = 23 - e;
      s_mul32 = (budget32 > 0 && p.w_dtype == 1 && m > 0.0 && sh32 <= 100 && sh32 >= -100) ? ldexp(1.0, sh32) : 0.0;
    }
    __syncthreads();
    int fails = 0, fails32 = 0;
#pragma unroll
    for (int i = 0; i < PB; ++i) {
      if (i < nkeep) {
        const double vs = wkeep[i] * s_mul;
        if (!((static_cast<double>(__double2ll_rn(vs)) == vs) & (fabs(vs) < s_limit))) ++fails;
        const double v32 = wkeep[i] * s_mul32;      // non-negative integer below 2^25 (NaN compares false)
        if (!((v32 >= 0.0) & (v32 < 33554432.0) & (static_cast<double>(__double2ll_rn(v32)) == v32))) ++fails32;
      }
    }
    int seen = nkeep;
    for (int o = 16; o > 0; o >>= 1) {
      fails += __shfl_xor_sync(0xffffffffu, fails, o); fails32 += __shfl_xor_sync(0xffffffffu, fails32, o);
      seen += __shfl_xor_sync(0xffffffffu, seen, o);
    }
    if ((tid & 31) == 0 && seen) { atomicAdd(&s_fail, fails); atomicAdd(&s_fail32, fails32); atomicAdd(&s_seen, seen); }
    __syncthreads();
    if (tid == 0) {
      if (s_mul32 > 0.0 && s_seen > 0 && static_cast<long long>(s_fail32) * 64 <= s_seen) {
        wout->fx_mul = s_mul32; wout->fx_unmul = 1.0 / s_mul32; wout->fx_limit = 33554432.0;
        wout->fx_ok = 1; wout->fx_mode = 32;
        s_budget = budget32;                 // 4-byte bins: the window may hold twice as many
      } else {
        wout->fx_mul = s_mul; wout->fx_unmul = 1.0 / s_mul; wout->fx_limit = s_limit;
        wout->fx_ok = (s_seen > 0 && static_cast<long long>(s_fail) * 64 <= s_seen) ? 1 : 0;
        wout->fx_mode = wout->fx_ok ? 64 : 0;
      }
    }
  } else if (tid == 0) { wout->fx_ok = 0; wout->fx_mode = 0; wout->fx_mul = 0.0; wout->fx_unmul = 0.0; wout->fx_limit = 0.0; }
  __syncthreads();
  budget = s_budget;

  // ---- window: warp 0 bisects a density threshold (density of slice s of variable k = marg * nb_k, equal for
  //      all slices of a uniform distribution); the box of variable k spans the slices at or above the
  //      threshold.  Then lane 0 grows the box greedily while it fits the budget.
  if (tid < 32) {
    const unsigned full = 0xffffffffu;
    auto box_at = [&](unsigned long long th, int* lo, int* len) -> long long {
      long long vol = 1;
      for (int k = 0; k < K; ++k) {
        int first = 0x7fffffff, last = -1;
        for (int s = tid; s < p.nb[k]; s += 32)
          if (static_cast<unsigned long long>(marg[s_moff[k] + s]) * p.nb[k] >= th) { first = min(first, s); last = max(last, s); }
        for (int o = 16; o > 0; o >>= 1) { first = min(first, __shfl_xor_sync(full, first, o)); last = max(last, __shfl_xor_sync(full, last, o)); }
        if (last < 0) { first = 0; last = 0; }   // fixed up below (arg max)
        lo[k] = first; len[k] = last - first + 1;
        vol *= len[k]; if (vol > (1ll << 40)) vol = 1ll << 40;
      }
      return vol;
    };
    unsigned long long dmax = 0;
    for (int k = 0; k < K; ++k)
      for (int s = tid; s < p.nb[k]; s += 32) { const unsigned long long d = static_cast<unsigned long long>(marg[s_moff[k] + s]) * p.nb[k]; if (d > dmax) dmax = d; }
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long t = __shfl_xor_sync(full, dmax, o); if (t > dmax) dmax = t; }
    int lo[XHK_MAX_VARS], len[XHK_MAX_VARS];
    unsigned long long tlo = 0, thi = dmax + 1;           // vol(tlo) > budget (unless everything fits), vol(thi) <= budget
    if (box_at(0, lo, len) > budget) {
      while (thi - tlo > 1) {
        const unsigned long long mid = tlo + (thi - tlo) / 2;
        if (box_at(mid, lo, len) <= budget) thi = mid; else tlo = mid;
      }
      long long vol = box_at(thi, lo, len);
      if (tid == 0) {
        // a variable with no slice at the threshold keeps its fullest slice
        for (int k = 0; k < K; ++k) {
          bool any = false;
          for (int s = lo[k]; s < lo[k] + len[k]; ++s) any = any || static_cast<unsigned long long>(marg[s_moff[k] + s]) * p.nb[k] >= thi;
          if (!any) { int arg = 0; unsigned best = 0; for (int s = 0; s < p.nb[k]; ++s) if (marg[s_moff[k] + s] > best) { best = marg[s_moff[k] + s]; arg = s; } lo[k] = arg; len[k] = 1; }
        }
        vol = 1; for (int k = 0; k < K; ++k) vol *= len[k];
        while (vol > budget) {   // cannot happen for a consistent threshold; kept as a guard
          int kb = 0; for (int k = 1; k < K; ++k) if (len[k] > len[kb]) kb = k;
          len[kb] -= 1; vol = 1; for (int k = 0; k < K; ++k) vol *= len[k];
        }
        while (true) {           // greedy growth: neighbouring slice with the most probe mass per added bin
          int bk = -1, bside = 0; double bgain = -1.0;
          for (int k = 0; k < K; ++k) {
            long long nv = 1; for (int q = 0; q < K; ++q) nv *= (q == k ? len[q] + 1 : len[q]);
            if (nv > budget) continue;
            if (lo[k] > 0) { const double g = (static_cast<double>(marg[s_moff[k] + lo[k] - 1]) + 1e-3) * len[k]; if (g > bgain) { bgain = g; bk = k; bside = -1; } }
            if (lo[k] + len[k] < p.nb[k]) { const double g = (static_cast<double>(marg[s_moff[k] + lo[k] + len[k]]) + 1e-3) * len[k]; if (g > bgain) { bgain = g; bk = k; bside = 1; } }
          }
          if (bk < 0) break;
          if (bside < 0) lo[bk] -= 1;
          len[bk] += 1;
        }
      }
    }
    if (tid == 0)
      for (int k = 0; k < XHK_MAX_VARS; ++k) { wout->lo[k] = k < K ? lo[k] : 0; wout->len[k] = k < K ? len[k] : 0; }
  }
}

// ---------------------------------------------------------------------------------------------
// column layout: reductions over LEADING axes.  The arrays are (n_outer, N, inner) C-contiguous blocks and the
// histogram is taken over N for every (a, m).  One thread owns one kept column m: it walks down the reduced axis
// (adjacent threads read adjacent addresses: coalesced) and accumulates into a thread-PRIVATE histogram
// hist[bin][thread] in shared memory — plain read-modify-write, no atomics, conflict-free banks; weighted sums are
// added in sample order like np.bincount.  The reference has to copy for this layout (np.moveaxis + reshape,
// core.py:218-226).  grid = (n_outer * column tiles, nsplit slices of the reduced axis).
// ---------------------------------------------------------------------------------------------
template <typename T, int W, int KT>
__global__ void __launch_bounds__(kMaxThreads, 1) k_hist_cols(const __grid_constant__ XhkParams p, long long inner, int tm, int accumulate) {
  using HT = typename std::conditional<W == 0, unsigned int, double>::type;
  using OT = typename std::conditional<W == 0, unsigned long long, double>::type;
  using WT = typename WType<W>::type;
  constexpr int KMAX = KT ? KT : XHK_MAX_VARS;
  extern __shared__ __align__(16) unsigned char smem[];
  const int K = KT ? KT : p.n_vars;
  const int tid = threadIdx.x, nthr = blockDim.x;
  T* sedges = reinterpret_cast<T*>(smem);
  const size_t edges_al = (static_cast<size_t>(p.n_edges_total) * sizeof(T) + 15) & ~static_cast<size_t>(15);
  unsigned short* slut = reinterpret_cast<unsigned short*>(smem + edges_al);
  HT* hist = reinterpret_cast<HT*>(smem + edges_al + ((static_cast<size_t>(p.n_lut_total) * 2 + 15) & ~static_cast<size_t>(15)));
  for (int i = tid; i < p.n_edges_total; i += nthr) sedges[i] = static_cast<const T*>(p.edges)[i];
  for (int i = tid; i < p.n_lut_total; i += nthr) slut[i] = p.lut[i];
  const int B = static_cast<int>(p.B);
  for (int i = tid; i < B * tm; i += nthr) hist[i] = HT(0);
  __syncthreads();

  const long long tiles = (inner + tm - 1) / tm;
  const long long a = blockIdx.x / tiles;
  const long long m0 = (blockIdx.x - a * tiles) * tm;
  const long long m = m0 + tid;
  const long long N = p.N;
  const long long n0 = N * blockIdx.y / gridDim.y, n1 = N * (blockIdx.y + 1ll) / gridDim.y;
  if (tid < tm && m < inner) {
    const T* px[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) px[k] = (k < K) ? static_cast<const T*>(p.data[k]) + a * N * inner + m : nullptr;
    const WT* pw = (W != 0) ? static_cast<const WT*>(p.w) + a * N * inner + m : nullptr;
    auto one = [&](const T (&x)[KMAX], WT wv) {
      int bin = 0; bool ok = true;
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        if (k < K) {
          const int j = exact_bin_inline<T>(p, k, sedges, slut, x[k]);
          ok = ok && j >= 0;
          bin = bin * p.nb[k] + j;
        }
      }
      if (ok) {
        HT* h = hist + static_cast<size_t>(bin) * tm + tid;
        if constexpr (W == 0) *h += 1u; else *h += static_cast<double>(wv);
      }
    };
    // rows of the reduced axis in flight per thread.  The loads are one element wide, so the bytes in flight — what
    // hides the DRAM latency here — are UN * (bytes per sample) per thread: 8 rows of a 4-byte record measured 0.35
    // of the HBM peak, 8 rows of an 8-byte record 0.62; hence up to 16 rows, aiming at 128 B per thread.
    constexpr int CREC = static_cast<int>(sizeof(T)) * KMAX + (W == 0 ? 0 : static_cast<int>(sizeof(WT)));
    constexpr int UN = CREC <= 8 ? 16 : CREC <= 16 ? 8 : 4;
    long long n = n0;
    for (; n + UN <= n1; n += UN) {
      T xv[UN][KMAX]; WT wv[UN];
      // (no L2 prefetch here: the kernel is issue-bound and the extra instructions cost 4-7 %, measured)
#pragma unroll
      for (int u = 0; u < UN; ++u) {
#pragma unroll
        for (int k = 0; k < KMAX; ++k) if (k < K) xv[u][k] = __ldcs(px[k] + (n + u) * inner);
        wv[u] = WT(1);
        if constexpr (W != 0) wv[u] = __ldcs(pw + (n + u) * inner);
      }
#pragma unroll
      for (int u = 0; u < UN; ++u) one(xv[u], wv[u]);
    }
    for (; n < n1; ++n) {
      T x[KMAX]; WT wv = WT(1);
#pragma unroll
      for (int k = 0; k < KMAX; ++k) if (k < K) x[k] = px[k][n * inner];
      if constexpr (W != 0) wv = pw[n * inner];
      one(x, wv);
    }
  }
  __syncthreads();
  // flush: consecutive threads write consecutive bins of one output row (coalesced); out row = a*inner + m
  OT* out = static_cast<OT*>(p.out) + (a * inner + m0) * B;
  const long long cols = (inner - m0 < tm) ? inner - m0 : tm;
  for (long long i = tid; i < cols * B; i += nthr) {
    const int ml = static_cast<int>(i / B), b = static_cast<int>(i - static_cast<long long>(ml) * B);
    const HT v = hist[static_cast<size_t>(b) * tm + ml];
    if (accumulate) { if (v != HT(0)) atomicAdd(out + i, static_cast<OT>(v)); }
    else out[i] = static_cast<OT>(v);
  }
}

// ---------------------------------------------------------------------------------------------
// several weight arrays over the SAME samples in one pass (the reference's tutorial computes a weighted mean
// as histogram(weights=w*a) / histogram(weights=w), two passes over identical samples, and lists "allow list of weights"
// as a TODO, xarray.py:106).  The samples are read and classified once; every weight array has its own plane of float64
// bins — [nw][B] in shared memory when that fits (red.shared.add.f64), else straight global REDs.  Work items are
// (row, chunk of the reduced axis); a CTA walks a contiguous run of items and flushes its planes whenever the row changes.
// out: [nw][M][B] float64, zero-filled by the host.
// ---------------------------------------------------------------------------------------------
template <typename T, typename WT, int KT>
__global__ void __launch_bounds__(kMaxThreads, 1) k_hist_mw(const __grid_constant__ XhkParams p, const __grid_constant__ XhkMultiWeights m) {
  constexpr int KMAX = KT ? KT : XHK_MAX_VARS;
  extern __shared__ __align__(16) unsigned char smem[];
  const int K = KT ? KT : p.n_vars, tid = threadIdx.x, nthr = blockDim.x;
  T* sedges = reinterpret_cast<T*>(smem);
  const size_t edges_al = (static_cast<size_t>(p.n_edges_total) * sizeof(T) + 15) & ~static_cast<size_t>(15);
  unsigned short* slut = reinterpret_cast<unsigned short*>(smem + edges_al);
  double* hist = reinterpret_cast<double*>(smem + edges_al + ((static_cast<size_t>(p.n_lut_total) * 2 + 15) & ~static_cast<size_t>(15)));
  for (int i = tid; i < p.n_edges_total; i += nthr) sedges[i] = static_cast<const T*>(p.edges)[i];
  for (int i = tid; i < p.n_lut_total; i += nthr) slut[i] = p.lut[i];
  const int B = static_cast<int>(p.B), nw = m.nw;
  const int hbins = m.use_smem ? nw * B : 0;
  for (int i = tid; i < hbins; i += nthr) hist[i] = 0.0;
  __syncthreads();
  const unsigned sh = static_cast<unsigned>(__cvta_generic_to_shared(hist));
  double* const out = static_cast<double*>(p.out);
  const long long plane = m.plane;
  const long long items_per_row = (p.N + m.chunk - 1) / m.chunk, items = p.M * items_per_row;
  const long long i0 = items * blockIdx.x / gridDim.x, i1 = items * (blockIdx.x + 1ll) / gridDim.x;
  long long cur_row = -1;
  auto flush = [&](long long r) {
    __syncthreads();
    for (int i = tid; i < hbins; i += nthr) {
      const double v = hist[i];
      if (v != 0.0) { const int q = i / B; atomicAdd(out + q * plane + r * p.B + (i - q * B), v); hist[i] = 0.0; }   // (NaN != 0: flushed)
    }
    __syncthreads();
  };
  auto one = [&](const T (&x)[KMAX], const WT (&w)[XHK_MAX_WEIGHTS], long long r) {
    int bin = 0; bool ok = true;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      if (k < K) {
        int j;
        bool fast = false;
        if constexpr (std::is_floating_point<T>::value) {
          if (p.uniform[k]) {
            // evenly spaced edges: the bin from the mantissa of r + magic, certain iff |frac - 0.5| <= chalf (see k_hist)
            const T r = fma_t(x[k] - Consts<T>::get(p, k, XHK_C_E0), Consts<T>::get(p, k, XHK_C_INV), T(-0.5));
            T jf; int jraw;
            RoundSplit<T>::run(r, jf, jraw);
            j = jraw - RoundSplit<T>::kBias;
            fast = (fabs(r - jf) <= Consts<T>::get(p, k, XHK_C_CHALF)) & (static_cast<unsigned>(j) < static_cast<unsigned>(p.nb[k])) & RoundSplit<T>::ok(r);
          }
        }
        if (!fast) j = exact_bin_inline<T>(p, k, sedges, slut, x[k]);
        ok = ok && j >= 0;
        bin = bin * p.nb[k] + j;
      }
    }
    if (!ok) return;
    if (m.use_smem) { for (int q = 0; q < nw; ++q) reds_add_f64(sh + 8u * static_cast<unsigned>(q * B + bin), static_cast<double>(w[q])); }
    else { for (int q = 0; q < nw; ++q) atomicAdd(out + q * plane + r * p.B + bin, static_cast<double>(w[q])); }
  };
  for (long long it = i0; it < i1; ++it) {
    const long long r = it / items_per_row, c0 = (it - r * items_per_row) * m.chunk;
    const long long c1 = c0 + m.chunk < p.N ? c0 + m.chunk : p.N;
    if (r != cur_row) { if (cur_row >= 0 && hbins) flush(cur_row); cur_row = r; }
    const T* px[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) px[k] = (k < K) ? static_cast<const T*>(p.data[k]) + r * p.stride[k] : nullptr;
    const WT* pw[XHK_MAX_WEIGHTS];
#pragma unroll
    for (int q = 0; q < XHK_MAX_WEIGHTS; ++q) pw[q] = (q < nw) ? static_cast<const WT*>(m.w[q]) + r * p.wstride : nullptr;
    bool vec = (KT != 0) && ((c0 & 3) == 0);
#pragma unroll
    for (int k = 0; k < KMAX; ++k) if (k < K) vec = vec && ((reinterpret_cast<uintptr_t>(px[k] + c0) & 15) == 0);
    for (int q = 0; q < nw; ++q) vec = vec && ((reinterpret_cast<uintptr_t>(pw[q] + c0) & 15) == 0);
    long long c = c0;
    if (vec) {
      const long long nv = (c1 - c0) >> 2;
      for (long long g = tid; g < nv; g += nthr) {
        T xv[KMAX][4]; WT wv[XHK_MAX_WEIGHTS][4];
#if XH_PREFETCH_DIST > 0
        {
          const long long gp = g + static_cast<long long>(XH_PREFETCH_DIST) * nthr;      // (may run into the next chunk of the row: harmless)
          if (p.prefetch && c0 + 4 * gp + 4 <= p.N) {
#pragma unroll
            for (int k = 0; k < KMAX; ++k) if (k < K) prefetch_l2(px[k] + c0 + 4 * gp);
#pragma unroll
            for (int q = 0; q < XHK_MAX_WEIGHTS; ++q) if (q < nw) prefetch_l2(pw[q] + c0 + 4 * gp);
          }
        }
#endif
#pragma unroll
        for (int k = 0; k < KMAX; ++k) if (k < K) load4(px[k] + c0, g, xv[k]);
#pragma unroll
        for (int q = 0; q < XHK_MAX_WEIGHTS; ++q) if (q < nw) load4(pw[q] + c0, g, wv[q]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          T x[KMAX]; WT w[XHK_MAX_WEIGHTS];
#pragma unroll
          for (int k = 0; k < KMAX; ++k) x[k] = (k < K) ? xv[k][e] : T(0);
#pragma unroll
          for (int q = 0; q < XHK_MAX_WEIGHTS; ++q) w[q] = (q < nw) ? wv[q][e] : WT(0);
          one(x, w, r);
        }
      }
      c = c0 + (nv << 2);
    }
    for (long long i = c + tid; i < c1; i += nthr) {
      T x[KMAX]; WT w[XHK_MAX_WEIGHTS];
#pragma unroll
      for (int k = 0; k < KMAX; ++k) x[k] = (k < K) ? px[k][i] : T(0);
#pragma unroll
      for (int q = 0; q < XHK_MAX_WEIGHTS; ++q) w[q] = (q < nw) ? pw[q][i] : WT(0);
      one(x, w, r);
    }
  }
  if (cur_row >= 0 && hbins) flush(cur_row);
}

template <typename T, typename WT>
XhkMwKernel pick_mw_k(int K) {
  switch (K) {
    case 1: return k_hist_mw<T, WT, 1>;
    case 2: return k_hist_mw<T, WT, 2>;
    case 3: return k_hist_mw<T, WT, 3>;
    default: return k_hist_mw<T, WT, 0>;
  }
}

template <typename T, int W, int MODE>
XhkHistKernel pick_k(int K) {
  switch (K) {
    case 1: return k_hist<T, W, 1, MODE>;
    case 2: return k_hist<T, W, 2, MODE>;
    case 3: return k_hist<T, W, 3, MODE>;
    case 4: return k_hist<T, W, 4, MODE>;
    default: return k_hist<T, W, 0, 0>;
  }
}
template <typename T, int W>
XhkHistKernel pick_m(int K, int mode) {
  if constexpr (std::is_floating_point<T>::value) {   // int64 data: general kernel only
    if (mode == 1) return pick_k<T, W, 1>(K);
    if (mode == 2) return pick_k<T, W, 2>(K);
    if (mode == 3) return pick_k<T, W, 3>(K);
    if (mode == 5) { if constexpr (W == 3) return pick_k<T, W, 5>(K); else return pick_k<T, W, 1>(K); }
  }
  return pick_k<T, W, 0>(K);
}


// packed counts (W = 4): fused fast path only (every variable uniform, no row tiling), floating-point data
template <typename T>
XhkHistKernel pick_pk(int K) {
  if constexpr (std::is_floating_point<T>::value) return pick_k<T, 4, 1>(K);
  else return nullptr;
}

// fx32 sibling (W = 3): floating-point data only (the host never asks for it with int64 data)
template <typename T>
XhkHistKernel pick_fx32(int K, int mode) {
  if constexpr (std::is_floating_point<T>::value) return pick_m<T, 3>(K, mode);
  else return nullptr;
}

template <typename T>
XhkWindowKernel pick_window_t(int K) {
  switch (K) {
    case 1: return k_window<T, 1>;
    case 2: return k_window<T, 2>;
    case 3: return k_window<T, 3>;
    case 4: return k_window<T, 4>;
    default: return k_window<T, 0>;
  }
}

template <typename T, int W>
XhkColsKernel pick_cols_k(int K) {
  switch (K) {
    case 1: return k_hist_cols<T, W, 1>;
    case 2: return k_hist_cols<T, W, 2>;
    case 3: return k_hist_cols<T, W, 3>;
    case 4: return k_hist_cols<T, W, 4>;
    default: return k_hist_cols<T, W, 0>;
  }
}
template <typename T>
XhkColsKernel pick_cols_w(int w, int K) { return w == 0 ? pick_cols_k<T, 0>(K) : w == 1 ? pick_cols_k<T, 1>(K) : pick_cols_k<T, 2>(K); }

}  // namespace

// one definition per data type (DT = f32 / f64 / i64)
#define XHK_DEFINE_PICKERS(DT, T, ALLOW_FAST)                                                          \
  XhkHistKernel xhk_pick_hist_##DT(int w, int K, int mode) {                                           \
    if (!(ALLOW_FAST)) mode = 0;                                                                        \
    if (w == 3) return pick_fx32<T>(K, mode);                                                           \
    if (w == 4) return pick_pk<T>(K);                                                                   \
    return w == 0 ? pick_m<T, 0>(K, mode) : w == 1 ? pick_m<T, 1>(K, mode) : pick_m<T, 2>(K, mode);     \
  }                                                                                                     \
  XhkWindowKernel xhk_pick_window_##DT(int K) { return pick_window_t<T>(K); }                           \
  XhkColsKernel xhk_pick_cols_##DT(int w, int K) { return pick_cols_w<T>(w, K); }                      \
  XhkMwKernel xhk_pick_mw_##DT(int w, int K) { return w == 1 ? pick_mw_k<T, float>(K) : pick_mw_k<T, double>(K); }
