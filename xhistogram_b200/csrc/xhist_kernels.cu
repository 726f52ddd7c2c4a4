// xhist_kernels.cu — hand-written sm_100a kernels of the histogram hot path.
//
// What the reference does per block (xhistogram/core.py:137-194, 73-83): per variable a
// searchsorted + right-edge fix (core.py:163-174), a ravel_multi_index (core.py:178-181), a
// row-offset np.bincount (core.py:80-81) and a slice that drops the under/overflow cells
// (core.py:191-192) — 5 numpy passes and 4 full-size temporaries per variable.  Here it is ONE
// persistent kernel: coalesced 16-byte streaming loads of the samples, per-sample classification
// against the edges staged in shared memory, accumulation into a privatised per-CTA
// shared-memory histogram, and a flush to global memory (plain stores for rows a CTA owns,
// RED atomics otherwise).  No tensor cores: this is a scatter/reduce, HBM-read bound.
//
// Design numbers (profiles/r1_microbench_mechanisms.log, B200): streaming 12 B/sample reaches
// ~7.0 TB/s; shared u32 atomics are free next to the stream (~590 Gsamples/s); shared f64 CAS adds
// reach ~340 Gsamples/s; global RED only 88 Gsamples/s — hence everything that can be privatised is.
// Source layout: the kernel templates live in xhist_kernels_impl.cuh and are instantiated per data type in
// xhist_k_f32.cu / xhist_k_f64.cu / xhist_k_i64.cu (parallel compilation); this file holds the small utility
// kernels and the host-callable launchers.
#include "xhist_kernels.cuh"
#include <algorithm>
#include <type_traits>

namespace {

// zero the rows of `out` that are shared between CTAs under the sample partition (XHK_FULL only)
template <typename OT>
__global__ void k_zero_shared_rows(const __grid_constant__ XhkParams p, int hist_grid) {
  const long long total = p.M * p.N;
  long long row = -1;
  if (p.M == 1) { if (blockIdx.x == 0) row = 0; }
  else if (blockIdx.x >= 1 && blockIdx.x < hist_grid) {
    const long long s = blockIdx.x * p.per_cta;
    if (s < total && (s % p.N) != 0) {
      row = s / p.N;
      const long long sp = (blockIdx.x - 1ll) * p.per_cta;         // previous boundary
      if (sp > row * p.N) row = -1;                                 // an earlier boundary already clears this row
    }
  }
  if (row < 0) return;
  OT* o = static_cast<OT*>(p.out) + row * p.B;
  for (long long b = threadIdx.x; b < p.B; b += blockDim.x) o[b] = OT(0);
}

// ---------------------------------------------------------------------------------------------
// utilities: counter-based synthetic data, min/max, L2 flush
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

template <typename T, int NORMAL>
__global__ void k_fill(T* p, long long n, unsigned long long seed, long long offset) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const unsigned long long h = mix64(mix64(seed) ^ static_cast<unsigned long long>(offset + i));
    if (NORMAL) {
      // Box-Muller on two 24/32-bit uniforms; only needs to be reproducible on the device
      const double u1 = (static_cast<double>(h >> 32) + 1.0) * (1.0 / 4294967297.0);
      const double u2 = static_cast<double>(h & 0xFFFFFFFFull) * (1.0 / 4294967296.0);
      p[i] = static_cast<T>(sqrt(-2.0 * log(u1)) * cospi(2.0 * u2));
    } else {
      if (sizeof(T) == 4) p[i] = static_cast<T>(static_cast<float>(h >> 40) * (1.0f / 16777216.0f));
      else p[i] = static_cast<T>(static_cast<double>(h >> 11) * (1.0 / 9007199254740992.0));
    }
  }
}

template <typename T>
__global__ void k_minmax(const T* __restrict__ d, long long n, double* out /* [grid][3] */) {
  double mn = INFINITY, mx = -INFINITY; int nan = 0;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double v = static_cast<double>(d[i]);
    if (v != v) nan = 1; else { mn = fmin(mn, v); mx = fmax(mx, v); }
  }
  __shared__ double smn[32], smx[32]; __shared__ int snan[32];
  for (int o = 16; o > 0; o >>= 1) {
    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    nan |= __shfl_xor_sync(0xffffffffu, nan, o);
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { smn[w] = mn; smx[w] = mx; snan[w] = nan; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < (blockDim.x >> 5); ++i) { mn = fmin(mn, smn[i]); mx = fmax(mx, smx[i]); nan |= snan[i]; }
    out[3 * blockIdx.x + 0] = mn; out[3 * blockIdx.x + 1] = mx; out[3 * blockIdx.x + 2] = nan ? 1.0 : 0.0;
  }
}

__global__ void k_flush(uint4* buf, size_t n16) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n16; i += stride) buf[i] = make_uint4(i, 0, 0, 0);
}

// ---------------------------------------------------------------------------------------------
// density (core.py:444-462): h = counts / bin_areas / rowsum, in that order, in float64.  bin_areas is the outer
// product of the per-variable widths as numpy forms it with functools.reduce(np.multiply.outer, widths): a product of
// two float32 operands is rounded to float32, anything else is a float64 product.  Counts arrive as int64 and are
// rewritten in place as float64.  One warp per row (B <= 1024) or one CTA per row.
// ---------------------------------------------------------------------------------------------
struct XhkDensity {
  const double* widths;              // device, all variables concatenated
  int off[XHK_MAX_VARS];
  int nb[XHK_MAX_VARS];
  int f32[XHK_MAX_VARS];
  int K;
};

__device__ __forceinline__ double density_area(const XhkDensity& q, long long b) {
  int idx[XHK_MAX_VARS];
  for (int k = q.K - 1; k >= 0; --k) { const long long t = b / q.nb[k]; idx[k] = static_cast<int>(b - t * q.nb[k]); b = t; }
  double acc = q.widths[q.off[0] + idx[0]];
  bool is32 = q.f32[0] != 0;
  for (int k = 1; k < q.K; ++k) {
    const double wk = q.widths[q.off[k] + idx[k]];
    if (is32 && q.f32[k]) acc = static_cast<double>(__fmul_rn(static_cast<float>(acc), static_cast<float>(wk)));
    else { acc = __dmul_rn(acc, wk); is32 = false; }
  }
  return acc;
}

template <bool COUNTS>
__device__ __forceinline__ double density_load(const void* row, long long b) {
  if (COUNTS) return static_cast<double>(static_cast<const long long*>(row)[b]);
  return static_cast<const double*>(row)[b];
}

// rows of at most 1024 bins: one warp per row, sum and scale in one kernel
template <bool COUNTS>
__global__ void __launch_bounds__(256) k_density_rows(void* out, long long M, long long B, const __grid_constant__ XhkDensity q) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (long long r = static_cast<long long>(blockIdx.x) * nwarp + warp; r < M; r += static_cast<long long>(gridDim.x) * nwarp) {
    unsigned char* row = static_cast<unsigned char*>(out) + static_cast<size_t>(r) * B * 8;
    double fs = 0.0; long long is = 0;     // int64 counts add exactly; float64 sums in a fixed tree order
    for (long long b = lane; b < B; b += 32) { if (COUNTS) is += reinterpret_cast<const long long*>(row)[b]; else fs += reinterpret_cast<const double*>(row)[b]; }
    for (int o = 16; o > 0; o >>= 1) { fs += __shfl_xor_sync(0xffffffffu, fs, o); is += __shfl_xor_sync(0xffffffffu, is, o); }
    const double total = COUNTS ? static_cast<double>(is) : fs;
    for (long long b = lane; b < B; b += 32) {
      const double v = density_load<COUNTS>(row, b);
      reinterpret_cast<double*>(row)[b] = __ddiv_rn(__ddiv_rn(v, density_area(q, b)), total);
    }
  }
}

// longer rows: (1) chunk sums added into sums[r] (u64 for counts: exact; float64 otherwise), (2) flat scaling pass
constexpr int kDensityChunk = 2048;
template <bool COUNTS>
__global__ void __launch_bounds__(256) k_density_sum(const void* out, long long B, int chunks, void* sums) {
  __shared__ double s_f[8];
  __shared__ long long s_i[8];
  const long long r = blockIdx.x / chunks;
  const long long b0 = static_cast<long long>(blockIdx.x % chunks) * kDensityChunk;
  const long long b1 = b0 + kDensityChunk < B ? b0 + kDensityChunk : B;
  const unsigned char* row = static_cast<const unsigned char*>(out) + static_cast<size_t>(r) * B * 8;
  double fs = 0.0; long long is = 0;
  for (long long b = b0 + threadIdx.x; b < b1; b += blockDim.x) { if (COUNTS) is += reinterpret_cast<const long long*>(row)[b]; else fs += reinterpret_cast<const double*>(row)[b]; }
  for (int o = 16; o > 0; o >>= 1) { fs += __shfl_xor_sync(0xffffffffu, fs, o); is += __shfl_xor_sync(0xffffffffu, is, o); }
  if ((threadIdx.x & 31) == 0) { s_f[threadIdx.x >> 5] = fs; s_i[threadIdx.x >> 5] = is; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < (blockDim.x >> 5); ++i) { fs += s_f[i]; is += s_i[i]; }
    if (COUNTS) atomicAdd(static_cast<unsigned long long*>(sums) + r, static_cast<unsigned long long>(is));
    else atomicAdd(static_cast<double*>(sums) + r, fs);
  }
}
template <bool COUNTS>
__global__ void __launch_bounds__(256) k_density_scale(void* out, long long M, long long B, const void* sums, const __grid_constant__ XhkDensity q) {
  const long long n = M * B, stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const long long r = i / B, b = i - r * B;
    const double total = COUNTS ? static_cast<double>(static_cast<const long long*>(sums)[r]) : static_cast<const double*>(sums)[r];
    const double v = density_load<COUNTS>(out, i);
    static_cast<double*>(out)[i] = __ddiv_rn(__ddiv_rn(v, density_area(q, b)), total);
  }
}

// ---------------------------------------------------------------------------------------------
// device transpose: out = C-contiguous copy of transpose(in, perm).  The reference brings non-trailing reduce axes
// to the end with np.moveaxis + reshape on the host (core.py:218-226); for device-resident inputs whose layout neither
// the row nor the column kernel reads in place, this is that copy, done in HBM.
//   k_permute_tiled : the innermost input axis (a) is not the innermost output axis (b): 32 x 32 tiles through shared
//                     memory, reads coalesced along a, writes coalesced along b; every other axis is batch
//   k_permute_rows  : the innermost axis stays innermost: element-wise gather, coalesced on both sides
// ---------------------------------------------------------------------------------------------
struct XhkPermute {
  int nd, da, db;
  long long shape[XHK_MAX_VARS];        // input shape
  long long sin[XHK_MAX_VARS];          // input strides (elements)
  long long sout[XHK_MAX_VARS];         // output stride of INPUT axis i (elements)
  long long oshape[XHK_MAX_VARS];       // output shape
  long long osin[XHK_MAX_VARS];         // input stride of OUTPUT axis j
  long long tiles_a, tiles_b, total;
};

template <typename E>
__global__ void __launch_bounds__(256) k_permute_tiled(const E* __restrict__ in, E* __restrict__ out, const __grid_constant__ XhkPermute q) {
  __shared__ E tile[32][33];
  long long t = blockIdx.x;
  const long long ta = t % q.tiles_a; t /= q.tiles_a;
  const long long tb = t % q.tiles_b; t /= q.tiles_b;
  long long bin = 0, bout = 0;
  for (int i = q.nd - 1; i >= 0; --i) {
    if (i == q.da || i == q.db) continue;
    const long long c = t % q.shape[i]; t /= q.shape[i];
    bin += c * q.sin[i]; bout += c * q.sout[i];
  }
  const long long a0 = ta * 32, b0 = tb * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8)
    if (a0 + tx < q.shape[q.da] && b0 + j < q.shape[q.db]) tile[j][tx] = in[bin + (b0 + j) * q.sin[q.db] + (a0 + tx)];
  __syncthreads();
  for (int j = ty; j < 32; j += 8)
    if (b0 + tx < q.shape[q.db] && a0 + j < q.shape[q.da]) out[bout + (a0 + j) * q.sout[q.da] + (b0 + tx)] = tile[tx][j];
}

template <typename E>
__global__ void __launch_bounds__(256) k_permute_rows(const E* __restrict__ in, E* __restrict__ out, const __grid_constant__ XhkPermute q) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < q.total; i += stride) {
    long long r = i, off = 0;
    for (int j = q.nd - 1; j >= 0; --j) { const long long c = r % q.oshape[j]; r /= q.oshape[j]; off += c * q.osin[j]; }
    out[i] = in[off];
  }
}

// ---------------------------------------------------------------------------------------------
// peer-memory all-reduce of partial histograms (the role of dask's .sum over chunk partials, core.py:439, for one rank
// per GPU on one NVSwitch node).  The partials are small (512 KB for 256 x 256 bins), so the cost of a collective is
// latency: instead of a ring / tree this is ONE kernel per rank — announce, wait, then read every peer's partial
// straight over NVLink and add in rank order.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

template <typename T>
__global__ void __launch_bounds__(512) k_peer_allreduce(const __grid_constant__ XhkPeerArgs a) {
  const int tid = threadIdx.x;
  if (blockIdx.x == 0 && tid < a.n) {
    __threadfence_system();                       // the partial written by the kernels before this one is visible first
    st_release_sys(a.flags[tid] + a.rank, a.seq);
  }
  if (tid < a.n) {
    const unsigned long long* mine = a.flags[a.rank] + tid;
    while (ld_acquire_sys(mine) < a.seq) { }
  }
  __syncthreads();
  const long long n2 = a.count >> 1;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  using V2 = typename std::conditional<std::is_same<T, double>::value, double2, longlong2>::type;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + tid; i < n2; i += stride) {
    V2 acc = __ldcv(reinterpret_cast<const V2*>(a.slot[0]) + i);
    for (int r = 1; r < a.n; ++r) { const V2 v = __ldcv(reinterpret_cast<const V2*>(a.slot[r]) + i); acc.x += v.x; acc.y += v.y; }
    reinterpret_cast<V2*>(a.out)[i] = acc;
  }
  if ((a.count & 1) && blockIdx.x == 0 && tid == 0) {
    T acc = __ldcv(static_cast<const T*>(a.slot[0]) + a.count - 1);
    for (int r = 1; r < a.n; ++r) acc += __ldcv(static_cast<const T*>(a.slot[r]) + a.count - 1);
    static_cast<T*>(a.out)[a.count - 1] = acc;
  }
}

// small results (M * B * 8 bytes up to a few MB): ONE launch, out of place.  G CTAs share a row; every one of them adds up
// the whole row (the same order in each, so they agree to the bit; the row sits in L2) and scales its own slice into `dst`,
// which may be pinned host memory — the finished density then lands on the host without a separate copy operation.
template <bool COUNTS>
__global__ void __launch_bounds__(1024) k_density_small(const void* src, double* dst, long long B, int G, const __grid_constant__ XhkDensity q) {
  __shared__ double s_f[32];
  __shared__ long long s_i[32];
  const long long r = blockIdx.x / G;
  const int part = blockIdx.x - static_cast<int>(r) * G;
  const unsigned char* row = static_cast<const unsigned char*>(src) + static_cast<size_t>(r) * B * 8;
  double fs = 0.0; long long is = 0;
  for (long long b = threadIdx.x; b < B; b += blockDim.x) { if (COUNTS) is += reinterpret_cast<const long long*>(row)[b]; else fs += reinterpret_cast<const double*>(row)[b]; }
  for (int o = 16; o > 0; o >>= 1) { fs += __shfl_xor_sync(0xffffffffu, fs, o); is += __shfl_xor_sync(0xffffffffu, is, o); }
  if ((threadIdx.x & 31) == 0) { s_f[threadIdx.x >> 5] = fs; s_i[threadIdx.x >> 5] = is; }
  __syncthreads();
  fs = 0.0; is = 0;
  for (int i = 0; i < (blockDim.x >> 5); ++i) { fs += s_f[i]; is += s_i[i]; }
  const double total = COUNTS ? static_cast<double>(is) : fs;
  const long long b0 = B * part / G, b1 = B * (part + 1ll) / G;
  for (long long b = b0 + threadIdx.x; b < b1; b += blockDim.x)
    dst[r * B + b] = __ddiv_rn(__ddiv_rn(density_load<COUNTS>(row, b), density_area(q, b)), total);
}

XhkHistKernel pick(int dtype, int w_dtype, int K, int mode) {
  return dtype == 1 ? xhk_pick_hist_f32(w_dtype, K, mode) : dtype == 2 ? xhk_pick_hist_f64(w_dtype, K, mode) : xhk_pick_hist_i64(w_dtype, K, mode);
}
XhkWindowKernel pick_window(int dtype, int K) {
  return dtype == 1 ? xhk_pick_window_f32(K) : dtype == 2 ? xhk_pick_window_f64(K) : xhk_pick_window_i64(K);
}
XhkColsKernel pick_cols(int dtype, int w, int K) {
  return dtype == 1 ? xhk_pick_cols_f32(w, K) : dtype == 2 ? xhk_pick_cols_f64(w, K) : xhk_pick_cols_i64(w, K);
}

// Mode 2 (branch-free classification for uniform / bounded-step-table variables) handles records above 16 bytes of
// data per sample two samples at a time (half groups) to stay inside the 64-register budget; see k_hist.
int kernel_mode(const XhkParams& p, int dtype, int w = 0) {
  if (dtype == 3) return 0;
  if (p.all_uniform) {
    if (p.tile_rows > 1) return 3;                       // row tiling is a compile-time variant of the fast kernel
    // one-limb weights over a short per-CTA range: dynamic dealing of the groups to the warps (see k_hist, MODE 5)
    // (also for data that spills a lot: warps that run long side loops no longer hold up their CTA — uniform data on config 3:
    //  4.9 -> 3.5 ms per 1e9 samples)
    if (w == 3 && p.partition == XHK_PART_SAMPLES && (p.per_cta <= (5ll << 19) || p.spilly)) return 5;
    return 1;
  }
  return (p.all_branch_free && p.n_vars <= 4) ? 2 : 0;
}


}  // namespace


cudaError_t xhk_set_smem_limits(int max_optin) {
  for (int dt = 1; dt <= 2; ++dt)
    for (int k = 1; k <= 4; ++k) {
      cudaError_t e = cudaFuncSetAttribute(reinterpret_cast<const void*>(pick(dt, 4, k, 1)), cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin);
      if (e != cudaSuccess) return e;
    }
  for (int dt = 1; dt <= 3; ++dt)
    for (int w = 0; w <= (dt == 3 ? 2 : 3); ++w)
      for (int k = 1; k <= 5; ++k)
        for (int f = 0; f <= 5; ++f) {
          if (f == 4) continue;
          cudaError_t e = cudaFuncSetAttribute(reinterpret_cast<const void*>(pick(dt, w, k, f)), cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin);
          if (e != cudaSuccess) return e;
        }
  // k_window has a little more static shared memory than k_hist
  for (int dt = 1; dt <= 3; ++dt)
    for (int w = 0; w <= 2; ++w)
      for (int k = 0; k <= 4; ++k) {
        cudaError_t e = cudaFuncSetAttribute(reinterpret_cast<const void*>(pick_cols(dt, w, k)), cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin);
        if (e != cudaSuccess) return e;
      }
  for (int w = 1; w <= 2; ++w)
    for (int k = 0; k <= 3; ++k) {
      const void* fns[3] = {reinterpret_cast<const void*>(xhk_pick_mw_f32(w, k)), reinterpret_cast<const void*>(xhk_pick_mw_f64(w, k)),
                            reinterpret_cast<const void*>(xhk_pick_mw_i64(w, k))};
      for (const void* fn : fns) { cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin); if (e != cudaSuccess) return e; }
    }
  for (int dt = 1; dt <= 3; ++dt)
    for (int k = 0; k <= 4; ++k) {
      cudaError_t e = cudaFuncSetAttribute(reinterpret_cast<const void*>(pick_window(dt, k)), cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin - 192);
      if (e != cudaSuccess) return e;
    }
  return cudaSuccess;
}

cudaError_t xhk_launch_hist(const XhkParams& p, const XhkLaunch& l) {
  XhkHistKernel k = pick(l.dtype, l.w_dtype, p.n_vars, kernel_mode(p, l.dtype, l.w_dtype));
  k<<<l.grid, l.threads, l.smem_bytes, l.stream>>>(p);
  return cudaGetLastError();
}

cudaError_t xhk_launch_hist_mw(const XhkParams& p, const XhkMultiWeights& m, const XhkLaunch& l) {
  XhkMwKernel k = l.dtype == 1 ? xhk_pick_mw_f32(l.w_dtype, p.n_vars) : l.dtype == 2 ? xhk_pick_mw_f64(l.w_dtype, p.n_vars) : xhk_pick_mw_i64(l.w_dtype, p.n_vars);
  k<<<l.grid, l.threads, l.smem_bytes, l.stream>>>(p, m);
  return cudaGetLastError();
}

cudaError_t xhk_launch_hist_cols(const XhkParams& p, const XhkLaunch& l, long long inner, int tm, int nsplit, int accumulate) {
  const long long outer = p.M / inner;
  dim3 grid(static_cast<unsigned>(((inner + tm - 1) / tm) * outer), static_cast<unsigned>(nsplit), 1);
  pick_cols(l.dtype, l.w_dtype, p.n_vars)<<<grid, l.threads, l.smem_bytes, l.stream>>>(p, inner, tm, accumulate);
  return cudaGetLastError();
}

size_t xhk_window_kernel_smem(const XhkParams& p) {
  size_t tsz = 8;  // upper bound on sizeof(T)
  size_t e = ((static_cast<size_t>(p.n_edges_total) * tsz + 15) & ~static_cast<size_t>(15)) + ((static_cast<size_t>(p.n_lut_total) * 2 + 15) & ~static_cast<size_t>(15));
  size_t m = 0; for (int k = 0; k < p.n_vars; ++k) m += static_cast<size_t>(p.nb[k]) * 4;
  return e + m;
}

cudaError_t xhk_launch_window(const XhkParams& p, const XhkLaunch& l, XhkWindow* window_dev, int budget_bins, int budget32_bins, int n_probe) {
  const size_t smem = xhk_window_kernel_smem(p);
  pick_window(l.dtype, p.n_vars)<<<1, XHK_THREADS, smem, l.stream>>>(p, window_dev, budget_bins, budget32_bins, n_probe);
  return cudaGetLastError();
}

cudaError_t xhk_launch_zero_shared_rows(const XhkParams& p, const XhkLaunch& l) {
  if (l.w_dtype == 0 || l.w_dtype == 4) k_zero_shared_rows<unsigned long long><<<l.grid, 256, 0, l.stream>>>(p, l.grid);
  else k_zero_shared_rows<double><<<l.grid, 256, 0, l.stream>>>(p, l.grid);
  return cudaGetLastError();
}

cudaError_t xhk_launch_density(void* out, long long M, long long B, int counts, const double* widths_dev, const int* nb, const int* f32,
                               int K, void* sums_dev, cudaStream_t s) {
  XhkDensity q;
  q.widths = widths_dev; q.K = K;
  int o = 0;
  for (int k = 0; k < XHK_MAX_VARS; ++k) { q.off[k] = o; q.nb[k] = k < K ? nb[k] : 1; q.f32[k] = k < K ? f32[k] : 0; if (k < K) o += nb[k]; }
  if (M <= 0 || B <= 0) return cudaSuccess;
  if (B <= 1024) {
    const int thr = 256, rows_per_cta = thr / 32;
    const int grid = static_cast<int>(std::min<long long>((M + rows_per_cta - 1) / rows_per_cta, 148ll * 16));
    if (counts) k_density_rows<true><<<grid, thr, 0, s>>>(out, M, B, q); else k_density_rows<false><<<grid, thr, 0, s>>>(out, M, B, q);
    return cudaGetLastError();
  }
  // sums_dev holds M 8-byte slots
  cudaError_t e = cudaMemsetAsync(sums_dev, 0, static_cast<size_t>(M) * 8, s);
  if (e != cudaSuccess) return e;
  const int chunks = static_cast<int>((B + kDensityChunk - 1) / kDensityChunk);
  const unsigned g1 = static_cast<unsigned>(M * chunks);
  const int g2 = static_cast<int>(std::min<long long>((M * B + 255) / 256, 148ll * 16));
  if (counts) { k_density_sum<true><<<g1, 256, 0, s>>>(out, B, chunks, sums_dev); k_density_scale<true><<<g2, 256, 0, s>>>(out, M, B, sums_dev, q); }
  else { k_density_sum<false><<<g1, 256, 0, s>>>(out, B, chunks, sums_dev); k_density_scale<false><<<g2, 256, 0, s>>>(out, M, B, sums_dev, q); }
  return cudaGetLastError();
}

cudaError_t xhk_launch_peer_allreduce(const XhkPeerArgs& a, int is_f64, cudaStream_t s) {
  const long long n2 = a.count >> 1;
  const int grid = static_cast<int>(std::max<long long>(1, std::min<long long>((n2 + 511) / 512, 64)));
  if (is_f64) k_peer_allreduce<double><<<grid, 512, 0, s>>>(a); else k_peer_allreduce<long long><<<grid, 512, 0, s>>>(a);
  return cudaGetLastError();
}

cudaError_t xhk_launch_permute(const void* in, void* out, int elem_size, int nd, const long long* shape, const int* perm, cudaStream_t s) {
  if (nd < 1 || nd > XHK_MAX_VARS || (elem_size != 4 && elem_size != 8)) return cudaErrorInvalidValue;
  XhkPermute q = {};
  q.nd = nd;
  long long st = 1, total = 1;
  for (int i = nd - 1; i >= 0; --i) { q.shape[i] = shape[i]; q.sin[i] = st; st *= shape[i]; total *= shape[i]; }
  st = 1;
  for (int j = nd - 1; j >= 0; --j) { q.oshape[j] = shape[perm[j]]; q.osin[j] = q.sin[perm[j]]; q.sout[perm[j]] = st; st *= shape[perm[j]]; }
  q.total = total;
  if (total == 0) return cudaSuccess;
  q.da = nd - 1; q.db = perm[nd - 1];
  if (q.da == q.db) {
    const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, 148ll * 32));
    if (elem_size == 4) k_permute_rows<uint32_t><<<grid, 256, 0, s>>>(static_cast<const uint32_t*>(in), static_cast<uint32_t*>(out), q);
    else k_permute_rows<uint64_t><<<grid, 256, 0, s>>>(static_cast<const uint64_t*>(in), static_cast<uint64_t*>(out), q);
    return cudaGetLastError();
  }
  q.tiles_a = (q.shape[q.da] + 31) / 32; q.tiles_b = (q.shape[q.db] + 31) / 32;
  long long blocks = q.tiles_a * q.tiles_b;
  for (int i = 0; i < nd; ++i) if (i != q.da && i != q.db) blocks *= q.shape[i];
  if (blocks > 2147483647ll) return cudaErrorInvalidValue;
  if (elem_size == 4) k_permute_tiled<uint32_t><<<static_cast<unsigned>(blocks), 256, 0, s>>>(static_cast<const uint32_t*>(in), static_cast<uint32_t*>(out), q);
  else k_permute_tiled<uint64_t><<<static_cast<unsigned>(blocks), 256, 0, s>>>(static_cast<const uint64_t*>(in), static_cast<uint64_t*>(out), q);
  return cudaGetLastError();
}

cudaError_t xhk_launch_density_small(const void* src, double* dst, long long M, long long B, int counts, const double* widths_dev, const int* nb,
                                     const int* f32, int K, cudaStream_t s) {
  XhkDensity q;
  q.widths = widths_dev; q.K = K;
  int o = 0;
  for (int k = 0; k < XHK_MAX_VARS; ++k) { q.off[k] = o; q.nb[k] = k < K ? nb[k] : 1; q.f32[k] = k < K ? f32[k] : 0; if (k < K) o += nb[k]; }
  if (M <= 0 || B <= 0) return cudaSuccess;
  const int G = static_cast<int>(std::max<long long>(1, std::min<long long>(std::min<long long>(16, (296 + M - 1) / M), (B + 1023) / 1024)));
  const int thr = B >= 1024 ? 1024 : static_cast<int>(std::max<long long>(32, (B + 31) / 32 * 32));
  const unsigned grid = static_cast<unsigned>(M * G);
  if (counts) k_density_small<true><<<grid, thr, 0, s>>>(src, dst, B, G, q); else k_density_small<false><<<grid, thr, 0, s>>>(src, dst, B, G, q);
  return cudaGetLastError();
}

cudaError_t xhk_launch_fill(void* ptr, int dtype, long long n, unsigned long long seed, long long offset, int normal, cudaStream_t s) {
  const int grid = 148 * 8, thr = 256;
  if (dtype == 1) { if (normal) k_fill<float, 1><<<grid, thr, 0, s>>>(static_cast<float*>(ptr), n, seed, offset); else k_fill<float, 0><<<grid, thr, 0, s>>>(static_cast<float*>(ptr), n, seed, offset); }
  else { if (normal) k_fill<double, 1><<<grid, thr, 0, s>>>(static_cast<double*>(ptr), n, seed, offset); else k_fill<double, 0><<<grid, thr, 0, s>>>(static_cast<double*>(ptr), n, seed, offset); }
  return cudaGetLastError();
}

cudaError_t xhk_launch_minmax(const void* data, int dtype, long long n, double* out_dev, cudaStream_t s) {
  const int grid = 296, thr = 256;  // out_dev holds grid*3 doubles
  if (dtype == 1) k_minmax<float><<<grid, thr, 0, s>>>(static_cast<const float*>(data), n, out_dev);
  else k_minmax<double><<<grid, thr, 0, s>>>(static_cast<const double*>(data), n, out_dev);
  return cudaGetLastError();
}

cudaError_t xhk_launch_flush(void* buf, size_t bytes, cudaStream_t s) {
  k_flush<<<148 * 4, 512, 0, s>>>(static_cast<uint4*>(buf), bytes / 16);
  return cudaGetLastError();
}
