// xhist_kernels.cuh — shared declarations between the kernels (xhist_kernels.cu) and the
// host side of the C-ABI (xhist_api.cu).  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define XHK_MAX_VARS 8
#define XHK_MAX_WEIGHTS 4
#define XHK_MAX_PEERS 16
#define XHK_PEER_FLAG_BYTES 4096   // head of every rank's symmetric buffer: one u64 arrival flag per peer
// threads of a k_hist CTA when one CTA owns an SM (register budget: 65536 / XHK_THREADS per thread)
#ifndef XHK_THREADS
#define XHK_THREADS 1024
#endif
// static shared memory of k_hist (window geometry + flags), rounded up; the host subtracts it from the dynamic budget
#define XHK_STATIC_SMEM 80
enum { XHK_C_LO = 0, XHK_C_HI = 1, XHK_C_E0 = 2, XHK_C_INV = 3, XHK_C_DELTA = 4, XHK_C_OMD = 5, XHK_C_CHALF = 6, XHK_C_SLOTS = 8 };

// How the per-CTA shared-memory histogram is used.
enum XhkHistMode {
  XHK_GLOBAL = 0,  // no shared histogram: every sample goes to a global RED (huge bin spaces / testing)
  XHK_FULL = 1,    // the whole (prod nb) bin space of one row is privatised in shared memory
  XHK_WINDOW = 2   // a hyper-rectangular window of the bin space is privatised; the rest spills to global RED
};

// How rows/samples are split over the persistent CTAs.
enum XhkPartition {
  XHK_PART_SAMPLES = 0,  // equal contiguous ranges of the flattened (M*N) sample space
  XHK_PART_ROWS = 1      // whole rows per CTA (no row is shared between CTAs -> no zero-fill, no atomics)
};

// Device-chosen window (written by the window kernel, read by the histogram kernel).
struct XhkWindow {
  int lo[XHK_MAX_VARS];
  int len[XHK_MAX_VARS];
  // exact fixed-point accumulation of the weights (chosen by the probe kernel):
  // v = w * fx_mul is accumulated as a 64-bit integer in two u32 shared limbs when it is an integer
  // below fx_limit; anything else goes to a float64 global RED.  fx_ok = 0 -> float64 shared adds instead.
  double fx_mul, fx_unmul, fx_limit;
  int fx_ok;
  // 64: two u32 limbs per bin (k_hist<W = 1 or 2>);  32: one u32 limb per bin, wraps go to the float64 output
  // (k_hist<W = 3>, fp32 weights that are non-negative multiples of a power of two within 25 bits);  0: float64 adds
  int fx_mode;
};

// Kernel parameters (passed by value, lives in the constant bank).
// T-typed classification constants are stored as raw 8-byte slots (float kernels read the low 4 bytes).
struct XhkParams {
  const void* data[XHK_MAX_VARS];
  long long stride[XHK_MAX_VARS];   // row stride in elements (0 = broadcast row)
  const void* w;
  long long wstride;
  // classification constants per variable, typed: cf for fp32 kernels, cd for fp64 kernels.
  //   [XHK_C_LO] smallest in-range value (effective first edge)   [XHK_C_HI] largest in-range value
  //   [XHK_C_E0],[XHK_C_INV] uniform path t = (x - e0) * inv
  //   [XHK_C_DELTA],[XHK_C_OMD] uniform path: bin certain iff delta <= frac(t) <= omd
  //   [XHK_C_CHALF] the same test in one compare: certain iff |frac(t) - 0.5| <= chalf  (chalf <= 0.5 - delta)
  float cf[XHK_MAX_VARS][XHK_C_SLOTS];
  double cd[XHK_MAX_VARS][XHK_C_SLOTS];
  long long ci[XHK_MAX_VARS][2];    // int64 kernels: [XHK_C_LO], [XHK_C_HI] only (no uniform path)
  int nb[XHK_MAX_VARS];             // bins of variable k  (= n_edges - 1)
  int uniform[XHK_MAX_VARS];        // 1: uniform fast path usable for variable k
  int all_uniform;                  // 1: every variable is uniform -> branch-free fast classification kernel
  int all_branch_free;              // 1: every variable is uniform or has lut_steps > 0 -> branch-free mixed kernel
  int eoff[XHK_MAX_VARS];           // offset of variable k's effective edges in `edges`
  long long gmul[XHK_MAX_VARS];     // C-order multipliers of the global bin index
  const void* edges;                // device, typed T, all variables concatenated
  int n_edges_total;
  // non-uniform variables: lookup table over lut_n[k] equal cells of [lo, hi]; entry c = bin at the left boundary
  // of cell c.  It brackets the binary search to the edges between cells c-1 and c+2 (usually 1-2 steps).
  const unsigned short* lut;        // device, all variables concatenated
  int n_lut_total;
  int lut_n[XHK_MAX_VARS];          // cells (0 = no table: plain binary search)
  int lut_steps[XHK_MAX_VARS];      // > 0: at most this many edges lie within any 3 consecutive cells -> the bin is the
                                    //      table entry of cell c-1 advanced by that many compare steps, no search loop
  int lut_off[XHK_MAX_VARS];
  float lut_invf[XHK_MAX_VARS];     // cells per unit of x, fp32 / fp64 kernels
  double lut_invd[XHK_MAX_VARS];
  int n_vars;
  void* out;                        // int64 (no weights) or double (weights), [M][B]
  long long B;                      // bins per row
  long long M, N;
  int hist_mode;                    // XhkHistMode
  int partition;                    // XhkPartition
  int hist_capacity;                // shared histogram capacity in bins
  int store_owned_rows;             // 1: a row wholly owned by one CTA is written with plain stores
  const XhkWindow* window;          // XHK_WINDOW: device-chosen window
  int wlo[XHK_MAX_VARS];            // XHK_FULL: 0 / nb ; ignored otherwise
  int wlen[XHK_MAX_VARS];
  long long per_cta;                // XHK_PART_SAMPLES: samples per CTA (multiple of 1024)
  int w_dtype;                      // 0 none, 1 fp32, 2 fp64 (for the probe kernel)
  // row tiling (many short rows): the launch sees tile_rows consecutive rows as one row of N = tile_rows*tile_n
  // samples and B = tile_rows * prod(nb) bins; the local row of a sample is (offset in the tile) / tile_n
  int tile_rows;                    // 1 = no tiling
  int tile_n;                       // samples of one real row
  unsigned tile_magic; int tile_shift;   // q = (umulhi(n, magic) + n) >> shift == n / tile_n for n < 2^31
  int fx_vbits;                     // fixed point: |v| < 2^fx_vbits keeps every per-flush bin sum below 2^63
  int prefetch;                     // 1: L2 prefetch two iterations ahead (default; XH_PREFETCH=0 switches it off for A/B runs — same box,
                                    //    config 4: 1.45 ms without, 1.42 with; config 2: 1.67 / 1.61; config 3: 2.25 / 1.99)
  int spilly;                       // host hint from the cached verdict: > 2 % of the samples left the fast path last time
  unsigned long long* stats;        // device counter: samples that left the fast path (window spills, weights outside the
                                    // fixed-point form) — lets the host notice a cached probe verdict that no longer fits
  int fx32_sibling;                 // 1: a k_hist<W = 3> launch of the same block precedes this one and does the work
                                    //    when the probe chose fx_mode 32 (this launch then returns at once)
};

struct XhkLaunch {
  int dtype;       // 1 f32, 2 f64, 3 int64 (xh_dtype)
  int w_dtype;     // 0 none, 1 f32, 2 f64, 3 f32 accumulated in one u32 limb per bin (fx32 sibling), 4 none + counts packed as 16-bit fields
  int grid, threads;
  size_t smem_bytes;
  cudaStream_t stream;
};

// kernel entry points by data type (defined in xhist_k_f32.cu / xhist_k_f64.cu / xhist_k_i64.cu)
typedef void (*XhkHistKernel)(const XhkParams);
typedef void (*XhkWindowKernel)(const XhkParams, XhkWindow*, int, int, int);
typedef void (*XhkColsKernel)(const XhkParams, long long, int, int);
// several weight arrays in one pass (k_hist_mw)
struct XhkMultiWeights {
  const void* w[XHK_MAX_WEIGHTS];   // device, addressed like XhkParams::w (row stride XhkParams::wstride)
  int nw;
  int use_smem;                     // 1: nw planes of B float64 bins in shared memory; 0: global REDs only
  long long chunk;                  // samples of the reduced axis per work item (multiple of 4)
  long long plane;                  // elements between two weight planes of the output
};
typedef void (*XhkMwKernel)(const XhkParams, const XhkMultiWeights);
#define XHK_DECLARE_PICKERS(DT)                                  \
  XhkHistKernel xhk_pick_hist_##DT(int w, int K, int mode);      \
  XhkWindowKernel xhk_pick_window_##DT(int K);                   \
  XhkColsKernel xhk_pick_cols_##DT(int w, int K);                \
  XhkMwKernel xhk_pick_mw_##DT(int w, int K);
XHK_DECLARE_PICKERS(f32)
XHK_DECLARE_PICKERS(f64)
XHK_DECLARE_PICKERS(i64)

// host-callable launchers (defined in xhist_kernels.cu)
cudaError_t xhk_launch_hist(const XhkParams& p, const XhkLaunch& l);
// column layout: p.M = n_outer * n_inner logical rows, p.N reduced length, inner = n_inner; tm columns per CTA,
// nsplit CTAs share the reduced axis of one column tile (nsplit > 1 -> atomic flush into a zeroed out)
cudaError_t xhk_launch_hist_cols(const XhkParams& p, const XhkLaunch& l, long long inner, int tm, int nsplit, int accumulate);
cudaError_t xhk_launch_hist_mw(const XhkParams& p, const XhkMultiWeights& m, const XhkLaunch& l);
cudaError_t xhk_launch_window(const XhkParams& p, const XhkLaunch& l, XhkWindow* window_dev, int budget_bins, int budget32_bins, int n_probe);
cudaError_t xhk_launch_zero_shared_rows(const XhkParams& p, const XhkLaunch& l);
cudaError_t xhk_set_smem_limits(int max_optin);
// density in place on the device: out[r][b] (int64 counts when `counts`, else float64) -> float64 h / area / rowsum
// (sums_dev: M 8-byte slots of workspace, used when B > 1024)
cudaError_t xhk_launch_density(void* out, long long M, long long B, int counts, const double* widths_dev, const int* nb, const int* f32,
                               int K, void* sums_dev, cudaStream_t s);
// Sum of the partial histograms of all ranks through peer memory (NVLink): slot[r] is rank r's partial (count 8-byte
// items, mapped into this process), flags[r] the head of rank r's symmetric buffer.  Every rank announces `seq` in
// every peer's flag array, waits until all peers have announced it in its own, then adds the partials in rank order
// (identical bits on every rank) into `out` (local).  is_f64: float64 sums, else int64.
struct XhkPeerArgs {
  const void* slot[XHK_MAX_PEERS];
  unsigned long long* flags[XHK_MAX_PEERS];
  int n, rank;
  unsigned long long seq;
  long long count;
  void* out;
};
cudaError_t xhk_launch_peer_allreduce(const XhkPeerArgs& a, int is_f64, cudaStream_t s);
// out = C-contiguous transpose(in, perm) of an nd-dimensional array of 4- or 8-byte elements
cudaError_t xhk_launch_permute(const void* in, void* out, int elem_size, int nd, const long long* shape, const int* perm, cudaStream_t s);
// the same out of place in one launch, for small results; dst may be pinned host memory (device-mapped)
cudaError_t xhk_launch_density_small(const void* src, double* dst, long long M, long long B, int counts, const double* widths_dev, const int* nb,
                                     const int* f32, int K, cudaStream_t s);
cudaError_t xhk_launch_fill(void* ptr, int dtype, long long n, unsigned long long seed, long long offset, int normal, cudaStream_t s);
cudaError_t xhk_launch_minmax(const void* data, int dtype, long long n, double* out2_dev, cudaStream_t s);
cudaError_t xhk_launch_flush(void* buf, size_t bytes, cudaStream_t s);
size_t xhk_window_kernel_smem(const XhkParams& p);
