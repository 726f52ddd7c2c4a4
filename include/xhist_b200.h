/*
 * xhist_b200.h — C-ABI of the B200-native histogram hot path.
 *
 * This is the drop-in boundary for ONE path of xgcm/xhistogram: the block-wise
 * digitize -> ravel_multi_index -> bincount loop of xhistogram/core.py
 * (reference: _bincount_2d_vectorized core.py:137-194, _bincount_2d core.py:73-83,
 * _bincount_loop core.py:86-99, _bincount core.py:197-247).  The reference has no
 * FFI of its own (it is pure Python over numpy), so every entry point below cites
 * the reference Python interface it replaces; INTEGRATION.md shows the ctypes stub a
 * maintainer would add inside xhistogram.core._bincount.
 *
 * Conventions
 *   - plain C types only; no ownership transfer of caller memory;
 *   - every function returns 0 on success or a negative xh_status; the message of
 *     the last failure on the calling thread is available through xh_last_error();
 *   - never aborts, never throws; safe to call concurrently from several threads
 *     (per-device state is mutex-guarded) — dask's threaded scheduler calls the
 *     reference's _bincount concurrently (core.py:429-437);
 *   - synchronous: on return the result is valid in `out`.
 */
#ifndef XHIST_B200_H
#define XHIST_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XH_VERSION_MAJOR 0
#define XH_VERSION_MINOR 1
#define XH_MAX_VARS 8
#define XH_MAX_WEIGHTS 4

typedef enum xh_status {
  XH_OK = 0,
  XH_ERR_INVALID = -1,     /* bad argument / unsupported combination            */
  XH_ERR_CUDA = -2,        /* CUDA runtime error (message has the detail)       */
  XH_ERR_NO_DEVICE = -3,   /* no usable sm_100 device                           */
  XH_ERR_NOMEM = -4,       /* device or pinned-host allocation failed           */
  XH_ERR_NCCL = -5,        /* NCCL missing or failed                            */
  XH_ERR_UNSUPPORTED = -6  /* valid request this build does not implement       */
} xh_status;

typedef enum xh_dtype { XH_NONE = 0, XH_F32 = 1, XH_F64 = 2, XH_I64 = 3 /* data only: int64, datetime64/timedelta64 ticks */ } xh_dtype;
typedef enum xh_mem { XH_HOST = 0, XH_DEVICE = 1 } xh_mem;

/* flags for xh_desc.flags */
#define XH_FLAG_NO_ZERO 1u      /* accumulate into `out` instead of zero-filling it first (device out only) */
#define XH_FLAG_FORCE_GLOBAL 2u /* testing: bypass the shared-memory histogram, global atomics only         */
#define XH_FLAG_FORCE_SEARCH 4u /* testing: bypass the uniform-edge fast path, binary search only           */
#define XH_FLAG_FORCE_WINDOW 8u /* testing: use the windowed shared-memory histogram even if all bins fit   */
#define XH_FLAG_FORCE_PACKED 512u /* testing: counts take the packed 16-bit shared histogram whenever it applies            */
#define XH_FLAG_ONE_PASS 1024u  /* several weight arrays on DEVICE data: take the one-pass kernel (k_hist_mw) although one fused pass
                                   per weight array is faster there (what the library does by default; host data always take the
                                   one-pass kernel — the samples cross PCIe once)                                              */
#define XH_FLAG_NO_FX32 16u     /* testing: fp32 weights never take the one-limb (4 bytes per bin) accumulation  */
#define XH_FLAG_ALLREDUCE 64u   /* sum the (n_rows, bins) result over the ranks of this device's communicator
                                   (xh_comm_init_rank) with ncclAllReduce before the density / the copy to `out`:
                                   histogram kernels, collective, density and D2H are one stream-ordered call — the
                                   role of dask's blockwise + .sum in core.py:429-439                              */
#define XH_FLAG_ASYNC 128u      /* device data and device out only: return once the work is enqueued on the stream; `out`
                                   is valid in stream order (xh_sync, or any later call on the same stream, orders after it) */
#define XH_FLAG_OUT_PINNED 256u  /* host `out` is page-locked memory from xh_host_alloc (device-mapped): the library lets the GPU
                                   write the result into it directly instead of landing it in its own pinned buffer and copying */
#define XH_FLAG_DENSITY 32u     /* finish the density on the device (core.py:444-462): out becomes float64
                                   counts / bin areas / row sum, also without weights; needs widths[]         */

/*
 * One histogram request over a logical (n_rows, n_cols) block: histogram along the
 * columns, independently for every row.  Replaces the reference call
 *     _bincount_2d_vectorized(*args, bins=, weights=, block_size=)   core.py:137-194
 * i.e. per variable k: searchsorted(edges_k, x_k, "right") with the last bin
 * right-inclusive (core.py:163-174), joint index (core.py:178-181), per-row
 * bincount (core.py:73-83) and removal of the under/overflow cells (core.py:191-192).
 *
 *   data[k]      element (r, c) of variable k is data[k][r*row_stride[k] + c];
 *                row_stride 0 broadcasts one row over all rows (core.py:366).
 *   weights      optional, same addressing with w_row_stride; any weights make the
 *                result float64 accumulated in float64 (np.bincount, core.py:81).
 *   edges[k]     HOST float64 (int64 in iedges[k] when dtype == XH_I64: integers and datetime64 ticks compare
 *                exactly, as numpy compares them), n_edges[k] >= 2 non-decreasing values; comparison
 *                semantics are numpy's (data promoted with the edges; fp32 data
 *                against fp32-representable edges compares in fp32 — identical
 *                results).
 *   n_inner      column layout for reductions over LEADING axes (np.moveaxis + reshape in core.py:218-226 would
 *                copy): the arrays are C-contiguous (n_outer, n_cols, n_inner) blocks, n_rows = n_outer*n_inner,
 *                logical row a*n_inner + m holds the samples data[k][(a*n_cols + c)*n_inner + m], c < n_cols;
 *                row strides are ignored.
 *   out          (n_rows, prod(n_edges[k]-1)) C-order; int64 without weights,
 *                float64 with weights; caller-owned, library zero-fills unless
 *                XH_FLAG_NO_ZERO.
 */
typedef struct xh_desc {
  int32_t n_vars;                     /* K, 1..XH_MAX_VARS                                   */
  int32_t dtype;                      /* xh_dtype of every data[k]                           */
  int32_t w_dtype;                    /* xh_dtype of weights, XH_NONE when unweighted        */
  int32_t mem;                        /* xh_mem of data[] and weights                        */
  int32_t out_mem;                    /* xh_mem of out                                       */
  int32_t device;                     /* CUDA device ordinal                                 */
  uint32_t flags;
  int32_t reserved;
  int64_t n_rows, n_cols;             /* M kept rows, N reduced columns                      */
  const void* data[XH_MAX_VARS];
  int64_t row_stride[XH_MAX_VARS];    /* in elements                                         */
  const void* weights;
  int64_t w_row_stride;
  const double* edges[XH_MAX_VARS];
  int32_t n_edges[XH_MAX_VARS];
  void* out;
  void* stream;                       /* cudaStream_t for device inputs; NULL = library stream */
  float* kernel_ms;                   /* optional: device time of the kernels of this call   */
  const int64_t* iedges[XH_MAX_VARS]; /* dtype == XH_I64: the edges as int64 (edges[] unused)  */
  int64_t n_inner;                    /* > 1: column layout (reduced axes lead, see below); 0/1: row layout */
  const double* widths[XH_MAX_VARS];  /* XH_FLAG_DENSITY: HOST np.diff(edges_k) as float64, n_edges[k]-1 values  */
  int32_t widths_f32[XH_MAX_VARS];    /* 1: numpy holds these widths as float32 (a product of two such is
                                         rounded to float32, as np.multiply.outer does in core.py:447-454)      */
  int32_t n_weights;                  /* 0 / 1: `weights` alone.  2..XH_MAX_WEIGHTS: several weight arrays over the same samples
                                         in ONE call (the reference needs one call per weight array: tutorial.ipynb:298-360,
                                         "TODO: allow list of weights" xarray.py:106): weights, weights_more[0..n_weights-2], all
                                         of w_dtype and addressed with w_row_stride; out is (n_weights, n_rows, bins) float64   */
  int32_t reserved2;
  const void* weights_more[XH_MAX_WEIGHTS - 1];
} xh_desc;

/* library / device lifecycle --------------------------------------------------------- */
int xh_version(void);                                   /* major*1000 + minor */
int xh_desc_size(void);                                 /* sizeof(xh_desc): lets a binding check its mirror of the struct */
int xh_device_count(int* count);
int xh_init(int device);                                /* create the per-device context (idempotent) */
int xh_shutdown(void);                                  /* release every context, workspace and communicator */
int xh_last_error(char* buf, size_t len);               /* copy the calling thread's last error message */
int xh_device_info(int device, int* sm_count, int* smem_optin_bytes, int64_t* total_mem, int* cc_major, int* cc_minor);

/* the hot path ------------------------------------------------------------------------ */
int xh_hist(const xh_desc* d);                          /* replaces core.py:137-194 for one block */

/* Host-side phase clock of the calling thread's last xh_hist, microseconds since entry: [0] edge tables ready (cache
 * hit or prepared + uploaded), [1] all work enqueued, [2] stream synchronised, [3] return.  For overhead accounting.  */
int xh_last_call_phases(double* us4);

/* Block-partitioned form: the same request sharded over `n_dev` devices of this process
 * (rows when n_rows >= n_dev, else columns) and, when columns are sharded, combined with
 * ncclAllReduce(sum) — the role dask's blockwise + .sum plays in core.py:429-439.
 * `devices` lists CUDA ordinals; data must be HOST memory; out is HOST memory.           */
int xh_hist_multi(const xh_desc* d, const int32_t* devices, int32_t n_dev);

/* min/max of a device or host array (np.histogram_bin_edges' range pass, core.py:383-388);
 * NaNs make both results NaN like numpy's a.min()/a.max().                                */
int xh_minmax(int device, const void* data, int dtype, int mem, int64_t n, double* mn, double* mx);

/* device buffers for the device-resident path ---------------------------------------- */
int xh_malloc(int device, size_t bytes, void** ptr);
int xh_free(int device, void* ptr);
int xh_host_alloc(size_t bytes, void** ptr);            /* pinned, device-mapped host memory: fast H2D source, direct result target */
int xh_host_free(void* ptr);
int xh_memcpy(int device, void* dst, const void* src, size_t bytes, int dst_mem, int src_mem);
int xh_memset(int device, void* dst, int value, size_t bytes);
int xh_sync(int device);
/* Order the library stream after the work already enqueued on `producer_stream` (a cudaStream_t; 1 / 2 = legacy /
 * per-thread default stream): the synchronisation a __cuda_array_interface__ consumer owes its exporter.            */
int xh_stream_wait(int device, void* producer_stream);
/* dst = C-contiguous copy of transpose(src, perm) for an ndim-dimensional device array of 4- or 8-byte elements,
 * enqueued on the library stream.  Replaces the host copy np.moveaxis(...).reshape(...) of core.py:218-226 for
 * device-resident inputs whose reduce axes neither trail nor form one leading block.                              */
int xh_permute(int device, const void* src, void* dst, int elem_size, int ndim, const int64_t* shape, const int32_t* perm);

/* synthetic data (counter-based: element i depends only on (seed, offset+i)) ---------- */
int xh_fill_normal(int device, void* ptr, int dtype, int64_t n, uint64_t seed, int64_t offset);
int xh_fill_uniform(int device, void* ptr, int dtype, int64_t n, uint64_t seed, int64_t offset);

/* device timing on the library stream (bench harness) --------------------------------- */
int xh_timer_start(int device);
int xh_timer_stop(int device, float* ms);               /* synchronises, returns elapsed ms */
int xh_flush_l2(int device);                            /* overwrite a >L2-sized scratch buffer */

/* multi-process partial-histogram reduction over NCCL (one rank per GPU) -------------- */
#define XH_NCCL_UNIQUE_ID_BYTES 128
int xh_comm_unique_id(void* id128);                     /* rank 0: create id, broadcast it out of band */
int xh_comm_init_rank(int device, const void* id128, int n_ranks, int rank);
int xh_comm_allreduce(int device, void* dev_buf, int64_t count, int dtype_is_f64); /* sum, int64 or float64, in place */
int xh_comm_destroy(int device);

#ifdef __cplusplus
}
#endif
#endif /* XHIST_B200_H */
