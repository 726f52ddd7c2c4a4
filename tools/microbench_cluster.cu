// microbench_cluster.cu — mechanism test for config 5 (3 x fp64, 50*60*70 = 210 000 float64 bins = 1.68 MB):
// can a thread-block CLUSTER of 8 CTAs hold the whole bin space in its 8 x 227 KB of shared memory and take the
// adds through distributed shared memory (DSMEM) faster than today's "shared window + 20 % spill to global RED"?
//
// Not part of the library; nothing here is used by the product path.  Written at the end of round 1 (no GPU minutes
// left to run it) as the first thing to measure in round 2 — see DESIGN.md section 8, item 1.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/microbench_cluster tools/microbench_cluster.cu
//   tools/bin/microbench_cluster            (one B200; prints Gsamples/s per variant and a checksum against variant 0)
//
// Stream per sample: one int32 joint bin (pre-classified, -1 = out of range) + one float64 weight = 12 B, so the
// numbers isolate the ACCUMULATION mechanism; classification cost is known from the library's own kernels.
// Variants:
//   0  global RED.F64 on every sample                                   (what a spill costs; ~90-140 Gsamples/s expected)
//   1  per-CTA shared float64 window (box of bins) + global RED spill   (today's k_hist design for config 5)
//   2  cluster of 8: bin space split in 8 slabs, float64 atomicAdd through DSMEM to the owner CTA
//   3  cluster of 8: the same with 64-bit fixed point in two u32 limbs (native integer atomics; carry via the returned value)
//   4  like 2 but every add goes to the CTA's OWN slab (wrong result by design): the local-atomic ceiling of the layout
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s failed: %s (line %d)\n", #x, cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int NB0 = 50, NB1 = 60, NB2 = 70, B = NB0 * NB1 * NB2;     // 210 000 bins
constexpr int CL = 8;                                                // CTAs per cluster
constexpr int SLAB = (B + CL - 1) / CL;                              // 26 250 bins per CTA = 210 000 B of float64
constexpr int THREADS = 1024;

__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
  z += 0x9E3779B97F4A7C15ull; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31);
}
// N(0,1)^3 over [-4, 4]^3 with uniform bins (the real config has non-uniform edges; the occupancy of the bin space is alike)
__global__ void k_gen(int* bin, double* w, long long n) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    int j[3]; bool ok = true;
    for (int k = 0; k < 3; ++k) {
      const unsigned long long h = mix64(mix64(17 + k) ^ static_cast<unsigned long long>(i));
      const double u1 = (static_cast<double>(h >> 32) + 1.0) / 4294967297.0, u2 = static_cast<double>(h & 0xFFFFFFFFull) / 4294967296.0;
      const double x = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
      const int nb = k == 0 ? NB0 : k == 1 ? NB1 : NB2;
      const int b = static_cast<int>(floor((x + 4.0) * (nb / 8.0)));
      ok = ok && b >= 0 && b < nb; j[k] = b;
    }
    bin[i] = ok ? (j[0] * NB1 + j[1]) * NB2 + j[2] : -1;
    w[i] = static_cast<double>(mix64(99 ^ static_cast<unsigned long long>(i)) >> 11) * (1.0 / 9007199254740992.0);
  }
}

struct Args { const int* bin; const double* w; long long n; double* out; int wlo[3], wlen[3]; double fx_mul, fx_unmul; };

__device__ __forceinline__ void load4(const Args& a, long long g, int (&b)[4], double (&w)[4]) {
  const int4 q = __ldcs(reinterpret_cast<const int4*>(a.bin) + g);
  const double2 w0 = __ldcs(reinterpret_cast<const double2*>(a.w) + 2 * g), w1 = __ldcs(reinterpret_cast<const double2*>(a.w) + 2 * g + 1);
  b[0] = q.x; b[1] = q.y; b[2] = q.z; b[3] = q.w; w[0] = w0.x; w[1] = w0.y; w[2] = w1.x; w[3] = w1.y;
}

template <int V>
__global__ void __launch_bounds__(THREADS, 1) k_acc(const __grid_constant__ Args a) {
  extern __shared__ __align__(16) unsigned char smem[];
  double* sh = reinterpret_cast<double*>(smem);                       // V=1: window, V=2,4: slab of doubles, V=3: [SLAB lo][SLAB hi] u32
  unsigned* lo32 = reinterpret_cast<unsigned*>(smem); unsigned* hi32 = lo32 + SLAB;
  const int tid = threadIdx.x;
  int wtot = 0;
  if (V == 1) { wtot = a.wlen[0] * a.wlen[1] * a.wlen[2]; for (int i = tid; i < wtot; i += THREADS) sh[i] = 0.0; }
  if (V >= 2) { for (int i = tid; i < SLAB; i += THREADS) sh[i] = 0.0; }   // (V=3: SLAB doubles = 2*SLAB u32)
  unsigned rank = 0;
  if (V >= 2) { cg::cluster_group cl = cg::this_cluster(); rank = cl.block_rank(); cl.sync(); } else __syncthreads();

  const long long n4 = a.n >> 2;
  const long long per = (n4 + gridDim.x - 1) / gridDim.x;
  const long long g0 = blockIdx.x * per, g1 = (g0 + per < n4) ? g0 + per : n4;
  for (long long g = g0 + tid; g < g1; g += THREADS) {
    int b[4]; double w[4];
    load4(a, g, b, w);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (b[e] < 0) continue;
      if (V == 0) { atomicAdd(a.out + b[e], w[e]); continue; }
      if (V == 1) {
        const int j2 = b[e] % NB2, t = b[e] / NB2, j1 = t % NB1, j0 = t / NB1;
        const unsigned q0 = j0 - a.wlo[0], q1 = j1 - a.wlo[1], q2 = j2 - a.wlo[2];
        if (q0 < static_cast<unsigned>(a.wlen[0]) && q1 < static_cast<unsigned>(a.wlen[1]) && q2 < static_cast<unsigned>(a.wlen[2]))
          atomicAdd(sh + (q0 * a.wlen[1] + q1) * a.wlen[2] + q2, w[e]);
        else atomicAdd(a.out + b[e], w[e]);
        continue;
      }
      const unsigned owner = (V == 4) ? rank : static_cast<unsigned>(b[e]) / SLAB;
      const unsigned idx = static_cast<unsigned>(b[e]) - (static_cast<unsigned>(b[e]) / SLAB) * SLAB;
      cg::cluster_group cl = cg::this_cluster();
      if (V == 2 || V == 4) {
        double* remote = cl.map_shared_rank(sh, owner);
        atomicAdd(remote + idx, w[e]);
      } else {   // V == 3: v = w * 2^s as a 64-bit integer in two u32 limbs of the owner's shared memory
        const long long v = __double2ll_rn(w[e] * a.fx_mul);
        unsigned* rlo = cl.map_shared_rank(lo32, owner); unsigned* rhi = cl.map_shared_rank(hi32, owner);
        const unsigned l = static_cast<unsigned>(v);
        const unsigned old = atomicAdd(rlo + idx, l);
        atomicAdd(rhi + idx, static_cast<unsigned>(static_cast<unsigned long long>(v) >> 32) + ((old + l < old) ? 1u : 0u));
      }
    }
  }
  if (V >= 2) cg::this_cluster().sync(); else __syncthreads();
  // flush
  if (V == 1) {
    for (int i = tid; i < wtot; i += THREADS) {
      const double v = sh[i]; if (v == 0.0) continue;
      const int q2 = i % a.wlen[2], t = i / a.wlen[2], q1 = t % a.wlen[1], q0 = t / a.wlen[1];
      atomicAdd(a.out + ((q0 + a.wlo[0]) * NB1 + q1 + a.wlo[1]) * NB2 + q2 + a.wlo[2], v);
    }
  } else if (V == 2 || V == 4) {
    for (int i = tid; i < SLAB; i += THREADS) { const double v = sh[i]; const long long gb = static_cast<long long>(rank) * SLAB + i; if (v != 0.0 && gb < B) atomicAdd(a.out + gb, v); }
  } else if (V == 3) {
    for (int i = tid; i < SLAB; i += THREADS) {
      const long long iv = static_cast<long long>((static_cast<unsigned long long>(hi32[i]) << 32) | lo32[i]);
      const long long gb = static_cast<long long>(rank) * SLAB + i;
      if (iv != 0 && gb < B) atomicAdd(a.out + gb, static_cast<double>(iv) * a.fx_unmul);
    }
  }
}

template <int V>
float run(const Args& a, int grid, size_t smem, bool cluster, int reps) {
  CK(cudaFuncSetAttribute(k_acc<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  if (cluster) CK(cudaFuncSetAttribute(k_acc<V>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cluster ? CL : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int r = 0; r < reps + 1; ++r) {
    CK(cudaMemset(a.out, 0, sizeof(double) * B));
    CK(cudaEventRecord(e0));
    CK(cudaLaunchKernelEx(&cfg, k_acc<V>, a));
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (r > 0 && ms < best) best = ms;
  }
  return best;
}

int main() {
  const long long n = 1ll << 27;
  int* bin; double *w, *out;
  CK(cudaMalloc(&bin, n * 4)); CK(cudaMalloc(&w, n * 8)); CK(cudaMalloc(&out, sizeof(double) * B));
  k_gen<<<148 * 8, 256>>>(bin, w, n); CK(cudaDeviceSynchronize());
  Args a = {}; a.bin = bin; a.w = w; a.n = n; a.out = out;
  // window of variant 1: the central box of ~28 000 bins (227 KB / 8 B), like the probe's choice for N(0,1)^3
  const int wl[3] = {26, 31, 35};
  for (int k = 0; k < 3; ++k) { const int nb = k == 0 ? NB0 : k == 1 ? NB1 : NB2; a.wlen[k] = wl[k]; a.wlo[k] = (nb - wl[k]) / 2; }
  a.fx_mul = ldexp(1.0, 53); a.fx_unmul = ldexp(1.0, -53);            // weights are k * 2^-53: exact; 2^10 adds of headroom per bin
  std::vector<double> ref(B), got(B);
  const char* names[5] = {"0 global RED f64", "1 shared window + spill", "2 cluster8 DSMEM f64", "3 cluster8 DSMEM 2 x u32 fixed point", "4 cluster8 local slab only (ceiling)"};
  for (int v = 0; v < 5; ++v) {
    float ms = 0;
    if (v == 0) ms = run<0>(a, 148 * 2, 0, false, 3);
    if (v == 1) ms = run<1>(a, 148, static_cast<size_t>(wl[0]) * wl[1] * wl[2] * 8, false, 3);
    if (v == 2) ms = run<2>(a, 144, static_cast<size_t>(SLAB) * 8, true, 3);
    if (v == 3) ms = run<3>(a, 144, static_cast<size_t>(SLAB) * 8, true, 3);
    if (v == 4) ms = run<4>(a, 144, static_cast<size_t>(SLAB) * 8, true, 3);
    CK(cudaMemcpy(got.data(), out, sizeof(double) * B, cudaMemcpyDeviceToHost));
    if (v == 0) ref = got;
    double err = 0, tot = 0; for (int i = 0; i < B; ++i) { err = fmax(err, fabs(got[i] - ref[i])); tot += ref[i]; }
    printf("%-40s %8.3f ms  %7.1f Gsamples/s  %7.1f GB/s   max |diff| vs variant 0 = %.3g (sum %.6g)%s\n", names[v], ms, n / ms / 1e6, n * 12.0 / ms / 1e6,
           err, tot, v == 4 ? "  [wrong by design]" : "");
  }
  return 0;
}
