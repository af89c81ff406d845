// exec.h -- execution layer of the fluca_b200 solver library.
//
// Product build (nvcc, sm_100a): every loop over cells / vector entries is a CUDA kernel launched
// on the solver's stream; reductions are block-reduced with warp shuffles, combined in a fixed
// order by the last block to finish (bitwise reproducible run to run), and read back through a
// pinned host buffer.
//
// Test-only build (-DFLUCA_HOSTEMU, plain g++, see tests/hostemu/): the same functors run in
// serial host loops so that the HOST LOGIC (step driver, Krylov and multigrid orchestration, slab
// partition, halo protocol) can be exercised on a machine without a GPU.  That build is a test
// double: the product loader (fluca_b200/_lib.py) never loads it and the product library refuses
// to run without a CUDA device.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#ifndef FLUCA_HOSTEMU
#include <cuda_runtime.h>
#define FL_HD __host__ __device__ __forceinline__
#define FL_LAMBDA [=] __host__ __device__
// read-only (non-coherent) load of coefficient-table entries; warp-uniform vote over the active lanes
#ifdef __CUDA_ARCH__
#define FL_LDG(p) __ldg(p)
#define FL_WARP_ALL(pred) __all_sync(__activemask(), (pred))
#else
#define FL_LDG(p) (*(p))
#define FL_WARP_ALL(pred) (pred)
#endif
#else
#define FL_HD inline
#define FL_LAMBDA [=]
#define FL_LDG(p) (*(p))
#define FL_WARP_ALL(pred) (pred)
#ifndef __restrict__
#define __restrict__
#endif
#endif

namespace fluca {

// Store of a kernel result (one place to change the store flavour of every stencil functor).  An asm store without a
// memory clobber was tried here so that the compiler could hoist the next plane's loads above it (the functors take their
// arrays through struct members, which carry no __restrict__ information); it changed neither the schedule nor the timing
// (DESIGN.md section 9), so this is a plain store.
FL_HD void fl_store(double *p, double v) { *p = v; }

// L2 prefetch of the element a direct-load stencil functor will need PF planes later.  The plane loop of k_box keeps
// only one plane of loads in flight per thread (profiles/r01l: FaceCombine issue slots 11 % busy, 48 warps stalled on
// the long scoreboard per issue); the prefetch costs no register and no scoreboard entry and turns the later demand
// loads into L2 hits.
#if !defined(FLUCA_HOSTEMU) && defined(__CUDA_ARCH__)
__device__ __forceinline__ void fl_prefetch(const double *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#else
inline void fl_prefetch(const double *) { }
#endif
static const int FL_PF = 2; // planes ahead

struct Error : public std::runtime_error {
  int code;
  Error(int c, const std::string &m) : std::runtime_error(m), code(c) { }
};

enum { FL_OK = 0, FL_ERR_ARG = 1, FL_ERR_CUDA = 2, FL_ERR_NCCL = 3, FL_ERR_DIVERGED = 4, FL_ERR_NODEVICE = 5, FL_ERR_INTERNAL = 6 };

#ifndef FLUCA_HOSTEMU
#define FL_CUDA(call) \
  do { \
    cudaError_t e_ = (call); \
    if (e_ != cudaSuccess) throw ::fluca::Error(::fluca::FL_ERR_CUDA, std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " + __FILE__ + ":" + std::to_string(__LINE__)); \
  } while (0)
typedef cudaStream_t Stream;
#else
typedef void *Stream;
#endif

// launch statistics (gpu_launches of bench.py counts our own kernels)
struct ExecStats {
  long launches = 0;
};

// kernel classes timed live with CUDA events (bench.py roofline): see KTimer below
enum { KT_MOMENTUM_APPLY = 0, KT_MOMENTUM_VEC, KT_POISSON_APPLY, KT_POISSON_VEC, KT_MG_SMOOTH, KT_MG_TRANSFER, KT_RHS_PROJECT, KT_OUTER, KT_HALO, KT_IBM, KT_NCLASS };

struct Exec {
  Stream    stream    = nullptr;
  ExecStats stats;
  // optional per-class event timing
  bool      ktime_on = false;
  bool      kt_grouped = false; // a KGroup is open: the launches inside share its one event pair
  int       kt_current = 7; // class charged by the launchers (KScope sets it); default KT_OUTER
#ifndef FLUCA_HOSTEMU
  std::vector<cudaEvent_t> kt_ev;   // pairs
  std::vector<int>         kt_cls;
  size_t                   kt_used = 0;
#endif
  double kt_ms[KT_NCLASS]    = {0};
  long   kt_count[KT_NCLASS] = {0};
  void   ktime_collect();
  // reduction scratch
  double      *d_partials = nullptr; // [max_blocks * MAXR]
  double      *d_result   = nullptr; // [MAXR]
  double      *d_carry    = nullptr; // [4][MAXR] partial sums handed from one launch of a reduction to the next
  unsigned    *d_ticket   = nullptr;
  double      *h_result   = nullptr; // pinned and mapped [MAXR]
  double      *h_result_dev = nullptr; // the device-side address of h_result (reduce_finish publishes the sums through it)
  long         max_blocks = 0;
  int          sm_count   = 148;
  static const int MAXR = 16;

  void init();
  void destroy();
  void sync();
};

// ---------------------------------------------------------------- memory
inline void *dev_alloc(size_t bytes)
{
  void *p = nullptr;
  if (bytes == 0) bytes = 8;
#ifndef FLUCA_HOSTEMU
  FL_CUDA(cudaMalloc(&p, bytes));
  // cudaMemset runs on the legacy default stream and may return before it completes; the solver's
  // stream is non-blocking, so without this synchronisation a later upload / kernel on that stream
  // could be overtaken by the zero-fill (seen on B200: multigrid tables zeroed after their upload)
  FL_CUDA(cudaMemset(p, 0, bytes));
  FL_CUDA(cudaDeviceSynchronize());
#else
  p = calloc(1, bytes);
  if (!p) throw Error(FL_ERR_INTERNAL, "host allocation failed");
#endif
  return p;
}
inline void dev_free(void *p)
{
  if (!p) return;
#ifndef FLUCA_HOSTEMU
  cudaFree(p);
#else
  free(p);
#endif
}
inline void dev_zero(Exec &ex, void *p, size_t bytes)
{
#ifndef FLUCA_HOSTEMU
  FL_CUDA(cudaMemsetAsync(p, 0, bytes, ex.stream));
#else
  (void)ex;
  memset(p, 0, bytes);
#endif
}
inline void copy_h2d(Exec &ex, void *dst, const void *src, size_t bytes)
{
#ifndef FLUCA_HOSTEMU
  FL_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ex.stream));
#else
  (void)ex;
  if (bytes) memcpy(dst, src, bytes); // an empty marker list hands a null pointer with zero bytes
#endif
}
inline void copy_d2h(Exec &ex, void *dst, const void *src, size_t bytes)
{
#ifndef FLUCA_HOSTEMU
  FL_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ex.stream));
#else
  (void)ex;
  if (bytes) memcpy(dst, src, bytes); // an empty marker list hands a null pointer with zero bytes
#endif
}
inline void copy_d2d(Exec &ex, void *dst, const void *src, size_t bytes)
{
#ifndef FLUCA_HOSTEMU
  FL_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ex.stream));
#else
  (void)ex;
  if (bytes) memmove(dst, src, bytes);
#endif
}
// strided 2-D copies (pitch in bytes), used to pad/unpad rows between the compact C-ABI layout
// and the padded device layout
inline void copy2d_h2d(Exec &ex, void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height)
{
  if (!width || !height) return;
#ifndef FLUCA_HOSTEMU
  FL_CUDA(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, cudaMemcpyHostToDevice, ex.stream));
#else
  (void)ex;
  for (size_t r = 0; r < height; ++r) memcpy((char *)dst + r * dpitch, (const char *)src + r * spitch, width);
#endif
}
inline void copy2d_d2h(Exec &ex, void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height)
{
  if (!width || !height) return;
#ifndef FLUCA_HOSTEMU
  FL_CUDA(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, cudaMemcpyDeviceToHost, ex.stream));
#else
  (void)ex;
  for (size_t r = 0; r < height; ++r) memcpy((char *)dst + r * dpitch, (const char *)src + r * spitch, width);
#endif
}
// the same for a whole field in one call: `depth` planes of `height` rows; a plane is `dslice` / `sslice` rows apart in memory.
// One enqueue per field instead of one per plane (512 per field at 512^3: the host needs 20-30 ms to enqueue the 3584 plane
// copies of a solution view, during which it cannot enqueue the next step).  kind 0: host to device, 1: device to host.
inline void copy3d(Exec &ex, void *stream, int kind, void *dst, size_t dpitch, size_t dslice, const void *src, size_t spitch, size_t sslice, size_t width, size_t height, size_t depth)
{
  if (!width || !height || !depth) return;
#ifndef FLUCA_HOSTEMU
  static const bool by_plane = getenv("FLUCA_B200_COPY_PLANES") != nullptr;
  cudaStream_t      st = stream ? (cudaStream_t)stream : ex.stream;
  if (by_plane) {
    for (size_t k = 0; k < depth; ++k)
      FL_CUDA(cudaMemcpy2DAsync((char *)dst + k * dslice * dpitch, dpitch, (const char *)src + k * sslice * spitch, spitch, width, height, kind ? cudaMemcpyDeviceToHost : cudaMemcpyHostToDevice, st));
    return;
  }
  cudaMemcpy3DParms p;
  memset(&p, 0, sizeof(p));
  p.srcPtr = make_cudaPitchedPtr(const_cast<void *>(src), spitch, width, sslice);
  p.dstPtr = make_cudaPitchedPtr(dst, dpitch, width, dslice);
  p.extent = make_cudaExtent(width, height, depth);
  p.kind   = kind ? cudaMemcpyDeviceToHost : cudaMemcpyHostToDevice;
  FL_CUDA(cudaMemcpy3DAsync(&p, st));
#else
  (void)ex, (void)stream, (void)kind;
  for (size_t k = 0; k < depth; ++k)
    for (size_t r = 0; r < height; ++r) memcpy((char *)dst + (k * dslice + r) * dpitch, (const char *)src + (k * sslice + r) * spitch, width);
#endif
}

inline void Exec::init()
{
#ifndef FLUCA_HOSTEMU
  int dev = 0;
  FL_CUDA(cudaGetDevice(&dev));
  FL_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
  FL_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  max_blocks = 32768; // BASELINE config 5 has 8192 tiles per plane (2048 x 1024 cells)
  d_partials = (double *)dev_alloc(sizeof(double) * max_blocks * MAXR);
  d_result   = (double *)dev_alloc(sizeof(double) * MAXR);
  d_carry    = (double *)dev_alloc(sizeof(double) * MAXR * 4);
  d_ticket   = (unsigned *)dev_alloc(sizeof(unsigned));
  FL_CUDA(cudaHostAlloc((void **)&h_result, sizeof(double) * MAXR, cudaHostAllocMapped));
  if (!getenv("FLUCA_B200_RESULT_MEMCPY")) FL_CUDA(cudaHostGetDevicePointer((void **)&h_result_dev, h_result, 0));
#else
  h_result = (double *)calloc(MAXR, sizeof(double));
  d_result = (double *)calloc(MAXR, sizeof(double));
  d_carry  = (double *)calloc(4 * MAXR, sizeof(double));
#endif
}
inline void Exec::destroy()
{
#ifndef FLUCA_HOSTEMU
  if (stream) cudaStreamSynchronize(stream);
  dev_free(d_partials);
  dev_free(d_result);
  dev_free(d_carry);
  dev_free(d_ticket);
  if (h_result) cudaFreeHost(h_result);
  for (cudaEvent_t e : kt_ev) cudaEventDestroy(e);
  kt_ev.clear();
  if (stream) cudaStreamDestroy(stream);
#else
  free(h_result);
  free(d_result);
  free(d_carry);
#endif
  d_partials = d_result = h_result = d_carry = h_result_dev = nullptr;
  d_ticket   = nullptr;
  stream     = nullptr;
}
inline void Exec::sync()
{
#ifndef FLUCA_HOSTEMU
  FL_CUDA(cudaStreamSynchronize(stream));
#endif
}

// Brackets one kernel launch (or a short launch sequence of one class) with a CUDA event pair on the
// solver's stream when ex.ktime_on is set.  Durations are summed per class by ktime_collect().
struct KTimer {
  Exec &ex;
  int   slot = -1;
  KTimer(Exec &e, int cls, bool active = true) : ex(e)
  {
#ifndef FLUCA_HOSTEMU
    if (!active || !ex.ktime_on || ex.kt_grouped) return;
    if (ex.kt_used + 2 > ex.kt_ev.size()) {
      if (ex.kt_ev.size() >= 200000) return;
      size_t old = ex.kt_ev.size();
      ex.kt_ev.resize(old + 1024);
      for (size_t i = old; i < ex.kt_ev.size(); ++i) cudaEventCreate(&ex.kt_ev[i]);
    }
    slot = (int)ex.kt_used;
    ex.kt_used += 2;
    ex.kt_cls.resize(ex.kt_used / 2);
    ex.kt_cls[slot / 2] = cls;
    cudaEventRecord(ex.kt_ev[slot], ex.stream);
#else
    (void)cls, (void)active;
#endif
  }
  ~KTimer()
  {
#ifndef FLUCA_HOSTEMU
    if (slot >= 0) cudaEventRecord(ex.kt_ev[slot + 1], ex.stream);
#endif
  }
};

// one event pair around a short sequence of launches that together form one operator application
struct KGroup {
  KTimer t;
  Exec  &ex;
  bool   prev;
  KGroup(Exec &e, int cls) : t(e, cls), ex(e), prev(e.kt_grouped) { ex.kt_grouped = true; }
  ~KGroup() { ex.kt_grouped = prev; } // runs before ~KTimer records the closing event
};

// selects the class that the launchers below charge their event pairs to
struct KScope {
  Exec &ex;
  int   prev;
  KScope(Exec &e, int cls) : ex(e), prev(e.kt_current) { ex.kt_current = cls; }
  ~KScope() { ex.kt_current = prev; }
};

inline void Exec::ktime_collect()
{
#ifndef FLUCA_HOSTEMU
  if (!kt_used) return;
  cudaStreamSynchronize(stream);
  for (size_t i = 0; i + 1 < kt_used; i += 2) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, kt_ev[i], kt_ev[i + 1]) == cudaSuccess) {
      kt_ms[kt_cls[i / 2]] += ms;
      kt_count[kt_cls[i / 2]]++;
    }
  }
  kt_used = 0;
#endif
}

// ---------------------------------------------------------------- iteration spaces
struct Box {
  int nx, ny, nz;
};

#ifndef FLUCA_HOSTEMU
// 256 threads: one warp = one 32-cell row segment (256 B, coalesced); 8 rows share their y-neighbours in L1.
// Every block marches a CONTIGUOUS chunk of planes: plane k+1 fetched for cell (i,j,k) is served from L1
// again as the centre plane of k+1 and the lower plane of k+2, so z-neighbours cost no extra DRAM traffic.
static const int BX = 32, BY = 8;

FL_HD void z_chunk(int nz, int nchunks, int c, int &k0, int &k1)
{
  const int base = nz / nchunks, rem = nz % nchunks;
  k0 = c * base + (c < rem ? c : rem);
  k1 = k0 + base + (c < rem ? 1 : 0);
}

// MINB: resident CTAs per SM the register allocation aims at (2 -> 128 registers, 3 -> 80, 4 -> 64).  The streaming functors
// with light stencils are latency-bound at 2 CTAs per SM (one plane of loads in flight per thread): they ask for more.
template <int U, int MINB, class F>
__global__ void __launch_bounds__(BX *BY, MINB) k_box(Box b, F f)
{
  const int i = blockIdx.x * BX + threadIdx.x;
  const int j = blockIdx.y * BY + threadIdx.y;
  if (i >= b.nx || j >= b.ny) return;
  int k0, k1;
  z_chunk(b.nz, gridDim.z, blockIdx.z, k0, k1);
#pragma unroll U
  for (int k = k0; k < k1; ++k) f(i, j, k);
}

template <class F>
__global__ void __launch_bounds__(256) k_range(long n, F f)
{
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) f(i);
}

template <int NR>
__device__ __forceinline__ void block_reduce_and_finish(double (&acc)[NR], const double *carry, double *partials, double *result, unsigned *ticket, unsigned nblocks, unsigned bid)
{
  // carry (optional): NR sums of an earlier launch of the same reduction (boundary planes), added last
  __shared__ double sm[NR][8];
  __shared__ bool   last;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    double v = acc[r];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) sm[r][warp] = v;
  }
  __syncthreads();
  if (tid == 0) {
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      double v = 0.;
      for (int w = 0; w < 8; ++w) v += sm[r][w];
      partials[(size_t)bid * NR + r] = v;
    }
    __threadfence();
    unsigned t = atomicAdd(ticket, 1u);
    last       = (t == nblocks - 1);
  }
  __syncthreads();
  if (last) {
    // fixed-order combination of the block partials: thread t sums partials t, t+256, ...; then one tree
    __threadfence();
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      double v = 0.;
      for (unsigned b = tid; b < nblocks; b += 256) v += ((volatile double *)partials)[(size_t)b * NR + r];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
      if (lane == 0) sm[r][warp] = v;
    }
    __syncthreads();
    if (tid == 0) {
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        double v = 0.;
        for (int w = 0; w < 8; ++w) v += sm[r][w];
        result[r] = carry ? v + carry[r] : v;
      }
      *ticket = 0u;
    }
  }
}

template <int NR, int MINB, class F>
__global__ void __launch_bounds__(BX *BY, MINB) k_box_reduce(Box b, F f, const double *carry, double *partials, double *result, unsigned *ticket)
{
  const int i = blockIdx.x * BX + threadIdx.x;
  const int j = blockIdx.y * BY + threadIdx.y;
  double    acc[NR];
#pragma unroll
  for (int r = 0; r < NR; ++r) acc[r] = 0.;
  if (i < b.nx && j < b.ny) {
    int k0, k1;
    z_chunk(b.nz, gridDim.z, blockIdx.z, k0, k1);
    for (int k = k0; k < k1; ++k) f(i, j, k, acc);
  }
  const unsigned nblocks = gridDim.x * gridDim.y * gridDim.z;
  const unsigned bid     = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
  block_reduce_and_finish<NR>(acc, carry, partials, result, ticket, nblocks, bid);
}

template <int NR, class F>
__global__ void __launch_bounds__(256) k_range_reduce(long n, F f, const double *carry, double *partials, double *result, unsigned *ticket)
{
  double acc[NR];
#pragma unroll
  for (int r = 0; r < NR; ++r) acc[r] = 0.;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) f(i, acc);
  block_reduce_and_finish<NR>(acc, carry, partials, result, ticket, gridDim.x, blockIdx.x);
}

inline dim3 box_grid(const Exec &ex, Box b, long cap_blocks)
{
  unsigned gx = (unsigned)((b.nx + BX - 1) / BX), gy = (unsigned)((b.ny + BY - 1) / BY);
  long     per_plane = (long)gx * gy;
  // enough z-blocks for ~16 resident CTAs' worth of work per SM, never more than the planes
  long want = (long)ex.sm_count * 16;
  long gz   = (want + per_plane - 1) / per_plane;
  if (gz > b.nz) gz = b.nz;
  if (gz < 1) gz = 1;
  if (cap_blocks > 0)
    while (gz > 1 && per_plane * gz > cap_blocks) --gz;
  if (gz > 65535) gz = 65535;
  return dim3(gx, gy, (unsigned)gz);
}
#endif // !FLUCA_HOSTEMU

// f(i, j, k) for every point of the box; U: planes per loop trip of a thread (loads of U planes in flight)
template <int U = 1, int MINB = 2, class F>
inline void for_box(Exec &ex, Box b, F f)
{
  if (b.nx <= 0 || b.ny <= 0 || b.nz <= 0) return;
  ex.stats.launches++;
#ifndef FLUCA_HOSTEMU
  dim3   g = box_grid(ex, b, 0);
  KTimer kt(ex, ex.kt_current);
  k_box<U, MINB><<<g, dim3(BX, BY, 1), 0, ex.stream>>>(b, f);
  FL_CUDA(cudaGetLastError());
#else
  for (int k = 0; k < b.nz; ++k)
    for (int j = 0; j < b.ny; ++j)
      for (int i = 0; i < b.nx; ++i) f(i, j, k);
#endif
}

// f(idx) for idx in [0, n)
template <class F>
inline void for_range(Exec &ex, long n, F f)
{
  if (n <= 0) return;
  ex.stats.launches++;
#ifndef FLUCA_HOSTEMU
  long blocks = (n + 255) / 256, cap = (long)ex.sm_count * 16;
  if (blocks > cap) blocks = cap;
  KTimer kt(ex, ex.kt_current);
  k_range<<<(unsigned)blocks, 256, 0, ex.stream>>>(n, f);
  FL_CUDA(cudaGetLastError());
#else
  for (long i = 0; i < n; ++i) f(i);
#endif
}

// f(i, j, k, acc[NR]) accumulates into acc; the NR sums land in ex.d_result (device)
// carry / result (device pointers, CUDA build): chain the sums of several launches of one reduction;
// the default writes ex.d_result
template <int NR, int MINB = 2, class F>
inline void for_box_reduce(Exec &ex, Box b, F f, const double *carry = nullptr, double *result = nullptr)
{
  static_assert(NR <= Exec::MAXR, "too many simultaneous reductions");
  ex.stats.launches++;
  if (!result) result = ex.d_result;
#ifndef FLUCA_HOSTEMU
  if (b.nx <= 0 || b.ny <= 0 || b.nz <= 0) {
    if (carry) copy_d2d(ex, result, carry, sizeof(double) * NR);
    else dev_zero(ex, result, sizeof(double) * NR);
    return;
  }
  dim3 g = box_grid(ex, b, ex.max_blocks);
  if ((long)g.x * g.y * g.z > ex.max_blocks) throw Error(FL_ERR_INTERNAL, "reduction grid exceeds partial buffer");
  KTimer kt(ex, ex.kt_current);
  k_box_reduce<NR, MINB><<<g, dim3(BX, BY, 1), 0, ex.stream>>>(b, f, carry, ex.d_partials, result, ex.d_ticket);
  FL_CUDA(cudaGetLastError());
#else
  double acc[NR];
  for (int r = 0; r < NR; ++r) acc[r] = 0.;
  for (int k = 0; k < b.nz; ++k)
    for (int j = 0; j < b.ny; ++j)
      for (int i = 0; i < b.nx; ++i) f(i, j, k, acc);
  for (int r = 0; r < NR; ++r) result[r] = acc[r] + (carry ? carry[r] : 0.);
#endif
}

template <int NR, class F>
inline void for_range_reduce(Exec &ex, long n, F f)
{
  static_assert(NR <= Exec::MAXR, "too many simultaneous reductions");
  ex.stats.launches++;
#ifndef FLUCA_HOSTEMU
  if (n <= 0) {
    dev_zero(ex, ex.d_result, sizeof(double) * NR);
    return;
  }
  long blocks = (n + 255) / 256, cap = (long)ex.sm_count * 8;
  if (blocks > cap) blocks = cap;
  if (blocks > ex.max_blocks) blocks = ex.max_blocks;
  KTimer kt(ex, ex.kt_current);
  k_range_reduce<NR><<<(unsigned)blocks, 256, 0, ex.stream>>>(n, f, nullptr, ex.d_partials, ex.d_result, ex.d_ticket);
  FL_CUDA(cudaGetLastError());
#else
  double acc[NR];
  for (int r = 0; r < NR; ++r) acc[r] = 0.;
  for (long i = 0; i < n; ++i) f(i, acc);
  for (int r = 0; r < NR; ++r) ex.d_result[r] = acc[r];
#endif
}

} // namespace fluca
