// tma.h -- TMA-staged plane pipeline for the 3-D stencil kernels (sm_100a).
//
// Every bandwidth-bound stencil operator of the step (momentum operator A, Poisson operator,
// multigrid smoother / residual) has the same shape: a CTA owns a TMX x TMY column of cells and
// marches it through a contiguous chunk of z-planes; for plane k it needs planes k-1, k, k+1 of
// NIN input fields with a one-cell halo in x and y.  Here the planes are moved HBM -> shared memory
// by the Tensor Memory Accelerator (cp.async.bulk.tensor.3d, one box of (TMX+2) x (TMY+2) x 1
// doubles per field and plane) into a ring of TMS slots, each guarded by an mbarrier that the
// loads complete with complete_tx::bytes.  One elected thread keeps the ring full (loads run two to three planes
// ahead of the plane being computed), so the bytes in flight per SM (2 CTAs x NIN x 2.9 KB x planes ahead)
// do not depend on how many registers the arithmetic needs -- the limit of the direct-load version
// (profiles/r01a_*: 128 registers, 16 warps/SM, long-scoreboard bound at 30 % of HBM peak).
//
// Out-of-range box elements (i < 0, j = -1, j = py) are zero-filled by the TMA unit; they only
// ever meet zero stencil weights, exactly like the clamped neighbour indices of the direct-load
// functors in stencil.h.  Tiles are always full: the last tile of a row / column is shifted back
// to end at nx / ny and the overlapped cells are computed by the tile that owns them only.
// Directions that are periodic in x or y, grids narrower than one tile, and 2-D meshes use the
// direct-load kernels of exec.h instead (tma_usable()).
#pragma once
#ifndef FLUCA_HOSTEMU
#include "exec.h"
#include <cuda.h>
#include <cstdint>

namespace fluca {

static const int TMX = 32, TMY = 8;             // cells per tile (one warp = one 32-cell row)
// box with halo.  The innermost start coordinate of a TMA box must be 16-byte aligned (measured on B200 with
// tools/tma_probe.cu: an odd fp64 start coordinate raises "illegal instruction", even ones -- negative or
// not -- are fine), so the x halo is two cells wide and tile origins are even.
static const int THX = 2;
static const int TLX = TMX + 2 * THX, TLY = TMY + 2;
static const int TILE_ELEMS = TLX * TLY;        // 360 doubles = 2880 B per field and plane
static const int TILE_STRIDE = 368;             // doubles between consecutive field tiles (2944 B, 128-B aligned)
static const int TMS = 4;                       // ring slots: planes k-1, k, k+1 live + one in flight (two while a plane is computed)

template <int NIN>
struct alignas(64) TmaIn {
  CUtensorMap m[NIN];
};

struct TmaGrid {
  int nx, ny;     // cells
  int kbeg, kend; // local planes [kbeg, kend) computed by this launch
  int ntx, nty, nchunk;
};

// tensor map of one field array laid out (px, py, nplanes) doubles; cached per (pointer, extents)
const CUtensorMap &tensor_map_for(const double *field, int px, int py, int nplanes);
void               tensor_map_forget(const double *field);

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
  // makes the initialised barriers visible to the async proxy (the TMA unit completes transactions on them)
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, unsigned bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, unsigned parity)
{
  unsigned ok;
  asm volatile("{\n\t.reg .pred P_OUT;\n\t"
               "mbarrier.try_wait.parity.shared::cta.b64 P_OUT, [%1], %2;\n\t"
               "selp.b32 %0, 1, 0, P_OUT;\n\t}"
               : "=r"(ok)
               : "r"(smem_u32(bar)), "r"(parity)
               : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
  // try_wait suspends the thread in hardware for a bounded time; a barrier that never completes (a protocol
  // bug) traps after ~seconds instead of hanging the GPU
  unsigned spins = 0;
  while (!mbar_try_wait(bar, parity))
    if (++spins > (1u << 22)) __trap();
}
// one (TLX x TLY x 1) box of doubles, global -> shared, completing on `bar`
__device__ __forceinline__ void tma_load_box(void *dst, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar)
{
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
               : "memory");
}

// view of the three live planes handed to an operator
struct TileView {
  const double *pm, *p0, *pp; // slot bases of planes k-1, k, k+1; field f starts at + f * TILE_STRIDE
  int           lc;           // element index of this thread's cell inside a field tile
};

// Op interface:
//   static const int NIN, NR, MINB;                                     inputs, reductions, resident CTAs per SM aimed at
//   struct Regs;                                                         per-thread operands that bypass the tiles
//   __device__ void prefetch(Regs &r, int i, int j, int kl) const;       global loads for plane kl, issued one plane ahead
//   __device__ void cell(const TileView &tv, const Regs &r, int i, int j, int kl, double *acc) const;   (acc: NR entries)
template <class Op>
__global__ void __launch_bounds__(TMX *TMY, Op::MINB) k_tma_march(const __grid_constant__ TmaIn<Op::NIN> in, const Op op, const TmaGrid tg, const double *carry, double *partials, double *result, unsigned *ticket)
{
  extern __shared__ __align__(128) unsigned char tma_smem[];
  double   *ring = reinterpret_cast<double *>(tma_smem);
  uint64_t *full = reinterpret_cast<uint64_t *>(tma_smem + (size_t)TMS * Op::NIN * TILE_STRIDE * sizeof(double));
  const int tid = threadIdx.x, tx = tid & (TMX - 1), ty = tid / TMX;
  const int ntile = tg.ntx * tg.nty;
  const int tile = blockIdx.x % ntile, bz = blockIdx.x / ntile;
  const int bx = tile % tg.ntx, by = tile / tg.ntx;
  const int i0 = min(bx * TMX, (tg.nx - TMX + 1) & ~1), j0 = min(by * TMY, tg.ny - TMY); // i0 even
  int       k0, k1;
  z_chunk(tg.kend - tg.kbeg, tg.nchunk, bz, k0, k1);
  k0 += tg.kbeg, k1 += tg.kbeg;
  const int  i = i0 + tx, j = j0 + ty;
  const bool owned = (i >= bx * TMX) && (i < tg.nx) && (j >= by * TMY);
  const int  nplanes = k1 - k0 + 2; // ring index r <-> local plane k0 - 1 + r
  constexpr unsigned TX_BYTES = Op::NIN * TILE_ELEMS * sizeof(double);

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < TMS; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0) {
    const int ahead = nplanes < TMS ? nplanes : TMS; // fill every slot: planes k0-1 .. k0+2
    for (int r = 0; r < ahead; ++r) {
      const int s = r % TMS;
      mbar_arrive_expect_tx(&full[s], TX_BYTES);
#pragma unroll
      for (int f = 0; f < Op::NIN; ++f) tma_load_box(ring + ((size_t)s * Op::NIN + f) * TILE_STRIDE, &in.m[f], i0 - THX, j0 - 1, k0 + r, &full[s]);
    }
  }
  double acc[Op::NR > 0 ? Op::NR : 1];
#pragma unroll
  for (int r = 0; r < (Op::NR > 0 ? Op::NR : 1); ++r) acc[r] = 0.;
  TileView tv;
  tv.lc     = (ty + 1) * TLX + tx + THX;
  int ready = -1;
  typename Op::Regs cur, nxt;
  if (owned && k0 < k1) op.prefetch(nxt, i, j, k0);
  for (int k = k0; k < k1; ++k) {
    const int rc = k - k0 + 1;
    cur = nxt;
    // operands read straight from global memory (one value per cell, no reuse) are requested a whole plane
    // ahead: their DRAM latency overlaps the arithmetic of the current plane
    if (owned && k + 1 < k1) op.prefetch(nxt, i, j, k + 1);
    while (ready < rc + 1) {
      ++ready;
      mbar_wait(&full[ready % TMS], (unsigned)(ready / TMS) & 1u);
    }
    tv.pm = ring + (size_t)((rc - 1) % TMS) * Op::NIN * TILE_STRIDE;
    tv.p0 = ring + (size_t)(rc % TMS) * Op::NIN * TILE_STRIDE;
    tv.pp = ring + (size_t)((rc + 1) % TMS) * Op::NIN * TILE_STRIDE;
    if (owned) op.cell(tv, cur, i, j, k, acc);
    __syncthreads(); // every thread is done with plane k-1: its slot may be refilled
    if (tid == 0) {
      const int r = rc - 1 + TMS;
      if (r < nplanes) {
        const int s = r % TMS;
        mbar_arrive_expect_tx(&full[s], TX_BYTES);
#pragma unroll
        for (int f = 0; f < Op::NIN; ++f) tma_load_box(ring + ((size_t)s * Op::NIN + f) * TILE_STRIDE, &in.m[f], i0 - THX, j0 - 1, k0 + r, &full[s]);
      }
    }
  }
  if (Op::NR > 0) block_reduce_and_finish<(Op::NR > 0 ? Op::NR : 1)>(acc, carry, partials, result, ticket, gridDim.x, blockIdx.x);
}

// number of z-chunks: fills the resident CTA slots in whole waves while keeping chunks long (every chunk
// re-reads its two boundary planes)
inline int tma_pick_chunks(int ntile, int nplanes, int slots, long max_blocks)
{
  int  best = 1;
  long best_cost = -1;
  for (int c = 1; c <= nplanes; ++c) {
    if ((long)ntile * c > max_blocks) break;
    const long waves = ((long)ntile * c + slots - 1) / slots;
    const long len   = (nplanes + c - 1) / c;
    if (len < 4 && c > 1) break;
    const long cost = waves * (len + 2);
    if (best_cost < 0 || cost < best_cost) best_cost = cost, best = c;
  }
  return best;
}

// launches Op over local planes [kbeg, kend); `fields` are the NIN input arrays in the order the operator expects
template <class Op>
inline void tma_launch(Exec &ex, const Op &op, const double *const *fields, int px, int py, int nplanes_alloc, int nx, int ny, int kbeg, int kend, const double *carry)
{
  if (kend <= kbeg) {
    if (Op::NR > 0) {
      if (carry) copy_d2d(ex, ex.d_result, carry, sizeof(double) * Op::NR);
      else dev_zero(ex, ex.d_result, sizeof(double) * Op::NR);
    }
    return;
  }
  TmaIn<Op::NIN> in;
  for (int f = 0; f < Op::NIN; ++f) in.m[f] = tensor_map_for(fields[f], px, py, nplanes_alloc);
  TmaGrid tg;
  tg.nx = nx, tg.ny = ny, tg.kbeg = kbeg, tg.kend = kend;
  tg.ntx = (nx + TMX - 1) / TMX, tg.nty = (ny + TMY - 1) / TMY;
  tg.nchunk = tma_pick_chunks(tg.ntx * tg.nty, kend - kbeg, Op::MINB * ex.sm_count, ex.max_blocks);
  const size_t smem = (size_t)TMS * Op::NIN * TILE_STRIDE * sizeof(double) + TMS * sizeof(uint64_t);
  static bool  configured = false; // per template instantiation
  if (!configured) {
    FL_CUDA(cudaFuncSetAttribute(k_tma_march<Op>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  ex.stats.launches++;
  KTimer kt(ex, ex.kt_current);
  k_tma_march<Op><<<(unsigned)(tg.ntx * tg.nty * tg.nchunk), TMX * TMY, smem, ex.stream>>>(in, op, tg, carry, ex.d_partials, ex.d_result, ex.d_ticket);
  FL_CUDA(cudaGetLastError());
}

} // namespace fluca
#endif // !FLUCA_HOSTEMU
