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
#ifndef FL_TILE_PLANES
#define FL_TILE_PLANES 1 // planes per CTA barrier of the 1-field tile operators; 2 measured 4-6 % SLOWER (profiles/r04: the kernels are issue-bound, not barrier-bound)
#endif
// box with halo.  The innermost start coordinate of a TMA box must be 16-byte aligned (measured on B200 with
// tools/tma_probe.cu: an odd fp64 start coordinate raises "illegal instruction", even ones -- negative or
// not -- are fine), so the x halo is two cells wide and tile origins are even.
static const int THX = 2;
static const int TLX = TMX + 2 * THX, TLY = TMY + 2;
static const int TILE_ELEMS = TLX * TLY;        // 360 doubles = 2880 B per field and plane
static const int TILE_STRIDE = 368;             // doubles between consecutive field tiles (2944 B, 128-B aligned)
// ring slots per operator (Op::STAGES): planes k-1, k, k+1 live + (STAGES - 3) in flight.  Operators with few input
// fields need a deep ring: with 4 slots a 1-field operator keeps 4 CTAs x 1 plane x 2.9 KB = 12 KB in flight per SM, far
// below the ~40 KB that saturate HBM (measured: the smoother ran at 45 % of peak); 8 slots give 58 KB.

template <int NIN>
struct alignas(64) TmaIn {
  CUtensorMap   m[NIN];
  const double *p[NIN]; // the same arrays as plain pointers (periodic-x wrap columns are patched in with ordinary loads)
};

struct TmaGrid {
  int px, py;     // padded row length / rows per plane of every field of the launch
  int nx, ny;     // cells
  int kbeg, kend; // local planes [kbeg, kend) computed by this launch
  int ntx, nty, nchunk;
  int perx;       // x is periodic: the halo columns of the first / last tile of a row hold the wrapped cells
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
  double       *scratch;      // Op::SCRATCH bytes of shared memory for exchanges between the threads of the CTA
};

// defaults of the optional parts of the operator interface
struct TileOpDefaults {
  static const int  ZALIGN  = 1;     // chunk boundaries are multiples of ZALIGN planes
  static const int  SCRATCH = 0;     // bytes of CTA scratch in shared memory
  static const bool POST    = false; // post(...) is called after the plane barrier (sees what every thread wrote to scratch)
  static const int  PLANES  = 1;     // planes computed per CTA barrier (needs STAGES >= PLANES + 3)
};

__device__ __forceinline__ void mbar_wait_addr(uint32_t bar, unsigned parity)
{
  unsigned ok, spins = 0;
  do {
    asm volatile("{\n\t.reg .pred P_OUT;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 P_OUT, [%1], %2;\n\t"
                 "selp.b32 %0, 1, 0, P_OUT;\n\t}"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
    if (!ok && ++spins > (1u << 22)) __trap(); // a protocol bug traps after ~seconds instead of hanging the GPU
  } while (!ok);
}

// Periodic x: the TMA unit zero-fills the halo columns left of column 0 and reads padding right of column nx - 1.  The CTAs
// that own the first / last tile of a row overwrite those two halo columns of a freshly arrived plane with the cells from the
// other end of the row (plain loads; 2 x TMY x NIN values), then the CTA synchronises.  Only 2 of the nx / 32 CTAs of a row
// pay for it.  (A second TMA box cannot land inside the halo columns of the tile: a box is dense in shared memory.)
template <int NIN>
__device__ __forceinline__ void wrap_fix(double *slot, const TmaIn<NIN> &in, const TmaGrid &tg, int i0, int j0, int zplane, bool wl, bool wr)
{
  constexpr int PER = NIN * TMY * 2;
  const int     n   = ((wl ? 1 : 0) + (wr ? 1 : 0)) * PER;
  for (int e = threadIdx.x; e < n; e += TMX * TMY) {
    const bool right = wl ? (e >= PER) : true;
    const int  r = e % PER, f = r / (TMY * 2), row = (r % (TMY * 2)) >> 1, cc = r & 1;
    const int  gcol = right ? cc : tg.nx - 2 + cc;
    const int  lcol = right ? THX + tg.nx - i0 + cc : cc;
    if (lcol < TLX) slot[f * TILE_STRIDE + (row + 1) * TLX + lcol] = in.p[f][(size_t)gcol + (size_t)tg.px * ((size_t)(j0 + row) + (size_t)tg.py * (size_t)zplane)];
  }
}

// Op interface:
//   static const int NIN, NR, MINB, STAGES;                             inputs, reductions, resident CTAs per SM aimed at, ring slots
//   struct Regs;                                                         per-thread operands that bypass the tiles
//   __device__ int  flags(int i, int j) const;                           per-thread constants of the column (wall tests), computed once
//   __device__ void prefetch(Regs &r, int off, int kl) const;            global loads for plane kl, issued one plane ahead
//   __device__ void cell(const TileView &tv, const Regs &r, int flags, int i, int j, int kl, int off, double *acc) const;
//   optional (TileOpDefaults): ZALIGN, SCRATCH, POST and
//   __device__ void post(const TileView &tv, int flags, bool owned, int i, int j, int kl, double &state) const;
// off = element offset of cell (i, j, kl) in the padded arrays of the launch (all fields of a launch share one layout,
// 32-bit: geom_build refuses slabs of 2^31 elements); the framework advances it by one plane per iteration, and every
// ring / barrier index is a running counter -- the plane loop of the 1-field operators was instruction-bound on index
// arithmetic (profiles/r01p: 180 warp instructions per warp-cell for 14 of arithmetic).
template <class Op>
__global__ void __launch_bounds__(TMX *TMY, Op::MINB) k_tma_march(const __grid_constant__ TmaIn<Op::NIN> in, const Op op, const TmaGrid tg, const double *carry, double *partials, double *result, unsigned *ticket)
{
  extern __shared__ __align__(128) unsigned char tma_smem[];
  constexpr int      TMS  = Op::STAGES;
  constexpr int      SLOT = Op::NIN * TILE_STRIDE; // doubles per ring slot
  constexpr unsigned TX_BYTES = Op::NIN * TILE_ELEMS * sizeof(double);
  double            *ring = reinterpret_cast<double *>(tma_smem);
  uint64_t          *full = reinterpret_cast<uint64_t *>(tma_smem + (size_t)TMS * SLOT * sizeof(double));
  double            *scratch = reinterpret_cast<double *>(tma_smem + (size_t)TMS * SLOT * sizeof(double) + TMS * sizeof(uint64_t));
  const uint32_t     ring_s = smem_u32(ring), full_s = smem_u32(full);
  const int tid = threadIdx.x, tx = tid & (TMX - 1), ty = tid / TMX;
  const int ntile = tg.ntx * tg.nty;
  const int tile = blockIdx.x % ntile, bz = blockIdx.x / ntile;
  const int bx = tile % tg.ntx, by = tile / tg.ntx;
  const int i0 = min(bx * TMX, (tg.nx - TMX + 1) & ~1), j0 = min(by * TMY, tg.ny - TMY); // i0 even
  int       k0, k1;
  z_chunk((tg.kend - tg.kbeg) / Op::ZALIGN, tg.nchunk, bz, k0, k1);
  k0 = k0 * Op::ZALIGN + tg.kbeg, k1 = k1 * Op::ZALIGN + tg.kbeg;
  const int  i = i0 + tx, j = j0 + ty;
  const bool owned = (i >= bx * TMX) && (i < tg.nx) && (j >= by * TMY);
  const int  nplanes = k1 - k0 + 2; // ring index r <-> local plane k0 - 1 + r
  const int  cx = i0 - THX, cy = j0 - 1;
  const bool wrap_l = tg.perx && i0 == 0, wrap_r = tg.perx && bx == tg.ntx - 1; // block-uniform

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < TMS; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0) {
    const int ahead = nplanes < TMS ? nplanes : TMS; // fill every slot
    for (int r = 0; r < ahead; ++r) {
      mbar_arrive_expect_tx(&full[r], TX_BYTES);
#pragma unroll
      for (int f = 0; f < Op::NIN; ++f) tma_load_box(ring + (size_t)r * SLOT + f * TILE_STRIDE, &in.m[f], cx, cy, k0 + r, &full[r]);
    }
  }
  double acc[Op::NR > 0 ? Op::NR : 1];
#pragma unroll
  for (int r = 0; r < (Op::NR > 0 ? Op::NR : 1); ++r) acc[r] = 0.;
  const int pstride = tg.px * tg.py;
  int       off = i + tg.px * (j + tg.py * (k0 + 1));
  const int fl  = owned ? op.flags(i, j) : 0;
  typename Op::Regs cur, nxt;
  if (owned && k0 < k1) op.prefetch(nxt, off, k0);
  if (k0 < k1) {
    mbar_wait_addr(full_s, 0);
    mbar_wait_addr(full_s + 8, 0);
    if (wrap_l || wrap_r) { // planes k0 - 1 and k0 (array planes k0, k0 + 1)
      wrap_fix<Op::NIN>(ring, in, tg, i0, j0, k0, wrap_l, wrap_r);
      wrap_fix<Op::NIN>(ring + SLOT, in, tg, i0, j0, k0 + 1, wrap_l, wrap_r);
      __syncthreads();
    }
  }
  TileView tv;
  tv.lc      = (ty + 1) * TLX + tx + THX;
  tv.scratch = scratch;
  double pstate = 0.;
  tv.p0 = ring;        // shifted into pm / p0 at the top of the first iteration
  tv.pp = ring + SLOT;
  int      sc = 1;          // ring slot of plane k+1 (after the advance at the top of the loop)
  unsigned par = 0;        // phase parity of slot sc
  int      zload = k0 + TMS; // TMA z coordinate of the next plane to request
  // PL planes per trip: one CTA barrier (and one refill round of the producer thread) per PL planes.  The 1-field operators
  // do so little per plane that the barrier was their main stall (profiles/r02n: PoissonTile 42 % of DRAM peak, 3.8 barrier
  // stall cycles per issue); the ring must hold the PL + 2 live planes plus what is in flight (STAGES >= PL + 3).
  constexpr int PL = Op::PLANES;
  static_assert(TMS >= PL + 3, "ring too small for the planes per trip");
  for (int k = k0; k < k1; k += PL) {
    const int np = (k1 - k) < PL ? (k1 - k) : PL;
#pragma unroll
    for (int q = 0; q < PL; ++q) {
      if (q < np) {
        cur = nxt;
        // operands read straight from global memory (one value per cell, no reuse) are requested a whole plane ahead
        if (owned && k + q + 1 < k1) op.prefetch(nxt, off + pstride, k + q + 1);
        sc = sc + 1 == TMS ? 0 : sc + 1;
        if (sc == 0) par ^= 1u;
        mbar_wait_addr(full_s + 8u * sc, par);
        if (wrap_l || wrap_r) { // plane k + q + 1 = array plane k + q + 2 just arrived in slot sc
          wrap_fix<Op::NIN>(ring + sc * SLOT, in, tg, i0, j0, k + q + 2, wrap_l, wrap_r);
          __syncthreads();
        }
        tv.pm = tv.p0;
        tv.p0 = tv.pp;
        tv.pp = ring + sc * SLOT;
        if (owned) op.cell(tv, cur, fl, i, j, k + q, off, acc);
        off += pstride;
      }
    }
    __syncthreads(); // every thread is done with planes k-1 .. k+np-2: their slots may be refilled
    if constexpr (Op::POST) {
#pragma unroll
      for (int q = 0; q < PL; ++q)
        if (q < np) op.post(tv, fl, owned, i, j, k + q, pstate);
    }
    if (tid == 0) {
      // the slot of plane k+np-1-m (m = 1 .. np) is m + 1 behind sc; refill the oldest first
      for (int m = np; m >= 1; --m) {
        if (zload < k0 + nplanes) {
          int sr = sc - 1 - m;
          if (sr < 0) sr += TMS;
          const uint32_t bar = full_s + 8u * sr, dst = ring_s + (uint32_t)(sr * SLOT * sizeof(double));
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(TX_BYTES) : "memory");
#pragma unroll
          for (int f = 0; f < Op::NIN; ++f)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst + (uint32_t)(f * TILE_STRIDE * sizeof(double))), "l"(&in.m[f]), "r"(cx), "r"(cy), "r"(zload), "r"(bar)
                         : "memory");
        }
        ++zload;
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
inline void tma_launch(Exec &ex, const Op &op, const double *const *fields, int px, int py, int nplanes_alloc, int nx, int ny, int kbeg, int kend, const double *carry, bool perx = false)
{
  if (kend <= kbeg) {
    if (Op::NR > 0) {
      if (carry) copy_d2d(ex, ex.d_result, carry, sizeof(double) * Op::NR);
      else dev_zero(ex, ex.d_result, sizeof(double) * Op::NR);
    }
    return;
  }
  TmaIn<Op::NIN> in;
  for (int f = 0; f < Op::NIN; ++f) in.m[f] = tensor_map_for(fields[f], px, py, nplanes_alloc), in.p[f] = fields[f];
  TmaGrid tg;
  tg.perx = perx ? 1 : 0;
  tg.px = px, tg.py = py, tg.nx = nx, tg.ny = ny, tg.kbeg = kbeg, tg.kend = kend;
  tg.ntx = (nx + TMX - 1) / TMX, tg.nty = (ny + TMY - 1) / TMY;
  if ((kend - kbeg) % Op::ZALIGN) throw Error(FL_ERR_INTERNAL, "plane range of a tile launch is not a multiple of the operator's alignment");
  tg.nchunk = tma_pick_chunks(tg.ntx * tg.nty, (kend - kbeg) / Op::ZALIGN, Op::MINB * ex.sm_count, ex.max_blocks);
  // one block partial per CTA: a reducing operator must not start more CTAs than ex.d_partials has room for
  if (Op::NR > 0 && (long)tg.ntx * tg.nty * tg.nchunk > ex.max_blocks) throw Error(FL_ERR_INTERNAL, "tile launch of a reducing operator exceeds the partial buffer");
  const size_t smem = (size_t)Op::STAGES * Op::NIN * TILE_STRIDE * sizeof(double) + Op::STAGES * sizeof(uint64_t) + Op::SCRATCH;
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
