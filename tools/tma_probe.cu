// tma_probe.cu -- stand-alone probe of the TMA box load used by fluca_b200/csrc/tma.h.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tma_probe tma_probe.cu
// run:   ./tma_probe <variant>   (each variant in its own process: a fault kills the context)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); return 2; } } while (0)

static int LX = 34, LY = 10;
__constant__ int dLX, dLY;

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int VAR>
__global__ void probe(const __grid_constant__ CUtensorMap map, const CUtensorMap *gmap, double *out, int c0, int c1, int c2)
{
  const int LX = dLX, LY = dLY;
  extern __shared__ __align__(128) unsigned char sm[];
  double   *tile = reinterpret_cast<double *>(sm);
  uint64_t *bar  = reinterpret_cast<uint64_t *>(sm + 4096);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(bar)), "r"(1) : "memory");
    if (VAR != 3) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (VAR == 6) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(bar)) : "memory");
    else
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(LX * LY * 8) : "memory");
    if (VAR == 6) { }
    else if (VAR == 9)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(s32(tile)), "l"(gmap), "r"(c0), "r"(c1), "r"(c2), "r"(s32(bar)) : "memory");
    else if (VAR == 2)
      asm volatile("cp.async.bulk.tensor.3d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(s32(tile)), "l"(&map), "r"(c0), "r"(c1), "r"(c2), "r"(s32(bar)) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(s32(tile)), "l"(&map), "r"(c0), "r"(c1), "r"(c2), "r"(s32(bar)) : "memory");
  }
  unsigned ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred P_OUT;\n\tmbarrier.try_wait.parity.shared::cta.b64 P_OUT, [%1], %2;\n\tselp.b32 %0, 1, 0, P_OUT;\n\t}" : "=r"(ok) : "r"(s32(bar)), "r"(0) : "memory");
  }
  for (int e = threadIdx.x; e < LX * LY; e += blockDim.x) out[e] = tile[e];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv)
{
  const int var = argc > 1 ? atoi(argv[1]) : 1;
  if (var == 8) LX = 32, LY = 8;
  if (var >= 10) LX = 36, LY = 10;
  cudaMemcpyToSymbol(dLX, &LX, 4); cudaMemcpyToSymbol(dLY, &LY, 4);
  const int px = 40, py = 14, nz = 5;
  std::vector<double> h((size_t)px * py * nz);
  for (size_t i = 0; i < h.size(); ++i) h[i] = 1.0 + (double)i;
  double *d, *out;
  CK(cudaMalloc(&d, h.size() * 8));
  CK(cudaMalloc(&out, LX * LY * 8));
  CK(cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice));
  void *fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  alignas(64) CUtensorMap m;
  cuuint64_t dims[3] = {(cuuint64_t)px, (cuuint64_t)py, (cuuint64_t)nz}, strides[2] = {(cuuint64_t)px * 8, (cuuint64_t)px * py * 8};
  cuuint32_t box[3] = {(cuuint32_t)LX, (cuuint32_t)LY, 1}, es[3] = {1, 1, 1};
  CUtensorMapDataType dt = var == 4 ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
  CUresult r = ((EncodeTiledFn)fp)(&m, dt, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, var == 5 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("variant %d: encode -> %d\n", var, (int)r);
  if (r != CUDA_SUCCESS) return 3;
  const int c0 = var == 7 ? 0 : (var == 10 ? -2 : (var == 11 ? 30 : (var == 12 ? 7 : -1))), c1 = var == 7 ? 0 : (var == 11 ? 9 : -1), c2 = var == 12 ? 4 : 2;
  CUtensorMap *gm; CK(cudaMalloc(&gm, 128)); CK(cudaMemcpy(gm, &m, 128, cudaMemcpyHostToDevice));
  if (var == 2) probe<2><<<1, 128, 8192>>>(m, gm, out, c0, c1, c2);
  else if (var == 3) probe<3><<<1, 128, 8192>>>(m, gm, out, c0, c1, c2);
  else if (var == 6) probe<6><<<1, 128, 8192>>>(m, gm, out, c0, c1, c2);
  else if (var == 9) probe<9><<<1, 128, 8192>>>(m, gm, out, c0, c1, c2);
  else probe<1><<<1, 128, 8192>>>(m, gm, out, c0, c1, c2);
  CK(cudaDeviceSynchronize());
  std::vector<double> o(LX * LY);
  CK(cudaMemcpy(o.data(), out, o.size() * 8, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int jj = 0; jj < LY; ++jj)
    for (int ii = 0; ii < LX; ++ii) {
      const int i = c0 + ii, j = c1 + jj;
      double want = (i < 0 || j < 0 || i >= px || j >= py) ? 0.0 : h[(size_t)i + px * (j + (size_t)py * c2)];
      if (o[jj * LX + ii] != want) ++bad;
    }
  printf("variant %d: %d mismatches of %d (corner %g, first interior %g)\n", var, bad, LX * LY, o[0], o[LX + 1]);
  return bad ? 1 : 0;
}
