// tma_probe.cu — stand-alone check of the bulk-tensor copy idiom K2 uses (descriptor array in global memory, 2-D int16 boxes, one issuing thread, mbarrier wait).
// nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu && ./tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>
#include <stdlib.h>
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, int bytes) { asm volatile("mbarrier.arrive.expect_tx.release.cta.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, int parity)
{
  asm volatile("{\n\t.reg .pred P1;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d_cta(void* smemDst, const CUtensorMap* map, int x, int y, unsigned long long* bar)
{
  asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               :: "r"(smem_u32(smemDst)), "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smemDst, const CUtensorMap* map, int x, int y, unsigned long long* bar)
{
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               :: "r"(smem_u32(smemDst)), "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}
constexpr int BW = 24, BH = 23;
struct MapPack { CUtensorMap m[3]; };
__constant__ CUtensorMap cMaps[3];
// variant 0: descriptor in global memory; 1: global memory + tensormap proxy fence; 2: __grid_constant__ parameter; 3: __constant__ memory
__global__ void probe(const CUtensorMap* gmaps, const __grid_constant__ MapPack pack, int variant_, int which, int x, int y, int16_t* out, unsigned* info)
{
  int variant = variant_;
  if (variant == 4 || variant == 5) variant = 2;
  const CUtensorMap* maps = variant == 2 ? pack.m : variant == 3 ? cMaps : gmaps;
  extern __shared__ __align__(128) int16_t smem[];
  __shared__ unsigned pad[25]; __shared__ int pad2[35];
  __shared__ __align__(8) unsigned long long sBar;
  pad[threadIdx.x % 25] = 0; pad2[threadIdx.x % 35] = 0;
  const int stage = variant >= 10 ? variant - 10 : 3; if (variant >= 10) variant = 0;
  if (threadIdx.x == 0) { info[0] = smem_u32(smem); info[1] = smem_u32(&sBar); mbar_init(&sBar, 1); }
  __syncthreads();
  if (stage == 0) return;
  int16_t* dst = smem + ((128 - (smem_u32(smem) & 127)) & 127) / 2;
  if (threadIdx.x == 0) {
    if (variant == 1) asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" :: "l"(maps + which) : "memory");
    mbar_expect_tx(&sBar, BW * BH * 2); if (stage >= 2) { if (variant_ == 4) tma_load_2d_cta(dst, maps + which, x, y, &sBar); else tma_load_2d(dst, maps + which, x, y, &sBar); }
  }
  if (stage >= 3) mbar_wait(&sBar, 0); else __nanosleep(20000);
  __syncthreads();
  for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = dst[i];
}
int main(int argc, char** argv)
{
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  const int W = 416, H = 240;
  std::vector<int16_t> h((size_t)W * H); for (size_t i = 0; i < h.size(); i++) h[i] = (int16_t)(i * 7 + 3);
  int16_t* d; cudaMalloc(&d, h.size() * 2); cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  printf("entry point: %d %d %p\n", (int)e, (int)q, fn);
  std::vector<CUtensorMap> maps(3);
  for (int i = 0; i < 3; i++) {
    cuuint64_t dims[2] = {W, H}, strides[1] = {W * 2}; cuuint32_t box[2] = {BW, BH}, es[2] = {1, 1};
    CUresult r = ((EncodeFn)fn)(&maps[i], CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode %d: %d\n", i, (int)r);
  }
  CUtensorMap* dm; cudaMalloc(&dm, sizeof(CUtensorMap) * 3); cudaMemcpy(dm, maps.data(), sizeof(CUtensorMap) * 3, cudaMemcpyHostToDevice);
  MapPack pack; for (int i = 0; i < 3; i++) pack.m[i] = maps[i];
  cudaMemcpyToSymbol(cMaps, maps.data(), sizeof(CUtensorMap) * 3);
  printf("variant %d\n", variant);
  int16_t* dout; cudaMalloc(&dout, BW * BH * 2); unsigned* dinfo; cudaMalloc(&dinfo, 16);
  for (int t = 0; t < 3; t++) {
    const int x = 5 + 13 * t, y = 3 + 50 * t;
    if (variant == 5) {
      cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(1); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = 4096;
      cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1; cfg.attrs = at; cfg.numAttrs = 1;
      e = cudaLaunchKernelEx(&cfg, probe, (const CUtensorMap*)dm, pack, variant, t, x, y, dout, dinfo); printf("launchEx: %s\n", cudaGetErrorString(e));
    } else probe<<<1, 64, 4096>>>(dm, pack, variant, t, x, y, dout, dinfo);
    e = cudaDeviceSynchronize(); printf("launch %d: %s\n", t, cudaGetErrorString(e)); if (e) return 1;
    std::vector<int16_t> o(BW * BH); unsigned info[2]; cudaMemcpy(o.data(), dout, o.size() * 2, cudaMemcpyDeviceToHost); cudaMemcpy(info, dinfo, 8, cudaMemcpyDeviceToHost);
    int bad = 0; for (int j = 0; j < BH; j++) for (int i = 0; i < BW; i++) bad += o[j * BW + i] != h[(size_t)(y + j) * W + x + i];
    printf("box at (%d,%d): %d wrong samples; dynamic smem at 0x%x, barrier at 0x%x\n", x, y, bad, info[0], info[1]);
  }
  return 0;
}
