// tma_probe_ref.cu — the same copy through libcu++'s own wrappers (cuda/barrier), to compare with tma_probe.cu.  usage: tma_probe_ref BW BH
#include <cuda.h>
#include <cuda/barrier>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <vector>
#include <dlfcn.h>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
__global__ void kernel(const __grid_constant__ CUtensorMap tensor_map, int x, int y, int n, int16_t* out)
{
  __shared__ alignas(128) int16_t smem_buffer[4096];
  #pragma nv_diag_suppress static_var_with_dynamic_init
  __shared__ barrier bar;
  if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
  __syncthreads();
  barrier::arrival_token token;
  if (threadIdx.x == 0) {
    cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, &tensor_map, x, y, bar);
    token = cuda::device::barrier_arrive_tx(bar, 1, n * 2);
  } else token = bar.arrive();
  bar.wait(std::move(token));
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = smem_buffer[i];
}
int main(int argc, char** argv)
{
  const int BW = argc > 1 ? atoi(argv[1]) : 24, BH = argc > 2 ? atoi(argv[2]) : 23, dt = argc > 3 ? atoi(argv[3]) : 1, sw = argc > 4 ? atoi(argv[4]) : 0, dl = argc > 5 ? atoi(argv[5]) : 0;
  const int W = 416, H = 240;
  std::vector<int16_t> h((size_t)W * H); for (size_t i = 0; i < h.size(); i++) h[i] = (int16_t)(i * 7 + 3);
  int16_t* d; cudaMalloc(&d, h.size() * 2); cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  if (dl) { void* h = dlopen("libcuda.so.1", RTLD_NOW); void* f2 = h ? dlsym(h, "cuTensorMapEncodeTiled") : nullptr; printf("dlsym %p vs entry point %p\n", f2, fn); if (f2) fn = f2; }
  CUtensorMap map;
  cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H}, strides[1] = {(cuuint64_t)W * 2}; cuuint32_t box[2] = {(cuuint32_t)BW, (cuuint32_t)BH}, es[2] = {1, 1};
  CUresult r = ((EncodeFn)fn)(&map, (CUtensorMapDataType)dt, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, (CUtensorMapSwizzle)sw, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("box %dx%d dtype %d swizzle %d encode: %d\n", BW, BH, dt, sw, (int)r);
  { const unsigned long long* w = (const unsigned long long*)&map; for (int i = 0; i < 16; i++) printf("%016llx%c", w[i], i % 4 == 3 ? '\n' : ' '); }
  int16_t* dout; cudaMalloc(&dout, 8192);
  const int x = 18, y = 53;
  kernel<<<1, 64>>>(map, x, y, BW * BH, dout);
  cudaError_t e = cudaDeviceSynchronize(); printf("launch: %s\n", cudaGetErrorString(e)); if (e) return 1;
  std::vector<int16_t> o(BW * BH); cudaMemcpy(o.data(), dout, o.size() * 2, cudaMemcpyDeviceToHost);
  int bad = 0; for (int j = 0; j < BH; j++) for (int i = 0; i < BW; i++) bad += o[j * BW + i] != h[(size_t)(y + j) * W + x + i];
  printf("%d wrong samples\n", bad);
  return 0;
}
