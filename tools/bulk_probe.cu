// bulk_probe.cu — 1-D bulk copy (cp.async.bulk, no tensor map) through an mbarrier: is the async copy unit usable at all on this box?
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void k(const int16_t* src, int16_t* out)
{
  __shared__ __align__(128) int16_t buf[512];
  __shared__ __align__(8) unsigned long long bar;
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)) : "memory"); asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.release.cta.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bar)), "r"(1024) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(smem_u32(buf)), "l"(src), "r"(1024), "r"(smem_u32(&bar)) : "memory");
  }
  asm volatile("{\n\t.reg .pred P1;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n\t@P1 bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" :: "r"(smem_u32(&bar)) : "memory");
  __syncthreads();
  for (int i = threadIdx.x; i < 512; i += blockDim.x) out[i] = buf[i];
}
int main()
{
  int16_t h[512]; for (int i = 0; i < 512; i++) h[i] = (int16_t)(i * 3 + 1);
  int16_t *d, *o; cudaMalloc(&d, 1024); cudaMalloc(&o, 1024); cudaMemcpy(d, h, 1024, cudaMemcpyHostToDevice);
  k<<<1, 64>>>(d, o); cudaError_t e = cudaDeviceSynchronize(); printf("bulk copy launch: %s\n", cudaGetErrorString(e)); if (e) return 1;
  int16_t r[512]; cudaMemcpy(r, o, 1024, cudaMemcpyDeviceToHost); int bad = 0; for (int i = 0; i < 512; i++) bad += r[i] != h[i]; printf("%d wrong\n", bad);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); printf("%s cc %d.%d\n", p.name, p.major, p.minor);
  return 0;
}
