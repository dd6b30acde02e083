// Micro-benchmark: back-to-back tcgen05.mma issue rate from shared memory (no loads),
// for the operand shapes the edge kernel can use.  nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>
#include "../chemeleon_b200/csrc/cb2_ptx.cuh"
using namespace cb2::ptx;

template <int N>
__global__ void __launch_bounds__(128, 1) k_mma(int iters, int a_rows_lbo, unsigned long long *cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const uint32_t sbase = smem_u32(smem);
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tslot), 512); tmem_relinquish(); }
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tslot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = idesc_f16_f32(128, N);
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
      // emulate a 32-wide K chunk: 2 k-steps, operands at different smem offsets
#pragma unroll
      for (int j = 0; j < 2; j++) {
        const uint64_t ad = smem_desc_kmajor(sbase + ((it & 3) * 8192) + 2 * j * 2048, 2048, 128);
#pragma unroll
        for (int nb = 0; nb < 512 / N; nb++) {
          const uint64_t bd = smem_desc_kmajor(sbase + 65536 + ((it % 3) * 32768) + 2 * j * 8192 + nb * N * 16, 8192, 128);
          umma_f16(tmem + nb * N, ad, bd, idesc, it > 0 || j > 0);
        }
      }
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) cycles[0] = (unsigned long long)(t1 - t0);
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

// transposed orientation: A = weights (M=128 channel block), B = edges (N=128)
__global__ void __launch_bounds__(128, 1) k_mma_t(int iters, unsigned long long *cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const uint32_t sbase = smem_u32(smem);
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tslot), 512); tmem_relinquish(); }
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tslot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = idesc_f16_f32(128, 128);
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int j = 0; j < 2; j++) {
        const uint64_t bd = smem_desc_kmajor(sbase + ((it & 3) * 8192) + 2 * j * 2048, 2048, 128);  // edges
#pragma unroll
        for (int mb = 0; mb < 4; mb++) {
          const uint64_t ad = smem_desc_kmajor(sbase + 65536 + ((it % 3) * 32768) + 2 * j * 8192 + mb * 2048, 8192, 128);
          umma_f16(tmem + mb * 128, ad, bd, idesc, it > 0 || j > 0);
        }
      }
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) cycles[0] = (unsigned long long)(t1 - t0);
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
  unsigned long long *d;
  cudaMalloc(&d, 8);
  const int smem_bytes = 200 * 1024;
  cudaFuncSetAttribute(k_mma<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  cudaFuncSetAttribute(k_mma<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  cudaFuncSetAttribute(k_mma_t, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  const int iters = 20000;
  for (int rep = 0; rep < 2; rep++) {
    for (int variant = 0; variant < 3; variant++) {
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      if (variant == 0) k_mma<256><<<148, 128, smem_bytes>>>(iters, 0, d);
      if (variant == 1) k_mma<128><<<148, 128, smem_bytes>>>(iters, 0, d);
      if (variant == 2) k_mma_t<<<148, 128, smem_bytes>>>(iters, d);
      cudaEventRecord(e1);
      cudaError_t err = cudaDeviceSynchronize();
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      unsigned long long cyc = 0;
      cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
      double flops = 148.0 * iters * 2.0 * (2.0 * 128 * 512 * 16);
      const char *nm[3] = {"M128 N256 (edges x channels)", "M128 N128 (edges x channels)", "M128 N128 transposed (channels x edges)"};
      printf("%-42s err=%d  %.3f ms  %.1f TFLOP/s  cycles/chunk(K=32,N=512)=%.1f\n", nm[variant], (int)err, ms,
             flops / (ms * 1e-3) / 1e12, (double)cyc / iters);
    }
  }
  return 0;
}
