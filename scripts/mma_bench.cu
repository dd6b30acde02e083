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

// variant with a second warp streaming 32 KB bulk copies (weights) into smem concurrently
template <int N, bool TRANSPOSED>
__global__ void __launch_bounds__(128, 1) k_mma_loaded(int iters, const uint8_t *wsrc, unsigned long long *cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, wbar[3];
  __shared__ uint32_t tslot;
  __shared__ volatile int done;
  const uint32_t sbase = smem_u32(smem);
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); for (int i = 0; i < 3; i++) mbar_init(smem_u32(&wbar[i]), 1); done = 0; fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tslot), 512); tmem_relinquish(); }
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tslot;
  if (threadIdx.x == 32) {
    // loader: keep three 32 KB copies in flight until the MMA thread is done
    uint32_t n = 0;
    for (int i = 0; i < 3; i++, n++) { mbar_arrive_expect_tx(smem_u32(&wbar[i]), 32768); bulk_g2s(sbase + 98304 + i * 32768, wsrc + (n % 40) * 32768, 32768, smem_u32(&wbar[i])); }
    while (!done) {
      const int s = n % 3;
      mbar_wait(smem_u32(&wbar[s]), ((n / 3) - 1) & 1);
      mbar_arrive_expect_tx(smem_u32(&wbar[s]), 32768);
      bulk_g2s(sbase + 98304 + s * 32768, wsrc + (n % 40) * 32768, 32768, smem_u32(&wbar[s]));
      n++;
    }
    for (int i = 0; i < 3; i++) { const uint32_t k = n - 3 + i; mbar_wait(smem_u32(&wbar[k % 3]), (k / 3) & 1); }
    if (blockIdx.x == 0) cycles[1] = n;
  }
  if (threadIdx.x == 0) {
    const uint32_t idesc = idesc_f16_f32(128, N);
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int j = 0; j < 2; j++) {
        if (TRANSPOSED) {
          const uint64_t bd = smem_desc_kmajor(sbase + ((it & 3) * 8192) + 2 * j * 2048, 2048, 128);
#pragma unroll
          for (int mb = 0; mb < 4; mb++) {
            const uint64_t ad = smem_desc_kmajor(sbase + 32768 + ((it & 1) * 32768) + 2 * j * 8192 + mb * 2048, 8192, 128);
            umma_f16(tmem + mb * 128, ad, bd, idesc, it > 0 || j > 0);
          }
        } else {
          const uint64_t ad = smem_desc_kmajor(sbase + ((it & 3) * 8192) + 2 * j * 2048, 2048, 128);
#pragma unroll
          for (int nb = 0; nb < 512 / N; nb++) {
            const uint64_t bd = smem_desc_kmajor(sbase + 32768 + ((it & 1) * 32768) + 2 * j * 8192 + nb * N * 16, 8192, 128);
            umma_f16(tmem + nb * N, ad, bd, idesc, it > 0 || j > 0);
          }
        }
      }
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    done = 1;
    if (blockIdx.x == 0) cycles[0] = (unsigned long long)(t1 - t0);
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

// per-chunk synchronisation cost: NW already-complete mbarrier waits + NC commits per 8 MMAs
template <int NW, int NC>
__global__ void __launch_bounds__(128, 1) k_mma_sync(int iters, unsigned long long *cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, dummy_full[2], dummy_empty[2];
  __shared__ uint32_t tslot;
  const uint32_t sbase = smem_u32(smem);
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    for (int i = 0; i < 2; i++) { mbar_init(smem_u32(&dummy_full[i]), 1); mbar_init(smem_u32(&dummy_empty[i]), 1); }
    fence_barrier_init();
  }
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tslot), 512); tmem_relinquish(); }
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tslot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = idesc_f16_f32(128, 128);
    const uint64_t d2k = smem_desc_kmajor(sbase, 2048, 128), d8k = smem_desc_kmajor(sbase, 8192, 128);
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int w = 0; w < NW; w++) mbar_wait(smem_u32(&dummy_full[w]), 1);   // fresh barrier: parity 1 is "complete"
      tc_fence_after_sync();
#pragma unroll
      for (int j = 0; j < 2; j++) {
        const uint64_t bd = d2k + (uint64_t)((((it & 3) * 8192) + 2 * j * 2048) >> 4);
#pragma unroll
        for (int mb = 0; mb < 4; mb++) {
          const uint64_t ad = d8k + (uint64_t)((65536 + ((it % 3) * 32768) + 2 * j * 8192 + mb * 2048) >> 4);
          umma_f16(tmem + mb * 128, ad, bd, idesc, it > 0 || j > 0);
        }
      }
#pragma unroll
      for (int c = 0; c < NC; c++) umma_commit(smem_u32(&dummy_empty[c]));
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) cycles[0] = (unsigned long long)(t1 - t0);
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

// MODE 0: tcgen05.fence::after_thread_sync only; 1: ld.acquire poll of a ready word + fence; 2: ld.volatile poll, no fence
template <int MODE>
__global__ void __launch_bounds__(128, 1) k_mma_poll(int iters, unsigned long long *cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, dummy_empty[2];
  __shared__ uint32_t tslot;
  __shared__ uint32_t ready;
  const uint32_t sbase = smem_u32(smem);
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    for (int i = 0; i < 2; i++) mbar_init(smem_u32(&dummy_empty[i]), 1);
    ready = 0x7fffffff;
    fence_barrier_init();
  }
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tslot), 512); tmem_relinquish(); }
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tslot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = idesc_f16_f32(128, 128);
    const uint64_t d2k = smem_desc_kmajor(sbase, 2048, 128), d8k = smem_desc_kmajor(sbase, 8192, 128);
    const uint32_t raddr = smem_u32(&ready);
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
      if (MODE == 1) {
        uint32_t v;
        do { asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(raddr) : "memory"); } while ((int)(v - (uint32_t)it) < 0);
        tc_fence_after_sync();
      } else if (MODE == 2) {
        uint32_t v;
        do { asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(raddr) : "memory"); } while ((int)(v - (uint32_t)it) < 0);
      } else {
        tc_fence_after_sync();
      }
#pragma unroll
      for (int j = 0; j < 2; j++) {
        const uint64_t bd = d2k + (uint64_t)((((it & 3) * 8192) + 2 * j * 2048) >> 4);
#pragma unroll
        for (int mb = 0; mb < 4; mb++) {
          const uint64_t ad = d8k + (uint64_t)((65536 + ((it % 3) * 32768) + 2 * j * 8192 + mb * 2048) >> 4);
          umma_f16(tmem + mb * 128, ad, bd, idesc, it > 0 || j > 0);
        }
      }
      umma_commit(smem_u32(&dummy_empty[0]));
      umma_commit(smem_u32(&dummy_empty[1]));
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) cycles[0] = (unsigned long long)(t1 - t0);
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

template <int MODE>
void run_poll(int iters, unsigned long long *d, int smem_bytes) {
  cudaFuncSetAttribute(k_mma_poll<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  for (int rep = 0; rep < 2; rep++) {
    k_mma_poll<MODE><<<148, 128, smem_bytes>>>(iters, d);
    cudaError_t err = cudaDeviceSynchronize();
    unsigned long long cyc = 0;
    cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
    if (rep) printf("POLL mode=%d (0 fence only, 1 ld.acquire+fence, 2 ld.volatile) + 2 commits: err=%d cycles/chunk=%.1f\n", MODE, (int)err, (double)cyc / iters);
  }
}

// transposed orientation with narrow N (edges): A = weights block (M=128), B = NE edges
template <int NE>
__global__ void __launch_bounds__(128, 1) k_mma_narrow(int iters, unsigned long long *cycles) {
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
    const uint32_t idesc = idesc_f16_f32(128, NE);
    const uint64_t d2k = smem_desc_kmajor(sbase, 2048, 128);
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
      // one 32 KB stage of W2-like weights: [16 k8][128 ch][16 B] -> 8 k-steps for one 128-channel unit
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const uint64_t ad = d2k + (uint64_t)((65536 + ((it % 3) * 32768) + 2 * j * 2048) >> 4);
        const uint64_t bd = d2k + (uint64_t)((((it & 3) * 16 + 2 * j) * 2048) >> 4);
        umma_f16(tmem + (it & 3) * 128, ad, bd, idesc, it > 3 || j > 0);
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

template <int NE>
void run_narrow(int iters, unsigned long long *d, int smem_bytes) {
  cudaFuncSetAttribute(k_mma_narrow<NE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  for (int rep = 0; rep < 2; rep++) {
    k_mma_narrow<NE><<<148, 128, smem_bytes>>>(iters, d);
    cudaError_t err = cudaDeviceSynchronize();
    unsigned long long cyc = 0;
    cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
    if (rep) printf("NARROW M128 N%d: err=%d cycles per 8 MMAs=%.1f  (ideal %d)\n", NE, (int)err, (double)cyc / iters, NE * 4);
  }
}

template <int NW, int NC>
void run_sync(int iters, unsigned long long *d, int smem_bytes) {
  cudaFuncSetAttribute(k_mma_sync<NW, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  for (int rep = 0; rep < 2; rep++) {
    k_mma_sync<NW, NC><<<148, 128, smem_bytes>>>(iters, d);
    cudaError_t err = cudaDeviceSynchronize();
    unsigned long long cyc = 0;
    cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
    if (rep) printf("SYNC waits=%d commits=%d: err=%d cycles/chunk=%.1f\n", NW, NC, (int)err, (double)cyc / iters);
  }
}

int main() {
  unsigned long long *d;
  cudaMalloc(&d, 16);
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
  run_narrow<128>(iters, d, smem_bytes);
  run_narrow<64>(iters, d, smem_bytes);
  run_narrow<32>(iters, d, smem_bytes);
  run_poll<0>(iters, d, smem_bytes);
  run_poll<1>(iters, d, smem_bytes);
  run_poll<2>(iters, d, smem_bytes);
  run_sync<0, 0>(iters, d, smem_bytes);
  run_sync<1, 0>(iters, d, smem_bytes);
  run_sync<2, 0>(iters, d, smem_bytes);
  run_sync<0, 1>(iters, d, smem_bytes);
  run_sync<0, 2>(iters, d, smem_bytes);
  run_sync<1, 1>(iters, d, smem_bytes);
  run_sync<2, 2>(iters, d, smem_bytes);
  // loaded variants
  uint8_t *w;
  cudaMalloc(&w, 40 * 32768);
  cudaMemset(w, 0x3c, 40 * 32768);
  cudaFuncSetAttribute(k_mma_loaded<256, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  cudaFuncSetAttribute(k_mma_loaded<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  for (int rep = 0; rep < 2; rep++) {
    for (int variant = 0; variant < 2; variant++) {
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      if (variant == 0) k_mma_loaded<256, false><<<148, 128, smem_bytes>>>(iters, w, d);
      if (variant == 1) k_mma_loaded<128, true><<<148, 128, smem_bytes>>>(iters, w, d);
      cudaEventRecord(e1);
      cudaError_t err = cudaDeviceSynchronize();
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      unsigned long long cyc[2] = {0, 0};
      cudaMemcpy(cyc, d, 16, cudaMemcpyDeviceToHost);
      double flops = 148.0 * iters * 2.0 * (2.0 * 128 * 512 * 16);
      printf("LOADED %-34s err=%d %.3f ms %.1f TFLOP/s cycles/chunk=%.1f copies/chunk=%.2f (%.1f GB/s/SM-agg)\n",
             variant == 0 ? "M128 N256" : "M128 N128 transposed", (int)err, ms, flops / (ms * 1e-3) / 1e12,
             (double)cyc[0] / iters, (double)cyc[1] / iters, 148.0 * cyc[1] * 32768 / (ms * 1e-3) / 1e9);
    }
  }
  return 0;
}
