// Microbenchmark: issue rate of tcgen05.mma (bf16, K = 16 per instruction) with operands resident in shared memory —
// cycles per MMA for M = 128 (cta_group::1) and M = 256 (cta_group::2), N in {64, 128, 256}, K-major and MN-major.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I multimodal_siamese_cd_b200/csrc -o /tmp/mma_rate tools/ubench/mma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace b200cd;

template <int N, int CG, int MAJOR>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(long long* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (64 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    if (CG == 2) { tmem_alloc_2cta(&tmem_slot, 512); tmem_relinquish_2cta(); }
    else { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const bool leader = CG == 1 || cluster_ctarank() == 0;
  long long t0 = 0, t1 = 0;
  if (warp == 1 && leader) {
    constexpr uint32_t idesc = make_idesc_bf16(CG == 2 ? 256 : 128, N, MAJOR, MAJOR);
    const uint32_t desc_hi = smem_desc_hi(1024);
    const uint32_t a_lo = smem_desc_lo(smem_u32(smem), MAJOR ? 8192 : 16);
    const uint32_t b_lo = smem_desc_lo(smem_u32(smem) + 16384, MAJOR ? 8192 : 16);
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t ka = MAJOR ? k * (2048 >> 4) : 2 * k;
          if (CG == 2) umma_bf16_2cta_lo(tmem_base, a_lo + ka, b_lo + ka, desc_hi, idesc, true);
          else umma_bf16_lo(tmem_base, a_lo + ka, b_lo + ka, desc_hi, idesc, true);
        }
      }
      __syncwarp();
    }
    if (elect_one_sync()) {
      if (CG == 2) umma_commit_2cta(&bar, 1);
      else umma_commit(&bar);
    }
    __syncwarp();
    while (!mbar_try_wait(&bar, 0)) {}
    t1 = clock64();
    if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_2cta(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

template <int N, int CG, int MAJOR>
void run(const char* name) {
  long long* d;
  cudaMalloc(&d, 8);
  const int iters = 2000;
  const int smem = 66 * 1024 + 1024;
  cudaFuncSetAttribute(mma_rate_kernel<N, CG, MAJOR>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(CG == 2 ? 148 : 148);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CG;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) cudaLaunchKernelEx(&cfg, mma_rate_kernel<N, CG, MAJOR>, d, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  const double per = double(h) / (4.0 * iters);
  const double m_per_sm = 128.0;
  printf("%-34s %7.1f cycles / MMA   (%5.1f %% of the 8192 flop/clk/SM peak)  %s\n", name, per,
         100.0 * (2.0 * m_per_sm * N * 16 / per) / 8192.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<64, 1, 0>("M128 N64  cta_group::1 K-major");
  run<128, 1, 0>("M128 N128 cta_group::1 K-major");
  run<256, 1, 0>("M128 N256 cta_group::1 K-major");
  run<64, 2, 0>("M256 N64  cta_group::2 K-major");
  run<128, 2, 0>("M256 N128 cta_group::2 K-major");
  run<256, 2, 0>("M256 N256 cta_group::2 K-major");
  run<64, 1, 1>("M128 N64  cta_group::1 MN-major");
  run<128, 1, 1>("M128 N128 cta_group::1 MN-major");
  run<256, 1, 1>("M128 N256 cta_group::1 MN-major");
  return 0;
}
