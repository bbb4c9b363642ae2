// Microbenchmark: L2 -> shared-memory rate of cp.async.bulk.tensor (2-D boxes of ROWS x 128 B, SWIZZLE_128B) per SM
// with all 148 SMs loading from an L2-resident buffer, as a function of box height, loads in flight and issuing warps.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I multimodal_siamese_cd_b200/csrc -o /tmp/tma_rate tools/ubench/tma_rate.cu
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace b200cd;

__global__ void __launch_bounds__(128, 1) tma_rate_kernel(const __grid_constant__ CUtensorMap map, long long* out,
                                                          int iters, int rows, int stages, int issuers, int total_rows) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t full[4][16];
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int w = 0; w < 4; ++w)
      for (int s = 0; s < 16; ++s) mbar_init(&full[w][s], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (warp < issuers) {
    const int box_bytes = rows * 128;
    uint8_t* ring = smem + warp * stages * box_bytes;
    uint32_t row = (blockIdx.x * 7919u + warp * 104729u) % (total_rows - rows);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int s = it % stages;
      if (it >= stages) {
        const uint32_t par = ((it / stages) - 1) & 1;
        while (!mbar_try_wait(&full[warp][s], par)) {}
      }
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&full[warp][s], box_bytes);
        tma_load_2d(ring + s * box_bytes, &map, &full[warp][s], 0, row);
      }
      __syncwarp();
      row = (row + rows * 13u + 17u) % (total_rows - rows);
    }
    for (int it = iters; it < iters + stages; ++it) {  // drain
      const int s = it % stages;
      const uint32_t par = ((it / stages) - 1) & 1;
      while (!mbar_try_wait(&full[warp][s], par)) {}
    }
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) out[blockIdx.x * 4 + warp] = t1 - t0;
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncodeFn encode = reinterpret_cast<EncodeFn>(fn);
  long long* out;
  cudaMalloc(&out, 148 * 4 * sizeof(long long));
  cudaFuncSetAttribute(tma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 + 1024);
  const int pitches[2] = {128, 512};
  for (int pi = 0; pi < 2; ++pi) {
    const int pitch = pitches[pi];
    const int total_rows = (32 << 20) / pitch;  // 32 MB: L2 resident
    void* buf;
    cudaMalloc(&buf, 32 << 20);
    cudaMemset(buf, 0, 32 << 20);
    const int cfgs[][3] = {{160, 3, 1}, {160, 3, 2}, {160, 6, 1}, {128, 6, 1}, {128, 6, 2}, {64, 12, 1}, {64, 12, 2},
                           {256, 3, 1}, {256, 3, 2}, {80, 6, 1}, {80, 6, 2}, {80, 4, 4}, {160, 2, 4}, {32, 16, 2}, {32, 16, 4}};
    for (auto& c : cfgs) {
      const int rows = c[0], stages = c[1], issuers = c[2];
      CUtensorMap map;
      cuuint64_t dims[2] = {64, (cuuint64_t)total_rows};
      cuuint64_t strides[1] = {(cuuint64_t)pitch};
      cuuint32_t box[2] = {64, (cuuint32_t)rows};
      cuuint32_t es[2] = {1, 1};
      CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
      const int iters = 4000;
      const size_t smem = (size_t)issuers * stages * rows * 128 + 1024;
      for (int rep = 0; rep < 2; ++rep)
        tma_rate_kernel<<<148, 128, smem>>>(map, out, iters, rows, stages, issuers, total_rows);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      long long h[148 * 4];
      cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      double worst = 0, sum = 0;
      for (int b = 0; b < 148; ++b)
        for (int w = 0; w < issuers; ++w) { worst = h[b * 4 + w] > worst ? h[b * 4 + w] : worst; sum += h[b * 4 + w]; }
      const double bytes = (double)iters * rows * 128 * issuers;
      printf("pitch %4d  box %3d rows  %2d in flight x %d issuers (%3zu KB smem): %.1f B/clk/SM (slowest SM), %.1f (mean)\n", pitch,
             rows, stages, issuers, smem >> 10, bytes / worst, bytes / (sum / (148 * issuers)));
    }
    cudaFree(buf);
  }
  return 0;
}
