// Hardware experiment (not product code): per-SM TMA load throughput and latency on B200 for the box shapes the
// conv kernels use.  One thread per CTA keeps `depth` tiled-2D TMA loads in flight over an L2-resident matrix.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../simple-hrnet_b200/csrc/ptx.cuh"

__global__ void __launch_bounds__(32, 1)
tma_bw_kernel(const __grid_constant__ CUtensorMap tm, int box_rows, int row_bytes, int depth, int iters, int total_rows,
              long long* cycles_out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (ptx::smem_u32(raw) + 1023u) & ~1023u;
  __shared__ uint64_t bars[16];
  const int slot_bytes = (box_rows * row_bytes + 1023) / 1024 * 1024;
  if (threadIdx.x == 0) {
    for (int i = 0; i < depth; ++i) ptx::mbar_init(ptx::smem_u32(&bars[i]), 1);
    ptx::fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int nbox = total_rows / box_rows;
    int boxi = (blockIdx.x * 977) % nbox;
    long long t0 = clock64();
    for (int i = 0; i < iters + depth; ++i) {
      const int slot = i % depth;
      if (i >= depth) ptx::mbar_wait(ptx::smem_u32(&bars[slot]), ((i / depth) - 1) & 1);
      if (i < iters) {
        ptx::mbar_expect_tx(ptx::smem_u32(&bars[slot]), (uint32_t)(box_rows * row_bytes));
        ptx::tma_load_2d(base + slot * slot_bytes, &tm, ptx::smem_u32(&bars[slot]), 0, boxi * box_rows);
        boxi += 37; if (boxi >= nbox) boxi -= nbox;
      }
    }
    cycles_out[blockIdx.x] = clock64() - t0;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// mat: [total_rows][pitch_elems] fp16 device; loads boxes of box_rows x (row_bytes/2) elements
extern "C" int exp_tma_bw(void* mat, int total_rows, int pitch_elems, int box_rows, int row_bytes, int depth, int iters,
                          int grid, long long* cycles_dev) {
  cudaDriverEntryPointQueryResult q; void* f = nullptr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || !f) return -1;
  EncodeTiledFn enc = (EncodeTiledFn)f;
  CUtensorMapSwizzle sw = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUtensorMap tm;
  cuuint64_t dims[2] = {(cuuint64_t)pitch_elems, (cuuint64_t)total_rows};
  cuuint64_t st[1] = {(cuuint64_t)pitch_elems * 2};
  cuuint32_t box[2] = {(cuuint32_t)(row_bytes / 2), (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, mat, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return -2;
  const int slot_bytes = (box_rows * row_bytes + 1023) / 1024 * 1024;
  const int smem = 1024 + depth * slot_bytes;
  if (cudaFuncSetAttribute(tma_bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -3;
  tma_bw_kernel<<<grid, 32, smem>>>(tm, box_rows, row_bytes, depth, iters, total_rows, cycles_dev);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { fprintf(stderr, "exp_tma_bw: %s\n", cudaGetErrorString(e)); return -4; }
  return 0;
}
