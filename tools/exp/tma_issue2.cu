// Hardware experiment: what bounds the TMA issue rate of one producer?  Each of `nwarps` warps (lane 0) runs its
// own ring of `depth` slots; per ring step it waits one mbarrier, arms it, and issues `k` tiled-2D loads of
// box_rows x 128 B.  Reports cycles per ring step.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../simple-hrnet_b200/csrc/ptx.cuh"

__global__ void __launch_bounds__(256, 1)
tma_issue_kernel(const __grid_constant__ CUtensorMap tm, int box_rows, int k, int depth, int iters, int total_rows,
                 int nwarps, int prefetch_desc, long long* cycles_out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (ptx::smem_u32(raw) + 1023u) & ~1023u;
  __shared__ uint64_t bars[8][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int load_bytes = box_rows * 128;
  const int slot_bytes = k * load_bytes;
  if (threadIdx.x == 0) {
    for (int w = 0; w < nwarps; ++w)
      for (int i = 0; i < depth; ++i) ptx::mbar_init(ptx::smem_u32(&bars[w][i]), 1);
    ptx::fence_mbar_init();
    if (prefetch_desc) ptx::prefetch_tmap(&tm);
  }
  __syncthreads();
  if (warp < nwarps && lane == 0) {
    const int nbox = total_rows / box_rows;
    int boxi = ((blockIdx.x * 8 + warp) * 977) % nbox;
    const uint32_t ring = base + warp * depth * slot_bytes;
    long long t0 = clock64();
    for (int i = 0; i < iters + depth; ++i) {
      const int slot = i % depth;
      const uint32_t bar = ptx::smem_u32(&bars[warp][slot]);
      if (i >= depth) ptx::mbar_wait(bar, ((i / depth) - 1) & 1);
      if (i < iters) {
        ptx::mbar_expect_tx(bar, (uint32_t)slot_bytes);
        for (int j = 0; j < k; ++j) {
          ptx::tma_load_2d(ring + slot * slot_bytes + j * load_bytes, &tm, bar, 0, boxi * box_rows);
          boxi += 37; if (boxi >= nbox) boxi -= nbox;
        }
      }
    }
    cycles_out[blockIdx.x * 8 + warp] = clock64() - t0;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

extern "C" int exp_tma_issue_pitch(void* mat, int total_rows, int pitch_elems, int box_rows, int k, int depth, int iters, int grid, int nwarps,
                             int prefetch_desc, long long* cycles_dev) {
  cudaDriverEntryPointQueryResult q; void* f = nullptr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || !f) return -1;
  EncodeTiledFn enc = (EncodeTiledFn)f;
  CUtensorMap tm;
  cuuint64_t dims[2] = {(cuuint64_t)pitch_elems, (cuuint64_t)total_rows};
  cuuint64_t st[1] = {(cuuint64_t)pitch_elems * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, mat, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return -2;
  const int smem = 1024 + nwarps * depth * k * box_rows * 128;
  if (smem > 227 * 1024) return -5;
  if (cudaFuncSetAttribute(tma_issue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -3;
  tma_issue_kernel<<<grid, 256, smem>>>(tm, box_rows, k, depth, iters, total_rows, nwarps, prefetch_desc, cycles_dev);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { fprintf(stderr, "exp_tma_issue: %s\n", cudaGetErrorString(e)); return -4; }
  return 0;
}
