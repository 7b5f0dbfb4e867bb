// Hardware experiment (not product code): what does a tcgen05.mma stream cost when TMA loads refill its operand
// stages at the same time?  One CTA per SM: warps 0-1 TMA producers (alternate stages), warp 2 MMA issuer, a ring of
// `stages` x (A 128 x 64 fp16 | B b_rows x 64 fp16) in shared memory, tiled loads from an L2-resident matrix, four
// K16 MMAs (M = 128, N = n_mma) per stage, accumulator in TMEM (never drained).  mode bit 0: MMAs on, bit 1: TMA on.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../simple-hrnet_b200/csrc/ptx.cuh"

__global__ void __launch_bounds__(384, 1)
mix_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmI,
           const __grid_constant__ CUtensorMap tmT, const __grid_constant__ CUtensorMap tmW, int n_mma, int b_rows, int stages,
           int iters, int mode, int total_rows, long long* out) {
  // mode bits 2-3: A loads 0 = random 2-D boxes, 1 = im2col over act[64][18][24][192] exactly like the C = 192 branch conv,
  // 2 = tiled 4-D box {64 ch, 24 w, 5 h, 1 n} (120 pixel rows, zero-filled halo); bit 4: B = the conv's weight k-blocks
  // (the same 27 boxes of w[192][1728] in every CTA) instead of random boxes
  const int amode = (mode >> 2) & 3, bmode = (mode >> 4) & 1;
  extern __shared__ uint8_t raw[];
  const uint32_t base = (ptx::smem_u32(raw) + 1023u) & ~1023u;
  __shared__ uint64_t full[8], empty[8], done, never;
  __shared__ volatile int stop;
  __shared__ uint32_t tslot;
  const int warp = ptx::warp_idx_uniform();
  const int b_bytes = (b_rows * 128 + 1023) / 1024 * 1024;
  const int nb_bytes = (n_mma * 128 + 1023) / 1024 * 1024;
  const int stage_bytes = 16384 + (b_bytes > nb_bytes ? b_bytes : nb_bytes);
  const bool do_mma = mode & 1, do_tma = mode & 2;
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) { ptx::mbar_init(ptx::smem_u32(&full[i]), 1); ptx::mbar_init(ptx::smem_u32(&empty[i]), 1); }
    ptx::mbar_init(ptx::smem_u32(&done), 1);
    ptx::mbar_init(ptx::smem_u32(&never), 1);
    stop = 0;
    ptx::fence_mbar_init();
  }
  if (warp == 2) ptx::tmem_alloc(ptx::smem_u32(&tslot), 256);
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem = tslot;
  const long long t0 = clock64();
  if (warp < 2) {
    if (do_tma && ptx::elect_one()) {
      const int nboxA = total_rows / 128, nboxB = total_rows / b_rows;
      int ba = (blockIdx.x * 977 + warp * 131) % nboxA, bb = (blockIdx.x * 613 + warp * 71) % nboxB;
      for (int i = warp; i < iters; i += 2) {
        const int s = i % stages;
        ptx::mbar_wait(ptx::smem_u32(&empty[s]), ((i / stages) & 1) ^ 1);
        const uint32_t f = ptx::smem_u32(&full[s]);
        const int a_bytes = amode == 2 ? 120 * 128 : 16384;
        ptx::mbar_expect_tx(f, (uint32_t)(a_bytes + b_rows * 128));
        const int kb = i % 27, tile = (blockIdx.x * 7 + i / 27) % 216;      // 216 M-tiles of 128 pixels, 27 k-blocks each
        const int tap = kb / 3, c0 = (kb % 3) * 64, r = tap / 3, sx = tap % 3;
        if (amode == 0) {
          ptx::tma_load_2d(base + s * stage_bytes, &tmA, f, 0, ba * 128);
        } else if (amode == 1) {
          const int m0 = tile * 128, img = m0 / 432, rem = m0 % 432, oh0 = rem / 24, ow0 = rem % 24;
          ptx::tma_load_im2col_4d(base + s * stage_bytes, &tmI, f, c0, ow0 - 1, oh0 - 1, img, (uint16_t)sx, (uint16_t)r);
        } else {
          const int img = (tile / 4) % 64, oh0 = (tile % 4) * 5;
          ptx::tma_load_4d(base + s * stage_bytes, &tmT, f, c0, sx - 1, oh0 + r - 1, img);
        }
        if (bmode == 0) ptx::tma_load_2d(base + s * stage_bytes + 16384, &tmB, f, 0, bb * b_rows);
        else ptx::tma_load_2d(base + s * stage_bytes + 16384, &tmW, f, tap * 192 + c0, 0);
        ba += 37; if (ba >= nboxA) ba -= nboxA;
        bb += 41; if (bb >= nboxB) bb -= nboxB;
      }
    }
    __syncwarp();
  } else if (warp == 2) {
    if (ptx::elect_one()) {
      const uint32_t idesc = ptx::umma_idesc_f16(128, n_mma);
      for (int i = 0; i < iters; ++i) {
        const int s = i % stages;
        if (do_tma) ptx::mbar_wait(ptx::smem_u32(&full[s]), (i / stages) & 1);
        ptx::tc_fence_after_sync();
        if (do_mma) {
          const uint64_t adesc = ptx::umma_desc_kmajor(base + s * stage_bytes, 128u, 1024u);
          const uint64_t bdesc = ptx::umma_desc_kmajor(base + s * stage_bytes + 16384, 128u, 1024u);
#pragma unroll
          for (int k = 0; k < 4; ++k) ptx::mma_f16_ss(tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, 1u);
          ptx::mma_commit(ptx::smem_u32(&empty[s]));
        } else {
          ptx::mbar_arrive(ptx::smem_u32(&empty[s]));
        }
      }
      if (do_mma) { ptx::mma_commit(ptx::smem_u32(&done)); ptx::mbar_wait(ptx::smem_u32(&done), 0); }
      out[blockIdx.x] = clock64() - t0;
      stop = 1;
    }
    __syncwarp();
  } else if (warp >= 4 && warp - 4 < ((mode >> 8) & 15)) {
    // polling warps: every lane spins on an mbarrier phase that never completes (what 256 epilogue threads waiting for
    // their accumulator do in the conv kernels)
    while (!stop) { if (ptx::mbar_try_wait(ptx::smem_u32(&never), 0)) break; }
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) { ptx::tc_fence_after_sync(); ptx::tmem_dealloc(tmem, 256); }
}

typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

extern "C" int exp_mix(void* mat, void* act, void* wgt, int total_rows, int n_mma, int b_rows, int stages, int iters, int mode, int grid, long long* out_dev) {
  cudaDriverEntryPointQueryResult q; void* f = nullptr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || !f) return -1;
  EncodeTiledFn enc = (EncodeTiledFn)f;
  CUtensorMap tmA, tmB;
  cuuint64_t dims[2] = {64, (cuuint64_t)total_rows};
  cuuint64_t st[1] = {128};
  cuuint32_t es[2] = {1, 1};
  cuuint32_t boxA[2] = {64, 128}, boxB[2] = {64, (cuuint32_t)b_rows};
  if (enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, mat, dims, st, boxA, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return -2;
  if (enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, mat, dims, st, boxB, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return -2;
  CUtensorMap tmI, tmT, tmW;
  {
    void* f2 = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &f2, cudaEnableDefault, &q) != cudaSuccess || !f2) return -1;
    EncodeIm2colFn enci = (EncodeIm2colFn)f2;
    cuuint64_t d4[4] = {192, 24, 18, 64};
    cuuint64_t s4[3] = {192 * 2, 24 * 192 * 2, 18 * 24 * 192 * 2};
    int lower[2] = {-1, -1}, upper[2] = {-1, -1};
    cuuint32_t e4[4] = {1, 1, 1, 1};
    if (enci(&tmI, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, act, d4, s4, lower, upper, 64, 128, e4, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return -6;
    cuuint32_t box4[4] = {64, 24, 5, 1};
    if (enc(&tmT, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, act, d4, s4, box4, e4, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return -7;
    cuuint64_t dw[2] = {1728, 192};
    cuuint64_t sw[1] = {1728 * 2};
    cuuint32_t boxw[2] = {64, (cuuint32_t)b_rows};
    if (enc(&tmW, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, wgt, dw, sw, boxw, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return -8;
  }
  const int b_bytes = (b_rows * 128 + 1023) / 1024 * 1024, nb_bytes = (n_mma * 128 + 1023) / 1024 * 1024;
  const int smem = 1024 + stages * (16384 + (b_bytes > nb_bytes ? b_bytes : nb_bytes));
  if (smem > 227 * 1024) return -5;
  if (cudaFuncSetAttribute(mix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -3;
  mix_kernel<<<grid, 384, smem>>>(tmA, tmB, tmI, tmT, tmW, n_mma, b_rows, stages, iters, mode, total_rows, out_dev);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { fprintf(stderr, "exp_mix: %s\n", cudaGetErrorString(e)); return -4; }
  return 0;
}
