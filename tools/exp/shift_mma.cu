// Hardware experiment (not product code): can a tcgen05 K-major SWIZZLE_128B shared-memory descriptor
// start at an address that is 128B- but not 1024B-aligned (a "row-shifted view" of a TMA-written tile),
// with a stride-byte-offset that is not 1024?  Needed for halo-patch reuse of the im2col operand.
//
//   X [R=256][64] fp16 --TMA SW128--> smem rows at 128 B pitch
//   B = I (64x64) --TMA SW128--> smem
//   D[m][n] = sum_k A[m][k] B[n][k], A = view(start = base + shift*128, SBO = sbo_bytes)
// expected: D[m][:] == X[(m/8)*(sbo/128) + (m%8) + shift][:]
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../simple-hrnet_b200/csrc/ptx.cuh"

__global__ void __launch_bounds__(128, 1)
shift_mma_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmB, float* D,
                 int shift, int sbo_bytes, int base_off_mode, int sw_bytes) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (ptx::smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* al = raw + (base - ptx::smem_u32(raw));
  const int row_bytes = sw_bytes;                       // K chunk = sw_bytes/2 fp16
  const uint32_t xs = base;                             // 256 rows
  const uint32_t bs = base + 256 * row_bytes;           // 64 rows
  uint64_t* bar = reinterpret_cast<uint64_t*>(al + 256 * row_bytes + 64 * row_bytes);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    ptx::mbar_init(ptx::smem_u32(&bar[0]), 1);
    ptx::mbar_init(ptx::smem_u32(&bar[1]), 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) ptx::tmem_alloc(ptx::smem_u32(tslot), 64);
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem = *tslot;
  if (threadIdx.x == 0) {
    ptx::mbar_expect_tx(ptx::smem_u32(&bar[0]), (uint32_t)(320 * row_bytes));
    ptx::tma_load_2d(xs, &tmX, ptx::smem_u32(&bar[0]), 0, 0);
    ptx::tma_load_2d(bs, &tmB, ptx::smem_u32(&bar[0]), 0, 0);
    ptx::mbar_wait(ptx::smem_u32(&bar[0]), 0);
    ptx::tc_fence_after_sync();
    const uint32_t a_addr = xs + (uint32_t)(shift * row_bytes);
    uint64_t adesc = ptx::umma_desc_kmajor(a_addr, (uint32_t)sw_bytes, (uint32_t)sbo_bytes);
    if (base_off_mode) adesc |= (uint64_t)((a_addr >> 7) & 7u) << 49;
    const uint64_t bdesc = ptx::umma_desc_kmajor(bs, (uint32_t)sw_bytes, 8u * sw_bytes);
    const uint32_t idesc = ptx::umma_idesc_f16(128, 64);
    const int nk = sw_bytes / 32;
    for (int k = 0; k < nk; ++k) ptx::mma_f16_ss(tmem, adesc + 2 * k, bdesc + 2 * k, idesc, k > 0);
    ptx::mma_commit(ptx::smem_u32(&bar[1]));
  }
  ptx::mbar_wait(ptx::smem_u32(&bar[1]), 0);
  ptx::tc_fence_after_sync();
  const int row = warp * 32 + lane;
  for (int c = 0; c < 64; c += 16) {
    uint32_t v[16];
    ptx::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
    ptx::tmem_ld_wait();
    for (int i = 0; i < 16; ++i) D[row * 64 + c + i] = __uint_as_float(v[i]);
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after_sync(); ptx::tmem_dealloc(tmem, 64); }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// X: [256][kc] fp16, Bm: [64][kc] fp16 (kc = sw_bytes/2), D: [128][64] f32 (device pointers)
extern "C" int exp_shift_mma(void* X, void* Bm, float* D, int shift, int sbo_bytes, int base_off_mode, int sw_bytes) {
  cudaDriverEntryPointQueryResult q; void* f = nullptr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || !f) return -1;
  EncodeTiledFn enc = (EncodeTiledFn)f;
  const int kc = sw_bytes / 2;
  CUtensorMapSwizzle sw = sw_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (sw_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUtensorMap tx, tb;
  cuuint32_t es[2] = {1, 1};
  { cuuint64_t dims[2] = {(cuuint64_t)kc, 256}; cuuint64_t st[1] = {(cuuint64_t)kc * 2}; cuuint32_t box[2] = {(cuuint32_t)kc, 256};
    if (enc(&tx, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, X, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return -2; }
  { cuuint64_t dims[2] = {(cuuint64_t)kc, 64}; cuuint64_t st[1] = {(cuuint64_t)kc * 2}; cuuint32_t box[2] = {(cuuint32_t)kc, 64};
    if (enc(&tb, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, Bm, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return -3; }
  const int smem = 1024 + 320 * sw_bytes + 64;
  cudaFuncSetAttribute(shift_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  shift_mma_kernel<<<1, 128, smem>>>(tx, tb, D, shift, sbo_bytes, base_off_mode, sw_bytes);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { fprintf(stderr, "exp_shift_mma: %s\n", cudaGetErrorString(e)); return -4; }
  return 0;
}
