// Hardware experiment: tcgen05.mma (cta_group::1, kind::f16, M=128, K=16, SS operands, SWIZZLE_128B) back-to-back
// issue rate vs N, operands resident in shared memory (no TMA traffic); A may start at any 128-byte row of its
// swizzle atom (`a_row_off`: the halo-patch kernel's tap shifts) with 8-row groups `a_stride_rows` rows apart.  One elected lane of a converged warp issues.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../simple-hrnet_b200/csrc/ptx.cuh"

__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int N, int iters, int a_stride_rows, int a_row_off, int d_col, int taps, int a_sw, long long* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (ptx::smem_u32(raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int warp = ptx::warp_idx_uniform();
  for (int i = threadIdx.x; i < (120 * 1024) / 4; i += 128) reinterpret_cast<uint32_t*>(raw + (base - ptx::smem_u32(raw)))[i] = 0;
  if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(&bar), 1); ptx::fence_mbar_init(); }
  if (warp == 0) ptx::tmem_alloc(ptx::smem_u32(&tslot), 256);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem = tslot;
  if (warp == 0) {
    const uint32_t idesc = ptx::umma_idesc_f16(128, N);
    const uint64_t adesc = ptx::umma_desc_kmajor(base + (uint32_t)a_row_off * (uint32_t)a_sw, (uint32_t)a_sw, (uint32_t)a_stride_rows * (uint32_t)a_sw);
    const uint64_t bdesc = ptx::umma_desc_kmajor(base + 32768u, 128u, 1024u);
    long long t0 = clock64();
    if (ptx::elect_one()) {
      if (!taps) {
        for (int i = 0; i < iters; ++i) ptx::mma_f16_ss(tmem + (uint32_t)d_col, adesc + (uint64_t)(2 * (i & (a_sw == 128 ? 3 : 1))), bdesc + (uint64_t)(2 * (i & 1)), idesc, 1u);
      } else {
        // the halo-patch kernel's sequence: 9 taps x 3 K16 steps, A shifted by (r * 10 + s) rows, B block t of 6 KB
        for (int i = 0; i < iters; i += 27)
#pragma unroll
          for (int t = 0; t < 9; ++t)
#pragma unroll
            for (int k = 0; k < 3; ++k)
              ptx::mma_f16_ss(tmem + (uint32_t)d_col, adesc + (uint64_t)(((t / 3) * 10 + (t % 3)) * 8 + 2 * k),
                              bdesc + (uint64_t)(t * (6144 >> 4) + 2 * k), idesc, (t | k) != 0 ? 1u : 0u);
      }
      ptx::mma_commit(ptx::smem_u32(&bar));
    }
    __syncwarp();
    long long t1 = clock64();
    ptx::mbar_wait(ptx::smem_u32(&bar), 0);
    long long t2 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after_sync(); ptx::tmem_dealloc(tmem, 256); }
}

extern "C" int exp_mma_rate(int N, int iters, int a_stride_rows, int a_row_off, int d_col, int taps, int a_sw, long long* out_dev) {
  const int smem = 1024 + 128 * 1024;
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  mma_rate_kernel<<<1, 128, smem>>>(N, iters, a_stride_rows, a_row_off, d_col, taps, a_sw, out_dev);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { fprintf(stderr, "exp_mma_rate: %s\n", cudaGetErrorString(e)); return -4; }
  return 0;
}
