// Hardware experiment: cost of the per-stage hand-off around tcgen05.mma.  G MMAs (M=128, K=16, N) per group, one
// tcgen05.commit per group onto a ring of mbarriers, R groups; optionally each group reads a different shared-memory
// stage (ring of `nstages` buffers of A 16 KB + B N*128 B) and optionally a second warp streams bulk copies
// (cp.async.bulk global -> shared, 16 KB each) into a separate buffer while the MMAs run.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../simple-hrnet_b200/csrc/ptx.cuh"

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__global__ void __launch_bounds__(128, 1) mma_commit_kernel(int N, int G, int R, int nstages, int commit_every, int copy_kb,
                                                            const uint8_t* gsrc, long long* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (ptx::smem_u32(raw) + 1023u) & ~1023u;
  __shared__ uint64_t bars[8];
  __shared__ uint64_t cbar;
  __shared__ uint32_t tslot;
  __shared__ volatile int stop;
  const int warp = ptx::warp_idx_uniform();
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) ptx::mbar_init(ptx::smem_u32(&bars[i]), 1);
    ptx::mbar_init(ptx::smem_u32(&cbar), 1);
    ptx::fence_mbar_init();
    stop = 0;
  }
  if (warp == 0) ptx::tmem_alloc(ptx::smem_u32(&tslot), 256);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem = tslot;
  const uint32_t stage_bytes = 16384u + (uint32_t)N * 128u;
  const uint32_t copy_dst = base + (uint32_t)nstages * stage_bytes;   // 16 KB landing buffer for the streaming copies
  if (warp == 0) {
    const uint32_t idesc = ptx::umma_idesc_f16(128, N);
    long long t0 = clock64();
    for (int r = 0; r < R; ++r) {
      const uint32_t a = base + (uint32_t)(r % nstages) * stage_bytes;
      const uint64_t adesc = ptx::umma_desc_kmajor(a, 128u, 1024u);
      const uint64_t bdesc = ptx::umma_desc_kmajor(a + 16384u, 128u, 1024u);
      if (ptx::elect_one()) {
        for (int i = 0; i < G; ++i)
          ptx::mma_f16_ss(tmem, adesc + (uint64_t)(2 * (i & 3)), bdesc + (uint64_t)(2 * (i & 3)), idesc, 1u);
        if (commit_every && (r % commit_every) == commit_every - 1) ptx::mma_commit(ptx::smem_u32(&bars[r & 7]));
      }
      __syncwarp();
    }
    if (ptx::elect_one()) ptx::mma_commit(ptx::smem_u32(&cbar));
    __syncwarp();
    long long t1 = clock64();
    ptx::mbar_wait(ptx::smem_u32(&cbar), 0);
    long long t2 = clock64();
    stop = 1;
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  } else if (warp == 1 && copy_kb > 0) {
    __shared__ uint64_t lbar;
    if (lane == 0) { ptx::mbar_init(ptx::smem_u32(&lbar), 1); ptx::fence_mbar_init(); }
    __syncwarp();
    uint32_t ph = 0;
    long long n = 0;
    while (!stop) {
      if (lane == 0) {
        ptx::mbar_expect_tx(ptx::smem_u32(&lbar), (uint32_t)copy_kb * 1024u);
        bulk_g2s(copy_dst, gsrc + ((n * copy_kb * 1024) & ((8 << 20) - 1)), (uint32_t)copy_kb * 1024u, ptx::smem_u32(&lbar));
      }
      __syncwarp();
      ptx::mbar_wait(ptx::smem_u32(&lbar), ph);
      ph ^= 1u; ++n;
    }
    if (lane == 0) out[2] = n;
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after_sync(); ptx::tmem_dealloc(tmem, 256); }
}

extern "C" int exp_mma_commit(int N, int G, int R, int nstages, int commit_every, int copy_kb, const void* gsrc, long long* out_dev) {
  const int smem = 1024 + nstages * (16384 + N * 128) + 32 * 1024;
  cudaFuncSetAttribute(mma_commit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  mma_commit_kernel<<<1, 128, smem>>>(N, G, R, nstages, commit_every, copy_kb, (const uint8_t*)gsrc, out_dev);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { fprintf(stderr, "exp_mma_commit: %s\n", cudaGetErrorString(e)); return -4; }
  return 0;
}
