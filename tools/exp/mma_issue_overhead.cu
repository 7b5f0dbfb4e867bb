// Hardware experiment: where do the ~550 clk per group of 4 tcgen05.mma (+ ~210 per tcgen05.commit) go?
// (profiles/r01_exp_mma_commit.log: a group of G MMAs costs ~360 clk + G x MMA time, whatever N.)
// Variants of the issuing structure, M=128, N=64, K=16, G MMAs per group, R groups, ring of 4 smem stages:
//   0: per group { descriptors; if (elect_one()) { G x mma; commit }; __syncwarp(); }            (kernel structure)
//   1: as 0 without the commit
//   2: as 0, descriptors from a precomputed table in shared memory (no address arithmetic in the loop)
//   3: ONE elect region around the whole loop: if (elect_one()) for r { descriptors; G x mma; commit }
//   4: as 3 without the commit
//   5: as 3 + an mbarrier try_wait (already completed phase) per group, as the real MMA warp does
//   6: as 0 + the same mbarrier wait per group, executed by the whole warp (kernel structure)
//   7: as 3 with plain `if (lane == 0)` instead of elect_one()
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../simple-hrnet_b200/csrc/ptx.cuh"

template <int V>
__global__ void __launch_bounds__(128, 1) k(int N, int G, int R, long long* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (ptx::smem_u32(raw) + 1023u) & ~1023u;
  __shared__ uint64_t bars[8];
  __shared__ uint64_t ready;       // completed once before the loop: waits on phase 0 return immediately
  __shared__ uint64_t cbar;
  __shared__ uint32_t tslot;
  __shared__ uint64_t dtab[8];
  const int warp = ptx::warp_idx_uniform();
  const int lane = threadIdx.x & 31;
  const uint32_t stage_bytes = 16384u + (uint32_t)N * 128u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) ptx::mbar_init(ptx::smem_u32(&bars[i]), 1);
    ptx::mbar_init(ptx::smem_u32(&cbar), 1);
    ptx::mbar_init(ptx::smem_u32(&ready), 1);
    ptx::fence_mbar_init();
    ptx::mbar_arrive(ptx::smem_u32(&ready));
    for (int i = 0; i < 4; ++i) {
      dtab[2 * i] = ptx::umma_desc_kmajor(base + (uint32_t)i * stage_bytes, 128u, 1024u);
      dtab[2 * i + 1] = ptx::umma_desc_kmajor(base + (uint32_t)i * stage_bytes + 16384u, 128u, 1024u);
    }
  }
  if (warp == 0) ptx::tmem_alloc(ptx::smem_u32(&tslot), 256);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem = tslot;
  if (warp == 0) {
    const uint32_t idesc = ptx::umma_idesc_f16(128, N);
    const uint32_t rdy = ptx::smem_u32(&ready);
    long long t0 = clock64();
    if constexpr (V == 0 || V == 1 || V == 2 || V == 6) {
      for (int r = 0; r < R; ++r) {
        if constexpr (V == 6) ptx::mbar_wait(rdy, 0);
        uint64_t adesc, bdesc;
        if constexpr (V == 2) { adesc = dtab[2 * (r & 3)]; bdesc = dtab[2 * (r & 3) + 1]; }
        else {
          const uint32_t a = base + (uint32_t)(r & 3) * stage_bytes;
          adesc = ptx::umma_desc_kmajor(a, 128u, 1024u);
          bdesc = ptx::umma_desc_kmajor(a + 16384u, 128u, 1024u);
        }
        if (ptx::elect_one()) {
#pragma unroll 4
          for (int i = 0; i < G; ++i)
            ptx::mma_f16_ss(tmem, adesc + (uint64_t)(2 * (i & 3)), bdesc + (uint64_t)(2 * (i & 3)), idesc, 1u);
          if constexpr (V != 1) ptx::mma_commit(ptx::smem_u32(&bars[r & 7]));
        }
        __syncwarp();
      }
    } else {
      bool me;
      if constexpr (V == 7) me = lane == 0; else me = ptx::elect_one();
      if (me) {
        for (int r = 0; r < R; ++r) {
          if constexpr (V == 5) ptx::mbar_wait(rdy, 0);
          const uint32_t a = base + (uint32_t)(r & 3) * stage_bytes;
          const uint64_t adesc = ptx::umma_desc_kmajor(a, 128u, 1024u);
          const uint64_t bdesc = ptx::umma_desc_kmajor(a + 16384u, 128u, 1024u);
#pragma unroll 4
          for (int i = 0; i < G; ++i)
            ptx::mma_f16_ss(tmem, adesc + (uint64_t)(2 * (i & 3)), bdesc + (uint64_t)(2 * (i & 3)), idesc, 1u);
          if constexpr (V != 4) ptx::mma_commit(ptx::smem_u32(&bars[r & 7]));
        }
      }
      __syncwarp();
    }
    if (ptx::elect_one()) ptx::mma_commit(ptx::smem_u32(&cbar));
    __syncwarp();
    long long t1 = clock64();
    ptx::mbar_wait(ptx::smem_u32(&cbar), 0);
    long long t2 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after_sync(); ptx::tmem_dealloc(tmem, 256); }
}

template <int V>
static int run(int N, int G, int R, long long* out_dev) {
  const int smem = 1024 + 4 * (16384 + N * 128);
  cudaFuncSetAttribute(k<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<V><<<1, 128, smem>>>(N, G, R, out_dev);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { fprintf(stderr, "exp_mma_issue_overhead v%d: %s\n", V, cudaGetErrorString(e)); return -4; }
  return 0;
}
extern "C" int exp_mma_issue_overhead(int variant, int N, int G, int R, long long* out_dev) {
  switch (variant) {
    case 0: return run<0>(N, G, R, out_dev);
    case 1: return run<1>(N, G, R, out_dev);
    case 2: return run<2>(N, G, R, out_dev);
    case 3: return run<3>(N, G, R, out_dev);
    case 4: return run<4>(N, G, R, out_dev);
    case 5: return run<5>(N, G, R, out_dev);
    case 6: return run<6>(N, G, R, out_dev);
    case 7: return run<7>(N, G, R, out_dev);
  }
  return -1;
}
