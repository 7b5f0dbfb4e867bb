// Hardware experiment (not product code): what does the thread-per-row epilogue of a narrow tile cost, piece by piece?
// 8 epilogue warps (two warpgroups of 128 accumulator rows, like the product kernels) loop over `tiles` tiles of a
// [rows][C] fp16 NHWC tensor; no MMAs run, the accumulators are whatever TMEM holds.  `mode` switches the pieces on:
//   bit 0  tcgen05.ld of the tile's C columns, two x16 loads per tcgen05.wait::ld (the product epilogue's pattern)
//   bit 1  (with bit 0) all x16 loads of the tile first, ONE wait
//   bit 2  BN fma + ReLU + fp16 conversion of the loaded columns
//   bit 3  thread-per-row global stores (C * 2 bytes per row as 16 B stores)
//   bit 4  thread-per-row residual loads (__ldg, issued one tile ahead like the product), added to the result
//   bit 5  only warpgroup 0 works (warpgroup 1 idles): contention between the two warpgroups = difference
//   bit 6  rows of a tile are 8-pixel segments of 4 image rows (halo-patch geometry) instead of 128 consecutive pixels
// Output: average clk per tile per warpgroup.  One CTA per SM on `grid` SMs, every CTA walks its own tiles.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../simple-hrnet_b200/csrc/ptx.cuh"

template <int C>
__global__ void __launch_bounds__(384, 1)
epilogue_cost_kernel(const __half* __restrict__ res, __half* __restrict__ out, int tiles_total, int W, int mode,
                     long long* cycles) {
  __shared__ uint32_t tslot;
  __shared__ float s_scale[256], s_bias[256];
  const int warp = ptx::warp_idx_uniform();
  const int lane = threadIdx.x & 31;
  if (warp == 2) ptx::tmem_alloc(ptx::smem_u32(&tslot), 512);
  for (int i = threadIdx.x; i < 256; i += 384) { s_scale[i] = 1.0f + 0.001f * i; s_bias[i] = 0.01f * i; }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem = tslot;
  if (warp >= 4) {
    const int g = (warp - 4) >> 2, q = warp & 3, row = q * 32 + lane;
    const bool active = !((mode & 32) && g == 1);
    const uint32_t t_row = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * 256);
    long long t0 = clock64();
    int done = 0;
    if (active) {
      uint4 rnext[C / 8];
      auto row_off = [&](int tile) -> size_t {
        if (mode & 64) {   // halo-patch geometry: tile = 8 (w) x 16 (h) pixels of a W-pixel-wide map
          const int tw = tile % (W / 8), th = tile / (W / 8);
          return ((size_t)(th * 16 + (row >> 3)) * W + (size_t)(tw * 8 + (row & 7))) * C;
        }
        return ((size_t)tile * 128 + row) * C;
      };
      int tile = blockIdx.x * 2 + g;
      if ((mode & 16) && tile < tiles_total) {
        const uint4* rp = reinterpret_cast<const uint4*>(res + row_off(tile));
#pragma unroll
        for (int i = 0; i < C / 8; ++i) rnext[i] = __ldg(rp + i);
      }
      for (; tile < tiles_total; tile += gridDim.x * 2, ++done) {
        uint4 rcur[C / 8];
#pragma unroll
        for (int i = 0; i < C / 8; ++i) rcur[i] = rnext[i];
        const int ntile = tile + gridDim.x * 2;
        if ((mode & 16) && ntile < tiles_total) {
          const uint4* rp = reinterpret_cast<const uint4*>(res + row_off(ntile));
#pragma unroll
          for (int i = 0; i < C / 8; ++i) rnext[i] = __ldg(rp + i);
        }
        uint32_t v[C];
#pragma unroll
        for (int i = 0; i < C; ++i) v[i] = 0x3f800000u + (uint32_t)(i + row);
        if (mode & 1) {
          if (mode & 2) {
#pragma unroll
            for (int c = 0; c < C; c += 16) ptx::tmem_ld16(t_row + (uint32_t)c, reinterpret_cast<uint32_t(&)[16]>(v[c]));
            ptx::tmem_ld_wait();
          } else {
#pragma unroll
            for (int c = 0; c < C; c += 32) {
              ptx::tmem_ld16(t_row + (uint32_t)c, reinterpret_cast<uint32_t(&)[16]>(v[c]));
              if (c + 16 < C) ptx::tmem_ld16(t_row + (uint32_t)(c + 16), reinterpret_cast<uint32_t(&)[16]>(v[c + 16]));
              ptx::tmem_ld_wait();
            }
          }
        }
        uint4 o[C / 8];
        __half2* oh = reinterpret_cast<__half2*>(o);
        const __half2* rh = reinterpret_cast<const __half2*>(rcur);
#pragma unroll
        for (int i = 0; i < C / 2; ++i) {
          float a = __uint_as_float(v[2 * i]), b = __uint_as_float(v[2 * i + 1]);
          if (mode & 4) {
            a = fmaxf(fmaf(a, s_scale[2 * i], s_bias[2 * i]), 0.f);
            b = fmaxf(fmaf(b, s_scale[2 * i + 1], s_bias[2 * i + 1]), 0.f);
          }
          if (mode & 16) { const float2 f = __half22float2(rh[i]); a += f.x; b += f.y; }
          oh[i] = __floats2half2_rn(a, b);
        }
        if (mode & 8) {
          uint4* op = reinterpret_cast<uint4*>(out + row_off(tile));
#pragma unroll
          for (int i = 0; i < C / 8; ++i) op[i] = o[i];
        } else {   // keep the work alive without global traffic
          uint32_t x = 0;
#pragma unroll
          for (int i = 0; i < C / 8; ++i) x ^= o[i].x ^ o[i].y ^ o[i].z ^ o[i].w;
          if (x == 0x12345678u) out[0] = __float2half(1.f);
        }
      }
    }
    const long long t1 = clock64();
    if (lane == 0 && q == 0) { cycles[(blockIdx.x * 2 + g) * 2] = t1 - t0; cycles[(blockIdx.x * 2 + g) * 2 + 1] = done; }
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) { ptx::tc_fence_after_sync(); ptx::tmem_dealloc(tmem, 512); }
}

extern "C" int exp_epilogue_cost(int C, const void* res, void* out, int tiles_total, int W, int mode, int grid,
                                 long long* cycles_dev) {
  switch (C) {
    case 48: epilogue_cost_kernel<48><<<grid, 384>>>((const __half*)res, (__half*)out, tiles_total, W, mode, cycles_dev); break;
    case 64: epilogue_cost_kernel<64><<<grid, 384>>>((const __half*)res, (__half*)out, tiles_total, W, mode, cycles_dev); break;
    case 96: epilogue_cost_kernel<96><<<grid, 384>>>((const __half*)res, (__half*)out, tiles_total, W, mode, cycles_dev); break;
    default: return -1;
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { fprintf(stderr, "exp_epilogue_cost: %s\n", cudaGetErrorString(e)); return -4; }
  return 0;
}
