// Shared epilogue of the tcgen05 conv kernels: TMEM accumulator row -> BN scale/bias (fp32) -> + residual ->
// ReLU -> fp16 / fp32 NHWC store.  One thread owns one accumulator row (= one output pixel); two epilogue
// warpgroups alternate tiles (warpgroup g always drains accumulator buffer g).
//
// Latency hiding (nothing else runs on these warps):
//   * the residual does not depend on the MMAs, so its first 64 channels are fetched into registers BEFORE the
//     thread blocks on the accumulator barrier, and each further 64-channel group one group ahead;
//   * TMEM is read 32 columns per tcgen05.wait::ld; scale / bias come from shared memory as float4.
#pragma once
#include "ptx.cuh"

namespace hrnet {

struct EpiRow {
  const float* s_scale;     // shared memory, indexed by absolute output channel (16-byte aligned base)
  const float* s_bias;
  const __half* residual;   // global NHWC fp16 or nullptr
  void* out;                // global NHWC fp16 / fp32
  size_t row_off;           // element offset of (this pixel, first channel of this tile)
  int ch0;                  // absolute first output channel of this tile (scale / bias index), multiple of 16
  int ncols;                // channels in this tile (multiple of 16)
  int relu, out_f32;
  bool valid;               // row maps to a real output pixel
};

// residual channels [c_begin, c_begin + 64) of this row -> r[0..7] (8 halves each)
__device__ __forceinline__ void epi_load_residual(uint4 (&r)[8], const EpiRow& e, int c_begin) {
  if (e.residual == nullptr || !e.valid) return;
  const uint4* rp = reinterpret_cast<const uint4*>(e.residual + e.row_off + c_begin);
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (c_begin + 8 * i < e.ncols) r[i] = __ldg(rp + i);
}

// 16 accumulator columns starting at tile column c: BN, residual (r0 | r1 = 16 halves), ReLU, store
__device__ __forceinline__ void epi_cols16(const uint32_t (&v)[16], const uint4& r0, const uint4& r1, const EpiRow& e,
                                           int c) {
  float y[16];
  const float4* sc = reinterpret_cast<const float4*>(e.s_scale + e.ch0 + c);
  const float4* bi = reinterpret_cast<const float4*>(e.s_bias + e.ch0 + c);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 s4 = sc[i], b4 = bi[i];
    y[4 * i + 0] = fmaf(__uint_as_float(v[4 * i + 0]), s4.x, b4.x);
    y[4 * i + 1] = fmaf(__uint_as_float(v[4 * i + 1]), s4.y, b4.y);
    y[4 * i + 2] = fmaf(__uint_as_float(v[4 * i + 2]), s4.z, b4.z);
    y[4 * i + 3] = fmaf(__uint_as_float(v[4 * i + 3]), s4.w, b4.w);
  }
  if (e.residual != nullptr) {
    const __half2* h0 = reinterpret_cast<const __half2*>(&r0);
    const __half2* h1 = reinterpret_cast<const __half2*>(&r1);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f0 = __half22float2(h0[i]), f1 = __half22float2(h1[i]);
      y[2 * i] += f0.x; y[2 * i + 1] += f0.y;
      y[8 + 2 * i] += f1.x; y[8 + 2 * i + 1] += f1.y;
    }
  }
  if (e.relu) {
#pragma unroll
    for (int i = 0; i < 16; ++i) y[i] = fmaxf(y[i], 0.f);
  }
  if (e.out_f32) {
    float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(e.out) + e.row_off + c);
#pragma unroll
    for (int i = 0; i < 4; ++i) op[i] = make_float4(y[4 * i], y[4 * i + 1], y[4 * i + 2], y[4 * i + 3]);
  } else {
    uint4 o[2];
    __half2* oh2 = reinterpret_cast<__half2*>(o);
#pragma unroll
    for (int i = 0; i < 8; ++i) oh2[i] = __floats2half2_rn(y[2 * i], y[2 * i + 1]);
    uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(e.out) + e.row_off + c);
    op[0] = o[0];
    op[1] = o[1];
  }
}

// `r` must hold the residual of channels [0, 64) on entry (epi_load_residual(r, e, 0) issued before the wait on the
// accumulator barrier).  t_row = TMEM address of (this warp's lane quarter, first column of the accumulator).
// Note: the y = acc * scale + bias rounding (one fma) differs from a separate mul + add by <= 1 ulp of fp32.
__device__ __forceinline__ void epi_store_row(uint4 (&r)[8], const EpiRow& e, uint32_t t_row) {
  for (int c64 = 0; c64 < e.ncols; c64 += 64) {
    uint4 cur[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) cur[i] = r[i];
    if (c64 + 64 < e.ncols) epi_load_residual(r, e, c64 + 64);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = c64 + 32 * h;
      if (c < e.ncols) {                       // warp-uniform
        uint32_t v0[16], v1[16];
        const bool two = c + 16 < e.ncols;     // warp-uniform
        ptx::tmem_ld16(t_row + (uint32_t)c, v0);
        if (two) ptx::tmem_ld16(t_row + (uint32_t)(c + 16), v1);
        ptx::tmem_ld_wait();
        if (e.valid) {
          epi_cols16(v0, cur[4 * h], cur[4 * h + 1], e, c);
          if (two) epi_cols16(v1, cur[4 * h + 2], cur[4 * h + 3], e, c + 16);
        }
      }
    }
  }
}

}  // namespace hrnet
