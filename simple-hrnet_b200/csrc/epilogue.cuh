// Shared epilogue of the tcgen05 conv kernels: TMEM accumulator row -> BN scale/bias (fp32) -> + residual ->
// ReLU -> fp16 / fp32 NHWC store.  One thread owns one accumulator row (= one output pixel).
//
// The residual does not depend on the MMAs, so it is fetched into registers BEFORE the thread blocks on the
// accumulator barrier (and, for rows wider than 128 channels, one 128-channel group ahead): with only four
// epilogue warps per SM nothing else hides the ~1 us global-load latency (the first version issued the loads after
// tcgen05.wait::ld and was latency bound at ~8k clk per 128x48 tile, profiles/r01_bench_v2_patch.json).
#pragma once
#include "ptx.cuh"

namespace hrnet {

struct EpiRow {
  const float* s_scale;     // shared memory, indexed by absolute output channel
  const float* s_bias;
  const __half* residual;   // global NHWC fp16 or nullptr
  void* out;                // global NHWC fp16 / fp32
  size_t row_off;           // element offset of (this pixel, first channel of this tile)
  int ch0;                  // absolute first output channel of this tile (scale / bias index)
  int ncols;                // channels in this tile (multiple of 16)
  int relu, out_f32;
  bool valid;               // row maps to a real output pixel
};

// load residual channels [c_begin, c_begin + 128) of this row into r[0..15] (8 halves each)
__device__ __forceinline__ void epi_load_residual(uint4 (&r)[16], const EpiRow& e, int c_begin) {
  if (e.residual == nullptr || !e.valid) return;
  const uint4* rp = reinterpret_cast<const uint4*>(e.residual + e.row_off + c_begin);
#pragma unroll
  for (int i = 0; i < 16; ++i)
    if (c_begin + 8 * i < e.ncols) r[i] = __ldg(rp + i);
}

// `r` must hold the residual of channels [0, 128) on entry (epi_load_residual(r, e, 0) issued before the wait on
// the accumulator barrier).  t_row = TMEM address of (this warp's lane quarter, first column of the accumulator).
__device__ __forceinline__ void epi_store_row(uint4 (&r)[16], const EpiRow& e, uint32_t t_row) {
  for (int sc = 0; sc < e.ncols; sc += 128) {
    uint4 cur[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) cur[i] = r[i];
    if (sc + 128 < e.ncols) epi_load_residual(r, e, sc + 128);
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const int c = sc + 16 * g;
      if (c < e.ncols) {   // warp-uniform
        uint32_t v[16];
        ptx::tmem_ld16(t_row + (uint32_t)c, v);
        ptx::tmem_ld_wait();
        if (e.valid) {
          float y[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) y[i] = __uint_as_float(v[i]) * e.s_scale[e.ch0 + c + i] + e.s_bias[e.ch0 + c + i];
          if (e.residual != nullptr) {
            const __half2* h0 = reinterpret_cast<const __half2*>(&cur[2 * g]);
            const __half2* h1 = reinterpret_cast<const __half2*>(&cur[2 * g + 1]);
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
      }
    }
  }
}

}  // namespace hrnet
