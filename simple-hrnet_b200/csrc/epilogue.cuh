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

// kEpi == 3 (opt-in, HRNET_TUNE_EPILOGUE = batch; in the GPU test matrix, no gain measured): the same thread-per-row epilogue for tiles at most 64
// channels wide with all tcgen05.ld of the tile issued before ONE tcgen05.wait::ld (epi_store_row waits once per 32
// columns: two TMEM round trips for a 48-channel tile).  `r` holds the residual of channels [0, 64) like above.
__device__ __forceinline__ void epi_store_row_batched(uint4 (&r)[8], const EpiRow& e, uint32_t t_row) {
  uint32_t v[4][16];
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (16 * k < e.ncols) ptx::tmem_ld16(t_row + (uint32_t)(16 * k), v[k]);     // warp-uniform
  ptx::tmem_ld_wait();
  if (e.valid) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (16 * k < e.ncols) epi_cols16(v[k], r[2 * k], r[2 * k + 1], e, 16 * k);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Shared-memory-staged epilogue with TMA stores (fp16 output, no sub-pixel remap).
//
// The thread-per-row stores above touch 32 different cache lines per warp instruction; for wide tiles (N = 256:
// layer1 conv3 / downsample) that, not HBM, bounds the kernel (19k clk of epilogue per 128x256 tile,
// profiles/r01_dbg_role_timers_other_layers.log).  Here each warpgroup owns two 16 KB staging tiles of
// [128 rows][64 channels] in the SWIZZLE_128B layout TMA expects:
//   residual: one elected thread TMA-loads the tile's 64-channel chunk (prefetched one chunk ahead), every thread then
//             reads its own row from shared memory;
//   output  : every thread writes its row's 64 channels, one elected thread TMA-stores the tile (rows / channels
//             outside the tensor are clipped by the descriptor, so no validity masks are needed).
struct EpiTma {
  const void* tm_out;        // 2-D {C, rows} box {64, 128}  or  4-D {C, W, H, N} box {64, 8, 16, 1}
  const void* tm_res;        // same geometry over the residual tensor (unused when has_res == 0)
  int dims4;                 // 0: 2-D coordinates (ch, row0); 1: 4-D coordinates (ch, w0, h0, img)
  int c_row0, c_w0, c_h0, c_img;
  int ch0, ncols;            // absolute first channel / channels of this tile
  int has_res, relu;
  bool store;                // false for ghost tiles (CTA-pair mode): compute but do not store
  const float* s_scale;
  const float* s_bias;
  uint32_t stage_out, stage_res;   // shared-memory addresses (1024 B aligned) of this warpgroup's staging tiles
  uint32_t res_bar;                // mbarrier for the residual loads of this warpgroup
  int bar_id;                      // named barrier id of this warpgroup (128 threads)
};

__device__ __forceinline__ void epi_tma_issue_residual(const EpiTma& e, int c64) {
  ptx::mbar_expect_tx(e.res_bar, 128u * 128u);
  if (e.dims4) ptx::tma_load_4d(e.stage_res, e.tm_res, e.res_bar, e.ch0 + c64, e.c_w0, e.c_h0, e.c_img);
  else ptx::tma_load_2d(e.stage_res, e.tm_res, e.res_bar, e.ch0 + c64, e.c_row0);
}

// row = this thread's accumulator row (0..127), leader = one fixed thread of the warpgroup.
// res_phase: running phase bit of e.res_bar (updated).  The first residual chunk must have been issued by the leader
// (epi_tma_issue_residual(e, 0)) before the wait on the accumulator barrier.
__device__ __forceinline__ void epi_tma_tile(const EpiTma& e, uint32_t t_row, int row, bool leader, uint32_t& res_phase) {
  const uint32_t my_out = e.stage_out + (uint32_t)row * 128u;
  const uint32_t my_res = e.stage_res + (uint32_t)row * 128u;
  const uint32_t sw = (uint32_t)(row & 7);
  for (int c64 = 0; c64 < e.ncols; c64 += 64) {
    const int nc = min(64, e.ncols - c64);     // 16, 32, 48 or 64 (warp-uniform)
    uint4 cur[8];
    if (e.has_res) {
      ptx::mbar_wait(e.res_bar, res_phase);
      res_phase ^= 1u;
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (8 * k < nc)
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(cur[k].x), "=r"(cur[k].y), "=r"(cur[k].z), "=r"(cur[k].w)
                       : "r"(my_res + (((uint32_t)k ^ sw) << 4)));
    }
    // The residual rows were read through the generic proxy and stage_res is about to be overwritten by the next TMA
    // load (async proxy): without this cross-proxy fence the overwrite is not ordered after the reads.  Observed on
    // B200 as run-to-run differences of layer1 conv3 inside the network for 3 <= batch < 64 (never in isolation);
    // CUTLASS's TMA epilogues fence the same way before releasing a TMA-loaded source buffer.
    if (e.has_res) ptx::fence_proxy_async_smem();
    if (leader) ptx::tma_store_wait_read();      // the previous chunk's store no longer reads stage_out
    ptx::bar_sync(e.bar_id, 128);                // residual tile consumed by everyone, stage_out free
    if (e.has_res && leader && c64 + 64 < e.ncols) epi_tma_issue_residual(e, c64 + 64);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = c64 + 32 * h;
      if (32 * h < nc) {                         // warp-uniform
        uint32_t v0[16], v1[16];
        const bool two = 32 * h + 16 < nc;
        ptx::tmem_ld16(t_row + (uint32_t)c, v0);
        if (two) ptx::tmem_ld16(t_row + (uint32_t)(c + 16), v1);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          if (half == 1 && !two) break;
          const uint32_t* v = half ? v1 : v0;
          const int cc = c + 16 * half;
          const float4* sc = reinterpret_cast<const float4*>(e.s_scale + e.ch0 + cc);
          const float4* bi = reinterpret_cast<const float4*>(e.s_bias + e.ch0 + cc);
          float y[16];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 s4 = sc[i], b4 = bi[i];
            y[4 * i + 0] = fmaf(__uint_as_float(v[4 * i + 0]), s4.x, b4.x);
            y[4 * i + 1] = fmaf(__uint_as_float(v[4 * i + 1]), s4.y, b4.y);
            y[4 * i + 2] = fmaf(__uint_as_float(v[4 * i + 2]), s4.z, b4.z);
            y[4 * i + 3] = fmaf(__uint_as_float(v[4 * i + 3]), s4.w, b4.w);
          }
          if (e.has_res) {
            const __half2* h0 = reinterpret_cast<const __half2*>(&cur[4 * h + 2 * half]);
            const __half2* h1 = reinterpret_cast<const __half2*>(&cur[4 * h + 2 * half + 1]);
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
          uint4 o[2];
          __half2* oh2 = reinterpret_cast<__half2*>(o);
#pragma unroll
          for (int i = 0; i < 8; ++i) oh2[i] = __floats2half2_rn(y[2 * i], y[2 * i + 1]);
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const uint32_t chunk = (uint32_t)(4 * h + 2 * half + k);
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};"
                         ::"r"(my_out + ((chunk ^ sw) << 4)), "r"(o[k].x), "r"(o[k].y), "r"(o[k].z), "r"(o[k].w) : "memory");
          }
        }
      }
    }
    ptx::fence_proxy_async_smem();               // generic-proxy writes -> visible to the TMA engine
    ptx::bar_sync(e.bar_id, 128);
    if (leader && e.store) {
      if (e.dims4) ptx::tma_store_4d(e.tm_out, e.stage_out, e.ch0 + c64, e.c_w0, e.c_h0, e.c_img);
      else ptx::tma_store_2d(e.tm_out, e.stage_out, e.ch0 + c64, e.c_row0);
      ptx::tma_store_commit();
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Per-warp TMA epilogue of the halo-patch kernel (fp16 output).  Same idea as above at warp granularity: every epilogue
// warp owns a 4 KB output and a 4 KB residual tile ([32 rows][64 channels], SWIZZLE_128B) and its own mbarrier, its 32
// accumulator rows are 4 rows x 8 pixels of the 8 x 16 output tile = one TMA box {64, 8, 4, 1}.  Only __syncwarp() and
// one elected lane's TMA instructions are needed -- no 128-thread barriers -- and the load/store unit sees 12 shared-
// memory accesses per thread instead of 12 scattered global ones (the thread-per-row stores cost ~1000 clk per tile,
// the residual loads ~500: more than a CTA pair's MMAs for the tile, profiles/r01_exp_patch_pair_nostore.log).
// elect.sync picks the same lane for the same mask every time, so the lane that commits a bulk group also waits on it.
struct EpiWarpTma {
  const void* tm_out;        // 4-D {C, W, H, N}, box {64, 8, 4, 1}
  const void* tm_res;
  int c_w0, c_h0, c_img;     // tile origin of THIS warp's 4 rows
  int ncols, has_res, relu;
  bool store;                // false for ghost tiles (CTA-pair mode)
  const float* s_scale;
  const float* s_bias;
  uint32_t stage_out, stage_res;   // 1024 B aligned shared-memory addresses of this warp's tiles
  uint32_t res_bar;
};

__device__ __forceinline__ void epi_wtma_issue_residual(const EpiWarpTma& e, int c64) {   // call from one elected lane
  ptx::mbar_expect_tx(e.res_bar, 32u * 128u);
  ptx::tma_load_4d(e.stage_res, e.tm_res, e.res_bar, c64, e.c_w0, e.c_h0, e.c_img);
}

// The residual of chunk 0 must have been requested before the wait on the accumulator barrier.
__device__ __forceinline__ void epi_wtma_tile(const EpiWarpTma& e, uint32_t t_row, int lane, uint32_t& res_phase) {
  const uint32_t my_out = e.stage_out + (uint32_t)lane * 128u;
  const uint32_t my_res = e.stage_res + (uint32_t)lane * 128u;
  const uint32_t sw = (uint32_t)(lane & 7);
  for (int c64 = 0; c64 < e.ncols; c64 += 64) {
    const int nc = min(64, e.ncols - c64);     // warp-uniform
    uint4 cur[8];
    if (e.has_res) {
      ptx::mbar_wait(e.res_bar, res_phase);
      res_phase ^= 1u;
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (8 * k < nc) cur[k] = ptx::lds128(my_res + (((uint32_t)k ^ sw) << 4));
    }
    if (e.has_res) ptx::fence_proxy_async_smem();       // generic-proxy reads of stage_res before its next TMA overwrite (see above)
    if (ptx::elect_one()) ptx::tma_store_wait_read();   // the previous store no longer reads stage_out
    __syncwarp();                                        // ... and every lane has read its residual row
    if (e.has_res && c64 + 64 < e.ncols) {
      if (ptx::elect_one()) epi_wtma_issue_residual(e, c64 + 64);
      __syncwarp();
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = c64 + 32 * h;
      if (32 * h < nc) {                         // warp-uniform
        uint32_t v0[16], v1[16];
        const bool two = 32 * h + 16 < nc;
        ptx::tmem_ld16(t_row + (uint32_t)c, v0);
        if (two) ptx::tmem_ld16(t_row + (uint32_t)(c + 16), v1);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          if (half == 1 && !two) break;
          const uint32_t* v = half ? v1 : v0;
          const int cc = c + 16 * half;
          const float4* sc = reinterpret_cast<const float4*>(e.s_scale + cc);
          const float4* bi = reinterpret_cast<const float4*>(e.s_bias + cc);
          float y[16];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 s4 = sc[i], b4 = bi[i];
            y[4 * i + 0] = fmaf(__uint_as_float(v[4 * i + 0]), s4.x, b4.x);
            y[4 * i + 1] = fmaf(__uint_as_float(v[4 * i + 1]), s4.y, b4.y);
            y[4 * i + 2] = fmaf(__uint_as_float(v[4 * i + 2]), s4.z, b4.z);
            y[4 * i + 3] = fmaf(__uint_as_float(v[4 * i + 3]), s4.w, b4.w);
          }
          if (e.has_res) {
            const __half2* h0 = reinterpret_cast<const __half2*>(&cur[4 * h + 2 * half]);
            const __half2* h1 = reinterpret_cast<const __half2*>(&cur[4 * h + 2 * half + 1]);
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
          uint4 o[2];
          __half2* oh2 = reinterpret_cast<__half2*>(o);
#pragma unroll
          for (int i = 0; i < 8; ++i) oh2[i] = __floats2half2_rn(y[2 * i], y[2 * i + 1]);
          ptx::sts128(my_out + ((((uint32_t)(4 * h + 2 * half)) ^ sw) << 4), o[0]);
          ptx::sts128(my_out + ((((uint32_t)(4 * h + 2 * half + 1)) ^ sw) << 4), o[1]);
        }
      }
    }
    ptx::fence_proxy_async_smem();               // generic-proxy writes -> visible to the TMA engine
    __syncwarp();
    if (e.store && ptx::elect_one()) {
      ptx::tma_store_4d(e.tm_out, e.stage_out, c64, e.c_w0, e.c_h0, e.c_img);
      ptx::tma_store_commit();
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Warp-staged coalesced epilogue (fp16 output, no sub-pixel remap).
//
// With one thread per output row every 16-byte global access of a warp touches 32 different cache lines; measured on
// the 48-channel halo-patch conv that costs ~1000 clk of load/store-unit time per tile for the stores and ~500 for the
// residual, more than the tile's MMAs in CTA-pair mode (profiles/r01_exp_patch_pair_nostore.log).  Here each epilogue
// warp owns a private [32 rows][64 channels + 16 B pad] staging tile in shared memory (pitch 144 B: both the row-per-
// thread and the coalesced access pattern are bank-conflict free) and only __syncwarp() is needed:
//   residual: cp.async 16-byte copies in coalesced order (consecutive lanes = consecutive bytes of a row) -> each
//             thread reads its own row;
//   output  : each thread writes its row, then the warp stores the tile in coalesced order.
// A warp instruction then touches 32 / (chunks per row) rows of 16 * (chunks per row) contiguous bytes: 4-6 lines.

struct EpiCoal {
  const float* s_scale;
  const float* s_bias;
  const __half* residual;   // global NHWC fp16 or nullptr
  __half* out;              // global NHWC fp16
  size_t row_off;           // element offset of (this thread's pixel, first channel of this tile)
  int ch0, ncols;           // absolute first channel / channels of this tile (multiple of 16)
  int relu;
  bool valid;               // this thread's row maps to a real output pixel
  uint32_t stage;           // shared-memory address of this warp's staging tile
};

// residual chunk [c64, c64 + nc) of the warp's 32 rows -> staging (asynchronous; complete with cp_async_wait_all)
__device__ __forceinline__ void epi_coal_fetch_residual(const EpiCoal& e, int c64, int lane) {
  const int nc = min(64, e.ncols - c64);
  const int cpr = nc >> 3;                               // 16-byte pieces per row: 2, 4, 6 or 8
  const uint32_t inv = (65536u + (uint32_t)cpr - 1u) / (uint32_t)cpr;
  const uint32_t vmask = __ballot_sync(0xffffffffu, e.valid);
  const unsigned long long off = (unsigned long long)e.row_off;
  for (int i = 0; i < cpr; ++i) {
    const uint32_t idx = (uint32_t)(i * 32 + lane);
    const uint32_t r = (idx * inv) >> 16;
    const uint32_t k = idx - r * (uint32_t)cpr;
    const unsigned long long roff = __shfl_sync(0xffffffffu, off, (int)r);
    if ((vmask >> r) & 1u)
      ptx::cp_async16(e.stage + r * (uint32_t)kCoalPitch + (k << 4), e.residual + roff + (size_t)(c64 + (int)(k << 3)));
  }
  ptx::cp_async_commit();
}

// t_row = TMEM address of (this warp's lane quarter, first column of the accumulator).  The residual of chunk 0 must
// have been requested (epi_coal_fetch_residual(e, 0, lane)) before the wait on the accumulator barrier.
__device__ __forceinline__ void epi_coal_tile(const EpiCoal& e, uint32_t t_row, int lane) {
  const uint32_t my = e.stage + (uint32_t)lane * (uint32_t)kCoalPitch;
  const uint32_t vmask = __ballot_sync(0xffffffffu, e.valid);
  const unsigned long long off = (unsigned long long)e.row_off;
  const bool has_res = e.residual != nullptr;
  for (int c64 = 0; c64 < e.ncols; c64 += 64) {
    const int nc = min(64, e.ncols - c64);     // warp-uniform
    const int cpr = nc >> 3;
    uint4 cur[8];
    if (has_res) {
      ptx::cp_async_wait_all();
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (k < cpr) cur[k] = ptx::lds128(my + (uint32_t)(k << 4));
      __syncwarp();                            // staging tile free for the outputs
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = c64 + 32 * h;
      if (32 * h < nc) {                       // warp-uniform
        uint32_t v0[16], v1[16];
        const bool two = 32 * h + 16 < nc;
        ptx::tmem_ld16(t_row + (uint32_t)c, v0);
        if (two) ptx::tmem_ld16(t_row + (uint32_t)(c + 16), v1);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          if (half == 1 && !two) break;
          const uint32_t* v = half ? v1 : v0;
          const int cc = c + 16 * half;
          const float4* sc = reinterpret_cast<const float4*>(e.s_scale + e.ch0 + cc);
          const float4* bi = reinterpret_cast<const float4*>(e.s_bias + e.ch0 + cc);
          float y[16];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 s4 = sc[i], b4 = bi[i];
            y[4 * i + 0] = fmaf(__uint_as_float(v[4 * i + 0]), s4.x, b4.x);
            y[4 * i + 1] = fmaf(__uint_as_float(v[4 * i + 1]), s4.y, b4.y);
            y[4 * i + 2] = fmaf(__uint_as_float(v[4 * i + 2]), s4.z, b4.z);
            y[4 * i + 3] = fmaf(__uint_as_float(v[4 * i + 3]), s4.w, b4.w);
          }
          if (has_res) {
            const __half2* h0 = reinterpret_cast<const __half2*>(&cur[4 * h + 2 * half]);
            const __half2* h1 = reinterpret_cast<const __half2*>(&cur[4 * h + 2 * half + 1]);
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
          uint4 o[2];
          __half2* oh2 = reinterpret_cast<__half2*>(o);
#pragma unroll
          for (int i = 0; i < 8; ++i) oh2[i] = __floats2half2_rn(y[2 * i], y[2 * i + 1]);
          ptx::sts128(my + (uint32_t)((4 * h + 2 * half) << 4), o[0]);
          ptx::sts128(my + (uint32_t)((4 * h + 2 * half + 1) << 4), o[1]);
        }
      }
    }
    __syncwarp();
    {   // coalesced store of the warp's 32 x nc tile
      const uint32_t inv = (65536u + (uint32_t)cpr - 1u) / (uint32_t)cpr;
      for (int i = 0; i < cpr; ++i) {
        const uint32_t idx = (uint32_t)(i * 32 + lane);
        const uint32_t r = (idx * inv) >> 16;
        const uint32_t k = idx - r * (uint32_t)cpr;
        const unsigned long long roff = __shfl_sync(0xffffffffu, off, (int)r);
        if ((vmask >> r) & 1u) {
          const uint4 v = ptx::lds128(e.stage + r * (uint32_t)kCoalPitch + (k << 4));
          *reinterpret_cast<uint4*>(e.out + roff + (size_t)(c64 + (int)(k << 3))) = v;
        }
      }
    }
    __syncwarp();                              // staging tile free again
    if (has_res && c64 + 64 < e.ncols) epi_coal_fetch_residual(e, c64 + 64, lane);
  }
}

}  // namespace hrnet
