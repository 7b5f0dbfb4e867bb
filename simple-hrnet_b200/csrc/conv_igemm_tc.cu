// Implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05) for sm_100a.
//
//   out[p, co] = act( scale[co] * sum_{r,s,ci} in[pix(p) + (r,s), ci] * w[co, r, s, ci] + bias[co] (+ residual[p, co]) )
//
// GEMM view: M = output pixels (n, oh, ow) flattened (NHWC), N = Cout, K = k*k*Cin (tap-major).
//   A (activations): never materialised.  Each k-block (one filter tap x `kc` input channels) of a
//     128-pixel M-tile is fetched by ONE TMA im2col load (cp.async.bulk.tensor.4d...im2col): the
//     hardware walks 128 consecutive output pixels across row / image boundaries, applies the tap
//     offset and the conv stride, and zero-fills the padding halo.  It lands in shared memory in
//     the 128B/64B/32B-swizzled K-major layout tcgen05.mma consumes.
//   B (weights [Cout][k*k*Cin] fp16): plain 2D tiled TMA, same swizzle.
//   D: fp32 accumulators in TMEM, double buffered (2 x n_tile columns) so the epilogue of tile i
//     overlaps the MMAs of tile i+1.
//
// Persistent, warp-specialised CTA (192 threads, 1 CTA/SM):
//   warp 0   : TMA producer (one lane)         -- smem ring: full[]/empty[] mbarriers
//   warp 1   : TMEM alloc + MMA issuer (one lane) -- tcgen05.mma / tcgen05.commit
//   warps 2-5: epilogue: tcgen05.ld -> BN scale/bias (fp32) -> +residual -> ReLU -> fp16/fp32 NHWC store
//
// Replaces, for the hot path, every nn.Conv2d + nn.BatchNorm2d (+ReLU, + `out += residual`) pair of
// reference models_/modules.py:56-72 (BasicBlock), :20-40 (Bottleneck) and models_/hrnet.py:23-51,
// 98-145 (fuse / transition convs).
#include "hrnet_internal.h"
#include "ptx.cuh"

namespace hrnet {

constexpr int kTileM = 128;
constexpr int kThreads = 192;
constexpr int kMaxStages = 8;

struct __align__(8) PipeBars {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
  uint32_t pad;
};

__global__ void __launch_bounds__(kThreads, 1)
conv_igemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const ConvTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  // stage buffers need swizzle-atom (1024 B) alignment
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_aligned = smem_raw + (smem_base - ptx::smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int a_stage_bytes = p.bps * p.a_blk_bytes;
  const int b_stage_bytes = p.bps * p.b_blk_bytes;
  const int stage_bytes = a_stage_bytes + b_stage_bytes;
  // layout: [stages x (A blocks | B blocks)] [scale Cout f32] [bias Cout f32] [barriers]
  float* s_scale = reinterpret_cast<float*>(smem_aligned + (size_t)p.stages * stage_bytes);
  float* s_bias = s_scale + p.Cout;
  PipeBars* bars = reinterpret_cast<PipeBars*>(s_bias + p.Cout);

  const int total_tiles = p.m_tiles * p.n_tiles;
  const int nstages_k = (p.nkb + p.bps - 1) / p.bps;  // pipeline stages consumed per tile

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    for (int i = 0; i < p.stages; ++i) {
      ptx::mbar_init(ptx::smem_u32(&bars->full[i]), 1);
      ptx::mbar_init(ptx::smem_u32(&bars->empty[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(ptx::smem_u32(&bars->tmem_full[i]), 1);
      ptx::mbar_init(ptx::smem_u32(&bars->tmem_empty[i]), 128);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(ptx::smem_u32(&bars->tmem_base), (uint32_t)p.tmem_cols);
  }
  if (warp >= 2) {
    for (int i = threadIdx.x - 64; i < p.Cout; i += 128) {
      s_scale[i] = p.scale[i];
      s_bias[i] = p.bias[i];
    }
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int mt = tile / p.n_tiles;
        const int nt = tile - mt * p.n_tiles;
        const int m0 = mt * kTileM;
        const int img = m0 / p.OHW;
        const int rem = m0 - img * p.OHW;
        const int oh0 = rem / p.OW;
        const int ow0 = rem - oh0 * p.OW;
        const int bw = ow0 * p.stride - p.pad_w;
        const int bh = oh0 * p.stride - p.pad_h;
        const int n0 = nt * p.n_tile;
        for (int ks = 0; ks < nstages_k; ++ks) {
          const int kb0 = ks * p.bps;
          const int nblk = min(p.bps, p.nkb - kb0);
          ptx::mbar_wait(ptx::smem_u32(&bars->empty[stage]), phase ^ 1u);
          const uint32_t full = ptx::smem_u32(&bars->full[stage]);
          ptx::mbar_expect_tx(full, (uint32_t)(nblk * (kTileM * p.kc * 2 + p.n_tile * p.kc * 2)));
          const uint32_t a_dst = smem_base + (uint32_t)(stage * stage_bytes);
          const uint32_t b_dst = a_dst + (uint32_t)a_stage_bytes;
          for (int j = 0; j < nblk; ++j) {
            const int kb = kb0 + j;
            const int tap = kb / p.cpt;
            const int c0 = (kb - tap * p.cpt) * p.kc;
            const int r = tap / p.ksize;
            const int s = tap - r * p.ksize;
            ptx::tma_load_im2col_4d(a_dst + (uint32_t)(j * p.a_blk_bytes), &tmA, full, c0, bw, bh, img,
                                    (uint16_t)s, (uint16_t)r);
            ptx::tma_load_2d(b_dst + (uint32_t)(j * p.b_blk_bytes), &tmB, full, tap * p.Cin + c0, n0);
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc = ptx::umma_idesc_f16(kTileM, p.n_tile);
      const uint32_t sw_bytes = (uint32_t)p.kc * 2u;
      const uint32_t sbo = 8u * sw_bytes;
      const int k16_per_blk = p.kc / 16;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        ptx::mbar_wait(ptx::smem_u32(&bars->tmem_empty[acc]), acc_phase ^ 1u);
        ptx::tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.n_tile);
        uint32_t accumulate = 0;
        for (int ks = 0; ks < nstages_k; ++ks) {
          const int nblk = min(p.bps, p.nkb - ks * p.bps);
          ptx::mbar_wait(ptx::smem_u32(&bars->full[stage]), phase);
          ptx::tc_fence_after_sync();
          const uint32_t a_src = smem_base + (uint32_t)(stage * stage_bytes);
          const uint32_t b_src = a_src + (uint32_t)a_stage_bytes;
          for (int j = 0; j < nblk; ++j) {
            const uint64_t adesc = ptx::umma_desc_kmajor(a_src + (uint32_t)(j * p.a_blk_bytes), sw_bytes, sbo);
            const uint64_t bdesc = ptx::umma_desc_kmajor(b_src + (uint32_t)(j * p.b_blk_bytes), sw_bytes, sbo);
            for (int k = 0; k < k16_per_blk; ++k) {
              // advancing K by 16 fp16 = 32 B inside the swizzle atom: +2 in the (addr >> 4) field
              ptx::mma_f16_ss(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, accumulate);
              accumulate = 1;
            }
          }
          ptx::mma_commit(ptx::smem_u32(&bars->empty[stage]));  // frees the smem slot when the MMAs retire
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
        ptx::mma_commit(ptx::smem_u32(&bars->tmem_full[acc]));  // accumulator ready for the epilogue
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    // ===================================================================== epilogue (warps 2..5)
    const int q = warp & 3;               // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;        // accumulator row == output pixel within the tile
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int mt = tile / p.n_tiles;
      const int nt = tile - mt * p.n_tiles;
      const int m = mt * kTileM + row;
      const int n0 = nt * p.n_tile;
      const bool valid = m < p.M_total;
      ptx::mbar_wait(ptx::smem_u32(&bars->tmem_full[acc]), acc_phase);
      ptx::tc_fence_after_sync();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.n_tile);
      size_t opix = (size_t)m;
      if (p.sub) {  // sub-pixel phase of a stride-2 transposed conv: (n, i, j) -> (n, 2i+a, 2j+b)
        const int img = m / p.OHW;
        const int rem = m - img * p.OHW;
        const int i = rem / p.OW;
        const int j = rem - i * p.OW;
        opix = ((size_t)img * (2 * p.OH) + (size_t)(2 * i + p.sub_a)) * (size_t)(2 * p.OW) + (size_t)(2 * j + p.sub_b);
      }
      const size_t row_off = opix * p.Cout + n0;
      for (int c = 0; c < p.n_tile; c += 16) {
        uint32_t v[16];
        ptx::tmem_ld16(t_row + (uint32_t)c, v);
        ptx::tmem_ld_wait();
        if (valid) {
          float y[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) y[i] = __uint_as_float(v[i]) * s_scale[n0 + c + i] + s_bias[n0 + c + i];
          if (p.residual != nullptr) {
            const uint4* rp = reinterpret_cast<const uint4*>(p.residual + row_off + c);
            uint4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
            const __half2* h0 = reinterpret_cast<const __half2*>(&r0);
            const __half2* h1 = reinterpret_cast<const __half2*>(&r1);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float2 f0 = __half22float2(h0[i]), f1 = __half22float2(h1[i]);
              y[2 * i] += f0.x; y[2 * i + 1] += f0.y;
              y[8 + 2 * i] += f1.x; y[8 + 2 * i + 1] += f1.y;
            }
          }
          if (p.relu) {
#pragma unroll
            for (int i = 0; i < 16; ++i) y[i] = fmaxf(y[i], 0.f);
          }
          if (p.out_f32) {
            float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + row_off + c);
#pragma unroll
            for (int i = 0; i < 4; ++i) op[i] = make_float4(y[4 * i], y[4 * i + 1], y[4 * i + 2], y[4 * i + 3]);
          } else {
            uint4 o[2];
            __half2* oh = reinterpret_cast<__half2*>(o);
#pragma unroll
            for (int i = 0; i < 8; ++i) oh[i] = __floats2half2_rn(y[2 * i], y[2 * i + 1]);
            uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.out) + row_off + c);
            op[0] = o[0];
            op[1] = o[1];
          }
        }
      }
      // all TMEM reads of this thread are complete (wait::ld above): release the accumulator
      ptx::tc_fence_before_sync();
      ptx::mbar_arrive(ptx::smem_u32(&bars->tmem_empty[acc]));
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

cudaError_t conv_tc_set_attributes(int max_smem) {
  return cudaFuncSetAttribute(conv_igemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
}

cudaError_t launch_conv_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvTcParams& p, int smem_bytes,
                           int grid, cudaStream_t st) {
  conv_igemm_tc_kernel<<<grid, kThreads, smem_bytes, st>>>(tmA, tmB, p);
  return cudaGetLastError();
}

}  // namespace hrnet
