// Exchange unit: every convolution of one StageModule's fuse layers (reference models_/hrnet.py:23-51, applied at :60-69)
// in ONE persistent kernel.
//
// A four-branch module has 16 of them -- six 1x1 "up" convs at the low resolutions and ten 3x3 stride-2 convs in six
// "down" chains of one to three steps -- each a few GFLOP at 64 crops, i.e. 1-3 microseconds of tensor-core time.
// Launched one by one (round 1) each cost 10-30 us of launch gap, prologue and tail, and the three-step chain 48 -> 48 ->
// 48 -> 384 put three such kernels in series on the critical path of every module (1.3 ms of serial kernel time per
// forward, 0.11 of the tensor peak).  Here the M-tiles of all member convs form one ticket sequence in dependency order
// (all first-level convs, then the second steps of the down chains, then the third); CTAs draw tickets from a global
// counter exactly like the branch chains (conv_chain.cu) and a conv of a later level starts when its producer conv has
// published all of its tiles.  The pipeline inside the CTA is the im2col pipeline of conv_igemm_body.cuh with per-op
// geometry (kernel size, stride, channels, map size, N tile) taken from an op table in the kernel parameters; outputs are
// the same fp16 terms as before (bit-identical to the per-conv launches).
//
// The sums of the unit (out_i = ReLU(sum_j nearest_up(term_ij)), models_/hrnet.py:60-69) run in the SAME kernel as "sum
// tickets": 1024 output pixels each, executed by the epilogue warpgroups (the other roles skip them), placed in the ticket
// sequence right behind the last conv level they depend on -- so the HBM-bound sums of the high-resolution outputs
// overlap the latency-bound down-chain convs of the others, and a module's exchange is one launch instead of 16 + 4.
// Same arithmetic as fuse_sum_kernel (fp32 sum in ascending branch order, one rounding), terms read with ld.global.cg.
#include <algorithm>

#include "chain_common.cuh"

namespace hrnet {

constexpr int kXThreads = 384;
constexpr uint32_t kXSumFlag = 1u << 27;      // ring entry: sum ticket (bits 28-29: output, bits 0-23: chunk)

// one sum ticket: pixels [chunk * kXSumChunk, ...) of output `sm`, by the 128 threads of one epilogue warpgroup.
// The sum is HBM-bound and only 256 threads per SM work on it, so every thread keeps four items (8 channels of one pixel
// each) x up to four sources = 16 independent 16-byte loads in flight (one item at a time ran 4x over the stand-alone
// fuse_sum_kernel's time: profiles/r02_s9_xunit_sum_tickets_v1_slow.log).
__device__ __forceinline__ void xunit_sum_chunk(const XSum& sm, int chunk, int tid128) {
  constexpr int kU = 4;
  const int CV = sm.C >> 3;
  const int p0 = chunk * kXSumChunk;
  const int npx = min(kXSumChunk, (int)sm.npix - p0);
  const int items = npx * CV;
  // item i = (pixel p0 + i / CV, channel group i % CV); a thread's items are 128 apart: (pixel, group) advance by
  // (128 / CV, 128 % CV) with carry, and (image, row, column) follow the pixel -- no division inside the loop
  const int dq = 128 / CV, dr = 128 - dq * CV;
  int px = tid128 / CV, cv = tid128 - px * CV;
  int n = (p0 + px) / (sm.H * sm.W);
  int rem = (p0 + px) - n * sm.H * sm.W;
  int h = rem / sm.W, w = rem - h * sm.W;
  for (int i0 = tid128; i0 < items; i0 += 128 * kU) {
    uint4 u[kU][4];
    size_t ooff[kU];
#pragma unroll
    for (int q = 0; q < kU; ++q) {
      ooff[q] = 0;
      if (i0 + 128 * q < items) {
        ooff[q] = (size_t)(p0 + px) * sm.C + (size_t)cv * 8;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (j < sm.nsrc) {
            const int sh = sm.shift[j];
            const int sH = sm.H >> sh, sW = sm.W >> sh;
            const size_t off = ((((size_t)n * sH + (h >> sh)) * sW + (w >> sh)) * sm.C) + (size_t)cv * 8;
            u[q][j] = __ldcg(reinterpret_cast<const uint4*>(sm.src[j] + off));
          }
        }
      }
      // advance to this thread's next item
      int adv = dq;
      cv += dr;
      if (cv >= CV) { cv -= CV; ++adv; }
      px += adv;
      w += adv;
      while (w >= sm.W) { w -= sm.W; ++h; }
      while (h >= sm.H) { h -= sm.H; ++n; }
    }
#pragma unroll
    for (int q = 0; q < kU; ++q) {
      if (i0 + 128 * q < items) {
        float acc[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (j < sm.nsrc) {
            const __half2* hh = reinterpret_cast<const __half2*>(&u[q][j]);
            float v[8];
#pragma unroll
            for (int k = 0; k < 4; ++k) { const float2 f = __half22float2(hh[k]); v[2 * k] = f.x; v[2 * k + 1] = f.y; }
            if (j == 0) {
#pragma unroll
              for (int k = 0; k < 8; ++k) acc[k] = v[k];
            } else {
#pragma unroll
              for (int k = 0; k < 8; ++k) acc[k] += v[k];
            }
          }
        }
        uint4 o;
        __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float a = acc[2 * k], b = acc[2 * k + 1];
          if (sm.relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
          oh[k] = __floats2half2_rn(a, b);
        }
        *reinterpret_cast<uint4*>(sm.out + ooff[q]) = o;
      }
    }
  }
}

constexpr int kXMaxKb = 384;     // k-blocks of all member convs together (W48: 167)

struct __align__(8) XUnitBars {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  ChainRing ring;
  uint32_t tmem_base;
  uint32_t pad;
  uint32_t kb_tab[kXMaxKb];
};

// (scale, bias) of absolute output channel i of one member conv, read straight from the parameter block
struct XSb {
  const XUnitParams& p;
  int off;
  __device__ __forceinline__ float2 operator[](int i) const { return p.sb[off + i]; }
};

__global__ void __launch_bounds__(kXThreads, 1)
conv_xunit_kernel(const __grid_constant__ XUnitMaps maps, const __grid_constant__ XUnitParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_aligned = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  const int warp = ptx::warp_idx_uniform();
  const int lane = threadIdx.x & 31;
  ptx::pdl_launch_dependents();

  const int stage_bytes = p.a_blk_bytes + p.b_blk_bytes;    // one k-block per stage
  XUnitBars* bars = reinterpret_cast<XUnitBars*>(smem_aligned + (size_t)p.stages * stage_bytes);

  if (warp == 0 && lane == 0) {
    for (int o = 0; o < p.nops; ++o) { ptx::prefetch_tmap(&maps.a[o]); ptx::prefetch_tmap(&maps.b[o]); }
    for (int i = 0; i < p.stages; ++i) {
      ptx::mbar_init(ptx::smem_u32(&bars->full[i]), 1);
      ptx::mbar_init(ptx::smem_u32(&bars->empty[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(ptx::smem_u32(&bars->tmem_full[i]), 1);
      ptx::mbar_init(ptx::smem_u32(&bars->tmem_empty[i]), 128u);
    }
    for (int i = 0; i < kChainRing; ++i) {
      ptx::mbar_init(ptx::smem_u32(&bars->ring.full[i]), 1);
      ptx::mbar_init(ptx::smem_u32(&bars->ring.empty[i]), 2u + 1u + 256u);   // producers, MMA issuer, epilogue threads
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) ptx::tmem_alloc(ptx::smem_u32(&bars->tmem_base), (uint32_t)p.tmem_cols);
  uint32_t* kb_tab = bars->kb_tab;
  if (warp == 3) {   // k-block tables of all ops: (weight K coordinate | channel offset << 15 | tap s << 27 | tap r << 29)
    for (int o = 0; o < p.nops; ++o) {
      const XOp& op = p.op[o];
      for (int kb = lane; kb < op.nkb; kb += 32) {
        const int tap = kb / op.cpt;
        const int c0 = (kb - tap * op.cpt) * kKC;
        const int r = tap / op.ksize;
        const int sx = tap - r * op.ksize;
        kb_tab[op.kb0 + kb] = (uint32_t)(tap * op.Cin + c0) | ((uint32_t)c0 << 15) | ((uint32_t)sx << 27) | ((uint32_t)r << 29);
      }
    }
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::pdl_wait();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;
  const uint32_t acc_stride = (uint32_t)(p.tmem_cols >> 1);      // two accumulator buffers
  const long long t_begin = p.dbg ? clock64() : 0;

  if (warp == 3) {
    // ===================================================================== scheduler
    if (ptx::elect_one()) {
      RingWriter rw; rw.init(&bars->ring);
      long long dbg_dep = 0; int dbg_tiles = 0;
      // Tickets interleave conv ops and sums in dependency order; `seg` walks the merged sequence (tickets only grow).
      // seg < nops + nsums; order[] is implicit: every op / sum carries its ticket0, so find the owner by range.
      uint32_t conv_ok = 0u, sum_ok = 0u;      // producers confirmed complete (per op / per sum), cached per CTA
      unsigned next = atomicAdd(&p.ctrl[0], 1u);
      for (;;) {
        const unsigned t = next;
        if (t >= (unsigned)p.total_tickets) { rw.acquire_slot(); rw.publish(kChainDone); break; }
        next = atomicAdd(&p.ctrl[0], 1u);
        int owner = -1;                       // conv op index, or nops + sum index
        for (int o = 0; o < p.nops; ++o)
          if ((int)t >= p.op[o].ticket0 && (int)t < p.op[o].ticket0 + p.op[o].m_tiles) owner = o;
        for (int q = 0; q < p.nsums; ++q)
          if ((int)t >= p.sum[q].ticket0 && (int)t < p.sum[q].ticket0 + p.sum[q].nchunks) owner = p.nops + q;
        if (owner < p.nops) {
          const XOp& op = p.op[owner];
          const int mt = (int)t - op.ticket0;
          if (!((conv_ok >> owner) & 1u)) {
            if (op.dep >= 0) {
              const long long tq = p.dbg ? clock64() : 0;
              chain_wait_counter(p.counters + op.dep, (unsigned)op.dep_need);
              fence_proxy_async_all();   // the acquired generic-proxy stores before the TMA (async-proxy) reads issued downstream
              if (p.dbg) dbg_dep += clock64() - tq;
            }
            conv_ok |= 1u << owner;
          }
          const int m0 = mt * kTileM;
          const int img = m0 / op.OHW;
          const int rem = m0 - img * op.OHW;
          const int oh0 = rem / op.OW;
          const uint32_t coord = ((uint32_t)img << 16) | ((uint32_t)oh0 << 8) | (uint32_t)(rem - oh0 * op.OW);
          for (int nt = 0; nt < op.n_tiles; ++nt) {
            rw.acquire_slot();
            rw.publish(((uint32_t)owner << 28) | ((uint32_t)nt << 24) | (uint32_t)mt, coord);
            ++dbg_tiles;
          }
        } else {
          const int q = owner - p.nops;
          const XSum& sm = p.sum[q];
          if (!((sum_ok >> q) & 1u)) {
            const long long tq = p.dbg ? clock64() : 0;
            for (int j = 0; j < sm.nsrc; ++j)
              if (sm.dep[j] >= 0) chain_wait_counter(p.counters + sm.dep[j], (unsigned)sm.dep_need[j]);
            if (p.dbg) dbg_dep += clock64() - tq;
            sum_ok |= 1u << q;
          }
          rw.acquire_slot();
          rw.publish(kXSumFlag | ((uint32_t)q << 28) | (uint32_t)((int)t - sm.ticket0));
        }
      }
      if (p.dbg) { p.dbg[blockIdx.x * 16 + 0] = dbg_dep; p.dbg[blockIdx.x * 16 + 1] = dbg_tiles; }
    }
    __syncwarp();
  } else if (warp < 2) {
    // ===================================================================== TMA producers (alternate stage loads)
    if (ptx::elect_one()) {
      RingReader rr; rr.init(&bars->ring);
      int L = 0;
      int stage = warp % p.stages;
      uint32_t phase = 0;
      for (;;) {
        uint32_t coord;
        const uint32_t info = rr.next(coord);
        if (info == kChainDone) break;
        if (info & kXSumFlag) continue;          // sum tickets belong to the epilogue warpgroups
        const int o = (int)(info >> 28), nt = (int)((info >> 24) & 7u);
        const XOp& op = p.op[o];
        const int img = (int)(coord >> 16);
        const int bw = (int)(coord & 255u) * op.stride - op.pad, bh = (int)((coord >> 8) & 255u) * op.stride - op.pad;
        const int n0 = nt * op.n_tile;
        const uint32_t tx = (uint32_t)(kTileM * kKC * 2 + op.n_tile * kKC * 2);
        for (int kb = 0; kb < op.nkb; ++kb, ++L) {
          if ((L & 1) != warp) continue;
          ptx::mbar_wait(ptx::smem_u32(&bars->empty[stage]), phase ^ 1u);
          const uint32_t full = ptx::smem_u32(&bars->full[stage]);
          const uint32_t a_dst = smem_base + (uint32_t)(stage * stage_bytes);
          const uint32_t b_dst = a_dst + (uint32_t)p.a_blk_bytes;
          ptx::mbar_expect_tx(full, tx);
          const uint32_t e = kb_tab[op.kb0 + kb];
          ptx::tma_load_im2col_4d(a_dst, &maps.a[o], full, (int)((e >> 15) & 0xfffu), bw, bh, img, (uint16_t)((e >> 27) & 3u),
                                  (uint16_t)(e >> 29));
          ptx::tma_load_2d(b_dst, &maps.b[o], full, (int)(e & 0x7fffu), n0);
          stage += 2;
          while (stage >= p.stages) { stage -= p.stages; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 2) {
    // ===================================================================== MMA issuer
    if (ptx::elect_one()) {
      RingReader rr; rr.init(&bars->ring);
      // lean issue loop (LeanPipe, conv_igemm_body.cuh): one k-block per stage, per-op k-block count / channel tail
      LeanPipe lp;
      lp.full0 = ptx::smem_u32(&bars->full[0]); lp.empty0 = ptx::smem_u32(&bars->empty[0]);
      lp.enc0 = (smem_base & 0x3FFFFu) >> 4;
      lp.enc_stage = (uint32_t)stage_bytes >> 4; lp.enc_b = (uint32_t)p.a_blk_bytes >> 4;
      lp.enc_ablk = 0u; lp.enc_bblk = 0u;
      lp.nst = p.stages; lp.reset();
      for (int it = 0;;) {
        const uint32_t info = rr.next();
        if (info == kChainDone) break;
        if (info & kXSumFlag) continue;
        const XOp& op = p.op[info >> 28];
        const uint32_t idesc = ptx::umma_idesc_f16(kTileM, op.n_tile);
        const int ctail = op.Cin - (op.cpt - 1) * kKC;
        const int acc = it & 1;
        ptx::mbar_wait(ptx::smem_u32(&bars->tmem_empty[acc]), (uint32_t)((it >> 1) & 1) ^ 1u);
        ptx::tc_fence_after_sync();
        lean_issue_tile_nk<false>((ctail + 15) / 16, lp, tmem_base + (uint32_t)acc * acc_stride, idesc, op.nkb, 1, op.cpt);
        ptx::mma_commit(ptx::smem_u32(&bars->tmem_full[acc]));
        ++it;
      }
    }
    __syncwarp();
  } else {
    // ===================================================================== epilogue (two warpgroups, alternating tiles)
    const int g = (warp - 4) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const bool leader = (q == 0) && (lane == 0);
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)g * acc_stride;
    RingReader rr; rr.init(&bars->ring);
    PendingPublish pend; pend.clear();
    int it = 0, its = 0;                 // conv tiles / sum tickets seen so far (both warpgroups count all of them)
    for (;;) {
      if (!rr.ready()) pend.flush(1 + g, leader);      // about to sleep on the ring: publish first (see PendingPublish)
      const uint32_t info = rr.next();
      if (info == kChainDone) break;
      if (info & kXSumFlag) {
        if ((its++ & 1) == g) {
          pend.flush(1 + g, leader);
          xunit_sum_chunk(p.sum[(info >> 28) & 3u], (int)(info & 0xffffffu), (int)threadIdx.x - 128 - 128 * g);
        }
        continue;
      }
      if ((it++ & 1) != g) continue;
      pend.flush(1 + g, leader);
      const uint32_t acc_phase = (uint32_t)(((it - 1) >> 1) & 1);
      const int o = (int)(info >> 28), nt = (int)((info >> 24) & 7u), mt = (int)(info & 0xffffffu);
      const XOp& op = p.op[o];
      const int m = mt * kTileM + row;
      const int n0 = nt * op.n_tile;
      EpiRow e;
      e.s_scale = op.scale; e.s_bias = op.bias; e.residual = nullptr; e.out = op.out;
      e.row_off = (size_t)m * op.Cout + n0;
      e.ch0 = n0; e.ncols = op.n_tile; e.relu = op.relu; e.out_f32 = 0; e.valid = m < op.M_total;
      U256 rres[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 8; ++i) rres[j].w[i] = 0u;
      ptx::mbar_wait(ptx::smem_u32(&bars->tmem_full[g]), acc_phase);
      ptx::tc_fence_after_sync();
      chain_store_row_c(rres, e, t_row, XSb{p, op.sb_off});      // BN constants from the kernel parameters (LDC)
      ptx::tc_fence_before_sync();
      ptx::mbar_arrive(ptx::smem_u32(&bars->tmem_empty[g]));
      pend.set(p.counters + o, false, false);
    }
    pend.flush(1 + g, leader);
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
  if (p.dbg && threadIdx.x == 0) p.dbg[blockIdx.x * 16 + 2] = clock64() - t_begin;
  chain_exit(p.ctrl, p.counters, 1, p.nops, p.nops, &bars->pad);
}

cudaError_t conv_xunit_set_attributes(int max_smem) {
  return cudaFuncSetAttribute(conv_xunit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
}

cudaError_t launch_xunit(const XUnitMaps& maps, const XUnitParams& p, int smem_bytes, int grid, cudaStream_t st) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kXThreads);
  cfg.dynamicSmemBytes = (size_t)smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  int na = 0;
  if (p.pdl) {
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = at; cfg.numAttrs = (unsigned)na;
  return cudaLaunchKernelEx(&cfg, conv_xunit_kernel, maps, p);
}

}  // namespace hrnet
