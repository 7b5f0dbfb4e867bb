// 3x3 stride-1 convolution with halo-patch operand reuse on tcgen05 (sm_100a).
//
// The im2col kernel (conv_igemm_tc.cu) re-fetches every input pixel nine times (once per filter tap) and
// streams the weights once per 128-pixel tile; measured on B200 it is bound by the L2->SM path
// (~28 B/clk/SM, ~4 clk per TMA row) at 5-26 % of tensor peak.  This kernel removes that traffic:
//
//   * output tile = 8 (w) x 16 (h) pixels of one image = 128 GEMM rows; its 10 x 18 input patch
//     (1-pixel halo, zero-filled out of bounds by TMA) is loaded ONCE per channel chunk by a tiled 4-D
//     TMA into swizzled shared memory, [18][10] pixels x (kc * 2) bytes.
//   * each of the 9 filter taps is then just a *row-shifted view* of that patch: the UMMA shared-memory
//     descriptor starts at patch + (r * 10 + s) pixel rows and steps 10 rows between 8-row core groups
//     (SBO = 10 * row bytes).  tcgen05 applies the swizzle to absolute shared-memory address bits, so a
//     descriptor that starts on a 128 B (not 1024 B) boundary reads exactly what TMA wrote
//     (profiles/r01_exp_shifted_umma_descriptor.log).
//   * the weights of all 9 taps stay resident in shared memory for the whole persistent CTA; when they do not fit
//     (e.g. transition1.0, 256 -> 48: 221 KB) the 9-tap block of a channel chunk is streamed with that chunk's patch.
//
// L2->SM traffic per tile drops from 9 x (128 x Cin) + 9 x Cin x Cout to 180 x Cin elements.
//
// Channel chunks: the patch of every chunk is a 64-channel (128 B per pixel, SWIZZLE_128B) slot; when Cin is not a
// multiple of 64 the last chunk overhangs the channel dimension, TMA zero-fills the tail and only the K16 steps
// holding real channels are issued (64 B / 32 B swizzled A operands measured ~3x slower per MMA).  The resident
// weight block of a chunk is as wide as its real channels need (64 / 32 / 16 -> 128 / 64 / 32 B swizzle).
// Warp roles (2 producers, 1 MMA issuer, 2 epilogue warpgroups), TMEM double buffering and the epilogue as in
// conv_igemm_tc.cu.
#pragma once
#include "hrnet_internal.h"
#include "epilogue.cuh"

namespace hrnet {

constexpr int kPThreads = 384;
constexpr int kPMaxSlots = 8;

struct __align__(8) PatchBars {
  uint64_t b_full;
  uint64_t a_full[kPMaxSlots];
  uint64_t a_empty[kPMaxSlots];
  uint64_t tmem_full[4];    // up to 4 accumulator buffers (p.nacc): the MMAs may run that many tiles ahead of the epilogue
  uint64_t tmem_empty[4];
  uint64_t wres_full[8];    // TMA-loaded residual tile of each epilogue warp (per-warp TMA epilogue)
  uint32_t tmem_base;
  uint32_t pad;
};

struct PatchMaps {
  CUtensorMap a;      // activations, box {64, 10, 18, 1}, SWIZZLE_128B
  CUtensorMap b[3];   // weights, box {64 | 32 | 16, Cout}
  CUtensorMap o, r;   // per-warp TMA epilogue: output / residual, box {64, 8, 4, 1}, SWIZZLE_128B
};

// All MMAs of one channel chunk: tap (r, s) = the same patch viewed from pixel row r * 10 + s (8-row core groups are one
// patch row apart); K advances 32 B (+2 in the descriptor's address field) per K16 step.
// kRowUnits: 16-byte units per patch row (8 = 128-byte rows, 4 = 64-byte rows of a 32-channel slot)
template <int NK, bool kPair, int kRowUnits = 8>
__device__ __forceinline__ void issue_taps(uint32_t d_tmem, uint64_t a0, uint64_t b0, uint32_t bstep, uint32_t idesc,
                                           uint32_t accumulate_first) {
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const uint64_t at = a0 + (uint64_t)(((t / 3) * kPatchPW + (t % 3)) * kRowUnits);
    const uint64_t bt = b0 + (uint64_t)(t * bstep);
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      if constexpr (kPair)
        ptx::mma_f16_ss_2cta(d_tmem, at + (uint64_t)(2 * k), bt + (uint64_t)(2 * k), idesc,
                             (t | k) != 0 ? 1u : accumulate_first);
      else
        ptx::mma_f16_ss(d_tmem, at + (uint64_t)(2 * k), bt + (uint64_t)(2 * k), idesc,
                        (t | k) != 0 ? 1u : accumulate_first);
    }
  }
}

// Body of one CTA working on problem `p` as CTA `cta` of `nctas` (its own persistent tile loop); shared by the
// single-problem kernel and the grouped multi-problem kernel (conv_group.cu).
// kEpi selects the epilogue at compile time (p.epi_tma must agree): 0 direct row-per-thread stores, 1 staged TMA stores,
// 2 warp-staged coalesced stores, 3 direct stores with batched TMEM loads for tiles <= 64 channels (epilogue.cuh).  Keeping several epilogues in one kernel cost the direct path
// registers (spills) and ~15 % of its speed.
// kPair (p.cs == 2, launched as clusters of two CTAs): `tcgen05.mma.cta_group::2`.  The two CTAs work on neighbouring
// tiles; one M = 256 instruction issued by the leader spans both, each CTA feeds its own patch and only HALF of the
// weight rows.  The per-MMA cost of this kernel is reading the 128 x 32 B A operand from shared memory (32 clk, more
// than the math for Cout <= 128): in pair mode both CTAs read their A in parallel and each reads half of B, so two
// tiles advance in little more than the time of one.  Barrier protocol as in conv_igemm_body.cuh (both CTAs' TMA loads
// credit the leader's `full` barriers, the leader's commits are multicast, both epilogues arrive on the leader's
// `tmem_empty`).  Not combined with streamed weights.
template <bool kPair, int kEpi>
__device__ __forceinline__ void conv3x3_patch_body(const PatchMaps& maps, const ConvPatchParams& p, const int cta,
                                                   const int nctas, uint8_t* smem_raw) {
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_aligned = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  const int warp = ptx::warp_idx_uniform();   // warp-uniform by construction (see ptx::elect_one)
  const int lane = threadIdx.x & 31;
  ptx::pdl_launch_dependents();               // the next kernel of the stream may begin its prologue
  if (p.dbg && threadIdx.x == 0) p.dbg[blockIdx.x * 32 + 16] = (long long)ptx::globaltimer();

  const uint32_t b_base = smem_base;                              // resident weights
  const uint32_t a_base = smem_base + (uint32_t)p.b_bytes;        // patch slots
  const uint32_t epi_base = a_base + (uint32_t)(p.nslots * p.slot_bytes);   // staged-epilogue tiles (1024 B aligned)
  float* s_scale = reinterpret_cast<float*>(smem_aligned + (size_t)p.b_bytes + (size_t)p.nslots * p.slot_bytes +
                                            (size_t)p.epi_bytes);
  float* s_bias = s_scale + p.Cout;
  PatchBars* bars = reinterpret_cast<PatchBars*>(s_bias + p.Cout);

  constexpr int cs = kPair ? 2 : 1;
  uint32_t crank = 0u;
  if constexpr (kPair) crank = ptx::cluster_ctarank();
  const int cluster_id = cta / cs;
  const int num_clusters = nctas / cs;
  const int total_pairs = (p.total_tiles + cs - 1) / cs;
  const uint16_t mc_mask = (uint16_t)((1u << cs) - 1u);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&maps.a);
    for (int i = 0; i < 3; ++i) ptx::prefetch_tmap(&maps.b[i]);
    if constexpr (kEpi == 1) { ptx::prefetch_tmap(&maps.o); ptx::prefetch_tmap(&maps.r); }
    ptx::mbar_init(ptx::smem_u32(&bars->b_full), (uint32_t)cs);      // pair: one expect_tx arrival per CTA (leader's barrier)
    for (int i = 0; i < p.nslots; ++i) {
      ptx::mbar_init(ptx::smem_u32(&bars->a_full[i]), (uint32_t)cs);
      ptx::mbar_init(ptx::smem_u32(&bars->a_empty[i]), 1);
    }
    for (int i = 0; i < 4; ++i) {
      ptx::mbar_init(ptx::smem_u32(&bars->tmem_full[i]), 1);
      ptx::mbar_init(ptx::smem_u32(&bars->tmem_empty[i]), 128u * (uint32_t)cs);   // pair: both CTAs' epilogues
    }
    for (int i = 0; i < 8; ++i) ptx::mbar_init(ptx::smem_u32(&bars->wres_full[i]), 1);
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    if constexpr (kPair) ptx::tmem_alloc_2cta(ptx::smem_u32(&bars->tmem_base), (uint32_t)p.tmem_cols);
    else ptx::tmem_alloc(ptx::smem_u32(&bars->tmem_base), (uint32_t)p.tmem_cols);
  }
  if (warp >= 4) {
    for (int i = threadIdx.x - 128; i < p.Cout; i += 256) {
      s_scale[i] = p.scale[i];
      s_bias[i] = p.bias[i];
    }
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if constexpr (kPair) ptx::cluster_sync_all();   // the peer's barriers must be initialised before any remote arrive
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;
  if (p.dbg && threadIdx.x == 0) p.dbg[blockIdx.x * 32 + 17] = (long long)ptx::globaltimer();
  const int tiles_per_img = p.tiles_w * p.tiles_h;

  if (warp < 2) {
    // ===================================================================== TMA producers (slot parity = warp)
    long long dbg_wait = 0, dbg_issue = 0, dbg_t0 = p.dbg ? clock64() : 0;
    if (warp == 0 && !p.b_stream) {
      // resident weights: 9 taps x nchunks blocks of [Cout rows x bkc channels]
      const int b_rows = p.Cout / cs;     // pair: this CTA keeps only its half of the weight rows
      uint32_t btx = 0;
      for (int j = 0; j < p.nchunks; ++j) btx += 9u * (uint32_t)(b_rows * p.bkc[j] * 2);
      const uint32_t bfull = ptx::smem_u32(&bars->b_full);
      if (ptx::elect_one()) {
        if constexpr (!kPair) {
          ptx::mbar_expect_tx(bfull, btx);
          for (int j = 0; j < p.nchunks; ++j)
            for (int t = 0; t < 9; ++t)
              ptx::tma_load_2d(b_base + (uint32_t)(p.boff[j] + t * p.bblk[j]), &maps.b[p.mapi[j]], bfull,
                               t * p.Cin + p.c0[j], 0);
        } else {
          const uint32_t lbfull = ptx::mapa_cluster(bfull, 0);
          ptx::mbar_expect_tx_cluster(lbfull, btx);
          for (int j = 0; j < p.nchunks; ++j)
            for (int t = 0; t < 9; ++t)
              ptx::tma_load_2d_2cta(b_base + (uint32_t)(p.boff[j] + t * p.bblk[j]), &maps.b[p.mapi[j]], lbfull,
                                    t * p.Cin + p.c0[j], (int)crank * b_rows);
        }
      }
      __syncwarp();
    }
    ptx::pdl_wait();                      // weights are constants; activations need the previous kernel
    // One issuer (p.mma_warps == 1): the two producers alternate the slot loads of every tile over the whole ring.
    // Two issuers: producer w, issuer w and epilogue warpgroup w form an independent pipeline over the tiles of parity
    // w with its own half of the slot ring (a barrier is then only ever waited on by one issuer, in consecutive phases).
    const int nw = p.mma_warps == 2 ? 2 : 1;
    const int ring = nw == 2 ? p.nslots / 2 : p.nslots;
    const int sbase = nw == 2 ? warp * ring : 0;
    int L = 0;                            // running slot-load index over the (tile, chunk) pairs of this ring
    for (int pi = cluster_id + (nw == 2 ? warp * num_clusters : 0); pi < total_pairs; pi += nw * num_clusters) {
      const int tile = min(pi * cs + (int)crank, p.total_tiles - 1);   // a ghost CTA redoes the last tile
      const int img = tile / tiles_per_img;
      const int rem = tile - img * tiles_per_img;
      const int th = rem / p.tiles_w;
      const int tw = rem - th * p.tiles_w;
      for (int j = 0; j < p.nchunks; ++j, ++L) {
        if (nw == 1 && (L & 1) != warp) continue;
        const int slot = sbase + L % ring;
        const uint32_t phase = (uint32_t)((L / ring) & 1);
        long long tq0 = 0; if (p.dbg) tq0 = clock64();
        ptx::mbar_wait(ptx::smem_u32(&bars->a_empty[slot]), phase ^ 1u);
        if (p.dbg) { const long long t = clock64(); dbg_wait += t - tq0; tq0 = t; }
        const uint32_t full = ptx::smem_u32(&bars->a_full[slot]);
        if (ptx::elect_one()) {
          const uint32_t slot_addr = a_base + (uint32_t)(slot * p.slot_bytes);
          if constexpr (kPair) {
            const uint32_t lfull = ptx::mapa_cluster(full, 0);   // the leader's barrier collects both CTAs' bytes
            ptx::mbar_expect_tx_cluster(lfull, (uint32_t)(kPatchRows * 128));
            ptx::tma_load_4d_2cta(slot_addr, &maps.a, lfull, p.c0[j], tw * kPatchTW - 1, th * kPatchTH - 1, img);
          } else {
          if (!p.b_stream) {
            ptx::mbar_expect_tx(full, (uint32_t)(kPatchRows * 128));
          } else {   // the chunk's 9-tap weight block rides in the same slot
            ptx::mbar_expect_tx(full, (uint32_t)(kPatchRows * 128 + 9 * p.Cout * p.bkc[j] * 2));
            for (int t = 0; t < 9; ++t)
              ptx::tma_load_2d(slot_addr + (uint32_t)(p.a_slot_bytes + t * p.bblk[j]), &maps.b[p.mapi[j]], full,
                               t * p.Cin + p.c0[j], 0);
          }
          ptx::tma_load_4d(slot_addr, &maps.a, full, p.c0[j], tw * kPatchTW - 1, th * kPatchTH - 1, img);
          }
        }
        __syncwarp();
        if (p.dbg) dbg_issue += clock64() - tq0;
      }
    }
    if (p.dbg && lane == 0) {
      p.dbg[blockIdx.x * 32 + 0 + 11 * warp] = dbg_wait;
      p.dbg[blockIdx.x * 32 + 1 + 11 * warp] = dbg_issue;
      p.dbg[blockIdx.x * 32 + 2 + 11 * warp] = clock64() - dbg_t0;
    }
  } else if ((warp == 2 || (warp == 3 && p.mma_warps == 2)) && (!kPair || crank == 0)) {
    // ===================================================================== MMA issuer(s) (pair mode: leader CTA only)
    // One elected thread runs the whole loop; the next slot's barrier is probed before the current chunk's MMAs are
    // issued so that its ~150 clk latency overlaps them (conv_igemm_body.cuh, profiles/r01_exp_mma_issue_overhead.log).
    // p.mma_warps == 2: warps 2 and 3 issue the MMAs of alternate tiles (tile parity = issuer = epilogue warpgroup =
    // accumulator parity), i.e. two independent MMA -> epilogue pipelines fed by the same producers.  A narrow-N MMA
    // occupies its issuing thread (~48 clk) longer than the tensor pipe (N = 48: 44 clk), and every tile adds ~800 clk
    // of serial barrier / commit latency to a single issuer; with two issuers those gaps are filled by the other tile.
    const int mw = warp - 2;
    const int nw = p.mma_warps == 2 ? 2 : 1;
    if (ptx::elect_one()) {
      const uint32_t idesc = ptx::umma_idesc_f16(kPair ? 256 : 128, p.Cout);
      if (!p.b_stream) ptx::mbar_wait(ptx::smem_u32(&bars->b_full), 0);
      // two issuers: this one owns the tiles of parity mw and the slot ring [sbase, sbase + ring) (see the producers)
      const int ring = nw == 2 ? p.nslots / 2 : p.nslots;
      const int sbase = nw == 2 ? mw * ring : 0;
      int slot = 0;                       // ring-local
      uint32_t phase = 0;
      bool ready = false;                 // result of the early probe of a_full[slot]
      bool first_chunk = true;
      int it = mw;
      const bool dbg_on = p.dbg != nullptr && mw == 0;
      long long dbg_wfull = 0, dbg_wtm = 0, dbg_mma = 0, dbg_t0 = dbg_on ? clock64() : 0;
      for (int pi = cluster_id + mw * num_clusters; pi < total_pairs; pi += nw * num_clusters, it += nw) {
        const int acc = it & (p.nacc - 1);
        const uint32_t acc_phase = (uint32_t)((it >> p.nacc_log2) & 1);
        long long tq0 = 0; if (dbg_on) tq0 = clock64();
        ptx::mbar_wait(ptx::smem_u32(&bars->tmem_empty[acc]), acc_phase ^ 1u);
        if (dbg_on) dbg_wtm += clock64() - tq0;
        ptx::tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.Cout);
        for (int j = 0; j < p.nchunks; ++j) {
          if (dbg_on) tq0 = clock64();
          if (!ready) ptx::mbar_wait(ptx::smem_u32(&bars->a_full[sbase + slot]), phase);
          if (dbg_on) { const long long t = clock64(); dbg_wfull += t - tq0; tq0 = t; if (first_chunk) p.dbg[blockIdx.x * 32 + 18] = (long long)ptx::globaltimer(); }
          first_chunk = false;
          ptx::tc_fence_after_sync();
          const uint32_t a_slot = a_base + (uint32_t)((sbase + slot) * p.slot_bytes);
          const uint32_t brow = (uint32_t)p.bkc[j] * 2u;          // weight block row bytes == its swizzle span
          const int nk = p.kreal[j] / 16;
          int nslot = slot + 1;
          uint32_t nphase = phase;
          if (nslot == ring) { nslot = 0; nphase ^= 1u; }
          const bool nready = ptx::mbar_test_wait(ptx::smem_u32(&bars->a_full[sbase + nslot]), nphase);
          // 9 taps x NK K16-steps, fully unrolled: every descriptor is base + compile-time offset
          const uint64_t a0 = ptx::umma_desc_kmajor(a_slot, 128u, (uint32_t)kPatchPW * 128u);
          const uint64_t b0 = ptx::umma_desc_kmajor(p.b_stream ? a_slot + (uint32_t)p.a_slot_bytes : b_base + (uint32_t)p.boff[j],
                                                    brow, 8u * brow);
          const uint32_t bstep = (uint32_t)p.bblk[j] >> 4;
          const uint32_t first = (uint32_t)(j != 0);
          switch (nk) {
            case 4: issue_taps<4, kPair>(d_tmem, a0, b0, bstep, idesc, first); break;
            case 3: issue_taps<3, kPair>(d_tmem, a0, b0, bstep, idesc, first); break;
            case 2: issue_taps<2, kPair>(d_tmem, a0, b0, bstep, idesc, first); break;
            default: issue_taps<1, kPair>(d_tmem, a0, b0, bstep, idesc, first); break;
          }
          // frees the patch slot (pair: in both CTAs) when the MMAs retire
          if constexpr (!kPair) ptx::mma_commit(ptx::smem_u32(&bars->a_empty[sbase + slot]));
          else ptx::mma_commit_2cta_mc(ptx::smem_u32(&bars->a_empty[sbase + slot]), mc_mask);
          slot = nslot; phase = nphase; ready = nready;
          if (dbg_on) dbg_mma += clock64() - tq0;
        }
        // accumulator ready (pair: for both CTAs' epilogues)
        if constexpr (!kPair) ptx::mma_commit(ptx::smem_u32(&bars->tmem_full[acc]));
        else ptx::mma_commit_2cta_mc(ptx::smem_u32(&bars->tmem_full[acc]), mc_mask);
      }
      if (dbg_on) {
        p.dbg[blockIdx.x * 32 + 19] = (long long)ptx::globaltimer();
        p.dbg[blockIdx.x * 32 + 4] = dbg_wfull; p.dbg[blockIdx.x * 32 + 5] = dbg_wtm;
        p.dbg[blockIdx.x * 32 + 6] = dbg_mma; p.dbg[blockIdx.x * 32 + 7] = clock64() - dbg_t0;
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================================================================== epilogue (two warpgroups, alternating tiles)
    const int g = (warp - 4) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int dh = row >> 3, dw = row & 7;
    ptx::pdl_wait();                      // residual reads / output writes need the previous kernel
    long long dbg_wacc = 0, dbg_work = 0, dbg_t0 = p.dbg ? clock64() : 0;
    uint32_t res_phase = 0;
    int it = 0;
    for (int pi = cluster_id; pi < total_pairs; pi += num_clusters, ++it) {
      if ((it & 1) != g) continue;
      const int acc = it & (p.nacc - 1);                   // warpgroup g drains the buffers of its parity
      const uint32_t acc_phase = (uint32_t)((it >> p.nacc_log2) & 1);
      const int tile_raw = pi * cs + (int)crank;
      const bool ghost = tile_raw >= p.total_tiles;        // odd tile count: the pair's second CTA recomputes, stores nothing
      const int tile = min(tile_raw, p.total_tiles - 1);
      const int img = tile / tiles_per_img;
      const int rem = tile - img * tiles_per_img;
      const int th = rem / p.tiles_w;
      const int tw = rem - th * p.tiles_w;
      if constexpr (kEpi == 1) {
        EpiWarpTma e;
        e.tm_out = &maps.o; e.tm_res = &maps.r;
        e.c_w0 = tw * kPatchTW; e.c_h0 = th * kPatchTH + 4 * q; e.c_img = img;   // this warp's 32 rows = 4 tile rows
        e.ncols = p.Cout; e.has_res = p.residual != nullptr; e.relu = p.relu; e.store = !ghost;
        e.s_scale = s_scale; e.s_bias = s_bias;
        e.stage_out = epi_base + (uint32_t)(warp - 4) * (uint32_t)(p.epi_bytes >> 3);
        e.stage_res = e.stage_out + 4096u;
        e.res_bar = ptx::smem_u32(&bars->wres_full[warp - 4]);
        if (e.has_res) {   // in flight while the MMAs of this tile finish
          if (ptx::elect_one()) epi_wtma_issue_residual(e, 0);
          __syncwarp();
        }
        long long tq0 = 0; if (p.dbg) tq0 = clock64();
        ptx::mbar_wait(ptx::smem_u32(&bars->tmem_full[acc]), acc_phase);
        if (p.dbg) { const long long t = clock64(); dbg_wacc += t - tq0; tq0 = t; }
        ptx::tc_fence_after_sync();
        epi_wtma_tile(e, tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.Cout), lane, res_phase);
        if (p.dbg) dbg_work += clock64() - tq0;
        ptx::tc_fence_before_sync();
        if (!kPair || crank == 0) ptx::mbar_arrive(ptx::smem_u32(&bars->tmem_empty[acc]));
        else if constexpr (kPair) ptx::mbar_arrive_cluster(ptx::smem_u32(&bars->tmem_empty[acc]), 0);   // the leader's MMA warp waits for both CTAs
      } else if constexpr (kEpi == 2) {
        const int oh = th * kPatchTH + dh, ow = tw * kPatchTW + dw;
        EpiCoal e;
        e.s_scale = s_scale; e.s_bias = s_bias; e.residual = p.residual; e.out = reinterpret_cast<__half*>(p.out);
        e.row_off = (((size_t)img * p.H + oh) * p.W + ow) * p.Cout;
        e.ch0 = 0; e.ncols = p.Cout; e.relu = p.relu;
        e.valid = oh < p.H && ow < p.W && !ghost;
        e.stage = epi_base + (uint32_t)(warp - 4) * (uint32_t)kCoalWarpBytes;
        if (e.residual != nullptr) epi_coal_fetch_residual(e, 0, lane);   // in flight while the MMAs of this tile finish
        long long tq0 = 0; if (p.dbg) tq0 = clock64();
        ptx::mbar_wait(ptx::smem_u32(&bars->tmem_full[acc]), acc_phase);
        if (p.dbg) { const long long t = clock64(); dbg_wacc += t - tq0; tq0 = t; }
        ptx::tc_fence_after_sync();
        epi_coal_tile(e, tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.Cout), lane);
        if (p.dbg) dbg_work += clock64() - tq0;
        ptx::tc_fence_before_sync();
        if (!kPair || crank == 0) ptx::mbar_arrive(ptx::smem_u32(&bars->tmem_empty[acc]));
        else if constexpr (kPair) ptx::mbar_arrive_cluster(ptx::smem_u32(&bars->tmem_empty[acc]), 0);
      } else {
      const int oh = th * kPatchTH + dh, ow = tw * kPatchTW + dw;
      const bool valid = oh < p.H && ow < p.W && !ghost;
      EpiRow e;
      e.s_scale = s_scale; e.s_bias = s_bias; e.residual = p.residual; e.out = p.out;
      e.row_off = (((size_t)img * p.H + oh) * p.W + ow) * p.Cout;
      e.ch0 = 0; e.ncols = p.Cout; e.relu = p.relu; e.out_f32 = p.out_f32; e.valid = valid;
      uint4 rres[8];
      epi_load_residual(rres, e, 0);            // in flight while the MMAs of this tile finish
      long long tq0 = 0; if (p.dbg) tq0 = clock64();
      ptx::mbar_wait(ptx::smem_u32(&bars->tmem_full[acc]), acc_phase);
      if (p.dbg) { const long long t = clock64(); dbg_wacc += t - tq0; tq0 = t; }
      ptx::tc_fence_after_sync();
      if constexpr (kEpi == 3) epi_store_row_batched(rres, e, tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.Cout));
      else epi_store_row(rres, e, tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.Cout));
      if (p.dbg) dbg_work += clock64() - tq0;
      ptx::tc_fence_before_sync();
      if (!kPair || crank == 0) ptx::mbar_arrive(ptx::smem_u32(&bars->tmem_empty[acc]));
      else if constexpr (kPair) ptx::mbar_arrive_cluster(ptx::smem_u32(&bars->tmem_empty[acc]), 0);
      }
    }
    if constexpr (kEpi == 1) {   // shared memory must outlive the bulk stores (same elected lane that committed them)
      if (ptx::elect_one()) ptx::tma_store_wait_all();
      __syncwarp();
    }
    if (p.dbg && threadIdx.x == 128) {
      p.dbg[blockIdx.x * 32 + 8] = dbg_wacc; p.dbg[blockIdx.x * 32 + 9] = dbg_work;
      p.dbg[blockIdx.x * 32 + 10] = clock64() - dbg_t0;
    }
  }

  if (p.dbg && threadIdx.x == 0) p.dbg[blockIdx.x * 32 + 20] = (long long)ptx::globaltimer();
  ptx::tc_fence_before_sync();
  __syncthreads();
  if constexpr (kPair) ptx::cluster_sync_all();   // no CTA may exit while its peer can still arrive on / read from it
  if (warp == 2) {
    ptx::tc_fence_after_sync();
    if constexpr (kPair) ptx::tmem_dealloc_2cta(tmem_base, (uint32_t)p.tmem_cols);
    else ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
  if (p.dbg && threadIdx.x == 0) p.dbg[blockIdx.x * 32 + 21] = (long long)ptx::globaltimer();
}


}  // namespace hrnet
