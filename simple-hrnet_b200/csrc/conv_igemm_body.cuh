// Implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05) for sm_100a.
//
//   out[p, co] = act( scale[co] * sum_{r,s,ci} in[pix(p) + (r,s), ci] * w[co, r, s, ci] + bias[co] (+ residual[p, co]) )
//
// GEMM view: M = output pixels (n, oh, ow) flattened (NHWC), N = Cout, K = k*k*Cin (tap-major).
//   A (activations): never materialised.  Each k-block (one filter tap x 64 input channels) of a 128-pixel M-tile
//     is fetched by ONE TMA im2col load (cp.async.bulk.tensor.4d...im2col): the hardware walks 128 consecutive
//     output pixels across row / image boundaries, applies the tap offset and the conv stride, and zero-fills the
//     padding halo.  It lands in shared memory in the 128B-swizzled K-major layout tcgen05.mma consumes.  When Cin
//     is not a multiple of 64 the last block of a tap overhangs the channel dimension: TMA zero-fills the tail and
//     the MMA loop only runs the K16 steps that hold real channels (64 B / 32 B swizzled operands measured ~3x
//     slower per MMA than 128 B ones, profiles/r01_dbg_role_timers_v2_uniform_issue.log).
//   B (weights [Cout][k*k*Cin] fp16): plain 2D tiled TMA, same swizzle.
//   CTA-pair mode (cs == 2, `tcgen05.mma.cta_group::2`): two CTAs of a cluster own neighbouring M-tiles of the same
//     N-tile; one MMA instruction issued by the leader spans both (M = 256), each CTA feeds its own A tile and only HALF
//     of the B tile from its shared memory -> per-SM weight traffic (TMA writes + MMA reads of shared memory, the
//     measured bound of this kernel) halves.  Both CTAs' TMA loads credit the leader's `full` barrier; the leader's
//     commits are multicast to both CTAs' `empty` / `tmem_full` barriers; both epilogues arrive on the leader's
//     `tmem_empty`.
//   D: fp32 accumulators in TMEM, double buffered (2 x n_tile columns) so the epilogue of tile i overlaps the MMAs of
//     tile i+1.
//
// Persistent, warp-specialised CTA (384 threads, 1 CTA/SM):
//   warps 0-1 : TMA producers, alternating pipeline stages (one producer's wait -> expect_tx -> issue chain costs
//               ~480 + 80/TMA clk per stage, profiles/r01_exp_tma_issue.log; two chains run concurrently)
//   warp  2   : TMEM alloc + MMA issuer (tcgen05.mma / tcgen05.commit)
//   warp  3   : k-block table; second MMA issuer (alternate tiles) when p.mma_warps == 2
//   warps 4-7 / 8-11 : two epilogue warpgroups, alternating tiles: tcgen05.ld -> BN scale/bias (fp32) -> +residual
//               -> ReLU -> fp16/fp32 NHWC store
// All role loops are warp-uniform with one elected lane issuing (ptx::elect_one).
//
// Replaces, for the hot path, every nn.Conv2d + nn.BatchNorm2d (+ReLU, + `out += residual`) pair of
// reference models_/modules.py:56-72 (BasicBlock), :20-40 (Bottleneck) and models_/hrnet.py:23-51,
// 98-145 (fuse / transition convs).
#pragma once
#include "hrnet_internal.h"
#include "epilogue.cuh"

namespace hrnet {

constexpr int kTileM = 128;
constexpr int kThreads = 384;
constexpr int kMaxStages = 8;
constexpr int kKC = 64;                 // channels per k-block: 128-byte swizzled rows
constexpr int kMaxKBlocks = 128;        // k-blocks per tile (taps x ceil(Cin / 64)): 9 x 512 / 64 = 72, 1 x 2048 / 64 = 32

struct __align__(8) PipeBars {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint64_t res_full[2];     // TMA-loaded residual tile of each epilogue warpgroup (staged epilogue)
  uint32_t tmem_base;
  uint32_t pad;
  uint32_t kb_tab[kMaxKBlocks];   // per k-block: weight K coordinate (15 b) | channel offset (12 b) << 15 | tap s << 27 | tap r << 29
};

// All K16 steps of one 64-channel k-block: K advances 32 B (+2 in the descriptor's address field) per step.
template <int NK, bool kPair>
__device__ __forceinline__ void issue_k16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate_first) {
#pragma unroll
  for (int k = 0; k < NK; ++k) {
    if constexpr (kPair)
      ptx::mma_f16_ss_2cta(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, k != 0 ? 1u : accumulate_first);
    else
      ptx::mma_f16_ss(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, k != 0 ? 1u : accumulate_first);
  }
}

// Lean issue loop.  What paced the wide-tile convs was not shared memory or L2 but the ISSUING THREAD: the general loop
// below spends ~750 clk per k-block on constant-bank reads, descriptor arithmetic and bookkeeping around four MMAs
// that need 384 clk of tensor time (N = 192); tools/exp/mma_tma_mix.cu runs the same loads and MMAs at 400-450 clk per
// k-block behind a minimal loop (profiles/r02_exp_mma_tma_mix.log, r02_s14_lean_issue_loop.log: C = 192 chain 20.4 k ->
// 15.0 k clk per tile).  Used whenever no debug timers are wanted.  Per stage: the next stage's barrier is probed first (its ~150 clk
// latency overlaps the MMAs), the descriptors of a stage differ from the previous one's by constants, one commit.
struct LeanPipe {
  uint32_t full0, empty0;      // `full` / `empty` barriers of stage 0 of this issuer's ring
  uint32_t enc0;               // (shared-memory address of ring stage 0) >> 4: the descriptor's start-address field
  uint32_t enc_stage, enc_b;   // >> 4: bytes per stage, offset of the B blocks inside a stage
  uint32_t enc_ablk, enc_bblk; // >> 4: bytes per A / B k-block (several k-blocks per stage)
  int nst;                     // stages in the ring
  int stage; uint32_t phase; bool ready; uint32_t enc;   // running state
  __device__ __forceinline__ void reset() { stage = 0; phase = 0u; ready = false; enc = enc0; }
};
// NKT = K16 steps of the LAST k-block of every tap (Cin % 64 real channels; all other k-blocks have four); cpt = k-blocks
// per tap.  cpt == 1: every block is a last block.
template <int NKT, bool kPair>
__device__ __forceinline__ void lean_issue_tile(LeanPipe& s, uint32_t d_tmem, uint32_t idesc, int nkb, int bps, int cpt) {
  const uint64_t desc_hi = ptx::umma_desc_kmajor(0u, 128u, 1024u);        // everything but the start address
  int kb = 0, cblk = 0;
#pragma unroll 1
  while (kb < nkb) {
    if (!s.ready) ptx::mbar_wait(s.full0 + 8u * (uint32_t)s.stage, s.phase);
    ptx::tc_fence_after_sync();
    const uint32_t a_enc = s.enc;
    const uint32_t cur_empty = s.empty0 + 8u * (uint32_t)s.stage;
    ++s.stage; s.enc += s.enc_stage;
    if (s.stage == s.nst) { s.stage = 0; s.phase ^= 1u; s.enc = s.enc0; }
    s.ready = ptx::mbar_test_wait(s.full0 + 8u * (uint32_t)s.stage, s.phase);   // answer needed one stage later
    {
      const uint64_t ad = desc_hi | (uint64_t)a_enc, bd = desc_hi | (uint64_t)(a_enc + s.enc_b);
      if (NKT == 4 || cblk != cpt - 1) issue_k16<4, kPair>(d_tmem, ad, bd, idesc, (uint32_t)(kb != 0));
      else issue_k16<NKT, kPair>(d_tmem, ad, bd, idesc, (uint32_t)(kb != 0));
      cblk = cblk + 1 == cpt ? 0 : cblk + 1;
    }
    if (bps == 2 && kb + 1 < nkb) {
      const uint64_t ad = desc_hi | (uint64_t)(a_enc + s.enc_ablk), bd = desc_hi | (uint64_t)(a_enc + s.enc_b + s.enc_bblk);
      if (NKT == 4 || cblk != cpt - 1) issue_k16<4, kPair>(d_tmem, ad, bd, idesc, 1u);
      else issue_k16<NKT, kPair>(d_tmem, ad, bd, idesc, 1u);
      cblk = cblk + 1 == cpt ? 0 : cblk + 1;
    }
    if constexpr (!kPair) ptx::mma_commit(cur_empty);
    else ptx::mma_commit_2cta_mc(cur_empty, (uint16_t)3);
    kb += bps;
  }
}
template <bool kPair>
__device__ __forceinline__ void lean_issue_tile_nk(int nk_tail, LeanPipe& s, uint32_t d_tmem, uint32_t idesc, int nkb, int bps, int cpt) {
  switch (nk_tail) {     // once per tile, not per k-block
    case 4: lean_issue_tile<4, kPair>(s, d_tmem, idesc, nkb, bps, cpt); break;
    case 3: lean_issue_tile<3, kPair>(s, d_tmem, idesc, nkb, bps, cpt); break;
    case 2: lean_issue_tile<2, kPair>(s, d_tmem, idesc, nkb, bps, cpt); break;
    default: lean_issue_tile<1, kPair>(s, d_tmem, idesc, nkb, bps, cpt); break;
  }
}

// kPair is a template parameter (not a run-time flag): a kernel that contains cta_group::2 instructions can only be
// launched as a cluster of CTA pairs ("cluster misconfiguration" otherwise).
// Body of one CTA working on problem `p` as CTA `cta` of `nctas`; shared by the single-problem kernels and the grouped
// multi-problem kernel (conv_group.cu, kPair = false only).
// kEpi selects the epilogue at compile time (p.epi_tma must agree): 0 direct row-per-thread stores, 1 staged TMA stores,
// 2 warp-staged coalesced stores, 3 direct stores with batched TMEM loads for tiles <= 64 channels (epilogue.cuh).
template <bool kPair, int kEpi>
__device__ __forceinline__ void conv_igemm_body(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap* tmO,
                                                const CUtensorMap* tmR, const ConvTcParams& p, const int cta,
                                                const int nctas, uint8_t* smem_raw) {
  // stage buffers need swizzle-atom (1024 B) alignment
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_aligned = smem_raw + (smem_base - ptx::smem_u32(smem_raw));

  const int warp = ptx::warp_idx_uniform();   // warp-uniform by construction (see ptx::elect_one)
  const int lane = threadIdx.x & 31;
  ptx::pdl_launch_dependents();               // the next kernel of the stream may begin its prologue
  if (p.dbg && threadIdx.x == 0) p.dbg[blockIdx.x * 32 + 16] = (long long)ptx::globaltimer();

  const int a_stage_bytes = p.bps * p.a_blk_bytes;
  const int b_stage_bytes = p.bps * p.b_blk_bytes;
  const int stage_bytes = a_stage_bytes + b_stage_bytes;
  // layout: [stages x (A blocks | B blocks)] [epilogue staging tiles] [scale Cout f32] [bias Cout f32] [barriers]
  const uint32_t epi_base = smem_base + (uint32_t)(p.stages * stage_bytes);
  float* s_scale = reinterpret_cast<float*>(smem_aligned + (size_t)p.stages * stage_bytes + (size_t)p.epi_bytes);
  float* s_bias = s_scale + p.Cout;
  PipeBars* bars = reinterpret_cast<PipeBars*>(s_bias + p.Cout);

  // cs == 2: CTA pair = two consecutive M-tiles of the same N-tile driven by cta_group::2 MMAs (see header).
  constexpr bool pair = kPair;
  constexpr int cs = kPair ? 2 : 1;
  uint32_t crank = 0u;
  if constexpr (kPair) crank = ptx::cluster_ctarank();
  const int cluster_id = cta / cs;
  const int num_clusters = nctas / cs;
  const int m_super = (p.m_tiles + cs - 1) / cs;
  const int total_super = m_super * p.n_tiles;
  const uint16_t mc_mask = (uint16_t)((1u << cs) - 1u);
  const int nstages_k = (p.nkb + p.bps - 1) / p.bps;  // pipeline stages consumed per tile

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    for (int i = 0; i < p.stages; ++i) {
      ptx::mbar_init(ptx::smem_u32(&bars->full[i]), (uint32_t)cs);   // pair: one expect_tx arrival per CTA (leader's barrier)
      ptx::mbar_init(ptx::smem_u32(&bars->empty[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(ptx::smem_u32(&bars->tmem_full[i]), 1);
      ptx::mbar_init(ptx::smem_u32(&bars->tmem_empty[i]), 128u * (uint32_t)cs);   // pair: both CTAs' epilogues
      ptx::mbar_init(ptx::smem_u32(&bars->res_full[i]), 1);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    if constexpr (kPair) ptx::tmem_alloc_2cta(ptx::smem_u32(&bars->tmem_base), (uint32_t)p.tmem_cols);
    else ptx::tmem_alloc(ptx::smem_u32(&bars->tmem_base), (uint32_t)p.tmem_cols);
  }
  uint32_t* kb_tab = bars->kb_tab;
  if (warp == 3) {   // k-block table: no divisions on the producers' issue path
    for (int kb = lane; kb < p.nkb; kb += 32) {
      const int tap = kb / p.cpt;
      const int c0 = (kb - tap * p.cpt) * kKC;
      const int r = tap / p.ksize;
      const int sx = tap - r * p.ksize;
      kb_tab[kb] = (uint32_t)(tap * p.Cin + c0) | ((uint32_t)c0 << 15) | ((uint32_t)sx << 27) | ((uint32_t)r << 29);
    }
  }
  if (warp >= 4) {
    for (int i = threadIdx.x - 128; i < p.Cout; i += 256) {   // constants: safe before pdl_wait
      s_scale[i] = p.scale[i];
      s_bias[i] = p.bias[i];
    }
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::pdl_wait();                       // from here on the previous kernel's outputs are visible
  if constexpr (kPair) ptx::cluster_sync_all();   // the peer's barriers must be initialised before any remote arrive
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;
  if (p.dbg && threadIdx.x == 0) p.dbg[blockIdx.x * 32 + 17] = (long long)ptx::globaltimer();

  // Issue-side costs measured on B200 (profiles/r01_exp_mma_issue_overhead.log): a tcgen05.mma takes ~48 clk of the
  // issuing thread whatever N, a tcgen05.commit ~90, an mbarrier try_wait ~150 even when the phase is already complete,
  // and every entry into an elect region ~40.  With one 64-channel k-block per stage that serial overhead (not the
  // math, not shared-memory bandwidth) set the pace: ~570 clk per k-block for every N <= 192.  Hence: ONE elected thread
  // runs each role's whole loop, the barrier state of the NEXT stage is probed (non-blocking test_wait) before the current stage's TMA / MMA
  // instructions are issued (the try_wait latency overlaps them), and small-N convs put two k-blocks in a stage.
  if (warp < 2) {
    // ===================================================================== TMA producers (stage parity = warp)
    if (ptx::elect_one()) {
      long long dbg_wait = 0, dbg_issue = 0, dbg_t0 = p.dbg ? clock64() : 0;
      const int b_rows = p.n_tile / cs;   // pair: this CTA stages only its half of the weight tile
      // One issuer (p.mma_warps == 1): the two producers alternate the stage loads of every tile over the whole ring.
      // Two issuers: producer w, issuer w and epilogue warpgroup w form an independent pipeline over the tiles of parity
      // w with its own half of the stage ring (a barrier is only ever waited on by one issuer, in consecutive phases).
      const int nw = p.mma_warps == 2 ? 2 : 1;
      const int ring = nw == 2 ? p.stages / 2 : p.stages;
      const int sbase = nw == 2 ? warp * ring : 0;
      const int step = nw == 2 ? 1 : 2;
      int L = 0;                            // running stage-load index over the tiles of this ring
      int stage = nw == 2 ? 0 : warp;       // ring-local stage / phase of the next load of THIS producer
      uint32_t phase = 0;
      bool ready = false;                   // result of the early try_wait on empty[stage]
      for (int st = cluster_id + (nw == 2 ? warp * num_clusters : 0); st < total_super; st += nw * num_clusters) {
        const int nt = st / m_super;
        const int mt = min((st - nt * m_super) * cs + (int)crank, p.m_tiles - 1);  // ghost CTAs redo the last tile
        const int m0 = mt * kTileM;
        const int img = m0 / p.OHW;
        const int rem = m0 - img * p.OHW;
        const int oh0 = rem / p.OW;
        const int ow0 = rem - oh0 * p.OW;
        const int bw = ow0 * p.stride - p.pad_w;
        const int bh = oh0 * p.stride - p.pad_h;
        const int n0 = nt * p.n_tile;
        for (int ks = 0; ks < nstages_k; ++ks, ++L) {
          if (nw == 1 && (L & 1) != warp) continue;
          const int kb0 = ks * p.bps;
          const int nblk = min(p.bps, p.nkb - kb0);
          long long tq0 = 0; if (p.dbg) tq0 = clock64();
          if (!ready) ptx::mbar_wait(ptx::smem_u32(&bars->empty[sbase + stage]), phase ^ 1u);
          if (p.dbg) { const long long t = clock64(); dbg_wait += t - tq0; tq0 = t; }
          const uint32_t full = ptx::smem_u32(&bars->full[sbase + stage]);
          const uint32_t a_dst = smem_base + (uint32_t)((sbase + stage) * stage_bytes);
          const uint32_t b_dst = a_dst + (uint32_t)a_stage_bytes;
          const uint32_t tx = (uint32_t)(nblk * (kTileM * kKC * 2 + b_rows * kKC * 2));
          // this producer's next stage: ask for its barrier state now, look at the answer after the TMA issue
          int nstage = stage + step;
          uint32_t nphase = phase;
          if (nstage >= ring) { nstage -= ring; nphase ^= 1u; }
          const bool nready = ptx::mbar_test_wait(ptx::smem_u32(&bars->empty[sbase + nstage]), nphase ^ 1u);
          if constexpr (!kPair) {
            ptx::mbar_expect_tx(full, tx);
            for (int j = 0; j < nblk; ++j) {
              const uint32_t e = kb_tab[kb0 + j];       // (weight K coordinate | channel offset | tap s | tap r)
              ptx::tma_load_im2col_4d(a_dst + (uint32_t)(j * p.a_blk_bytes), &tmA, full, (int)((e >> 15) & 0xfffu), bw, bh, img,
                                      (uint16_t)((e >> 27) & 3u), (uint16_t)(e >> 29));
              ptx::tma_load_2d(b_dst + (uint32_t)(j * p.b_blk_bytes), &tmB, full, (int)(e & 0x7fffu), n0);
            }
          } else {
            const uint32_t lfull = ptx::mapa_cluster(full, 0);   // the leader's barrier collects both CTAs' bytes
            ptx::mbar_expect_tx_cluster(lfull, tx);
            for (int j = 0; j < nblk; ++j) {
              const uint32_t e = kb_tab[kb0 + j];
              ptx::tma_load_im2col_4d_2cta(a_dst + (uint32_t)(j * p.a_blk_bytes), &tmA, lfull, (int)((e >> 15) & 0xfffu), bw, bh,
                                           img, (uint16_t)((e >> 27) & 3u), (uint16_t)(e >> 29));
              ptx::tma_load_2d_2cta(b_dst + (uint32_t)(j * p.b_blk_bytes), &tmB, lfull, (int)(e & 0x7fffu),
                                    n0 + (int)crank * b_rows);
            }
          }
          stage = nstage; phase = nphase; ready = nready;
          if (p.dbg) dbg_issue += clock64() - tq0;
        }
      }
      if (p.dbg) {
        p.dbg[blockIdx.x * 32 + 0 + 11 * warp] = dbg_wait;
        p.dbg[blockIdx.x * 32 + 1 + 11 * warp] = dbg_issue;
        p.dbg[blockIdx.x * 32 + 2 + 11 * warp] = clock64() - dbg_t0;
      }
    }
    __syncwarp();
  } else if ((warp == 2 || (warp == 3 && p.mma_warps == 2)) && (!pair || crank == 0)) {
    // ===================================================================== MMA issuer(s) (pair mode: leader CTA only)
    // p.mma_warps == 2: warps 2 and 3 issue the MMAs of alternate tiles (tile parity = issuer = accumulator = epilogue
    // warpgroup): two independent MMA -> epilogue pipelines fed by the same producers (conv3x3_patch_body.cuh).
    const int mw = warp - 2;
    const int nw = p.mma_warps == 2 ? 2 : 1;
    if (ptx::elect_one()) {
      const uint32_t idesc = ptx::umma_idesc_f16(pair ? 2 * kTileM : kTileM, p.n_tile);
      const int ctail = p.Cin - (p.cpt - 1) * kKC;          // real channels in the last k-block of a tap
      // two issuers: this one owns the tiles of parity mw and the stage ring [sbase, sbase + ring) (see the producers)
      const int ring = nw == 2 ? p.stages / 2 : p.stages;
      const int sbase = nw == 2 ? mw * ring : 0;
      int stage = 0;                        // ring-local
      uint32_t phase = 0;
      bool ready = false;                   // result of the early try_wait on full[stage]
      int it = mw;
      bool first_stage = true;
      const bool dbg_on = p.dbg != nullptr && mw == 0;
      long long dbg_wfull = 0, dbg_wtm = 0, dbg_mma = 0, dbg_t0 = dbg_on ? clock64() : 0;
      // the lean loop (see LeanPipe) unless debug timers are wanted
      const int lean_nk = (ctail + 15) / 16;          // K16 steps of the last k-block of a tap
      if (p.dbg == nullptr && p.bps <= 2) {
        LeanPipe lp;
        lp.full0 = ptx::smem_u32(&bars->full[sbase]); lp.empty0 = ptx::smem_u32(&bars->empty[sbase]);
        lp.enc0 = ((smem_base + (uint32_t)(sbase * stage_bytes)) & 0x3FFFFu) >> 4;
        lp.enc_stage = (uint32_t)stage_bytes >> 4; lp.enc_b = (uint32_t)a_stage_bytes >> 4;
        lp.enc_ablk = (uint32_t)p.a_blk_bytes >> 4; lp.enc_bblk = (uint32_t)p.b_blk_bytes >> 4;
        lp.nst = ring; lp.reset();
        for (int st = cluster_id + mw * num_clusters; st < total_super; st += nw * num_clusters, it += nw) {
          const int acc = it & 1;
          ptx::mbar_wait(ptx::smem_u32(&bars->tmem_empty[acc]), (uint32_t)((it >> 1) & 1) ^ 1u);
          ptx::tc_fence_after_sync();
          lean_issue_tile_nk<kPair>(lean_nk, lp, tmem_base + (uint32_t)(acc * p.n_tile), idesc, p.nkb, p.bps, p.cpt);
          if constexpr (!kPair) ptx::mma_commit(ptx::smem_u32(&bars->tmem_full[acc]));
          else ptx::mma_commit_2cta_mc(ptx::smem_u32(&bars->tmem_full[acc]), mc_mask);
        }
      } else
      for (int st = cluster_id + mw * num_clusters; st < total_super; st += nw * num_clusters, it += nw) {
        const int acc = it & 1;
        const uint32_t acc_phase = (uint32_t)((it >> 1) & 1);
        long long tq0 = 0; if (dbg_on) tq0 = clock64();
        ptx::mbar_wait(ptx::smem_u32(&bars->tmem_empty[acc]), acc_phase ^ 1u);
        if (dbg_on) dbg_wtm += clock64() - tq0;
        ptx::tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.n_tile);
        int cblk = 0;                         // channel block of the stage's first k-block within its tap (kb % cpt, no division)
        for (int ks = 0; ks < nstages_k; ++ks) {
          const int kb0 = ks * p.bps;
          const int nblk = min(p.bps, p.nkb - kb0);
          if (dbg_on) tq0 = clock64();
          if (!ready) ptx::mbar_wait(ptx::smem_u32(&bars->full[sbase + stage]), phase);
          if (dbg_on) { const long long t = clock64(); dbg_wfull += t - tq0; tq0 = t; if (first_stage) p.dbg[blockIdx.x * 32 + 18] = (long long)ptx::globaltimer(); }
          first_stage = false;
          ptx::tc_fence_after_sync();
          const uint32_t a_src = smem_base + (uint32_t)((sbase + stage) * stage_bytes);
          const uint32_t b_src = a_src + (uint32_t)a_stage_bytes;
          // the next stage's barrier state is requested before this stage's MMAs are issued
          int nstage = stage + 1;
          uint32_t nphase = phase;
          if (nstage == ring) { nstage = 0; nphase ^= 1u; }
          const bool nready = ptx::mbar_test_wait(ptx::smem_u32(&bars->full[sbase + nstage]), nphase);
          for (int j = 0; j < nblk; ++j) {
            const int nk = (cblk == p.cpt - 1 ? ctail : kKC) / 16;
            cblk = cblk + 1 == p.cpt ? 0 : cblk + 1;
            const uint64_t adesc = ptx::umma_desc_kmajor(a_src + (uint32_t)(j * p.a_blk_bytes), 128u, 1024u);
            const uint64_t bdesc = ptx::umma_desc_kmajor(b_src + (uint32_t)(j * p.b_blk_bytes), 128u, 1024u);
            const uint32_t first = (uint32_t)((kb0 + j) != 0);
            switch (nk) {   // fully unrolled K16 steps: descriptors are base + compile-time offset
              case 4: issue_k16<4, kPair>(d_tmem, adesc, bdesc, idesc, first); break;
              case 3: issue_k16<3, kPair>(d_tmem, adesc, bdesc, idesc, first); break;
              case 2: issue_k16<2, kPair>(d_tmem, adesc, bdesc, idesc, first); break;
              default: issue_k16<1, kPair>(d_tmem, adesc, bdesc, idesc, first); break;
            }
          }
          // frees the smem slot (pair: in both CTAs) when the MMAs retire
          if constexpr (!kPair) ptx::mma_commit(ptx::smem_u32(&bars->empty[sbase + stage]));
          else ptx::mma_commit_2cta_mc(ptx::smem_u32(&bars->empty[sbase + stage]), mc_mask);
          stage = nstage; phase = nphase; ready = nready;
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
    const int g = (warp - 4) >> 2;        // warpgroup == accumulator buffer it drains
    const int q = warp & 3;               // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;        // accumulator row == output pixel within the tile
    long long dbg_wacc = 0, dbg_work = 0, dbg_t0 = p.dbg ? clock64() : 0;
    const bool leader = (q == 0) && (lane == 0);
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * p.n_tile);
    uint32_t res_phase = 0;
    int it = 0;
    for (int st = cluster_id; st < total_super; st += num_clusters, ++it) {
      if ((it & 1) != g) continue;
      const uint32_t acc_phase = (uint32_t)((it >> 1) & 1);
      const int nt = st / m_super;
      const int mt_raw = (st - nt * m_super) * cs + (int)crank;
      const int mt = min(mt_raw, p.m_tiles - 1);
      const int m = mt * kTileM + row;
      const int n0 = nt * p.n_tile;
      if constexpr (kEpi == 1) {
        EpiTma e;
        e.tm_out = tmO; e.tm_res = tmR; e.dims4 = 0; e.c_row0 = mt * kTileM; e.c_w0 = e.c_h0 = e.c_img = 0;
        e.ch0 = n0; e.ncols = p.n_tile; e.has_res = p.residual != nullptr; e.relu = p.relu;
        e.store = mt_raw < p.m_tiles;
        e.s_scale = s_scale; e.s_bias = s_bias;
        e.stage_out = epi_base + (uint32_t)g * (uint32_t)(p.epi_bytes >> 1);
        e.stage_res = e.stage_out + 16384u;
        e.res_bar = ptx::smem_u32(&bars->res_full[g]); e.bar_id = 1 + g;
        if (e.has_res && leader) epi_tma_issue_residual(e, 0);   // in flight while the MMAs of this tile finish
        long long tq0 = 0; if (p.dbg) tq0 = clock64();
        ptx::mbar_wait(ptx::smem_u32(&bars->tmem_full[g]), acc_phase);
        if (p.dbg) { const long long t = clock64(); dbg_wacc += t - tq0; tq0 = t; }
        ptx::tc_fence_after_sync();
        epi_tma_tile(e, t_row, row, leader, res_phase);
        if (p.dbg) dbg_work += clock64() - tq0;
        ptx::tc_fence_before_sync();
        if (!pair || crank == 0) ptx::mbar_arrive(ptx::smem_u32(&bars->tmem_empty[g]));
        else if constexpr (kPair) ptx::mbar_arrive_cluster(ptx::smem_u32(&bars->tmem_empty[g]), 0);
      } else if constexpr (kEpi == 2) {
        EpiCoal e;
        e.s_scale = s_scale; e.s_bias = s_bias; e.residual = p.residual; e.out = reinterpret_cast<__half*>(p.out);
        e.row_off = (size_t)m * p.Cout + n0;
        e.ch0 = n0; e.ncols = p.n_tile; e.relu = p.relu;
        e.valid = m < p.M_total && mt_raw < p.m_tiles;
        e.stage = epi_base + (uint32_t)(warp - 4) * (uint32_t)kCoalWarpBytes;
        if (e.residual != nullptr) epi_coal_fetch_residual(e, 0, lane);   // in flight while the MMAs of this tile finish
        long long tq0 = 0; if (p.dbg) tq0 = clock64();
        ptx::mbar_wait(ptx::smem_u32(&bars->tmem_full[g]), acc_phase);
        if (p.dbg) { const long long t = clock64(); dbg_wacc += t - tq0; tq0 = t; }
        ptx::tc_fence_after_sync();
        epi_coal_tile(e, t_row, lane);
        if (p.dbg) dbg_work += clock64() - tq0;
        ptx::tc_fence_before_sync();
        if (!pair || crank == 0) ptx::mbar_arrive(ptx::smem_u32(&bars->tmem_empty[g]));
        else if constexpr (kPair) ptx::mbar_arrive_cluster(ptx::smem_u32(&bars->tmem_empty[g]), 0);
      } else {
      const bool valid = m < p.M_total && mt_raw < p.m_tiles;
      size_t opix = (size_t)m;
      if (p.sub) {  // sub-pixel phase of a stride-2 transposed conv: (n, i, j) -> (n, 2i+a, 2j+b)
        const int img = m / p.OHW;
        const int rem = m - img * p.OHW;
        const int i = rem / p.OW;
        const int j = rem - i * p.OW;
        opix = ((size_t)img * (2 * p.OH) + (size_t)(2 * i + p.sub_a)) * (size_t)(2 * p.OW) + (size_t)(2 * j + p.sub_b);
      }
      EpiRow e;
      e.s_scale = s_scale; e.s_bias = s_bias; e.residual = p.residual; e.out = p.out;
      e.row_off = opix * p.Cout + n0;
      e.ch0 = n0; e.ncols = p.n_tile; e.relu = p.relu; e.out_f32 = p.out_f32; e.valid = valid;
      uint4 rres[8];
      epi_load_residual(rres, e, 0);            // in flight while the MMAs of this tile finish
      long long tq0 = 0; if (p.dbg) tq0 = clock64();
      ptx::mbar_wait(ptx::smem_u32(&bars->tmem_full[g]), acc_phase);
      if (p.dbg) { const long long t = clock64(); dbg_wacc += t - tq0; tq0 = t; }
      ptx::tc_fence_after_sync();
      if constexpr (kEpi == 3) epi_store_row_batched(rres, e, t_row);
      else epi_store_row(rres, e, t_row);
      if (p.dbg) dbg_work += clock64() - tq0;
      // all TMEM reads of this thread are complete (wait::ld inside): release the accumulator
      ptx::tc_fence_before_sync();
      if (!pair || crank == 0) ptx::mbar_arrive(ptx::smem_u32(&bars->tmem_empty[g]));
      else if constexpr (kPair) ptx::mbar_arrive_cluster(ptx::smem_u32(&bars->tmem_empty[g]), 0);   // the leader's MMA warp waits for both CTAs
      }
    }
    if (kEpi == 1 && leader) ptx::tma_store_wait_all();   // shared memory must outlive the bulk stores
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
}


}  // namespace hrnet
