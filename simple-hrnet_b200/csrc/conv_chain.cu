// Branch chains: the eight 3x3 stride-1 convs (4 BasicBlocks, reference models_/modules.py:56-72 instantiated at
// models_/hrnet.py:15-20) of one StageModule branch run in ONE persistent kernel instead of eight launches.
//
// Why: at 64 crops a branch conv is 23 / 6.5 / 1.5 / 0.7 tiles per SM.  Launched one by one, every
// conv pays a 3-5 us prologue (barrier / TMEM setup, resident weights, first operands) and a tail that cannot overlap
// its successor (one CTA fills an SM), and the 216- / 108-tile convs of the 24x18 / 12x9 maps waste half of their
// second / only wave.  Here the tiles of all convs of the chain form ONE ordered ticket sequence (conv-major):
//
//   * a CTA draws tickets from a global counter (dynamic scheduling: a CTA only ever waits for tiles with a LOWER
//     ticket, which are owned by CTAs that are already running -> no deadlock whatever subset of the grid is resident,
//     e.g. when two forwards or the four branch chains of a module share the SMs);
//   * tile (conv k, position t) may start when the tiles of conv k-1 that cover its 3x3 halo are stored.  The output
//     map is cut into UNITS -- a row of 8x16 tiles of one image (halo-patch kernel) or a 128-pixel M-tile (im2col
//     kernel) -- and `counters[k][u]` counts the finished tiles of units u-1, u, u+1 of conv k: after its stores
//     (+ __threadfence) an epilogue warpgroup adds 1 to the (up to) three counters its tile belongs to (red.add, fire
//     and forget).  A ticket is a CHUNK of tiles of one unit (three neighbouring patch tiles / all N-tiles of an
//     M-tile); the scheduler warp needs ONE ld.acquire.gpu of counters[k-1][u] per chunk.  The last CTA to exit clears
//     the counters and the ticket (every other CTA is gone by then), so nothing has to be cleared between forwards or
//     graph replays and any batch size can follow any other;
//   * write-after-read hazards on the rotating t / y0 / y1 buffers are covered by the same chain of dependencies (a
//     tile of conv k+1 depends on every tile of conv k that read the region it overwrites);
//   * data written earlier in the same launch is read through TMA (L2) or ld.global.cg -- never through L1 / ld.nc.
//
// Inside the CTA the pipelines are the ones of conv_igemm_body.cuh / conv3x3_patch_body.cuh (same MMA order, same
// epilogue arithmetic -> results are bit-identical to the per-conv launches, which is what tests/test_gpu_chain.py
// asserts), plus a scheduler (warp 3 of the im2col kernel, warp 1 of the halo-patch kernel) that feeds a small ring of
// tile descriptors to the producer / MMA / epilogue roles.
#include <algorithm>

#include "conv3x3_patch_body.cuh"
#include "chain_common.cuh"

namespace hrnet {

// =====================================================================================================================
// im2col chain (branches whose map does not tile into 8x16 patches / whose weights do not fit: C = 192, 384 at W48)
// =====================================================================================================================
constexpr int kCIThreads = 384;

struct __align__(8) ChainIgemmBars {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  ChainRing ring;
  uint32_t tmem_base;
  uint32_t pad;
  uint32_t kb_tab[kMaxKBlocks];
};

// kPair (cta_group::2): the two CTAs of a cluster own the two M-tiles of a unit.  One MMA instruction issued by the leader
// spans both (M = 256); each CTA stages its own A tile and only HALF of the weight tile (rows [crank * n_tile / 2, ...)),
// so the shared-memory port of an SM carries 16 + 12 KB of TMA writes and 28 KB of operand reads per 128 x 192 x 64
// step instead of 40 + 40 KB (the measured bound of the single-CTA pipeline) -- and unlike the
// two-tiles-per-ticket mode the accumulators stay double buffered (each CTA's TMEM holds its own 128 rows).
//   * tickets are drawn by the LEADER's scheduler; it writes every descriptor into both CTAs' rings (st.shared::cluster +
//     a release.cluster arrive on the peer's `full` barrier); all consumers of both CTAs arrive on the leader's `empty`;
//   * both CTAs' TMA loads credit the leader's stage barrier; the leader's commits are multicast to both CTAs' `empty` /
//     `tmem_full` barriers; both CTAs' epilogues arrive on the leader's `tmem_empty` (as in conv_igemm_body.cuh);
//   * each CTA's epilogue drains and publishes its own M-tile (the unit counters expect two tiles per N-tile).
template <bool kPair>
__global__ void __launch_bounds__(kCIThreads, 1)
conv_chain_igemm_kernel(const __grid_constant__ ChainIgemmMaps maps, const __grid_constant__ ChainIgemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_aligned = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  const int warp = ptx::warp_idx_uniform();
  const int lane = threadIdx.x & 31;
  ptx::pdl_launch_dependents();

  // Two M-tiles per ticket (p.m2 == 2): both tiles consume every weight k-block of a stage, so a stage carries A0 | A1 | B
  // and the TMA bytes per output drop by 30 % (weights fetched once per 256 rows).  What paced the single-tile pipeline was
  // the per-SM TMA delivery rate: 40 KB per 128 x 192 x 64 step in ~800 clk against 384 clk of math
  // (profiles/r02_s3_*.log: 17.9 k clk per 27-k-block tile).  The two accumulators fill the 2 x n_tile TMEM columns the
  // single-tile mode uses for double buffering, so the MMAs of a ticket wait for the epilogue of the previous one; the
  // operand loads keep streaming meanwhile.
  const int m2 = kPair ? 1 : p.m2;        // M-tiles this CTA owns per ticket
  const int tpu = kPair ? 2 : p.m2;       // M-tiles per unit (= per ticket)
  uint32_t crank = 0u;
  if constexpr (kPair) crank = ptx::cluster_ctarank();
  const bool peer = kPair && crank != 0u;
  const int a_stage_bytes = m2 * p.bps * p.a_blk_bytes;
  const int b_stage_bytes = p.bps * p.b_blk_bytes;
  const int stage_bytes = a_stage_bytes + b_stage_bytes;
  ChainIgemmBars* bars = reinterpret_cast<ChainIgemmBars*>(smem_aligned + (size_t)p.stages * stage_bytes);
  const int nstages_k = (p.nkb + p.bps - 1) / p.bps;

  if (warp == 0 && lane == 0) {
    for (int k = 0; k < p.nconv; ++k) { ptx::prefetch_tmap(&maps.a[k]); ptx::prefetch_tmap(&maps.b[k]); }
    for (int i = 0; i < p.stages; ++i) {
      ptx::mbar_init(ptx::smem_u32(&bars->full[i]), kPair ? 2u : 1u);      // pair: one expect_tx arrival per CTA
      ptx::mbar_init(ptx::smem_u32(&bars->empty[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(ptx::smem_u32(&bars->tmem_full[i]), 1);
      ptx::mbar_init(ptx::smem_u32(&bars->tmem_empty[i]), 128u * (uint32_t)tpu);   // pair: both CTAs' epilogues
    }
    for (int i = 0; i < kChainRing; ++i) {
      ptx::mbar_init(ptx::smem_u32(&bars->ring.full[i]), 1);
      // producers, MMA issuer, epilogue threads (pair: + the peer's producers and epilogue threads, on the leader's barrier)
      ptx::mbar_init(ptx::smem_u32(&bars->ring.empty[i]), 2u + 1u + 256u + (kPair ? 2u + 256u : 0u));
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    if constexpr (kPair) ptx::tmem_alloc_2cta(ptx::smem_u32(&bars->tmem_base), (uint32_t)p.tmem_cols);
    else ptx::tmem_alloc(ptx::smem_u32(&bars->tmem_base), (uint32_t)p.tmem_cols);
  }
  uint32_t* kb_tab = bars->kb_tab;
  if (warp == 3) {
    for (int kb = lane; kb < p.nkb; kb += 32) {
      const int tap = kb / p.cpt;
      const int c0 = (kb - tap * p.cpt) * kKC;
      const int r = tap / 3;
      const int sx = tap - r * 3;
      kb_tab[kb] = (uint32_t)(tap * p.C + c0) | ((uint32_t)c0 << 15) | ((uint32_t)sx << 27) | ((uint32_t)r << 29);
    }
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::pdl_wait();
  if constexpr (kPair) ptx::cluster_sync_all();   // the peer's barriers must be initialised before any remote arrive
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;
  const long long t_begin = p.dbg ? clock64() : 0;

  if (warp == 3) {
    // ===================================================================== scheduler (pair: the leader's, for both CTAs)
    if (!peer && ptx::elect_one()) {
      RingWriter rw; rw.init(&bars->ring, kPair);
      long long dbg_dep = 0; int dbg_tiles = 0;
      // one ticket = one unit = m2 neighbouring M-tiles (all of their N-tiles: same dependencies, same A operands)
      const unsigned total_chunks = (unsigned)(p.nconv * p.units);
      unsigned next = atomicAdd(&p.ctrl[0], 1u);
      for (;;) {
        const unsigned t = next;
        if (t >= total_chunks) { rw.acquire_slot(); rw.publish(kChainDone); break; }
        next = atomicAdd(&p.ctrl[0], 1u);            // in flight while this ticket's dependency is polled
        const int k = (int)t / p.units;
        const int u = (int)t - k * p.units;
        if (k > 0) {
          // rows [m0 - (OW + 1), m_last + OW + 1] of conv k-1: OW + 1 < 128, so the neighbouring units (all N-tiles)
          const long long tq = p.dbg ? clock64() : 0;
          const int nb = 1 + (u > 0 ? 1 : 0) + (u < p.units - 1 ? 1 : 0);
          chain_wait_counter(p.counters + (size_t)(k - 1) * p.unit_stride + u, (unsigned)(nb * p.n_tiles * tpu));
          fence_proxy_async_all();   // the acquired generic-proxy stores before the TMA (async-proxy) reads issued downstream
          if (p.dbg) dbg_dep += clock64() - tq;
        }
        uint32_t coord[2] = {0u, 0u};
        for (int i = 0; i < tpu; ++i) {
          const int m0 = (u * tpu + i) * kTileM;    // (a tile past the last image loads zeros and stores nothing)
          const int img = m0 / p.OHW;
          const int rem = m0 - img * p.OHW;
          const int oh0 = rem / p.OW;
          coord[i] = ((uint32_t)img << 16) | ((uint32_t)oh0 << 8) | (uint32_t)(rem - oh0 * p.OW);
        }
        for (int nt = 0; nt < p.n_tiles; ++nt) {
          rw.acquire_slot();
          rw.publish(((uint32_t)k << 28) | ((uint32_t)nt << 24) | (uint32_t)u, coord[0], coord[1]);
          ++dbg_tiles;
        }
      }
      if (p.dbg) { p.dbg[blockIdx.x * 16 + 0] = dbg_dep; p.dbg[blockIdx.x * 16 + 1] = dbg_tiles * m2; }
      if (p.dbg && kPair) p.dbg[(blockIdx.x + 1) * 16 + 1] = dbg_tiles;
    }
    __syncwarp();
  } else if (warp < 2) {
    // ===================================================================== TMA producers (alternate stage loads)
    if (ptx::elect_one()) {
      RingReader rr; rr.init(&bars->ring, peer);
      const int b_rows = kPair ? p.n_tile / 2 : p.n_tile;   // pair: this CTA stages only its half of the weight tile
      int L = 0;
      int stage = warp % p.stages;
      uint32_t phase = (uint32_t)(warp / p.stages) & 1u;
      long long dbg_slot = 0;
      for (;;) {
        uint32_t coord, coord2;
        const uint32_t info = rr.next(coord, coord2);
        if (info == kChainDone) break;
        const int k = (int)(info >> 28), nt = (int)((info >> 24) & 15u);
        if (peer) coord = coord2;                 // pair: the peer owns the second M-tile of the unit
        const int img0 = (int)(coord >> 16), bw0 = (int)(coord & 255u) - 1, bh0 = (int)((coord >> 8) & 255u) - 1;
        const int img1 = (int)(coord2 >> 16), bw1 = (int)(coord2 & 255u) - 1, bh1 = (int)((coord2 >> 8) & 255u) - 1;
        const int n0 = nt * p.n_tile;
        for (int ks = 0; ks < nstages_k; ++ks, ++L) {
          if ((L & 1) != warp) continue;
          const int kb0 = ks * p.bps;
          const int nblk = min(p.bps, p.nkb - kb0);
          const long long tq = p.dbg ? clock64() : 0;
          ptx::mbar_wait(ptx::smem_u32(&bars->empty[stage]), phase ^ 1u);
          if (p.dbg) dbg_slot += clock64() - tq;
          const uint32_t full = ptx::smem_u32(&bars->full[stage]);
          const uint32_t a_dst = smem_base + (uint32_t)(stage * stage_bytes);
          const uint32_t b_dst = a_dst + (uint32_t)a_stage_bytes;
          const uint32_t tx = (uint32_t)(nblk * (m2 * kTileM * kKC * 2 + b_rows * kKC * 2));
          if constexpr (!kPair) {
            ptx::mbar_expect_tx(full, tx);
            for (int j = 0; j < nblk; ++j) {
              const uint32_t e = kb_tab[kb0 + j];
              ptx::tma_load_im2col_4d(a_dst + (uint32_t)(j * p.a_blk_bytes), &maps.a[k], full, (int)((e >> 15) & 0xfffu), bw0, bh0,
                                      img0, (uint16_t)((e >> 27) & 3u), (uint16_t)(e >> 29));
              if (m2 == 2)
                ptx::tma_load_im2col_4d(a_dst + (uint32_t)((p.bps + j) * p.a_blk_bytes), &maps.a[k], full, (int)((e >> 15) & 0xfffu),
                                        bw1, bh1, img1, (uint16_t)((e >> 27) & 3u), (uint16_t)(e >> 29));
              ptx::tma_load_2d(b_dst + (uint32_t)(j * p.b_blk_bytes), &maps.b[k], full, (int)(e & 0x7fffu), n0);
            }
          } else {
            const uint32_t lfull = ptx::mapa_cluster(full, 0);   // the leader's barrier collects both CTAs' bytes
            ptx::mbar_expect_tx_cluster(lfull, tx);
            for (int j = 0; j < nblk; ++j) {
              const uint32_t e = kb_tab[kb0 + j];
              ptx::tma_load_im2col_4d_2cta(a_dst + (uint32_t)(j * p.a_blk_bytes), &maps.a[k], lfull, (int)((e >> 15) & 0xfffu), bw0,
                                           bh0, img0, (uint16_t)((e >> 27) & 3u), (uint16_t)(e >> 29));
              ptx::tma_load_2d_2cta(b_dst + (uint32_t)(j * p.b_blk_bytes), &maps.b[k], lfull, (int)(e & 0x7fffu),
                                    n0 + (int)crank * b_rows);
            }
          }
          // this producer's next load is two stages further round the ring
          stage += 2;
          while (stage >= p.stages) { stage -= p.stages; phase ^= 1u; }
        }
      }
      if (p.dbg) p.dbg[blockIdx.x * 16 + 11 + warp] = dbg_slot;
    }
    __syncwarp();
  } else if (warp == 2) {
    // ===================================================================== MMA issuer (pair: the leader CTA only)
    if (!peer && ptx::elect_one()) {
      RingReader rr; rr.init(&bars->ring, false);
      const uint32_t idesc = ptx::umma_idesc_f16(kPair ? 2 * kTileM : kTileM, p.n_tile);
      const int ctail = p.C - (p.cpt - 1) * kKC;
      int stage = 0;
      uint32_t phase = 0;
      bool ready = false;
      long long dbg_wfull = 0, dbg_wtm = 0, dbg_ring = 0;
      // Lean issue loop (LeanPipe, conv_igemm_body.cuh) for every production chain: one M-tile per CTA and ticket, 64 real
      // channels per k-block.  (skip == 9: experiment, the general loop.)
      if (m2 == 1 && p.bps <= 2 && (p.C & 63) == 0 && p.skip != 9) {
        LeanPipe lp;
        lp.full0 = ptx::smem_u32(&bars->full[0]); lp.empty0 = ptx::smem_u32(&bars->empty[0]);
        lp.enc0 = (smem_base & 0x3FFFFu) >> 4;
        lp.enc_stage = (uint32_t)stage_bytes >> 4; lp.enc_b = (uint32_t)a_stage_bytes >> 4;
        lp.enc_ablk = (uint32_t)p.a_blk_bytes >> 4; lp.enc_bblk = (uint32_t)p.b_blk_bytes >> 4;
        lp.nst = p.stages; lp.reset();
        for (int it = 0;; ++it) {
          const uint32_t info = rr.next();
          if (info == kChainDone) break;
          const int acc = it & 1;
          ptx::mbar_wait(ptx::smem_u32(&bars->tmem_empty[acc]), (uint32_t)((it >> 1) & 1) ^ 1u);
          ptx::tc_fence_after_sync();
          lean_issue_tile<4, kPair>(lp, tmem_base + (uint32_t)(acc * p.n_tile), idesc, p.nkb, p.bps, p.cpt);
          if constexpr (!kPair) ptx::mma_commit(ptx::smem_u32(&bars->tmem_full[acc]));
          else ptx::mma_commit_2cta_mc(ptx::smem_u32(&bars->tmem_full[acc]), (uint16_t)3);
        }
      } else
      for (int it = 0;; ++it) {
        long long tr = p.dbg ? clock64() : 0;
        const uint32_t info = rr.next();
        if (p.dbg) dbg_ring += clock64() - tr;
        if (info == kChainDone) break;
        // single tile: two accumulator buffers alternate; two tiles: both halves belong to this ticket
        const int acc = m2 == 2 ? 0 : (it & 1);
        const uint32_t acc_phase = (uint32_t)((m2 == 2 ? it : (it >> 1)) & 1);
        long long tq = p.dbg ? clock64() : 0;
        ptx::mbar_wait(ptx::smem_u32(&bars->tmem_empty[acc]), acc_phase ^ 1u);
        if (p.dbg) dbg_wtm += clock64() - tq;
        ptx::tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.n_tile);
        int cblk = 0;
        for (int ks = 0; ks < nstages_k; ++ks) {
          const int kb0 = ks * p.bps;
          const int nblk = min(p.bps, p.nkb - kb0);
          if (p.dbg) tq = clock64();
          if (!ready) ptx::mbar_wait(ptx::smem_u32(&bars->full[stage]), phase);
          if (p.dbg) dbg_wfull += clock64() - tq;
          ptx::tc_fence_after_sync();
          const uint32_t a_src = smem_base + (uint32_t)(stage * stage_bytes);
          const uint32_t b_src = a_src + (uint32_t)a_stage_bytes;
          int nstage = stage + 1;
          uint32_t nphase = phase;
          if (nstage == p.stages) { nstage = 0; nphase ^= 1u; }
          const bool nready = ptx::mbar_test_wait(ptx::smem_u32(&bars->full[nstage]), nphase);
          for (int j = 0; j < nblk; ++j) {
            const int nk = (cblk == p.cpt - 1 ? ctail : kKC) / 16;
            cblk = cblk + 1 == p.cpt ? 0 : cblk + 1;
            const uint64_t bdesc = ptx::umma_desc_kmajor(b_src + (uint32_t)(j * p.b_blk_bytes), 128u, 1024u);
            const uint32_t first = (uint32_t)((kb0 + j) != 0);
            for (int i = 0; i < m2; ++i) {
              const uint64_t adesc = ptx::umma_desc_kmajor(a_src + (uint32_t)((i * p.bps + j) * p.a_blk_bytes), 128u, 1024u);
              const uint32_t d = d_tmem + (uint32_t)(i * p.n_tile);
              switch (nk) {
                case 4: issue_k16<4, kPair>(d, adesc, bdesc, idesc, first); break;
                case 3: issue_k16<3, kPair>(d, adesc, bdesc, idesc, first); break;
                case 2: issue_k16<2, kPair>(d, adesc, bdesc, idesc, first); break;
                default: issue_k16<1, kPair>(d, adesc, bdesc, idesc, first); break;
              }
            }
          }
          if constexpr (!kPair) ptx::mma_commit(ptx::smem_u32(&bars->empty[stage]));
          else ptx::mma_commit_2cta_mc(ptx::smem_u32(&bars->empty[stage]), (uint16_t)3);   // frees the stage in both CTAs
          stage = nstage; phase = nphase; ready = nready;
        }
        if constexpr (!kPair) ptx::mma_commit(ptx::smem_u32(&bars->tmem_full[acc]));
        else ptx::mma_commit_2cta_mc(ptx::smem_u32(&bars->tmem_full[acc]), (uint16_t)3);
      }
      if (p.dbg) { p.dbg[blockIdx.x * 16 + 5] = dbg_wfull; p.dbg[blockIdx.x * 16 + 6] = dbg_wtm; p.dbg[blockIdx.x * 16 + 7] = dbg_ring; }
    }
    __syncwarp();
  } else {
    // ===================================================================== epilogue (two warpgroups)
    // single tile per ticket: warpgroup g drains buffer g of alternate tickets; two tiles: warpgroup g drains tile g of
    // every ticket
    const int g = (warp - 4) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const bool leader = (q == 0) && (lane == 0);
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * p.n_tile);
    RingReader rr; rr.init(&bars->ring, peer);
    PendingPublish pend; pend.clear();
    long long dbg_wait = 0, dbg_work = 0;
    for (int it = 0;; ++it) {
      if (!rr.ready()) pend.flush(1 + g, leader);      // about to sleep on the ring: publish first (see PendingPublish)
      const uint32_t info = rr.next();
      if (info == kChainDone) break;
      if (m2 == 1 && (it & 1) != g) continue;
      pend.flush(1 + g, leader);                        // the previous tile of this warpgroup
      const int accb = m2 == 2 ? 0 : g;
      const uint32_t acc_phase = (uint32_t)((m2 == 2 ? it : (it >> 1)) & 1);
      const int k = (int)(info >> 28), nt = (int)((info >> 24) & 15u), u = (int)(info & 0xffffffu);
      const ChainConv& cv = p.conv[k];
      const int m = (u * tpu + (kPair ? (int)crank : (m2 == 2 ? g : 0))) * kTileM + row;
      const int n0 = nt * p.n_tile;
      EpiRow e;
      e.s_scale = cv.scale; e.s_bias = cv.bias; e.residual = cv.residual; e.out = cv.out;
      e.row_off = (size_t)m * p.C + n0;
      e.ch0 = n0; e.ncols = p.n_tile; e.relu = cv.relu; e.out_f32 = 0; e.valid = m < p.M_total;
      U256 rres[4];
      const int eskip = p.skip >= 8 ? 0 : p.skip;      // (8, 9 select other experiments)
      if (eskip == 0 || eskip == 3) chain_load_residual(rres, e, 0);
      long long tq = p.dbg ? clock64() : 0;
      ptx::mbar_wait(ptx::smem_u32(&bars->tmem_full[accb]), acc_phase);
      if (p.dbg) { const long long t = clock64(); dbg_wait += t - tq; tq = t; }
      ptx::tc_fence_after_sync();
      if (eskip == 0) {
        chain_store_row_c(rres, e, t_row, p.sb[k]);
      } else if (eskip == 2) {            // experiment (results invalid): accumulator loads only
        uint32_t acc = 0;
        for (int c = 0; c < e.ncols; c += 16) {
          uint32_t v[16];
          ptx::tmem_ld16(t_row + (uint32_t)c, v);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) acc ^= v[i];
        }
        if (acc == 0x12345679u && e.valid) reinterpret_cast<__half*>(e.out)[e.row_off] = __float2half(0.f);
      } else if (eskip == 3) {            // experiment (results invalid): residual loads and output stores only
        if (e.valid)
          for (int c = 0; c < e.ncols; c += 64) {
            if (c) chain_load_residual(rres, e, c);
#pragma unroll
            for (int i = 0; i < 4; ++i) if (c + 16 * i < e.ncols) stg256(reinterpret_cast<__half*>(e.out) + e.row_off + c + 16 * i, rres[i]);
          }
      }                                    // skip == 1: nothing at all
      ptx::tc_fence_before_sync();
      if (!peer) ptx::mbar_arrive(ptx::smem_u32(&bars->tmem_empty[accb]));
      else ptx::mbar_arrive_cluster(ptx::smem_u32(&bars->tmem_empty[accb]), 0);   // the leader's MMA warp waits for both CTAs
      pend.set(p.counters + (size_t)k * p.unit_stride + u, u > 0, u < p.units - 1);
      if (p.dbg) dbg_work += clock64() - tq;
    }
    pend.flush(1 + g, leader);
    if (p.dbg && threadIdx.x == 128) { p.dbg[blockIdx.x * 16 + 3] = dbg_wait; p.dbg[blockIdx.x * 16 + 4] = dbg_work; }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if constexpr (kPair) ptx::cluster_sync_all();   // no CTA may exit while its peer can still arrive on / read from it
  if (warp == 2) {
    ptx::tc_fence_after_sync();
    if constexpr (kPair) ptx::tmem_dealloc_2cta(tmem_base, (uint32_t)p.tmem_cols);
    else ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
  if (p.dbg && threadIdx.x == 0) p.dbg[blockIdx.x * 16 + 2] = clock64() - t_begin;
  chain_exit(p.ctrl, p.counters, p.nconv, p.units, p.unit_stride, &bars->pad);
}

// =====================================================================================================================
// halo-patch chain (C = 48, 96 at W48; C = 32, 64 at W32): resident weights are swapped at every conv boundary
// =====================================================================================================================
// Warp 0 producer (one TMA per channel chunk of a tile: a single issuing thread keeps up, unlike the im2col kernel's
// per-k-block loads), warp 1 scheduler, warps 2-3 MMA issuers (3 only when pp.mma_warps == 2), warps 4-19 epilogue:
// FOUR epilogue warpgroups, warpgroup g drains the tiles it = g (mod 4); tile it accumulates in buffer it & (nacc - 1).
// EIGHT accumulator buffers when Cout <= 64: an issuer's next tile needs a buffer back from the epilogue, and a tile's
// epilogue takes ~3,000 clk (longer than two tiles' MMAs) -- with four buffers (two per issuer) the C = 48 issuers spent
// 19 % of their time waiting for TMEM (profiles/r02_s13_*.log: wait_tmem 209 k of 1,079 k clk).  The thread-per-row
// epilogue of a narrow tile is latency-bound (profiles/r01_exp_epilogue_cost.log: one or two warps per scheduler
// partition expose every TMEM-load / shared-load / FMA -> MAX -> CVT chain) and paced the C = 48 convs at 2,150 clk per
// tile where the MMAs need 1,190; four warps per partition hide those latencies behind each other.  20 warps cap a
// thread at 96 registers (5 warps on a partition share its 16 K registers).
constexpr int kCPThreads = 640;
constexpr int kCPEpiGroups = 4;

struct __align__(8) ChainPatchBars {
  uint64_t b_full;          // resident weights of the current conv have landed (one phase per conv of this CTA)
  uint64_t b_empty;         // every MMA issuer has retired its MMAs on the previous conv's weights
  uint64_t a_full[kPMaxSlots];
  uint64_t a_empty[kPMaxSlots];
  uint64_t tmem_full[8];
  uint64_t tmem_empty[8];
  ChainRing ring;
  uint32_t tmem_base;
  uint32_t pad;
};

__global__ void __launch_bounds__(kCPThreads, 1)
conv_chain_patch_kernel(const __grid_constant__ ChainPatchMaps maps, const __grid_constant__ ChainPatchParams cp) {
  extern __shared__ uint8_t smem_raw[];
  const ConvPatchParams& p = cp.pp;
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_aligned = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  const int warp = ptx::warp_idx_uniform();
  const int lane = threadIdx.x & 31;
  ptx::pdl_launch_dependents();

  const uint32_t b_base = smem_base;
  const uint32_t a_base = smem_base + (uint32_t)p.b_bytes;
  // tail64: two 128-byte-row slots + one 64-byte-row slot (180 x 64 B rounded up to 12 KB) instead of p.nslots equal slots
  const size_t slots_bytes = cp.tail64 ? (size_t)2 * p.slot_bytes + 12288u : (size_t)p.nslots * p.slot_bytes;
  ChainPatchBars* bars = reinterpret_cast<ChainPatchBars*>(smem_aligned + (size_t)p.b_bytes + slots_bytes);
  const int n_issuers = p.mma_warps == 2 ? 2 : 1;

  if (warp == 0 && lane == 0) {
    for (int k = 0; k < cp.nconv; ++k) {
      ptx::prefetch_tmap(&maps.a[k]); ptx::prefetch_tmap(&maps.b[k][0]); ptx::prefetch_tmap(&maps.b[k][1]);
    }
    ptx::mbar_init(ptx::smem_u32(&bars->b_full), 1);
    ptx::mbar_init(ptx::smem_u32(&bars->b_empty), (uint32_t)n_issuers);
    for (int i = 0; i < (cp.tail64 ? 3 : p.nslots); ++i) {
      ptx::mbar_init(ptx::smem_u32(&bars->a_full[i]), 1);
      ptx::mbar_init(ptx::smem_u32(&bars->a_empty[i]), 1);
    }
    for (int i = 0; i < 8; ++i) {
      ptx::mbar_init(ptx::smem_u32(&bars->tmem_full[i]), 1);
      ptx::mbar_init(ptx::smem_u32(&bars->tmem_empty[i]), 128u);
    }
    for (int i = 0; i < kChainRing; ++i) {
      ptx::mbar_init(ptx::smem_u32(&bars->ring.full[i]), 1);
      ptx::mbar_init(ptx::smem_u32(&bars->ring.empty[i]), 1u + (uint32_t)n_issuers + 128u * kCPEpiGroups);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) ptx::tmem_alloc(ptx::smem_u32(&bars->tmem_base), (uint32_t)p.tmem_cols);
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;
  const long long t_begin = cp.dbg ? clock64() : 0;

  // bytes of one conv's resident weight set
  uint32_t btx = 0;
  for (int j = 0; j < p.nchunks; ++j) btx += 9u * (uint32_t)(p.Cout * p.bkc[j] * 2);

  if (warp == 1) {
    // ===================================================================== scheduler
    ptx::pdl_wait();
    if (ptx::elect_one()) {
      RingWriter rw; rw.init(&bars->ring);
      long long dbg_dep = 0; int dbg_tiles = 0;
      // one ticket = one tile row of one image (tiles_w neighbouring tiles: same dependencies, overlapping input patches)
      const int units = p.N * p.tiles_h;                         // tile rows of this launch
      const unsigned total_chunks = (unsigned)(cp.nconv * units);
      long long dbg_tk = 0, dbg_ring = 0;
      unsigned next = atomicAdd(&cp.ctrl[0], 1u);
      for (;;) {
        long long tq = cp.dbg ? clock64() : 0;
        const unsigned t = next;                                 // (first use of the atomic's result: waits for it)
        if (t >= total_chunks) { rw.acquire_slot(); rw.publish(kChainDone); break; }
        next = atomicAdd(&cp.ctrl[0], 1u);
        const int k = (int)t / units;
        const int u = (int)t - k * units;                        // unit = img * tiles_h + th
        const int img = u / p.tiles_h;
        const int th = u - img * p.tiles_h;
        if (cp.dbg) { const long long tn = clock64(); dbg_tk += tn - tq; tq = tn; }
        if (k > 0) {
          // the 10 x 18 input patch of an 8 x 16 tile touches the 3 x 3 neighbouring tiles of the same image: tile rows
          // th-1 .. th+1 of that image must be complete
          const int nb = 1 + (th > 0 ? 1 : 0) + (th < p.tiles_h - 1 ? 1 : 0);
          chain_wait_counter(cp.counters + (size_t)(k - 1) * cp.unit_stride + u, (unsigned)(nb * p.tiles_w));
          fence_proxy_async_all();   // the acquired generic-proxy stores before the TMA (async-proxy) reads issued downstream
          if (cp.dbg) { const long long tn = clock64(); dbg_dep += tn - tq; tq = tn; }
        }
        const uint32_t c0 = ((uint32_t)img << 16) | ((uint32_t)th << 8);
        for (int tw = 0; tw < p.tiles_w; ++tw) {
          rw.acquire_slot();
          rw.publish(((uint32_t)k << 28) | (uint32_t)(u * p.tiles_w + tw), c0 | (uint32_t)tw);
          ++dbg_tiles;
        }
        if (cp.dbg) dbg_ring += clock64() - tq;
      }
      if (cp.dbg) { cp.dbg[blockIdx.x * 16 + 8] = dbg_tk; cp.dbg[blockIdx.x * 16 + 9] = dbg_ring; }
      if (cp.dbg) { cp.dbg[blockIdx.x * 16 + 0] = dbg_dep; cp.dbg[blockIdx.x * 16 + 1] = dbg_tiles; }
    }
    __syncwarp();
  } else if (warp == 0) {
    // ===================================================================== TMA producer
    // Also owns the resident weights: conv 0's before the grid dependency resolves (constants), then a reload whenever
    // the ticket stream of this CTA moves on to another conv.
    if (ptx::elect_one()) {
      const uint32_t bfull = ptx::smem_u32(&bars->b_full);
      const uint32_t bempty = ptx::smem_u32(&bars->b_empty);
      auto load_weights = [&](int k) {
        ptx::mbar_expect_tx(bfull, btx);
        for (int j = 0; j < p.nchunks; ++j)
          for (int t = 0; t < 9; ++t)
            ptx::tma_load_2d(b_base + (uint32_t)(p.boff[j] + t * p.bblk[j]), &maps.b[k][p.mapi[j] ? 1 : 0], bfull,
                             t * p.Cin + p.c0[j], 0);
      };
      // The first conv this CTA works on is not known before its first ticket; conv 0 is the common case (always true
      // for the first gridDim tickets), so its weights are requested ahead of the grid dependency.
      int cur_k = 0;
      int nswitch = 0;                      // weight reloads so far
      load_weights(0);
      ptx::pdl_wait();
      RingReader rr; rr.init(&bars->ring);
      // Two issuers: tiles of parity w run through pipeline w (issuer w, epilogue warpgroup w, slots [w * ring, (w+1) * ring)).
      const int nw = n_issuers;
      const int ring = nw == 2 ? p.nslots / 2 : p.nslots;
      int L0 = 0, L1 = 0;                   // chunk loads issued into each pipeline's slot ring
      long long dbg_pring = 0, dbg_pslot = 0;
      for (int it = 0;; ++it) {
        long long tq = cp.dbg ? clock64() : 0;
        uint32_t coord;
        const uint32_t info = rr.next(coord);
        if (cp.dbg) dbg_pring += clock64() - tq;
        if (info == kChainDone) break;
        const int k = (int)(info >> 28);
        if (k != cur_k) {
          cur_k = k;
          // the MMAs of every earlier tile of this CTA (all issuers) must have retired before the overwrite
          ptx::mbar_wait(bempty, (uint32_t)(nswitch & 1));
          load_weights(k);
          ++nswitch;
        }
        const int w = nw == 2 ? (it & 1) : 0;
        const int sbase = w * ring;
        const int img = (int)(coord >> 16), th = (int)((coord >> 8) & 255u), tw = (int)(coord & 255u);
        if (cp.tail64) {
          // three slots: chunk 0 (64 channels, 128-byte rows) alternates between slots 0 and 2, chunk 1 (32 channels,
          // 64-byte rows) lives in slot 1 -- the next tile's chunk 0 loads while this tile's is still being multiplied
          const int s0 = (it & 1) ? 2 : 0;
          const uint32_t a1_off = 2u * (uint32_t)p.slot_bytes;                       // slot 1 sits behind the two wide slots
          ptx::mbar_wait(ptx::smem_u32(&bars->a_empty[s0]), (uint32_t)((it >> 1) & 1) ^ 1u);
          uint32_t full = ptx::smem_u32(&bars->a_full[s0]);
          ptx::mbar_expect_tx(full, (uint32_t)(kPatchRows * 128));
          ptx::tma_load_4d(a_base + (uint32_t)((s0 >> 1) * p.slot_bytes), &maps.a[k], full, p.c0[0], tw * kPatchTW - 1,
                           th * kPatchTH - 1, img);
          ptx::mbar_wait(ptx::smem_u32(&bars->a_empty[1]), (uint32_t)(it & 1) ^ 1u);
          full = ptx::smem_u32(&bars->a_full[1]);
          ptx::mbar_expect_tx(full, (uint32_t)(kPatchRows * 64));
          ptx::tma_load_4d(a_base + a1_off, &maps.a2[k], full, p.c0[1], tw * kPatchTW - 1, th * kPatchTH - 1, img);
          continue;
        }
        for (int j = 0; j < p.nchunks; ++j) {
          const int L = w ? L1++ : L0++;
          const int slot = sbase + L % ring;
          const uint32_t phase = (uint32_t)((L / ring) & 1);
          if (cp.dbg) tq = clock64();
          ptx::mbar_wait(ptx::smem_u32(&bars->a_empty[slot]), phase ^ 1u);
          if (cp.dbg) dbg_pslot += clock64() - tq;
          const uint32_t full = ptx::smem_u32(&bars->a_full[slot]);
          ptx::mbar_expect_tx(full, (uint32_t)(kPatchRows * 128));
          ptx::tma_load_4d(a_base + (uint32_t)(slot * p.slot_bytes), &maps.a[k], full, p.c0[j], tw * kPatchTW - 1,
                           th * kPatchTH - 1, img);
        }
      }
      if (cp.dbg) { cp.dbg[blockIdx.x * 16 + 10] = dbg_pring; cp.dbg[blockIdx.x * 16 + 11] = dbg_pslot; }
    }
    __syncwarp();
  } else if (warp == 2 || (warp == 3 && n_issuers == 2)) {
    // ===================================================================== MMA issuer(s)
    const int mw = warp - 2;
    if (ptx::elect_one()) {
      RingReader rr; rr.init(&bars->ring);
      const uint32_t idesc = ptx::umma_idesc_f16(128, p.Cout);
      const uint32_t bfull = ptx::smem_u32(&bars->b_full);
      const uint32_t bempty = ptx::smem_u32(&bars->b_empty);
      const int nw = n_issuers;
      const int ring = nw == 2 ? p.nslots / 2 : p.nslots;
      const int sbase = nw == 2 ? mw * ring : 0;
      int slot = 0;
      uint32_t phase = 0;
      bool ready = false;
      int cur_k = 0;
      int nswitch = 0;
      const bool dbg_on = cp.dbg != nullptr && mw == 0;
      long long dbg_wfull = 0, dbg_wtm = 0, dbg_wsw = 0;
      ptx::mbar_wait(bfull, 0);               // conv 0's weights
      // Lean issue loop (one or two channel chunks per tile, no timers): the chunk constants live in registers, the
      // descriptors of consecutive slots differ by a constant, and the barriers this issuer will need NEXT (the next
      // patch slot, the accumulator of its next tile) are probed before the MMAs of the current chunk are issued, so
      // their ~150 clk answers arrive under the MMAs.  The general loop below re-reads every chunk constant from the
      // constant bank and pays each barrier round trip serially: ~800 clk per 27-MMA tile (profiles/r02_s16_*.log:
      // a single issuer needed 2,220 clk per C = 48 tile, 82 clk per MMA, where the tensor pipe needs 44).
      if (cp.tail64) {
        // three-slot layout (see the producer): chunk 0 from slot 0 / 2 (tile parity), chunk 1 from the 64-byte-row slot 1
        const uint64_t a_hi = ptx::umma_desc_kmajor(0u, 128u, (uint32_t)kPatchPW * 128u);
        const uint64_t a_hi64 = ptx::umma_desc_kmajor(0u, 64u, (uint32_t)kPatchPW * 64u);
        const uint32_t enc_s0 = (a_base & 0x3FFFFu) >> 4, enc_s2 = ((a_base + (uint32_t)p.slot_bytes) & 0x3FFFFu) >> 4;
        const uint32_t enc_s1 = ((a_base + 2u * (uint32_t)p.slot_bytes) & 0x3FFFFu) >> 4;
        const uint32_t brow0 = (uint32_t)p.bkc[0] * 2u, brow1 = (uint32_t)p.bkc[1] * 2u;
        const uint64_t bd0 = ptx::umma_desc_kmajor(b_base + (uint32_t)p.boff[0], brow0, 8u * brow0);
        const uint64_t bd1 = ptx::umma_desc_kmajor(b_base + (uint32_t)p.boff[1], brow1, 8u * brow1);
        const uint32_t bs0 = (uint32_t)p.bblk[0] >> 4, bs1 = (uint32_t)p.bblk[1] >> 4;
        const uint32_t afull0 = ptx::smem_u32(&bars->a_full[0]), aempty0 = ptx::smem_u32(&bars->a_empty[0]);
        const uint32_t tfull0 = ptx::smem_u32(&bars->tmem_full[0]), tempty0 = ptx::smem_u32(&bars->tmem_empty[0]);
        const int nacc_mask = p.nacc - 1, nacc_log2 = p.nacc_log2;
        for (int it = 0;; ++it) {
          const uint32_t info = rr.next();
          if (info == kChainDone) break;
          const int k = (int)(info >> 28);
          if (k != cur_k) {                   // (see the general loop)
            cur_k = k;
            ptx::mma_commit(bempty);
            ++nswitch;
            ptx::mbar_wait(bfull, (uint32_t)(nswitch & 1));
            ptx::tc_fence_after_sync();
          }
          const int acc = it & nacc_mask;
          ptx::mbar_wait(tempty0 + 8u * (uint32_t)acc, (uint32_t)((it >> nacc_log2) & 1) ^ 1u);
          ptx::tc_fence_after_sync();
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.Cout);
          const int s0 = (it & 1) ? 2 : 0;
          ptx::mbar_wait(afull0 + 8u * (uint32_t)s0, (uint32_t)((it >> 1) & 1));
          ptx::tc_fence_after_sync();
          issue_taps<4, false>(d_tmem, a_hi | (uint64_t)(s0 ? enc_s2 : enc_s0), bd0, bs0, idesc, 0u);
          ptx::mma_commit(aempty0 + 8u * (uint32_t)s0);
          ptx::mbar_wait(afull0 + 8u, (uint32_t)(it & 1));
          ptx::tc_fence_after_sync();
          issue_taps<2, false, 4>(d_tmem, a_hi64 | (uint64_t)enc_s1, bd1, bs1, idesc, 1u);
          ptx::mma_commit(aempty0 + 8u);
          ptx::mma_commit(tfull0 + 8u * (uint32_t)acc);
        }
      } else
      if (p.nchunks <= 2 && cp.skip != 9) {      // (skip == 9: experiment, the general loop)
        const int nch = p.nchunks;
        const uint64_t a_hi = ptx::umma_desc_kmajor(0u, 128u, (uint32_t)kPatchPW * 128u);
        const uint32_t a_enc0 = ((a_base + (uint32_t)(sbase * p.slot_bytes)) & 0x3FFFFu) >> 4, a_enc_slot = (uint32_t)p.slot_bytes >> 4;
        const uint32_t brow0 = (uint32_t)p.bkc[0] * 2u, brow1 = (uint32_t)p.bkc[nch - 1] * 2u;
        const uint64_t bd0 = ptx::umma_desc_kmajor(b_base + (uint32_t)p.boff[0], brow0, 8u * brow0);
        const uint64_t bd1 = ptx::umma_desc_kmajor(b_base + (uint32_t)p.boff[nch - 1], brow1, 8u * brow1);
        const uint32_t bs0 = (uint32_t)p.bblk[0] >> 4, bs1 = (uint32_t)p.bblk[nch - 1] >> 4;
        const int nk0 = p.kreal[0] / 16, nk1 = p.kreal[nch - 1] / 16;
        const uint32_t afull0 = ptx::smem_u32(&bars->a_full[sbase]), aempty0 = ptx::smem_u32(&bars->a_empty[sbase]);
        const uint32_t tfull0 = ptx::smem_u32(&bars->tmem_full[0]), tempty0 = ptx::smem_u32(&bars->tmem_empty[0]);
        const int nacc_mask = p.nacc - 1, nacc_log2 = p.nacc_log2;
        uint32_t a_enc = a_enc0;
        bool tm_ready = false;                // probe of the accumulator of this issuer's next tile
        for (int it = 0;; ++it) {
          const uint32_t info = rr.next();
          if (info == kChainDone) break;
          const int k = (int)(info >> 28);
          if (k != cur_k) {                   // (see the general loop)
            cur_k = k;
            ptx::mma_commit(bempty);
            ++nswitch;
            ptx::mbar_wait(bfull, (uint32_t)(nswitch & 1));
            ptx::tc_fence_after_sync();
          }
          if (nw == 2 && (it & 1) != mw) continue;
          const int acc = it & nacc_mask;
          if (!tm_ready) ptx::mbar_wait(tempty0 + 8u * (uint32_t)acc, (uint32_t)((it >> nacc_log2) & 1) ^ 1u);
          ptx::tc_fence_after_sync();
          {
            const int nit = it + nw;
            tm_ready = ptx::mbar_test_wait(tempty0 + 8u * (uint32_t)(nit & nacc_mask), (uint32_t)((nit >> nacc_log2) & 1) ^ 1u);
          }
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.Cout);
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            if (j < nch) {
              if (!ready) ptx::mbar_wait(afull0 + 8u * (uint32_t)slot, phase);
              ptx::tc_fence_after_sync();
              const uint64_t a0 = a_hi | (uint64_t)a_enc;
              const uint32_t cur_empty = aempty0 + 8u * (uint32_t)slot;
              ++slot; a_enc += a_enc_slot;
              if (slot == ring) { slot = 0; phase ^= 1u; a_enc = a_enc0; }
              ready = ptx::mbar_test_wait(afull0 + 8u * (uint32_t)slot, phase);
              switch (j == 0 ? nk0 : nk1) {
                case 4: issue_taps<4, false>(d_tmem, a0, j == 0 ? bd0 : bd1, j == 0 ? bs0 : bs1, idesc, (uint32_t)j); break;
                case 3: issue_taps<3, false>(d_tmem, a0, j == 0 ? bd0 : bd1, j == 0 ? bs0 : bs1, idesc, (uint32_t)j); break;
                case 2: issue_taps<2, false>(d_tmem, a0, j == 0 ? bd0 : bd1, j == 0 ? bs0 : bs1, idesc, (uint32_t)j); break;
                default: issue_taps<1, false>(d_tmem, a0, j == 0 ? bd0 : bd1, j == 0 ? bs0 : bs1, idesc, (uint32_t)j); break;
              }
              ptx::mma_commit(cur_empty);
            }
          }
          ptx::mma_commit(tfull0 + 8u * (uint32_t)acc);
        }
      } else
      for (int it = 0;; ++it) {
        const uint32_t info = rr.next();
        if (info == kChainDone) {
          if (dbg_on) { cp.dbg[blockIdx.x * 16 + 5] = dbg_wfull; cp.dbg[blockIdx.x * 16 + 6] = dbg_wtm; cp.dbg[blockIdx.x * 16 + 7] = dbg_wsw; }
          break;
        }
        const int k = (int)(info >> 28);
        if (k != cur_k) {
          const long long tsw = dbg_on ? clock64() : 0;
          // every issuer walks every ring entry, so both see every conv switch of this CTA, in order: retire the
          // MMAs that read the old weights, then wait for the new set (each b_full phase is waited exactly once)
          cur_k = k;
          ptx::mma_commit(bempty);
          ++nswitch;
          ptx::mbar_wait(bfull, (uint32_t)(nswitch & 1));
          ptx::tc_fence_after_sync();
          if (dbg_on) dbg_wsw += clock64() - tsw;
        }
        if (nw == 2 && (it & 1) != mw) continue;
        const int acc = it & (p.nacc - 1);
        const uint32_t acc_phase = (uint32_t)((it >> p.nacc_log2) & 1);
        long long tq = dbg_on ? clock64() : 0;
        ptx::mbar_wait(ptx::smem_u32(&bars->tmem_empty[acc]), acc_phase ^ 1u);
        if (dbg_on) { const long long t = clock64(); dbg_wtm += t - tq; }
        ptx::tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.Cout);
        for (int j = 0; j < p.nchunks; ++j) {
          if (dbg_on) tq = clock64();
          if (!ready) ptx::mbar_wait(ptx::smem_u32(&bars->a_full[sbase + slot]), phase);
          if (dbg_on) dbg_wfull += clock64() - tq;
          ptx::tc_fence_after_sync();
          const uint32_t a_slot = a_base + (uint32_t)((sbase + slot) * p.slot_bytes);
          const uint32_t brow = (uint32_t)p.bkc[j] * 2u;
          const int nk = p.kreal[j] / 16;
          int nslot = slot + 1;
          uint32_t nphase = phase;
          if (nslot == ring) { nslot = 0; nphase ^= 1u; }
          const bool nready = ptx::mbar_test_wait(ptx::smem_u32(&bars->a_full[sbase + nslot]), nphase);
          const uint64_t a0 = ptx::umma_desc_kmajor(a_slot, 128u, (uint32_t)kPatchPW * 128u);
          const uint64_t b0 = ptx::umma_desc_kmajor(b_base + (uint32_t)p.boff[j], brow, 8u * brow);
          const uint32_t bstep = (uint32_t)p.bblk[j] >> 4;
          const uint32_t first = (uint32_t)(j != 0);
          switch (nk) {
            case 4: issue_taps<4, false>(d_tmem, a0, b0, bstep, idesc, first); break;
            case 3: issue_taps<3, false>(d_tmem, a0, b0, bstep, idesc, first); break;
            case 2: issue_taps<2, false>(d_tmem, a0, b0, bstep, idesc, first); break;
            default: issue_taps<1, false>(d_tmem, a0, b0, bstep, idesc, first); break;
          }
          ptx::mma_commit(ptx::smem_u32(&bars->a_empty[sbase + slot]));
          slot = nslot; phase = nphase; ready = nready;
        }
        ptx::mma_commit(ptx::smem_u32(&bars->tmem_full[acc]));
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================================================================== epilogue (four warpgroups, tile it -> it & 3)
    const int g = (warp - 4) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int dh = row >> 3, dw = row & 7;
    const bool leader = (q == 0) && (lane == 0);
    ptx::pdl_wait();
    RingReader rr; rr.init(&bars->ring);
    PendingPublish pend; pend.clear();
    long long dbg_work = 0, dbg_wait = 0;
    for (int it = 0;; ++it) {
      if (!rr.ready()) pend.flush(1 + g, leader);      // about to sleep on the ring: publish first (see PendingPublish)
      uint32_t coord;
      const uint32_t info = rr.next(coord);
      if (info == kChainDone) break;
      if ((it & (kCPEpiGroups - 1)) != g) continue;
      pend.flush(1 + g, leader);                        // the previous tile of this warpgroup
      const int k = (int)(info >> 28);
      const ChainConv& cv = cp.conv[k];
      const int acc = it & (p.nacc - 1);
      const uint32_t acc_phase = (uint32_t)((it >> p.nacc_log2) & 1);
      const int img = (int)(coord >> 16), th = (int)((coord >> 8) & 255u), tw = (int)(coord & 255u);
      const int oh = th * kPatchTH + dh, ow = tw * kPatchTW + dw;
      EpiRow e;
      e.s_scale = nullptr; e.s_bias = nullptr; e.residual = cv.residual; e.out = cv.out;      // BN constants: cp.sb[k] (constant bank)
      e.row_off = (((size_t)img * p.H + oh) * p.W + ow) * p.Cout;
      e.ch0 = 0; e.ncols = p.Cout; e.relu = cv.relu; e.out_f32 = 0; e.valid = oh < p.H && ow < p.W;
      U256 rres[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int i = 0; i < 8; ++i) rres[j].w[i] = 0u;
        // in flight while the MMAs of this tile finish
        if (e.residual != nullptr && e.valid && 16 * j < e.ncols && cp.skip != 1) rres[j] = ldg256_cg(e.residual + e.row_off + 16 * j);
      }
      long long tq = cp.dbg ? clock64() : 0;
      ptx::mbar_wait(ptx::smem_u32(&bars->tmem_full[acc]), acc_phase);
      if (cp.dbg) { const long long t = clock64(); dbg_wait += t - tq; tq = t; }
      ptx::tc_fence_after_sync();
      if (cp.skip == 2) {            // experiment (results invalid): accumulator loads only
        uint32_t x = 0;
        for (int c = 0; c < e.ncols; c += 16) {
          uint32_t v[16];
          ptx::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.Cout + c), v);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) x ^= v[i];
        }
        if (x == 0x12345679u && e.valid) reinterpret_cast<__half*>(e.out)[e.row_off] = __float2half(0.f);
      } else if (cp.skip == 3) {     // experiment (results invalid): residual loads and output stores only
        if (e.valid)
          for (int j = 0; j < 4; ++j) if (16 * j < e.ncols) stg256(reinterpret_cast<__half*>(e.out) + e.row_off + 16 * j, rres[j]);
      } else if (cp.skip != 1)       // (skip == 1: experiment, no epilogue work -- results invalid)
        chain_store_row_lean_c(e, tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.Cout), rres, cp.sb[k]);
      ptx::tc_fence_before_sync();
      ptx::mbar_arrive(ptx::smem_u32(&bars->tmem_empty[acc]));
      pend.set(cp.counters + (size_t)k * cp.unit_stride + img * p.tiles_h + th, th > 0, th < p.tiles_h - 1);
      if (cp.dbg) dbg_work += clock64() - tq;
    }
    pend.flush(1 + g, leader);
    if (cp.dbg && threadIdx.x == 128) { cp.dbg[blockIdx.x * 16 + 3] = dbg_wait; cp.dbg[blockIdx.x * 16 + 4] = dbg_work; }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
  if (cp.dbg && threadIdx.x == 0) cp.dbg[blockIdx.x * 16 + 2] = clock64() - t_begin;
  chain_exit(cp.ctrl, cp.counters, cp.nconv, p.N * p.tiles_h, cp.unit_stride, &bars->pad);
}

// ---------------------------------------------------------------------------------------------------------------------
cudaError_t conv_chain_set_attributes(int max_smem) {
  cudaError_t e = cudaFuncSetAttribute(conv_chain_igemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_chain_igemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_chain_patch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
  return e;
}

template <typename K, typename M, typename P>
static cudaError_t launch_chain(K kernel, const M& maps, const P& p, int threads, int smem_bytes, int grid, cudaStream_t st,
                                int cluster = 1) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = (size_t)smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  int na = 0;
  if (p.pdl) {
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (cluster > 1) {
    at[na].id = cudaLaunchAttributeClusterDimension;
    at[na].val.clusterDim.x = (unsigned)cluster; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = at; cfg.numAttrs = (unsigned)na;
  return cudaLaunchKernelEx(&cfg, kernel, maps, p);
}

cudaError_t launch_chain_igemm(const ChainIgemmMaps& maps, const ChainIgemmParams& p, int smem_bytes, int grid, cudaStream_t st) {
  if (p.pair) return launch_chain(conv_chain_igemm_kernel<true>, maps, p, kCIThreads, smem_bytes, grid / 2 * 2, st, 2);
  return launch_chain(conv_chain_igemm_kernel<false>, maps, p, kCIThreads, smem_bytes, grid, st);
}
cudaError_t launch_chain_patch(const ChainPatchMaps& maps, const ChainPatchParams& p, int smem_bytes, int grid, cudaStream_t st) {
  return launch_chain(conv_chain_patch_kernel, maps, p, kCPThreads, smem_bytes, grid, st);
}

}  // namespace hrnet
