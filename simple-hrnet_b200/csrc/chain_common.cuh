// Device-side building blocks shared by the persistent ticket-scheduled kernels (conv_chain.cu: branch chains,
// conv_xunit.cu: exchange units): the descriptor ring between the scheduler warp and the other roles, dependency
// counters in global memory (acquire / deferred release), 256-bit global accesses and the thread-per-row epilogue.
#pragma once
#include "conv_igemm_body.cuh"

namespace hrnet {

constexpr int kChainRing = 8;
constexpr uint32_t kChainDone = 0xffffffffu;

struct ChainRing {
  uint64_t full[kChainRing];
  uint64_t empty[kChainRing];
  uint32_t info[kChainRing];     // conv | tile (kChainDone after the last one)
  uint32_t coord[kChainRing];    // the tile's coordinates, decoded once by the scheduler (no divisions in the other roles)
  uint32_t coord2[kChainRing];   // im2col chain, two M-tiles per ticket: coordinates of the second tile
};

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_add_relaxed_gpu(unsigned* p, unsigned v) {
  asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// 256-bit global accesses (sm_100: LDG / STG.E.256).  The thread-per-row epilogue is bound by the load/store unit's sector
// rate -- every lane touches its own 32-byte sector, ~1 sector per clock per SM (profiles/r02_s2_chain_sched_v2_roles.log:
// four epilogue warpgroups took exactly as long per tile as two) -- so moving a whole sector per access instead of half
// of one halves the epilogue's load/store time.  The load is GPU-coherent (.cg: L2, never a stale L1 line).
struct __align__(32) U256 { uint32_t w[8]; };
__device__ __forceinline__ U256 ldg256_cg(const void* p) {
  U256 v;
  asm volatile("ld.global.cg.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v.w[0]), "=r"(v.w[1]), "=r"(v.w[2]), "=r"(v.w[3]), "=r"(v.w[4]), "=r"(v.w[5]), "=r"(v.w[6]), "=r"(v.w[7])
               : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void stg256(void* p, const U256& v) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"l"(p), "r"(v.w[0]), "r"(v.w[1]), "r"(v.w[2]), "r"(v.w[3]), "r"(v.w[4]), "r"(v.w[5]), "r"(v.w[6]), "r"(v.w[7]) : "memory");
}
// 16 accumulator columns starting at tile column c: the arithmetic of epi_cols16 (epilogue.cuh) with the residual arriving
// as and the result leaving as ONE 32-byte access.  fp16 outputs only.
__device__ __forceinline__ void chain_cols16(const uint32_t (&v)[16], const U256& r, const EpiRow& e, int c) {
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
    const __half2* h = reinterpret_cast<const __half2*>(&r);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float2 f = __half22float2(h[i]);
      y[2 * i] += f.x; y[2 * i + 1] += f.y;
    }
  }
  if (e.relu) {
#pragma unroll
    for (int i = 0; i < 16; ++i) y[i] = fmaxf(y[i], 0.f);
  }
  U256 o;
  __half2* oh2 = reinterpret_cast<__half2*>(&o);
#pragma unroll
  for (int i = 0; i < 8; ++i) oh2[i] = __floats2half2_rn(y[2 * i], y[2 * i + 1]);
  stg256(reinterpret_cast<__half*>(e.out) + e.row_off + c, o);
}

// The same with the BN constants read from the KERNEL PARAMETERS (constant bank): `sb` must be a reference into the
// __grid_constant__ parameter block ((scale, bias) per absolute output channel), so every access is an LDC -- no
// shared-memory or L1 wavefront.  Why: the tensor core's operand reads, LDS and LDG / STG share the SM's L1 data pipe,
// and a warp-wide LDS.128 costs four wavefronts even when all lanes read the same address: the constants of a 128 x 48
// tile cost 452 wavefronts next to the 1,188 of its MMAs, and the pipe was 84 % busy -- the bound of the halo-patch
// chains (profiles/r02_s11_ncu_chain_kernels.txt; without any epilogue work the C = 48 chain runs 1,345 instead of
// 2,150 clk per tile, profiles/r02_s18_*.log).  Same arithmetic, same results.
template <typename SB>
__device__ __forceinline__ void chain_cols16_c(const uint32_t (&v)[16], const U256& r, const EpiRow& e, int c, const SB& sb) {
  float y[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float2 s = sb[e.ch0 + c + i];
    y[i] = fmaf(__uint_as_float(v[i]), s.x, s.y);
  }
  if (e.residual != nullptr) {
    const __half2* h = reinterpret_cast<const __half2*>(&r);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float2 f = __half22float2(h[i]);
      y[2 * i] += f.x; y[2 * i + 1] += f.y;
    }
  }
  if (e.relu) {
#pragma unroll
    for (int i = 0; i < 16; ++i) y[i] = fmaxf(y[i], 0.f);
  }
  U256 o;
  __half2* oh2 = reinterpret_cast<__half2*>(&o);
#pragma unroll
  for (int i = 0; i < 8; ++i) oh2[i] = __floats2half2_rn(y[2 * i], y[2 * i + 1]);
  stg256(reinterpret_cast<__half*>(e.out) + e.row_off + c, o);
}

// generic-proxy accesses before / async-proxy (TMA) accesses after, all state spaces
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ uint32_t lds_volatile_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts_volatile_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// A dependency that never arrives would hang the GPU: after ~2 s of polling the CTA traps (launch failure on the host).
static __device__ __noinline__ void chain_wait_counter_slow(const unsigned* c, unsigned want) {
  const long long t0 = clock64();
  while (ld_acquire_gpu(c) < want) {
    __nanosleep(32);
    if (clock64() - t0 > 4000000000ll) {
      printf("hrnet_b200: chain dependency timeout (block %d counter %p want %u have %u)\n", (int)blockIdx.x, (const void*)c,
             want, ld_acquire_gpu(c));
      __trap();
    }
  }
}
// counters[k-1][u] has reached `want`: every tile of units u-1 .. u+1 of the previous conv is stored and visible
__device__ __forceinline__ void chain_wait_counter(const unsigned* c, unsigned want) {
  if (ld_acquire_gpu(c) < want) chain_wait_counter_slow(c, want);
}
// Ring consumer: every consuming thread walks every entry (tile descriptor or kChainDone), in order.
// CTA pairs (conv_chain.cu, cta_group::2): the LEADER's scheduler fills both CTAs' rings.  A consumer of the peer CTA
// (`peer`) waits on its own CTA's `full` barrier with cluster-scope acquire (the descriptor words were stored remotely)
// and releases the slot on the leader's `empty` barrier.
struct RingReader {
  uint32_t full0, empty0, info0, coord0;
  int i;
  bool peer;
  __device__ __forceinline__ void init(ChainRing* r, bool peer_cta = false) {
    full0 = ptx::smem_u32(&r->full[0]); empty0 = ptx::smem_u32(&r->empty[0]); info0 = ptx::smem_u32(&r->info[0]);
    coord0 = ptx::smem_u32(&r->coord[0]);
    i = 0;
    peer = peer_cta;
  }
  __device__ __forceinline__ uint32_t next(uint32_t& coord, uint32_t& coord2) {
    const uint32_t slot = (uint32_t)(i % kChainRing);
    const uint32_t ph = (uint32_t)((i / kChainRing) & 1);
    if (peer) ptx::mbar_wait_cluster(full0 + 8u * slot, ph);
    else ptx::mbar_wait(full0 + 8u * slot, ph);
    const uint32_t v = lds_volatile_u32(info0 + 4u * slot);
    coord = lds_volatile_u32(coord0 + 4u * slot);
    coord2 = lds_volatile_u32(coord0 + 4u * (kChainRing + slot));
    if (peer) ptx::mbar_arrive_cluster(empty0 + 8u * slot, 0);
    else ptx::mbar_arrive(empty0 + 8u * slot);
    ++i;
    return v;
  }
  // true when the next entry is already published (next() would not block)
  __device__ __forceinline__ bool ready() const {      // (a hint only: no acquire semantics needed)
    const uint32_t slot = (uint32_t)(i % kChainRing);
    const uint32_t ph = (uint32_t)((i / kChainRing) & 1);
    return ptx::mbar_test_wait(full0 + 8u * slot, ph);
  }
  __device__ __forceinline__ uint32_t next(uint32_t& coord) { uint32_t c2; return next(coord, c2); }
  __device__ __forceinline__ uint32_t next() { uint32_t c, c2; return next(c, c2); }
};

// Ring producer side of the scheduler thread.
struct RingWriter {
  uint32_t full0, empty0, info0, coord0;
  int i;
  bool pair;      // also fill the ring of CTA 1 of the cluster (same shared-memory offsets)
  __device__ __forceinline__ void init(ChainRing* r, bool pair_mode = false) {
    full0 = ptx::smem_u32(&r->full[0]); empty0 = ptx::smem_u32(&r->empty[0]); info0 = ptx::smem_u32(&r->info[0]);
    coord0 = ptx::smem_u32(&r->coord[0]);
    i = 0;
    pair = pair_mode;
  }
  __device__ __forceinline__ void acquire_slot() {
    const uint32_t slot = (uint32_t)(i % kChainRing);
    const uint32_t ph = (uint32_t)((i / kChainRing) & 1);
    ptx::mbar_wait(empty0 + 8u * slot, ph ^ 1u);
  }
  __device__ __forceinline__ void publish(uint32_t v, uint32_t coord = 0u, uint32_t coord2 = 0u) {
    const uint32_t slot = (uint32_t)(i % kChainRing);
    sts_volatile_u32(info0 + 4u * slot, v);
    sts_volatile_u32(coord0 + 4u * slot, coord);
    sts_volatile_u32(coord0 + 4u * (kChainRing + slot), coord2);
    if (pair) {
      ptx::st_shared_cluster_u32(ptx::mapa_cluster(info0 + 4u * slot, 1), v);
      ptx::st_shared_cluster_u32(ptx::mapa_cluster(coord0 + 4u * slot, 1), coord);
      ptx::st_shared_cluster_u32(ptx::mapa_cluster(coord0 + 4u * (kChainRing + slot), 1), coord2);
      ptx::mbar_arrive_cluster(full0 + 8u * slot, 1);   // release at cluster scope
    }
    ptx::mbar_arrive(full0 + 8u * slot);     // release at CTA scope: the descriptor is visible to the waiters
    ++i;
  }
};

// Block-wide, after the final __syncthreads: the last CTA of the launch clears the unit counters it and the others used
// (`nconv` x `units`, row pitch `stride`) and re-arms the ticket.  `flag` is a shared-memory word.
__device__ __forceinline__ void chain_exit(unsigned* ctrl, unsigned* counters, int nconv, int units, int stride, uint32_t* flag) {
  if (threadIdx.x == 0) {
    __threadfence();
    *flag = atomicAdd(&ctrl[1], 1u) == gridDim.x - 1 ? 1u : 0u;
  }
  __syncthreads();
  if (*flag == 0u) return;
  __threadfence();
  for (int k = 0; k < nconv; ++k)
    for (int u = threadIdx.x; u < units; u += blockDim.x) counters[(size_t)k * stride + u] = 0u;
  if (threadIdx.x == 0) { ctrl[0] = 0u; ctrl[1] = 0u; }
}

// Epilogue of one row of the im2col chain: the residual of the next 64 channels is in flight while the current 64 are
// converted (4 x 256-bit loads; written earlier in this launch by another SM, hence the coherent loads).
__device__ __forceinline__ void chain_load_residual(U256 (&r)[4], const EpiRow& e, int c_begin) {
  if (e.residual == nullptr || !e.valid) return;
  const __half* rp = e.residual + e.row_off + c_begin;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (c_begin + 16 * i < e.ncols) r[i] = ldg256_cg(rp + 16 * i);
}
__device__ __forceinline__ void chain_store_row(U256 (&r)[4], const EpiRow& e, uint32_t t_row) {
  for (int c64 = 0; c64 < e.ncols; c64 += 64) {
    U256 cur[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) cur[i] = r[i];
    if (c64 + 64 < e.ncols) chain_load_residual(r, e, c64 + 64);
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
          chain_cols16(v0, cur[2 * h], e, c);
          if (two) chain_cols16(v1, cur[2 * h + 1], e, c + 16);
        }
      }
    }
  }
}

template <typename SB>
__device__ __forceinline__ void chain_store_row_c(U256 (&r)[4], const EpiRow& e, uint32_t t_row, const SB& sb) {
  for (int c64 = 0; c64 < e.ncols; c64 += 64) {
    U256 cur[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) cur[i] = r[i];
    if (c64 + 64 < e.ncols) chain_load_residual(r, e, c64 + 64);
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
          chain_cols16_c(v0, cur[2 * h], e, c, sb);
          if (two) chain_cols16_c(v1, cur[2 * h + 1], e, c + 16, sb);
        }
      }
    }
  }
}
template <typename SB>
__device__ __forceinline__ void chain_store_row_lean_c(const EpiRow& e, uint32_t t_row, U256 (&r)[4], const SB& sb) {
  const bool has_res = e.residual != nullptr && e.valid;
  const __half* rp = e.residual + e.row_off;
  for (int c64 = 0; c64 < e.ncols; c64 += 64) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c64 + 16 * j;
      if (c < e.ncols) {                         // warp-uniform
        uint32_t v[16];
        ptx::tmem_ld16(t_row + (uint32_t)c, v);
        ptx::tmem_ld_wait();
        if (e.valid) chain_cols16_c(v, r[j], e, c, sb);
        if (has_res && c + 64 < e.ncols) r[j] = ldg256_cg(rp + c + 64);
      }
    }
  }
}

// Register-lean variant for the 20-warp halo-patch chain (96 registers per thread): 16 columns per step.  r[0..3] hold the
// residual of columns [0, 64), requested before the wait on the accumulator barrier; for wider tiles r[j] is re-filled
// with the columns 64 further on right after it has been consumed (four steps of prefetch distance).  Same arithmetic as
// epi_cols16, so results do not change.
__device__ __forceinline__ void chain_store_row_lean(const EpiRow& e, uint32_t t_row, U256 (&r)[4]) {
  const bool has_res = e.residual != nullptr && e.valid;
  const __half* rp = e.residual + e.row_off;
  for (int c64 = 0; c64 < e.ncols; c64 += 64) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c64 + 16 * j;
      if (c < e.ncols) {                         // warp-uniform
        uint32_t v[16];
        ptx::tmem_ld16(t_row + (uint32_t)c, v);
        ptx::tmem_ld_wait();
        if (e.valid) chain_cols16(v, r[j], e, c);
        if (has_res && c + 64 < e.ncols) r[j] = ldg256_cg(rp + c + 64);
      }
    }
  }
}

// Publishing a finished tile of unit u: make the warpgroup's stores visible GPU-wide (__threadfence by its leader after a
// warpgroup barrier) and count the tile in the neighbourhood counters of units u-1 (if has_lo), u and u+1 (if has_hi).
// The fence waits for the leader's own outstanding stores (an L2 round trip) and the other 127 threads would meet it
// again at the next tile's barrier, so the publication is DEFERRED: a warpgroup publishes tile t when it starts its
// next tile (the stores have long landed then) -- unless the ring is empty, in which case it publishes at once: a
// warpgroup never sleeps on an unpublished tile (another CTA's scheduler may be waiting for exactly that counter).
struct PendingPublish {
  unsigned* counter;     // nullptr: nothing pending
  bool has_lo, has_hi;
  __device__ __forceinline__ void clear() { counter = nullptr; has_lo = has_hi = false; }
  __device__ __forceinline__ void set(unsigned* c, bool lo, bool hi) { counter = c; has_lo = lo; has_hi = hi; }
  // all 128 threads of the warpgroup call this (warpgroup-uniform state)
  __device__ __forceinline__ void flush(int bar_id, bool leader) {
    if (counter == nullptr) return;
    ptx::bar_sync(bar_id, 128);                  // all four warps of the warpgroup have issued the tile's stores
    if (leader) {
      fence_proxy_async_all();                   // generic-proxy stores before later async-proxy (TMA) reads
      __threadfence();
      red_add_relaxed_gpu(counter, 1u);
      if (has_lo) red_add_relaxed_gpu(counter - 1, 1u);
      if (has_hi) red_add_relaxed_gpu(counter + 1, 1u);
    }
    counter = nullptr;
  }
};

}  // namespace hrnet
