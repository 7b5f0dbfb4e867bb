// Stem conv1 (models_/hrnet.py:158-160: 3x3 stride 2, 3 -> 64, + BN + ReLU) on the tensor cores.
//
// The SIMT version spent 885 us at N=64 (profiles/r01_layer_breakdown_v3.txt): 1728 FMAs per output pixel with every
// input re-loaded by four threads.  Here a CTA turns 128 output pixels into one small GEMM
//     D[128 x 64] = A[128 x K] * W[64 x K]^T,   K = 27 taps*channels padded to 32
// whose A rows the threads build straight from the NCHW fp32 input (one thread = one pixel = 27 loads).
// To keep fp32-conv accuracy on fp16 tensor cores both operands are split x = hi + lo (two fp16 each):
//     D = A_hi W_hi^T + A_lo W_hi^T + A_hi W_lo^T          (the lo*lo term is ~2^-22 relative and dropped)
// laid out as two K-major SWIZZLE_128B k-blocks: block 0 = [hi | lo] x [W_hi | W_hi] (K = 64), block 1 = [hi] x [W_lo]
// (K = 32) -> 6 tcgen05.mma of M=128, N=64, K=16, fp32 accumulate in TMEM.  The operands are written to shared
// memory by ordinary stores (generic proxy), so a fence.proxy.async precedes the MMAs.
// Persistent CTAs (4 per SM, ~50 KB shared memory and 64 TMEM columns each, so the others' loads hide one's latency):
// barrier / TMEM setup and the hi / lo split of the weights happen once per CTA, then it loops over its tiles
// (one tile per CTA cost 176 us at N=64: 13824 CTA prologues; profiles/r01_op_roofline_v9.txt).
#include <algorithm>

#include "hrnet_internal.h"
#include "ptx.cuh"

namespace hrnet {

__device__ __forceinline__ void split_pack8(const float* x, uint4& hi, uint4& lo) {
  __half2* h = reinterpret_cast<__half2*>(&hi);
  __half2* l = reinterpret_cast<__half2*>(&lo);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half a = __float2half_rn(x[2 * i]), b = __float2half_rn(x[2 * i + 1]);
    h[i] = __halves2half2(a, b);
    l[i] = __floats2half2_rn(x[2 * i] - __half2float(a), x[2 * i + 1] - __half2float(b));
  }
}

// row-major [rows][64 halves] K-major SWIZZLE_128B: logical 16-byte chunk c of row m lives at chunk c ^ (m & 7)
__device__ __forceinline__ void st_chunk(uint8_t* blk, int row, int chunk, const uint4& v) {
  *reinterpret_cast<uint4*>(blk + row * 128 + ((chunk ^ (row & 7)) << 4)) = v;
}

// kU8 = true: the input is the reference's *pre-transform* image batch, NHWC BGR uint8 at network resolution, and the
// kernel applies cvtColor(BGR2RGB) + ToTensor (/255) + Normalize(mean, std) (SimpleHRNet.py:222,149-153) on the fly with
// the same fp32 operation order (IEEE div / sub / div), so the result is bit-identical to feeding the host-normalised
// fp32 tensor -- while the host->device copy shrinks 4x (3 B instead of 12 B per input pixel).
template <bool kU8>
__global__ void __launch_bounds__(128, 4)      // four CTAs per SM: 128 registers per thread
stem_conv3x3s2_tc_kernel(const void* __restrict__ in_any, const float* __restrict__ w, const float* __restrict__ scale,
                         const float* __restrict__ bias, __half* __restrict__ out, int N, int H, int W) {
  const float* in = static_cast<const float*>(in_any);
  const uint8_t* in8 = static_cast<const uint8_t*>(in_any);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (sbase - ptx::smem_u32(smem_raw));
  uint8_t* A0 = sm;                 // 128 x 128 B  [hi | lo]
  uint8_t* A1 = sm + 16384;         // 128 x 128 B  [hi | - ]
  uint8_t* B0 = sm + 32768;         //  64 x 128 B  [W_hi | W_hi]
  uint8_t* B1 = sm + 40960;         //  64 x 128 B  [W_lo | - ]
  float* s_scale = reinterpret_cast<float*>(sm + 49152);
  float* s_bias = s_scale + 64;
  uint64_t* bar = reinterpret_cast<uint64_t*>(s_bias + 64);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bar + 1);

  const int warp = ptx::warp_idx_uniform();
  const int tid = threadIdx.x;
  if (tid == 0) { ptx::mbar_init(ptx::smem_u32(bar), 1); ptx::fence_mbar_init(); }
  if (warp == 0) ptx::tmem_alloc(ptx::smem_u32(tslot), 64);
  if (tid < 64) { s_scale[tid] = scale[tid]; s_bias[tid] = bias[tid]; }

  const int OH = H / 2, OW = W / 2;
  const long total = (long)N * OH * OW;

  // ---- B rows: one output channel per thread (weights [co][r][s][ci] fp32 -> K index (r*3+s)*3+ci)
  if (tid < 64) {
    float x[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) x[i] = i < 27 ? __ldg(w + tid * 27 + i) : 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint4 hi, lo;
      split_pack8(x + 8 * c, hi, lo);
      st_chunk(B0, tid, c, hi);
      st_chunk(B0, tid, 4 + c, hi);
      st_chunk(B1, tid, c, lo);
    }
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem = *tslot;
  uint32_t phase = 0;
  const long ntiles = (total + 127) / 128;
  // The 27 scattered input loads of a pixel are latency, not bandwidth: the NEXT tile's are issued right after this
  // tile's MMAs and land while the MMA round trip and the epilogue run (software pipeline, 32 more registers).
  auto gather = [&](long tile, float (&x)[32]) {
    const long pix = tile * 128 + tid;
#pragma unroll
    for (int i = 0; i < 32; ++i) x[i] = 0.f;
    if (tile < ntiles && pix < total) {
      const int n = (int)(pix / (OH * OW));
      const int rem = (int)(pix - (long)n * OH * OW);
      const int oh = rem / OW, ow = rem - oh * OW;
      const float* ip = in + (size_t)n * 3 * H * W;
      const uint8_t* ip8 = in8 + (size_t)n * H * W * 3;
      const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};   // RGB, SimpleHRNet.py:152
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int ih = oh * 2 - 1 + r;
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const int iw = ow * 2 - 1 + s;
          const bool ok = ih >= 0 && ih < H && iw >= 0 && iw < W;
#pragma unroll
          for (int ci = 0; ci < 3; ++ci) {
            float v = 0.f;    // zero padding applies to the normalised tensor
            if (ok) {
              if constexpr (kU8) {
                const float u = (float)__ldg(ip8 + ((size_t)ih * W + iw) * 3 + (2 - ci));   // BGR -> RGB
                v = __fdiv_rn(__fsub_rn(__fdiv_rn(u, 255.f), mean[ci]), stdv[ci]);
              } else {
                v = __ldg(ip + ((size_t)ci * H + ih) * W + iw);
              }
            }
            x[(r * 3 + s) * 3 + ci] = v;
          }
        }
      }
    }
  };
  float xc[32];
  gather((long)blockIdx.x, xc);
  for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long pix = tile * 128 + tid;
    const bool valid = pix < total;
    // ---- A rows: one output pixel per thread
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint4 hi, lo;
      split_pack8(xc + 8 * c, hi, lo);
      st_chunk(A0, tid, c, hi);
      st_chunk(A0, tid, 4 + c, lo);
      st_chunk(A1, tid, c, hi);
    }
    ptx::fence_proxy_async_smem();     // generic-proxy stores -> visible to the tensor core (async proxy)
    ptx::tc_fence_before_sync();
    __syncthreads();
    ptx::tc_fence_after_sync();
    if (warp == 0) {
      if (ptx::elect_one()) {
        const uint32_t idesc = ptx::umma_idesc_f16(128, 64);
        const uint64_t a0 = ptx::umma_desc_kmajor(sbase, 128u, 1024u);
        const uint64_t a1 = ptx::umma_desc_kmajor(sbase + 16384u, 128u, 1024u);
        const uint64_t b0 = ptx::umma_desc_kmajor(sbase + 32768u, 128u, 1024u);
        const uint64_t b1 = ptx::umma_desc_kmajor(sbase + 40960u, 128u, 1024u);
  #pragma unroll
        for (int k = 0; k < 4; ++k) ptx::mma_f16_ss(tmem, a0 + (uint64_t)(2 * k), b0 + (uint64_t)(2 * k), idesc, (uint32_t)(k != 0));
  #pragma unroll
        for (int k = 0; k < 2; ++k) ptx::mma_f16_ss(tmem, a1 + (uint64_t)(2 * k), b1 + (uint64_t)(2 * k), idesc, 1u);
        ptx::mma_commit(ptx::smem_u32(bar));
      }
      __syncwarp();
    }
    gather(tile + gridDim.x, xc);      // the next tile's inputs: in flight during the MMA round trip and the epilogue
    ptx::mbar_wait(ptx::smem_u32(bar), phase);
    phase ^= 1u;
    ptx::tc_fence_after_sync();

    // ---- epilogue: thread = accumulator row = output pixel; 64 channels = 128 contiguous bytes
    const uint32_t t_row = tmem + ((uint32_t)(warp * 32) << 16);
    uint4 o[8];
    __half2* oh2 = reinterpret_cast<__half2*>(o);
  #pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint32_t v0[16], v1[16];
      ptx::tmem_ld16(t_row + (uint32_t)(32 * h), v0);
      ptx::tmem_ld16(t_row + (uint32_t)(32 * h + 16), v1);
      ptx::tmem_ld_wait();
  #pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = 32 * h + 2 * i;
        oh2[16 * h + i] = __floats2half2_rn(fmaxf(fmaf(__uint_as_float(v0[2 * i]), s_scale[c], s_bias[c]), 0.f),
                                           fmaxf(fmaf(__uint_as_float(v0[2 * i + 1]), s_scale[c + 1], s_bias[c + 1]), 0.f));
        oh2[16 * h + 8 + i] = __floats2half2_rn(fmaxf(fmaf(__uint_as_float(v1[2 * i]), s_scale[c + 16], s_bias[c + 16]), 0.f),
                                               fmaxf(fmaf(__uint_as_float(v1[2 * i + 1]), s_scale[c + 17], s_bias[c + 17]), 0.f));
      }
    }
    if (valid) {
      uint4* op = reinterpret_cast<uint4*>(out + (size_t)pix * 64);
  #pragma unroll
      for (int i = 0; i < 8; ++i) op[i] = o[i];
    }
    ptx::tc_fence_before_sync();   // this tile's TMEM reads are ordered before the next tile's MMAs by the __syncthreads above
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after_sync(); ptx::tmem_dealloc(tmem, 64); }
}

constexpr int kStemSmem = 1024 + 49152 + 512 + 64;
cudaError_t stem_tc_set_attributes() {
  cudaError_t e = cudaFuncSetAttribute(stem_conv3x3s2_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStemSmem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(stem_conv3x3s2_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStemSmem);
  return e;
}

static cudaError_t launch_stem_tc_any(const void* in, bool u8, const float* w, const float* scale, const float* bias,
                                      __half* out, int N, int H, int W, int num_sms, cudaStream_t st) {
  const long total = (long)N * (H / 2) * (W / 2);
  if (total == 0) return cudaSuccess;
  const unsigned grid = (unsigned)std::min<long>((total + 127) / 128, (long)num_sms * 4);
  if (u8) stem_conv3x3s2_tc_kernel<true><<<grid, 128, kStemSmem, st>>>(in, w, scale, bias, out, N, H, W);
  else stem_conv3x3s2_tc_kernel<false><<<grid, 128, kStemSmem, st>>>(in, w, scale, bias, out, N, H, W);
  return cudaGetLastError();
}

cudaError_t launch_stem_tc(const float* in_nchw, const float* w, const float* scale, const float* bias, __half* out,
                           int N, int H, int W, int num_sms, cudaStream_t st) {
  return launch_stem_tc_any(in_nchw, false, w, scale, bias, out, N, H, W, num_sms, st);
}

cudaError_t launch_stem_tc_u8(const uint8_t* in_nhwc_bgr, const float* w, const float* scale, const float* bias,
                              __half* out, int N, int H, int W, int num_sms, cudaStream_t st) {
  return launch_stem_tc_any(in_nhwc_bgr, true, w, scale, bias, out, N, H, W, num_sms, st);
}

}  // namespace hrnet
