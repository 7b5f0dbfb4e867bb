// HBM-bound / odd-shaped kernels of the hot path that do not belong on the tensor cores:
//   stem_conv3x3s2   : models_/hrnet.py:158-160  (3->64, reads the NCHW fp32 input, writes NHWC fp16)
//   stem_conv7x7s2   : models_/poseresnet.py:109-111 (3->64, 7x7)
//   maxpool3x3s2     : models_/poseresnet.py:112
//   fuse_sum         : models_/hrnet.py:60-69    (sum of identity / nearest-upsampled / down-sampled terms + ReLU)
//   head_conv1x1     : models_/hrnet.py:187      (c->J 1x1 conv with bias, fp32 NCHW heatmaps)
//   argmax_decode    : SimpleHRNet.py:296-308    (np.argmax first-occurrence + (y, x, conf) scaling)
//   final_preds      : misc/utils.py:125-182     (get_max_preds + quarter-pixel refine + inverse affine, evaluation side)
//   flip_average     : training/COCO.py:206-212  (flip-test average of the heat-maps)
//   conv_simt        : generic direct conv used as the debug cross-check of the tcgen05 path and for shapes
//                      the implicit GEMM does not cover (Cin or Cout not a multiple of 16)
#include "hrnet_internal.h"

namespace hrnet {

// ------------------------------------------------------------------------------------------------ stem 3x3 s2
// block = 256 threads = 64 output pixels x 4 channel groups of 16.  Weights sit in shared memory transposed to
// [tap][cout] so a thread fetches its 16 output channels of one tap with four 128-bit loads (the first version
// read them one float at a time and was shared-memory-issue bound: 842 us at N=64, profiles/r01_launches_v1.csv).
__global__ void __launch_bounds__(256)
stem_conv3x3s2_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ scale,
                      const float* __restrict__ bias, __half* __restrict__ out, int N, int H, int W) {
  __shared__ __align__(16) float sw[27 * 64];  // [t = (r*3+s)*3+ci][co]
  __shared__ __align__(16) float ss[64], sb[64];
  for (int i = threadIdx.x; i < 64 * 27; i += 256) {
    const int co = i / 27, t = i - co * 27;   // global layout [co][r][s][ci]
    sw[t * 64 + co] = w[i];
  }
  if (threadIdx.x < 64) { ss[threadIdx.x] = scale[threadIdx.x]; sb[threadIdx.x] = bias[threadIdx.x]; }
  __syncthreads();
  const int OH = H / 2, OW = W / 2;
  const long total = (long)N * OH * OW;
  const long pix = (long)blockIdx.x * 64 + (threadIdx.x >> 2);
  const int cg = threadIdx.x & 3;
  if (pix >= total) return;
  const int n = (int)(pix / (OH * OW));
  const int rem = (int)(pix - (long)n * OH * OW);
  const int oh = rem / OW, ow = rem - oh * OW;
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int ih = oh * 2 - 1 + r;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const int iw = ow * 2 - 1 + s;
      const bool ok = ih >= 0 && ih < H && iw >= 0 && iw < W;
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) {
        const float x = ok ? __ldg(in + (((size_t)n * 3 + ci) * H + ih) * W + iw) : 0.f;
        const float4* wp = reinterpret_cast<const float4*>(sw + ((r * 3 + s) * 3 + ci) * 64 + cg * 16);
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const float4 wv = wp[v];
          acc[4 * v + 0] = fmaf(x, wv.x, acc[4 * v + 0]);
          acc[4 * v + 1] = fmaf(x, wv.y, acc[4 * v + 1]);
          acc[4 * v + 2] = fmaf(x, wv.z, acc[4 * v + 2]);
          acc[4 * v + 3] = fmaf(x, wv.w, acc[4 * v + 3]);
        }
      }
    }
  }
  uint4 o[2];
  __half2* oh2 = reinterpret_cast<__half2*>(o);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int co = cg * 16 + 2 * i;
    oh2[i] = __floats2half2_rn(fmaxf(acc[2 * i] * ss[co] + sb[co], 0.f),
                               fmaxf(acc[2 * i + 1] * ss[co + 1] + sb[co + 1], 0.f));
  }
  uint4* op = reinterpret_cast<uint4*>(out + (size_t)pix * 64 + cg * 16);
  op[0] = o[0];
  op[1] = o[1];
}

cudaError_t launch_stem(const float* in_nchw, const float* w, const float* scale, const float* bias, __half* out,
                        int N, int H, int W, cudaStream_t st) {
  const long total = (long)N * (H / 2) * (W / 2);
  if (total == 0) return cudaSuccess;
  stem_conv3x3s2_kernel<<<(unsigned)((total + 63) / 64), 256, 0, st>>>(in_nchw, w, scale, bias, out, N, H, W);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ stem 7x7 s2 p3
// one thread = one output pixel x 16 channels; weights [co][r][s][ci] fp32 in shared memory (64*147 floats)
__global__ void __launch_bounds__(256)
stem_conv7x7s2_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ scale,
                      const float* __restrict__ bias, __half* __restrict__ out, int N, int H, int W) {
  extern __shared__ float sw7[];  // 64*147
  __shared__ float ss[64], sb[64];
  for (int i = threadIdx.x; i < 64 * 147; i += 256) sw7[i] = w[i];
  if (threadIdx.x < 64) { ss[threadIdx.x] = scale[threadIdx.x]; sb[threadIdx.x] = bias[threadIdx.x]; }
  __syncthreads();
  const int OH = H / 2, OW = W / 2;
  const long total = (long)N * OH * OW;
  const long pix = (long)blockIdx.x * 64 + (threadIdx.x >> 2);
  const int cg = threadIdx.x & 3;
  if (pix >= total) return;
  const int n = (int)(pix / (OH * OW));
  const int rem = (int)(pix - (long)n * OH * OW);
  const int oh = rem / OW, ow = rem - oh * OW;
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  for (int r = 0; r < 7; ++r) {
    const int ih = oh * 2 - 3 + r;
    if (ih < 0 || ih >= H) continue;
    for (int s = 0; s < 7; ++s) {
      const int iw = ow * 2 - 3 + s;
      if (iw < 0 || iw >= W) continue;
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) {
        const float x = __ldg(in + (((size_t)n * 3 + ci) * H + ih) * W + iw);
        const int t = (r * 7 + s) * 3 + ci;
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fmaf(x, sw7[(cg * 16 + i) * 147 + t], acc[i]);
      }
    }
  }
  uint4 o[2];
  __half2* oh2 = reinterpret_cast<__half2*>(o);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int co = cg * 16 + 2 * i;
    oh2[i] = __floats2half2_rn(fmaxf(acc[2 * i] * ss[co] + sb[co], 0.f),
                               fmaxf(acc[2 * i + 1] * ss[co + 1] + sb[co + 1], 0.f));
  }
  uint4* op = reinterpret_cast<uint4*>(out + (size_t)pix * 64 + cg * 16);
  op[0] = o[0];
  op[1] = o[1];
}

cudaError_t launch_stem7(const float* in_nchw, const float* w, const float* scale, const float* bias, __half* out,
                         int N, int H, int W, cudaStream_t st) {
  const long total = (long)N * (H / 2) * (W / 2);
  if (total == 0) return cudaSuccess;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(stem_conv7x7s2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 147 * 4);
    attr = true;
  }
  stem_conv7x7s2_kernel<<<(unsigned)((total + 63) / 64), 256, 64 * 147 * 4, st>>>(in_nchw, w, scale, bias, out, N,
                                                                                  H, W);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ maxpool 3x3 s2 p1
__global__ void __launch_bounds__(256)
maxpool3x3s2_kernel(const __half* __restrict__ in, __half* __restrict__ out, int N, int IH, int IW, int C) {
  const int OH = IH / 2, OW = IW / 2, CV = C / 8;
  const long total = (long)N * OH * OW * CV;
  const long i = (long)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  const int cv = (int)(i % CV);
  const long pix = i / CV;
  const int n = (int)(pix / (OH * OW));
  const int rem = (int)(pix - (long)n * OH * OW);
  const int oh = rem / OW, ow = rem - oh * OW;
  __half2 m[4];
  const __half2 ninf = __float2half2_rn(-65504.f);
#pragma unroll
  for (int k = 0; k < 4; ++k) m[k] = ninf;
  for (int r = 0; r < 3; ++r) {
    const int ih = oh * 2 - 1 + r;
    if (ih < 0 || ih >= IH) continue;
    for (int s = 0; s < 3; ++s) {
      const int iw = ow * 2 - 1 + s;
      if (iw < 0 || iw >= IW) continue;
      uint4 v = __ldg(reinterpret_cast<const uint4*>(in + (((size_t)n * IH + ih) * IW + iw) * C + cv * 8));
      const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
      for (int k = 0; k < 4; ++k) m[k] = __hmax2(m[k], h[k]);
    }
  }
  uint4 o;
  __half2* oh2 = reinterpret_cast<__half2*>(&o);
#pragma unroll
  for (int k = 0; k < 4; ++k) oh2[k] = m[k];
  *reinterpret_cast<uint4*>(out + (size_t)pix * C + cv * 8) = o;
}

cudaError_t launch_maxpool(const __half* in, __half* out, int N, int IH, int IW, int C, cudaStream_t st) {
  const long total = (long)N * (IH / 2) * (IW / 2) * (C / 8);
  if (total == 0) return cudaSuccess;
  maxpool3x3s2_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, out, N, IH, IW, C);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ fuse sum
// out[n,h,w,c] = act( sum_j src_j[n, h >> shift_j, w >> shift_j, c] ), fp32 sum in ascending j like hrnet.py:61-66.
// one thread = 8 channels (16 B fp16 / 32 B fp32 per source)
// (32-bit index math only: image n comes from blockIdx.y; the 64-bit divisions of a flat index cost more than the loads)
__device__ __forceinline__ void fuse_sum_one(const FuseParams& p, int n, int t, int CV) {
  const int rem = t / CV;                 // pixel within the image
  const int cv = t - rem * CV;
  const long pix = (long)n * (p.H * p.W) + rem;
  const int h = rem / p.W, w = rem - h * p.W;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (j >= p.nsrc) break;
    const int sh = p.shift[j];
    const int sH = p.H >> sh, sW = p.W >> sh;
    const size_t off = ((((size_t)n * sH + (h >> sh)) * sW + (w >> sh)) * p.C) + (size_t)cv * 8;
    float v[8];
    if (p.f32[j]) {
      const float4* sp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.src[j]) + off);
      float4 a = __ldg(sp), b = __ldg(sp + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
      uint4 u = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(p.src[j]) + off));
      const __half2* hh = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int k = 0; k < 4; ++k) { float2 f = __half22float2(hh[k]); v[2 * k] = f.x; v[2 * k + 1] = f.y; }
    }
    if (j == 0) {
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = v[k];
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += v[k];
    }
  }
  uint4 o;
  __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float a = acc[2 * k], b = acc[2 * k + 1];
    if (p.relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
    oh[k] = __floats2half2_rn(a, b);
  }
  *reinterpret_cast<uint4*>(p.out + (size_t)pix * p.C + (size_t)cv * 8) = o;
}

__global__ void __launch_bounds__(256) fuse_sum_kernel(const __grid_constant__ FuseParams p) {
  const int CV = p.C / 8;
  const int per_image = p.H * p.W * CV;
  const int t = (int)blockIdx.x * 256 + (int)threadIdx.x;
  if (t >= per_image) return;
  fuse_sum_one(p, (int)blockIdx.y, t, CV);   // (two elements per thread measured no faster: 34.9 vs 32.5 us for the 96x72x48 sum)
}

cudaError_t launch_fuse(const FuseParams& p, cudaStream_t st) {
  const int per_image = p.H * p.W * (p.C / 8);
  if (per_image == 0 || p.N == 0) return cudaSuccess;
  fuse_sum_kernel<<<dim3((unsigned)((per_image + 255) / 256), (unsigned)p.N), 256, 0, st>>>(p);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ head 1x1 + bias
// in NHWC fp16 [N*HW, Cin], w fp32 [J][Cin], out NCHW fp32 [N, J, HW]; one thread = one pixel, all joints
constexpr int kHeadMaxJ = 32;
__global__ void __launch_bounds__(128)
head_conv1x1_kernel(const __half* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                    float* __restrict__ out, int N, int HW, int Cin, int J) {
  extern __shared__ float shw[];  // [J][Cin] + [J]
  float* sbias = shw + J * Cin;
  for (int i = threadIdx.x; i < J * Cin; i += 128) shw[i] = w[i];
  for (int i = threadIdx.x; i < J; i += 128) sbias[i] = bias[i];
  __syncthreads();
  const long pix = (long)blockIdx.x * 128 + threadIdx.x;
  if (pix >= (long)N * HW) return;
  const int n = (int)(pix / HW);
  const int hw = (int)(pix - (long)n * HW);
  float acc[kHeadMaxJ];
#pragma unroll
  for (int j = 0; j < kHeadMaxJ; ++j) acc[j] = 0.f;
  const __half* ip = in + (size_t)pix * Cin;
  for (int c = 0; c < Cin; c += 8) {
    uint4 u = __ldg(reinterpret_cast<const uint4*>(ip + c));
    const __half2* hh = reinterpret_cast<const __half2*>(&u);
    float x[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) { float2 f = __half22float2(hh[k]); x[2 * k] = f.x; x[2 * k + 1] = f.y; }
#pragma unroll
    for (int j = 0; j < kHeadMaxJ; ++j) {
      if (j < J) {   // two 16-byte broadcast reads per joint instead of eight scalar ones (the loop was LDS-bound)
        const float4 w0 = *reinterpret_cast<const float4*>(shw + j * Cin + c);
        const float4 w1 = *reinterpret_cast<const float4*>(shw + j * Cin + c + 4);
        float a = acc[j];
        a = fmaf(x[0], w0.x, a); a = fmaf(x[1], w0.y, a); a = fmaf(x[2], w0.z, a); a = fmaf(x[3], w0.w, a);
        a = fmaf(x[4], w1.x, a); a = fmaf(x[5], w1.y, a); a = fmaf(x[6], w1.z, a); a = fmaf(x[7], w1.w, a);
        acc[j] = a;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < kHeadMaxJ; ++j)
    if (j < J) out[((size_t)n * J + j) * HW + hw] = acc[j] + sbias[j];
}

cudaError_t launch_head(const __half* in, const float* w, const float* bias, float* out_nchw, int N, int HW, int Cin,
                        int J, cudaStream_t st) {
  const long total = (long)N * HW;
  if (total == 0) return cudaSuccess;
  if (J > kHeadMaxJ) return cudaErrorInvalidValue;
  const int smem = (J * Cin + J) * 4;
  if (smem > 48 * 1024) {   // (17 joints x 2048 channels is 139 KB; per device, so no cached state)
    cudaError_t e = cudaFuncSetAttribute(head_conv1x1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
  }
  head_conv1x1_kernel<<<(unsigned)((total + 127) / 128), 128, smem, st>>>(in, w, bias, out_nchw, N, HW, Cin, J);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ argmax decode
// One 256-thread block per (person, joint) heatmap.  np.argmax semantics: flat row-major index of the first
// maximum; a NaN counts as the maximum (first NaN wins).  Warp-shuffle reduction, then one warp reduces the
// 8 per-warp candidates.  (y, x) follow SimpleHRNet.py:306-307 evaluated in float64 and rounded to float32.
__device__ __forceinline__ bool better(float a, int ia, float b, int ib) {
  const bool an = a != a, bn = b != b;
  if (an != bn) return an;        // NaN beats non-NaN
  if (an) return ia < ib;         // both NaN: first
  if (a != b) return a > b;
  return ia < ib;                 // tie: first occurrence
}

__global__ void __launch_bounds__(256)
argmax_decode_kernel(const float* __restrict__ hm, int J, int Hh, int Wh, const float* __restrict__ boxes,
                     float* __restrict__ joints, int32_t* __restrict__ idx_out) {
  const int pj = blockIdx.x;  // person * J + joint
  const int HW = Hh * Wh;
  const float* p = hm + (size_t)pj * HW;
  float best = 0.f;
  int bi = 0x7fffffff;
  // vectorised scan when every map is 16-byte aligned (HW % 4 == 0 holds for every supported resolution; the base
  // pointer of hrnet_argmax / hrnet_forward(heatmaps=...) is the caller's and may be a view at any 4-byte offset)
  if ((HW & 3) == 0 && (reinterpret_cast<uintptr_t>(hm) & 15u) == 0) {
    const float4* p4 = reinterpret_cast<const float4*>(p);
    for (int i = threadIdx.x; i < HW / 4; i += 256) {
      float4 v = __ldg(p4 + i);
      const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (bi == 0x7fffffff || better(e[k], 4 * i + k, best, bi)) { best = e[k]; bi = 4 * i + k; }
    }
  } else {
    for (int i = threadIdx.x; i < HW; i += 256) {
      const float v = __ldg(p + i);
      if (bi == 0x7fffffff || better(v, i, best, bi)) { best = v; bi = i; }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (oi != 0x7fffffff && (bi == 0x7fffffff || better(ob, oi, best, bi))) { best = ob; bi = oi; }
  }
  __shared__ float sv[8];
  __shared__ int si[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sv[warp] = best; si[warp] = bi; }
  __syncthreads();
  if (warp == 0) {
    best = lane < 8 ? sv[lane] : 0.f;
    bi = lane < 8 ? si[lane] : 0x7fffffff;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (oi != 0x7fffffff && (bi == 0x7fffffff || better(ob, oi, best, bi))) { best = ob; bi = oi; }
    }
    if (lane == 0) {
      const int person = pj / J;
      const int r = bi / Wh, c = bi - r * Wh;
      float x1 = 0.f, y1 = 0.f, x2 = (float)(Wh * 4), y2 = (float)(Hh * 4);
      if (boxes != nullptr) {
        x1 = boxes[person * 4 + 0]; y1 = boxes[person * 4 + 1];
        x2 = boxes[person * 4 + 2]; y2 = boxes[person * 4 + 3];
      }
      const float dy = y2 - y1, dx = x2 - x1;  // float32 subtraction like the numpy f32 box array
      const double y = (double)r * 1.0 / (double)Hh * (double)dy + (double)y1;
      const double x = (double)c * 1.0 / (double)Wh * (double)dx + (double)x1;
      joints[(size_t)pj * 3 + 0] = (float)y;
      joints[(size_t)pj * 3 + 1] = (float)x;
      joints[(size_t)pj * 3 + 2] = best;
      if (idx_out != nullptr) idx_out[pj] = bi;
    }
  }
}

// ------------------------------------------------------------------------------------------------ head (+ argmax)
// The head above is bound by the shared-memory broadcast reads of its weights (17 x 48 x 4 B per pixel-thread: an LDS.128
// costs four L1 wavefronts even when every lane reads the same address -> 0.23 of the HBM roofline); here the weights
// are kernel parameters: every FMA's second operand comes from the constant bank.  Same accumulation order as
// head_conv1x1_kernel -> bit-identical heat-map values.  grid = (blocks of 128 pixels, person).
// head_conv1x1_kernel is bound by the shared-memory broadcast reads of its weights (816 L1 wavefronts per 32 pixels = 42 us
// at W48 / 64 crops), not by HBM.  What was tried against that, all slower on B200 (profiles/r02_s28_head_stem.log,
// r02_s29_head_four_pixels.log): weights from the constant bank (one uniform load per FMA: 86 us; four pixels per thread:
// 64 us -- the uniform loads form a latency chain) and four pixels per thread with shared-memory weights (76 us: 135
// registers, 896 blocks = 2.02 waves).  Kept: the original mapping (kHeadPix = 1) with the argmax candidates fused in.
constexpr int kHeadPix = 1;
template <int CIN, int NJ>
__global__ void __launch_bounds__(128) head_c_kernel(const __grid_constant__ HeadParams p) {
  const int n = blockIdx.y;
  const int hw0 = (int)blockIdx.x * (128 * kHeadPix) + (int)threadIdx.x;      // pixels hw0 + 128 * i
  __shared__ __align__(16) float shw[NJ * CIN];
  for (int i = threadIdx.x; i < NJ * CIN; i += 128) shw[i] = __ldg(p.w_dev + i);      // (per-lane indexed reads of the constant bank serialise)
  __syncthreads();
  float acc[kHeadPix][NJ];
#pragma unroll
  for (int i = 0; i < kHeadPix; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j) acc[i][j] = 0.f;
#pragma unroll
  for (int c = 0; c < CIN; c += 8) {
    float x[kHeadPix][8];
#pragma unroll
    for (int i = 0; i < kHeadPix; ++i) {
      const int hw = hw0 + 128 * i;
      uint4 u = make_uint4(0u, 0u, 0u, 0u);
      if (hw < p.HW) u = __ldg(reinterpret_cast<const uint4*>(p.in + ((size_t)n * p.HW + hw) * CIN + c));
      const __half2* hh = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int k = 0; k < 4; ++k) { const float2 f = __half22float2(hh[k]); x[i][2 * k] = f.x; x[i][2 * k + 1] = f.y; }
    }
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      // two 16-byte broadcast reads feed 8 x kHeadPix FMAs per lane (a broadcast LDS.128 costs four L1 wavefronts)
      const float4 w0 = *reinterpret_cast<const float4*>(shw + j * CIN + c);
      const float4 w1 = *reinterpret_cast<const float4*>(shw + j * CIN + c + 4);
      const float wk[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int k = 0; k < 8; ++k) {
#pragma unroll
        for (int i = 0; i < kHeadPix; ++i) acc[i][j] = fmaf(x[i][k], wk[k], acc[i][j]);     // same order per pixel as head_conv1x1_kernel
      }
    }
  }
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const float bj = p.bias[j];      // (compile-time index: a constant-bank operand)
#pragma unroll
    for (int i = 0; i < kHeadPix; ++i) acc[i][j] += bj;
  }
  if (p.out != nullptr) {
#pragma unroll
    for (int i = 0; i < kHeadPix; ++i) {
      const int hw = hw0 + 128 * i;
      if (hw < p.HW) {
#pragma unroll
        for (int j = 0; j < NJ; ++j) p.out[((size_t)n * NJ + j) * p.HW + hw] = acc[i][j];
      }
    }
  }
  if (p.pval == nullptr) return;
  // per-joint candidate of this block: np.argmax order (first maximum, a NaN is the maximum) is a total order on
  // (value, index), so partial results combine in any grouping
  __shared__ float sv[4][NJ];
  __shared__ int si[4][NJ];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // warp stage with the hardware reduction (redux.sync) on an order-preserving 32-bit key: NaN -> the largest key (a NaN
  // is np.argmax's maximum), -0.0 == +0.0 (same key), ties -> the lowest lane = the lowest pixel index of the warp.
  // (Five shuffle rounds of (value, index) pairs per joint cost more than the 816 FMAs of the head itself.)
  static_assert(kHeadPix == 1, "the warp stage assumes one pixel per lane, consecutive lanes = consecutive pixels");
  const bool valid0 = hw0 < p.HW;
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const float v = acc[0][j];
    const uint32_t u = __float_as_uint(v + 0.f);                      // -0.0 + 0.0 = +0.0
    uint32_t key = (u & 0x80000000u) ? ~u : (u | 0x80000000u);        // monotone in v for non-NaN
    if (v != v) key = 0xffffffffu;
    if (!valid0) key = 0u;                                            // (below every real key: ~(-NaN bits) is >= 0x00400000)
    const uint32_t kmax = __reduce_max_sync(0xffffffffu, key);
    const uint32_t who = __ballot_sync(0xffffffffu, key == kmax && valid0);
    if (who != 0u && lane == __ffs((int)who) - 1) { sv[warp][j] = v; si[warp][j] = hw0; }
    else if (who == 0u && lane == 0) { sv[warp][j] = 0.f; si[warp][j] = 0x7fffffff; }
  }
  __syncthreads();
  if ((int)threadIdx.x < NJ) {
    const int j = threadIdx.x;
    float best = sv[0][j];
    int bi = si[0][j];
#pragma unroll
    for (int w = 1; w < 4; ++w) {
      const float ob = sv[w][j];
      const int oi = si[w][j];
      if (oi != 0x7fffffff && (bi == 0x7fffffff || better(ob, oi, best, bi))) { best = ob; bi = oi; }
    }
    const size_t o = ((size_t)n * NJ + j) * p.nblk + blockIdx.x;
    p.pval[o] = best;
    p.pidx[o] = bi;
  }
}

int head_c_blocks(int hw) { return (hw + 128 * kHeadPix - 1) / (128 * kHeadPix); }
bool head_c_supported(int cin, int nj) { return nj == 17 && (cin == 48 || cin == 32); }
cudaError_t launch_head_c(const HeadParams& p, cudaStream_t st) {
  if (p.N == 0 || p.HW == 0) return cudaSuccess;
  const dim3 grid((unsigned)p.nblk, (unsigned)p.N);
  if (p.J == 17 && p.Cin == 48) head_c_kernel<48, 17><<<grid, 128, 0, st>>>(p);
  else if (p.J == 17 && p.Cin == 32) head_c_kernel<32, 17><<<grid, 128, 0, st>>>(p);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

// one warp per (person, joint): reduces the block candidates and decodes like argmax_decode_kernel
__global__ void __launch_bounds__(128)
head_argmax_finish_kernel(const float* __restrict__ pval, const int* __restrict__ pidx, int NJ, int J, int nblk, int Hh, int Wh,
                          const float* __restrict__ boxes, float* __restrict__ joints, int32_t* __restrict__ idx_out) {
  const int pj = (int)blockIdx.x * 4 + (int)(threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (pj >= NJ) return;
  float best = 0.f;
  int bi = 0x7fffffff;
  for (int i = lane; i < nblk; i += 32) {
    const float v = pval[(size_t)pj * nblk + i];
    const int vi = pidx[(size_t)pj * nblk + i];
    if (vi != 0x7fffffff && (bi == 0x7fffffff || better(v, vi, best, bi))) { best = v; bi = vi; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (oi != 0x7fffffff && (bi == 0x7fffffff || better(ob, oi, best, bi))) { best = ob; bi = oi; }
  }
  if (lane == 0) {
    const int person = pj / J;
    const int r = bi / Wh, c = bi - r * Wh;
    float x1 = 0.f, y1 = 0.f, x2 = (float)(Wh * 4), y2 = (float)(Hh * 4);
    if (boxes != nullptr) {
      x1 = boxes[person * 4 + 0]; y1 = boxes[person * 4 + 1];
      x2 = boxes[person * 4 + 2]; y2 = boxes[person * 4 + 3];
    }
    const float dy = y2 - y1, dx = x2 - x1;  // float32 subtraction like the numpy f32 box array
    const double y = (double)r * 1.0 / (double)Hh * (double)dy + (double)y1;
    const double x = (double)c * 1.0 / (double)Wh * (double)dx + (double)x1;
    joints[(size_t)pj * 3 + 0] = (float)y;
    joints[(size_t)pj * 3 + 1] = (float)x;
    joints[(size_t)pj * 3 + 2] = best;
    if (idx_out != nullptr) idx_out[pj] = bi;
  }
}

cudaError_t launch_head_argmax_finish(const float* pval, const int* pidx, int N, int J, int nblk, int Hh, int Wh,
                                      const float* boxes, float* joints, int32_t* idx, cudaStream_t st) {
  if (N * J == 0) return cudaSuccess;
  head_argmax_finish_kernel<<<(unsigned)((N * J + 3) / 4), 128, 0, st>>>(pval, pidx, N * J, J, nblk, Hh, Wh, boxes, joints, idx);
  return cudaGetLastError();
}

cudaError_t launch_argmax(const float* hm, int N, int J, int Hh, int Wh, const float* boxes, float* joints,
                          int32_t* idx, cudaStream_t st) {
  if (N * J == 0) return cudaSuccess;
  argmax_decode_kernel<<<N * J, 256, 0, st>>>(hm, J, Hh, Wh, boxes, joints, idx);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ evaluation decode
// get_max_preds + the quarter-pixel refinement + the inverse affine of get_final_preds (misc/utils.py:125-182), one
// 256-thread block per (person, joint) heat-map.  Same first-maximum / NaN rule as the argmax above (torch.max == np.argmax).
//   preds (x, y) = (idx % W, floor(idx / W)) as float32, zeroed when the maximum is not > 0 (`preds *= pred_mask`);
//   post_processing: px = floor(x + .5), py likewise; if 1 < px < W-1 and 1 < py < H-1 the joint moves a quarter pixel
//     towards the higher neighbour: += sign(hm[py][px+1] - hm[py][px-1]) * .25 (torch.sign: 0 for 0 and for NaN);
//   trans (optional, [n][2][3] float64 = cv2.getAffineTransform of misc/utils.py:46-79 with inv = 1, built on the host):
//     (x, y) <- t . (x, y, 1) evaluated in float64 without contraction, stored as float32 (np.dot of transform_preds).
__global__ void __launch_bounds__(256)
final_preds_kernel(const float* __restrict__ hm, int J, int Hh, int Wh, int post, const double* __restrict__ trans,
                   float* __restrict__ preds, float* __restrict__ maxvals) {
  const int pj = blockIdx.x;  // person * J + joint
  const int HW = Hh * Wh;
  const float* p = hm + (size_t)pj * HW;
  float best = 0.f;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < HW; i += 256) {
    const float v = __ldg(p + i);
    if (bi == 0x7fffffff || better(v, i, best, bi)) { best = v; bi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (oi != 0x7fffffff && (bi == 0x7fffffff || better(ob, oi, best, bi))) { best = ob; bi = oi; }
  }
  __shared__ float sv[8];
  __shared__ int si[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sv[warp] = best; si[warp] = bi; }
  __syncthreads();
  if (warp != 0) return;
  best = lane < 8 ? sv[lane] : 0.f;
  bi = lane < 8 ? si[lane] : 0x7fffffff;
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (oi != 0x7fffffff && (bi == 0x7fffffff || better(ob, oi, best, bi))) { best = ob; bi = oi; }
  }
  if (lane != 0) return;
  const int r = bi / Wh, c = bi - r * Wh;
  float x = (float)c, y = (float)r;
  if (!(best > 0.f)) { x = 0.f; y = 0.f; }
  if (post) {
    const int px = (int)floorf(x + 0.5f), py = (int)floorf(y + 0.5f);
    if (1 < px && px < Wh - 1 && 1 < py && py < Hh - 1) {
      const float dx = __fsub_rn(__ldg(p + py * Wh + px + 1), __ldg(p + py * Wh + px - 1));
      const float dy = __fsub_rn(__ldg(p + (py + 1) * Wh + px), __ldg(p + (py - 1) * Wh + px));
      x += dx > 0.f ? 0.25f : (dx < 0.f ? -0.25f : 0.f);
      y += dy > 0.f ? 0.25f : (dy < 0.f ? -0.25f : 0.f);
    }
  }
  if (trans != nullptr) {
    const double* t = trans + (size_t)(pj / J) * 6;
    const double xd = (double)x, yd = (double)y;
    const double nx = __dadd_rn(__dadd_rn(__dmul_rn(t[0], xd), __dmul_rn(t[1], yd)), t[2]);
    const double ny = __dadd_rn(__dadd_rn(__dmul_rn(t[3], xd), __dmul_rn(t[4], yd)), t[5]);
    x = (float)nx; y = (float)ny;
  }
  preds[(size_t)pj * 2 + 0] = x;
  preds[(size_t)pj * 2 + 1] = y;
  maxvals[pj] = best;
}

cudaError_t launch_final_preds(const float* hm, int N, int J, int Hh, int Wh, int post, const double* trans, float* preds,
                               float* maxvals, cudaStream_t st) {
  if (N * J == 0) return cudaSuccess;
  final_preds_kernel<<<N * J, 256, 0, st>>>(hm, J, Hh, Wh, post, trans, preds, maxvals);
  return cudaGetLastError();
}

// Flip test (training/COCO.py:206-212, misc/utils.py:9-29): out = (a + flip_back(b)) * 0.5 where flip_back mirrors every
// map along x and swaps the left / right joints back: out[n, j, h, w] = (a[n, j, h, w] + b[n, perm[j], h, W-1-w]) * 0.5.
// HBM-bound (three maps of 4 * J * Hh * Wh bytes per person); one thread per 4 consecutive output pixels when W % 4 == 0.
struct FlipPerm { int v[32]; };
__global__ void __launch_bounds__(256)
flip_average_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, const FlipPerm perm,
                    int J, int Hh, int Wh, long total_rows, int vec) {
  const int wq = vec ? Wh / 4 : Wh;
  const long i = (long)blockIdx.x * 256 + threadIdx.x;
  if (i >= total_rows * wq) return;
  const long row = i / wq;                 // (n * J + j) * Hh + h
  const int wi = (int)(i - row * wq);
  const long nj = row / Hh;
  const int h = (int)(row - nj * Hh);
  const long n = nj / J;
  const int j = (int)(nj - n * J);
  const long brow = ((n * J + perm.v[j]) * Hh + h) * (long)Wh;
  if (vec) {
    const float4 av = __ldg(reinterpret_cast<const float4*>(a + row * Wh) + wi);
    const float4 bv = __ldg(reinterpret_cast<const float4*>(b + brow) + (wq - 1 - wi));
    float4 o;
    o.x = __fmul_rn(__fadd_rn(av.x, bv.w), 0.5f); o.y = __fmul_rn(__fadd_rn(av.y, bv.z), 0.5f);
    o.z = __fmul_rn(__fadd_rn(av.z, bv.y), 0.5f); o.w = __fmul_rn(__fadd_rn(av.w, bv.x), 0.5f);
    reinterpret_cast<float4*>(out + row * Wh)[wi] = o;
  } else {
    out[row * Wh + wi] = __fmul_rn(__fadd_rn(__ldg(a + row * Wh + wi), __ldg(b + brow + (Wh - 1 - wi))), 0.5f);
  }
}

cudaError_t launch_flip_average(const float* a, const float* b, float* out, const int* perm, int N, int J, int Hh, int Wh,
                                cudaStream_t st) {
  if ((long)N * J * Hh * Wh == 0) return cudaSuccess;
  FlipPerm fp;
  for (int j = 0; j < 32; ++j) fp.v[j] = j < J ? perm[j] : j;
  const int vec = (Wh % 4 == 0) && ((((uintptr_t)a | (uintptr_t)b | (uintptr_t)out) & 15u) == 0);
  const long rows = (long)N * J * Hh;
  const long work = rows * (vec ? Wh / 4 : Wh);
  flip_average_kernel<<<(unsigned)((work + 255) / 256), 256, 0, st>>>(a, b, out, fp, J, Hh, Wh, rows, vec);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ generic SIMT conv
// one thread = one output pixel x 8 output channels; fp32 accumulate.  Debug cross-check, not the hot path.
__global__ void __launch_bounds__(128) conv_simt_kernel(const ConvSimtParams p) {
  const int CO8 = p.Cout / 8;
  const long total = (long)p.N * p.OH * p.OW * CO8;
  const long i = (long)blockIdx.x * 128 + threadIdx.x;
  if (i >= total) return;
  const int cg = (int)(i % CO8);
  const long pix = i / CO8;
  const int n = (int)(pix / (p.OH * p.OW));
  const int rem = (int)(pix - (long)n * p.OH * p.OW);
  const int oh = rem / p.OW, ow = rem - oh * p.OW;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  const int K = p.ksize * p.ksize * p.Cin;
  for (int r = 0; r < p.ksize; ++r) {
    const int ih = oh * p.stride - p.pad + r;
    if (ih < 0 || ih >= p.IH) continue;
    for (int s = 0; s < p.ksize; ++s) {
      const int iw = ow * p.stride - p.pad + s;
      if (iw < 0 || iw >= p.IW) continue;
      const __half* ip = p.in + (((size_t)n * p.IH + ih) * p.IW + iw) * p.Cin;
      const __half* wp = p.w + (size_t)(cg * 8) * K + (size_t)(r * p.ksize + s) * p.Cin;
      for (int c = 0; c < p.Cin; c += 8) {
        uint4 u = __ldg(reinterpret_cast<const uint4*>(ip + c));
        const __half2* xh = reinterpret_cast<const __half2*>(&u);
        float x[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) { float2 f = __half22float2(xh[k]); x[2 * k] = f.x; x[2 * k + 1] = f.y; }
#pragma unroll
        for (int o = 0; o < 8; ++o) {
          uint4 wu = __ldg(reinterpret_cast<const uint4*>(wp + (size_t)o * K + c));
          const __half2* wh = reinterpret_cast<const __half2*>(&wu);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float2 f = __half22float2(wh[k]);
            acc[o] = fmaf(x[2 * k], f.x, acc[o]);
            acc[o] = fmaf(x[2 * k + 1], f.y, acc[o]);
          }
        }
      }
    }
  }
  const size_t off = (size_t)pix * p.Cout + (size_t)cg * 8;
  float y[8];
#pragma unroll
  for (int o = 0; o < 8; ++o) y[o] = acc[o] * p.scale[cg * 8 + o] + p.bias[cg * 8 + o];
  if (p.residual != nullptr) {
    uint4 u = __ldg(reinterpret_cast<const uint4*>(p.residual + off));
    const __half2* rh = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int k = 0; k < 4; ++k) { float2 f = __half22float2(rh[k]); y[2 * k] += f.x; y[2 * k + 1] += f.y; }
  }
  if (p.relu) {
#pragma unroll
    for (int o = 0; o < 8; ++o) y[o] = fmaxf(y[o], 0.f);
  }
  if (p.out_f32) {
    float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + off);
    op[0] = make_float4(y[0], y[1], y[2], y[3]);
    op[1] = make_float4(y[4], y[5], y[6], y[7]);
  } else {
    uint4 o4;
    __half2* oh2 = reinterpret_cast<__half2*>(&o4);
#pragma unroll
    for (int k = 0; k < 4; ++k) oh2[k] = __floats2half2_rn(y[2 * k], y[2 * k + 1]);
    *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.out) + off) = o4;
  }
}

cudaError_t launch_conv_simt(const ConvSimtParams& p, cudaStream_t st) {
  const long total = (long)p.N * p.OH * p.OW * (p.Cout / 8);
  if (total == 0) return cudaSuccess;
  conv_simt_kernel<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(p);
  return cudaGetLastError();
}


// ------------------------------------------------------------------------------------------------ cubic resize (uint8)
// OpenCV's own INTER_CUBIC kernel for 8-bit images (imgproc/src/resize.cpp: HResizeCubic + VResizeCubic), which is what
// `cv2.resize(image, (W, H), interpolation=cv2.INTER_CUBIC)` of reference SimpleHRNet.py:216-220 / :356-360 computes when
// OpenCV's vendor-optimised (IPP) path is off: per axis four taps sx-1 .. sx+2 (indices clamped to the image), Keys cubic
// weights (A = -0.75) rounded to 11-bit fixed point; the horizontal pass is exact in int32.  The vertical pass has two
// forms in OpenCV and both are restated: the vector loop (VResizeCubicVec_32s8u, the first 8 * floor(3 * dw / 8)
// elements of a row with the baseline 128-bit lanes) evaluates s0*b0 + (s1*b1 + (s2*b2 + s3*b3)) in float32 with
// b = weight / 2^22, unfused, rounds to nearest-even and saturates; the scalar tail adds 1 << 21 to the int32 sum,
// shifts by 22 and saturates.  The tap tables are built on the host in float32 exactly like OpenCV does
// (simple_hrnet_b200/preprocess.py).
//   src [n, sh, sw, 3] uint8, dst [n, dh, dw, 3] uint8; xofs [dw] / yofs [dh] first tap (sx - 1), xcoef [dw][4] / ycoef [dh][4]
__global__ void __launch_bounds__(256)
resize_cubic_u8_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int n, int sh, int sw, int dh, int dw,
                       const int32_t* __restrict__ xofs, const int16_t* __restrict__ xcoef,
                       const int32_t* __restrict__ yofs, const int16_t* __restrict__ ycoef) {
  const long total = (long)n * dh * dw;
  const long i = (long)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  const int dx = (int)(i % dw);
  const long t = i / dw;
  const int dy = (int)(t % dh);
  const int img = (int)(t / dh);
  const int x0 = xofs[dx], y0 = yofs[dy];
  int xa[4], xi[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) { xa[k] = xcoef[dx * 4 + k]; xi[k] = min(max(x0 + k, 0), sw - 1); }
  int h[4][3];
  const uint8_t* base = src + (size_t)img * sh * sw * 3;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int yy = min(max(y0 + r, 0), sh - 1);
    const uint8_t* row = base + (size_t)yy * sw * 3;
    h[r][0] = h[r][1] = h[r][2] = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint8_t* px = row + xi[k] * 3;
      h[r][0] += (int)px[0] * xa[k]; h[r][1] += (int)px[1] * xa[k]; h[r][2] += (int)px[2] * xa[k];
    }
  }
  int yb[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) yb[r] = ycoef[dy * 4 + r];
  const int nvec = (3 * dw) & ~7;          // elements of a row the 8-lane vector loop covers
  const float scale = 1.f / (2048.f * 2048.f);
  uint8_t* o = dst + (size_t)i * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    int v;
    if (3 * dx + c < nvec) {
      float f = __fmul_rn((float)h[3][c], __fmul_rn((float)yb[3], scale));
      f = __fadd_rn(__fmul_rn((float)h[2][c], __fmul_rn((float)yb[2], scale)), f);
      f = __fadd_rn(__fmul_rn((float)h[1][c], __fmul_rn((float)yb[1], scale)), f);
      f = __fadd_rn(__fmul_rn((float)h[0][c], __fmul_rn((float)yb[0], scale)), f);
      v = __float2int_rn(f);
    } else {
      v = (h[0][c] * yb[0] + h[1][c] * yb[1] + h[2][c] * yb[2] + h[3][c] * yb[3] + (1 << 21)) >> 22;
    }
    o[c] = (uint8_t)min(max(v, 0), 255);
  }
}

cudaError_t launch_resize_cubic_u8(const uint8_t* src, uint8_t* dst, int n, int sh, int sw, int dh, int dw, const int32_t* xofs,
                                   const int16_t* xcoef, const int32_t* yofs, const int16_t* ycoef, cudaStream_t st) {
  const long total = (long)n * dh * dw;
  if (total == 0) return cudaSuccess;
  resize_cubic_u8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(src, dst, n, sh, sw, dh, dw, xofs, xcoef, yofs, ycoef);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ multi-person crops
// `ToPILImage -> Resize((H, W))` of the reference's crop transform (SimpleHRNet.py:166-171) = Pillow's antialiased
// bilinear resample of an 8-bit image (libImaging/Resample.c: horizontal pass rounded to uint8, then vertical pass; both
// with 22-bit fixed-point coefficients and a rounding shift) applied to `image[y1:y2, x1:x2]` (+ zero padding,
// SimpleHRNet.py:262-270).  One thread = one output pixel (3 channels); the horizontal pass of the rows a pixel needs is
// recomputed (ky x kx taps, a few dozen integer MACs).  desc[m][kCropDescInts]:
//   0 frame | 1,2 x0, y0: frame coordinates of the padded crop's top-left | 3..6 the frame rectangle [vx0, vx1) x [vy0, vy1)
//   that holds image data (zero outside) | 7,8 offsets (ints) of the x / y tables | 9,10 taps per output coordinate kx, ky
// tables (int32): per axis [out][2] (first source index, tap count) followed by [out][k] coefficients
// (simple_hrnet_b200.preprocess.pil_bilinear_tables restates Pillow's precompute_coeffs / normalize_coeffs_8bpc).
constexpr int kCropDescInts = 12;
__global__ void __launch_bounds__(256)
crop_resize_bilinear_u8_kernel(const uint8_t* __restrict__ frames, int FH, int FW, const int32_t* __restrict__ desc,
                               const int32_t* __restrict__ tables, uint8_t* __restrict__ out, int OH, int OW) {
  const int m = blockIdx.y;
  const int t = (int)blockIdx.x * 256 + (int)threadIdx.x;
  if (t >= OH * OW) return;
  const int oy = t / OW, ox = t - oy * OW;
  const int32_t* d = desc + (size_t)m * kCropDescInts;
  const int frame = d[0], x0 = d[1], y0 = d[2], vx0 = d[3], vy0 = d[4], vx1 = d[5], vy1 = d[6];
  const int kx = d[9], ky = d[10];
  const int32_t* bx = tables + d[7];
  const int32_t* cx = bx + 2 * OW + (size_t)ox * kx;
  const int32_t* by = tables + d[8];
  const int32_t* cy = by + 2 * OH + (size_t)oy * ky;
  const int xmin = bx[2 * ox], nx = bx[2 * ox + 1];
  const int ymin = by[2 * oy], ny = by[2 * oy + 1];
  const uint8_t* img = frames + (size_t)frame * FH * FW * 3;
  constexpr int kHalf = 1 << 21;
  int acc[3] = {kHalf, kHalf, kHalf};
  for (int ty = 0; ty < ny; ++ty) {
    const int fy = y0 + ymin + ty;
    int h[3] = {kHalf, kHalf, kHalf};
    if (fy >= vy0 && fy < vy1) {
      const uint8_t* row = img + (size_t)fy * FW * 3;
      for (int tx = 0; tx < nx; ++tx) {
        const int fx = x0 + xmin + tx;
        if (fx >= vx0 && fx < vx1) {
          const int k = __ldg(cx + tx);
          h[0] += (int)__ldg(row + 3 * fx + 0) * k;
          h[1] += (int)__ldg(row + 3 * fx + 1) * k;
          h[2] += (int)__ldg(row + 3 * fx + 2) * k;
        }
      }
    }
    const int k = __ldg(cy + ty);
#pragma unroll
    for (int c = 0; c < 3; ++c) acc[c] += min(max(h[c] >> 22, 0), 255) * k;
  }
  uint8_t* o = out + (((size_t)m * OH + oy) * OW + ox) * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) o[c] = (uint8_t)min(max(acc[c] >> 22, 0), 255);
}

cudaError_t launch_crop_resize_bilinear_u8(const uint8_t* frames, int FH, int FW, const int32_t* desc, const int32_t* tables,
                                           int m, uint8_t* out, int OH, int OW, cudaStream_t st) {
  if (m == 0) return cudaSuccess;
  crop_resize_bilinear_u8_kernel<<<dim3((unsigned)((OH * OW + 255) / 256), (unsigned)m), 256, 0, st>>>(frames, FH, FW, desc, tables,
                                                                                                      out, OH, OW);
  return cudaGetLastError();
}

}  // namespace hrnet
