// Grouped launch: ONE persistent kernel runs up to four independent convolutions (the k-th conv of every branch of a
// StageModule, models_/hrnet.py:12-21) side by side, each on its own contiguous range of CTAs sized by estimated work.
//
// Why: launched one after the other, every branch conv pays its own prologue / pipeline fill / drain (~6 us inside the
// kernel plus the launch gap) and the low-resolution branches leave SMs idle (216 or 108 tiles for 148 SMs: 27 % of the
// machine waits for the last tile round).  Side by side the 4 problems share one prologue and the tile rounds of one
// branch fill the gaps of the others (profiles/r01_layer_breakdown_v7.txt: 157 us per 4-conv level before grouping).
//
// Slots: up to two halo-patch problems (high-resolution branches) and two TMA-im2col problems (low-resolution ones);
// each CTA picks its problem from its block index and runs that problem's unchanged single-problem body.
#include <cstdlib>

#include "conv3x3_patch_body.cuh"
#include "conv_igemm_body.cuh"

namespace hrnet {

struct GroupArgs {
  int cta_end[4];          // exclusive CTA prefix: [patch0 | patch1 | igemm0 | igemm1]
  PatchMaps pm[2];
  ConvPatchParams pp[2];
  CUtensorMap ia[2], ib[2];
  ConvTcParams ip[2];
};

__global__ void __launch_bounds__(384, 1) conv_group_kernel(const __grid_constant__ GroupArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const int b = (int)blockIdx.x;
  if (b < a.cta_end[0]) {
    conv3x3_patch_body<false, 0>(a.pm[0], a.pp[0], b, a.cta_end[0], smem_raw);
  } else if (b < a.cta_end[1]) {
    conv3x3_patch_body<false, 0>(a.pm[1], a.pp[1], b - a.cta_end[0], a.cta_end[1] - a.cta_end[0], smem_raw);
  } else if (b < a.cta_end[2]) {
    conv_igemm_body<false, 0>(a.ia[0], a.ib[0], nullptr, nullptr, a.ip[0], b - a.cta_end[1], a.cta_end[2] - a.cta_end[1], smem_raw);
  } else {
    conv_igemm_body<false, 0>(a.ia[1], a.ib[1], nullptr, nullptr, a.ip[1], b - a.cta_end[2], a.cta_end[3] - a.cta_end[2], smem_raw);
  }
}

cudaError_t conv_group_set_attributes(int max_smem) {
  return cudaFuncSetAttribute(conv_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
}

cudaError_t launch_conv_group(const GroupLaunch& g, cudaStream_t st) {
  GroupArgs a;
  int end = 0;
  for (int i = 0; i < 2; ++i) {
    if (i < g.n_patch) {
      a.pm[i].a = g.patch_maps_a[i][0];
      for (int k = 0; k < 3; ++k) a.pm[i].b[k] = g.patch_maps_b[i][k];
      a.pm[i].o = a.pm[i].r = a.pm[i].a;   // grouped launches use the direct-store epilogue
      a.pp[i] = g.pp[i];
      end += g.patch_ctas[i];
    } else {
      a.pm[i] = a.pm[0];
      a.pp[i] = g.pp[0];
    }
    a.cta_end[i] = end;
  }
  for (int i = 0; i < 2; ++i) {
    if (i < g.n_igemm) {
      a.ia[i] = g.igemm_map_a[i]; a.ib[i] = g.igemm_map_b[i]; a.ip[i] = g.ip[i];
      end += g.igemm_ctas[i];
    } else {
      a.ia[i] = g.igemm_map_a[0]; a.ib[i] = g.igemm_map_b[0]; a.ip[i] = g.ip[0];
    }
    a.cta_end[2 + i] = end;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)end);
  cfg.blockDim = dim3(384);
  cfg.dynamicSmemBytes = (size_t)g.smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  int na = 0;
  if (g.pdl) {
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = at; cfg.numAttrs = (unsigned)na;
  return cudaLaunchKernelEx(&cfg, conv_group_kernel, a);
}

}  // namespace hrnet
