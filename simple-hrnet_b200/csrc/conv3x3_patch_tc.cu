// Single-problem launch of the halo-patch 3x3 conv (body and documentation: conv3x3_patch_body.cuh).
#include <cstdlib>

#include "conv3x3_patch_body.cuh"

namespace hrnet {

template <bool kEpiTma>
__global__ void __launch_bounds__(kPThreads, 1)
conv3x3_patch_tc_kernel(const __grid_constant__ PatchMaps maps, const ConvPatchParams p) {
  extern __shared__ uint8_t smem_raw[];
  conv3x3_patch_body<kEpiTma>(maps, p, (int)blockIdx.x, (int)gridDim.x, smem_raw);
}

cudaError_t conv_patch_set_attributes(int max_smem) {
  cudaError_t e = cudaFuncSetAttribute(conv3x3_patch_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(conv3x3_patch_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
}

cudaError_t launch_conv_patch(const CUtensorMap* tmA3, const CUtensorMap* tmB3, const CUtensorMap* tmOR,
                              const ConvPatchParams& p, int smem_bytes, int grid, cudaStream_t st) {
  PatchMaps m;
  m.a = tmA3[0];
  for (int i = 0; i < 3; ++i) m.b[i] = tmB3[i];
  m.o = p.epi_tma ? tmOR[0] : tmA3[0];   // never dereferenced without the staged epilogue
  m.r = p.epi_tma ? tmOR[1] : tmA3[0];
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kPThreads);
  cfg.dynamicSmemBytes = (size_t)smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  int na = 0;
  static int pdl = -1;
  if (pdl < 0) pdl = getenv("HRNET_B200_NO_PDL") ? 0 : 1;
  if (pdl) {   // resident-weight loads and the rest of the prologue overlap the previous kernel's tail
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = at; cfg.numAttrs = (unsigned)na;
  if (p.epi_tma) return cudaLaunchKernelEx(&cfg, conv3x3_patch_tc_kernel<true>, m, p);
  return cudaLaunchKernelEx(&cfg, conv3x3_patch_tc_kernel<false>, m, p);
}

}  // namespace hrnet
