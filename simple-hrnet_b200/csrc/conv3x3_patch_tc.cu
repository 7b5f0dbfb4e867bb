// Single-problem launch of the halo-patch 3x3 conv (body and documentation: conv3x3_patch_body.cuh).
#include <algorithm>

#include "conv3x3_patch_body.cuh"

namespace hrnet {

template <bool kPair, int kEpi>
__global__ void __launch_bounds__(kPThreads, 1)
conv3x3_patch_tc_kernel(const __grid_constant__ PatchMaps maps, const ConvPatchParams p) {
  extern __shared__ uint8_t smem_raw[];
  conv3x3_patch_body<kPair, kEpi>(maps, p, (int)blockIdx.x, (int)gridDim.x, smem_raw);
}

template <bool kPair, int kEpi>
static cudaError_t set_attr(int max_smem) {
  return cudaFuncSetAttribute(conv3x3_patch_tc_kernel<kPair, kEpi>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
}
cudaError_t conv_patch_set_attributes(int max_smem) {
  cudaError_t e = set_attr<false, 0>(max_smem);
  if (e == cudaSuccess) e = set_attr<false, 1>(max_smem);
  if (e == cudaSuccess) e = set_attr<false, 2>(max_smem);
  if (e == cudaSuccess) e = set_attr<false, 3>(max_smem);
  if (e == cudaSuccess) e = set_attr<true, 0>(max_smem);
  if (e == cudaSuccess) e = set_attr<true, 1>(max_smem);
  if (e == cudaSuccess) e = set_attr<true, 2>(max_smem);
  return e;
}

// persistent grid: one CTA per SM, or as many CTA pairs as can be co-resident (cudaOccupancyMaxActiveClusters)
int conv_patch_grid(const ConvPatchParams& p, int num_sms) {
  if (p.cs < 2) return std::min(p.total_tiles, num_sms);
  static int max_clusters = 0;
  if (max_clusters == 0) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(num_sms / 2 * 2));
    cfg.blockDim = dim3(kPThreads);
    cfg.dynamicSmemBytes = 200 * 1024;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int nc = 0;
    if (cudaOccupancyMaxActiveClusters(&nc, conv3x3_patch_tc_kernel<true, 0>, &cfg) == cudaSuccess && nc > 0) max_clusters = nc;
    else { cudaGetLastError(); max_clusters = num_sms / 2; }
  }
  const int pairs = (p.total_tiles + 1) / 2;
  return std::min(pairs, std::min(max_clusters, num_sms / 2)) * 2;
}

cudaError_t launch_conv_patch(const CUtensorMap* tmA3, const CUtensorMap* tmB3, const CUtensorMap* tmOR,
                              const ConvPatchParams& p, int smem_bytes, int grid, cudaStream_t st) {
  PatchMaps m;
  m.a = tmA3[0];
  for (int i = 0; i < 3; ++i) m.b[i] = tmB3[i];
  m.o = p.epi_tma == 1 ? tmOR[0] : tmA3[0];   // never dereferenced without the staged epilogue
  m.r = p.epi_tma == 1 ? tmOR[1] : tmA3[0];
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kPThreads);
  cfg.dynamicSmemBytes = (size_t)smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  int na = 0;
  if (p.cs > 1) {
    at[na].id = cudaLaunchAttributeClusterDimension;
    at[na].val.clusterDim.x = 2; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
    ++na;
  }
  if (p.pdl) {   // resident-weight loads and the rest of the prologue overlap the previous kernel's tail
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = at; cfg.numAttrs = (unsigned)na;
  if (p.cs > 1) {
    if (p.epi_tma == 1) return cudaLaunchKernelEx(&cfg, conv3x3_patch_tc_kernel<true, 1>, m, p);
    if (p.epi_tma == 2) return cudaLaunchKernelEx(&cfg, conv3x3_patch_tc_kernel<true, 2>, m, p);
    return cudaLaunchKernelEx(&cfg, conv3x3_patch_tc_kernel<true, 0>, m, p);
  }
  if (p.epi_tma == 1) return cudaLaunchKernelEx(&cfg, conv3x3_patch_tc_kernel<false, 1>, m, p);
  if (p.epi_tma == 2) return cudaLaunchKernelEx(&cfg, conv3x3_patch_tc_kernel<false, 2>, m, p);
  if (p.epi_tma == 3) return cudaLaunchKernelEx(&cfg, conv3x3_patch_tc_kernel<false, 3>, m, p);
  return cudaLaunchKernelEx(&cfg, conv3x3_patch_tc_kernel<false, 0>, m, p);
}

}  // namespace hrnet
