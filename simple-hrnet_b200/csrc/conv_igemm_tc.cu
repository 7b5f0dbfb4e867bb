// Single-problem launches of the TMA-im2col implicit-GEMM conv (body and documentation: conv_igemm_body.cuh).
#include <algorithm>

#include "conv_igemm_body.cuh"

namespace hrnet {

template <bool kPair, int kEpi>
__global__ void __launch_bounds__(kThreads, 1)
conv_igemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmR,
                     const ConvTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  conv_igemm_body<kPair, kEpi>(tmA, tmB, &tmO, &tmR, p, (int)blockIdx.x, (int)gridDim.x, smem_raw);
}

template <bool kPair, int kEpi>
static cudaError_t set_attr(int max_smem) {
  return cudaFuncSetAttribute(conv_igemm_tc_kernel<kPair, kEpi>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
}
cudaError_t conv_tc_set_attributes(int max_smem) {
  cudaError_t e = set_attr<false, 0>(max_smem);
  if (e == cudaSuccess) e = set_attr<false, 1>(max_smem);
  if (e == cudaSuccess) e = set_attr<false, 2>(max_smem);
  if (e == cudaSuccess) e = set_attr<false, 3>(max_smem);
  if (e == cudaSuccess) e = set_attr<true, 0>(max_smem);
  if (e == cudaSuccess) e = set_attr<true, 1>(max_smem);
  if (e == cudaSuccess) e = set_attr<true, 2>(max_smem);
  return e;
}

// grid size for a persistent launch: as many clusters as can be co-resident (cudaOccupancyMaxActiveClusters), capped
// by the work available
int conv_tc_grid(const ConvTcParams& p, int smem_bytes, int num_sms) {
  const int cs = p.cs;
  const int m_super = (p.m_tiles + cs - 1) / cs;
  const int need = m_super * p.n_tiles;
  int max_clusters = num_sms / cs;
  if (cs > 1) {
    static int cached[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (cached[cs] == 0) {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3((unsigned)(num_sms / cs * cs));
      cfg.blockDim = dim3(kThreads);
      cfg.dynamicSmemBytes = 200 * 1024;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = (unsigned)cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int nc = 0;
      if (cudaOccupancyMaxActiveClusters(&nc, conv_igemm_tc_kernel<true, 0>, &cfg) == cudaSuccess && nc > 0) cached[cs] = nc;
      else { cudaGetLastError(); cached[cs] = num_sms / cs; }
    }
    max_clusters = std::min(cached[cs], num_sms / cs);   // num_sms may be a caller-imposed cap
  }
  return std::min(need, max_clusters) * cs;
}

cudaError_t launch_conv_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap* tmOR, const ConvTcParams& p,
                           int smem_bytes, int grid, cudaStream_t st) {
  // without the staged epilogue the two extra maps are never dereferenced: pass any valid descriptor
  const CUtensorMap& tmO = p.epi_tma == 1 ? tmOR[0] : tmA;
  const CUtensorMap& tmR = p.epi_tma == 1 ? tmOR[1] : tmA;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = (size_t)smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  int na = 0;
  if (p.cs > 1) {
    at[na].id = cudaLaunchAttributeClusterDimension;
    at[na].val.clusterDim.x = (unsigned)p.cs; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
    ++na;
  }
  if (p.pdl) {   // prologue overlaps the tail of the previous kernel in the stream (ptx::pdl_wait)
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = at; cfg.numAttrs = (unsigned)na;
  if (p.cs > 1) {
    if (p.epi_tma == 1) return cudaLaunchKernelEx(&cfg, conv_igemm_tc_kernel<true, 1>, tmA, tmB, tmO, tmR, p);
    if (p.epi_tma == 2) return cudaLaunchKernelEx(&cfg, conv_igemm_tc_kernel<true, 2>, tmA, tmB, tmO, tmR, p);
    return cudaLaunchKernelEx(&cfg, conv_igemm_tc_kernel<true, 0>, tmA, tmB, tmO, tmR, p);
  }
  if (p.epi_tma == 1) return cudaLaunchKernelEx(&cfg, conv_igemm_tc_kernel<false, 1>, tmA, tmB, tmO, tmR, p);
  if (p.epi_tma == 2) return cudaLaunchKernelEx(&cfg, conv_igemm_tc_kernel<false, 2>, tmA, tmB, tmO, tmR, p);
  if (p.epi_tma == 3) return cudaLaunchKernelEx(&cfg, conv_igemm_tc_kernel<false, 3>, tmA, tmB, tmO, tmR, p);
  return cudaLaunchKernelEx(&cfg, conv_igemm_tc_kernel<false, 0>, tmA, tmB, tmO, tmR, p);
}

}  // namespace hrnet
