// Internal types shared by the plan builder, the kernel launchers and the C-ABI layer.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

namespace hrnet {

enum OpKind : int {
  OP_STEM = 0,     // 3x3 s2 conv on the NCHW fp32 network input -> NHWC fp16 (+BN+ReLU), SIMT
  OP_CONV = 1,     // kxk conv (k in {1,3}, stride in {1,2}) NHWC fp16 -> NHWC fp16/fp32, BN, +residual, ReLU
  OP_FUSE = 2,     // out = ReLU(sum_j nearest_up(src_j)) (exchange-unit sum, hrnet.py:60-69)
  OP_HEAD = 3,     // 1x1 conv + bias, NHWC fp16 -> NCHW fp32 heatmaps
  OP_ARGMAX = 4,   // per-joint argmax decode -> (y, x, conf) + flat index
  OP_MAXPOOL = 5,  // 3x3 s2 p1 max pool NHWC fp16 (PoseResNet)
  OP_STEM7 = 6,    // 7x7 s2 p3 conv on the NCHW fp32 input (PoseResNet)
};

enum DType : int { DT_F16 = 0, DT_F32 = 1 };

struct TensorInfo {
  int C = 0, H = 0, W = 0;
  int dtype = DT_F16;
  size_t offset = 0;  // byte offset in the activation workspace; (size_t)-1 for external tensors
  size_t bytes(int n) const { return (size_t)n * C * H * W * (dtype == DT_F16 ? 2 : 4); }
};

// One weight-carrying layer: what the Python packer must write where.
struct ParamInfo {
  std::string conv_key;  // state_dict prefix of the conv ("stage2.0.branches.0.0.conv1"); weight = key + ".weight"
  std::string bn_key;    // state_dict prefix of the BN ("...bn1"), empty -> scale = 1, bias = conv bias (or 0)
  int cout = 0, cin = 0, kh = 0, kw = 0;
  int transposed = 0;    // 1: ConvTranspose2d weight [Cin, Cout, kh, kw]
  int has_bias = 0;      // conv has its own bias (final_layer)
  int w_f32 = 0;         // weights stored as fp32 (stem conv1, head) instead of fp16
  size_t w_offset = 0;   // bytes into the packed weight buffer: [cout][kh][kw][cin]
  size_t scale_offset = 0, bias_offset = 0;  // fp32 [cout] each
};

struct ConvTcCfg {
  int kc = 0;        // channels per k-block (64 / 32 / 16) -> swizzle 128 / 64 / 32 B
  int bps = 0;       // k-blocks per pipeline stage
  int n_tile = 0;    // output channels per CTA tile
  int stages = 0;
  int smem_bytes = 0;
  int tmem_cols = 0;
  int cs = 1;        // 2: CTA-pair mode (cta_group::2)
  int epi = 0;       // epilogue: 0 direct stores, 1 staged TMA stores, 2 warp-staged coalesced stores
  int epi_bytes = 0; // shared memory of the epilogue staging tiles
  int mma_warps = 1; // 2: two MMA-issuing warps on alternate tiles
};


// warp-staged coalesced epilogue (epilogue.cuh): staging tile of one epilogue warp
constexpr int kCoalPitch = 144;                       // bytes per staged row (64 fp16 channels + 16 B pad)
constexpr int kCoalWarpBytes = 32 * kCoalPitch;       // 4608 B per epilogue warp

// device-side parameter block for the tcgen05 implicit-GEMM conv
struct ConvTcParams {
  int M_total, OH, OW, OHW;
  int ksize, stride, pad_h, pad_w;   // pad_* = leading (top / left) padding
  int sub, sub_a, sub_b;             // transposed-conv sub-pixel phase: out pixel (2i+a, 2j+b)
  int Cin, Cout;
  int kc, cpt, nkb, bps;
  int n_tile, n_tiles, m_tiles;
  int cs;            // cluster size along M (weights multicast): 1, 2 or 4
  int stages;
  int mma_warps;     // 1, or 2: warps 2 and 3 issue the MMAs of alternate tiles
  int relu, out_f32;
  int tmem_cols;
  int a_blk_bytes, b_blk_bytes;
  int epi_tma, epi_bytes;   // epilogue (epilogue.cuh): 0 direct, 1 staged TMA stores, 2 warp-staged coalesced; staging follows the stages
  int pdl;                  // host side: launch with programmatic stream serialization
  const float* scale;
  const float* bias;
  const __half* residual;
  void* out;
  long long* dbg;    // optional per-CTA role timers (16 x int64 per CTA), nullptr in production
};

// device-side parameter block for the halo-patch 3x3 stride-1 conv (conv3x3_patch_tc.cu)
constexpr int kPatchTW = 8, kPatchTH = 16;                 // output tile: 8 wide x 16 high = 128 GEMM rows
constexpr int kPatchPW = kPatchTW + 2, kPatchPH = kPatchTH + 2;  // input patch with the 1-pixel halo
constexpr int kPatchRows = kPatchPW * kPatchPH;            // 180 pixels
struct ConvPatchParams {
  int N, H, W, Cin, Cout;
  int tiles_w, tiles_h, total_tiles;
  int nchunks;                 // channel chunks of the K dimension: 64-channel patch slots (zero-filled tail)
  int c0[4], kreal[4];         // first channel, real channels in the chunk (multiple of 16)
  int bkc[4], mapi[4];         // weight block width (64 / 32 / 16) and its tensor-map index (0 / 1 / 2)
  int boff[4], bblk[4];        // resident-weight block offset / per-tap block size in shared memory
  int b_bytes;                 // shared memory reserved for the resident weights (0 when streamed)
  int b_stream;                // 1: the 9-tap weight block of a chunk travels with its patch slot (weights too big to stay)
  int a_slot_bytes;            // bytes of the patch part of a slot (weights of a streamed chunk follow it)
  int slot_bytes, nslots;      // ring of patch slots (one channel chunk of one tile each)
  int nacc, nacc_log2;         // TMEM accumulator buffers (2 or 4)
  int mma_warps;               // 1, or 2: warps 2 and 3 issue the MMAs of alternate tiles (two MMA -> epilogue pipelines)
  int cs;                      // 2: CTA-pair mode (cta_group::2), each CTA keeps Cout / 2 weight rows; else 1
  int relu, out_f32, tmem_cols;
  int epi_tma, epi_bytes;      // epilogue: 0 direct, 1 staged TMA stores, 2 warp-staged coalesced; staging follows the patch slots
  int pdl;                     // host side: launch with programmatic stream serialization
  const float* scale;
  const float* bias;
  const __half* residual;
  void* out;
  long long* dbg;    // optional per-CTA role timers (16 x int64 per CTA), nullptr in production
};

// device-side parameter block for the SIMT fallback conv (debug path / odd shapes)
struct ConvSimtParams {
  int N, IH, IW, OH, OW, Cin, Cout, ksize, stride, pad, relu, out_f32;
  const __half* in;
  const __half* w;  // [Cout][k][k][Cin]
  const float* scale;
  const float* bias;
  const __half* residual;
  void* out;
};

// ---- branch chains (conv_chain.cu): the eight 3x3 convs of one StageModule branch in ONE persistent kernel ----------
// Tiles of all convs form one ordered ticket sequence (conv-major); a CTA draws tickets from a global counter and a tile
// of conv k starts once the tiles of conv k-1 covering its 3x3 halo carry this launch's epoch stamp.
constexpr int kChainMaxConv = 8;
struct ChainConv {
  const float* scale;
  const float* bias;
  const __half* residual;    // nullptr or NHWC fp16 (written earlier in the same launch: read with ld.global.cg)
  __half* out;
  int relu;
  int pad_;
};
// control words of one chain in global memory: [0] ticket counter, [1] exited-CTA counter.
// counters[conv * unit_stride + unit] = finished tiles of that unit (a row of 8x16 tiles of one image / a 128-pixel
// M-tile); the last CTA of a launch clears what the launch used.
constexpr int kChainMaxCoutI = 384;   // widest im2col chain / halo-patch chain whose BN constants travel in the kernel parameters
constexpr int kChainMaxCoutP = 128;
struct ChainIgemmParams {
  int nconv;
  int M_total, OH, OW, OHW, C;
  int cpt, nkb, bps, n_tile, n_tiles, m_tiles, stages, tmem_cols, a_blk_bytes, b_blk_bytes;
  int m2;                    // M-tiles per ticket: 1, or 2 (both tiles share every weight k-block of a pipeline stage)
  int skip;                  // experiments only (HRNET_TUNE_CHAIN_SKIP, results invalid): 1 no epilogue work, 2 accumulator loads
                             // only, 3 residual loads + stores only
  int pair;                  // CTA pairs (cta_group::2): a unit = two M-tiles, one per CTA of the cluster (m2 == 1 then)
  int units;                 // tickets per conv = ceil(m_tiles / tiles per unit)
  int unit_stride;           // counters per conv (units at max batch)
  int pdl;                   // host side: launch with programmatic stream serialization
  unsigned* ctrl;
  unsigned* counters;
  long long* dbg;            // optional per-CTA counters (8 x int64 per CTA)
  ChainConv conv[kChainMaxConv];
  float2 sb[kChainMaxConv][kChainMaxCoutI];   // (BN scale, bias) per conv and output channel: read as constants (LDC) by the epilogue
};
struct ChainIgemmMaps { CUtensorMap a[kChainMaxConv], b[kChainMaxConv]; };
struct ChainPatchParams {
  int nconv;
  ConvPatchParams pp;        // geometry / shared-memory layout common to all convs of the chain (scale .. out unused)
  int unit_stride;           // counters per conv (tile rows at max batch)
  int chunk;                 // tiles per ticket: a divisor of pp.tiles_w (neighbouring tiles of one tile row)
  int skip;                  // experiments only (HRNET_TUNE_CHAIN_SKIP): 9 = the general MMA issue loop
  int tail64;                // two-chunk chains whose second chunk has 32 channels: THREE slots -- two 128-byte-row slots for chunk 0
                             // (alternate tiles) and one 64-byte-row slot for chunk 1 (conv_chain.cu)
  int pdl;                   // host side: launch with programmatic stream serialization
  unsigned* ctrl;
  unsigned* counters;
  long long* dbg;
  ChainConv conv[kChainMaxConv];
  float2 sb[kChainMaxConv][kChainMaxCoutP];   // (BN scale, bias) per conv and output channel (see ChainIgemmParams)
};
struct ChainPatchMaps { CUtensorMap a[kChainMaxConv]; CUtensorMap b[kChainMaxConv][2]; CUtensorMap a2[kChainMaxConv]; };   // a2: 32-channel patch box (tail64)
cudaError_t launch_chain_igemm(const ChainIgemmMaps& maps, const ChainIgemmParams& p, int smem_bytes, int grid, cudaStream_t st);
cudaError_t launch_chain_patch(const ChainPatchMaps& maps, const ChainPatchParams& p, int smem_bytes, int grid, cudaStream_t st);
cudaError_t conv_chain_set_attributes(int max_smem);

// ---- exchange unit (conv_xunit.cu): every conv of one StageModule's fuse layers (reference models_/hrnet.py:23-51) in ONE
// persistent kernel: tickets = M-tiles of the member convs in dependency order, a conv of a down-chain starts when its
// producer conv is complete.
constexpr int kXMaxOps = 16;
constexpr int kXMaxSb = 2560;      // output channels of all member convs together (W48 stage 4: 2,400)
struct XOp {
  int M_total, OH, OW, OHW;      // output geometry (flattened NHWC rows)
  int ksize, stride, pad;
  int Cin, Cout, cpt, nkb;       // channel blocks per tap, k-blocks per tile
  int n_tile, n_tiles, m_tiles;
  int relu;
  int dep, dep_need;             // producer op within the unit (-1: a module input) and its tile count
  int kb0;                       // first entry of this op in the k-block table
  int ticket0;                   // first ticket (one ticket = one M-tile, all of its N-tiles)
  int sb_off;                    // first (scale, bias) pair of this op in XUnitParams::sb
  int pad_;
  const float* scale;
  const float* bias;
  __half* out;
};
// one output of the exchange unit: out = ReLU(sum_j nearest_up(src_j)), models_/hrnet.py:60-69 (fp32 sum in ascending j)
constexpr int kXSumChunk = 1024;   // output pixels per sum ticket
struct XSum {
  int H, W, C, nsrc, relu;
  int shift[4];
  int dep[4], dep_need[4];       // member conv producing src j (-1: a module input, e.g. the identity term) and its tile count
  long long npix;                // N * H * W
  int ticket0, nchunks;
  const __half* src[4];
  __half* out;
};
struct XUnitParams {
  int nops, total_tickets, total_kb;
  int nsums;
  XSum sum[4];
  int stages, tmem_cols, a_blk_bytes, b_blk_bytes, pdl;
  unsigned* ctrl;                // [0] ticket counter, [1] exited-CTA counter
  unsigned* counters;            // [nops] finished tiles per op
  long long* dbg;
  XOp op[kXMaxOps];
  float2 sb[kXMaxSb];            // BN (scale, bias) of every member conv's output channels: constants (LDC) for the epilogue
};
struct XUnitMaps { CUtensorMap a[kXMaxOps], b[kXMaxOps]; };
cudaError_t launch_xunit(const XUnitMaps& maps, const XUnitParams& p, int smem_bytes, int grid, cudaStream_t st);
cudaError_t conv_xunit_set_attributes(int max_smem);

struct FuseParams {
  int N, H, W, C, nsrc, relu;
  const void* src[4];
  int shift[4];
  int f32[4];
  __half* out;
};

struct Op {
  int kind = 0;
  std::string name;
  int in = -1, out = -1, res = -1;  // tensor ids
  int param = -1;
  int cin = 0, cout = 0, k = 1, stride = 1, pad = 0;
  int relu = 0;
  // fuse
  int nsrc = 0;
  int src[4] = {-1, -1, -1, -1};
  int shift[4] = {0, 0, 0, 0};
  // scheduling
  int stream = 0;
  std::vector<int> deps;       // ops that must complete before this one
  bool needs_event = false;    // some dependant runs on another stream
  int grp = -1;                // grouped launch id: the k-th convs of all branches of a StageModule run in one kernel
  double work = 0;             // estimated SM-cycles (CTA split of grouped launches)
  float sm_frac = 1.f;         // share of the SMs this op's persistent grid may occupy (branch-level SM partitioning)
  int group = -1;              // ops with the same group id run concurrently on different streams and split the SMs
  int chain = -1, chain_pos = 0;   // branch chain (conv_chain.cu) this conv belongs to / its position in it
  int xunit = -1, xpos = 0, xlevel = 0;   // exchange unit (conv_xunit.cu) / position in its ticket order / step in its down-chain
  // tcgen05 path
  bool use_tc = false;
  ConvTcCfg tc;
  CUtensorMap tmA, tmB;
  // halo-patch path (3x3 stride 1 with shared-memory-resident weights)
  bool use_patch = false;
  ConvPatchParams pp{};
  int patch_smem = 0;
  CUtensorMap tmPA[3], tmPB[3];
  // TMA-store epilogue: output / residual maps (2-D {C, pixels} for the im2col kernel, 4-D {C, W, H, N} for the patch kernel)
  CUtensorMap tmOR[2];
};

// host-side description of one grouped launch (conv_group.cu): up to 2 halo-patch + 2 im2col problems
struct GroupLaunch {
  int n_patch = 0, n_igemm = 0;
  int patch_ctas[2] = {0, 0}, igemm_ctas[2] = {0, 0};
  const CUtensorMap* patch_maps_a[2] = {nullptr, nullptr};   // -> Op::tmPA
  const CUtensorMap* patch_maps_b[2] = {nullptr, nullptr};   // -> Op::tmPB
  ConvPatchParams pp[2];
  CUtensorMap igemm_map_a[2], igemm_map_b[2];
  ConvTcParams ip[2];
  int smem_bytes = 0;
  int pdl = 1;
};
cudaError_t launch_conv_group(const GroupLaunch& g, cudaStream_t st);
cudaError_t conv_group_set_attributes(int max_smem);

// launchers (implemented in the .cu files)
cudaError_t launch_conv_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap* tmOR, const ConvTcParams& p,
                           int smem_bytes, int grid, cudaStream_t st);
cudaError_t launch_conv_simt(const ConvSimtParams& p, cudaStream_t st);
cudaError_t launch_stem(const float* in_nchw, const float* w, const float* scale, const float* bias, __half* out,
                        int N, int H, int W, cudaStream_t st);
cudaError_t stem_tc_set_attributes();     // once per device, before the first launch (hrnet_plan_bind)
cudaError_t launch_stem_tc(const float* in_nchw, const float* w, const float* scale, const float* bias, __half* out,
                           int N, int H, int W, int num_sms, cudaStream_t st);
cudaError_t launch_stem_tc_u8(const uint8_t* in_nhwc_bgr, const float* w, const float* scale, const float* bias,
                              __half* out, int N, int H, int W, int num_sms, cudaStream_t st);
cudaError_t launch_fuse(const FuseParams& p, cudaStream_t st);
// Head 1x1 conv with the weights as kernel parameters (constant bank) and, optionally, the per-joint argmax fused in:
// each 128-pixel block reduces its 17 maxima to one (value, index) candidate per joint (`pval` / `pidx`,
// [N][J][nblk]); head_argmax_finish reduces the candidates and decodes the joints -- the heat-maps are then never
// written (72 MB) nor re-read (30 MB) at W48 / 64 crops.  `out` (NCHW fp32 heat-maps) and the partials are each optional.
constexpr int kHeadMaxW = 6144;      // J * Cin floats that fit the parameter block (W48: 816, PoseResNet-50: 4,352)
constexpr int kHeadMaxJc = 32;
struct HeadParams {
  const __half* in;
  float* out;
  float* pval;
  int* pidx;
  int N, HW, Cin, J, nblk, pad_;
  const float* w_dev;                // [J][Cin] fp32 weights in the plan's weight buffer (staged in shared memory per block)
  float bias[kHeadMaxJc];
};
int head_c_blocks(int hw);                   // argmax candidates per (person, joint): blocks of 512 pixels
bool head_c_supported(int cin, int nj);     // shapes with a compiled specialisation (HRNet-W48 / W32, 17 joints)
cudaError_t launch_head_c(const HeadParams& p, cudaStream_t st);
cudaError_t launch_head_argmax_finish(const float* pval, const int* pidx, int N, int J, int nblk, int Hh, int Wh,
                                      const float* boxes, float* joints, int32_t* idx, cudaStream_t st);
cudaError_t launch_head(const __half* in, const float* w, const float* bias, float* out_nchw, int N, int HW, int Cin,
                        int J, cudaStream_t st);
cudaError_t launch_argmax(const float* hm, int N, int J, int Hh, int Wh, const float* boxes, float* joints,
                          int32_t* idx, cudaStream_t st);
cudaError_t launch_final_preds(const float* hm, int N, int J, int Hh, int Wh, int post, const double* trans, float* preds,
                               float* maxvals, cudaStream_t st);
cudaError_t launch_flip_average(const float* a, const float* b, float* out, const int* perm, int N, int J, int Hh, int Wh,
                                cudaStream_t st);
cudaError_t launch_crop_resize_bilinear_u8(const uint8_t* frames, int FH, int FW, const int32_t* desc, const int32_t* tables,
                                           int m, uint8_t* out, int OH, int OW, cudaStream_t st);
cudaError_t launch_resize_cubic_u8(const uint8_t* src, uint8_t* dst, int n, int sh, int sw, int dh, int dw, const int32_t* xofs,
                                   const int16_t* xcoef, const int32_t* yofs, const int16_t* ycoef, cudaStream_t st);
cudaError_t launch_maxpool(const __half* in, __half* out, int N, int IH, int IW, int C, cudaStream_t st);
cudaError_t launch_stem7(const float* in_nchw, const float* w, const float* scale, const float* bias, __half* out,
                         int N, int H, int W, cudaStream_t st);
cudaError_t conv_tc_set_attributes(int max_smem);
int conv_tc_grid(const ConvTcParams& p, int smem_bytes, int num_sms);
cudaError_t launch_conv_patch(const CUtensorMap* tmA3, const CUtensorMap* tmB3, const CUtensorMap* tmOR,
                              const ConvPatchParams& p, int smem_bytes, int grid, cudaStream_t st);
cudaError_t conv_patch_set_attributes(int max_smem);
int conv_patch_grid(const ConvPatchParams& p, int num_sms);

}  // namespace hrnet
