// Plan = the static execution schedule of one network instance (arch, width, resolution, max batch):
// tensors placed in a caller-owned activation workspace, weight-carrying layers placed in a
// caller-owned packed weight buffer, an ordered op list with stream assignment and dependencies,
// TMA descriptors for every tensor-core conv, and a cached CUDA graph per batch size.
//
// The graph that is built restates the *structure* of the reference modules
//   HRNet.forward        models_/hrnet.py:157-189   (stem, layer1, transitions, stages, head)
//   StageModule.forward  models_/hrnet.py:55-71     (branches, exchange unit)
//   Bottleneck / BasicBlock  models_/modules.py:20-40 / 56-72
//   PoseResNet.forward   models_/poseresnet.py:108-122
// with BN folded to a per-channel fp32 (scale, bias) epilogue, ReLU / residual-add fused into the
// producing conv, and the four HRNet branches running concurrently on forked streams.
#include <algorithm>
#include <array>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <sstream>

#include "../../include/hrnet_b200.h"
#include "hrnet_internal.h"

namespace hrnet {

thread_local std::string g_last_error;
int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// tuning knobs of the single-op entry points (hrnet_debug_set_tune); plans carry their own copy in HrnetDesc::tune
int32_t g_single_op_tune[HRNET_TUNE_COUNT] = {0};

}  // namespace hrnet

using namespace hrnet;

// One branch chain (conv_chain.cu): the eight 3x3 convs of a StageModule branch issued as one persistent kernel.
struct ChainInfo {
  std::vector<int> ops;            // member convs, in execution order
  int module = -1, branch = 0;
  bool enabled = false;            // eligible and not disabled by a flag
  bool patch = false;              // halo-patch kernel (else im2col kernel)
  int smem = 0;
  int share = 0;                   // per mille of the SMs inside the forward: a module's chains run side by side
  int grid = 0;                    // ... resolved to CTAs by hrnet_plan_bind
  double cost = 0;                 // estimated SM-cycles of the whole chain (grid split)
  int flag_stride = 0;             // unit counters per conv (at max batch)
  int m2 = 1, stages = 0;          // im2col chain: M-tiles per ticket, pipeline stages
  bool pair = false;               // im2col chain on CTA pairs (cta_group::2): a unit = two M-tiles, one per CTA
  bool tail64 = false;             // halo-patch chain with a 64 + 32 channel split: three patch slots (conv_chain.cu)
  std::vector<float> sb;           // host copy of the member convs' BN (scale, bias) pairs [conv][cout][2], read back by
                                   // hrnet_plan_bind: they travel to the chain kernels as kernel parameters (constant bank)
  size_t ctrl_off = 0, flags_off = 0;   // bytes into the activation workspace
  ChainIgemmMaps imaps;            // filled by hrnet_plan_bind
  ChainPatchMaps pmaps;
};

// One exchange unit (conv_xunit.cu): every conv of a StageModule's fuse layers issued as one persistent kernel.
struct XUnitInfo {
  int stream = 0;                  // the launch runs on this stream (per-source units: the source branch's, right behind its chain)
  std::vector<float> sb;           // host copy of the member convs' BN (scale, bias) pairs, in member order (hrnet_plan_bind)
  std::vector<int> sums;           // the module's exchange sums (OP_FUSE ops), executed as sum tickets of the same kernel
  std::vector<int> ops;            // member convs in ticket (dependency) order
  int module = -1;
  bool enabled = false;
  int smem = 0, stages = 0, tmem_cols = 0, b_blk = 0;
  size_t ctrl_off = 0, counters_off = 0;
  XUnitMaps maps;                  // filled by hrnet_plan_bind
};

struct HrnetPlan {
  HrnetDesc desc{};
  std::vector<TensorInfo> tensors;
  std::vector<ParamInfo> params;
  std::vector<int> param_kind, param_a, param_b;  // deconv sub-pixel phases
  std::vector<Op> ops;
  std::vector<ChainInfo> chains;
  std::vector<XUnitInfo> xunits;
  size_t sync_off = 0, sync_bytes = 0;   // chain control words + tile flags (zeroed by hrnet_plan_bind)
  size_t act_bytes = 0, weight_bytes = 0;
  int t_input = -1, t_heatmaps = -1;
  int Hh = 0, Wh = 0;
  // staging for hrnet_forward_host (inside the workspace)
  size_t off_in_stage = 0, off_joints = 0, off_idx = 0, off_boxes = 0;
  size_t off_pval = 0, off_pidx = 0;      // per-block argmax candidates of the fused head [max_batch][J][head_nblk]
  int head_nblk = 0;
  std::vector<float> head_w, head_b;      // host copy of the head's fp32 weights / bias (kernel parameters), hrnet_plan_bind
  // bound state
  bool bound = false;
  uint8_t* wbase = nullptr;
  uint8_t* abase = nullptr;
  int num_sms = 0;
  cudaStream_t side[3] = {nullptr, nullptr, nullptr};
  cudaStream_t cap = nullptr;  // capture origin: the caller's stream may be the legacy default stream, which cannot capture
  std::vector<cudaEvent_t> events;
  cudaEvent_t fork_ev = nullptr;
  std::map<int, cudaGraphExec_t> graphs;
  int launch_count = 0;
};

namespace {

// ------------------------------------------------------------------------------------------------
// Builder
// ------------------------------------------------------------------------------------------------
struct Builder {
  HrnetPlan& P;
  int maxb;
  bool fuse_f16;
  // arenas: 0 = pre (stem / layer1 / staging), 1 and 2 alternate per stage module
  size_t arena_off[3] = {0, 0, 0};   // bump pointer (bytes at max batch)
  size_t arena_max[3] = {0, 0, 0};
  std::vector<int> tensor_arena;
  std::vector<int> producer;         // tensor id -> op that writes it (-1 external)
  std::vector<int> last_on_stream = std::vector<int>(4, -1);
  size_t wcur = 0;

  explicit Builder(HrnetPlan& p) : P(p), maxb(p.desc.max_batch), fuse_f16((p.desc.flags & HRNET_FLAG_FUSE_F32) == 0) {}

  int new_tensor(int arena, int C, int H, int W, int dtype = DT_F16) {
    TensorInfo t;
    t.C = C; t.H = H; t.W = W; t.dtype = dtype;
    if (arena >= 0) {
      t.offset = arena_off[arena];
      arena_off[arena] += align_up(t.bytes(maxb), 1024);
      arena_max[arena] = std::max(arena_max[arena], arena_off[arena]);
    } else {
      t.offset = (size_t)-1;
    }
    P.tensors.push_back(t);
    tensor_arena.push_back(arena);
    producer.push_back(-1);
    return (int)P.tensors.size() - 1;
  }
  void reset_arena(int a) { arena_off[a] = 0; }

  int new_param(const std::string& conv_key, const std::string& bn_key, int cout, int cin, int kh, int kw,
                bool has_bias = false, bool w_f32 = false, int kind = 0, int sa = 0, int sb = 0) {
    ParamInfo pi;
    pi.conv_key = conv_key; pi.bn_key = bn_key;
    pi.cout = cout; pi.cin = cin; pi.kh = kh; pi.kw = kw;
    pi.has_bias = has_bias; pi.w_f32 = w_f32; pi.transposed = kind;
    wcur = align_up(wcur, 1024);
    pi.w_offset = wcur;
    wcur += (size_t)cout * cin * kh * kw * (w_f32 ? 4 : 2);
    wcur = align_up(wcur, 256);
    pi.scale_offset = wcur; wcur += (size_t)cout * 4;
    wcur = align_up(wcur, 256);
    pi.bias_offset = wcur; wcur += (size_t)cout * 4;
    P.params.push_back(pi);
    P.param_kind.push_back(kind); P.param_a.push_back(sa); P.param_b.push_back(sb);
    return (int)P.params.size() - 1;
  }

  int push(Op op) {
    const int id = (int)P.ops.size();
    auto dep_on = [&](int t) {
      if (t < 0) return;
      int pr = producer[t];
      if (pr >= 0 && std::find(op.deps.begin(), op.deps.end(), pr) == op.deps.end()) op.deps.push_back(pr);
    };
    dep_on(op.in); dep_on(op.res);
    for (int j = 0; j < op.nsrc; ++j) dep_on(op.src[j]);
    if (P.desc.flags & HRNET_FLAG_SERIAL) op.stream = 0;
    if (op.out >= 0) producer[op.out] = id;
    last_on_stream[op.stream] = id;
    P.ops.push_back(op);
    return id;
  }

  // conv + BN (+res) (+relu); returns the output tensor
  int conv(const std::string& name, const std::string& conv_key, const std::string& bn_key, int in, int cout, int k,
           int stride, bool relu, int res, int out_tensor, int arena, int stream, int out_dtype = DT_F16) {
    const TensorInfo& ti = P.tensors[in];
    const int OH = ti.H / stride, OW = ti.W / stride;
    int out = out_tensor >= 0 ? out_tensor : new_tensor(arena, cout, OH, OW, out_dtype);
    Op op;
    op.kind = OP_CONV; op.name = name; op.in = in; op.out = out; op.res = res;
    op.cin = ti.C; op.cout = cout; op.k = k; op.stride = stride; op.pad = k / 2; op.relu = relu ? 1 : 0;
    op.stream = stream;
    op.param = new_param(conv_key, bn_key, cout, ti.C, k, k);
    push(op);
    return out;
  }
};

void choose_tc_cfg(Op& op, uint32_t flags, const int32_t* tune, int max_nt = 256) {
  op.use_tc = false;
  if (flags & HRNET_FLAG_FORCE_SIMT) return;
  if (op.kind != OP_CONV) return;
  if (op.cin % 16 || op.cout % 16) return;
  if (!(op.k == 1 || op.k == 3 || op.k == 2)) return;
  ConvTcCfg c;
  c.kc = 64;   // 128-byte swizzled k-blocks; a channel tail is zero-filled by TMA and skipped by the MMA loop
  int nt = op.cout;
  if (nt > max_nt) {
    int d = 2;
    while (op.cout % d || (op.cout / d) > max_nt || (op.cout / d) % 16) ++d;
    nt = op.cout / d;
  }
  c.n_tile = nt;
  // k-blocks per pipeline stage: every stage costs the issuing threads a barrier wait, an expect_tx and a commit
  // (~300 clk, conv_igemm_body.cuh); with narrow tiles that is more than the stage's MMAs, so two k-blocks share a
  // stage when four such stages still fit (HRNET_TUNE_BPS overrides).
  c.bps = (nt <= 64 && op.k * op.k * ((op.cin + 63) / 64) >= 4) ? 2 : 1;
  if (tune[HRNET_TUNE_BPS] == 1 || tune[HRNET_TUNE_BPS] == 2) c.bps = tune[HRNET_TUNE_BPS];
  // CTA-pair mode (tcgen05 cta_group::2): each CTA stages half of the weight tile.  Measured on B200 it makes the
  // large-N kernels faster in isolation (stage-4 C=192/384 convs 46 -> 42.7 us, 2.3x fewer clk per GEMM row) but the
  // whole W48/64 forward slower (9.76 vs 9.46 ms: pair clusters co-schedule worse with the other branches' kernels,
  // profiles/r01_exp_pair_mode_scope.log), so it is opt-in: HRNET_TUNE_IGEMM_PAIR = 2 (optionally _PAIR_MIN_K).
  c.cs = tune[HRNET_TUNE_IGEMM_PAIR] == 2 ? 2 : 1;
  if (op.k * op.k * op.cin < tune[HRNET_TUNE_IGEMM_PAIR_MIN_K]) c.cs = 1;   // pair mode only for convs with at least this K
  if (nt % 16 || (nt / 2) % 8) c.cs = 1;
  // two MMA-issuing warps on alternate tiles (conv_igemm_body.cuh): HRNET_TUNE_IGEMM_MMA2 = 1 (all convs) or = <max N>
  // (only tiles at most that wide, where the issue side rather than shared-memory bandwidth sets the pace)
  const int a_blk = (int)align_up((size_t)128 * c.kc * 2, 1024);
  const int b_blk = (int)align_up((size_t)(nt / c.cs) * c.kc * 2, 1024);
  const int stage = c.bps * (a_blk + b_blk);
  const int fixed = 1024 + 2 * op.cout * 4 + 1024;   // alignment slack, BN scale / bias, barriers + k-block table
  const int budget = 208 * 1024;
  const int nkb = op.k * op.k * ((op.cin + c.kc - 1) / c.kc);
  const int kstages = (nkb + c.bps - 1) / c.bps;
  c.stages = std::max(2, std::min({8, (budget - fixed) / stage, std::max(2, kstages * 2)}));
  c.stages &= ~1;   // two producer warps alternate stages
  // two MMA-issuing warps on alternate tiles, each with half of the stage ring (conv_igemm_body.cuh): on for tiles at
  // most 96 channels wide, where per-tile barrier / commit latencies rather than shared-memory bandwidth set the pace
  // (forward 7.81 -> 7.71 ms); wider tiles lose more from the halved ring (C = 384 branch conv 35 -> 52 us).
  // HRNET_TUNE_IGEMM_MMA2 = -1 (off) / 1 (all convs) / <max N> overrides.
  int mma2_max_n = 96;
  if (tune[HRNET_TUNE_IGEMM_MMA2]) { const int v = tune[HRNET_TUNE_IGEMM_MMA2]; mma2_max_n = v == 1 ? 1 << 30 : v; }
  c.mma_warps = nt <= mma2_max_n ? 2 : 1;
  if (c.cs == 2 || c.stages < 4) c.mma_warps = 1;
  c.smem_bytes = fixed + c.stages * stage;
  int cols = 32;
  while (cols < 2 * nt) cols *= 2;
  c.tmem_cols = cols;
  op.tc = c;
  op.use_tc = true;
}

// Halo-patch path: 3x3 stride-1 convs whose 9-tap weights fit in shared memory next to >= 2 patch slots and
// whose map tiles into 8x16 output tiles with little waste.
constexpr int kMaxDynSmem = 226 * 1024;
constexpr int kChainMaxSmem = 227 * 1024;   // the most a block may opt in to on sm_100 (228 KB per SM, 1 KB reserved per block)
bool choose_patch_cfg(Op& op, int H, int W, uint32_t flags, const int32_t* tune) {
  op.use_patch = false;
  if (flags & (HRNET_FLAG_FORCE_SIMT | HRNET_FLAG_NO_PATCH)) return false;
  if (op.kind != OP_CONV || op.k != 3 || op.stride != 1 || op.pad != 1) return false;
  if (op.cin % 16 || op.cout % 16 || op.cout > 256 || op.cin > 256) return false;
  const int tw = (W + kPatchTW - 1) / kPatchTW, th = (H + kPatchTH - 1) / kPatchTH;
  if ((double)(H * W) / (double)(tw * kPatchTW * th * kPatchTH) < 0.85) return false;
  ConvPatchParams p{};
  p.H = H; p.W = W; p.Cin = op.cin; p.Cout = op.cout; p.tiles_w = tw; p.tiles_h = th;
  // CTA-pair mode (conv3x3_patch_body.cuh): HRNET_TUNE_PATCH_PAIR_MIN_COUT enables it for convs with at least that
  // many (and at most HRNET_TUNE_PATCH_PAIR_MAX_COUT) output channels; each CTA then keeps Cout / 2 weight rows.
  p.cs = 1;
  if (!(flags & HRNET_FLAG_GROUP)) {
    const int pair_min = tune[HRNET_TUNE_PATCH_PAIR_MIN_COUT];
    const int pair_max = tune[HRNET_TUNE_PATCH_PAIR_MAX_COUT] > 0 ? tune[HRNET_TUNE_PATCH_PAIR_MAX_COUT] : 1 << 30;
    if (pair_min > 0 && op.cout >= pair_min && op.cout <= pair_max && op.cout % 16 == 0) p.cs = 2;
  }
  const int b_rows = op.cout / p.cs;
  int c = 0, n = 0;
  size_t boff = 0;
  while (c < op.cin) {
    const int real = std::min(64, op.cin - c);
    if (n == 4) return false;
    const int bkc = real > 32 ? 64 : (real > 16 ? 32 : 16);
    p.c0[n] = c; p.kreal[n] = real; p.bkc[n] = bkc; p.mapi[n] = bkc == 64 ? 0 : (bkc == 32 ? 1 : 2);
    p.bblk[n] = (int)align_up((size_t)b_rows * bkc * 2, 1024);
    p.boff[n] = (int)boff;
    boff += (size_t)9 * p.bblk[n];
    c += 64; ++n;
  }
  p.nchunks = n;
  p.a_slot_bytes = (int)align_up((size_t)kPatchRows * 128, 1024);
  const int fixed = 1024 + 2 * op.cout * 4 + 512;
  p.b_stream = 0;
  p.b_bytes = (int)boff;
  p.slot_bytes = p.a_slot_bytes;
  int avail = kMaxDynSmem - fixed - p.b_bytes;
  if (avail < 2 * p.slot_bytes && p.cs == 2) return false;   // (not reached for the HRNet shapes)
  if (avail < 2 * p.slot_bytes) {
    // weights too big to stay resident: stream each chunk's 9-tap block with its patch
    int maxblk = 0;
    for (int j = 0; j < n; ++j) maxblk = std::max(maxblk, p.bblk[j]);
    p.b_stream = 1;
    p.b_bytes = 0;
    p.slot_bytes = p.a_slot_bytes + 9 * maxblk;
    avail = kMaxDynSmem - fixed;
    if (avail < 2 * p.slot_bytes) return false;
  }
  p.nslots = std::min(8, avail / p.slot_bytes) & ~1;   // two producer warps alternate slots
  // accumulator buffers: 4 when they fit the 512 TMEM columns (HRNET_TUNE_PATCH_NACC = 2 forces double buffering)
  p.nacc = 4 * op.cout <= 512 ? 4 : 2;
  if (tune[HRNET_TUNE_PATCH_NACC] == 2) p.nacc = 2;
  p.nacc_log2 = p.nacc == 4 ? 2 : 1;
  // two MMA-issuing warps on alternate tiles, each with half of the slot ring (conv3x3_patch_body.cuh): on when every
  // ring keeps two slots (C = 48 branch convs 39.7 -> 33.9 us, layer1 conv2 47.8 -> 44.8; with one slot per ring the
  // C = 96 convs and the streamed-weight transition conv get slower, profiles/r01_exp_mma2_issuers.log).
  // HRNET_TUNE_PATCH_MMA2 = 1 (one issuer) / 2 (two) overrides.
  p.mma_warps = p.nslots >= 4 ? 2 : 1;
  if (tune[HRNET_TUNE_PATCH_MMA2] == 1 || tune[HRNET_TUNE_PATCH_MMA2] == 2) p.mma_warps = tune[HRNET_TUNE_PATCH_MMA2];
  if (p.cs == 2) p.mma_warps = 1;
  int cols = 32;
  while (cols < p.nacc * op.cout) cols *= 2;
  p.tmem_cols = cols;
  p.relu = op.relu;
  op.pp = p;
  op.patch_smem = fixed + p.b_bytes + p.nslots * p.slot_bytes;
  op.use_patch = true;
  return true;
}

// Epilogue selection (epilogue.cuh), fp16 outputs only and not for the sub-pixel phases of a transposed conv:
//   1 staged TMA stores        : tiles at least 256 channels wide (layer1 conv3 / downsample: 136 -> 104 us)
//   2 warp-staged coalesced    : everything else, when the 36 KB of staging fit next to the pipeline
//   0 direct row-per-thread    : fp32 outputs, sub-pixel phases, or no shared memory left (C = 96 halo-patch conv)
// HRNET_TUNE_EPILOGUE = 0 auto | 1 direct | 2 tma (TMA wherever eligible) | 3 coal (coalesced wherever it fits) |
// 4 / 5 TMA on every halo-patch / im2col conv | 6 batch (kind 3: direct stores with batched TMEM loads on tiles <= 64
// channels); read at plan time.  The staging tiles are carved out of the pipeline's shared memory.
int epi_policy(const int32_t* tune) {
  const int v = tune[HRNET_TUNE_EPILOGUE];
  return v >= 0 && v <= 6 ? v : 0;
}
// The epilogue staging tiles are carved out of the pipeline's shared memory: re-check that every issuer's ring keeps its
// minimum depth once the final slot / stage counts are known.
void finalize_mma_warps(Op& op, const int32_t* tune) {
  if (!op.use_tc) return;
  if (op.use_patch) {
    if (op.pp.mma_warps == 2 && op.pp.nslots < 4 && !tune[HRNET_TUNE_PATCH_MMA2]) op.pp.mma_warps = 1;
  } else if (op.tc.mma_warps == 2 && op.tc.stages < 4) {
    op.tc.mma_warps = 1;
  }
}

void choose_epi_impl(Op& op, bool out_f32, bool sub, bool has_res, const int32_t* tune);
void choose_epi(Op& op, bool out_f32, bool sub, bool has_res, const int32_t* tune) {
  choose_epi_impl(op, out_f32, sub, has_res, tune);
  finalize_mma_warps(op, tune);
}
void choose_epi_impl(Op& op, bool out_f32, bool sub, bool has_res, const int32_t* tune) {
  const int policy = epi_policy(tune);
  if (!op.use_tc || policy == 1 || out_f32 || sub) return;
  const int width = op.use_patch ? op.pp.Cout : op.tc.n_tile;
  if (policy == 6 && width <= 64 && (op.use_patch ? op.pp.cs : op.tc.cs) == 1) {   // no staging memory needed
    if (op.use_patch) op.pp.epi_tma = 3; else op.tc.epi = 3;
    return;
  }
  // CTA-pair halo-patch convs need the cheap per-warp TMA epilogue to keep up with the halved MMA time
  const bool want_tma = policy == 2 || (policy == 4 && op.use_patch) || (policy == 5 && !op.use_patch) || width >= 256 ||
                        (op.use_patch && op.pp.cs == 2);
  for (int kind = want_tma ? 1 : 2; kind <= 2; ++kind) {   // TMA first if wanted, else / then coalesced
    if (kind == 2 && policy != 3) break;   // the coalesced epilogue is opt-in only
    // kind 1: im2col kernel: per warpgroup an output (+ residual) tile of 128 x 64 fp16; halo-patch kernel: per warp
    //         an output + a residual tile of 32 x 64 fp16;  kind 2: per epilogue warp 32 rows x 144 B
    const int reserve = kind == 1 ? (op.use_patch ? 8 * 8192 : (has_res ? 65536 : 32768)) : 8 * kCoalWarpBytes;
    if (op.use_patch) {
      ConvPatchParams& p = op.pp;
      const int fixed = 1024 + 2 * op.cout * 4 + 512;
      const int avail = kMaxDynSmem - fixed - p.b_bytes - reserve;
      const int ns = std::min(8, avail / p.slot_bytes) & ~1;
      if (ns < 2 || (ns < 4 && p.nchunks > 1 && !p.b_stream)) continue;   // keep two tiles in flight
      p.nslots = std::min(p.nslots, ns);
      p.epi_tma = kind; p.epi_bytes = reserve;
      op.patch_smem = fixed + p.b_bytes + p.nslots * p.slot_bytes + reserve;
      return;
    } else {
      ConvTcCfg& c = op.tc;
      if (kind == 1 && op.cout / c.n_tile > 1 && c.n_tile % 64) continue;   // a partial last 64-channel chunk would spill into the next N-tile
      const int fixed = 1024 + 2 * op.cout * 4 + 1024;   // alignment slack, BN scale / bias, barriers + k-block table
      const int stage = (c.smem_bytes - fixed) / c.stages;
      const int ns = std::min(c.stages, ((kMaxDynSmem - fixed - reserve) / stage) & ~1);
      if (ns < std::min(c.stages, 4)) continue;
      c.stages = ns; c.epi = kind; c.epi_bytes = reserve;
      c.smem_bytes = fixed + ns * stage + reserve;
      return;
    }
  }
}

// ---- HRNet ------------------------------------------------------------------------------------
int build_hrnet(HrnetPlan& P) {
  const HrnetDesc& d = P.desc;
  const int c = d.c, J = d.nof_joints, H = d.height, W = d.width;
  Builder b(P);
  const int dt_term = b.fuse_f16 ? DT_F16 : DT_F32;

  P.t_input = b.new_tensor(-1, 3, H, W, DT_F32);
  // stem (hrnet.py:158-163)
  int s1 = b.new_tensor(0, 64, H / 2, W / 2);
  {
    Op op; op.kind = OP_STEM; op.name = "conv1"; op.in = P.t_input; op.out = s1; op.cin = 3; op.cout = 64; op.k = 3;
    op.stride = 2; op.pad = 1; op.relu = 1; op.stream = 0;
    op.param = b.new_param("conv1", "bn1", 64, 3, 3, 3, false, true);
    b.push(op);
  }
  int x = b.conv("conv2", "conv2", "bn2", s1, 64, 3, 2, true, -1, -1, 0, 0);
  // layer1: 4 x Bottleneck (hrnet.py:86-95, modules.py:20-40)
  {
    const int h4 = H / 4, w4 = W / 4;
    int ta = b.new_tensor(0, 64, h4, w4), tb = b.new_tensor(0, 64, h4, w4);
    int ty[2] = {b.new_tensor(0, 256, h4, w4), b.new_tensor(0, 256, h4, w4)};
    int tr = b.new_tensor(0, 256, h4, w4);
    for (int k = 0; k < 4; ++k) {
      const std::string p = "layer1." + std::to_string(k);
      b.conv(p + ".conv1", p + ".conv1", p + ".bn1", x, 64, 1, 1, true, -1, ta, 0, 0);
      b.conv(p + ".conv2", p + ".conv2", p + ".bn2", ta, 64, 3, 1, true, -1, tb, 0, 0);
      int res = x;
      if (k == 0) {
        b.conv(p + ".downsample", p + ".downsample.0", p + ".downsample.1", x, 256, 1, 1, false, -1, tr, 0, 0);
        res = tr;
      }
      b.conv(p + ".conv3", p + ".conv3", p + ".bn3", tb, 256, 1, 1, true, res, ty[k & 1], 0, 0);
      x = ty[k & 1];
    }
  }
  // transition1 (hrnet.py:98-109) -> arena 2 (the "previous" arena of module 0, which uses arena 1)
  std::vector<int> xs;
  xs.push_back(b.conv("transition1.0", "transition1.0.0", "transition1.0.1", x, c, 3, 1, true, -1, -1, 2, 0));
  xs.push_back(b.conv("transition1.1", "transition1.1.0.0", "transition1.1.0.1", x, 2 * c, 3, 2, true, -1, -1, 2, 1));

  int module_idx = 0;
  int grp_counter = 0;
  auto stage_module = [&](const std::string& prefix, int S, int O) {
    const int arena = 1 + (module_idx & 1);
    b.reset_arena(arena);
    ++module_idx;
    std::vector<int> ys(S);
    {
      // branch convs are emitted level by level (k-th conv of every branch next to each other) so that one level can
      // be issued as ONE grouped launch (conv_group.cu); each branch keeps its private t / y0 / y1 rotation.
      std::vector<int> tbuf(S), cur(S);
      std::vector<std::array<int, 2>> ybuf(S);
      for (int i = 0; i < S; ++i) {
        const TensorInfo ti = P.tensors[xs[i]];
        tbuf[i] = b.new_tensor(arena, ti.C, ti.H, ti.W);
        ybuf[i] = {b.new_tensor(arena, ti.C, ti.H, ti.W), b.new_tensor(arena, ti.C, ti.H, ti.W)};
        cur[i] = xs[i];
      }
      const int chain_base = (int)P.chains.size();
      for (int i = 0; i < S; ++i) {
        ChainInfo ci; ci.module = module_idx; ci.branch = i;
        P.chains.push_back(ci);
      }
      for (int k = 0; k < 4; ++k) {
        for (int half = 0; half < 2; ++half) {
          const int grp = grp_counter++;
          for (int i = 0; i < S; ++i) {
            const int C = P.tensors[xs[i]].C;
            const std::string p = prefix + ".branches." + std::to_string(i) + "." + std::to_string(k);
            if (half == 0)
              b.conv(p + ".conv1", p + ".conv1", p + ".bn1", cur[i], C, 3, 1, true, -1, tbuf[i], arena, i);
            else
              b.conv(p + ".conv2", p + ".conv2", p + ".bn2", tbuf[i], C, 3, 1, true, cur[i], ybuf[i][k & 1], arena, i);
            P.ops.back().group = module_idx;
            P.ops.back().grp = grp;
            P.ops.back().chain = chain_base + i;
            P.ops.back().chain_pos = 2 * k + half;
            P.chains[chain_base + i].ops.push_back((int)P.ops.size() - 1);
          }
        }
        for (int i = 0; i < S; ++i) cur[i] = ybuf[i][k & 1];
      }
      for (int i = 0; i < S; ++i) ys[i] = cur[i];
    }
    std::vector<int> outs(O);
    // Exchange units.  Default / HRNET_TUNE_XUNIT = 1: ONE UNIT PER SOURCE BRANCH j -- the 1x1 up convs and the 3x3 s2 down
    // chains that read branch j's output -- launched on branch j's stream right behind its chain: it starts when THAT chain
    // is done and runs under the tails of the others (a unit of all 16 convs had to wait for the slowest chain, which ate
    // what it saved: profiles/r02_s7_*.log, r02_s21_*.log).  HRNET_TUNE_XUNIT = 2: one unit per module that also executes
    // the sums (measured slower, kept as an experiment).
    const bool per_source = P.desc.tune[HRNET_TUNE_XUNIT] != 2;
    const int xid = (int)P.xunits.size();
    for (int j = 0; j < (per_source ? S : 1); ++j) { XUnitInfo xu; xu.module = module_idx; xu.stream = per_source ? j : 0; P.xunits.push_back(xu); }
    auto mark_x = [&](int level, int src_branch) {
      Op& o = P.ops.back();
      o.xunit = xid + (per_source ? src_branch : 0); o.xlevel = level;
      P.xunits[o.xunit].ops.push_back((int)P.ops.size() - 1);
    };
    for (int i = 0; i < O; ++i) {
      const TensorInfo ti = P.tensors[ys[i]];
      Op f; f.kind = OP_FUSE; f.name = prefix + ".fuse." + std::to_string(i); f.relu = 1; f.stream = i; f.nsrc = S;
      f.cout = ti.C;
      for (int j = 0; j < S; ++j) {
        const std::string p = prefix + ".fuse_layers." + std::to_string(i) + "." + std::to_string(j);
        if (j == i) {
          f.src[j] = ys[j]; f.shift[j] = 0;
        } else if (j > i) {
          // 1x1 conv + BN at the low resolution; the nearest upsample is folded into the fuse read (hrnet.py:30-35)
          f.src[j] = b.conv(p, p + ".0", p + ".1", ys[j], ti.C, 1, 1, false, -1, -1, arena, i, dt_term);
          mark_x(0, j);
          f.shift[j] = j - i;
        } else {
          int t = ys[j];
          const int cj = P.tensors[ys[j]].C;
          for (int k = 0; k < i - j - 1; ++k) {
            t = b.conv(p + "." + std::to_string(k), p + "." + std::to_string(k) + ".0", p + "." + std::to_string(k) + ".1",
                       t, cj, 3, 2, true, -1, -1, arena, i);
            mark_x(k, j);
          }
          const int k = i - j - 1;
          f.src[j] = b.conv(p + "." + std::to_string(k), p + "." + std::to_string(k) + ".0",
                            p + "." + std::to_string(k) + ".1", t, ti.C, 3, 2, false, -1, -1, arena, i, dt_term);
          mark_x(k, j);
          f.shift[j] = 0;
        }
      }
      f.out = b.new_tensor(arena, ti.C, ti.H, ti.W);
      outs[i] = f.out;
      b.push(f);
      if (!per_source) {
        P.ops.back().xunit = xid;
        P.xunits[xid].sums.push_back((int)P.ops.size() - 1);
      }
    }
    xs = outs;
    return arena;
  };

  int ar = stage_module("stage2.0", 2, 2);
  xs.push_back(b.conv("transition2.2", "transition2.2.0.0", "transition2.2.0.1", xs.back(), 4 * c, 3, 2, true, -1, -1,
                      ar, 2));
  for (int m = 0; m < 4; ++m) ar = stage_module("stage3." + std::to_string(m), 3, 3);
  xs.push_back(b.conv("transition3.3", "transition3.3.0.0", "transition3.3.0.1", xs.back(), 8 * c, 3, 2, true, -1, -1,
                      ar, 3));
  stage_module("stage4.0", 4, 4);
  stage_module("stage4.1", 4, 4);
  stage_module("stage4.2", 4, 1);

  // head (hrnet.py:155,187) + decode (SimpleHRNet.py:296-308)
  P.Hh = H / 4; P.Wh = W / 4;
  P.t_heatmaps = b.new_tensor(0, J, P.Hh, P.Wh, DT_F32);
  {
    Op op; op.kind = OP_HEAD; op.name = "final_layer"; op.in = xs[0]; op.out = P.t_heatmaps; op.cin = c; op.cout = J;
    op.k = 1; op.stream = 0;
    op.param = b.new_param("final_layer", "", J, c, 1, 1, true, true);
    b.push(op);
  }
  {
    Op op; op.kind = OP_ARGMAX; op.name = "argmax_decode"; op.in = P.t_heatmaps; op.cout = J; op.stream = 0;
    b.push(op);
  }
  // host staging (forward_host): input, joints, idx, boxes
  const size_t pre = b.arena_max[0];
  size_t cur = pre;
  P.off_in_stage = cur; cur += align_up((size_t)b.maxb * 3 * H * W * 4, 1024);
  P.off_joints = cur;   cur += align_up((size_t)b.maxb * J * 3 * 4, 1024);
  P.off_idx = cur;      cur += align_up((size_t)b.maxb * J * 4, 1024);
  P.off_boxes = cur;    cur += align_up((size_t)b.maxb * 4 * 4, 1024);
  P.head_nblk = head_c_blocks(P.Hh * P.Wh);
  P.off_pval = cur;     cur += align_up((size_t)b.maxb * J * P.head_nblk * 4, 1024);
  P.off_pidx = cur;     cur += align_up((size_t)b.maxb * J * P.head_nblk * 4, 1024);
  const size_t a0 = cur, a1 = b.arena_max[1], a2 = b.arena_max[2];
  for (size_t i = 0; i < P.tensors.size(); ++i) {
    const int a = b.tensor_arena[i];
    if (a == 1) P.tensors[i].offset += a0;
    if (a == 2) P.tensors[i].offset += a0 + a1;
  }
  P.act_bytes = a0 + a1 + a2;
  P.weight_bytes = align_up(b.wcur, 1024);
  return 0;
}

// ---- PoseResNet (models_/poseresnet.py) -------------------------------------------------------
int build_poseresnet(HrnetPlan& P) {
  const HrnetDesc& d = P.desc;
  const int J = d.nof_joints, H = d.height, W = d.width;
  int layers[4];
  if (d.c == 50) { int l[4] = {3, 4, 6, 3}; memcpy(layers, l, sizeof l); }
  else if (d.c == 101) { int l[4] = {3, 4, 23, 3}; memcpy(layers, l, sizeof l); }
  else if (d.c == 152) { int l[4] = {3, 8, 36, 3}; memcpy(layers, l, sizeof l); }
  else return fail(HRNET_E_INVALID, "PoseResNet size must be 50, 101 or 152 (18/34 are broken in the reference, modules.py:51)");
  Builder b(P);
  P.t_input = b.new_tensor(-1, 3, H, W, DT_F32);
  int s1 = b.new_tensor(0, 64, H / 2, W / 2);
  {
    Op op; op.kind = OP_STEM7; op.name = "conv1"; op.in = P.t_input; op.out = s1; op.cin = 3; op.cout = 64; op.k = 7;
    op.stride = 2; op.pad = 3; op.relu = 1;
    op.param = b.new_param("conv1", "bn1", 64, 3, 7, 7, false, true);
    b.push(op);
  }
  int x = b.new_tensor(0, 64, H / 4, W / 4);
  {
    Op op; op.kind = OP_MAXPOOL; op.name = "maxpool"; op.in = s1; op.out = x; op.cin = 64; op.cout = 64; op.k = 3;
    op.stride = 2; op.pad = 1;
    b.push(op);
  }
  int inplanes = 64;
  const int planes_l[4] = {64, 128, 256, 512};
  for (int li = 0; li < 4; ++li) {
    const int planes = planes_l[li];
    const int stride = li == 0 ? 1 : 2;
    for (int k = 0; k < layers[li]; ++k) {
      const std::string p = "layer" + std::to_string(li + 1) + "." + std::to_string(k);
      const int s = k == 0 ? stride : 1;
      const bool down = k == 0 && (stride != 1 || inplanes != planes * 4);
      int a = b.conv(p + ".conv1", p + ".conv1", p + ".bn1", x, planes, 1, 1, true, -1, -1, 0, 0);
      int bb = b.conv(p + ".conv2", p + ".conv2", p + ".bn2", a, planes, 3, s, true, -1, -1, 0, 0);
      int res = x;
      if (down) res = b.conv(p + ".downsample", p + ".downsample.0", p + ".downsample.1", x, planes * 4, 1, s, false, -1, -1, 0, 0);
      x = b.conv(p + ".conv3", p + ".conv3", p + ".bn3", bb, planes * 4, 1, 1, true, res, -1, 0, 0);
      inplanes = planes * 4;
    }
  }
  // 3 x ConvTranspose2d(4, s2, p1) + BN + ReLU as four 2x2 sub-pixel convs each (poseresnet.py:81-106)
  for (int dl = 0; dl < 3; ++dl) {
    const TensorInfo ti = P.tensors[x];
    int out = b.new_tensor(0, 256, ti.H * 2, ti.W * 2);
    for (int a = 0; a < 2; ++a)
      for (int bb = 0; bb < 2; ++bb) {
        Op op; op.kind = OP_CONV; op.name = "deconv" + std::to_string(dl) + "." + std::to_string(a) + std::to_string(bb);
        op.in = x; op.out = out; op.cin = ti.C; op.cout = 256; op.k = 2; op.stride = 1;
        op.pad = 100 + a * 2 + bb;  // marks a sub-pixel phase: pad_lo = 1 - a (rows), 1 - b (cols)
        op.relu = 1; op.stream = 0;
        op.param = b.new_param("deconv_layers." + std::to_string(3 * dl), "deconv_layers." + std::to_string(3 * dl + 1),
                               256, ti.C, 2, 2, false, false, 1, a, bb);
        b.push(op);
      }
    x = out;
  }
  P.Hh = H / 4; P.Wh = W / 4;
  P.t_heatmaps = b.new_tensor(0, J, P.Hh, P.Wh, DT_F32);
  {
    Op op; op.kind = OP_HEAD; op.name = "final_layer"; op.in = x; op.out = P.t_heatmaps; op.cin = 256; op.cout = J; op.k = 1;
    op.param = b.new_param("final_layer", "", J, 256, 1, 1, true, true);
    b.push(op);
  }
  {
    Op op; op.kind = OP_ARGMAX; op.name = "argmax_decode"; op.in = P.t_heatmaps; op.cout = J;
    b.push(op);
  }
  size_t cur = b.arena_max[0];
  P.off_in_stage = cur; cur += align_up((size_t)b.maxb * 3 * H * W * 4, 1024);
  P.off_joints = cur;   cur += align_up((size_t)b.maxb * J * 3 * 4, 1024);
  P.off_idx = cur;      cur += align_up((size_t)b.maxb * J * 4, 1024);
  P.off_boxes = cur;    cur += align_up((size_t)b.maxb * 4 * 4, 1024);
  P.head_nblk = head_c_blocks(P.Hh * P.Wh);
  P.off_pval = cur;     cur += align_up((size_t)b.maxb * J * P.head_nblk * 4, 1024);
  P.off_pidx = cur;     cur += align_up((size_t)b.maxb * J * P.head_nblk * 4, 1024);
  P.act_bytes = cur;
  P.weight_bytes = align_up(b.wcur, 1024);
  return 0;
}

// Branch chains: which chains run as one kernel, their shared-memory needs, the tile-flag region in the workspace, the
// "all chains of a module start together" dependencies and each chain's share of the SMs.
void plan_chains(HrnetPlan& P) {
  const uint32_t off_flags = HRNET_FLAG_NO_CHAIN | HRNET_FLAG_FORCE_SIMT | HRNET_FLAG_GROUP | HRNET_FLAG_PARTITION;
  const bool want = !(P.desc.flags & off_flags);
  size_t cur = (P.act_bytes + 1023) / 1024 * 1024;
  P.sync_off = cur;
  for (auto& ch : P.chains) {
    ch.enabled = false;
    if (!want || ch.ops.empty() || (int)ch.ops.size() > kChainMaxConv) continue;
    const Op& o0 = P.ops[ch.ops[0]];
    bool ok = o0.use_tc && o0.k == 3 && o0.stride == 1 && o0.pad == 1 && o0.cin == o0.cout;
    for (int i : ch.ops) {
      const Op& o = P.ops[i];
      ok = ok && o.use_tc && o.use_patch == o0.use_patch && o.cin == o0.cin && o.cout == o0.cout && o.k == 3 &&
           o.stride == 1 && P.tensors[o.in].H == P.tensors[o0.in].H && P.tensors[o.in].W == P.tensors[o0.in].W &&
           P.tensors[o.out].dtype == DT_F16;
      if (o.use_patch) ok = ok && o.pp.cs == 1 && !o.pp.b_stream && o.pp.epi_tma == 0 && o.pp.nacc == 4;
      else ok = ok && o.tc.cs == 1 && o.tc.epi == 0 && o.tc.n_tile <= 256 && o.cout / o.tc.n_tile <= 15;
    }
    ok = ok && o0.cout <= (o0.use_patch ? kChainMaxCoutP : kChainMaxCoutI);
    if (!ok) continue;
    ch.enabled = true;
    ch.patch = o0.use_patch;
    const TensorInfo& ti = P.tensors[o0.in];
    const double k16 = 9.0 * ((o0.cin + 15) / 16);
    if (ch.patch) {
      ch.smem = o0.patch_smem + (int)(ch.ops.size() - 1) * 2 * o0.cout * 4;   // (room the BN constants used to need)
      if (ch.smem > kMaxDynSmem) { ch.enabled = false; continue; }
      // 64 + 32 input channels (C = 96), one issuer, two slots: the second chunk only needs 64-byte rows -- a third slot
      // fits (2 x 23 KB + 12 KB) and the next tile's chunk 0 loads while this tile's is multiplied (the reload of the only
      // chunk-0 slot cost ~700 clk per 4,380-clk tile).  HRNET_TUNE_CHAIN_SKIP = 8 keeps the two-slot layout (experiment).
      ch.tail64 = o0.pp.nchunks == 2 && o0.pp.kreal[0] == 64 && o0.pp.kreal[1] == 32 && o0.pp.mma_warps == 1 && o0.pp.nslots == 2 &&
                  P.desc.tune[HRNET_TUNE_CHAIN_SKIP] != 8;
      if (ch.tail64) {
        const int smem3 = o0.patch_smem + 12288;      // (kPatchRows x 64 B = 11,520 rounded up: the third, narrow slot)
        if (smem3 <= kMaxDynSmem) ch.smem = std::max(ch.smem, smem3); else ch.tail64 = false;
      }
      ch.flag_stride = P.desc.max_batch * o0.pp.tiles_h;             // one counter per row of tiles of an image
      // per-MMA costs calibrated on the in-situ module times of SM-split sweeps (profiles/r02_s26_split_sweep.log: per
      // mille 375 / 235 / 195 / 195 for C = 48 / 96 / 192 / 384 (after the three-slot layout of the C = 96 chains, r02_s35) is 6 % faster than the isolated per-tile times suggest --
      // 1,973 clk per 128 x 48 x 432 tile, 4,380 per 128 x 96 x 864 tile -- because the chains also share L2 and power)
      ch.cost = (double)ch.ops.size() * P.desc.max_batch * o0.pp.tiles_w * o0.pp.tiles_h * k16 * (o0.cout <= 64 ? 67.3 : 75.8);
    } else {
      for (int i : ch.ops) P.ops[i].tc.mma_warps = 1;
      const int n_tiles = o0.cout / o0.tc.n_tile;
      const int m_tiles = (P.desc.max_batch * ti.H * ti.W + 127) / 128;
      // two M-tiles per ticket share every weight k-block (conv_chain.cu); their accumulators need 2 x n_tile TMEM columns
      // (opt-in, HRNET_TUNE_CHAIN_M2 = 2: measured equal inside the forward -- 7.228 vs 7.229 ms -- and slower alone on
      // the GPU, where half as many tickets per conv leave CTAs idle: profiles/r02_s5_chain_deferred_publish_m1_vs_m2.log)
      ch.m2 = (P.desc.tune[HRNET_TUNE_CHAIN_M2] == 2 && 2 * o0.tc.n_tile <= 512) ? 2 : 1;
      // CTA pairs (HRNET_TUNE_CHAIN_PAIR = 2): one cta_group::2 MMA spans the two M-tiles of a unit, each CTA stages half
      // of the weight tile -> 56 instead of 80 KB of shared-memory traffic per 128 x 192 x 64 step and SM
      ch.pair = P.desc.tune[HRNET_TUNE_CHAIN_PAIR] == 2 && o0.tc.n_tile % 32 == 0 && 2 * o0.tc.n_tile <= 512;
      if (ch.pair) ch.m2 = 1;
      const int tpu = ch.pair ? 2 : ch.m2;                           // M-tiles per unit
      {
        const int a_blk = (int)align_up((size_t)128 * o0.tc.kc * 2, 1024);
        const int b_blk = (int)align_up((size_t)(o0.tc.n_tile / (ch.pair ? 2 : 1)) * o0.tc.kc * 2, 1024);
        const int stage = o0.tc.bps * (ch.m2 * a_blk + b_blk);
        const int fixed = 1024 + 2048;                               // alignment slack + barriers, ring, k-block table
        // as deep as shared memory allows: under load a stage refill (commit -> empty -> TMA -> full) takes ~2.2 k clk, i.e.
        // 5-7 k-blocks of MMA time (tools/exp/mma_tma_mix.cu: 148 CTAs x 160 KB in flight move 72 B/clk per SM)
        ch.stages = std::max(2, std::min(P.desc.tune[HRNET_TUNE_CHAIN_STAGES] > 0 ? P.desc.tune[HRNET_TUNE_CHAIN_STAGES] : 8 /* kMaxStages of conv_igemm_body.cuh */,
                                         (kChainMaxSmem - fixed) / stage));
        ch.smem = fixed + ch.stages * stage;
      }
      ch.flag_stride = (m_tiles + tpu - 1) / tpu;                    // one counter per ticket (tpu M-tiles x n_tiles arrivals)
      // measured inside the per-conv kernels: ~200 clk per K16 step of a 128 x 192 tile (profiles/r01_exp_gridcap_pair_sweep.log)
      // ... and ~490 clk per 128 x 192 x 64 k-block of the im2col chains (four K16 steps; CTA pairs ~470)
      ch.cost = (double)ch.ops.size() * m_tiles * n_tiles * k16 * ((ch.pair ? 135.0 : 140.0) * o0.tc.n_tile / 192.0);
    }
    ch.ctrl_off = cur; cur += 256;
    ch.flags_off = cur; cur += ((size_t)ch.ops.size() * ch.flag_stride * 4 + 255) / 256 * 256;
  }
  P.sync_bytes = cur - P.sync_off;
  P.act_bytes = cur;
  for (auto& ch : P.chains)
    if (!ch.enabled) for (int i : ch.ops) { P.ops[i].chain = -1; P.ops[i].chain_pos = 0; }
  // The chains of a module share the SMs (grids sum to the SM count), so they must start together: while some of them
  // still waited for the previous module's exchange unit, that unit's ordinary persistent kernels would only find the
  // SMs the early chains left over.  Every chain's first conv therefore depends on everything any of them depends on.
  std::map<int, std::vector<int>> by_module;
  for (size_t c = 0; c < P.chains.size(); ++c) if (P.chains[c].enabled) by_module[P.chains[c].module].push_back((int)c);
  for (auto& kv : by_module) {
    std::vector<int> deps;
    for (int c : kv.second)
      for (int dpi : P.ops[P.chains[c].ops[0]].deps)
        if (std::find(deps.begin(), deps.end(), dpi) == deps.end()) deps.push_back(dpi);
    // (HRNET_TUNE_CHAIN_EARLY = 1, experiment: every chain starts as soon as ITS input is summed)
    if (P.desc.tune[HRNET_TUNE_CHAIN_EARLY] != 1)
      for (int c : kv.second) P.ops[P.chains[c].ops[0]].deps = deps;
    // grid split: proportional to the estimated cost (HRNET_TUNE_CHAIN_SHARE* overrides, per mille of the SM count);
    // resolved to CTA counts at bind time, when the SM count is known
    double tot = 0;
    for (int c : kv.second) tot += P.chains[c].cost;
    double tsum = 0;
    bool tuned = true;
    for (int c : kv.second) { const int t = P.desc.tune[HRNET_TUNE_CHAIN_SHARE0 + std::min(P.chains[c].branch, 3)]; if (t <= 0) tuned = false; tsum += t; }
    for (int c : kv.second) {
      ChainInfo& ch = P.chains[c];
      const double share = tuned ? P.desc.tune[HRNET_TUNE_CHAIN_SHARE0 + std::min(ch.branch, 3)] / tsum : ch.cost / tot;
      ch.share = std::max(1, (int)std::lround(share * 1000.0));
    }
  }
}

// Exchange units: which modules' fuse-layer convs run as one kernel, ticket order, shared-memory needs, control words.
void plan_xunits(HrnetPlan& P) {
  // Opt-in (HRNET_TUNE_XUNIT): measured on B200 the single kernel cuts the exchange convs' serial kernel time 2.6x (1,350 ->
  // 525 us per forward) but not the forward (7.37 vs 7.33 ms, 7.70 vs 7.35 ms in a second session): launched one by one on
  // four streams the small convs already overlap each other and the tails of the unevenly finishing branch chains, while
  // one kernel waits for the slowest chain; sums as tickets (256 epilogue threads per SM) cannot match fuse_sum_kernel's
  // bandwidth (8.80 ms).  profiles/r02_s7_xunit_first_contact.log, r02_s9_xunit_sum_tickets_v2.log
  const uint32_t off_flags = HRNET_FLAG_FORCE_SIMT | HRNET_FLAG_GROUP | HRNET_FLAG_PARTITION | HRNET_FLAG_FUSE_F32;
  const int mode = P.desc.tune[HRNET_TUNE_XUNIT];      // 0 / 1: per-source units (default), 2: module units with sum tickets, 3: off
  const bool want = mode != 3 && !(P.desc.flags & off_flags);
  size_t cur = (P.act_bytes + 255) / 256 * 256;
  const size_t begin = cur;
  for (size_t x = 0; x < P.xunits.size(); ++x) {
    XUnitInfo& xu = P.xunits[x];
    xu.enabled = false;
    bool ok = want && xu.ops.size() >= 2 && (int)xu.ops.size() <= kXMaxOps;
    int max_nt = 0, total_kb = 0, total_cout = 0;
    for (int i : xu.ops) total_cout += P.ops[i].cout;
    ok = ok && total_cout <= kXMaxSb;
    for (int i : xu.ops) {
      const Op& o = P.ops[i];
      ok = ok && o.use_tc && !o.use_patch && o.tc.cs == 1 && P.tensors[o.out].dtype == DT_F16 && o.res < 0 && o.pad < 100 &&
           o.tc.n_tile <= 256 && o.cout / o.tc.n_tile <= 15;
      max_nt = std::max(max_nt, o.tc.n_tile);
      total_kb += o.k * o.k * ((o.cin + 63) / 64);
    }
    if (mode != 2) {      // sums as separate launches
      for (int i : xu.sums) { P.ops[i].xunit = -1; P.ops[i].xpos = 0; }
      xu.sums.clear();
    }
    if (xu.sums.size() > 4) ok = false;
    for (int i : xu.sums)
      for (int j = 0; j < P.ops[i].nsrc; ++j) ok = ok && P.tensors[P.ops[i].src[j]].dtype == DT_F16;
    if (!ok || total_kb > 384) continue;
    xu.enabled = true;
    // ticket order = dependency order: all first steps (1x1 up convs, first 3x3 s2 of every down chain), then the second
    // steps, then the third; within a level the plan's order
    std::stable_sort(xu.ops.begin(), xu.ops.end(), [&](int a, int b) { return P.ops[a].xlevel < P.ops[b].xlevel; });
    xu.b_blk = (int)align_up((size_t)max_nt * 64 * 2, 1024);
    const int stage = 16384 + xu.b_blk;
    xu.stages = std::min(4, (kChainMaxSmem - 4096) / stage);
    xu.smem = 1024 + xu.stages * stage + 3072;      // alignment slack, pipeline, barriers + ring + k-block table
    int cols = 32;
    while (cols < 2 * max_nt) cols *= 2;
    xu.tmem_cols = cols;
    xu.ctrl_off = cur; cur += 256;
    xu.counters_off = cur; cur += 256;
    // one kernel on the caller's stream: the member convs lose their per-output streams; the first member (in plan order)
    // stands for the launch and waits for everything any member waits for
    int first = *std::min_element(xu.ops.begin(), xu.ops.end());
    std::vector<int> deps;
    std::vector<int> members = xu.ops;
    members.insert(members.end(), xu.sums.begin(), xu.sums.end());      // the sums run as tickets of the same kernel
    for (int i : members) {
      P.ops[i].stream = xu.stream;
      for (int dpi : P.ops[i].deps)
        if (P.ops[dpi].xunit != (int)x && std::find(deps.begin(), deps.end(), dpi) == deps.end()) deps.push_back(dpi);
    }
    P.ops[first].deps = deps;
    int pos = 1;
    for (int i : members) P.ops[i].xpos = i == first ? 0 : pos++;
  }
  for (size_t x = 0; x < P.xunits.size(); ++x)
    if (!P.xunits[x].enabled) {
      for (int i : P.xunits[x].ops) { P.ops[i].xunit = -1; P.ops[i].xpos = 0; }
      for (int i : P.xunits[x].sums) { P.ops[i].xunit = -1; P.ops[i].xpos = 0; }
    }
  if (cur > begin) {
    if (P.sync_bytes == 0) P.sync_off = begin;
    P.sync_bytes = cur - P.sync_off;
    P.act_bytes = cur;
  }
}

void finalize_schedule(HrnetPlan& P) {
  const bool chain_cfg = !(P.desc.flags & (HRNET_FLAG_NO_CHAIN | HRNET_FLAG_FORCE_SIMT | HRNET_FLAG_GROUP | HRNET_FLAG_PARTITION));
  for (auto& op : P.ops) {
    // chain members keep N tiles <= 192 wide (two accumulators in TMEM, direct epilogue) and the direct epilogue
    const bool cand = chain_cfg && op.chain >= 0;
    choose_tc_cfg(op, P.desc.flags, P.desc.tune, cand ? 192 : 256);
    if (op.use_tc && op.in >= 0) choose_patch_cfg(op, P.tensors[op.in].H, P.tensors[op.in].W, P.desc.flags, P.desc.tune);
    if (op.use_tc && op.out >= 0 && !(P.desc.flags & HRNET_FLAG_GROUP) && !cand)
      choose_epi(op, P.tensors[op.out].dtype == DT_F32, op.pad >= 100, op.res >= 0, P.desc.tune);
  }
  // (Opt-in, HRNET_FLAG_PARTITION; measured SLOWER than letting every kernel use all SMs: 12.85 vs 9.9 ms per
  // W48/64 forward, profiles/r01_exp_variants_partition_pdl.log -- total work is unchanged and the low-resolution
  // branches lose more to extra tile rounds than the overlap of prologues wins back.)
  // Branch-level SM partitioning: the S branch chains of a StageModule run concurrently on S streams; capping each
  // branch's persistent grid to its share of the SMs (proportional to its estimated work) lets the chains progress
  // side by side, so per-kernel prologues / tails / wave quantisation of one branch hide behind the others' MMAs.
  if ((P.desc.flags & HRNET_FLAG_PARTITION) && !(P.desc.flags & HRNET_FLAG_SERIAL)) {
    std::map<int, std::map<int, double>> work;   // group -> stream -> estimated SM-cycles of one conv
    auto est = [&](const Op& op) {
      const TensorInfo& ti = P.tensors[op.in];
      const double px = (double)P.desc.max_batch * ti.H * ti.W;
      const double tiles = std::ceil(px / 128.0);
      const int n = op.use_patch ? op.cout : op.tc.n_tile;
      const double nsplit = op.use_patch ? 1.0 : (double)op.cout / n;
      const double k16 = 9.0 * op.cin / 16.0;
      const double mma = std::max(n / 2.0, 32.0 + n / 4.0) + (op.use_patch ? 6.0 : (128.0 + n) / 4.0);
      return tiles * nsplit * (k16 * mma + 900.0);   // calibrated on profiles/r01_dbg_role_timers_v3_kernels_v2.log
    };
    for (auto& op : P.ops)
      if (op.group >= 0 && op.use_tc) work[op.group][op.stream] = est(op);
    for (auto& op : P.ops) {
      if (op.group < 0 || !op.use_tc) continue;
      double tot = 0;
      for (auto& kv : work[op.group]) tot += kv.second;
      op.sm_frac = (float)(work[op.group][op.stream] / tot);
    }
  }
  // Grouped launches (conv_group.cu): the k-th convs of all branches of a StageModule form one launch when they fit the
  // kernel's slots (<= 2 halo-patch + <= 2 im2col problems, no CTA-pair mode).  Per-tile cost model (clk) for the CTA
  // split, calibrated on profiles/r01_dbg_role_timers_v3_kernels_v2.log / r01_exp_mma_rate.log.
  {
    std::map<int, std::vector<int>> groups;
    for (size_t i = 0; i < P.ops.size(); ++i) {
      Op& op = P.ops[i];
      if (op.kind != OP_CONV || !op.use_tc) { op.grp = -1; continue; }
      const int n = op.use_patch ? op.cout : op.tc.n_tile;
      const double k16 = (double)op.k * op.k * ((op.cin + 15) / 16);
      const double mma = std::max(n / 2.0, 32.0 + n / 4.0);
      // wall clk per tile incl. barrier waits, measured inside a grouped launch (profiles/r01_dbg_group_split_v1.log)
      op.work = op.use_patch ? k16 * mma * 1.45 + 650.0 : k16 * (mma + (128.0 + n) / 4.0) * 1.5 + 1500.0;
      if (op.grp >= 0) groups[op.grp].push_back((int)i);
    }
    // opt-in (HRNET_FLAG_GROUP): measured equal to per-branch launches on four streams, whose kernels the hardware
    // already co-schedules (9.42 vs 9.35 ms per W48/64 forward, profiles/r01_exp_variants_group.log)
    const bool enabled = (P.desc.flags & HRNET_FLAG_GROUP) && !(P.desc.flags & (HRNET_FLAG_FORCE_SIMT | HRNET_FLAG_SERIAL));
    for (auto& kv : groups) {
      int np = 0, ni = 0;
      bool ok = enabled && kv.second.size() >= 2;
      for (int i : kv.second) {
        const Op& op = P.ops[i];
        if (op.use_patch) ++np; else ++ni;
        if (!op.use_patch && op.tc.cs != 1) ok = false;
      }
      for (size_t k = 1; k < kv.second.size(); ++k) if (kv.second[k] != kv.second[k - 1] + 1) ok = false;
      if (np > 2 || ni > 2) ok = false;
      for (int i : kv.second) {
        if (!ok) P.ops[i].grp = -1;
        else P.ops[i].stream = 0;      // the whole level is one kernel on the caller's stream
      }
    }
  }
  plan_chains(P);
  plan_xunits(P);
  // every stream's last op must be joined back into stream 0 before the head runs
  int head = -1;
  for (size_t i = 0; i < P.ops.size(); ++i) if (P.ops[i].kind == OP_HEAD) head = (int)i;
  std::vector<int> last(4, -1);
  for (int i = 0; i < head; ++i) last[P.ops[i].stream] = i;
  for (int s = 1; s < 4; ++s)
    if (last[s] >= 0 && std::find(P.ops[head].deps.begin(), P.ops[head].deps.end(), last[s]) == P.ops[head].deps.end())
      P.ops[head].deps.push_back(last[s]);
  for (auto& op : P.ops)
    for (int dpi : op.deps)
      if (P.ops[dpi].stream != op.stream) P.ops[dpi].needs_event = true;
  P.launch_count = 0;
  for (size_t i = 0; i < P.ops.size(); ++i) {
    if (P.ops[i].chain >= 0) { if (P.ops[i].chain_pos == 0) ++P.launch_count; continue; }   // one kernel per chain
    if (P.ops[i].xunit >= 0) { if (P.ops[i].xpos == 0) ++P.launch_count; continue; }        // one kernel per exchange unit
    if (P.ops[i].grp < 0 || i == 0 || P.ops[i - 1].grp != P.ops[i].grp) ++P.launch_count;
  }
}

// ------------------------------------------------------------------------------------------------
// TMA descriptor encoding (driver entry points resolved at run time: no link-time libcuda dependency,
// so the library loads on a CPU-only box)
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode_tiled = nullptr;
EncodeIm2colFn g_encode_im2col = nullptr;
int g_driver_version = 0;

int load_driver_fns() {
  if (g_encode_tiled && g_encode_im2col) return 0;
  cudaDriverEntryPointQueryResult q;
  void* f = nullptr;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !f)
    return fail(HRNET_E_CUDA, std::string("cuTensorMapEncodeTiled unavailable: ") + cudaGetErrorString(e));
  g_encode_tiled = (EncodeTiledFn)f;
  f = nullptr;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &f, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !f)
    return fail(HRNET_E_CUDA, std::string("cuTensorMapEncodeIm2col unavailable: ") + cudaGetErrorString(e));
  g_encode_im2col = (EncodeIm2colFn)f;
  cudaDriverGetVersion(&g_driver_version);
  return 0;
}

CUtensorMapSwizzle swizzle_for(int kc) {
  return kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (kc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

// weights [Cout][K] fp16, box {kc, n_tile}
int encode_weights(CUtensorMap* tm, const void* w, int cout, int K, int kc, int n_tile) {
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)cout};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {(cuuint32_t)kc, (cuuint32_t)n_tile};
  cuuint32_t es[2] = {1, 1};
  CUresult r = g_encode_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(w), dims, strides, box, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(kc), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(HRNET_E_CUDA, "cuTensorMapEncodeTiled(weights) failed: " + std::to_string((int)r));
  return 0;
}

// activations NHWC fp16 [N, IH, IW, C] as an im2col map: 128 output pixels x kc channels per load.
// pad_lo / pad_hi per spatial dim allow the asymmetric 2x2 sub-pixel phases of the transposed conv.
int encode_im2col(CUtensorMap* tm, const void* act, int N, int IH, int IW, int C, int kc, int ksize, int stride,
                  int pad_lo_h, int pad_hi_h, int pad_lo_w, int pad_hi_w) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)IW, (cuuint64_t)IH, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)IW * C * 2, (cuuint64_t)IH * IW * C * 2};
  int lower[2] = {-pad_lo_w, -pad_lo_h};
  int upper[2] = {pad_hi_w - (ksize - 1), pad_hi_h - (ksize - 1)};
  cuuint32_t es[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = g_encode_im2col(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(act), dims, strides, lower,
                               upper, (cuuint32_t)kc, 128, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(kc),
                               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(HRNET_E_CUDA, "cuTensorMapEncodeIm2col failed: " + std::to_string((int)r));
  // Known driver issue (<= 13.1) for im2col maps over tensors smaller than 128 KiB: clear bit 21 of
  // the second descriptor word (same workaround NVIDIA's own CuTe im2col descriptor builder applies).
  if (g_driver_version <= 13010 && (size_t)N * IH * IW * C * 2 < 131072)
    reinterpret_cast<uint64_t*>(tm)[1] &= ~(1ull << 21);
  return 0;
}

// activations NHWC fp16 as a tiled 4-D map whose box is one 10 x 18 halo patch of kc channels
int encode_patch(CUtensorMap* tm, const void* act, int N, int H, int W, int C, int kc) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)kc, (cuuint32_t)kPatchPW, (cuuint32_t)kPatchPH, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = g_encode_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(act), dims, strides, box, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(kc), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(HRNET_E_CUDA, "cuTensorMapEncodeTiled(patch) failed: " + std::to_string((int)r));
  return 0;
}

// output / residual NHWC fp16 as seen by the staged epilogue: 2-D {C, pixels}, box 64 channels x 128 pixels
int encode_out2d(CUtensorMap* tm, const void* ptr, size_t rows, int C) {
  cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)C * 2};
  cuuint32_t box[2] = {64, 128};
  cuuint32_t es[2] = {1, 1};
  CUresult r = g_encode_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(HRNET_E_CUDA, "cuTensorMapEncodeTiled(out 2d) failed: " + std::to_string((int)r));
  return 0;
}
// ... and for the halo-patch kernel: 4-D {C, W, H, N}, box = the 8 x 4 pixels (32 accumulator rows) of one epilogue warp
int encode_out4d(CUtensorMap* tm, const void* ptr, int N, int H, int W, int C) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)kPatchTW, (cuuint32_t)(kPatchTH / 4), 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = g_encode_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(HRNET_E_CUDA, "cuTensorMapEncodeTiled(out 4d) failed: " + std::to_string((int)r));
  return 0;
}
// maps of the staged epilogue of `op` (residual == nullptr: the output map stands in, it is never used)
int encode_epi_maps(Op& op, const void* out, const void* residual, int N, int OH, int OW) {
  const void* res = residual ? residual : out;
  int rc;
  if (op.use_patch) {
    if (op.pp.epi_tma != 1) return 0;
    rc = encode_out4d(&op.tmOR[0], out, N, OH, OW, op.cout);
    if (!rc) rc = encode_out4d(&op.tmOR[1], res, N, OH, OW, op.cout);
  } else {
    if (op.tc.epi != 1) return 0;
    rc = encode_out2d(&op.tmOR[0], out, (size_t)N * OH * OW, op.cout);
    if (!rc) rc = encode_out2d(&op.tmOR[1], res, (size_t)N * OH * OW, op.cout);
  }
  return rc;
}

// all tensor maps of a patch op: one 64-channel activation map and one weight map per block width in use
int encode_patch_maps(Op& op, const void* act, const void* w, int N) {
  const ConvPatchParams& p = op.pp;
  int rc = encode_patch(&op.tmPA[0], act, N, p.H, p.W, p.Cin, 64);
  if (rc) return rc;
  op.tmPA[1] = op.tmPA[2] = op.tmPA[0];
  bool have[3] = {false, false, false};
  for (int j = 0; j < p.nchunks; ++j) have[p.mapi[j]] = true;
  const int kcs[3] = {64, 32, 16};
  int first = -1;
  for (int i = 0; i < 3; ++i) {
    if (!have[i]) continue;
    rc = encode_weights(&op.tmPB[i], w, p.Cout, 9 * p.Cin, kcs[i], p.Cout / p.cs);
    if (rc) return rc;
    if (first < 0) first = i;
  }
  for (int i = 0; i < 3; ++i)
    if (!have[i]) op.tmPB[i] = op.tmPB[first];
  return 0;
}

void conv_geometry(const Op& op, int& pad_lo_h, int& pad_hi_h, int& pad_lo_w, int& pad_hi_w, int& sub, int& sa, int& sb) {
  sub = 0; sa = sb = 0;
  if (op.pad >= 100) {  // transposed-conv sub-pixel phase
    sub = 1; sa = (op.pad - 100) / 2; sb = (op.pad - 100) % 2;
    pad_lo_h = 1 - sa; pad_hi_h = sa; pad_lo_w = 1 - sb; pad_hi_w = sb;
  } else {
    pad_lo_h = pad_hi_h = pad_lo_w = pad_hi_w = op.pad;
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

const char* hrnet_last_error(void) { return g_last_error.c_str(); }

void hrnet_debug_set_tune(const int32_t* tune) {
  for (int i = 0; i < HRNET_TUNE_COUNT; ++i) g_single_op_tune[i] = tune ? tune[i] : 0;
}

int hrnet_plan_create(const HrnetDesc* desc, HrnetPlan** out) {
  if (!desc || !out) return fail(HRNET_E_INVALID, "null argument");
  if (desc->height <= 0 || desc->width <= 0 || desc->height % 32 || desc->width % 32)
    return fail(HRNET_E_INVALID, "resolution must be a positive multiple of 32 (exact x8 down/up-sampling)");
  if (desc->max_batch <= 0) return fail(HRNET_E_INVALID, "max_batch must be positive");
  if (desc->nof_joints <= 0 || desc->nof_joints > 32) return fail(HRNET_E_INVALID, "nof_joints must be in [1, 32]");
  HrnetPlan* P = new HrnetPlan();
  P->desc = *desc;
  int rc;
  if (desc->arch == HRNET_ARCH_HRNET) {
    if (desc->c <= 0 || desc->c % 16) { delete P; return fail(HRNET_E_INVALID, "HRNet width c must be a positive multiple of 16"); }
    rc = build_hrnet(*P);
  } else if (desc->arch == HRNET_ARCH_POSERESNET) {
    rc = build_poseresnet(*P);
  } else {
    delete P;
    return fail(HRNET_E_INVALID, "Wrong model name.");  // SimpleHRNet.py:114
  }
  if (rc) { delete P; return rc; }
  finalize_schedule(*P);
  *out = P;
  return HRNET_OK;
}

void hrnet_plan_destroy(HrnetPlan* P) {
  if (!P) return;
  for (auto& kv : P->graphs) cudaGraphExecDestroy(kv.second);
  for (auto e : P->events) if (e) cudaEventDestroy(e);
  if (P->fork_ev) cudaEventDestroy(P->fork_ev);
  for (auto s : P->side) if (s) cudaStreamDestroy(s);
  if (P->cap) cudaStreamDestroy(P->cap);
  delete P;
}

int hrnet_plan_workspace_bytes(const HrnetPlan* P, size_t* act_bytes, size_t* weight_bytes) {
  if (!P) return fail(HRNET_E_INVALID, "null plan");
  if (act_bytes) *act_bytes = P->act_bytes;
  if (weight_bytes) *weight_bytes = P->weight_bytes;
  return HRNET_OK;
}

int hrnet_plan_num_params(const HrnetPlan* P) { return P ? (int)P->params.size() : HRNET_E_INVALID; }

int hrnet_plan_launch_count(const HrnetPlan* P) { return P ? P->launch_count : HRNET_E_INVALID; }

int hrnet_plan_param_info(const HrnetPlan* P, int i, HrnetParamInfo* out) {
  if (!P || !out || i < 0 || i >= (int)P->params.size()) return fail(HRNET_E_INVALID, "bad param index");
  const ParamInfo& pi = P->params[i];
  memset(out, 0, sizeof(*out));
  snprintf(out->conv_key, sizeof(out->conv_key), "%s", pi.conv_key.c_str());
  snprintf(out->bn_key, sizeof(out->bn_key), "%s", pi.bn_key.c_str());
  out->cout = pi.cout; out->cin = pi.cin; out->kh = pi.kh; out->kw = pi.kw;
  out->kind = P->param_kind[i]; out->sub_a = P->param_a[i]; out->sub_b = P->param_b[i];
  out->has_bias = pi.has_bias; out->w_f32 = pi.w_f32;
  out->w_offset = pi.w_offset; out->scale_offset = pi.scale_offset; out->bias_offset = pi.bias_offset;
  return HRNET_OK;
}

int hrnet_plan_describe(const HrnetPlan* P, char* buf, size_t cap, size_t* needed) {
  if (!P) return fail(HRNET_E_INVALID, "null plan");
  std::ostringstream o;
  o << "{\"arch\":" << P->desc.arch << ",\"c\":" << P->desc.c << ",\"nof_joints\":" << P->desc.nof_joints
    << ",\"height\":" << P->desc.height << ",\"width\":" << P->desc.width << ",\"max_batch\":" << P->desc.max_batch
    << ",\"act_bytes\":" << P->act_bytes << ",\"weight_bytes\":" << P->weight_bytes << ",\"input\":" << P->t_input
    << ",\"heatmaps\":" << P->t_heatmaps << ",\"tensors\":[";
  for (size_t i = 0; i < P->tensors.size(); ++i) {
    const TensorInfo& t = P->tensors[i];
    if (i) o << ",";
    o << "{\"C\":" << t.C << ",\"H\":" << t.H << ",\"W\":" << t.W << ",\"f32\":" << (t.dtype == DT_F32 ? 1 : 0)
      << ",\"offset\":" << (t.offset == (size_t)-1 ? -1LL : (long long)t.offset) << "}";
  }
  o << "],\"ops\":[";
  for (size_t i = 0; i < P->ops.size(); ++i) {
    const Op& op = P->ops[i];
    if (i) o << ",";
    o << "{\"kind\":" << op.kind << ",\"name\":\"" << op.name << "\",\"in\":" << op.in << ",\"out\":" << op.out
      << ",\"res\":" << op.res << ",\"param\":" << op.param << ",\"cin\":" << op.cin << ",\"cout\":" << op.cout
      << ",\"k\":" << op.k << ",\"stride\":" << op.stride << ",\"pad\":" << op.pad << ",\"relu\":" << op.relu
      << ",\"stream\":" << op.stream << ",\"chain\":" << op.chain << ",\"chain_pos\":" << op.chain_pos
      << ",\"xunit\":" << op.xunit << ",\"xpos\":" << op.xpos << ",\"grp\":" << op.grp << ",\"sm_frac\":" << op.sm_frac << ",\"use_tc\":" << (op.use_tc ? 1 : 0) << ",\"use_patch\":" << (op.use_patch ? 1 : 0) << ",\"nsrc\":" << op.nsrc << ",\"src\":["
      << op.src[0] << "," << op.src[1] << "," << op.src[2] << "," << op.src[3] << "],\"shift\":[" << op.shift[0] << ","
      << op.shift[1] << "," << op.shift[2] << "," << op.shift[3] << "],\"deps\":[";
    for (size_t k = 0; k < op.deps.size(); ++k) o << (k ? "," : "") << op.deps[k];
    o << "],\"tc\":{\"kc\":" << op.tc.kc << ",\"bps\":" << op.tc.bps << ",\"n_tile\":" << op.tc.n_tile
      << ",\"cs\":" << op.tc.cs << ",\"stages\":" << op.tc.stages << ",\"smem\":" << op.tc.smem_bytes << ",\"tmem_cols\":" << op.tc.tmem_cols
      << ",\"epi\":" << (op.use_patch ? op.pp.epi_tma : op.tc.epi)
      << ",\"patch_cs\":" << (op.use_patch ? op.pp.cs : 0) << ",\"patch_slots\":" << (op.use_patch ? op.pp.nslots : 0)
      << ",\"mma_warps\":" << (op.use_patch ? op.pp.mma_warps : op.tc.mma_warps) << "}}";
  }
  o << "],\"chains\":[";
  bool first_chain = true;
  for (const auto& ch : P->chains) {
    if (!ch.enabled) continue;
    if (!first_chain) o << ",";
    first_chain = false;
    o << "{\"module\":" << ch.module << ",\"branch\":" << ch.branch << ",\"patch\":" << (ch.patch ? 1 : 0) << ",\"smem\":" << ch.smem
      << ",\"share_permille\":" << ch.share << ",\"grid\":" << ch.grid << ",\"m2\":" << ch.m2 << ",\"pair\":" << (ch.pair ? 1 : 0) << ",\"tail64\":" << (ch.tail64 ? 1 : 0) << ",\"stages\":" << ch.stages << ",\"flag_stride\":" << ch.flag_stride << ",\"ctrl_off\":" << ch.ctrl_off
      << ",\"flags_off\":" << ch.flags_off << ",\"ops\":[";
    for (size_t k = 0; k < ch.ops.size(); ++k) o << (k ? "," : "") << ch.ops[k];
    o << "]}";
  }
  o << "],\"xunits\":[";
  bool first_x = true;
  for (const auto& xu : P->xunits) {
    if (!xu.enabled) continue;
    if (!first_x) o << ",";
    first_x = false;
    o << "{\"module\":" << xu.module << ",\"smem\":" << xu.smem << ",\"stages\":" << xu.stages << ",\"ops\":[";
    for (size_t k = 0; k < xu.ops.size(); ++k) o << (k ? "," : "") << xu.ops[k];
    o << "],\"sums\":[";
    for (size_t k = 0; k < xu.sums.size(); ++k) o << (k ? "," : "") << xu.sums[k];
    o << "]}";
  }
  o << "],\"sync_off\":" << P->sync_off << ",\"sync_bytes\":" << P->sync_bytes << "}";
  const std::string s = o.str();
  if (needed) *needed = s.size() + 1;
  if (buf && cap) {
    const size_t n = std::min(cap - 1, s.size());
    memcpy(buf, s.data(), n);
    buf[n] = 0;
  }
  return HRNET_OK;
}

int hrnet_plan_bind(HrnetPlan* P, void* weights_dev, size_t weight_bytes, void* workspace_dev, size_t act_bytes) {
  if (!P || !weights_dev || !workspace_dev) return fail(HRNET_E_INVALID, "null argument");
  if (weight_bytes < P->weight_bytes || act_bytes < P->act_bytes) return fail(HRNET_E_NOMEM, "buffer too small for this plan");
  if (((uintptr_t)weights_dev | (uintptr_t)workspace_dev) & 1023) return fail(HRNET_E_INVALID, "buffers must be 1024-byte aligned");
  int rc = load_driver_fns();
  if (rc) return rc;
  int dev = 0;
  cudaDeviceProp prop;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess)
    return fail(HRNET_E_CUDA, "no CUDA device");
  if (prop.major != 10) return fail(HRNET_E_CUDA, "this library contains sm_100a code only; device is sm_" + std::to_string(prop.major * 10 + prop.minor));
  P->num_sms = prop.multiProcessorCount;
  P->wbase = (uint8_t*)weights_dev;
  P->abase = (uint8_t*)workspace_dev;
  for (auto& kv : P->graphs) cudaGraphExecDestroy(kv.second);
  P->graphs.clear();
  int max_smem = 0, max_patch_smem = 0;
  for (auto& op : P->ops) {
    if (!op.use_tc) continue;
    const TensorInfo& ti = P->tensors[op.in];
    const ParamInfo& pi = P->params[op.param];
    int plh, phh, plw, phw, sub, sa, sb;
    conv_geometry(op, plh, phh, plw, phw, sub, sa, sb);
    rc = encode_im2col(&op.tmA, P->abase + ti.offset, P->desc.max_batch, ti.H, ti.W, ti.C, op.tc.kc, op.k, op.stride,
                       plh, phh, plw, phw);
    if (rc) return rc;
    rc = encode_weights(&op.tmB, P->wbase + pi.w_offset, op.cout, op.k * op.k * op.cin, op.tc.kc, op.tc.n_tile / op.tc.cs);
    if (rc) return rc;
    max_smem = std::max(max_smem, op.tc.smem_bytes);
    if (op.use_patch) {
      rc = encode_patch_maps(op, P->abase + ti.offset, P->wbase + pi.w_offset, P->desc.max_batch);
      if (rc) return rc;
      max_patch_smem = std::max(max_patch_smem, op.patch_smem);
    }
    {
      const TensorInfo& to = P->tensors[op.out];
      rc = encode_epi_maps(op, P->abase + to.offset, op.res >= 0 ? P->abase + P->tensors[op.res].offset : nullptr,
                           P->desc.max_batch, to.H, to.W);
      if (rc) return rc;
    }
  }
  {
    cudaError_t e = stem_tc_set_attributes();
    if (e != cudaSuccess) return fail(HRNET_E_CUDA, std::string("cudaFuncSetAttribute(stem): ") + cudaGetErrorString(e));
  }
  // head: fp32 weights / bias -> host, they travel as kernel parameters (head_c_kernel) when they fit
  P->head_w.clear(); P->head_b.clear();
  for (auto& op : P->ops) {
    if (op.kind != OP_HEAD) continue;
    const ParamInfo& pi = P->params[op.param];
    if (op.cout > kHeadMaxJc || op.cout * op.cin > kHeadMaxW || !head_c_supported(op.cin, op.cout)) break;
    P->head_w.resize((size_t)op.cout * op.cin); P->head_b.resize(op.cout);
    if (cudaMemcpy(P->head_w.data(), P->wbase + pi.w_offset, P->head_w.size() * 4, cudaMemcpyDeviceToHost) != cudaSuccess ||
        cudaMemcpy(P->head_b.data(), P->wbase + pi.bias_offset, P->head_b.size() * 4, cudaMemcpyDeviceToHost) != cudaSuccess)
      return fail(HRNET_E_CUDA, "cudaMemcpy(head weights)");
  }
  // branch chains: tensor maps of the member convs side by side, grid split of each module's chains, cleared flags
  {
    std::map<int, std::vector<int>> by_module;
    for (size_t c = 0; c < P->chains.size(); ++c) {
      ChainInfo& ch = P->chains[c];
      if (!ch.enabled) continue;
      by_module[ch.module].push_back((int)c);
      for (size_t k = 0; k < ch.ops.size(); ++k) {
        const Op& op = P->ops[ch.ops[k]];
        if (ch.patch) {
          int tail = 0;
          for (int j = 0; j < op.pp.nchunks; ++j) if (op.pp.mapi[j]) tail = op.pp.mapi[j];
          ch.pmaps.a[k] = op.tmPA[0];
          ch.pmaps.b[k][0] = op.tmPB[0];
          ch.pmaps.b[k][1] = op.tmPB[tail];
          ch.pmaps.a2[k] = op.tmPA[0];
          if (ch.tail64) {      // 32-channel patch box with 64-byte swizzled rows for the second chunk
            const TensorInfo& ti = P->tensors[op.in];
            rc = encode_patch(&ch.pmaps.a2[k], P->abase + ti.offset, P->desc.max_batch, op.pp.H, op.pp.W, op.pp.Cin, 32);
            if (rc) return rc;
          }
        } else {
          ch.imaps.a[k] = op.tmA;
          ch.imaps.b[k] = op.tmB;
          if (ch.pair) {      // each CTA of a pair loads half of the weight tile's rows
            const ParamInfo& pi = P->params[op.param];
            rc = encode_weights(&ch.imaps.b[k], P->wbase + pi.w_offset, op.cout, op.k * op.k * op.cin, op.tc.kc, op.tc.n_tile / 2);
            if (rc) return rc;
          }
        }
      }
      {   // BN constants of the member convs -> host (the caller's weight buffer must hold the packed weights now)
        const int cout = P->ops[ch.ops[0]].cout;
        ch.sb.assign(ch.ops.size() * (size_t)cout * 2, 0.f);
        std::vector<float> tmp((size_t)cout * 2);
        for (size_t k = 0; k < ch.ops.size(); ++k) {
          const ParamInfo& pi = P->params[P->ops[ch.ops[k]].param];
          if (cudaMemcpy(tmp.data(), P->wbase + pi.scale_offset, (size_t)cout * 4, cudaMemcpyDeviceToHost) != cudaSuccess ||
              cudaMemcpy(tmp.data() + cout, P->wbase + pi.bias_offset, (size_t)cout * 4, cudaMemcpyDeviceToHost) != cudaSuccess)
            return fail(HRNET_E_CUDA, "cudaMemcpy(BN constants of a branch chain)");
          for (int c = 0; c < cout; ++c) { ch.sb[(k * cout + c) * 2] = tmp[c]; ch.sb[(k * cout + c) * 2 + 1] = tmp[cout + c]; }
        }
      }
      for (size_t k = ch.ops.size(); k < (size_t)kChainMaxConv; ++k) {   // unused slots: any valid descriptor
        if (ch.patch) { ch.pmaps.a[k] = ch.pmaps.a[0]; ch.pmaps.a2[k] = ch.pmaps.a2[0]; ch.pmaps.b[k][0] = ch.pmaps.b[0][0]; ch.pmaps.b[k][1] = ch.pmaps.b[0][1]; }
        else { ch.imaps.a[k] = ch.imaps.a[0]; ch.imaps.b[k] = ch.imaps.b[0]; }
      }
    }
    const int cap = P->desc.tune[HRNET_TUNE_CHAIN_GRID_CAP] > 0 ? P->desc.tune[HRNET_TUNE_CHAIN_GRID_CAP] : P->num_sms;
    for (auto& kv : by_module) {
      int left = P->num_sms, big = kv.second[0];
      for (int c : kv.second) {
        ChainInfo& ch = P->chains[c];
        ch.grid = std::max(1, (int)std::lround(ch.share * 0.001 * P->num_sms));
        left -= ch.grid;
        if (ch.share > P->chains[big].share) big = c;
      }
      P->chains[big].grid = std::max(1, P->chains[big].grid + left);    // rounding remainder to the largest chain
      for (int c : kv.second) {
        ChainInfo& ch = P->chains[c];
        ch.grid = std::min(ch.grid, cap);
        if (ch.pair) ch.grid = std::max(2, ch.grid / 2 * 2);
      }
    }
    for (auto& xu : P->xunits) {
      if (!xu.enabled) continue;
      xu.sb.clear();
      for (int i : xu.ops) {        // BN constants -> host (they travel as kernel parameters)
        const Op& op = P->ops[i];
        const ParamInfo& pi = P->params[op.param];
        std::vector<float> tmp((size_t)op.cout * 2);
        if (cudaMemcpy(tmp.data(), P->wbase + pi.scale_offset, (size_t)op.cout * 4, cudaMemcpyDeviceToHost) != cudaSuccess ||
            cudaMemcpy(tmp.data() + op.cout, P->wbase + pi.bias_offset, (size_t)op.cout * 4, cudaMemcpyDeviceToHost) != cudaSuccess)
          return fail(HRNET_E_CUDA, "cudaMemcpy(BN constants of an exchange unit)");
        for (int c = 0; c < op.cout; ++c) { xu.sb.push_back(tmp[c]); xu.sb.push_back(tmp[op.cout + c]); }
      }
      for (size_t k = 0; k < (size_t)kXMaxOps; ++k) {
        const Op& op = P->ops[xu.ops[std::min(k, xu.ops.size() - 1)]];     // unused slots: any valid descriptor
        xu.maps.a[k] = op.tmA;
        xu.maps.b[k] = op.tmB;
      }
    }
    if (P->sync_bytes) {
      cudaError_t e = conv_xunit_set_attributes(kChainMaxSmem);
      if (e != cudaSuccess) return fail(HRNET_E_CUDA, std::string("cudaFuncSetAttribute(xunit): ") + cudaGetErrorString(e));
      e = cudaMemset(P->abase + P->sync_off, 0, P->sync_bytes);
      if (e != cudaSuccess) return fail(HRNET_E_CUDA, std::string("cudaMemset(chain flags): ") + cudaGetErrorString(e));
      e = conv_chain_set_attributes(kChainMaxSmem);
      if (e != cudaSuccess) return fail(HRNET_E_CUDA, std::string("cudaFuncSetAttribute(chain): ") + cudaGetErrorString(e));
    }
  }
  {
    cudaError_t e = conv_group_set_attributes(kMaxDynSmem);
    if (e != cudaSuccess) return fail(HRNET_E_CUDA, std::string("cudaFuncSetAttribute(group): ") + cudaGetErrorString(e));
  }
  if (max_patch_smem) {
    cudaError_t e = conv_patch_set_attributes(kMaxDynSmem);
    if (e != cudaSuccess) return fail(HRNET_E_CUDA, std::string("cudaFuncSetAttribute(patch): ") + cudaGetErrorString(e));
  }
  if (max_smem) {
    cudaError_t e = conv_tc_set_attributes(kMaxDynSmem);
    if (e != cudaSuccess) return fail(HRNET_E_CUDA, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e));
  }
  if (!P->side[0]) {
    for (int i = 0; i < 3; ++i)
      if (cudaStreamCreateWithFlags(&P->side[i], cudaStreamNonBlocking) != cudaSuccess) return fail(HRNET_E_CUDA, "stream create failed");
    if (cudaStreamCreateWithFlags(&P->cap, cudaStreamNonBlocking) != cudaSuccess) return fail(HRNET_E_CUDA, "stream create failed");
    P->events.resize(P->ops.size(), nullptr);
    for (size_t i = 0; i < P->ops.size(); ++i)
      if (P->ops[i].needs_event && cudaEventCreateWithFlags(&P->events[i], cudaEventDisableTiming) != cudaSuccess)
        return fail(HRNET_E_CUDA, "event create failed");
    if (cudaEventCreateWithFlags(&P->fork_ev, cudaEventDisableTiming) != cudaSuccess) return fail(HRNET_E_CUDA, "event create failed");
  }
  P->bound = true;
  return HRNET_OK;
}

}  // extern "C"

namespace {

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess) return fail(HRNET_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
  } while (0)

ConvPatchParams fill_patch_params(HrnetPlan* P, const Op& op, int n) {
  const TensorInfo& to = P->tensors[op.out];
  const ParamInfo& pi = P->params[op.param];
  ConvPatchParams p = op.pp;
  p.N = n; p.total_tiles = n * p.tiles_w * p.tiles_h;
  p.out_f32 = to.dtype == DT_F32;
  p.scale = (const float*)(P->wbase + pi.scale_offset);
  p.bias = (const float*)(P->wbase + pi.bias_offset);
  p.residual = op.res >= 0 ? (const __half*)(P->abase + P->tensors[op.res].offset) : nullptr;
  p.out = P->abase + to.offset;
  p.pdl = P->desc.tune[HRNET_TUNE_NO_PDL] ? 0 : 1;
  return p;
}

ConvTcParams fill_tc_params(HrnetPlan* P, const Op& op, int n) {
  const TensorInfo& ti = P->tensors[op.in];
  const TensorInfo& to = P->tensors[op.out];
  const ParamInfo& pi = P->params[op.param];
  int plh, phh, plw, phw, sub, sa, sb;
  conv_geometry(op, plh, phh, plw, phw, sub, sa, sb);
  const int OH = sub ? ti.H : ti.H / op.stride, OW = sub ? ti.W : ti.W / op.stride;
  ConvTcParams p{};
  p.M_total = n * OH * OW; p.OH = OH; p.OW = OW; p.OHW = OH * OW;
  p.ksize = op.k; p.stride = op.stride; p.pad_h = plh; p.pad_w = plw;
  p.sub = sub; p.sub_a = sa; p.sub_b = sb;
  p.Cin = op.cin; p.Cout = op.cout;
  p.kc = op.tc.kc; p.cpt = (op.cin + op.tc.kc - 1) / op.tc.kc; p.nkb = op.k * op.k * p.cpt; p.bps = op.tc.bps;
  p.n_tile = op.tc.n_tile; p.n_tiles = op.cout / op.tc.n_tile; p.m_tiles = (p.M_total + 127) / 128;
  p.cs = op.tc.cs;
  p.stages = op.tc.stages; p.mma_warps = op.tc.mma_warps; p.relu = op.relu; p.out_f32 = to.dtype == DT_F32; p.tmem_cols = op.tc.tmem_cols;
  p.a_blk_bytes = (int)align_up((size_t)128 * p.kc * 2, 1024);
  p.b_blk_bytes = (int)align_up((size_t)(p.n_tile / p.cs) * p.kc * 2, 1024);
  p.scale = (const float*)(P->wbase + pi.scale_offset);
  p.bias = (const float*)(P->wbase + pi.bias_offset);
  p.residual = op.res >= 0 ? (const __half*)(P->abase + P->tensors[op.res].offset) : nullptr;
  p.out = P->abase + to.offset;
  p.epi_tma = op.tc.epi; p.epi_bytes = op.tc.epi_bytes;
  p.pdl = P->desc.tune[HRNET_TUNE_NO_PDL] ? 0 : 1;
  return p;
}

// ops [first, last) form one grouped launch; returns the per-member estimated cost (for time attribution)
int launch_group(HrnetPlan* P, int first, int last, int n, cudaStream_t st, std::vector<double>* cost_out) {
  GroupLaunch g;
  std::vector<double> units, ucost;
  std::vector<int*> slot;
  for (int i = first; i < last; ++i) {
    const Op& op = P->ops[i];
    if (op.use_patch) {
      const int k = g.n_patch++;
      g.pp[k] = fill_patch_params(P, op, n);
      g.patch_maps_a[k] = op.tmPA; g.patch_maps_b[k] = op.tmPB;
      units.push_back((double)g.pp[k].total_tiles); slot.push_back(&g.patch_ctas[k]);
      g.smem_bytes = std::max(g.smem_bytes, op.patch_smem);
    } else {
      const int k = g.n_igemm++;
      g.ip[k] = fill_tc_params(P, op, n);
      g.igemm_map_a[k] = op.tmA; g.igemm_map_b[k] = op.tmB;
      units.push_back((double)g.ip[k].m_tiles * g.ip[k].n_tiles); slot.push_back(&g.igemm_ctas[k]);
      g.smem_bytes = std::max(g.smem_bytes, op.tc.smem_bytes);
    }
    ucost.push_back(op.work);
  }
  // CTA split: greedily hand each SM to the problem that currently finishes last (whole tile rounds)
  const int np = (int)units.size();
  std::vector<int> c(np, 1);
  auto t = [&](int k) { return std::ceil(units[k] / c[k]) * ucost[k]; };
  for (int k = 0; k < np; ++k) c[k] = 1;
  int left = P->num_sms - np;
  while (left > 0) {
    int worst = 0;
    for (int k = 1; k < np; ++k) if (t(k) > t(worst)) worst = k;
    if (c[worst] >= (int)units[worst]) {            // cannot use more CTAs than tiles: give it to the next worst that can
      int alt = -1;
      for (int k = 0; k < np; ++k) if (c[k] < (int)units[k] && (alt < 0 || t(k) > t(alt))) alt = k;
      if (alt < 0) break;
      worst = alt;
    }
    ++c[worst]; --left;
  }
  for (int k = 0; k < np; ++k) *slot[k] = c[k];
  if (cost_out) { cost_out->clear(); for (int k = 0; k < np; ++k) cost_out->push_back(units[k] * ucost[k]); }
  g.pdl = P->desc.tune[HRNET_TUNE_NO_PDL] ? 0 : 1;
  if (P->desc.tune[HRNET_TUNE_DEBUG] & 4) {   // debug: per-problem finish times of one grouped launch
    long long* dev = nullptr;
    const int grid = P->num_sms;
    if (cudaMalloc(&dev, (size_t)grid * 32 * sizeof(long long)) == cudaSuccess) {
      cudaMemset(dev, 0, (size_t)grid * 32 * sizeof(long long));
      for (int k = 0; k < g.n_patch; ++k) g.pp[k].dbg = dev;
      for (int k = 0; k < g.n_igemm; ++k) g.ip[k].dbg = dev;
      CK(launch_conv_group(g, st));
      cudaStreamSynchronize(st);
      std::vector<long long> h((size_t)grid * 32);
      cudaMemcpy(h.data(), dev, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
      cudaFree(dev);
      long long t0 = LLONG_MAX;
      for (int b = 0; b < grid; ++b) if (h[(size_t)b * 32 + 16]) t0 = std::min(t0, h[(size_t)b * 32 + 16]);
      int begin = 0;
      fprintf(stderr, "[dbg-group] first=%s n=%d:", P->ops[first].name.c_str(), n);
      for (int k = 0; k < np; ++k) {
        long long done = 0, ops0 = 0; double mma = 0;
        for (int b = begin; b < begin + c[k]; ++b) {
          done = std::max(done, h[(size_t)b * 32 + 20] - t0);
          ops0 = std::max(ops0, h[(size_t)b * 32 + 18] - t0);
          mma += (double)h[(size_t)b * 32 + 6] / c[k];
        }
        fprintf(stderr, " [%s ctas=%d units=%.0f ucost=%.0f est=%.0f clk | first_operands=%lld ns roles_done=%lld ns mma_issue=%.0f clk]",
                P->ops[first + k].use_patch ? "patch" : "igemm", c[k], units[k], ucost[k], std::ceil(units[k] / c[k]) * ucost[k], ops0, done, mma);
        begin += c[k];
      }
      fprintf(stderr, "\n");
      return 0;
    }
  }
  CK(launch_conv_group(g, st));
  return 0;
}

// One launch for the whole branch chain `c` at batch n on `grid` CTAs (<= 0: the chain's share inside the forward).
int launch_chain_op(HrnetPlan* P, int c, int n, int grid, cudaStream_t st, long long* dbg = nullptr) {
  const ChainInfo& ch = P->chains[c];
  const Op& o0 = P->ops[ch.ops[0]];
  if (grid <= 0) grid = ch.grid;
  auto fill_convs = [&](ChainConv* cv) {
    for (size_t k = 0; k < ch.ops.size(); ++k) {
      const Op& op = P->ops[ch.ops[k]];
      const ParamInfo& pi = P->params[op.param];
      cv[k].scale = (const float*)(P->wbase + pi.scale_offset);
      cv[k].bias = (const float*)(P->wbase + pi.bias_offset);
      cv[k].residual = op.res >= 0 ? (const __half*)(P->abase + P->tensors[op.res].offset) : nullptr;
      cv[k].out = (__half*)(P->abase + P->tensors[op.out].offset);
      cv[k].relu = op.relu;
      cv[k].pad_ = 0;
    }
  };
  auto fill_sb = [&](auto& sb) {
    const int cout = o0.cout;
    for (size_t k = 0; k < ch.ops.size(); ++k)
      for (int c = 0; c < cout; ++c) sb[k][c] = make_float2(ch.sb[(k * cout + c) * 2], ch.sb[(k * cout + c) * 2 + 1]);
  };
  if (ch.patch) {
    static thread_local ChainPatchParams p;      // (the parameter blocks are tens of KB: keep them off the stack)
    p = ChainPatchParams{};
    p.nconv = (int)ch.ops.size();
    fill_sb(p.sb);
    p.pp = o0.pp;
    p.pp.N = n; p.pp.total_tiles = n * p.pp.tiles_w * p.pp.tiles_h;
    if (8 * p.pp.Cout <= 512 && P->desc.tune[HRNET_TUNE_PATCH_NACC] == 0) {      // eight accumulator buffers (conv_chain.cu)
      p.pp.nacc = 8; p.pp.nacc_log2 = 3;
      int cols = 32;
      while (cols < 8 * p.pp.Cout) cols *= 2;
      p.pp.tmem_cols = cols;
    }
    p.unit_stride = ch.flag_stride;
    p.chunk = p.pp.tiles_w;   // one ticket = one tile row
    p.skip = P->desc.tune[HRNET_TUNE_CHAIN_SKIP];
    p.tail64 = ch.tail64 ? 1 : 0;
    p.pdl = P->desc.tune[HRNET_TUNE_NO_PDL] ? 0 : 1;
    p.ctrl = (unsigned*)(P->abase + ch.ctrl_off);
    p.counters = (unsigned*)(P->abase + ch.flags_off);
    p.dbg = dbg;
    fill_convs(p.conv);
    if (p.pp.total_tiles == 0) return 0;
    grid = std::max(1, std::min(grid, p.pp.total_tiles));
    CK(launch_chain_patch(ch.pmaps, p, ch.smem, grid, st));
  } else {
    const ConvTcParams t = fill_tc_params(P, o0, n);
    static thread_local ChainIgemmParams p;
    p = ChainIgemmParams{};
    p.nconv = (int)ch.ops.size();
    fill_sb(p.sb);
    p.M_total = t.M_total; p.OH = t.OH; p.OW = t.OW; p.OHW = t.OHW; p.C = o0.cin;
    p.cpt = t.cpt; p.nkb = t.nkb; p.bps = t.bps; p.n_tile = t.n_tile; p.n_tiles = t.n_tiles; p.m_tiles = t.m_tiles;
    p.stages = ch.stages; p.tmem_cols = t.tmem_cols; p.a_blk_bytes = t.a_blk_bytes; p.b_blk_bytes = t.b_blk_bytes;
    p.pdl = t.pdl;
    p.m2 = ch.m2; p.pair = ch.pair ? 1 : 0;
    p.skip = P->desc.tune[HRNET_TUNE_CHAIN_SKIP];
    const int tpu = ch.pair ? 2 : ch.m2;
    p.units = (p.m_tiles + tpu - 1) / tpu;
    if (ch.pair) p.b_blk_bytes = (int)align_up((size_t)(t.n_tile / 2) * t.kc * 2, 1024);
    p.unit_stride = ch.flag_stride;
    p.ctrl = (unsigned*)(P->abase + ch.ctrl_off);
    p.counters = (unsigned*)(P->abase + ch.flags_off);
    p.dbg = dbg;
    fill_convs(p.conv);
    if (p.m_tiles * p.n_tiles == 0) return 0;
    grid = std::max(1, std::min(grid, p.units * (ch.pair ? 2 : 1)));
    if (ch.pair) grid = std::max(2, grid / 2 * 2);
    CK(launch_chain_igemm(ch.imaps, p, ch.smem, grid, st));
  }
  return 0;
}

// in_u8: `in_ext` holds NHWC BGR uint8 images (hrnet_forward_u8) instead of the NCHW fp32 network input
// One launch for the whole exchange unit `x` at batch n.
int launch_xunit_op(HrnetPlan* P, int x, int n, cudaStream_t st, long long* dbg = nullptr) {
  const XUnitInfo& xu = P->xunits[x];
  static thread_local XUnitParams p;      // (20 KB of BN constants: off the stack)
  p = XUnitParams{};
  p.nops = (int)xu.ops.size();
  for (size_t i = 0; i + 1 < xu.sb.size() && i / 2 < (size_t)kXMaxSb; i += 2) p.sb[i / 2] = make_float2(xu.sb[i], xu.sb[i + 1]);
  p.stages = xu.stages; p.tmem_cols = xu.tmem_cols; p.a_blk_bytes = 16384; p.b_blk_bytes = xu.b_blk;
  p.pdl = P->desc.tune[HRNET_TUNE_NO_PDL] ? 0 : 1;
  p.ctrl = (unsigned*)(P->abase + xu.ctrl_off);
  p.counters = (unsigned*)(P->abase + xu.counters_off);
  p.dbg = dbg;
  int kb = 0, sb_off = 0;
  for (int k = 0; k < p.nops; ++k) {
    const Op& op = P->ops[xu.ops[k]];
    const ConvTcParams t = fill_tc_params(P, op, n);
    XOp& o = p.op[k];
    o.sb_off = sb_off; sb_off += op.cout;
    o.M_total = t.M_total; o.OH = t.OH; o.OW = t.OW; o.OHW = t.OHW;
    o.ksize = t.ksize; o.stride = t.stride; o.pad = t.pad_h;
    o.Cin = t.Cin; o.Cout = t.Cout; o.cpt = t.cpt; o.nkb = t.nkb;
    o.n_tile = t.n_tile; o.n_tiles = t.n_tiles; o.m_tiles = t.m_tiles;
    o.relu = t.relu;
    o.dep = -1; o.dep_need = 0;
    for (int j = 0; j < p.nops; ++j)
      if (P->ops[xu.ops[j]].out == op.in) { o.dep = j; }
    o.kb0 = kb; kb += o.nkb;
    o.scale = t.scale; o.bias = t.bias; o.out = (__half*)t.out;
  }
  for (int k = 0; k < p.nops; ++k)
    if (p.op[k].dep >= 0) p.op[k].dep_need = p.op[p.op[k].dep].m_tiles * p.op[p.op[k].dep].n_tiles;
  // the sums: one XSum per OP_FUSE of the module, its level = the deepest down-chain step it reads
  p.nsums = (int)xu.sums.size();
  int sum_level[4] = {-1, -1, -1, -1};
  for (int q = 0; q < p.nsums; ++q) {
    const Op& f = P->ops[xu.sums[q]];
    const TensorInfo& to = P->tensors[f.out];
    XSum& sm = p.sum[q];
    sm.H = to.H; sm.W = to.W; sm.C = to.C; sm.nsrc = f.nsrc; sm.relu = f.relu;
    sm.npix = (long long)n * to.H * to.W;
    sm.nchunks = (int)((sm.npix + kXSumChunk - 1) / kXSumChunk);
    sm.out = (__half*)(P->abase + to.offset);
    for (int j = 0; j < 4; ++j) { sm.shift[j] = 0; sm.dep[j] = -1; sm.dep_need[j] = 0; sm.src[j] = nullptr; }
    for (int j = 0; j < f.nsrc; ++j) {
      sm.src[j] = (const __half*)(P->abase + P->tensors[f.src[j]].offset);
      sm.shift[j] = f.shift[j];
      for (int k = 0; k < p.nops; ++k)
        if (P->ops[xu.ops[k]].out == f.src[j]) {
          sm.dep[j] = k; sm.dep_need[j] = p.op[k].m_tiles * p.op[k].n_tiles;
          sum_level[q] = std::max(sum_level[q], P->ops[xu.ops[k]].xlevel);
        }
    }
  }
  // ticket sequence: convs of level L (plan order), then the sums whose deepest source is level L
  int ticket = 0;
  for (int q = 0; q < p.nsums; ++q)
    if (sum_level[q] < 0) { p.sum[q].ticket0 = ticket; ticket += p.sum[q].nchunks; }
  for (int L = 0; L < 8; ++L) {
    for (int k = 0; k < p.nops; ++k)
      if (P->ops[xu.ops[k]].xlevel == L) { p.op[k].ticket0 = ticket; ticket += p.op[k].m_tiles; }
    for (int q = 0; q < p.nsums; ++q)
      if (sum_level[q] == L) { p.sum[q].ticket0 = ticket; ticket += p.sum[q].nchunks; }
  }
  p.total_tickets = ticket; p.total_kb = kb;
  if (ticket == 0) return 0;
  CK(launch_xunit(xu.maps, p, xu.smem, std::min(P->num_sms, ticket), st));
  return 0;
}

int launch_op(HrnetPlan* P, const Op& op, int n, const float* in_ext, float* hm_ext, float* joints, int32_t* idx,
              const float* boxes, cudaStream_t st, bool in_u8 = false, bool no_hm = false) {
  auto tptr = [&](int t) -> uint8_t* { return P->abase + P->tensors[t].offset; };
  switch (op.kind) {
    case OP_STEM:
    case OP_STEM7: {
      const ParamInfo& pi = P->params[op.param];
      // conv1 runs on the tensor cores unless the SIMT cross-check path is forced
      if (in_u8) {
        if (op.kind != OP_STEM) return fail(HRNET_E_INVALID, "uint8 image input is implemented for the HRNet stem only");
        CK(launch_stem_tc_u8((const uint8_t*)in_ext, (const float*)(P->wbase + pi.w_offset),
                             (const float*)(P->wbase + pi.scale_offset), (const float*)(P->wbase + pi.bias_offset),
                             (__half*)tptr(op.out), n, P->desc.height, P->desc.width, P->num_sms, st));
        return 0;
      }
      if (op.kind == OP_STEM && !(P->desc.flags & HRNET_FLAG_FORCE_SIMT)) {
        CK(launch_stem_tc(in_ext, (const float*)(P->wbase + pi.w_offset), (const float*)(P->wbase + pi.scale_offset),
                          (const float*)(P->wbase + pi.bias_offset), (__half*)tptr(op.out), n, P->desc.height, P->desc.width,
                          P->num_sms, st));
        return 0;
      }
      auto fn = op.kind == OP_STEM ? launch_stem : launch_stem7;
      CK(fn(in_ext, (const float*)(P->wbase + pi.w_offset), (const float*)(P->wbase + pi.scale_offset),
            (const float*)(P->wbase + pi.bias_offset), (__half*)tptr(op.out), n, P->desc.height, P->desc.width, st));
      return 0;
    }
    case OP_MAXPOOL: {
      const TensorInfo& ti = P->tensors[op.in];
      CK(launch_maxpool((const __half*)tptr(op.in), (__half*)tptr(op.out), n, ti.H, ti.W, ti.C, st));
      return 0;
    }
    case OP_CONV: {
      const TensorInfo& ti = P->tensors[op.in];
      const TensorInfo& to = P->tensors[op.out];
      const ParamInfo& pi = P->params[op.param];
      int plh, phh, plw, phw, sub, sa, sb;
      conv_geometry(op, plh, phh, plw, phw, sub, sa, sb);
      const int OH = sub ? ti.H : ti.H / op.stride, OW = sub ? ti.W : ti.W / op.stride;
      if (op.use_patch) {
        ConvPatchParams p = op.pp;
        p.N = n; p.total_tiles = n * p.tiles_w * p.tiles_h;
        p.out_f32 = to.dtype == DT_F32;
        p.scale = (const float*)(P->wbase + pi.scale_offset);
        p.bias = (const float*)(P->wbase + pi.bias_offset);
        p.residual = op.res >= 0 ? (const __half*)tptr(op.res) : nullptr;
        p.out = tptr(op.out);
        p.pdl = P->desc.tune[HRNET_TUNE_NO_PDL] ? 0 : 1;
        if (p.total_tiles == 0) return 0;
        const int cap = std::max(1, (int)std::lround(op.sm_frac * P->num_sms));
        CK(launch_conv_patch(op.tmPA, op.tmPB, op.tmOR, p, op.patch_smem, std::min(std::max(p.cs, cap / p.cs * p.cs), conv_patch_grid(p, P->num_sms)), st));
      } else if (op.use_tc) {
        ConvTcParams p{};
        p.M_total = n * OH * OW; p.OH = OH; p.OW = OW; p.OHW = OH * OW;
        p.ksize = op.k; p.stride = op.stride; p.pad_h = plh; p.pad_w = plw;
        p.sub = sub; p.sub_a = sa; p.sub_b = sb;
        p.Cin = op.cin; p.Cout = op.cout;
        p.kc = op.tc.kc; p.cpt = (op.cin + op.tc.kc - 1) / op.tc.kc; p.nkb = op.k * op.k * p.cpt; p.bps = op.tc.bps;
        p.n_tile = op.tc.n_tile; p.n_tiles = op.cout / op.tc.n_tile; p.m_tiles = (p.M_total + 127) / 128;
        p.cs = op.tc.cs;
        p.stages = op.tc.stages; p.mma_warps = op.tc.mma_warps; p.relu = op.relu; p.out_f32 = to.dtype == DT_F32; p.tmem_cols = op.tc.tmem_cols;
        p.a_blk_bytes = (int)align_up((size_t)128 * p.kc * 2, 1024);
        p.b_blk_bytes = (int)align_up((size_t)(p.n_tile / p.cs) * p.kc * 2, 1024);
        p.scale = (const float*)(P->wbase + pi.scale_offset);
        p.bias = (const float*)(P->wbase + pi.bias_offset);
        p.residual = op.res >= 0 ? (const __half*)tptr(op.res) : nullptr;
        p.out = tptr(op.out);
        p.epi_tma = op.tc.epi; p.epi_bytes = op.tc.epi_bytes;
        p.pdl = P->desc.tune[HRNET_TUNE_NO_PDL] ? 0 : 1;
        const int tiles = p.m_tiles * p.n_tiles;
        if (tiles == 0) return 0;
        const int cap = std::max(p.cs, (int)std::lround(op.sm_frac * P->num_sms) / p.cs * p.cs);
        CK(launch_conv_tc(op.tmA, op.tmB, op.tmOR, p, op.tc.smem_bytes, std::min(cap, conv_tc_grid(p, op.tc.smem_bytes, P->num_sms)), st));
      } else {
        if (sub) return fail(HRNET_E_INVALID, "transposed-conv phases are not wired to the SIMT kernel yet");
        ConvSimtParams p{};
        p.N = n; p.IH = ti.H; p.IW = ti.W; p.OH = OH; p.OW = OW; p.Cin = op.cin; p.Cout = op.cout; p.ksize = op.k;
        p.stride = op.stride; p.pad = op.pad; p.relu = op.relu; p.out_f32 = to.dtype == DT_F32;
        p.in = (const __half*)tptr(op.in); p.w = (const __half*)(P->wbase + pi.w_offset);
        p.scale = (const float*)(P->wbase + pi.scale_offset); p.bias = (const float*)(P->wbase + pi.bias_offset);
        p.residual = op.res >= 0 ? (const __half*)tptr(op.res) : nullptr;
        p.out = tptr(op.out);
        CK(launch_conv_simt(p, st));
      }
      return 0;
    }
    case OP_FUSE: {
      const TensorInfo& to = P->tensors[op.out];
      FuseParams p{};
      p.N = n; p.H = to.H; p.W = to.W; p.C = to.C; p.nsrc = op.nsrc; p.relu = op.relu;
      for (int j = 0; j < op.nsrc; ++j) {
        p.src[j] = tptr(op.src[j]); p.shift[j] = op.shift[j]; p.f32[j] = P->tensors[op.src[j]].dtype == DT_F32;
      }
      p.out = (__half*)tptr(op.out);
      CK(launch_fuse(p, st));
      return 0;
    }
    case OP_HEAD: {
      const TensorInfo& ti = P->tensors[op.in];
      const ParamInfo& pi = P->params[op.param];
      float* out = hm_ext ? hm_ext : (float*)tptr(op.out);
      if (!P->head_w.empty()) {
        // weights as kernel parameters; when nobody wants the heat-maps (no_hm) they are not written at all: the
        // block-level argmax candidates go to the finishing kernel launched for OP_ARGMAX
        static thread_local HeadParams hp;
        hp.in = (const __half*)tptr(op.in);
        hp.out = no_hm ? nullptr : out;
        hp.pval = no_hm ? (float*)(P->abase + P->off_pval) : nullptr;
        hp.pidx = no_hm ? (int*)(P->abase + P->off_pidx) : nullptr;
        hp.N = n; hp.HW = ti.H * ti.W; hp.Cin = op.cin; hp.J = op.cout; hp.nblk = P->head_nblk; hp.pad_ = 0;
        memcpy(hp.bias, P->head_b.data(), P->head_b.size() * 4);
        hp.w_dev = (const float*)(P->wbase + pi.w_offset);
        CK(launch_head_c(hp, st));
        return 0;
      }
      CK(launch_head((const __half*)tptr(op.in), (const float*)(P->wbase + pi.w_offset),
                     (const float*)(P->wbase + pi.bias_offset), out, n, ti.H * ti.W, op.cin, op.cout, st));
      return 0;
    }
    case OP_ARGMAX: {
      if (no_hm && !P->head_w.empty()) {
        CK(launch_head_argmax_finish((const float*)(P->abase + P->off_pval), (const int*)(P->abase + P->off_pidx), n, op.cout,
                                     P->head_nblk, P->Hh, P->Wh, boxes, joints, idx, st));
        return 0;
      }
      const float* hm = hm_ext ? hm_ext : (const float*)tptr(op.in);
      CK(launch_argmax(hm, n, op.cout, P->Hh, P->Wh, boxes, joints, idx, st));
      return 0;
    }
  }
  return fail(HRNET_E_INVALID, "unknown op kind");
}

// launches ops [first, last) with stream fork/join according to the dependency lists
int run_range(HrnetPlan* P, int first, int last, int n, const float* in_ext, float* hm_ext, float* joints,
              int32_t* idx, const float* boxes, cudaStream_t s0, bool in_u8 = false) {
  auto stream_of = [&](int s) { return s == 0 ? s0 : P->side[s - 1]; };
  for (int i = first; i < last; ++i) {
    const Op& op = P->ops[i];
    cudaStream_t st = stream_of(op.stream);
    if (op.chain >= 0) {
      // the whole branch chain is issued where its first conv stands; the other members are part of that launch
      if (op.chain_pos != 0) continue;
      const ChainInfo& ch = P->chains[op.chain];
      for (int m : ch.ops)
        for (int dpi : P->ops[m].deps)
          if (P->ops[dpi].stream != op.stream) CK(cudaStreamWaitEvent(st, P->events[dpi], 0));
      int rc = launch_chain_op(P, op.chain, n, (P->desc.flags & HRNET_FLAG_SERIAL) ? P->num_sms : 0, st);
      if (rc) return rc;
      for (int m : ch.ops)
        if (P->ops[m].needs_event) CK(cudaEventRecord(P->events[m], st));
      continue;
    }
    if (op.xunit >= 0) {
      // the whole exchange unit is issued where its first conv stands; the other members are part of that launch
      if (op.xpos != 0) continue;
      const XUnitInfo& xu = P->xunits[op.xunit];
      for (int dpi : op.deps)               // (the first member carries every member's external dependencies)
        if (P->ops[dpi].stream != op.stream) CK(cudaStreamWaitEvent(st, P->events[dpi], 0));
      int rc = launch_xunit_op(P, op.xunit, n, st);
      if (rc) return rc;
      for (int m : xu.ops)
        if (P->ops[m].needs_event) CK(cudaEventRecord(P->events[m], st));
      for (int m : xu.sums)
        if (P->ops[m].needs_event) CK(cudaEventRecord(P->events[m], st));
      continue;
    }
    if (op.grp >= 0) {
      int j = i;
      while (j < last && P->ops[j].grp == op.grp) ++j;
      for (int m = i; m < j; ++m)
        for (int dpi : P->ops[m].deps)
          if (P->ops[dpi].stream != P->ops[m].stream) CK(cudaStreamWaitEvent(st, P->events[dpi], 0));
      int rc = launch_group(P, i, j, n, st, nullptr);
      if (rc) return rc;
      for (int m = i; m < j; ++m)
        if (P->ops[m].needs_event) CK(cudaEventRecord(P->events[m], st));
      i = j - 1;
      continue;
    }
    for (int dpi : op.deps)
      if (P->ops[dpi].stream != op.stream) CK(cudaStreamWaitEvent(st, P->events[dpi], 0));
    int rc = launch_op(P, op, n, in_ext, hm_ext, joints, idx, boxes, st, in_u8);
    if (rc) return rc;
    if (op.needs_event) CK(cudaEventRecord(P->events[i], st));
    // debug (HRNET_FLAG_NO_GRAPH only): synchronise after every op and print a hash of the n images of its output, so
    // two runs can be diffed op by op (tools/dbg_invariance.py)
    if ((P->desc.flags & HRNET_FLAG_NO_GRAPH) && op.out >= 0 && P->tensors[op.out].offset != (size_t)-1) {
      if (P->desc.tune[HRNET_TUNE_DEBUG] & 2) {
        CK(cudaStreamSynchronize(st));
        const size_t bytes = P->tensors[op.out].bytes(n);
        std::vector<uint8_t> h(bytes);
        CK(cudaMemcpy(h.data(), P->abase + P->tensors[op.out].offset, bytes, cudaMemcpyDeviceToHost));
        uint64_t hash = 1469598103934665603ull;
        const uint64_t* w = reinterpret_cast<const uint64_t*>(h.data());
        for (size_t k = 0; k < bytes / 8; ++k) { hash ^= w[k]; hash *= 1099511628211ull; }
        fprintf(stderr, "[cs] %d %s %016llx\n", i, op.name.c_str(), (unsigned long long)hash);
      }
    }
  }
  return 0;
}

}  // namespace

// keep_hm: the caller will read the plan's internal heat-map tensor afterwards (host variants with heatmaps_h); with
// neither that nor an external `heatmaps` buffer the head runs fused with the argmax and writes no heat-maps
static int forward_impl(HrnetPlan* P, const float* in, int n, float* heatmaps, float* joints, int32_t* argmax_idx,
                        const float* boxes, void* stream, bool in_u8, bool keep_hm = false) {
  if (!P) return fail(HRNET_E_INVALID, "null plan");
  if (!P->bound) return fail(HRNET_E_STATE, "hrnet_plan_bind must be called before hrnet_forward");
  if (n < 0 || n > P->desc.max_batch) return fail(HRNET_E_INVALID, "n out of range [0, max_batch]");
  if (n == 0) return HRNET_OK;
  if (!in || !joints) return fail(HRNET_E_INVALID, "null input / joints pointer");
  cudaStream_t s0 = (cudaStream_t)stream;
  const int nops = (int)P->ops.size();
  // ops touching caller pointers stay outside the graph: first (stem) and the last two (head, argmax)
  const int g_first = 1, g_last = nops - 2;
  int rc = run_range(P, 0, g_first, n, in, heatmaps, joints, argmax_idx, boxes, s0, in_u8);
  if (rc) return rc;
  if (P->desc.flags & HRNET_FLAG_NO_GRAPH) {
    rc = run_range(P, g_first, g_last, n, in, heatmaps, joints, argmax_idx, boxes, s0);
    if (rc) return rc;
  } else {
    auto it = P->graphs.find(n);
    if (it == P->graphs.end()) {
      // head deps on side streams are recorded inside the capture and joined by the explicit waits below
      cudaGraph_t g = nullptr;
      cudaStream_t cs = P->cap;
      CK(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
      rc = run_range(P, g_first, g_last, n, in, heatmaps, joints, argmax_idx, boxes, cs);
      if (!rc) {
        // join every side stream back into s0 so the capture can end
        const Op& head = P->ops[g_last];
        for (int dpi : head.deps)
          if (P->ops[dpi].stream != 0) {
            cudaError_t e = cudaStreamWaitEvent(cs, P->events[dpi], 0);
            if (e != cudaSuccess) { rc = fail(HRNET_E_CUDA, std::string("join: ") + cudaGetErrorString(e)); break; }
          }
      }
      cudaError_t e = cudaStreamEndCapture(cs, &g);
      if (rc) { if (g) cudaGraphDestroy(g); return rc; }
      if (e != cudaSuccess) return fail(HRNET_E_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
      cudaGraphExec_t ge = nullptr;
      e = cudaGraphInstantiate(&ge, g, 0);
      cudaGraphDestroy(g);
      if (e != cudaSuccess) return fail(HRNET_E_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
      it = P->graphs.emplace(n, ge).first;
    }
    CK(cudaGraphLaunch(it->second, s0));
  }
  // head + argmax on s0 (side streams were joined either by the graph or by the waits in run_range)
  for (int i = g_last; i < nops; ++i) {
    const Op& op = P->ops[i];
    if (P->desc.flags & HRNET_FLAG_NO_GRAPH)
      for (int dpi : op.deps)
        if (P->ops[dpi].stream != 0) CK(cudaStreamWaitEvent(s0, P->events[dpi], 0));
    rc = launch_op(P, op, n, in, heatmaps, joints, argmax_idx, boxes, s0, false, heatmaps == nullptr && !keep_hm);
    if (rc) return rc;
  }
  return HRNET_OK;
}

extern "C" {

int hrnet_forward(HrnetPlan* P, const float* in, int n, float* heatmaps, float* joints, int32_t* argmax_idx,
                  const float* boxes, void* stream) {
  return forward_impl(P, in, n, heatmaps, joints, argmax_idx, boxes, stream, false);
}

int hrnet_forward_u8(HrnetPlan* P, const uint8_t* images_nhwc_bgr, int n, float* heatmaps, float* joints,
                     int32_t* argmax_idx, const float* boxes, void* stream) {
  if (!P) return fail(HRNET_E_INVALID, "null plan");
  if (P->desc.arch != HRNET_ARCH_HRNET) return fail(HRNET_E_INVALID, "uint8 image input is implemented for HRNet only");
  return forward_impl(P, reinterpret_cast<const float*>(images_nhwc_bgr), n, heatmaps, joints, argmax_idx, boxes, stream, true);
}

static int forward_host_u8_impl(HrnetPlan* P, const uint8_t* images_h, int n, float* heatmaps_h, float* joints_h,
                                int32_t* idx_h, const float* boxes_h, void* stream, bool sync) {
  if (!P) return fail(HRNET_E_INVALID, "null plan");
  if (!P->bound) return fail(HRNET_E_STATE, "hrnet_plan_bind must be called before hrnet_forward_host_u8");
  if (n < 0 || n > P->desc.max_batch) return fail(HRNET_E_INVALID, "n out of range [0, max_batch]");
  if (n == 0) return HRNET_OK;
  if (!images_h || !joints_h) return fail(HRNET_E_INVALID, "null input / joints pointer");
  cudaStream_t s0 = (cudaStream_t)stream;
  const int J = P->desc.nof_joints;
  uint8_t* in_d = P->abase + P->off_in_stage;
  float* joints_d = (float*)(P->abase + P->off_joints);
  int32_t* idx_d = (int32_t*)(P->abase + P->off_idx);
  float* boxes_d = boxes_h ? (float*)(P->abase + P->off_boxes) : nullptr;
  CK(cudaMemcpyAsync(in_d, images_h, (size_t)n * 3 * P->desc.height * P->desc.width, cudaMemcpyHostToDevice, s0));
  if (boxes_h) CK(cudaMemcpyAsync(boxes_d, boxes_h, (size_t)n * 16, cudaMemcpyHostToDevice, s0));
  if (P->desc.arch != HRNET_ARCH_HRNET) return fail(HRNET_E_INVALID, "uint8 image input is implemented for HRNet only");
  int rc = forward_impl(P, reinterpret_cast<const float*>(in_d), n, nullptr, joints_d, idx_d, boxes_d, stream, true, heatmaps_h != nullptr);
  if (rc) return rc;
  CK(cudaMemcpyAsync(joints_h, joints_d, (size_t)n * J * 12, cudaMemcpyDeviceToHost, s0));
  if (idx_h) CK(cudaMemcpyAsync(idx_h, idx_d, (size_t)n * J * 4, cudaMemcpyDeviceToHost, s0));
  if (heatmaps_h)
    CK(cudaMemcpyAsync(heatmaps_h, P->abase + P->tensors[P->t_heatmaps].offset, (size_t)n * J * P->Hh * P->Wh * 4,
                       cudaMemcpyDeviceToHost, s0));
  if (sync) CK(cudaStreamSynchronize(s0));
  return HRNET_OK;
}

int hrnet_forward_host_u8(HrnetPlan* P, const uint8_t* images_h, int n, float* heatmaps_h, float* joints_h,
                          int32_t* idx_h, const float* boxes_h, void* stream) {
  return forward_host_u8_impl(P, images_h, n, heatmaps_h, joints_h, idx_h, boxes_h, stream, true);
}

int hrnet_forward_host_u8_async(HrnetPlan* P, const uint8_t* images_h, int n, float* heatmaps_h, float* joints_h,
                                int32_t* idx_h, const float* boxes_h, void* stream) {
  return forward_host_u8_impl(P, images_h, n, heatmaps_h, joints_h, idx_h, boxes_h, stream, false);
}

int hrnet_profile_ops(HrnetPlan* P, const float* in, int n, float* usec_per_op, int iters, void* stream) {
  if (!P || !in || !usec_per_op || iters <= 0) return fail(HRNET_E_INVALID, "bad argument");
  if (!P->bound) return fail(HRNET_E_STATE, "hrnet_plan_bind must be called first");
  if (n <= 0 || n > P->desc.max_batch) return fail(HRNET_E_INVALID, "n out of range");
  cudaStream_t s0 = (cudaStream_t)stream;
  const int nops = (int)P->ops.size();
  float* joints = (float*)(P->abase + P->off_joints);
  int32_t* idx = (int32_t*)(P->abase + P->off_idx);
  std::vector<cudaEvent_t> ev(nops + 1);
  for (auto& e : ev) CK(cudaEventCreate(&e));
  std::vector<std::vector<float>> t(nops, std::vector<float>(iters));
  std::vector<cudaEvent_t> xunit_ev(2 * P->xunits.size(), nullptr);
  for (size_t x = 0; x < P->xunits.size(); ++x)
    if (P->xunits[x].enabled) { CK(cudaEventCreate(&xunit_ev[2 * x])); CK(cudaEventCreate(&xunit_ev[2 * x + 1])); }
  std::vector<cudaEvent_t> chain_ev(2 * P->chains.size(), nullptr);
  for (size_t c = 0; c < P->chains.size(); ++c)
    if (P->chains[c].enabled) { CK(cudaEventCreate(&chain_ev[2 * c])); CK(cudaEventCreate(&chain_ev[2 * c + 1])); }
  for (int it = -1; it < iters; ++it) {   // iteration -1 = warm-up
    CK(cudaEventRecord(ev[0], s0));
    std::vector<std::pair<int, int>> spans;          // grouped launches: [first, last)
    std::vector<std::vector<double>> span_cost;
    for (int i = 0; i < nops; ++i) {
      const Op& op = P->ops[i];
      if (op.chain >= 0) {
        // Branch chains are timed the way they run inside a forward: the chains of one StageModule are launched
        // together on their own streams with their in-forward grids (they share the SMs) where the module's first conv
        // stands; the time from the first launch to the last completion is split evenly between the module's branch convs
        // (equal FLOPs) below.  chain_ev[2c], [2c+1] of the module's FIRST chain bracket the module.
        const int mod = P->chains[op.chain].module;
        int first_c = -1;
        std::vector<int> mod_chains;
        for (size_t c = 0; c < P->chains.size(); ++c)
          if (P->chains[c].enabled && P->chains[c].module == mod) { if (first_c < 0) first_c = (int)c; mod_chains.push_back((int)c); }
        if (op.chain == first_c && op.chain_pos == 0) {
          CK(cudaEventRecord(chain_ev[2 * first_c], s0));
          for (int c : mod_chains) {
            const int br = P->ops[P->chains[c].ops[0]].stream;
            cudaStream_t sc = br == 0 ? s0 : P->side[br - 1];
            if (sc != s0) CK(cudaStreamWaitEvent(sc, chain_ev[2 * first_c], 0));
            int rc = launch_chain_op(P, c, n, 0, sc);
            if (rc) return rc;
            if (sc != s0) {
              CK(cudaEventRecord(chain_ev[2 * c + 1], sc));
              CK(cudaStreamWaitEvent(s0, chain_ev[2 * c + 1], 0));
            }
          }
          CK(cudaEventRecord(chain_ev[2 * first_c + 1], s0));
        }
        CK(cudaEventRecord(ev[i + 1], s0));
        continue;
      }
      if (op.xunit >= 0) {
        // an exchange unit is one kernel: timed where its first conv stands, split by FLOPs between the members below
        if (op.xpos == 0) {
          CK(cudaEventRecord(xunit_ev[2 * op.xunit], s0));
          int rc = launch_xunit_op(P, op.xunit, n, s0);
          if (rc) return rc;
          CK(cudaEventRecord(xunit_ev[2 * op.xunit + 1], s0));
        }
        CK(cudaEventRecord(ev[i + 1], s0));
        continue;
      }
      if (op.grp >= 0) {
        int j = i;
        while (j < nops && P->ops[j].grp == op.grp) ++j;
        std::vector<double> cost;
        int rc = launch_group(P, i, j, n, s0, &cost);
        if (rc) return rc;
        for (int m = i; m < j; ++m) CK(cudaEventRecord(ev[m + 1], s0));
        spans.push_back({i, j}); span_cost.push_back(cost);
        i = j - 1;
        continue;
      }
      int rc = launch_op(P, op, n, in, nullptr, joints, idx, nullptr, s0, false, true);
      if (rc) return rc;
      CK(cudaEventRecord(ev[i + 1], s0));
    }
    CK(cudaStreamSynchronize(s0));
    if (it >= 0) {
      for (int i = 0; i < nops; ++i) { CK(cudaEventElapsedTime(&t[i][it], ev[i], ev[i + 1])); }
      {
        std::map<int, std::vector<int>> by_module;
        for (size_t c = 0; c < P->chains.size(); ++c) if (P->chains[c].enabled) by_module[P->chains[c].module].push_back((int)c);
        for (auto& kv : by_module) {
          float total = 0.f;
          CK(cudaEventElapsedTime(&total, chain_ev[2 * kv.second[0]], chain_ev[2 * kv.second[0] + 1]));
          size_t members = 0;
          for (int c : kv.second) members += P->chains[c].ops.size();
          for (int c : kv.second) for (int m : P->chains[c].ops) t[m][it] = total / (float)members;
        }
      }
      for (size_t x = 0; x < P->xunits.size(); ++x) {
        const XUnitInfo& xu = P->xunits[x];
        if (!xu.enabled) continue;
        float total = 0.f;
        CK(cudaEventElapsedTime(&total, xunit_ev[2 * x], xunit_ev[2 * x + 1]));
        // split by a simple cost model: convs by FLOPs at 500 TFLOP/s, sums by bytes at 3 TB/s
        double fsum = 0;
        std::vector<double> fl;
        std::vector<int> mem = xu.ops;
        mem.insert(mem.end(), xu.sums.begin(), xu.sums.end());
        for (int m : mem) {
          const Op& o = P->ops[m];
          const TensorInfo& to = P->tensors[o.out];
          double c;
          if (o.kind == OP_CONV) c = 2.0 * to.H * to.W * o.k * o.k * o.cin * o.cout / 500e12;
          else {
            double bytes = 2.0 * to.H * to.W * to.C;
            for (int j = 0; j < o.nsrc; ++j) { const TensorInfo& ts = P->tensors[o.src[j]]; bytes += 2.0 * ts.H * ts.W * ts.C; }
            c = bytes / 3e12;
          }
          fl.push_back(c);
          fsum += c;
        }
        for (size_t k = 0; k < mem.size(); ++k) t[mem[k]][it] = (float)(total * fl[k] / fsum);
      }
      for (size_t gi = 0; gi < spans.size(); ++gi) {   // one kernel for the whole span: split its time by estimated cost
        float total = 0.f;
        CK(cudaEventElapsedTime(&total, ev[spans[gi].first], ev[spans[gi].second]));
        double csum = 0;
        for (double c : span_cost[gi]) csum += c;
        for (int m = spans[gi].first; m < spans[gi].second; ++m)
          t[m][it] = (float)(total * span_cost[gi][m - spans[gi].first] / csum);
      }
    }
  }
  for (int i = 0; i < nops; ++i) {
    std::sort(t[i].begin(), t[i].end());
    usec_per_op[i] = t[i][iters / 2] * 1000.f;
  }
  if (P->desc.tune[HRNET_TUNE_CHAIN_DEBUG]) {
    // development aid: every chain once more with per-CTA role timers, alone on the whole GPU and on its in-forward grid
    long long* dev = nullptr;
    const int maxg = P->num_sms;
    if (cudaMalloc(&dev, (size_t)maxg * 16 * sizeof(long long)) == cudaSuccess) {
      for (size_t c = 0; c < P->chains.size(); ++c) {
        const ChainInfo& ch = P->chains[c];
        if (!ch.enabled) continue;
        for (int g : {P->num_sms, ch.grid}) {
          cudaMemset(dev, 0, (size_t)maxg * 16 * sizeof(long long));
          cudaEvent_t a, b;
          cudaEventCreate(&a); cudaEventCreate(&b);
          cudaEventRecord(a, s0);
          int rc = launch_chain_op(P, (int)c, n, g, s0, dev);
          cudaEventRecord(b, s0);
          if (rc || cudaStreamSynchronize(s0) != cudaSuccess) { cudaFree(dev); return fail(HRNET_E_CUDA, "chain debug launch failed"); }
          float ms = 0.f;
          cudaEventElapsedTime(&ms, a, b);
          std::vector<long long> h((size_t)maxg * 16);
          cudaMemcpy(h.data(), dev, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
          double avg[16] = {0};
          int act = 0;
          for (int bb = 0; bb < maxg; ++bb) if (h[(size_t)bb * 16 + 2]) { ++act; for (int k = 0; k < 16; ++k) avg[k] += (double)h[(size_t)bb * 16 + k]; }
          for (double& v : avg) v /= std::max(1, act);
          fprintf(stderr, "[chain-dbg] %s %s grid=%d (ctas seen %d) %.1f us | per CTA: tiles %.1f, cta clk %.0f, sched: ticket %.0f dep-wait %.0f "
                  "ring %.0f | producer: ring-wait %.0f slot-wait %.0f | epi(wg0) wait_acc %.0f work %.0f | mma0 wait_full %.0f wait_tmem %.0f "
                  "wait_weights / ring %.0f\n",
                  P->ops[ch.ops[0]].name.c_str(), ch.patch ? "patch" : "im2col", g, act, ms * 1000.f, avg[1], avg[2], avg[8], avg[0], avg[9],
                  avg[10], avg[11], avg[3], avg[4], avg[5], avg[6], avg[7]);
          cudaEventDestroy(a); cudaEventDestroy(b);
        }
      }
      cudaFree(dev);
    }
  }
  for (auto& e : ev) cudaEventDestroy(e);
  for (auto& e : chain_ev) if (e) cudaEventDestroy(e);
  for (auto& e : xunit_ev) if (e) cudaEventDestroy(e);
  return HRNET_OK;
}

int hrnet_forward_host(HrnetPlan* P, const float* in_h, int n, float* heatmaps_h, float* joints_h, int32_t* idx_h,
                       const float* boxes_h, void* stream) {
  if (!P) return fail(HRNET_E_INVALID, "null plan");
  if (!P->bound) return fail(HRNET_E_STATE, "hrnet_plan_bind must be called before hrnet_forward_host");
  if (n < 0 || n > P->desc.max_batch) return fail(HRNET_E_INVALID, "n out of range [0, max_batch]");
  if (n == 0) return HRNET_OK;
  if (!in_h || !joints_h) return fail(HRNET_E_INVALID, "null input / joints pointer");
  cudaStream_t s0 = (cudaStream_t)stream;
  const int J = P->desc.nof_joints;
  float* in_d = (float*)(P->abase + P->off_in_stage);
  float* joints_d = (float*)(P->abase + P->off_joints);
  int32_t* idx_d = (int32_t*)(P->abase + P->off_idx);
  float* boxes_d = boxes_h ? (float*)(P->abase + P->off_boxes) : nullptr;
  CK(cudaMemcpyAsync(in_d, in_h, (size_t)n * 3 * P->desc.height * P->desc.width * 4, cudaMemcpyHostToDevice, s0));
  if (boxes_h) CK(cudaMemcpyAsync(boxes_d, boxes_h, (size_t)n * 16, cudaMemcpyHostToDevice, s0));
  int rc = forward_impl(P, in_d, n, nullptr, joints_d, idx_d, boxes_d, stream, false, heatmaps_h != nullptr);
  if (rc) return rc;
  CK(cudaMemcpyAsync(joints_h, joints_d, (size_t)n * J * 12, cudaMemcpyDeviceToHost, s0));
  if (idx_h) CK(cudaMemcpyAsync(idx_h, idx_d, (size_t)n * J * 4, cudaMemcpyDeviceToHost, s0));
  if (heatmaps_h)
    CK(cudaMemcpyAsync(heatmaps_h, P->abase + P->tensors[P->t_heatmaps].offset, (size_t)n * J * P->Hh * P->Wh * 4,
                       cudaMemcpyDeviceToHost, s0));
  CK(cudaStreamSynchronize(s0));
  return HRNET_OK;
}


// Debug only (HRNET_TUNE_DEBUG through hrnet_debug_set_tune, single-op entry points): per-CTA role timers of the tcgen05 kernels.
struct DbgTimers {
  long long* dev = nullptr;
  int grid = 0;
  bool on() const { return dev != nullptr; }
  void begin(int g) {
    if (!(g_single_op_tune[HRNET_TUNE_DEBUG] & 1)) return;
    grid = g;
    if (cudaMalloc(&dev, (size_t)g * 32 * sizeof(long long)) != cudaSuccess) { dev = nullptr; return; }
    cudaMemset(dev, 0, (size_t)g * 32 * sizeof(long long));
  }
  void end(cudaStream_t st, const char* what, int tiles) {
    if (!dev) return;
    cudaStreamSynchronize(st);
    std::vector<long long> h((size_t)grid * 32);
    cudaMemcpy(h.data(), dev, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaFree(dev);
    double a[16] = {0};
    for (int b = 0; b < grid; ++b) for (int k = 0; k < 16; ++k) a[k] += (double)h[(size_t)b * 32 + k] / grid;
    // wall-clock (globaltimer, ns) milestones relative to the earliest CTA entry
    long long t0 = LLONG_MAX, tmax[6] = {0, 0, 0, 0, 0, 0};
    double tavg[6] = {0, 0, 0, 0, 0, 0};
    for (int b = 0; b < grid; ++b) t0 = std::min(t0, h[(size_t)b * 32 + 16]);
    for (int b = 0; b < grid; ++b)
      for (int k = 0; k < 6; ++k) { const long long v = h[(size_t)b * 32 + 16 + k] - t0; tmax[k] = std::max(tmax[k], v); tavg[k] += (double)v / grid; }
    fprintf(stderr, "[dbg-ns] %s avg/max ns since first CTA entry: entry %.0f/%lld setup_done %.0f/%lld first_operands %.0f/%lld "
            "mma_loop_end %.0f/%lld roles_done %.0f/%lld exit %.0f/%lld\n", what, tavg[0], tmax[0], tavg[1], tmax[1], tavg[2], tmax[2],
            tavg[3], tmax[3], tavg[4], tmax[4], tavg[5], tmax[5]);
    fprintf(stderr, "[dbg] %s grid=%d tiles=%d (%.2f/CTA) cycles/CTA: producer0 wait_empty=%.0f issue=%.0f total=%.0f | "
            "producer1 wait_empty=%.0f issue=%.0f | mma wait_full=%.0f wait_tmem=%.0f issue=%.0f total=%.0f | "
            "epilogue(wg0) wait_acc=%.0f work=%.0f total=%.0f\n",
            what, grid, tiles, (double)tiles / grid, a[0], a[1], a[2], a[11], a[12], a[4], a[5], a[6], a[7], a[8], a[9], a[10]);
  }
};

// ---- single-op entry points ---------------------------------------------------------------------
static int conv_single(const void* in, const void* w, const float* scale, const float* bias, const void* residual,
                       void* out, int n, int ih, int iw, int cin, int cout, int ksize, int stride, int relu,
                       int out_f32, int use_tc, cudaStream_t st) {
  if (!in || !w || !scale || !bias || !out) return fail(HRNET_E_INVALID, "null argument");
  if (!(ksize == 1 || ksize == 3) || !(stride == 1 || stride == 2)) return fail(HRNET_E_INVALID, "ksize in {1,3}, stride in {1,2}");
  if (cin % 8 || cout % 8) return fail(HRNET_E_INVALID, "cin and cout must be multiples of 8");
  if (ih % stride || iw % stride) return fail(HRNET_E_INVALID, "spatial size must be divisible by the stride");
  Op op;
  op.kind = OP_CONV; op.cin = cin; op.cout = cout; op.k = ksize; op.stride = stride; op.pad = ksize / 2; op.relu = relu;
  const int OH = ih / stride, OW = iw / stride;
  if (use_tc == 2) {
    op.use_tc = true;
    if (!choose_patch_cfg(op, ih, iw, 0, g_single_op_tune))
      return fail(HRNET_E_INVALID, "shape not eligible for the halo-patch path (3x3 s1, cin/cout % 16, weights must fit in smem, map must tile 8x16)");
    int rc = load_driver_fns();
    if (rc) return rc;
    rc = encode_patch_maps(op, in, w, n);
    if (rc) return rc;
    choose_epi(op, out_f32 != 0, false, residual != nullptr, g_single_op_tune);
    rc = encode_epi_maps(op, out, residual, n, OH, OW);
    if (rc) return rc;
    CK(conv_patch_set_attributes(kMaxDynSmem));
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    ConvPatchParams p = op.pp;
    p.N = n; p.total_tiles = n * p.tiles_w * p.tiles_h; p.relu = relu; p.out_f32 = out_f32;
    p.scale = scale; p.bias = bias; p.residual = (const __half*)residual; p.out = out;
    if (p.total_tiles == 0) return HRNET_OK;
    p.pdl = g_single_op_tune[HRNET_TUNE_NO_PDL] ? 0 : 1;
    if (g_single_op_tune[HRNET_TUNE_DEBUG] & 8) p.H = 0;   // experiments: every output row invalid -> epilogue without global traffic
    if (g_single_op_tune[HRNET_TUNE_GRID_CAP] > 0) sms = std::max(1, std::min(sms, (int)g_single_op_tune[HRNET_TUNE_GRID_CAP]));   // experiments
    const int pgrid = conv_patch_grid(p, sms);
    DbgTimers dt; dt.begin(pgrid); p.dbg = dt.dev;
    CK(launch_conv_patch(op.tmPA, op.tmPB, op.tmOR, p, op.patch_smem, pgrid, st));
    dt.end(st, "patch", p.total_tiles);
    return HRNET_OK;
  }
  if (use_tc) {
    choose_tc_cfg(op, 0, g_single_op_tune);
    if (!op.use_tc) return fail(HRNET_E_INVALID, "shape not supported by the tcgen05 path (cin, cout must be multiples of 16)");
    int rc = load_driver_fns();
    if (rc) return rc;
    choose_epi(op, out_f32 != 0, false, residual != nullptr, g_single_op_tune);
    rc = encode_epi_maps(op, out, residual, n, OH, OW);
    if (rc) return rc;
    rc = encode_im2col(&op.tmA, in, n, ih, iw, cin, op.tc.kc, ksize, stride, op.pad, op.pad, op.pad, op.pad);
    if (rc) return rc;
    rc = encode_weights(&op.tmB, w, cout, ksize * ksize * cin, op.tc.kc, op.tc.n_tile / op.tc.cs);
    if (rc) return rc;
    CK(conv_tc_set_attributes(kMaxDynSmem));
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (g_single_op_tune[HRNET_TUNE_GRID_CAP] > 0) sms = std::max(2, std::min(sms, (int)g_single_op_tune[HRNET_TUNE_GRID_CAP]));   // experiments
    ConvTcParams p{};
    p.M_total = n * OH * OW; p.OH = OH; p.OW = OW; p.OHW = OH * OW;
    p.ksize = ksize; p.stride = stride; p.pad_h = op.pad; p.pad_w = op.pad; p.Cin = cin; p.Cout = cout;
    p.kc = op.tc.kc; p.cpt = (cin + op.tc.kc - 1) / op.tc.kc; p.nkb = ksize * ksize * p.cpt; p.bps = op.tc.bps;
    p.n_tile = op.tc.n_tile; p.n_tiles = cout / op.tc.n_tile; p.m_tiles = (p.M_total + 127) / 128;
    p.cs = op.tc.cs;
    p.stages = op.tc.stages; p.mma_warps = op.tc.mma_warps; p.relu = relu; p.out_f32 = out_f32; p.tmem_cols = op.tc.tmem_cols;
    p.a_blk_bytes = (int)align_up((size_t)128 * p.kc * 2, 1024);
    p.b_blk_bytes = (int)align_up((size_t)(p.n_tile / p.cs) * p.kc * 2, 1024);
    p.scale = scale; p.bias = bias; p.residual = (const __half*)residual; p.out = out;
    p.epi_tma = op.tc.epi; p.epi_bytes = op.tc.epi_bytes;
    p.pdl = g_single_op_tune[HRNET_TUNE_NO_PDL] ? 0 : 1;
    const int tiles = p.m_tiles * p.n_tiles;
    if (tiles == 0) return HRNET_OK;
    DbgTimers dt; dt.begin(conv_tc_grid(p, op.tc.smem_bytes, sms)); p.dbg = dt.dev;
    CK(launch_conv_tc(op.tmA, op.tmB, op.tmOR, p, op.tc.smem_bytes, conv_tc_grid(p, op.tc.smem_bytes, sms), st));
    dt.end(st, "im2col", tiles);
  } else {
    ConvSimtParams p{};
    p.N = n; p.IH = ih; p.IW = iw; p.OH = OH; p.OW = OW; p.Cin = cin; p.Cout = cout; p.ksize = ksize; p.stride = stride;
    p.pad = op.pad; p.relu = relu; p.out_f32 = out_f32;
    p.in = (const __half*)in; p.w = (const __half*)w; p.scale = scale; p.bias = bias;
    p.residual = (const __half*)residual; p.out = out;
    CK(launch_conv_simt(p, st));
  }
  return HRNET_OK;
}

int hrnet_conv_bn_act(const void* in, const void* w, const float* scale, const float* bias, const void* residual,
                      void* out, int n, int ih, int iw, int cin, int cout, int ksize, int stride, int relu, int out_f32,
                      int use_tc, void* stream) {
  return conv_single(in, w, scale, bias, residual, out, n, ih, iw, cin, cout, ksize, stride, relu, out_f32, use_tc,
                     (cudaStream_t)stream);
}

int hrnet_conv_bench(const void* in, const void* w, const float* scale, const float* bias, const void* residual,
                     void* out, int n, int ih, int iw, int cin, int cout, int ksize, int stride, int relu, int use_tc,
                     int iters, float* usec_out, void* stream) {
  if (!usec_out || iters <= 0) return fail(HRNET_E_INVALID, "bad iters / output");
  cudaStream_t st = (cudaStream_t)stream;
  std::vector<cudaEvent_t> ev(iters + 1);
  for (auto& e : ev) CK(cudaEventCreate(&e));
  for (int i = 0; i < 3; ++i) {
    int rc = conv_single(in, w, scale, bias, residual, out, n, ih, iw, cin, cout, ksize, stride, relu, 0, use_tc, st);
    if (rc) return rc;
  }
  CK(cudaStreamSynchronize(st));
  CK(cudaEventRecord(ev[0], st));
  for (int i = 0; i < iters; ++i) {
    int rc = conv_single(in, w, scale, bias, residual, out, n, ih, iw, cin, cout, ksize, stride, relu, 0, use_tc, st);
    if (rc) return rc;
    CK(cudaEventRecord(ev[i + 1], st));
  }
  CK(cudaStreamSynchronize(st));
  std::vector<float> t(iters);
  for (int i = 0; i < iters; ++i) { CK(cudaEventElapsedTime(&t[i], ev[i], ev[i + 1])); t[i] *= 1000.f; }
  std::sort(t.begin(), t.end());
  *usec_out = t[iters / 2];
  for (auto& e : ev) cudaEventDestroy(e);
  return HRNET_OK;
}

int hrnet_fuse(const void* const* srcs, const int* shifts, const int* is_f32, int nsrc, void* out, int n, int h, int w,
               int c, int relu, void* stream) {
  if (!srcs || !shifts || !is_f32 || !out || nsrc < 1 || nsrc > 4) return fail(HRNET_E_INVALID, "bad argument (1 <= nsrc <= 4)");
  if (c % 8) return fail(HRNET_E_INVALID, "c must be a multiple of 8");
  FuseParams p{};
  p.N = n; p.H = h; p.W = w; p.C = c; p.nsrc = nsrc; p.relu = relu;
  for (int j = 0; j < nsrc; ++j) {
    if ((h % (1 << shifts[j])) || (w % (1 << shifts[j]))) return fail(HRNET_E_INVALID, "map size not divisible by the upsample factor");
    p.src[j] = srcs[j]; p.shift[j] = shifts[j]; p.f32[j] = is_f32[j];
  }
  p.out = (__half*)out;
  CK(launch_fuse(p, (cudaStream_t)stream));
  return HRNET_OK;
}

int hrnet_argmax(const float* heatmaps, int n, int J, int hh, int wh, const float* boxes, float* joints,
                 int32_t* argmax_idx, void* stream) {
  if (n < 0 || J <= 0 || hh <= 0 || wh <= 0) return fail(HRNET_E_INVALID, "bad shape");
  if (n == 0) return HRNET_OK;
  if (!heatmaps || !joints) return fail(HRNET_E_INVALID, "null argument");
  CK(launch_argmax(heatmaps, n, J, hh, wh, boxes, joints, argmax_idx, (cudaStream_t)stream));
  return HRNET_OK;
}

int hrnet_resize_cubic_u8(const uint8_t* src, int n, int sh, int sw, uint8_t* dst, int dh, int dw, const int32_t* xofs,
                          const int16_t* xcoef, const int32_t* yofs, const int16_t* ycoef, void* stream) {
  if (n < 0 || sh <= 0 || sw <= 0 || dh <= 0 || dw <= 0) return fail(HRNET_E_INVALID, "bad shape");
  if (n == 0) return HRNET_OK;
  if (!src || !dst || !xofs || !xcoef || !yofs || !ycoef) return fail(HRNET_E_INVALID, "null argument");
  CK(launch_resize_cubic_u8(src, dst, n, sh, sw, dh, dw, xofs, xcoef, yofs, ycoef, (cudaStream_t)stream));
  return HRNET_OK;
}

int hrnet_crop_resize_bilinear_u8(const uint8_t* frames, int n_frames, int frame_h, int frame_w, const int32_t* crops,
                                  const int32_t* tables, int m, uint8_t* out, int out_h, int out_w, void* stream) {
  if (n_frames <= 0 || frame_h <= 0 || frame_w <= 0 || m < 0 || out_h <= 0 || out_w <= 0) return fail(HRNET_E_INVALID, "bad shape");
  if (m == 0) return HRNET_OK;
  if (!frames || !crops || !tables || !out) return fail(HRNET_E_INVALID, "null argument");
  CK(launch_crop_resize_bilinear_u8(frames, frame_h, frame_w, crops, tables, m, out, out_h, out_w, (cudaStream_t)stream));
  return HRNET_OK;
}

int hrnet_final_preds(const float* heatmaps, int n, int J, int hh, int wh, int post_processing, const double* trans_2x3,
                      float* preds_xy, float* maxvals, void* stream) {
  if (n < 0 || J <= 0 || hh <= 0 || wh <= 0) return fail(HRNET_E_INVALID, "bad shape");
  if (n == 0) return HRNET_OK;
  if (!heatmaps || !preds_xy || !maxvals) return fail(HRNET_E_INVALID, "null argument");
  CK(launch_final_preds(heatmaps, n, J, hh, wh, post_processing ? 1 : 0, trans_2x3, preds_xy, maxvals, (cudaStream_t)stream));
  return HRNET_OK;
}

int hrnet_flip_average(const float* output, const float* output_flipped, const int32_t* joint_perm_host, int n, int J, int hh,
                       int wh, float* averaged, void* stream) {
  if (n < 0 || J <= 0 || J > 32 || hh <= 0 || wh <= 0) return fail(HRNET_E_INVALID, "bad shape (nof_joints <= 32)");
  if (n == 0) return HRNET_OK;
  if (!output || !output_flipped || !joint_perm_host || !averaged) return fail(HRNET_E_INVALID, "null argument");
  for (int j = 0; j < J; ++j)
    if (joint_perm_host[j] < 0 || joint_perm_host[j] >= J) return fail(HRNET_E_INVALID, "joint_perm entry out of range");
  CK(launch_flip_average(output, output_flipped, averaged, joint_perm_host, n, J, hh, wh, (cudaStream_t)stream));
  return HRNET_OK;
}

}  // extern "C"
