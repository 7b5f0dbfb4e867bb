/*
 * hrnet_b200 -- C ABI of the B200-native HRNet inference hot path.
 *
 * Drop-in boundary: the reference calls its network exactly once per
 * <= max_batch_size chunk as `self.model(images)` (SimpleHRNet.py:286,294,421,429), copies the
 * heat-maps to the host (:296,:431) and decodes them in a Python double loop (:299-308,:434-443).
 * The reference already swaps `self.model` for an opaque engine at this seam when
 * `enable_tensorrt=True` (SimpleHRNet.py:143-147).  This library is the engine that plugs into
 * the same seam: plain pointers and sizes, no torch / C++ types in any signature, bound from
 * Python with ctypes (INTEGRATION.md shows the reference-side stub).
 *
 * Conventions
 *   - every entry point returns 0 on success or a negative HRNET_E_* code; the message of the
 *     last failure on the calling thread is available through hrnet_last_error().
 *   - the library never allocates device memory and never synchronises unless stated: the caller
 *     (PyTorch in this repo) owns the packed weights, the activation workspace and every
 *     input / output buffer; pointers are BORROWED and must stay valid while bound.
 *   - all device work is enqueued on the caller's stream (`stream` is a cudaStream_t passed as
 *     void*); branch-level concurrency uses internal streams forked from / joined to it.
 *   - a plan is not thread-safe; use one plan per (process, device).
 */
#ifndef HRNET_B200_H_
#define HRNET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HRNET_OK 0
#define HRNET_E_INVALID (-1)   /* bad argument / unsupported configuration */
#define HRNET_E_CUDA (-2)      /* a CUDA runtime / driver call failed */
#define HRNET_E_STATE (-3)     /* call order violated (e.g. forward before bind) */
#define HRNET_E_NOMEM (-4)     /* caller-provided buffer too small */

#define HRNET_ARCH_HRNET 0      /* models_/hrnet.py:74  HRNet(c, nof_joints)            */
#define HRNET_ARCH_POSERESNET 1 /* models_/poseresnet.py:16 PoseResNet(resnet_size, ..) */

/* plan flags */
#define HRNET_FLAG_FORCE_SIMT 1u   /* debug: run every conv on the SIMT cross-check kernel          */
#define HRNET_FLAG_NO_GRAPH 2u     /* debug: launch kernels directly instead of replaying a graph   */
#define HRNET_FLAG_FUSE_F32 4u     /* store exchange-unit partial terms in fp32 (default fp16: same accuracy, faster) */
#define HRNET_FLAG_SERIAL 8u       /* debug: run all branches on the caller's stream                */
#define HRNET_FLAG_NO_PATCH 16u    /* debug: disable the halo-patch 3x3 kernel (im2col kernel everywhere) */
#define HRNET_FLAG_GROUP 64u       /* experiment: issue the k-th convs of all branches as one multi-problem kernel (same speed) */
#define HRNET_FLAG_PARTITION 32u   /* experiment: split the SMs between the concurrent branch chains (slower) */
#define HRNET_FLAG_NO_CHAIN 128u   /* debug / cross-check: launch the eight convs of a StageModule branch one by one
                                      instead of as one persistent branch-chain kernel (bit-identical results)     */


/* Tuning / diagnostic knobs of a plan (HrnetDesc.tune[]); 0 = the library's default.  They select between kernel
 * variants that compute the same function (results agree to accumulation-order noise at most; the default set is what
 * the parity tests pin).  Nothing in the library reads the process environment. */
#define HRNET_TUNE_CHAIN_SHARE0 0  /* [0..3] per-mille share of the SMs given to the branch-0..3 chain kernel of a
                                      StageModule with four branches (three / two branches: renormalised)          */
#define HRNET_TUNE_CHAIN_GRID_CAP 4 /* upper bound on any chain kernel's grid (experiments)                         */
#define HRNET_TUNE_CHAIN_DEBUG 5    /* != 0: hrnet_profile_ops also runs every chain once with per-CTA role timers and
                                      prints their averages to stderr (development aid; allocates a scratch buffer)  */
#define HRNET_TUNE_IGEMM_PAIR 6     /* 2: im2col kernel in CTA-pair mode (tcgen05 cta_group::2)                      */
#define HRNET_TUNE_IGEMM_PAIR_MIN_K 7 /* ... only for convs with at least this GEMM K                                */
#define HRNET_TUNE_PATCH_PAIR_MIN_COUT 8  /* > 0: halo-patch kernel in CTA-pair mode for convs with Cout in            */
#define HRNET_TUNE_PATCH_PAIR_MAX_COUT 9  /*      [MIN_COUT, MAX_COUT] (MAX 0 = no upper bound)                        */
#define HRNET_TUNE_EPILOGUE 10      /* 0 auto, 1 direct, 2 staged TMA wherever eligible, 3 warp-staged coalesced wherever
                                      it fits, 4 TMA on every halo-patch conv, 5 TMA on every im2col conv, 6 batched
                                      TMEM loads on tiles <= 64 channels                                              */
#define HRNET_TUNE_BPS 11           /* k-blocks per pipeline stage of the im2col kernel: 1 or 2                       */
#define HRNET_TUNE_IGEMM_MMA2 12    /* two MMA-issuing warps for N tiles up to this width (default 96); -1 never, 1 always */
#define HRNET_TUNE_PATCH_MMA2 13    /* halo-patch kernel: 1 one MMA issuer, 2 two issuers (default: two when each keeps two slots) */
#define HRNET_TUNE_PATCH_NACC 14    /* 2: only two TMEM accumulator buffers in the halo-patch kernel                  */
#define HRNET_TUNE_NO_PDL 15        /* != 0: launch without programmatic dependent launch                             */
#define HRNET_TUNE_DEBUG 16         /* bit 0: per-CTA role timers on the single-op entry points (stderr); bit 1: per-op
                                      output checksums with HRNET_FLAG_NO_GRAPH (stderr); bit 2: finish times of grouped
                                      launches; bit 3: single-op halo-patch conv without global stores (timing only)  */
#define HRNET_TUNE_GRID_CAP 17      /* single-op entry points: cap on the persistent grid (experiments)              */
#define HRNET_TUNE_CHAIN_M2 18       /* im2col chains: 2 = two M-tiles per ticket sharing every weight k-block (30 % fewer TMA
                                      bytes, no accumulator double buffering); default one tile per ticket             */
#define HRNET_TUNE_XUNIT 19          /* exchange-unit kernel (conv_xunit.cu: several fuse-layer convs as one ticket-scheduled kernel).
                                      0 / 1 = one unit per StageModule and SOURCE branch, launched behind that branch's chain
                                      (default); 2 = one unit per module that also executes the sums as sum tickets (measured
                                      slower); 3 = off: every fuse-layer conv its own launch.  Bit-identical results.          */
#define HRNET_TUNE_CHAIN_PAIR 20     /* im2col chains on CTA pairs (tcgen05 cta_group::2: a unit = two M-tiles, each CTA stages half of
                                      the weight tile): 2 = on                                                          */
#define HRNET_TUNE_CHAIN_SKIP 21     /* experiments only, RESULTS INVALID: the im2col chain's epilogue does 1 = nothing, 2 = only its
                                      accumulator loads, 3 = only its residual loads and stores (what slows the MMA stream?) */
#define HRNET_TUNE_CHAIN_STAGES 22   /* im2col chains: pipeline stages (default: as many as fit in shared memory, <= 8)                 */
#define HRNET_TUNE_CHAIN_EARLY 23    /* 1 = a module's branch chains do not wait for each other's inputs (default: they start together) */
#define HRNET_TUNE_COUNT 24

typedef struct HrnetPlan HrnetPlan;

/* Mirrors the constructor arguments that select the network in SimpleHRNet.__init__
 * (SimpleHRNet.py:21-37: c, nof_joints, model_name, resolution, max_batch_size). */
typedef struct HrnetDesc {
  int32_t arch;        /* HRNET_ARCH_*                                              */
  int32_t c;           /* HRNet width (32 / 48) or PoseResNet size (50 / 101 / 152) */
  int32_t nof_joints;  /* <= 32                                                     */
  int32_t height;      /* network input height, multiple of 32                      */
  int32_t width;       /* network input width, multiple of 32                       */
  int32_t max_batch;   /* largest n accepted by hrnet_forward                       */
  uint32_t flags;      /* HRNET_FLAG_*                                              */
  int32_t tune[HRNET_TUNE_COUNT]; /* HRNET_TUNE_*; zero-initialise for the defaults        */
} HrnetDesc;

/* One weight-carrying layer of the plan; tells the host-side packer which state_dict entries
 * (reference naming) go where in the packed weight buffer. */
typedef struct HrnetParamInfo {
  char conv_key[96];    /* "<key>.weight" is the conv weight, "<key>.bias" its bias if has_bias   */
  char bn_key[96];      /* "<key>.{weight,bias,running_mean,running_var}"; empty = no BatchNorm   */
  int32_t cout, cin, kh, kw;
  int32_t kind;         /* 0 conv [cout,cin,kh,kw]; 1 transposed-conv sub-pixel phase (see sub_a/b) */
  int32_t sub_a, sub_b; /* kind 1: output parity (row, col) this 2x2 sub-kernel produces          */
  int32_t has_bias;
  int32_t w_f32;        /* weights stored fp32 instead of fp16                                     */
  uint64_t w_offset;    /* bytes: packed as [cout][kh][kw][cin]                                    */
  uint64_t scale_offset;/* bytes: fp32 [cout]  gamma / sqrt(var + 1e-5)      (1 without BN)        */
  uint64_t bias_offset; /* bytes: fp32 [cout]  beta - mean * scale           (conv bias without BN)*/
} HrnetParamInfo;

/* ---- plan construction: host only, no CUDA calls (usable on a CPU-only box) ------------------ */
int hrnet_plan_create(const HrnetDesc* desc, HrnetPlan** out);
void hrnet_plan_destroy(HrnetPlan* plan);
const char* hrnet_last_error(void);

/* Bytes the caller must provide: activation workspace and packed weights (both 1024 B aligned). */
int hrnet_plan_workspace_bytes(const HrnetPlan* plan, size_t* act_bytes, size_t* weight_bytes);
int hrnet_plan_num_params(const HrnetPlan* plan);
int hrnet_plan_param_info(const HrnetPlan* plan, int index, HrnetParamInfo* out);
/* JSON description of tensors / ops (used by the CPU graph-emulation test and for debugging).
 * Writes at most cap bytes (NUL-terminated); *needed receives the full length incl. NUL. */
int hrnet_plan_describe(const HrnetPlan* plan, char* buf, size_t cap, size_t* needed);

/* ---- binding and execution: need a CUDA device ---------------------------------------------- */
/* Creates the TMA descriptors over the caller's buffers.  Must be repeated if either moves -- or if the weights are
 * changed in place: `weights_dev` must already hold the packed weights, because the BN constants of the chained convs
 * and the head's weights are read back here (synchronous device-to-host copies) and later passed to the kernels as
 * launch parameters (constant bank). */
int hrnet_plan_bind(HrnetPlan* plan, void* weights_dev, size_t weight_bytes, void* workspace_dev, size_t act_bytes);

/* The hot path: `self.model(images)` + the argmax decode.
 *   in_nchw_f32 : [n,3,H,W] fp32, ImageNet-normalised RGB, device (what the reference feeds self.model)
 *   heatmaps    : [n,J,H/4,W/4] fp32 device, or NULL: the head then runs fused with the argmax and NO heat-maps are written
 *                 (same joints / indices bit for bit)
 *   joints      : [n,J,3] fp32 device, (y, x, confidence) exactly as SimpleHRNet.py:306-308
 *   argmax_idx  : [n,J] int32 device flat np.argmax index, or NULL
 *   boxes       : [n,4] fp32 device (x1,y1,x2,y2), or NULL = [0,0,W,H] (multiperson=False crops)
 * 0 <= n <= max_batch. Asynchronous on `stream`. */
int hrnet_forward(HrnetPlan* plan, const float* in_nchw_f32, int n, float* heatmaps, float* joints,
                  int32_t* argmax_idx, const float* boxes, void* stream);

/* Same call with HOST buffers (pinned for full speed): H2D of the input, forward, D2H of the
 * joints (and of heat-maps / indices when requested).  Staging buffers live in the workspace.
 * Synchronises `stream` before returning. */
int hrnet_forward_host(HrnetPlan* plan, const float* in_nchw_f32_host, int n, float* heatmaps_host,
                       float* joints_host, int32_t* argmax_idx_host, const float* boxes_host, void* stream);

/* Same two calls fed with the reference's pre-transform images: [n,H,W,3] BGR uint8 at network resolution (what
 * SimpleHRNet hands to cv2.cvtColor + ToTensor + Normalize, SimpleHRNet.py:222,149-153).  The stem kernel applies
 * BGR->RGB, /255 and the ImageNet mean/std with the same fp32 operation order, so results are bit-identical to the
 * fp32 entry points while the host->device copy is 4x smaller.  HRNet only. */
int hrnet_forward_u8(HrnetPlan* plan, const uint8_t* images_nhwc_bgr_u8, int n, float* heatmaps, float* joints,
                     int32_t* argmax_idx, const float* boxes, void* stream);
int hrnet_forward_host_u8(HrnetPlan* plan, const uint8_t* images_nhwc_bgr_u8_host, int n, float* heatmaps_host,
                          float* joints_host, int32_t* argmax_idx_host, const float* boxes_host, void* stream);
/* Same, but only ENQUEUES the copies and the forward on `stream` and returns: the host buffers must be pinned and stay
 * untouched until the caller synchronises `stream`.  With two plans bound to the same weight buffer (own workspaces) on
 * two streams this double-buffers a serving loop: batch i+1 is copied to the device while batch i computes (replaces the
 * reference's synchronous images.to(device) -> model -> .cpu() sequence, SimpleHRNet.py:282-296). */
int hrnet_forward_host_u8_async(HrnetPlan* plan, const uint8_t* images_nhwc_bgr_u8_host, int n, float* heatmaps_host,
                                float* joints_host, int32_t* argmax_idx_host, const float* boxes_host, void* stream);

/* Per-op device time: runs the plan's ops one after the other on `stream` (no graph, no branch concurrency) with a
 * CUDA event pair around every launch; usec_per_op[i] = median over `iters` passes for op i of hrnet_plan_describe.
 * Used by bench.py for the per-layer-class breakdown.  Synchronises. */
int hrnet_profile_ops(HrnetPlan* plan, const float* in_nchw_f32, int n, float* usec_per_op, int iters, void* stream);

/* Number of kernels one hrnet_forward(n) launches (for bench.py's gpu_launches). */
int hrnet_plan_launch_count(const HrnetPlan* plan);

/* ---- single-op entry points (tests, ncu) ------------------------------------------------------ */
/* The single-op entry points below have no HrnetDesc: they use this process-wide copy of the tuning knobs
 * (HRNET_TUNE_COUNT int32 values, NULL = all defaults).  Plans never look at it. */
void hrnet_debug_set_tune(const int32_t* tune);
/* kxk conv + BN(scale,bias) (+residual) (+ReLU) on NHWC fp16; weights [cout][k][k][cin] fp16.
 * use_tc: 0 = SIMT cross-check kernel, 1 = tcgen05 implicit GEMM (TMA im2col), 2 = tcgen05 halo-patch
 * kernel (3x3 stride 1 with shared-memory-resident weights).  out_f32 selects fp32 output. */
int hrnet_conv_bn_act(const void* in_nhwc_f16, const void* w_f16, const float* scale, const float* bias,
                      const void* residual_f16, void* out, int n, int ih, int iw, int cin, int cout, int ksize,
                      int stride, int relu, int out_f32, int use_tc, void* stream);
/* out = act(sum_j nearest_up(src_j, 2^shift_j)); srcs NHWC fp16 (is_f32[j]=0) or fp32. */
int hrnet_fuse(const void* const* srcs, const int* shifts, const int* is_f32, int nsrc, void* out_f16, int n, int h,
               int w, int c, int relu, void* stream);
/* argmax decode of [n,J,Hh,Wh] fp32 heat-maps (SimpleHRNet.py:296-308). */
int hrnet_argmax(const float* heatmaps, int n, int nof_joints, int hh, int wh, const float* boxes, float* joints,
                 int32_t* argmax_idx, void* stream);
/* Evaluation-side decode, misc/utils.py:125-182 (get_max_preds + get_final_preds; training/COCO.py:218-224 is the caller):
 * per joint the first maximum of the map (torch.max), (x, y) = (idx % wh, idx / wh) zeroed when the maximum is not > 0,
 * with post_processing the quarter-pixel shift towards the higher neighbour, then -- when trans_2x3 is given -- the
 * inverse affine `np.dot(t, [x, y, 1.])` of transform_preds in float64 (t = the host-built cv2.getAffineTransform
 * matrix of misc/utils.py:46-79, one [2][3] float64 block per person, device memory).
 *   preds_xy [n,J,2] fp32 (x, y)   maxvals [n,J] fp32 (the [n,J,1] tensor of the reference) */
int hrnet_final_preds(const float* heatmaps, int n, int nof_joints, int hh, int wh, int post_processing,
                      const double* trans_2x3, float* preds_xy, float* maxvals, void* stream);
/* Flip test, training/COCO.py:206-212 with misc/utils.py:19-29: averaged = (output + flip_back(output_flipped)) * 0.5 on
 * [n,J,hh,wh] fp32 device maps.  joint_perm_host[j] (HOST array of J ints, J <= 32) = the joint whose flipped map becomes
 * joint j (identity outside the left/right pairs).  averaged may alias output. */
int hrnet_flip_average(const float* output, const float* output_flipped, const int32_t* joint_perm_host, int n,
                       int nof_joints, int hh, int wh, float* averaged, void* stream);
/* Device-side image resize of the reference's preprocessing (`cv2.resize(image, (W, H), interpolation=cv2.INTER_CUBIC)`,
 * SimpleHRNet.py:216-220 and :356-360): OpenCV's own 8-bit cubic kernel (4 taps per axis, Keys weights A = -0.75 in
 * 11-bit fixed point, int32 accumulation, rounding shift by 22, saturation).  Bit-identical to cv2.resize with OpenCV's
 * vendor-optimised paths disabled (cv2.setUseOptimized(False)); the IPP path most cv2 wheels take by default differs
 * from OpenCV's own kernel by one grey level on ~3.5 % of the pixels.
 *   src [n,sh,sw,3] uint8, dst [n,dh,dw,3] uint8 (device); per axis the first tap index sx-1 (xofs [dw], yofs [dh]) and the
 *   four 11-bit weights (xcoef [dw][4], ycoef [dh][4], int16), device arrays built by the host
 *   (simple_hrnet_b200.preprocess.cubic_tables restates OpenCV's float32 table computation). */
int hrnet_resize_cubic_u8(const uint8_t* src, int n, int sh, int sw, uint8_t* dst, int dh, int dw, const int32_t* xofs,
                          const int16_t* xcoef, const int32_t* yofs, const int16_t* ycoef, void* stream);
/* Multi-person crops on the device: `ToPILImage -> Resize((H, W))` of the reference's crop transform
 * (SimpleHRNet.py:166-171) applied to `image[y1:y2, x1:x2]` (+ the zero padding of SimpleHRNet.py:262-270) -- Pillow's
 * antialiased bilinear resample of an 8-bit image restated bit for bit (horizontal pass rounded to uint8, vertical pass,
 * 22-bit fixed-point coefficients).  The output is the input of hrnet_forward_u8.
 *   frames [n_frames, frame_h, frame_w, 3] uint8 (device); out [m, out_h, out_w, 3] uint8 (device)
 *   crops  [m][12] int32 (device): frame | x0 y0 = frame coordinates of the padded crop's top-left | vx0 vy0 vx1 vy1 = the
 *          frame rectangle that holds image data (zero outside) | offsets (in int32 units) of the crop's x and y tables in
 *          `tables` | taps per output coordinate kx ky | 0
 *   tables int32 (device): per axis [out][2] (first source index, tap count) then [out][k] coefficients, as built by
 *          simple_hrnet_b200.preprocess.pil_bilinear_tables (= Pillow's precompute_coeffs + normalize_coeffs_8bpc). */
int hrnet_crop_resize_bilinear_u8(const uint8_t* frames, int n_frames, int frame_h, int frame_w, const int32_t* crops,
                                  const int32_t* tables, int m, uint8_t* out, int out_h, int out_w, void* stream);
/* median device time in microseconds of `iters` launches of one conv (CUDA events on `stream`),
 * same arguments as hrnet_conv_bn_act; used by bench.py for the per-kernel roofline. */
int hrnet_conv_bench(const void* in_nhwc_f16, const void* w_f16, const float* scale, const float* bias,
                     const void* residual_f16, void* out, int n, int ih, int iw, int cin, int cout, int ksize,
                     int stride, int relu, int use_tc, int iters, float* usec_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HRNET_B200_H_ */
