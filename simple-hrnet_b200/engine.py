"""B200Engine: the object that replaces `self.model` at the reference's engine seam
(SimpleHRNet.py:143-147, where a torch2trt TRTModule is swapped in) and additionally exposes the
fused forward+decode entry point.  PyTorch is used only to own device memory and streams."""
import ctypes
import json
import math
from collections import OrderedDict

import numpy as np
import torch

from . import _lib
from ._lib import HrnetDesc, HrnetParamInfo, HrnetError, check, load_library

BN_EPS = 1e-5  # models_/hrnet.py:33,42,49,80; nn.BatchNorm2d default in models_/modules.py


def _param_infos(lib, plan):
    out = []
    for i in range(check(lib.hrnet_plan_num_params(plan), lib)):
        pi = HrnetParamInfo()
        check(lib.hrnet_plan_param_info(plan, i, ctypes.byref(pi)), lib)
        out.append(pi)
    return out


def expected_state_dict_keys(infos):
    """Keys the reference's strict load_state_dict (SimpleHRNet.py:117-121) would require."""
    req, opt = OrderedDict(), set()
    for pi in infos:
        ck, bk = pi.conv_key.decode(), pi.bn_key.decode()
        req[ck + ".weight"] = None
        if pi.has_bias:
            req[ck + ".bias"] = None
        if bk:
            for s in ("weight", "bias", "running_mean", "running_var"):
                req[bk + "." + s] = None
            opt.add(bk + ".num_batches_tracked")
    return list(req.keys()), opt


def pack_state_dict(state_dict, infos, weight_bytes):
    """state_dict (reference names) -> packed host buffer laid out by the plan:
    per layer fp16 (or fp32) weights as [cout][kh][kw][cin] plus fp32 BN-folded (scale, bias).
    BN stays an fp32 epilogue: scale = gamma / sqrt(var + eps), bias = beta - mean * scale are NOT
    folded into the fp16 weights (the extra rounding would eat the 1e-3 heat-map margin)."""
    if "model" in state_dict and not torch.is_tensor(state_dict["model"]):
        state_dict = state_dict["model"]          # checkpoint saved by training (SimpleHRNet.py:118-119)
    req, opt = expected_state_dict_keys(infos)
    missing = [k for k in req if k not in state_dict]
    unexpected = [k for k in state_dict if k not in set(req) and k not in opt]
    if missing or unexpected:
        raise RuntimeError("Error(s) in loading state_dict: "
                           + (f"Missing key(s) in state_dict: {missing[:8]}{'...' if len(missing) > 8 else ''}. " if missing else "")
                           + (f"Unexpected key(s) in state_dict: {unexpected[:8]}{'...' if len(unexpected) > 8 else ''}." if unexpected else ""))
    buf = np.zeros(weight_bytes, dtype=np.uint8)
    for pi in infos:
        ck, bk = pi.conv_key.decode(), pi.bn_key.decode()
        w = state_dict[ck + ".weight"].detach().to("cpu", torch.float32)
        if pi.kind == 0:
            if tuple(w.shape) != (pi.cout, pi.cin, pi.kh, pi.kw):
                raise RuntimeError(f"size mismatch for {ck}.weight: {tuple(w.shape)} vs {(pi.cout, pi.cin, pi.kh, pi.kw)}")
            wp = w.permute(0, 2, 3, 1).contiguous()
        else:
            # ConvTranspose2d(4, stride 2, pad 1) weight [cin, cout, 4, 4] -> the 2x2 sub-kernel that
            # produces output pixels (2i+a, 2j+b): tap r reads input row i - (1-a) + r with kh = (3-a) - 2r
            if tuple(w.shape) != (pi.cin, pi.cout, 4, 4):
                raise RuntimeError(f"size mismatch for {ck}.weight: {tuple(w.shape)}")
            a, b = pi.sub_a, pi.sub_b
            khs = [(3 - a) - 2 * r for r in range(2)]
            kws = [(3 - b) - 2 * s for s in range(2)]
            sub = w[:, :, khs][:, :, :, kws]             # [cin, cout, 2, 2]
            wp = sub.permute(1, 2, 3, 0).contiguous()    # [cout, r, s, cin]
        raw = (wp.numpy().astype(np.float32) if pi.w_f32 else wp.to(torch.float16).numpy()).tobytes()
        buf[pi.w_offset:pi.w_offset + len(raw)] = np.frombuffer(raw, dtype=np.uint8)
        if bk:
            g = state_dict[bk + ".weight"].detach().to("cpu", torch.float64)
            be = state_dict[bk + ".bias"].detach().to("cpu", torch.float64)
            mu = state_dict[bk + ".running_mean"].detach().to("cpu", torch.float64)
            var = state_dict[bk + ".running_var"].detach().to("cpu", torch.float64)
            scale = g / torch.sqrt(var + BN_EPS)
            bias = be - mu * scale
        else:
            scale = torch.ones(pi.cout, dtype=torch.float64)
            bias = (state_dict[ck + ".bias"].detach().to("cpu", torch.float64) if pi.has_bias
                    else torch.zeros(pi.cout, dtype=torch.float64))
        s_raw = scale.to(torch.float32).numpy().tobytes()
        b_raw = bias.to(torch.float32).numpy().tobytes()
        buf[pi.scale_offset:pi.scale_offset + len(s_raw)] = np.frombuffer(s_raw, dtype=np.uint8)
        buf[pi.bias_offset:pi.bias_offset + len(b_raw)] = np.frombuffer(b_raw, dtype=np.uint8)
    return buf


def _aligned_empty(nbytes, device, zero=False, align=1024):
    """(base, view): a uint8 device buffer whose view starts on an `align`-byte boundary.  The torch caching allocator
    only guarantees 512-byte granularity (a block split off a cached segment can start at an odd 512 multiple), while
    hrnet_plan_bind wants 1024 (swizzle-atom aligned TMA boxes): over-allocate and bind the aligned slice; `base`
    keeps the allocation alive."""
    base = (torch.zeros if zero else torch.empty)(nbytes + align, dtype=torch.uint8, device=device)
    off = (-base.data_ptr()) % align
    return base, base[off:off + nbytes]


class Plan:
    """Host-only view of a plan (no CUDA needed): layer inventory, workspace sizes, op list."""

    def __init__(self, arch, c, nof_joints, resolution, max_batch, flags=0, tune=None):
        self.lib = load_library()
        self.arch = {"hrnet": _lib.ARCH_HRNET, "poseresnet": _lib.ARCH_POSERESNET}.get(str(arch).lower())
        if self.arch is None:
            raise ValueError('Wrong model name.')          # SimpleHRNet.py:114
        self.c, self.J = int(c), int(nof_joints)
        self.H, self.W = int(resolution[0]), int(resolution[1])
        self.max_batch = int(max_batch)
        self.flags = int(flags)
        self._plan = ctypes.c_void_p()
        self.tune = dict(tune or {})                     # {HRNET_TUNE_* index: value}; speed knobs, never results
        desc = HrnetDesc(self.arch, self.c, self.J, self.H, self.W, self.max_batch, int(flags))
        for k, v in self.tune.items():
            desc.tune[int(k)] = int(v)
        check(self.lib.hrnet_plan_create(ctypes.byref(desc), ctypes.byref(self._plan)), self.lib)
        a, w = ctypes.c_size_t(), ctypes.c_size_t()
        check(self.lib.hrnet_plan_workspace_bytes(self._plan, ctypes.byref(a), ctypes.byref(w)), self.lib)
        self.act_bytes, self.weight_bytes = a.value, w.value
        self.infos = _param_infos(self.lib, self._plan)

    def describe(self):
        need = ctypes.c_size_t()
        check(self.lib.hrnet_plan_describe(self._plan, None, 0, ctypes.byref(need)), self.lib)
        buf = ctypes.create_string_buffer(need.value)
        check(self.lib.hrnet_plan_describe(self._plan, buf, need.value, ctypes.byref(need)), self.lib)
        return json.loads(buf.value.decode())

    @property
    def launch_count(self):
        return check(self.lib.hrnet_plan_launch_count(self._plan), self.lib)

    def close(self):
        if getattr(self, "_plan", None) is not None and self._plan.value:
            self.lib.hrnet_plan_destroy(self._plan)
            self._plan = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class B200Engine(Plan):
    """One plan (arch, width, resolution, max batch) bound to one GPU.

    engine(images)                      -> heat-maps [n,J,H/4,W/4] fp32 (drop-in for `self.model(images)`)
    engine.forward_decode(images, ...)  -> (joints [n,J,3] (y,x,conf), argmax idx [n,J], heat-maps or None)
    """

    def __init__(self, arch, c, nof_joints, resolution, max_batch, device, flags=0, tune=None):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise HrnetError("B200Engine is CUDA-only (sm_100a); there is no CPU fallback")
        if not torch.cuda.is_available():
            raise HrnetError("no CUDA device available: the B200 engine cannot run (no CPU fallback)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        super().__init__(arch, c, nof_joints, resolution, max_batch, flags, tune)
        self._weights = None
        self._workspace = None

    # -- weights ----------------------------------------------------------------------------
    def load_state_dict(self, state_dict):
        host = pack_state_dict(state_dict, self.infos, self.weight_bytes)
        with torch.cuda.device(self.device):
            self._weights_base, self._weights = _aligned_empty(self.weight_bytes, self.device)
            self._weights.copy_(torch.from_numpy(host))
            # hrnet_plan_bind reads the BN constants / head weights back with synchronous copies on the legacy default
            # stream, which does not wait for torch's non-blocking streams: the upload must have landed
            torch.cuda.current_stream(self.device).synchronize()
            if self._workspace is None:
                self._workspace_base, self._workspace = _aligned_empty(self.act_bytes, self.device, zero=True)
            check(self.lib.hrnet_plan_bind(self._plan, self._weights.data_ptr(), self.weight_bytes,
                                           self._workspace.data_ptr(), self.act_bytes), self.lib)
            # the zero fill ran on the current stream; other streams (HostPipeline) may use the plan right away
            torch.cuda.current_stream(self.device).synchronize()
        return self

    def clone_shared(self):
        """A second engine of the same plan bound to the SAME device weights with its own workspace (for double
        buffering: `HostPipeline`)."""
        if self._weights is None:
            raise HrnetError("load_state_dict must be called before clone_shared")
        other = B200Engine("hrnet" if self.arch == _lib.ARCH_HRNET else "poseresnet", self.c, self.J, (self.H, self.W),
                           self.max_batch, self.device, flags=self.flags, tune=self.tune)
        with torch.cuda.device(self.device):
            other._weights_base, other._weights = self._weights_base, self._weights
            other._workspace_base, other._workspace = _aligned_empty(other.act_bytes, self.device, zero=True)
            check(other.lib.hrnet_plan_bind(other._plan, other._weights.data_ptr(), other.weight_bytes,
                                            other._workspace.data_ptr(), other.act_bytes), other.lib)
            torch.cuda.current_stream(self.device).synchronize()
        return other

    # -- execution --------------------------------------------------------------------------
    def _check_input(self, images):
        if self._weights is None:
            raise HrnetError("load_state_dict must be called before running the engine")
        if images.dim() != 4 or images.shape[1] != 3 or images.shape[2] != self.H or images.shape[3] != self.W:
            raise ValueError(f"expected [n,3,{self.H},{self.W}] input, got {tuple(images.shape)}")
        if images.shape[0] > self.max_batch:
            raise ValueError(f"batch {images.shape[0]} > max_batch {self.max_batch}")
        if images.device != self.device:
            raise ValueError(f"input on {images.device}, engine on {self.device}")
        return images.to(torch.float32).contiguous()

    def _outputs(self, n, boxes, return_heatmaps, joints_out):
        """Joints / idx / heat-map buffers and the device copy of the boxes, validated the same way for every entry point."""
        if joints_out is not None:
            if not (joints_out.is_cuda and joints_out.device == self.device and joints_out.is_contiguous()
                    and joints_out.dtype == torch.float32 and joints_out.numel() == n * self.J * 3):
                raise ValueError(f"joints_out must be a contiguous float32 cuda tensor with {n * self.J * 3} elements on {self.device}")
            joints = joints_out
        else:
            joints = torch.empty(n, self.J, 3, dtype=torch.float32, device=self.device)
        idx = torch.empty(n, self.J, dtype=torch.int32, device=self.device)
        hm = torch.empty(n, self.J, self.H // 4, self.W // 4, dtype=torch.float32, device=self.device) if return_heatmaps else None
        bx = None
        if boxes is not None:
            bx = torch.as_tensor(boxes, dtype=torch.float32).to(self.device).contiguous()
            if tuple(bx.shape) != (n, 4):
                raise ValueError(f"boxes must have shape ({n}, 4), got {tuple(bx.shape)}")
        return joints, idx, hm, bx

    def forward_decode(self, images, boxes=None, return_heatmaps=False, joints_out=None):
        """images [n,3,H,W] fp32 cuda -> joints [n,J,3] fp32 (y,x,conf), idx [n,J] int32, heat-maps|None.
        `joints_out` lets the caller have the decode write straight into a slice of a larger
        (e.g. all-gather) buffer."""
        x = self._check_input(images)
        n = x.shape[0]
        with torch.cuda.device(self.device):
            joints, idx, hm, bx = self._outputs(n, boxes, return_heatmaps, joints_out)
            stream = torch.cuda.current_stream(self.device).cuda_stream
            check(self.lib.hrnet_forward(self._plan, x.data_ptr(), n, hm.data_ptr() if hm is not None else None,
                                         joints.data_ptr(), idx.data_ptr(), bx.data_ptr() if bx is not None else None,
                                         stream), self.lib)
        return joints, idx, hm

    def forward_decode_u8(self, images_u8, boxes=None, return_heatmaps=False, joints_out=None):
        """Same as forward_decode, fed with the reference's pre-transform images: [n,H,W,3] BGR uint8 on the device at
        network resolution.  BGR->RGB, /255 and the ImageNet mean/std run inside the stem kernel (bit-identical to the
        host transform of SimpleHRNet.py:149-153)."""
        if self._weights is None:
            raise HrnetError("load_state_dict must be called before running the engine")
        x = images_u8
        if x.dtype != torch.uint8 or x.dim() != 4 or tuple(x.shape[1:]) != (self.H, self.W, 3):
            raise ValueError(f"expected uint8 [n,{self.H},{self.W},3], got {x.dtype} {tuple(x.shape)}")
        if x.shape[0] > self.max_batch:
            raise ValueError(f"batch {x.shape[0]} > max_batch {self.max_batch}")
        if x.device != self.device:
            raise ValueError(f"input on {x.device}, engine on {self.device}")
        x = x.contiguous()
        n = x.shape[0]
        with torch.cuda.device(self.device):
            joints, idx, hm, bx = self._outputs(n, boxes, return_heatmaps, joints_out)
            stream = torch.cuda.current_stream(self.device).cuda_stream
            check(self.lib.hrnet_forward_u8(self._plan, x.data_ptr(), n, hm.data_ptr() if hm is not None else None,
                                            joints.data_ptr(), idx.data_ptr(), bx.data_ptr() if bx is not None else None,
                                            stream), self.lib)
        return joints, idx, hm

    def forward_host_u8(self, images_u8_host, boxes_host=None, want_heatmaps=False):
        """End-to-end call with HOST uint8 images [n,H,W,3] BGR (ideally pinned): H2D (3 B/pixel) + forward + D2H."""
        if self._weights is None:
            raise HrnetError("load_state_dict must be called before running the engine")
        x = np.ascontiguousarray(images_u8_host, dtype=np.uint8)
        if x.ndim != 4 or x.shape[1:] != (self.H, self.W, 3):
            raise ValueError(f"expected uint8 [n,{self.H},{self.W},3], got {x.shape}")
        n = x.shape[0]
        self._check_host(n, boxes_host)
        joints = np.empty((n, self.J, 3), dtype=np.float32)
        idx = np.empty((n, self.J), dtype=np.int32)
        hm = np.empty((n, self.J, self.H // 4, self.W // 4), dtype=np.float32) if want_heatmaps else None
        bx = np.ascontiguousarray(boxes_host, dtype=np.float32) if boxes_host is not None else None
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            check(self.lib.hrnet_forward_host_u8(self._plan, x.ctypes.data, n, hm.ctypes.data if hm is not None else None,
                                                 joints.ctypes.data, idx.ctypes.data,
                                                 bx.ctypes.data if bx is not None else None, stream), self.lib)
        return joints, idx, hm

    def _check_host(self, n, boxes_host):
        if n > self.max_batch:
            raise ValueError(f"batch {n} > max_batch {self.max_batch}")
        if boxes_host is not None and np.asarray(boxes_host).shape != (n, 4):
            raise ValueError(f"boxes must have shape ({n}, 4), got {np.asarray(boxes_host).shape}")

    def profile_ops(self, images, iters=5):
        """[(op name, op dict, usec)] with one CUDA-event pair per kernel (serial execution, no graph)."""
        x = self._check_input(images)
        desc = self.describe()
        out = (ctypes.c_float * len(desc["ops"]))()
        with torch.cuda.device(self.device):
            check(self.lib.hrnet_profile_ops(self._plan, x.data_ptr(), x.shape[0], out, int(iters),
                                             torch.cuda.current_stream(self.device).cuda_stream), self.lib)
        return [(op["name"], op, float(t)) for op, t in zip(desc["ops"], out)], desc

    def __call__(self, images):
        """Drop-in for `self.model(images)` (SimpleHRNet.py:286): returns the heat-maps."""
        return self.forward_decode(images, return_heatmaps=True)[2]

    def forward_host(self, images_host, boxes_host=None, want_heatmaps=False):
        """End-to-end call with HOST buffers (numpy, ideally pinned): H2D + forward + decode + D2H."""
        if self._weights is None:
            raise HrnetError("load_state_dict must be called before running the engine")
        x = np.ascontiguousarray(images_host, dtype=np.float32)
        if x.ndim != 4 or x.shape[1:] != (3, self.H, self.W):     # the C side copies n*3*H*W*4 bytes from this pointer
            raise ValueError(f"expected float32 [n,3,{self.H},{self.W}], got {x.shape}")
        n = x.shape[0]
        self._check_host(n, boxes_host)
        joints = np.empty((n, self.J, 3), dtype=np.float32)
        idx = np.empty((n, self.J), dtype=np.int32)
        hm = np.empty((n, self.J, self.H // 4, self.W // 4), dtype=np.float32) if want_heatmaps else None
        bx = np.ascontiguousarray(boxes_host, dtype=np.float32) if boxes_host is not None else None
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            check(self.lib.hrnet_forward_host(self._plan, x.ctypes.data, n, hm.ctypes.data if hm is not None else None,
                                              joints.ctypes.data, idx.ctypes.data,
                                              bx.ctypes.data if bx is not None else None, stream), self.lib)
        return joints, idx, hm

    # seam compatibility with TRTModule usage in the reference (`.cuda().eval()`)
    def cuda(self):
        return self

    def eval(self):
        return self

    def to(self, *_a, **_k):
        return self


class HostPipeline:
    """Double-buffered end-to-end runner for a stream of uint8 crop batches held in HOST memory.

    The reference moves a batch to the device, runs the model and copies the result back one after the other
    (SimpleHRNet.py:282-296).  Here `depth` engines share one set of device weights; batch i+1 is copied to the device
    on its own stream while batch i computes, and the joints of batch i come back while batch i+1 computes.

        pipe = HostPipeline(engine)                  # engine: a loaded B200Engine
        for joints in pipe.run(batches):             # batches: iterable of uint8 [n,H,W,3] BGR arrays (numpy / torch, CPU)
            ...                                       # joints: numpy [n,J,3] (y, x, confidence), in submission order

    Inputs are staged through the pipeline's own pinned buffers unless they already are pinned torch tensors.
    """

    def __init__(self, engine, depth=2):
        self.engines = [engine] + [engine.clone_shared() for _ in range(depth - 1)]
        self.device = engine.device
        e = engine
        with torch.cuda.device(self.device):
            self.streams = [torch.cuda.Stream(self.device) for _ in self.engines]
        self._in = [torch.empty(e.max_batch, e.H, e.W, 3, dtype=torch.uint8).pin_memory() for _ in self.engines]
        self._joints = [torch.empty(e.max_batch, e.J, 3, dtype=torch.float32).pin_memory() for _ in self.engines]
        self._idx = [torch.empty(e.max_batch, e.J, dtype=torch.int32).pin_memory() for _ in self.engines]
        self._boxes = [torch.empty(e.max_batch, 4, dtype=torch.float32).pin_memory() for _ in self.engines]
        self._pending = [None] * len(self.engines)   # batch size in flight per slot
        self._next = 0

    def submit(self, images_u8, boxes=None):
        """Enqueue one batch; returns the slot index.  Blocks only if that slot's previous batch is still in flight
        and has not been collected (collect it first with `collect`)."""
        slot = self._next
        if self._pending[slot] is not None:
            raise HrnetError("HostPipeline slot still holds an uncollected batch: call collect() first")
        eng = self.engines[slot]
        x = torch.as_tensor(images_u8)
        if x.dtype != torch.uint8 or x.dim() != 4 or tuple(x.shape[1:]) != (eng.H, eng.W, 3):
            raise ValueError(f"expected uint8 [n,{eng.H},{eng.W},3], got {x.dtype} {tuple(x.shape)}")
        n = x.shape[0]
        if n > eng.max_batch:
            raise ValueError(f"batch {n} > max_batch {eng.max_batch}")
        if not (x.is_pinned() and x.is_contiguous()):
            self._in[slot][:n].copy_(x)
            x = self._in[slot][:n]
        bx = None
        if boxes is not None:
            self._boxes[slot][:n].copy_(torch.as_tensor(boxes, dtype=torch.float32))
            bx = self._boxes[slot]
        with torch.cuda.device(self.device):
            check(eng.lib.hrnet_forward_host_u8_async(eng._plan, x.data_ptr(), n, None, self._joints[slot].data_ptr(),
                                                      self._idx[slot].data_ptr(), bx.data_ptr() if bx is not None else None,
                                                      self.streams[slot].cuda_stream), eng.lib)
        self._pending[slot] = (n, x)       # keep the input alive until the copy has run
        self._next = (slot + 1) % len(self.engines)
        return slot

    def collect(self, slot):
        """Wait for the batch in `slot`; returns (joints [n,J,3] numpy copy, argmax idx [n,J] numpy copy)."""
        if self._pending[slot] is None:
            raise HrnetError("HostPipeline slot is empty")
        n, _ = self._pending[slot]
        self.streams[slot].synchronize()
        self._pending[slot] = None
        return self._joints[slot][:n].numpy().copy(), self._idx[slot][:n].numpy().copy()

    def run(self, batches, boxes=None):
        """Generator over an iterable of batches: yields joints per batch, in order, keeping `depth` batches in flight."""
        order = []
        for i, b in enumerate(batches):
            slot = self._next
            if self._pending[slot] is not None:
                order.remove(slot)
                yield self.collect(slot)[0]
            self.submit(b, None if boxes is None else boxes[i])
            order.append(slot)
        for slot in list(order):
            yield self.collect(slot)[0]
