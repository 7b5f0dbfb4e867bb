"""CPU emulation of a plan's op list (hrnet_plan_describe JSON) with torch ops, reading weights
back out of the PACKED buffer the product packer produced.  Test infrastructure only: it checks,
without a GPU, that the C++ graph builder (op order, wiring, fusion shifts, residuals, parameter
manifest) and the Python weight packer restate the reference network.  Rounding points mirror the
CUDA path: fp16 storage of activations, fp32 accumulate / BN / residual / fusion math."""
import numpy as np
import torch
import torch.nn.functional as F

OP_STEM, OP_CONV, OP_FUSE, OP_HEAD, OP_ARGMAX, OP_MAXPOOL, OP_STEM7 = range(7)


def _f16(t):
    return t.to(torch.float16).to(torch.float32)


def _param(buf, pi):
    n = pi.cout * pi.cin * pi.kh * pi.kw
    if pi.w_f32:
        w = np.frombuffer(buf, dtype=np.float32, count=n, offset=pi.w_offset)
    else:
        w = np.frombuffer(buf, dtype=np.float16, count=n, offset=pi.w_offset).astype(np.float32)
    w = torch.from_numpy(w.copy()).view(pi.cout, pi.kh, pi.kw, pi.cin).permute(0, 3, 1, 2).contiguous()
    scale = torch.from_numpy(np.frombuffer(buf, dtype=np.float32, count=pi.cout, offset=pi.scale_offset).copy())
    bias = torch.from_numpy(np.frombuffer(buf, dtype=np.float32, count=pi.cout, offset=pi.bias_offset).copy())
    return w, scale, bias


@torch.no_grad()
def emulate(desc, infos, packed, x, exact=False):
    """desc: plan JSON; infos: HrnetParamInfo list; packed: np.uint8 buffer; x: [n,3,H,W] fp32.
    exact=True skips the fp16 activation rounding (graph/packing check against the fp32 oracle with
    only the weight rounding left).  Returns (heatmaps [n,J,h,w] fp32, taps dict name->tensor)."""
    rnd = (lambda t: t) if exact else _f16
    T = {desc["input"]: x}
    named = {}
    for op in desc["ops"]:
        k = op["kind"]
        if k in (OP_STEM, OP_STEM7):
            w, s, b = _param(packed, infos[op["param"]])
            y = F.conv2d(T[op["in"]], w, stride=2, padding=op["pad"])
            T[op["out"]] = rnd(F.relu(y * s.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)))
        elif k == OP_MAXPOOL:
            T[op["out"]] = F.max_pool2d(T[op["in"]], 3, 2, 1)
        elif k == OP_CONV:
            w, s, b = _param(packed, infos[op["param"]])
            xin = T[op["in"]]
            tout = desc["tensors"][op["out"]]
            if op["pad"] >= 100:  # transposed-conv sub-pixel phase
                a, bb = divmod(op["pad"] - 100, 2)
                xp = F.pad(xin, (1 - bb, bb, 1 - a, a))
                y = F.conv2d(xp, w)
            else:
                y = F.conv2d(xin, w, stride=op["stride"], padding=op["pad"])
            y = y * s.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)
            if op["res"] >= 0:
                y = y + T[op["res"]]
            if op["relu"]:
                y = F.relu(y)
            if not tout["f32"]:
                y = rnd(y)
            if op["pad"] >= 100:
                full = T.get(op["out"])
                if full is None:
                    full = torch.zeros(xin.shape[0], tout["C"], tout["H"], tout["W"])
                    T[op["out"]] = full
                full[:, :, a::2, bb::2] = y
            else:
                T[op["out"]] = y
        elif k == OP_FUSE:
            acc = None
            for j in range(op["nsrc"]):
                t = T[op["src"][j]]
                sh = op["shift"][j]
                if sh:
                    t = F.interpolate(t, scale_factor=float(2 ** sh), mode="nearest")
                acc = t if acc is None else acc + t
            T[op["out"]] = rnd(F.relu(acc) if op["relu"] else acc)
        elif k == OP_HEAD:
            w, s, b = _param(packed, infos[op["param"]])
            T[op["out"]] = F.conv2d(T[op["in"]], w) + b.view(1, -1, 1, 1)
        elif k == OP_ARGMAX:
            pass
        else:
            raise ValueError(k)
        if op["out"] >= 0:
            named[op["name"]] = T[op["out"]]
    return T[desc["heatmaps"]], named
