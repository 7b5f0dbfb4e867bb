"""Device-side image resize of the reference's preprocessing: `cv2.resize(image, (W, H),
interpolation=cv2.INTER_CUBIC)` of SimpleHRNet.py:216-220 / :356-360 as a CUDA kernel (hrnet_resize_cubic_u8).

The kernel evaluates OpenCV's own 8-bit cubic algorithm; this module builds its tap tables the way OpenCV does
(imgproc/src/resize.cpp, resize(): per destination coordinate `f = (float)((d + 0.5) * scale - 0.5)`, first tap floor(f) - 1,
Keys weights with A = -0.75 evaluated in float32, `saturate_cast<short>(w * 2048)`; vertical pass = OpenCV's 8-lane vector
loop in float32 plus its integer tail).  Result: bit-identical to cv2.resize with cv2.setUseOptimized(False)
(tests/test_preprocess_cpu.py pins the host restatement, tests/test_gpu_kernels.py the kernel); the Intel IPP path that
the cv2 wheels take by default rounds differently -- one grey level on about 3.5 % of the pixels
(tests/test_preprocess_cpu.py measures it) -- which is why the device resize is opt-in
(`SimpleHRNet(..., device_resize=True)`) and the default stays the reference's host cv2 call."""
import ctypes

import numpy as np
import torch

from ._lib import check, load_library

_vp = ctypes.c_void_p


def _keys_weights(x):
    """Four cubic-convolution weights for the fractional offset x in [0, 1), float32 arithmetic (A = -0.75)."""
    x = np.float32(x)
    a, one = np.float32(-0.75), np.float32(1)
    w0 = ((a * (x + one) - np.float32(5) * a) * (x + one) + np.float32(8) * a) * (x + one) - np.float32(4) * a
    w1 = ((a + np.float32(2)) * x - (a + np.float32(3))) * x * x + one
    w2 = ((a + np.float32(2)) * (one - x) - (a + np.float32(3))) * (one - x) * (one - x) + one
    w3 = one - w0 - w1 - w2
    return np.array([w0, w1, w2, w3], dtype=np.float32)


def cubic_tables(src_size, dst_size):
    """(first tap index [dst] int32, weights [dst, 4] int16) for one axis."""
    scale = float(src_size) / float(dst_size)
    ofs = np.empty(dst_size, dtype=np.int32)
    coef = np.empty((dst_size, 4), dtype=np.int16)
    for d in range(dst_size):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(np.floor(f))
        w = _keys_weights(np.float32(f - np.float32(s)))
        coef[d] = np.rint(w * np.float32(2048)).astype(np.int16)          # cvRound: ties to even
        ofs[d] = s - 1
    return ofs, coef


def resize_cubic_reference(images_u8, dh, dw, lanes=8):
    """Host restatement of the kernel (numpy): used by the CPU tests to pin the tables and the two vertical-pass forms
    (vector loop in float32 for the first `lanes * floor(3 * dw / lanes)` elements of a row, integer tail) against cv2."""
    img = np.asarray(images_u8)
    n, sh, sw, _ = img.shape
    xo, xa = cubic_tables(sw, dw)
    yo, yb = cubic_tables(sh, dh)
    xi = np.clip(xo[:, None] + np.arange(4)[None, :], 0, sw - 1)
    yi = np.clip(yo[:, None] + np.arange(4)[None, :], 0, sh - 1)
    h = (img.astype(np.int64)[:, :, xi, :] * xa.astype(np.int64)[None, None, :, :, None]).sum(3)      # [n, sh, dw, 3]
    rows = h[:, yi]                                                                                    # [n, dh, 4, dw, 3]
    vi = np.clip(((rows * yb.astype(np.int64)[None, :, :, None, None]).sum(2) + (1 << 21)) >> 22, 0, 255)
    b = (yb.astype(np.float32) * (np.float32(1.0) / np.float32(2048 * 2048)))[None, :, :, None, None]
    rf = rows.astype(np.float32)
    t = rf[:, :, 3] * b[:, :, 3]
    for k in (2, 1, 0):
        t = rf[:, :, k] * b[:, :, k] + t
    vf = np.clip(np.rint(t), 0, 255)
    width = 3 * dw
    nvec = width - width % lanes
    out = vi.reshape(n, dh, width).copy()
    out[:, :, :nvec] = vf.reshape(n, dh, width)[:, :, :nvec]
    return out.reshape(n, dh, dw, 3).astype(np.uint8)


class CubicResizer:
    """Resizes uint8 [n,sh,sw,3] image stacks on the device; tap tables are cached per (source size, target size)."""

    def __init__(self, device):
        self.device = torch.device(device)
        self._tables = {}

    def _get(self, sh, sw, dh, dw):
        key = (sh, sw, dh, dw)
        if key not in self._tables:
            xo, xa = cubic_tables(sw, dw)
            yo, ya = cubic_tables(sh, dh)
            self._tables[key] = tuple(torch.from_numpy(a).to(self.device) for a in (xo, xa, yo, ya))
        return self._tables[key]

    def __call__(self, images_u8, dh, dw):
        x = torch.as_tensor(images_u8)
        if x.dtype != torch.uint8 or x.dim() != 4 or x.shape[3] != 3:
            raise ValueError(f"expected uint8 [n,h,w,3], got {x.dtype} {tuple(x.shape)}")
        x = x.to(self.device).contiguous()
        n, sh, sw, _ = x.shape
        out = torch.empty(n, dh, dw, 3, dtype=torch.uint8, device=self.device)
        xo, xa, yo, ya = self._get(sh, sw, dh, dw)
        with torch.cuda.device(self.device):
            lib = load_library()
            check(lib.hrnet_resize_cubic_u8(_vp(x.data_ptr()), n, sh, sw, _vp(out.data_ptr()), dh, dw, _vp(xo.data_ptr()),
                                            _vp(xa.data_ptr()), _vp(yo.data_ptr()), _vp(ya.data_ptr()),
                                            _vp(torch.cuda.current_stream(self.device).cuda_stream)), lib)
        return out


# ------------------------------------------------------------------------------------------------------------------
# Multi-person crops: `ToPILImage -> Resize((H, W))` of SimpleHRNet.py:166-171 is Pillow's
# antialiased bilinear resample of an 8-bit image (ImagingResample, libImaging/Resample.c): per axis a triangle filter
# whose support grows with the down-scale factor, coefficients normalised in double precision and rounded to 22-bit
# fixed point, a horizontal pass rounded to uint8, then a vertical pass.  `pil_bilinear_tables` restates the coefficient
# builder (precompute_coeffs + normalize_coeffs_8bpc), `resize_pil_bilinear_reference` the two passes; the CPU tests pin
# both against Pillow itself, the CUDA kernel (hrnet_crop_resize_bilinear_u8) evaluates the same integers.
PIL_PRECISION_BITS = 32 - 8 - 2


def pil_bilinear_tables(in_size, out_size):
    """(bounds [out, 2] int32 = (first source index, tap count), coefficients [out, ksize] int32) for one axis."""
    scale = float(in_size) / float(out_size)
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale                      # bilinear: support 1.0
    ksize = int(np.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)           # C cast: truncation (the argument is never below -0.5 ... see max)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = np.zeros(ksize, dtype=np.float64)
        ww = 0.0
        for x in range(xmax):
            a = (x + xmin - center + 0.5) * ss
            if a < 0.0:
                a = -a
            v = 1.0 - a if a < 1.0 else 0.0
            w[x] = v
            ww += v
        if ww != 0.0:
            for x in range(xmax):
                w[x] /= ww
        for x in range(ksize):
            p = w[x] * float(1 << PIL_PRECISION_BITS)
            kk[xx, x] = int(-0.5 + p) if w[x] < 0 else int(0.5 + p)
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def resize_pil_bilinear_reference(img_u8, out_h, out_w):
    """numpy restatement of Image.resize((out_w, out_h), BILINEAR) for an HxWxC uint8 image (horizontal pass first)."""
    img = np.asarray(img_u8)
    h, w, c = img.shape
    half = 1 << (PIL_PRECISION_BITS - 1)
    tmp = img
    if out_w != w:
        bx, kx = pil_bilinear_tables(w, out_w)
        tmp = np.empty((h, out_w, c), dtype=np.uint8)
        for xx in range(out_w):
            x0, n = int(bx[xx, 0]), int(bx[xx, 1])
            acc = (img[:, x0:x0 + n, :].astype(np.int64) * kx[xx, :n].astype(np.int64)[None, :, None]).sum(1) + half
            tmp[:, xx, :] = np.clip(acc >> PIL_PRECISION_BITS, 0, 255)
    out = tmp
    if out_h != h:
        by, ky = pil_bilinear_tables(h, out_h)
        out = np.empty((out_h, tmp.shape[1], c), dtype=np.uint8)
        for yy in range(out_h):
            y0, n = int(by[yy, 0]), int(by[yy, 1])
            acc = (tmp[y0:y0 + n].astype(np.int64) * ky[yy, :n].astype(np.int64)[:, None, None]).sum(0) + half
            out[yy] = np.clip(acc >> PIL_PRECISION_BITS, 0, 255)
    return out


class CropResizer:
    """Device-side `image[y1:y2, x1:x2]` (+ zero padding) -> Pillow bilinear resize to the network resolution for a list of
    boxes (hrnet_crop_resize_bilinear_u8).  Tables are cached per source extent."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.lib = load_library()
        self._cache = {}

    def _axis(self, in_size, out_size):
        key = (in_size, out_size)
        if key not in self._cache:
            b, k = pil_bilinear_tables(in_size, out_size)
            self._cache[key] = (np.concatenate([b.reshape(-1), k.reshape(-1)]).astype(np.int32), k.shape[1])
        return self._cache[key]

    def __call__(self, frames_dev, specs, out_h, out_w):
        """frames_dev: uint8 [F, FH, FW, 3] on the device.  specs: list of (frame, x1, y1, x2, y2, pad_top, pad_bottom,
        pad_left, pad_right) with the box inside the frame.  Returns uint8 [m, out_h, out_w, 3] on the device."""
        m = len(specs)
        out = torch.empty((m, out_h, out_w, 3), dtype=torch.uint8, device=self.device)
        if m == 0:
            return out
        F, FH, FW, _ = frames_dev.shape
        desc = np.zeros((m, 12), dtype=np.int32)
        chunks, off = [], 0
        for i, (f, x1, y1, x2, y2, pt, pb, pl, pr) in enumerate(specs):
            if not (0 <= x1 < x2 <= FW and 0 <= y1 < y2 <= FH and 0 <= f < F):
                raise ValueError("crop box outside the frame")
            cw, ch = (x2 - x1) + pl + pr, (y2 - y1) + pt + pb
            tx, kx = self._axis(cw, out_w)
            ty, ky = self._axis(ch, out_h)
            desc[i] = [f, x1 - pl, y1 - pt, x1, y1, x2, y2, off, off + tx.size, kx, ky, 0]
            chunks += [tx, ty]
            off += tx.size + ty.size
        tables = torch.from_numpy(np.concatenate(chunks)).to(self.device)
        desc_d = torch.from_numpy(desc).to(self.device)
        with torch.cuda.device(self.device):
            check(self.lib.hrnet_crop_resize_bilinear_u8(frames_dev.data_ptr(), F, FH, FW, desc_d.data_ptr(), tables.data_ptr(), m,
                                                         out.data_ptr(), out_h, out_w,
                                                         torch.cuda.current_stream(self.device).cuda_stream), self.lib)
        return out
