"""Evaluation-side decode of the reference (`misc/utils.py`), re-hosted on the B200 kernels: same function names,
argument meaning and return shapes, tensors stay on the GPU.

  get_max_preds(batch_heatmaps)                                   misc/utils.py:125-151
  get_final_preds(post_processing, batch_heatmaps, center, scale, pixel_std)   misc/utils.py:154-182
  flip_back(output_flipped, matched_parts)                        misc/utils.py:19-29
  flip_average(output, output_flipped, matched_parts)             training/COCO.py:206-212 `(output + flip_back(...)) * 0.5`

Host work is limited to what the reference also does on the host with cv2: the 2x3 inverse affine of each person
(`get_affine_transform(..., inv=1)`, misc/utils.py:46-79).  CUDA-only: there is no CPU fallback."""
import ctypes

import numpy as np
import torch

from ._lib import check, load_library

_vp = ctypes.c_void_p


def _f32_cuda(t, what):
    if not isinstance(t, torch.Tensor):
        raise AssertionError(f'{what} should be torch.Tensor')           # misc/utils.py:131
    if not t.is_cuda:
        raise ValueError(f"{what} must live on a CUDA device (the B200 engine has no CPU path)")
    return t.detach().to(torch.float32).contiguous()


def _stream(t):
    return _vp(torch.cuda.current_stream(t.device).cuda_stream)


def _get_3rd_point(a, b):
    d = a - b
    return b + np.array([-d[1], d[0]], dtype=np.float32)


def get_affine_transform(center, scale, pixel_std, rot, output_size, shift=np.array([0, 0], dtype=np.float32), inv=0):
    """misc/utils.py:46-79: crop <-> image affine from three point pairs (cv2.getAffineTransform, float64 2x3)."""
    import cv2
    if not isinstance(scale, np.ndarray) and not isinstance(scale, list):
        scale = np.array([scale, scale])
    scale_tmp = scale * 1.0 * pixel_std
    src_w = scale_tmp[0]
    dst_w, dst_h = output_size[0], output_size[1]
    rot_rad = np.pi * rot / 180
    sn, cs = np.sin(rot_rad), np.cos(rot_rad)
    p0, p1 = 0, src_w * -0.5
    src_dir = [p0 * cs - p1 * sn, p0 * sn + p1 * cs]
    dst_dir = np.array([0, dst_w * -0.5], np.float32)
    src = np.zeros((3, 2), dtype=np.float32)
    dst = np.zeros((3, 2), dtype=np.float32)
    src[0, :] = center + scale_tmp * shift
    src[1, :] = center + src_dir + scale_tmp * shift
    dst[0, :] = [dst_w * 0.5, dst_h * 0.5]
    dst[1, :] = np.array([dst_w * 0.5, dst_h * 0.5]) + dst_dir
    src[2:, :] = _get_3rd_point(src[0, :], src[1, :])
    dst[2:, :] = _get_3rd_point(dst[0, :], dst[1, :])
    if inv:
        return cv2.getAffineTransform(np.float32(dst), np.float32(src))
    return cv2.getAffineTransform(np.float32(src), np.float32(dst))


def _decode(batch_heatmaps, post_processing, trans):
    hm = _f32_cuda(batch_heatmaps, 'batch_heatmaps')
    assert hm.dim() == 4, 'batch_images should be 4-ndim'               # misc/utils.py:132
    n, J, Hh, Wh = hm.shape
    preds = torch.empty(n, J, 2, dtype=torch.float32, device=hm.device)
    maxvals = torch.empty(n, J, 1, dtype=torch.float32, device=hm.device)
    t_dev = None
    if trans is not None:
        t_dev = torch.from_numpy(np.ascontiguousarray(trans, dtype=np.float64)).to(hm.device)
        assert t_dev.shape == (n, 2, 3)
    with torch.cuda.device(hm.device):
        lib = load_library()
        check(lib.hrnet_final_preds(_vp(hm.data_ptr()), n, J, Hh, Wh, int(bool(post_processing)),
                                    _vp(t_dev.data_ptr()) if t_dev is not None else None, _vp(preds.data_ptr()),
                                    _vp(maxvals.data_ptr()), _stream(hm)), lib)
    return preds, maxvals


def get_max_preds(batch_heatmaps):
    """heatmaps [n,J,H,W] -> (preds [n,J,2] as (x, y) heat-map pixels, maxvals [n,J,1])."""
    return _decode(batch_heatmaps, False, None)


def get_final_preds(post_processing, batch_heatmaps, center, scale, pixel_std):
    """Quarter-pixel refined joints transformed back to image coordinates; center / scale: [n,2] numpy arrays."""
    n, _, Hh, Wh = batch_heatmaps.shape
    trans = np.stack([get_affine_transform(center[i], scale[i], pixel_std, 0, [Wh, Hh], inv=1) for i in range(n)]) \
        if n else np.zeros((0, 2, 3))
    return _decode(batch_heatmaps, post_processing, trans)


def _perm(matched_parts, J):
    perm = list(range(J))
    for a, b in matched_parts:          # sequential swaps, like the reference's loop
        perm[a], perm[b] = perm[b], perm[a]
    return (ctypes.c_int32 * J)(*perm)


def flip_average(output, output_flipped, matched_parts):
    """`(output + flip_back(output_flipped, matched_parts)) * 0.5` in one pass over the maps."""
    a = _f32_cuda(output, 'output')
    b = _f32_cuda(output_flipped, 'output_flipped')
    assert a.dim() == 4 and a.shape == b.shape, 'output_flipped has to be [batch_size, num_joints, height, width]'
    n, J, Hh, Wh = a.shape
    out = torch.empty_like(a)
    with torch.cuda.device(a.device):
        lib = load_library()
        check(lib.hrnet_flip_average(_vp(a.data_ptr()), _vp(b.data_ptr()), _perm(matched_parts, J), n, J, Hh, Wh,
                                     _vp(out.data_ptr()), _stream(a)), lib)
    return out


def flip_back(output_flipped, matched_parts):
    """Mirror along x and swap the left / right joints back (== 2 * flip_average(0, output_flipped) without rounding:
    (0 + x) * 0.5 * 2 is exact in binary floating point)."""
    b = _f32_cuda(output_flipped, 'output_flipped')
    return flip_average(torch.zeros_like(b), b, matched_parts) * 2
