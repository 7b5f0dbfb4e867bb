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


def crop_to_image_affines(center, scale, pixel_std, heatmap_size):
    """[n,2,3] float64 matrices that map heat-map pixels back to image pixels -- what the reference obtains per person
    from `get_affine_transform(center, scale, pixel_std, 0, [Wh, Hh], inv=1)` (misc/utils.py:46-79; the evaluation
    path never rotates).

    An axis-aligned box of `scale * pixel_std` pixels around `center` corresponds to the heat-map; the reference pins the
    map with three point pairs (box centre, the point half a box width above it, and that offset turned by 90 degrees),
    stores them as float32 and lets `cv2.getAffineTransform` solve for the matrix in double precision.  The same three
    pairs are built here for all persons at once, with the same float32 / float64 roundings (the arithmetic keeps the
    caller's dtype of `scale` exactly as NumPy does in the reference), so the matrices agree bit for bit
    (tests/test_eval_cpu.py)."""
    import cv2
    center = np.asarray(center)
    scale = np.asarray(scale)
    n = center.shape[0]
    box = scale * 1.0 * pixel_std                         # [n,2] box size in image pixels (dtype of `scale`)
    up_img = box[:, 0] * -0.5                             # half a box WIDTH, upwards
    wm, hm = heatmap_size[0], heatmap_size[1]
    img = np.zeros((n, 3, 2), dtype=np.float32)
    img[:, 0] = center
    img[:, 1, 0] = center[:, 0] + np.zeros(n)             # float64 sums, rounded once when stored (as in the reference)
    img[:, 1, 1] = center[:, 1] + up_img.astype(np.float64)
    d = img[:, 0] - img[:, 1]                             # float32
    img[:, 2, 0] = img[:, 1, 0] - d[:, 1]
    img[:, 2, 1] = img[:, 1, 1] + d[:, 0]
    hmp = np.zeros((3, 2), dtype=np.float32)
    hmp[0] = [wm * 0.5, hm * 0.5]
    hmp[1] = np.array([wm * 0.5, hm * 0.5]) + np.array([0, wm * -0.5], np.float32)
    dh = hmp[0] - hmp[1]
    hmp[2] = hmp[1] + np.array([-dh[1], dh[0]], dtype=np.float32)
    return np.stack([cv2.getAffineTransform(hmp, img[i]) for i in range(n)]) if n else np.zeros((0, 2, 3))


def get_affine_transform(center, scale, pixel_std, rot, output_size, shift=None, inv=0):
    """Reference signature (misc/utils.py:46) for the one case the evaluation decode uses: rot = 0, no shift, inv = 1."""
    if rot != 0 or not inv or (shift is not None and np.any(np.asarray(shift) != 0)):
        raise NotImplementedError("only the inverse, unrotated, unshifted crop affine of the evaluation decode is provided")
    if not isinstance(scale, (np.ndarray, list)):
        scale = np.array([scale, scale])
    return crop_to_image_affines(np.asarray(center)[None], np.asarray(scale)[None], pixel_std, output_size)[0]


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
    return _decode(batch_heatmaps, post_processing, crop_to_image_affines(center, scale, pixel_std, [Wh, Hh]))


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
