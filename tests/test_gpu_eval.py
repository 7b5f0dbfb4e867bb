"""`-m gpu` parity of the rows either side of the hot path, through the C ABI:
  * hrnet_final_preds / hrnet_flip_average vs the fixtures generated from the reference's misc/utils.py
    (tests/golden/make_golden_eval.py): index / float work, so the bar is bit-exact (NaNs compare equal);
  * at BASELINE's full size (64 persons x 17 joints x 96x72) vs the numpy oracle and through size-independent properties
    (flip twice = identity, average with itself mirrored is symmetric);
  * SimpleHRNet(multiperson=True, detector=...) vs the reference's _predict_single / _predict_batch fixtures: boxes
    exact, heat-maps within the 2.5e-3 * max|hm| bar of the randomised-BN weight set, joints exactly the decode of the
    engine's own heat-maps in the adapted boxes."""
import os

import numpy as np
import pytest
import torch

from oracle import hrnet_oracle as O
from simple_hrnet_b200 import SimpleHRNet
from simple_hrnet_b200 import evaluation as E

pytestmark = pytest.mark.gpu


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"), allow_pickle=False)


def _eq(a, b):
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


def test_final_preds_fixture_bit_exact(golden_dir):
    g = _load(golden_dir, "eval_decode")
    hm = torch.from_numpy(g["heatmaps"]).cuda()
    preds, maxvals = E.get_max_preds(hm)
    assert preds.shape == (8, 17, 2) and maxvals.shape == (8, 17, 1)
    assert _eq(preds.cpu(), g["max_preds"]) and _eq(maxvals.cpu(), g["maxvals"])
    for post in (0, 1):
        fp, fm = E.get_final_preds(bool(post), hm, g["center"], g["scale"], int(g["pixel_std"]))
        assert _eq(fp.cpu(), g[f"final_preds_post{post}"]), f"post_processing={post}"
        assert _eq(fm.cpu(), g["maxvals"])


def test_flip_average_fixture_bit_exact(golden_dir):
    g = _load(golden_dir, "eval_decode")
    a, b = torch.from_numpy(g["heatmaps"]).cuda(), torch.from_numpy(g["heatmaps_flipped"]).cuda()
    pairs = g["flip_pairs"].tolist()
    assert _eq(E.flip_average(a, b, pairs).cpu(), g["flip_avg"])
    fb = E.flip_back(b, pairs).cpu().numpy()
    assert _eq(fb, O.flip_average(np.zeros_like(g["heatmaps_flipped"]), g["heatmaps_flipped"], pairs) * 2)
    # odd width / unaligned view -> scalar path of the kernel
    a2, b2 = a[:, :, :, :17].contiguous(), b[:, :, :, :17].contiguous()
    ref = O.flip_average(a2.cpu().numpy(), b2.cpu().numpy(), pairs)
    assert _eq(E.flip_average(a2, b2, pairs).cpu(), ref)


def test_final_preds_full_size_vs_oracle_and_properties():
    """BASELINE size: 64 persons, 17 joints, 96x72 maps."""
    gen = torch.Generator().manual_seed(9)
    hm = torch.randn(64, 17, 96, 72, generator=gen) * 0.2
    hm[3, 4] = -hm[3, 4].abs() - 0.1              # masked joint
    hm[5, 6, 50, 30] = hm[5, 6, 10, 60] = 7.0     # tie
    hmn = hm.numpy()
    d = hm.cuda()
    preds, maxvals = E.get_max_preds(d)
    op, om = O.get_max_preds(hmn)
    assert _eq(preds.cpu(), op) and _eq(maxvals.cpu(), om)
    center = (np.random.default_rng(1).random((64, 2)) * 400 + 20).astype(np.float32)
    scale = (np.random.default_rng(2).random((64, 2)) * 2 + 0.3).astype(np.float32)
    fp, _ = E.get_final_preds(True, d, center, scale, 200)
    ofp, _ = O.get_final_preds(True, hmn, center, scale, 200)
    assert _eq(fp.cpu(), ofp)
    pairs = [[1, 2], [3, 4], [5, 6], [7, 8], [9, 10], [11, 12], [13, 14], [15, 16]]
    twice = E.flip_back(E.flip_back(d, pairs), pairs)
    assert _eq(twice.cpu(), hmn)                  # mirroring twice and swapping twice is the identity
    avg = E.flip_average(d, d, pairs)
    assert _eq(E.flip_back(avg, pairs).cpu(), avg.cpu())   # (a + flip(a)) / 2 is invariant under flip_back


class _Detector:
    def __init__(self, dets):
        self.dets = torch.from_numpy(dets)

    def predict_single(self, image):
        return self.dets

    def predict(self, images):
        return [self.dets[:2], None, self.dets[2:]]


def test_multiperson_api_matches_reference_fixture(golden_dir):
    g = _load(golden_dir, "multiperson")
    sd = O.make_state_dict(O.hrnet_param_spec(32, 17), seed=int(g["wseed"]), bn=str(g["bn"]))
    img = np.random.default_rng(int(g["iseed"])).integers(0, 256, tuple(g["img_shape"]), dtype=np.uint8)
    res = tuple(int(v) for v in g["res"])
    m = SimpleHRNet(32, 17, sd, resolution=res, multiperson=True, return_heatmaps=True, return_bounding_boxes=True,
                    max_batch_size=2, device=torch.device("cuda:0"), detector=_Detector(g["detections"]))
    hm, boxes, pts = m.predict(img)
    assert boxes.dtype == np.int32 and np.array_equal(boxes, g["boxes"])
    ref = g["heatmaps"]
    tol = 2.5e-3 * max(1.0, float(np.abs(ref).max()))
    err = float(np.abs(hm - ref).max())
    assert err <= tol, (err, tol)
    opts, _ = O.decode_joints(hm, boxes)                      # the decode is exact on the engine's own heat-maps
    assert np.array_equal(pts.view(np.uint32), opts.view(np.uint32))
    gap_ok = np.abs(pts[..., 2] - g["pts"][..., 2]) <= tol    # confidences within the heat-map bar
    assert gap_ok.all()
    imgs = np.stack([img, img[::-1].copy(), np.ascontiguousarray(img[:, ::-1])])
    bhm, bboxes, bpts = m.predict(imgs)
    assert len(bpts) == 3 and bpts[1].shape == (0, 17, 3) and bboxes[1].shape == (0, 4)
    assert np.array_equal(np.concatenate([bboxes[0], bboxes[2]]), g["batch_boxes"])
    assert float(np.abs(np.concatenate([bhm[0], bhm[2]]) - g["batch_heatmaps"]).max()) <= tol
    bp = np.concatenate([bpts[0], bpts[2]])
    obp, _ = O.decode_joints(np.concatenate([bhm[0], bhm[2]]), g["batch_boxes"])
    assert np.array_equal(bp.view(np.uint32), obp.view(np.uint32))
