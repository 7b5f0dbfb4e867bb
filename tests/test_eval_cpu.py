"""CPU tests for the rows either side of the hot path:
  * the oracle restatements of get_max_preds / get_final_preds / flip test / multi-person crops replayed against the
    fixtures generated from the real reference (tests/golden/make_golden_eval.py) -- bit-exact;
  * the host glue of simple_hrnet_b200.SimpleHRNet(multiperson=True) (box adaptation, crops, batch re-packing) with the
    GPU call replaced by the oracle, against the same fixtures: boxes, network inputs and return shapes are exact;
  * the host-side affine of simple_hrnet_b200.evaluation against the reference matrices."""
import os
import types

import numpy as np
import pytest
import torch

from oracle import hrnet_oracle as O


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"), allow_pickle=False)


def test_oracle_eval_decode_matches_reference_fixture(golden_dir):
    g = _load(golden_dir, "eval_decode")
    hm = g["heatmaps"]
    preds, maxvals = O.get_max_preds(hm)
    assert np.array_equal(preds, g["max_preds"]) and np.array_equal(maxvals, g["maxvals"], equal_nan=True)
    for post in (0, 1):
        fp, _ = O.get_final_preds(bool(post), hm, g["center"], g["scale"], int(g["pixel_std"]))
        assert np.array_equal(fp, g[f"final_preds_post{post}"])
    assert not np.array_equal(g["final_preds_post0"], g["final_preds_post1"])     # the refinement does something
    avg = O.flip_average(hm, g["heatmaps_flipped"], g["flip_pairs"].tolist())
    assert np.array_equal(avg, g["flip_avg"], equal_nan=True)


def test_host_affine_matches_reference_matrices(golden_dir):
    from simple_hrnet_b200 import evaluation as E
    g = _load(golden_dir, "eval_decode")
    Hh, Wh = g["heatmaps"].shape[2:]
    for i in range(len(g["center"])):
        t = E.get_affine_transform(g["center"][i], g["scale"][i], int(g["pixel_std"]), 0, [Wh, Hh], inv=1)
        assert np.array_equal(t, g["trans"][i])


def test_evaluation_module_refuses_cpu_tensors():
    from simple_hrnet_b200 import evaluation as E
    with pytest.raises(ValueError):
        E.get_max_preds(torch.zeros(1, 17, 8, 8))
    with pytest.raises(AssertionError):
        E.get_max_preds(np.zeros((1, 17, 8, 8), dtype=np.float32))     # misc/utils.py:131


def test_oracle_multiperson_matches_reference_fixture(golden_dir):
    g = _load(golden_dir, "multiperson")
    img = np.random.default_rng(int(g["iseed"])).integers(0, 256, tuple(g["img_shape"]), dtype=np.uint8)
    res = tuple(int(v) for v in g["res"])
    x, boxes = O.multiperson_crops(img, g["detections"], res)
    assert np.array_equal(boxes, g["boxes"]) and boxes.dtype == np.int32
    assert np.array_equal(x.numpy(), g["inputs"])
    sd = O.make_state_dict(O.hrnet_param_spec(32, 17), seed=int(g["wseed"]), bn=str(g["bn"]))
    hm = O.hrnet_forward(sd, x).numpy()
    assert np.array_equal(hm, g["heatmaps"])
    pts, _ = O.decode_joints(hm, boxes)
    assert np.array_equal(pts, g["pts"])


class _Detector:
    def __init__(self, dets):
        self.dets = torch.from_numpy(dets)

    def predict_single(self, image):
        return self.dets

    def predict(self, images):
        return [self.dets[:2], None, self.dets[2:]]


def _cpu_api(g, sd, **kw):
    """A SimpleHRNet object whose GPU call (`_run`) is replaced by the oracle: everything else is the product's host code."""
    from simple_hrnet_b200.api import SimpleHRNet
    m = SimpleHRNet.__new__(SimpleHRNet)
    m.nof_joints = 17; m.resolution = tuple(int(v) for v in g["res"]); m.multiperson = True
    m.return_heatmaps = True; m.return_bounding_boxes = True; m.max_batch_size = 2
    m.detector = _Detector(g["detections"]); m._mp_transform = None
    for k, v in kw.items():
        setattr(m, k, v)
    seen = {}

    def _run(self, x, boxes):
        seen["x"], seen["boxes"] = x.clone(), np.array(boxes)
        hm = O.hrnet_forward(sd, x).numpy()
        return hm, O.decode_joints(hm, boxes)[0]

    m._run = types.MethodType(_run, m)
    return m, seen


def test_multiperson_host_glue_matches_reference_fixture(golden_dir):
    g = _load(golden_dir, "multiperson")
    sd = O.make_state_dict(O.hrnet_param_spec(32, 17), seed=int(g["wseed"]), bn=str(g["bn"]))
    img = np.random.default_rng(int(g["iseed"])).integers(0, 256, tuple(g["img_shape"]), dtype=np.uint8)
    m, seen = _cpu_api(g, sd)
    hm, boxes, pts = m.predict(img)
    assert boxes.dtype == np.int32 and np.array_equal(boxes, g["boxes"])
    assert np.array_equal(seen["x"].numpy(), g["inputs"])
    assert np.array_equal(hm, g["heatmaps"]) and np.array_equal(pts, g["pts"])
    # batch entry point: three frames, the middle one without detections
    imgs = np.stack([img, img[::-1].copy(), np.ascontiguousarray(img[:, ::-1])])
    bhm, bboxes, bpts = m.predict(imgs)
    assert len(bpts) == len(bboxes) == len(bhm) == 3
    assert bpts[1].shape == (0, 17, 3) and bboxes[1].shape == (0, 4) and bhm[1].shape[0] == 0
    assert np.array_equal(np.concatenate([bboxes[0], bboxes[2]]), g["batch_boxes"])
    assert np.array_equal(np.concatenate([bpts[0], bpts[2]]), g["batch_pts"])
    assert np.array_equal(np.concatenate([bhm[0], bhm[2]]), g["batch_heatmaps"])


def test_multiperson_no_detections_return_formats(golden_dir):
    g = _load(golden_dir, "multiperson")
    m, _ = _cpu_api(g, None)
    m.detector = types.SimpleNamespace(predict_single=lambda image: None, predict=lambda images: [None, None])
    img = np.zeros((40, 30, 3), dtype=np.uint8)
    hm, boxes, pts = m.predict(img)
    assert pts.shape == (0, 0, 3) and boxes.shape == (0, 4) and hm.shape[0] == 0          # SimpleHRNet.py:331
    hm, boxes, pts = m.predict(np.stack([img, img]))
    assert boxes.shape == (0,) and boxes.dtype == np.int32 and len(pts) == 2 and pts[0].shape == (0, 17, 3)   # :478-482


def test_multiperson_without_detector_raises():
    from simple_hrnet_b200.api import SimpleHRNet
    with pytest.raises(NotImplementedError):
        SimpleHRNet(32, 17, {}, multiperson=True, device=torch.device("cuda"))
    with pytest.raises(ValueError, match="Unsopported YOLO version."):
        SimpleHRNet(32, 17, {}, multiperson=True, yolo_version="v9", device=torch.device("cuda"))


@pytest.mark.parametrize("tag", ["4x3", "1x1"])
def test_box_adaptation_many_cases_match_reference(golden_dir, tag):
    """120 random detections per network aspect ratio (some touching the frame border): the boxes produced by the product's
    single-frame (zero padding) and stacked-frame (clamped enlargement) glue equal those of the reference's
    _predict_single / _predict_batch (network stubbed out when the fixture was generated)."""
    g = _load(golden_dir, "multiperson_boxes")
    res = tuple(int(v) for v in g[f"res_{tag}"])
    dets = g[f"dets_{tag}"]
    img = np.zeros(tuple(g["img_shape"]), dtype=np.uint8)
    m, seen = _cpu_api({"res": np.asarray(res), "detections": dets}, None)
    J = 17

    def _run(self, x, boxes):                       # network stubbed out like in the fixture
        return (np.zeros((len(boxes), J, res[0] // 4, res[1] // 4), dtype=np.float32),
                np.zeros((len(boxes), J, 3), dtype=np.float32))

    m._run = types.MethodType(_run, m)
    m.max_batch_size = 64
    m.detector = types.SimpleNamespace(predict_single=lambda image: torch.from_numpy(dets),
                                       predict=lambda images: [torch.from_numpy(dets[:60]), torch.from_numpy(dets[60:])])
    _, boxes, _ = m.predict(img)
    assert np.array_equal(boxes, g[f"boxes_single_{tag}"])
    oboxes = np.asarray([O.adapt_box(*[int(round(float(v))) for v in d[:4]], res)[0] for d in dets], dtype=np.int32)
    assert np.array_equal(oboxes, g[f"boxes_single_{tag}"])
    _, bboxes, _ = m.predict(np.stack([img, img]))
    assert np.array_equal(np.concatenate(bboxes), g[f"boxes_batch_{tag}"])
