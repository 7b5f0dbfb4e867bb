"""SimpleHRNet: the reference's user-facing class (SimpleHRNet.py:12-496) re-hosted on the B200
engine.  Same constructor arguments, same `predict()` input/return formats, same error strings.

multiperson=False is the hot path named by BASELINE.json.  multiperson=True runs the
reference's glue around the path -- detection boxes -> aspect-ratio adaptation -> crops -> network -> joints rescaled
into the boxes by the decode kernel -- with a person detector supplied by the caller (`detector=`: any object with the
reference detectors' `predict_single(image)` / `predict(images)` methods, models_/detectors/YOLOv3.py:88,117): the YOLO
networks themselves are out of scope, so without a detector the constructor raises
NotImplementedError instead of silently doing something else.  The TensorRT loader raises as well."""
import numpy as np
import torch

from .engine import B200Engine

IMAGENET_MEAN = (0.485, 0.456, 0.406)   # SimpleHRNet.py:152
IMAGENET_STD = (0.229, 0.224, 0.225)


class SimpleHRNet:
    def __init__(self, c, nof_joints, checkpoint_path, model_name='HRNet', resolution=(384, 288),
                 interpolation=None, multiperson=True, return_heatmaps=False, return_bounding_boxes=False,
                 max_batch_size=32, yolo_version='v3', yolo_model_def=None, yolo_class_path=None,
                 yolo_weights_path=None, device=torch.device("cuda"), enable_tensorrt=False, engine_flags=0,
                 device_preprocess=True, detector=None, device_resize=False, device_crops=False):
        self.c = c
        self.nof_joints = nof_joints
        self.checkpoint_path = checkpoint_path
        self.model_name = model_name
        self.resolution = resolution          # (height, width)
        self.interpolation = interpolation
        self.multiperson = multiperson
        self.return_heatmaps = return_heatmaps
        self.return_bounding_boxes = return_bounding_boxes
        self.max_batch_size = max_batch_size
        self.yolo_version = yolo_version
        self.device = torch.device(device) if not isinstance(device, torch.device) else device
        self.enable_tensorrt = enable_tensorrt

        if self.multiperson:
            if self.yolo_version not in ('v3', 'v5'):
                raise ValueError('Unsopported YOLO version.')           # SimpleHRNet.py:107 (sic)
            if detector is None:
                raise NotImplementedError("multiperson=True needs a person detector and the reference's YOLO networks are "
                                          "outside the B200 hot path: pass detector=<object with predict_single(image) / "
                                          "predict(images)>, or crop people yourself and call predict() on the crops")
        self.detector = detector
        if self.enable_tensorrt:
            raise NotImplementedError("enable_tensorrt: the B200 engine already occupies the TensorRT seam")
        if model_name in ('HRNet', 'hrnet'):
            arch = "hrnet"
        elif model_name in ('PoseResNet', 'poseresnet', 'ResNet', 'resnet'):
            arch = "poseresnet"
        else:
            raise ValueError('Wrong model name.')                       # SimpleHRNet.py:114
        dev = str(self.device)
        if dev == 'cpu':
            raise ValueError("Wrong device name. The B200 engine is CUDA-only: use torch.device('cuda[:k]')")
        if not dev.startswith('cuda') or ',' in dev:
            raise ValueError('Wrong device name.')                      # SimpleHRNet.py:139
        self.model = B200Engine(arch, c, nof_joints, resolution, max_batch_size, self.device, flags=engine_flags)
        self.device = self.model.device
        checkpoint = checkpoint_path if isinstance(checkpoint_path, dict) else \
            torch.load(checkpoint_path, map_location="cpu")
        if 'model' in checkpoint:                                        # SimpleHRNet.py:118-121
            self.model.load_state_dict(checkpoint['model'])
        else:
            self.model.load_state_dict(checkpoint)
        # device-side transform (uint8 in) for HRNet at a fixed resolution; set device_preprocess=False to feed the
        # host-normalised fp32 tensor exactly like the reference does
        self._u8_path = bool(device_preprocess) and arch == "hrnet" and self.resolution is not None and not self.multiperson
        # device_resize=True: the images go to the device at their ORIGINAL size and the cubic resize of
        # SimpleHRNet.py:216-220 runs there too (OpenCV's own 8-bit cubic kernel, see preprocess.py: bit-identical to cv2
        # without its vendor path, one grey level away from the default cv2 call on a few per cent of the pixels -- hence opt-in)
        self._device_resize = bool(device_resize) and self._u8_path
        self._resizer = None
        # device_crops=True (multiperson, HRNet): the frames go to the device as they are and the per-person crops -- slice,
        # zero padding, `ToPILImage -> Resize` (Pillow's antialiased bilinear, restated bit for bit), ToTensor, Normalize --
        # are made there (hrnet_crop_resize_bilinear_u8 + the uint8 stem); boxes outside the frame fall back to the host path
        self._device_crops = bool(device_crops) and arch == "hrnet" and self.resolution is not None and bool(self.multiperson)
        self._cropper = None
        self._mp_transform = None
        self._mean = torch.tensor(IMAGENET_MEAN, dtype=torch.float32).view(3, 1, 1)
        self._std = torch.tensor(IMAGENET_STD, dtype=torch.float32).view(3, 1, 1)

    # ToTensor + Normalize of SimpleHRNet.py:149-153 on one RGB uint8 HxWx3 image
    def _transform(self, rgb_u8):
        t = torch.from_numpy(np.ascontiguousarray(rgb_u8)).permute(2, 0, 1).to(torch.float32).div(255)
        return (t - self._mean) / self._std

    def _prep(self, image):
        import cv2
        interp = cv2.INTER_CUBIC if self.interpolation is None else self.interpolation
        if self.resolution is not None:
            image = cv2.resize(image, (self.resolution[1], self.resolution[0]), interpolation=interp)
        return self._transform(cv2.cvtColor(image, cv2.COLOR_BGR2RGB))

    def predict(self, image):
        """Same contract as SimpleHRNet.predict (SimpleHRNet.py:174-210): HxWx3 or NxHxWx3 BGR uint8;
        returns pts [(n,1,)| (1,)] [J,3] as (y, x, confidence) and optionally heat-maps / boxes."""
        if len(image.shape) == 3:
            return self._predict_single(image)
        elif len(image.shape) == 4:
            return self._predict_batch(image)
        else:
            raise ValueError('Wrong image format.')

    def _resize_only(self, image):
        import cv2
        interp = cv2.INTER_CUBIC if self.interpolation is None else self.interpolation
        if self.resolution is not None:
            image = cv2.resize(image, (self.resolution[1], self.resolution[0]), interpolation=interp)
        return image

    def _to_network_size(self, images_u8):
        """uint8 [n,h,w,3] at any size -> network resolution: on the device (device_resize) or with cv2 like the reference."""
        if self.resolution is None or tuple(images_u8.shape[1:3]) == tuple(self.resolution):
            return images_u8
        if self._device_resize and (self.interpolation is None or self.interpolation == 2):   # cv2.INTER_CUBIC == 2
            if self._resizer is None:
                from .preprocess import CubicResizer
                self._resizer = CubicResizer(self.device)
            return self._resizer(np.ascontiguousarray(images_u8), self.resolution[0], self.resolution[1])
        return np.stack([self._resize_only(im) for im in images_u8])

    def _run_u8(self, images_u8, boxes):
        """HRNet fast path: the uint8 BGR crops go to the device as they are (4x fewer bytes than the fp32 tensor) and
        cvtColor + ToTensor + Normalize happen inside the stem kernel, bit-identically to the host transform."""
        n = images_u8.shape[0]
        J, Hh, Wh = self.nof_joints, self.resolution[0] // 4, self.resolution[1] // 4
        pts = np.empty((n, J, 3), dtype=np.float32)
        heatmaps = np.zeros((n, J, Hh, Wh), dtype=np.float32)
        dev = images_u8 if torch.is_tensor(images_u8) else torch.from_numpy(np.ascontiguousarray(images_u8)).to(self.device)
        for i in range(0, n, self.max_batch_size):
            sl = slice(i, min(n, i + self.max_batch_size))
            joints, _idx, hm = self.model.forward_decode_u8(dev[sl], boxes=boxes[sl], return_heatmaps=self.return_heatmaps)
            pts[sl] = joints.cpu().numpy()
            if hm is not None:
                heatmaps[sl] = hm.cpu().numpy()
        return heatmaps, pts

    def _run(self, images, boxes):
        n = images.shape[0]
        J, Hh, Wh = self.nof_joints, self.resolution[0] // 4, self.resolution[1] // 4
        pts = np.empty((n, J, 3), dtype=np.float32)
        heatmaps = np.zeros((n, J, Hh, Wh), dtype=np.float32)
        images = images.to(self.device)
        for i in range(0, n, self.max_batch_size):                      # SimpleHRNet.py:288-294
            sl = slice(i, min(n, i + self.max_batch_size))
            joints, _idx, hm = self.model.forward_decode(images[sl], boxes=boxes[sl],
                                                         return_heatmaps=self.return_heatmaps)
            pts[sl] = joints.cpu().numpy()
            if hm is not None:
                heatmaps[sl] = hm.cpu().numpy()
        return heatmaps, pts

    def _pack(self, heatmaps, boxes, pts):
        res = list()
        if self.return_heatmaps:
            res.append(heatmaps)
        if self.return_bounding_boxes:
            res.append(boxes)
        res.append(pts)
        return res if len(res) > 1 else res[0]

    # ---- multi-person glue (SimpleHRNet.py:227-278 single image, :376-412 batch) -------------------------------------
    def _crop_to_input(self, crop_rgb):
        """ToPILImage -> Resize((H, W)) -> ToTensor -> Normalize of SimpleHRNet.py:166-171 (torchvision, on the host)."""
        if self._mp_transform is None:
            from torchvision.transforms import transforms
            self._mp_transform = transforms.Compose([
                transforms.ToPILImage(), transforms.Resize((self.resolution[0], self.resolution[1])),
                transforms.ToTensor(), transforms.Normalize(mean=list(IMAGENET_MEAN), std=list(IMAGENET_STD))])
        return self._mp_transform(crop_rgb)

    def _crops_on_device(self, frames, specs):
        """uint8 network inputs [m, H, W, 3] (BGR, device) for `specs` = (frame, x1, y1, x2, y2, pads...), or None when a box
        does not lie inside its frame (numpy's slicing semantics for such boxes stay on the host path)."""
        FH, FW = frames.shape[1], frames.shape[2]
        for (_f, x1, y1, x2, y2, *_p) in specs:
            if not (0 <= x1 < x2 <= FW and 0 <= y1 < y2 <= FH):
                return None
        if self._cropper is None:
            from .preprocess import CropResizer
            self._cropper = CropResizer(self.device)
        dev = torch.from_numpy(np.ascontiguousarray(frames)).to(self.device)
        return self._cropper(dev, specs, self.resolution[0], self.resolution[1])

    @staticmethod
    def _rounded_box(det):
        x1, y1, x2, y2 = [int(round(v.item() if hasattr(v, "item") else float(v))) for v in det[:4]]
        return x1, y1, x2, y2

    def _aspect(self, x1, y1, x2, y2):
        return self.resolution[0] / self.resolution[1] * (x2 - x1) / (y2 - y1)

    def _predict_single_multiperson(self, image):
        detections = self.detector.predict_single(image)
        m = len(detections) if detections is not None else 0
        boxes = np.empty((m, 4), dtype=np.int32)
        x = torch.empty((m, 3, self.resolution[0], self.resolution[1]))
        specs, crops = [], []
        for i in range(m):
            x1, y1, x2, y2 = self._rounded_box(detections[i])
            cf = self._aspect(x1, y1, x2, y2)
            nx1, ny1, nx2, ny2, pad = x1, y1, x2, y2, None
            if cf > 1:       # too wide for the network: grow the box in y and zero-pad the crop (no neighbours pulled in)
                mid, length = y1 + (y2 - y1) // 2, int(round((y2 - y1) * cf))
                ny1, ny2 = int(mid - length // 2), int(mid + length // 2)
                pad = ((int(abs(ny1 - y1)), int(abs(ny2 - y2))), (0, 0), (0, 0))
            elif cf < 1:     # too tall: same along x
                mid, length = x1 + (x2 - x1) // 2, int(round((x2 - x1) * 1 / cf))
                nx1, nx2 = int(mid - length // 2), int(mid + length // 2)
                pad = ((0, 0), (abs(nx1 - x1), int(abs(nx2 - x2))), (0, 0))
            p = pad if pad is not None else ((0, 0), (0, 0), (0, 0))
            specs.append((0, x1, y1, x2, y2, p[0][0], p[0][1], p[1][0], p[1][1]))
            crops.append((x1, y1, x2, y2, pad))
            boxes[i] = [nx1, ny1, nx2, ny2]
        x_u8 = self._crops_on_device(image[None], specs) if (getattr(self, "_device_crops", False) and m > 0) else None
        if x_u8 is None:
            for i, (x1, y1, x2, y2, pad) in enumerate(crops):
                crop = image[y1:y2, x1:x2, ::-1]
                if pad is not None:
                    crop = np.pad(crop, pad)
                x[i] = self._crop_to_input(crop)
        if m > 0:
            heatmaps, pts = self._run_u8(x_u8, boxes) if x_u8 is not None else self._run(x, boxes)
        else:
            heatmaps = np.zeros((0, self.nof_joints, self.resolution[0] // 4, self.resolution[1] // 4), dtype=np.float32)
            pts = np.empty((0, 0, 3), dtype=np.float32)                  # SimpleHRNet.py:331
        return self._pack(heatmaps, boxes, pts)

    def _predict_batch_multiperson(self, images):
        image_detections = self.detector.predict(images)
        m = int(np.sum([len(d) for d in image_detections if d is not None]))
        boxes = np.empty((m, 4), dtype=np.int32)
        x = torch.empty((m, 3, self.resolution[0], self.resolution[1]))
        base = 0
        specs = []
        for d, detections in enumerate(image_detections):
            image = images[d]
            if detections is None or len(detections) == 0:
                continue
            for i in range(len(detections)):
                x1, y1, x2, y2 = self._rounded_box(detections[i])
                cf = self._aspect(x1, y1, x2, y2)
                if cf > 1:       # the batch path enlarges the box (clamped to the frame) instead of padding (:395-405)
                    mid, length = y1 + (y2 - y1) // 2, int(round((y2 - y1) * cf))
                    y1, y2 = max(0, mid - length // 2), min(image.shape[0], mid + length // 2)
                elif cf < 1:
                    mid, length = x1 + (x2 - x1) // 2, int(round((x2 - x1) * 1 / cf))
                    x1, x2 = max(0, mid - length // 2), min(image.shape[1], mid + length // 2)
                boxes[base + i] = [x1, y1, x2, y2]
                specs.append((d, x1, y1, x2, y2, 0, 0, 0, 0))
            base += len(detections)
        x_u8 = self._crops_on_device(images, specs) if (getattr(self, "_device_crops", False) and m > 0) else None
        if x_u8 is None:
            for i, (d, x1, y1, x2, y2, *_p) in enumerate(specs):
                x[i] = self._crop_to_input(images[d][y1:y2, x1:x2, ::-1])
        J, Hh, Wh = self.nof_joints, self.resolution[0] // 4, self.resolution[1] // 4
        if m > 0:
            heatmaps, pts = self._run_u8(x_u8, boxes) if x_u8 is not None else self._run(x, boxes)
            hm_b, box_b, pts_b, index = [], [], [], 0
            for detections in image_detections:                          # re-add the per-frame axis (:448-470)
                k = len(detections) if detections is not None else 0
                if detections is not None:
                    pts_b.append(pts[index:index + k]); hm_b.append(heatmaps[index:index + k])
                    box_b.append(boxes[index:index + k])
                else:
                    pts_b.append(np.zeros((0, J, 3), dtype=np.float32))
                    hm_b.append(np.zeros((0, J, Hh, Wh), dtype=np.float32))
                    box_b.append(np.zeros((0, 4), dtype=np.float32))
                index += k
            return self._pack(hm_b, box_b, pts_b)
        heatmaps = np.zeros((0, J, Hh, Wh), dtype=np.float32)
        boxes = np.asarray([], dtype=np.int32)                            # :478
        pts = [np.zeros((0, J, 3), dtype=np.float32) for _ in range(len(image_detections))]
        return self._pack(heatmaps, boxes, pts)

    def _predict_single(self, image):
        if self.multiperson:
            return self._predict_single_multiperson(image)
        old_res = image.shape
        boxes = np.asarray([[0, 0, old_res[1], old_res[0]]], dtype=np.float32)   # [x1, y1, x2, y2]
        if self._u8_path:
            heatmaps, pts = self._run_u8(self._to_network_size(image[None]), boxes)
        else:
            heatmaps, pts = self._run(self._prep(image).unsqueeze(dim=0), boxes)
        return self._pack(heatmaps, boxes, pts)

    def _predict_batch(self, images):
        if self.multiperson:
            return self._predict_batch_multiperson(images)
        if images.shape[0] == 0:
            raise ValueError  # the reference reaches `raise ValueError` for an empty non-multiperson batch (:487)
        old_res = images[0].shape
        boxes = np.repeat(np.asarray([[0, 0, old_res[1], old_res[0]]], dtype=np.float32), len(images), axis=0)
        if self._u8_path:
            heatmaps, pts = self._run_u8(self._to_network_size(images), boxes)
        else:
            x = torch.empty(images.shape[0], 3, self.resolution[0], self.resolution[1])
            for i, image in enumerate(images):
                x[i] = self._prep(image)
            heatmaps, pts = self._run(x, boxes)
        pts = np.expand_dims(pts, axis=1)                                # SimpleHRNet.py:475
        return self._pack(heatmaps, boxes, pts)
