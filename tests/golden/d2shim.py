"""Minimal stand-in for the detectron2 (v0.6) / fvcore modules the reference imports, installed into ``sys.modules`` so
that the UNMODIFIED reference files under /root/reference can be imported and executed in this container
(detectron2 / fvcore / yacs are not installable here).  TEST INFRASTRUCTURE ONLY - used by
``tests/golden/make_golden_ref.py`` to generate ``golden_ref_v1.npz``; nothing in the product imports it and it does
not import ``oracle/`` (the fixtures must not be generated from the oracle).

Every class/function restates the public detectron2 v0.6 behaviour of the same name (the reference pins the
cu113/torch1.10 wheel index = v0.6, /root/reference/README.md:24); the numeric kernels are the real binaries
(``torchvision.ops.roi_align`` / ``nms`` / ``batched_nms``, ATen).  Only what the RoI hot path touches is filled in;
everything else the reference imports at module top (losses, registries, event storage) is a recording stub.
"""
from __future__ import annotations

import itertools
import math
import sys
import types
from collections import namedtuple
from typing import Any, Dict, List, Tuple, Union

import numpy as np
import torch
import torchvision
from torch import nn
from torch.nn import functional as F


# ------------------------------------------------------------------------------------------------ detectron2.layers
def cat(tensors: List[torch.Tensor], dim: int = 0):
    assert isinstance(tensors, (list, tuple))
    if len(tensors) == 1:
        return tensors[0]
    return torch.cat(tensors, dim)


def nonzero_tuple(x):
    return x.nonzero(as_tuple=True)


def batched_nms(boxes: torch.Tensor, scores: torch.Tensor, idxs: torch.Tensor, iou_threshold: float):
    # v0.6: torchvision picks coordinate-trick vs per-class loop itself; fp16 lacks the range, hence float()
    assert boxes.shape[-1] == 4
    return torchvision.ops.boxes.batched_nms(boxes.float(), scores, idxs, iou_threshold)


def cross_entropy(input, target, *, reduction="mean", **kwargs):
    if target.numel() == 0 and reduction == "mean":
        return input.sum() * 0.0
    return F.cross_entropy(input, target, reduction=reduction, **kwargs)


class ShapeSpec(namedtuple("_ShapeSpec", ["channels", "height", "width", "stride"])):
    def __new__(cls, channels=None, height=None, width=None, stride=None):
        return super().__new__(cls, channels, height, width, stride)


class Conv2d(torch.nn.Conv2d):
    def __init__(self, *args, **kwargs):
        norm = kwargs.pop("norm", None)
        activation = kwargs.pop("activation", None)
        super().__init__(*args, **kwargs)
        self.norm = norm
        self.activation = activation

    def forward(self, x):
        x = F.conv2d(x, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)
        if self.norm is not None:
            x = self.norm(x)
        if self.activation is not None:
            x = self.activation(x)
        return x


class ROIAlign(nn.Module):
    def __init__(self, output_size, spatial_scale, sampling_ratio, aligned=True):
        super().__init__()
        self.output_size = output_size
        self.spatial_scale = spatial_scale
        self.sampling_ratio = sampling_ratio
        self.aligned = aligned

    def forward(self, input, rois):
        assert rois.dim() == 2 and rois.size(1) == 5
        return torchvision.ops.roi_align(input, rois.to(dtype=input.dtype), self.output_size, self.spatial_scale,
                                         self.sampling_ratio, self.aligned)


def _not_on_path(name):
    def f(*a, **k):
        raise NotImplementedError(f"d2shim: {name} is not on the RoI hot path")
    f.__name__ = name
    return f


# -------------------------------------------------------------------------------------------- detectron2.structures
class Boxes:
    def __init__(self, tensor: torch.Tensor):
        device = tensor.device if isinstance(tensor, torch.Tensor) else torch.device("cpu")
        tensor = torch.as_tensor(tensor, dtype=torch.float32, device=device)
        if tensor.numel() == 0:
            tensor = tensor.reshape((-1, 4)).to(dtype=torch.float32, device=device)
        assert tensor.dim() == 2 and tensor.size(-1) == 4, tensor.size()
        self.tensor = tensor

    def clone(self) -> "Boxes":
        return Boxes(self.tensor.clone())

    def to(self, device):
        return Boxes(self.tensor.to(device=device))

    def area(self) -> torch.Tensor:
        box = self.tensor
        return (box[:, 2] - box[:, 0]) * (box[:, 3] - box[:, 1])

    def clip(self, box_size: Tuple[int, int]) -> None:
        assert torch.isfinite(self.tensor).all(), "Box tensor contains infinite or NaN!"
        h, w = box_size
        x1 = self.tensor[:, 0].clamp(min=0, max=w)
        y1 = self.tensor[:, 1].clamp(min=0, max=h)
        x2 = self.tensor[:, 2].clamp(min=0, max=w)
        y2 = self.tensor[:, 3].clamp(min=0, max=h)
        self.tensor = torch.stack((x1, y1, x2, y2), dim=-1)

    def nonempty(self, threshold: float = 0.0) -> torch.Tensor:
        box = self.tensor
        widths = box[:, 2] - box[:, 0]
        heights = box[:, 3] - box[:, 1]
        return (widths > threshold) & (heights > threshold)

    def __getitem__(self, item) -> "Boxes":
        if isinstance(item, int):
            return Boxes(self.tensor[item].view(1, -1))
        b = self.tensor[item]
        assert b.dim() == 2, "Indexing on Boxes with {} failed to return a matrix!".format(item)
        return Boxes(b)

    def __len__(self) -> int:
        return self.tensor.shape[0]

    def __repr__(self) -> str:
        return "Boxes(" + str(self.tensor) + ")"

    def get_centers(self) -> torch.Tensor:
        return (self.tensor[:, :2] + self.tensor[:, 2:]) / 2

    @classmethod
    def cat(cls, boxes_list: List["Boxes"]) -> "Boxes":
        assert isinstance(boxes_list, (list, tuple))
        if len(boxes_list) == 0:
            return cls(torch.empty(0))
        assert all([isinstance(box, Boxes) for box in boxes_list])
        return cls(torch.cat([b.tensor for b in boxes_list], dim=0))

    @property
    def device(self):
        return self.tensor.device

    def __iter__(self):
        yield from self.tensor


def pairwise_intersection(boxes1: Boxes, boxes2: Boxes) -> torch.Tensor:
    boxes1, boxes2 = boxes1.tensor, boxes2.tensor
    width_height = torch.min(boxes1[:, None, 2:], boxes2[:, 2:]) - torch.max(boxes1[:, None, :2], boxes2[:, :2])
    width_height.clamp_(min=0)
    return width_height.prod(dim=2)


def pairwise_iou(boxes1: Boxes, boxes2: Boxes) -> torch.Tensor:
    area1 = boxes1.area()
    area2 = boxes2.area()
    inter = pairwise_intersection(boxes1, boxes2)
    return torch.where(inter > 0, inter / (area1[:, None] + area2 - inter),
                       torch.zeros(1, dtype=inter.dtype, device=inter.device))


class Instances:
    def __init__(self, image_size: Tuple[int, int], **kwargs: Any):
        self._image_size = image_size
        self._fields: Dict[str, Any] = {}
        for k, v in kwargs.items():
            self.set(k, v)

    @property
    def image_size(self) -> Tuple[int, int]:
        return self._image_size

    def __setattr__(self, name: str, val: Any) -> None:
        if name.startswith("_"):
            super().__setattr__(name, val)
        else:
            self.set(name, val)

    def __getattr__(self, name: str) -> Any:
        if name == "_fields" or name not in self._fields:
            raise AttributeError("Cannot find field '{}' in the given Instances!".format(name))
        return self._fields[name]

    def set(self, name: str, value: Any) -> None:
        data_len = len(value)
        if len(self._fields):
            assert len(self) == data_len, "Adding a field of length {} to a Instances of length {}".format(data_len, len(self))
        self._fields[name] = value

    def has(self, name: str) -> bool:
        return name in self._fields

    def remove(self, name: str) -> None:
        del self._fields[name]

    def get(self, name: str) -> Any:
        return self._fields[name]

    def get_fields(self) -> Dict[str, Any]:
        return self._fields

    def to(self, *args: Any, **kwargs: Any) -> "Instances":
        ret = Instances(self._image_size)
        for k, v in self._fields.items():
            if hasattr(v, "to"):
                v = v.to(*args, **kwargs)
            ret.set(k, v)
        return ret

    def __getitem__(self, item: Union[int, slice, torch.BoolTensor]) -> "Instances":
        if type(item) == int:
            if item >= len(self) or item < -len(self):
                raise IndexError("Instances index out of range!")
            else:
                item = slice(item, None, len(self))
        ret = Instances(self._image_size)
        for k, v in self._fields.items():
            ret.set(k, v[item])
        return ret

    def __len__(self) -> int:
        for v in self._fields.values():
            return v.__len__()
        raise NotImplementedError("Empty Instances does not support __len__!")

    def __iter__(self):
        raise NotImplementedError("`Instances` object is not iterable!")

    @staticmethod
    def cat(instance_lists: List["Instances"]) -> "Instances":
        assert all(isinstance(i, Instances) for i in instance_lists)
        assert len(instance_lists) > 0
        if len(instance_lists) == 1:
            return instance_lists[0]
        image_size = instance_lists[0].image_size
        for i in instance_lists[1:]:
            assert i.image_size == image_size
        ret = Instances(image_size)
        for k in instance_lists[0]._fields.keys():
            values = [i.get(k) for i in instance_lists]
            v0 = values[0]
            if isinstance(v0, torch.Tensor):
                values = torch.cat(values, dim=0)
            elif isinstance(v0, list):
                values = list(itertools.chain(*values))
            elif hasattr(type(v0), "cat"):
                values = type(v0).cat(values)
            else:
                raise ValueError("Unsupported type {} for concatenation".format(type(v0)))
            ret.set(k, values)
        return ret


class ImageList:
    def __init__(self, tensor: torch.Tensor, image_sizes: List[Tuple[int, int]]):
        self.tensor = tensor
        self.image_sizes = image_sizes

    def __len__(self):
        return len(self.image_sizes)


# ------------------------------------------------------------------------------------------------ detectron2.config
class CfgNode(dict):
    """Attribute-access dict; enough for the reference's ``from_config`` classmethods and ``add_openset_rcnn_config``."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


def _called_with_cfg(*args, **kwargs):
    if len(args) and isinstance(args[0], CfgNode):
        return True
    if isinstance(kwargs.pop("cfg", None), CfgNode):
        return True
    return False


def _get_args_from_config(from_config_func, *args, **kwargs):
    import inspect
    signature = inspect.signature(from_config_func)
    if list(signature.parameters.keys())[0] != "cfg":
        raise TypeError("from_config must take 'cfg' as the first argument!")
    support_var_arg = any(p.kind in [p.VAR_POSITIONAL, p.VAR_KEYWORD] for p in signature.parameters.values())
    if support_var_arg:
        ret = from_config_func(*args, **kwargs)
    else:
        supported = set(signature.parameters.keys())
        extra = {k: kwargs.pop(k) for k in list(kwargs.keys()) if k not in supported}
        ret = from_config_func(*args, **kwargs)
        ret.update(extra)
    return ret


def configurable(init_func=None, *, from_config=None):
    import functools
    assert init_func is not None and from_config is None, "d2shim: only the @configurable __init__ form is used"
    assert init_func.__name__ == "__init__"

    @functools.wraps(init_func)
    def wrapped(self, *args, **kwargs):
        try:
            from_config_func = type(self).from_config
        except AttributeError as e:
            raise AttributeError("Class with @configurable must have a 'from_config' classmethod.") from e
        if _called_with_cfg(*args, **kwargs):
            explicit_args = _get_args_from_config(from_config_func, *args, **kwargs)
            init_func(self, **explicit_args)
        else:
            init_func(self, *args, **kwargs)
    return wrapped


# -------------------------------------------------------------------------------------------------- detectron2.data
class _Metadata(types.SimpleNamespace):
    pass


class _MetadataCatalog:
    def __init__(self):
        self._d: Dict[str, _Metadata] = {}

    def get(self, name):
        if name not in self._d:
            self._d[name] = _Metadata(name=name)
        return self._d[name]


MetadataCatalog = _MetadataCatalog()


# ------------------------------------------------------------------------------------------- detectron2.utils.events
class EventStorage:
    def __init__(self):
        self.scalars: Dict[str, list] = {}

    def put_scalar(self, name, value, smoothing_hint=True):
        self.scalars.setdefault(name, []).append(float(value))

    def __enter__(self):
        _STORAGE_STACK.append(self)
        return self

    def __exit__(self, *a):
        _STORAGE_STACK.pop()


_STORAGE_STACK: List[EventStorage] = [EventStorage()]


def get_event_storage():
    return _STORAGE_STACK[-1]


def retry_if_cuda_oom(func):
    return func


# ------------------------------------------------------------------------------------------------ registries
class Registry:
    def __init__(self, name):
        self._name = name
        self._obj_map: Dict[str, Any] = {}

    def register(self, obj=None):
        if obj is None:
            def deco(func_or_class):
                self._obj_map[func_or_class.__name__] = func_or_class
                return func_or_class
            return deco
        self._obj_map[obj.__name__] = obj

    def get(self, name):
        return self._obj_map[name]


RPN_HEAD_REGISTRY = Registry("RPN_HEAD")
PROPOSAL_GENERATOR_REGISTRY = Registry("PROPOSAL_GENERATOR")
ROI_HEADS_REGISTRY = Registry("ROI_HEADS")


# --------------------------------------------------------------------------------- detectron2.modeling.anchor_generator
class DefaultAnchorGenerator(nn.Module):
    box_dim = 4

    def __init__(self, *, sizes, aspect_ratios, strides, offset=0.0):
        super().__init__()
        self.strides = strides
        self.num_features = len(strides)
        sizes = list(sizes) * self.num_features if len(sizes) == 1 else sizes
        aspect_ratios = list(aspect_ratios) * self.num_features if len(aspect_ratios) == 1 else aspect_ratios
        self.cell_anchors = [self.generate_cell_anchors(s, a).float() for s, a in zip(sizes, aspect_ratios)]
        self.offset = offset
        assert 0.0 <= self.offset < 1.0, self.offset

    @property
    def num_cell_anchors(self):
        return [len(c) for c in self.cell_anchors]

    @staticmethod
    def generate_cell_anchors(sizes=(32, 64, 128, 256, 512), aspect_ratios=(0.5, 1, 2)):
        anchors = []
        for size in sizes:
            area = size ** 2.0
            for aspect_ratio in aspect_ratios:
                w = math.sqrt(area / aspect_ratio)
                h = aspect_ratio * w
                x0, y0, x1, y1 = -w / 2.0, -h / 2.0, w / 2.0, h / 2.0
                anchors.append([x0, y0, x1, y1])
        return torch.tensor(anchors)

    def _grid_anchors(self, grid_sizes):
        anchors = []
        for size, stride, base_anchors in zip(grid_sizes, self.strides, self.cell_anchors):
            grid_height, grid_width = size
            shifts_x = torch.arange(self.offset * stride, grid_width * stride, step=stride, dtype=torch.float32)
            shifts_y = torch.arange(self.offset * stride, grid_height * stride, step=stride, dtype=torch.float32)
            shift_y, shift_x = torch.meshgrid(shifts_y, shifts_x, indexing="ij")
            shift_x = shift_x.reshape(-1)
            shift_y = shift_y.reshape(-1)
            shifts = torch.stack((shift_x, shift_y, shift_x, shift_y), dim=1)
            anchors.append((shifts.view(-1, 1, 4) + base_anchors.view(1, -1, 4)).reshape(-1, 4))
        return anchors

    def forward(self, features: List[torch.Tensor]):
        grid_sizes = [feature_map.shape[-2:] for feature_map in features]
        return [Boxes(x) for x in self._grid_anchors(grid_sizes)]


# --------------------------------------------------------------------------------- detectron2.modeling.box_regression
_DEFAULT_SCALE_CLAMP = math.log(1000.0 / 16)


class Box2BoxTransform(object):
    def __init__(self, weights: Tuple[float, float, float, float], scale_clamp: float = _DEFAULT_SCALE_CLAMP):
        self.weights = weights
        self.scale_clamp = scale_clamp

    def get_deltas(self, src_boxes, target_boxes):
        src_widths = src_boxes[:, 2] - src_boxes[:, 0]
        src_heights = src_boxes[:, 3] - src_boxes[:, 1]
        src_ctr_x = src_boxes[:, 0] + 0.5 * src_widths
        src_ctr_y = src_boxes[:, 1] + 0.5 * src_heights
        target_widths = target_boxes[:, 2] - target_boxes[:, 0]
        target_heights = target_boxes[:, 3] - target_boxes[:, 1]
        target_ctr_x = target_boxes[:, 0] + 0.5 * target_widths
        target_ctr_y = target_boxes[:, 1] + 0.5 * target_heights
        wx, wy, ww, wh = self.weights
        dx = wx * (target_ctr_x - src_ctr_x) / src_widths
        dy = wy * (target_ctr_y - src_ctr_y) / src_heights
        dw = ww * torch.log(target_widths / src_widths)
        dh = wh * torch.log(target_heights / src_heights)
        deltas = torch.stack((dx, dy, dw, dh), dim=1)
        assert (src_widths > 0).all().item(), "Input boxes to Box2BoxTransform are not valid!"
        return deltas

    def apply_deltas(self, deltas, boxes):
        deltas = deltas.float()  # ensure fp32 for decoding precision
        boxes = boxes.to(deltas.dtype)
        widths = boxes[:, 2] - boxes[:, 0]
        heights = boxes[:, 3] - boxes[:, 1]
        ctr_x = boxes[:, 0] + 0.5 * widths
        ctr_y = boxes[:, 1] + 0.5 * heights
        wx, wy, ww, wh = self.weights
        dx = deltas[:, 0::4] / wx
        dy = deltas[:, 1::4] / wy
        dw = deltas[:, 2::4] / ww
        dh = deltas[:, 3::4] / wh
        dw = torch.clamp(dw, max=self.scale_clamp)
        dh = torch.clamp(dh, max=self.scale_clamp)
        pred_ctr_x = dx * widths[:, None] + ctr_x[:, None]
        pred_ctr_y = dy * heights[:, None] + ctr_y[:, None]
        pred_w = torch.exp(dw) * widths[:, None]
        pred_h = torch.exp(dh) * heights[:, None]
        x1 = pred_ctr_x - 0.5 * pred_w
        y1 = pred_ctr_y - 0.5 * pred_h
        x2 = pred_ctr_x + 0.5 * pred_w
        y2 = pred_ctr_y + 0.5 * pred_h
        pred_boxes = torch.stack((x1, y1, x2, y2), dim=-1)
        return pred_boxes.reshape(deltas.shape)


class Box2BoxTransformLinear(object):
    def __init__(self, normalize_by_size=True):
        self.normalize_by_size = normalize_by_size

    def apply_deltas(self, deltas, boxes):
        deltas = F.relu(deltas)
        boxes = boxes.to(deltas.dtype)
        ctr_x = 0.5 * (boxes[:, 0] + boxes[:, 2])
        ctr_y = 0.5 * (boxes[:, 1] + boxes[:, 3])
        if self.normalize_by_size:
            stride_w = boxes[:, 2] - boxes[:, 0]
            stride_h = boxes[:, 3] - boxes[:, 1]
            strides = torch.stack([stride_w, stride_h, stride_w, stride_h], axis=1)
            deltas = deltas * strides
        l = deltas[:, 0::4]
        t = deltas[:, 1::4]
        r = deltas[:, 2::4]
        b = deltas[:, 3::4]
        pred_boxes = torch.zeros_like(deltas)
        pred_boxes[:, 0::4] = ctr_x[:, None] - l  # x1
        pred_boxes[:, 1::4] = ctr_y[:, None] - t  # y1
        pred_boxes[:, 2::4] = ctr_x[:, None] + r  # x2
        pred_boxes[:, 3::4] = ctr_y[:, None] + b  # y2
        return pred_boxes


# ---------------------------------------------------------------------------- detectron2.modeling.matcher / .sampling
class Matcher(object):
    def __init__(self, thresholds: List[float], labels: List[int], allow_low_quality_matches: bool = False):
        thresholds = thresholds[:]
        assert thresholds[0] > 0
        thresholds.insert(0, -float("inf"))
        thresholds.append(float("inf"))
        assert all([low <= high for (low, high) in zip(thresholds[:-1], thresholds[1:])])
        assert all([l in [-1, 0, 1] for l in labels])
        assert len(labels) == len(thresholds) - 1
        self.thresholds = thresholds
        self.labels = labels
        self.allow_low_quality_matches = allow_low_quality_matches

    def __call__(self, match_quality_matrix):
        assert match_quality_matrix.dim() == 2
        if match_quality_matrix.numel() == 0:
            default_matches = match_quality_matrix.new_full((match_quality_matrix.size(1),), 0, dtype=torch.int64)
            default_match_labels = match_quality_matrix.new_full((match_quality_matrix.size(1),), self.labels[0],
                                                                 dtype=torch.int8)
            return default_matches, default_match_labels
        assert torch.all(match_quality_matrix >= 0)
        matched_vals, matches = match_quality_matrix.max(dim=0)
        match_labels = matches.new_full(matches.size(), 1, dtype=torch.int8)
        for (l, low, high) in zip(self.labels, self.thresholds[:-1], self.thresholds[1:]):
            low_high = (matched_vals >= low) & (matched_vals < high)
            match_labels[low_high] = l
        if self.allow_low_quality_matches:
            self.set_low_quality_matches_(match_labels, match_quality_matrix)
        return matches, match_labels

    def set_low_quality_matches_(self, match_labels, match_quality_matrix):
        highest_quality_foreach_gt, _ = match_quality_matrix.max(dim=1)
        _, pred_inds_with_highest_quality = nonzero_tuple(match_quality_matrix == highest_quality_foreach_gt[:, None])
        match_labels[pred_inds_with_highest_quality] = 1


def subsample_labels(labels: torch.Tensor, num_samples: int, positive_fraction: float, bg_label: int):
    positive = nonzero_tuple((labels != -1) & (labels != bg_label))[0]
    negative = nonzero_tuple(labels == bg_label)[0]
    num_pos = int(num_samples * positive_fraction)
    num_pos = min(positive.numel(), num_pos)
    num_neg = num_samples - num_pos
    num_neg = min(negative.numel(), num_neg)
    perm1 = torch.randperm(positive.numel(), device=positive.device)[:num_pos]
    perm2 = torch.randperm(negative.numel(), device=negative.device)[:num_neg]
    return positive[perm1], negative[perm2]


# ------------------------------------------------------------------ detectron2.modeling.proposal_generator.proposal_utils
def add_ground_truth_to_proposals_single_image(gt, proposals: Instances) -> Instances:
    if isinstance(gt, Boxes):
        gt = Instances(proposals.image_size, gt_boxes=gt)
    gt_boxes = gt.gt_boxes
    device = proposals.objectness_logits.device
    # objectness logit of an appended GT box: P(object) = sigmoid(logit) =~ 1
    gt_logit_value = math.log((1.0 - 1e-10) / (1 - (1.0 - 1e-10)))
    gt_logits = gt_logit_value * torch.ones(len(gt_boxes), device=device)
    gt_proposal = Instances(proposals.image_size, **gt.get_fields())
    gt_proposal.proposal_boxes = gt_boxes
    gt_proposal.objectness_logits = gt_logits
    for key in proposals.get_fields().keys():
        assert gt_proposal.has(key), "The attribute '{}' in `proposals` does not exist in `gt`".format(key)
    return Instances.cat([proposals, gt_proposal])


def add_ground_truth_to_proposals(gt, proposals: List[Instances]) -> List[Instances]:
    assert gt is not None
    if len(proposals) != len(gt):
        raise ValueError("proposals and gt should have the same length as the number of images!")
    if len(proposals) == 0:
        return proposals
    return [add_ground_truth_to_proposals_single_image(gt_i, proposals_i) for gt_i, proposals_i in zip(gt, proposals)]


# ---------------------------------------------------------------------------------------- detectron2.modeling.poolers
def assign_boxes_to_levels(box_lists: List[Boxes], min_level: int, max_level: int, canonical_box_size: int,
                           canonical_level: int):
    box_sizes = torch.sqrt(cat([boxes.area() for boxes in box_lists]))
    # Eqn.(1) in FPN paper
    level_assignments = torch.floor(canonical_level + torch.log2(box_sizes / canonical_box_size + 1e-8))
    level_assignments = torch.clamp(level_assignments, min=min_level, max=max_level)
    return level_assignments.to(torch.int64) - min_level


def _fmt_box_list(box_tensor, batch_index: int):
    repeated_index = torch.full_like(box_tensor[:, :1], batch_index, dtype=box_tensor.dtype, device=box_tensor.device)
    return cat((repeated_index, box_tensor), dim=1)


def convert_boxes_to_pooler_format(box_lists: List[Boxes]):
    return cat([_fmt_box_list(box_list.tensor, i) for i, box_list in enumerate(box_lists)], dim=0)


class ROIPooler(nn.Module):
    def __init__(self, output_size, scales, sampling_ratio, pooler_type, canonical_box_size=224, canonical_level=4):
        super().__init__()
        if isinstance(output_size, int):
            output_size = (output_size, output_size)
        assert len(output_size) == 2
        self.output_size = output_size
        if pooler_type == "ROIAlign":
            self.level_poolers = nn.ModuleList(
                ROIAlign(output_size, spatial_scale=scale, sampling_ratio=sampling_ratio, aligned=False) for scale in scales)
        elif pooler_type == "ROIAlignV2":
            self.level_poolers = nn.ModuleList(
                ROIAlign(output_size, spatial_scale=scale, sampling_ratio=sampling_ratio, aligned=True) for scale in scales)
        else:
            raise ValueError("Unknown pooler type: {}".format(pooler_type))
        min_level = -(math.log2(scales[0]))
        max_level = -(math.log2(scales[-1]))
        assert math.isclose(min_level, int(min_level)) and math.isclose(max_level, int(max_level))
        self.min_level = int(min_level)
        self.max_level = int(max_level)
        assert len(scales) == self.max_level - self.min_level + 1
        assert 0 <= self.min_level and self.min_level <= self.max_level
        self.canonical_level = canonical_level
        assert canonical_box_size > 0
        self.canonical_box_size = canonical_box_size

    def forward(self, x: List[torch.Tensor], box_lists: List[Boxes]):
        num_level_assignments = len(self.level_poolers)
        assert isinstance(x, list) and isinstance(box_lists, list)
        assert len(x) == num_level_assignments
        assert len(box_lists) == x[0].size(0)
        if len(box_lists) == 0:
            return torch.zeros((0, x[0].shape[1]) + self.output_size, device=x[0].device, dtype=x[0].dtype)
        pooler_fmt_boxes = convert_boxes_to_pooler_format(box_lists)
        if num_level_assignments == 1:
            return self.level_poolers[0](x[0], pooler_fmt_boxes)
        level_assignments = assign_boxes_to_levels(box_lists, self.min_level, self.max_level, self.canonical_box_size,
                                                   self.canonical_level)
        num_boxes = pooler_fmt_boxes.size(0)
        num_channels = x[0].shape[1]
        output_size = self.output_size[0]
        dtype, device = x[0].dtype, x[0].device
        output = torch.zeros((num_boxes, num_channels, output_size, output_size), dtype=dtype, device=device)
        for level, pooler in enumerate(self.level_poolers):
            inds = nonzero_tuple(level_assignments == level)[0]
            pooler_fmt_boxes_level = pooler_fmt_boxes[inds]
            output.index_put_((inds,), pooler(x[level], pooler_fmt_boxes_level))
        return output


# ------------------------------------------------------------------------------ detectron2.modeling.roi_heads.roi_heads
class ROIHeads(nn.Module):
    @configurable
    def __init__(self, *, num_classes, batch_size_per_image, positive_fraction, proposal_matcher,
                 proposal_append_gt=True):
        super().__init__()
        self.batch_size_per_image = batch_size_per_image
        self.positive_fraction = positive_fraction
        self.num_classes = num_classes
        self.proposal_matcher = proposal_matcher
        self.proposal_append_gt = proposal_append_gt

    @classmethod
    def from_config(cls, cfg):
        return {
            "batch_size_per_image": cfg.MODEL.ROI_HEADS.BATCH_SIZE_PER_IMAGE,
            "positive_fraction": cfg.MODEL.ROI_HEADS.POSITIVE_FRACTION,
            "num_classes": cfg.MODEL.ROI_HEADS.NUM_CLASSES,
            "proposal_append_gt": cfg.MODEL.ROI_HEADS.PROPOSAL_APPEND_GT,
            "proposal_matcher": Matcher(cfg.MODEL.ROI_HEADS.IOU_THRESHOLDS, cfg.MODEL.ROI_HEADS.IOU_LABELS,
                                        allow_low_quality_matches=False),
        }

    def _sample_proposals(self, matched_idxs, matched_labels, gt_classes):
        has_gt = gt_classes.numel() > 0
        if has_gt:
            gt_classes = gt_classes[matched_idxs]
            gt_classes[matched_labels == 0] = self.num_classes
            gt_classes[matched_labels == -1] = -1
        else:
            gt_classes = torch.zeros_like(matched_idxs) + self.num_classes
        sampled_fg_idxs, sampled_bg_idxs = subsample_labels(gt_classes, self.batch_size_per_image,
                                                            self.positive_fraction, self.num_classes)
        sampled_idxs = torch.cat([sampled_fg_idxs, sampled_bg_idxs], dim=0)
        return sampled_idxs, gt_classes[sampled_idxs]


# ------------------------------------------------------------------------------- detectron2.modeling.roi_heads.box_head
class FastRCNNConvFCHead(nn.Sequential):
    """fc-only form (the shipped configs: NUM_CONV 0, NUM_FC 2, FC_DIM 1024): flatten, (Linear, ReLU) x len(fc_dims)."""

    def __init__(self, input_shape: ShapeSpec, *, conv_dims: List[int], fc_dims: List[int], conv_norm=""):
        super().__init__()
        assert len(conv_dims) == 0, "d2shim: conv layers of the box head are not on the shipped path"
        assert len(fc_dims) > 0
        self._output_size = (input_shape.channels, input_shape.height, input_shape.width)
        self.fcs = []
        for k, fc_dim in enumerate(fc_dims):
            if k == 0:
                self.add_module("flatten", nn.Flatten())
            fc = nn.Linear(int(np.prod(self._output_size)), fc_dim)
            self.add_module("fc{}".format(k + 1), fc)
            self.add_module("fc_relu{}".format(k + 1), nn.ReLU())
            self.fcs.append(fc)
            self._output_size = fc_dim
        for layer in self.fcs:  # fvcore c2_xavier_fill
            nn.init.kaiming_uniform_(layer.weight, a=1)
            nn.init.constant_(layer.bias, 0)

    def forward(self, x):
        for layer in self:
            x = layer(x)
        return x

    @property
    def output_shape(self):
        o = self._output_size
        if isinstance(o, int):
            return ShapeSpec(channels=o)
        return ShapeSpec(channels=o[0], height=o[1], width=o[2])


def build_box_head(cfg, input_shape):
    num_fc = cfg.MODEL.ROI_BOX_HEAD.NUM_FC
    fc_dim = cfg.MODEL.ROI_BOX_HEAD.FC_DIM
    return FastRCNNConvFCHead(input_shape, conv_dims=[], fc_dims=[fc_dim] * num_fc)


# ------------------------------------------------------------------------------------------------------- fvcore.nn
def smooth_l1_loss(input, target, beta: float, reduction: str = "none"):
    if beta < 1e-5:
        loss = torch.abs(input - target)
    else:
        n = torch.abs(input - target)
        cond = n < beta
        loss = torch.where(cond, 0.5 * n ** 2 / beta, n - 0.5 * beta)
    if reduction == "mean":
        loss = loss.mean() if loss.numel() > 0 else 0.0 * loss.sum()
    elif reduction == "sum":
        loss = loss.sum()
    return loss


# ------------------------------------------------------------------------------------------------------- install
def _mod(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []  # behave as a package so that sub-module imports resolve through sys.modules
    sys.modules[name] = m
    return m


def install() -> None:
    """Register the stand-in modules.  Refuses to shadow a real detectron2."""
    if "detectron2" in sys.modules and not getattr(sys.modules["detectron2"], "_IS_D2SHIM", False):
        raise RuntimeError("a real detectron2 is already imported; use it instead of the shim")
    _mod("detectron2", _IS_D2SHIM=True, __version__="0.6+shim")
    _mod("detectron2.config", configurable=configurable, CfgNode=CfgNode)
    _mod("detectron2.layers", batched_nms=batched_nms, cat=cat, nonzero_tuple=nonzero_tuple, cross_entropy=cross_entropy,
         ShapeSpec=ShapeSpec, Conv2d=Conv2d, ROIAlign=ROIAlign, ciou_loss=_not_on_path("ciou_loss"),
         diou_loss=_not_on_path("diou_loss"))
    _mod("detectron2.structures", Boxes=Boxes, Instances=Instances, ImageList=ImageList, pairwise_iou=pairwise_iou,
         pairwise_intersection=pairwise_intersection)
    _mod("detectron2.data", MetadataCatalog=MetadataCatalog)
    _mod("detectron2.utils")
    _mod("detectron2.utils.events", get_event_storage=get_event_storage, EventStorage=EventStorage)
    _mod("detectron2.utils.memory", retry_if_cuda_oom=retry_if_cuda_oom)
    _mod("detectron2.modeling", build_anchor_generator=_not_on_path("build_anchor_generator"),
         build_rpn_head=_not_on_path("build_rpn_head"), RPN_HEAD_REGISTRY=RPN_HEAD_REGISTRY,
         PROPOSAL_GENERATOR_REGISTRY=PROPOSAL_GENERATOR_REGISTRY, ROI_HEADS_REGISTRY=ROI_HEADS_REGISTRY)
    _mod("detectron2.modeling.anchor_generator", DefaultAnchorGenerator=DefaultAnchorGenerator)
    _mod("detectron2.modeling.box_regression", Box2BoxTransform=Box2BoxTransform,
         Box2BoxTransformLinear=Box2BoxTransformLinear,
         _dense_box_regression_loss=_not_on_path("_dense_box_regression_loss"))
    _mod("detectron2.modeling.matcher", Matcher=Matcher)
    _mod("detectron2.modeling.sampling", subsample_labels=subsample_labels)
    _mod("detectron2.modeling.poolers", ROIPooler=ROIPooler, assign_boxes_to_levels=assign_boxes_to_levels,
         convert_boxes_to_pooler_format=convert_boxes_to_pooler_format)
    _mod("detectron2.modeling.proposal_generator")
    _mod("detectron2.modeling.proposal_generator.proposal_utils",
         add_ground_truth_to_proposals=add_ground_truth_to_proposals,
         add_ground_truth_to_proposals_single_image=add_ground_truth_to_proposals_single_image)
    _mod("detectron2.modeling.roi_heads")
    _mod("detectron2.modeling.roi_heads.roi_heads", ROIHeads=ROIHeads, ROI_HEADS_REGISTRY=ROI_HEADS_REGISTRY)
    _mod("detectron2.modeling.roi_heads.box_head", build_box_head=build_box_head, FastRCNNConvFCHead=FastRCNNConvFCHead)
    _mod("fvcore")
    _mod("fvcore.nn", smooth_l1_loss=smooth_l1_loss, giou_loss=_not_on_path("giou_loss"))


def import_reference(ref_root: str = "/root/reference"):
    """Import the unmodified reference modules of the RoI path from ``ref_root`` (file by file: the reference's own
    ``openset_rcnn/modeling/__init__.py`` pulls names that its empty sub-package ``__init__``s do not export)."""
    import importlib.util
    import os
    install()
    pkgs = ["openset_rcnn", "openset_rcnn.data", "openset_rcnn.modeling", "openset_rcnn.modeling.proposal_generator",
            "openset_rcnn.modeling.roi_heads"]
    for p in pkgs:
        m = types.ModuleType(p)
        m.__path__ = [os.path.join(ref_root, *p.split("."))]
        sys.modules[p] = m

    def load(modname):
        path = os.path.join(ref_root, *modname.split(".")) + ".py"
        spec = importlib.util.spec_from_file_location(modname, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[modname] = mod
        spec.loader.exec_module(mod)
        return mod

    ref = types.SimpleNamespace()
    ref.graspnet_meta = load("openset_rcnn.data.graspnet_meta")
    ref.box_regression_w_iou = load("openset_rcnn.modeling.box_regression_w_iou")
    ref.find_top_proposals = load("openset_rcnn.modeling.find_top_proposals")
    ref.classification_free_rpn = load("openset_rcnn.modeling.proposal_generator.classification_free_rpn")
    ref.prototype_learning_network = load("openset_rcnn.modeling.roi_heads.prototype_learning_network")
    ref.softmax_classifier = load("openset_rcnn.modeling.roi_heads.softmax_classifier")
    ref.osrcnn_fast_rcnn = load("openset_rcnn.modeling.roi_heads.osrcnn_fast_rcnn")
    ref.osrcnn_roi_heads = load("openset_rcnn.modeling.roi_heads.osrcnn_roi_heads")
    return ref
