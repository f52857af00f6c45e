"""Container types at the drop-in boundary.

If detectron2 is importable its own ``Boxes`` / ``Instances`` are used, so objects
produced here flow straight into the reference's code
(``find_top_proposals.py:122-127`` builds ``Instances`` with ``proposal_boxes`` /
``objectness_logits``; ``osrcnn_roi_heads.py:306`` passes ``[x.proposal_boxes for x in
proposals]`` to the pooler).  detectron2 is not installed in this image, so the shims
below provide the small part of that surface the RoI path needs.
"""
from __future__ import annotations

from typing import Any, Dict, List, Tuple

import torch

try:  # pragma: no cover - detectron2 is absent in this image
    from detectron2.structures import Boxes, Instances  # type: ignore

    HAVE_DETECTRON2 = True
except Exception:  # noqa: BLE001
    HAVE_DETECTRON2 = False

    class Boxes:  # type: ignore[no-redef]
        """(K,4) fp32 xyxy absolute-pixel boxes."""

        def __init__(self, tensor: torch.Tensor):
            if not isinstance(tensor, torch.Tensor):
                tensor = torch.as_tensor(tensor, dtype=torch.float32)
            tensor = tensor.to(torch.float32)
            if tensor.numel() == 0:
                tensor = tensor.reshape((-1, 4))
            assert tensor.dim() == 2 and tensor.size(-1) == 4, tensor.size()
            self.tensor = tensor

        def clone(self) -> "Boxes":
            return Boxes(self.tensor.clone())

        def to(self, *args, **kwargs) -> "Boxes":
            return Boxes(self.tensor.to(*args, **kwargs))

        def area(self) -> torch.Tensor:
            b = self.tensor
            return (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])

        def clip(self, box_size: Tuple[int, int]) -> None:
            assert torch.isfinite(self.tensor).all(), "Box tensor contains infinite or NaN!"
            h, w = box_size
            t = self.tensor
            self.tensor = torch.stack(
                (t[:, 0].clamp(min=0, max=w), t[:, 1].clamp(min=0, max=h),
                 t[:, 2].clamp(min=0, max=w), t[:, 3].clamp(min=0, max=h)), dim=-1)

        def nonempty(self, threshold: float = 0.0) -> torch.Tensor:
            b = self.tensor
            return ((b[:, 2] - b[:, 0]) > threshold) & ((b[:, 3] - b[:, 1]) > threshold)

        def __getitem__(self, item) -> "Boxes":
            if isinstance(item, int):
                return Boxes(self.tensor[item].view(1, -1))
            return Boxes(self.tensor[item])

        def __len__(self) -> int:
            return self.tensor.shape[0]

        def __repr__(self) -> str:
            return "Boxes(" + str(self.tensor) + ")"

        @property
        def device(self):
            return self.tensor.device

        @classmethod
        def cat(cls, boxes_list: List["Boxes"]) -> "Boxes":
            if len(boxes_list) == 0:
                return cls(torch.empty(0, 4))
            return cls(torch.cat([b.tensor for b in boxes_list], dim=0))

    class Instances:  # type: ignore[no-redef]
        """image_size + equally long named fields."""

        def __init__(self, image_size: Tuple[int, int], **kwargs: Any):
            object.__setattr__(self, "_image_size", image_size)
            object.__setattr__(self, "_fields", {})
            for k, v in kwargs.items():
                self.set(k, v)

        @property
        def image_size(self) -> Tuple[int, int]:
            return self._image_size

        def __setattr__(self, name: str, val: Any) -> None:
            if name.startswith("_"):
                object.__setattr__(self, name, val)
            else:
                self.set(name, val)

        def __getattr__(self, name: str) -> Any:
            if name == "_fields" or name not in self._fields:
                raise AttributeError("Cannot find field '{}' in the given Instances!".format(name))
            return self._fields[name]

        def set(self, name: str, value: Any) -> None:
            data_len = len(value)
            if len(self._fields):
                assert len(self) == data_len, "Adding a field of length {} to a Instances of length {}".format(
                    data_len, len(self))
            self._fields[name] = value

        def has(self, name: str) -> bool:
            return name in self._fields

        def remove(self, name: str) -> None:
            del self._fields[name]

        def get(self, name: str) -> Any:
            return self._fields[name]

        def get_fields(self) -> Dict[str, Any]:
            return self._fields

        def to(self, *args, **kwargs) -> "Instances":
            ret = Instances(self._image_size)
            for k, v in self._fields.items():
                if hasattr(v, "to"):
                    v = v.to(*args, **kwargs)
                ret.set(k, v)
            return ret

        def __len__(self) -> int:
            for v in self._fields.values():
                return v.__len__()
            raise NotImplementedError("Empty Instances does not support __len__!")

        def __getitem__(self, item) -> "Instances":
            ret = Instances(self._image_size)
            for k, v in self._fields.items():
                ret.set(k, v[item])
            return ret

        def __repr__(self) -> str:
            return "Instances(num_instances={}, image_size={}, fields=[{}])".format(
                len(self) if len(self._fields) else 0, self._image_size, ", ".join(self._fields.keys()))


def cat_rows(tensors) -> torch.Tensor:
    """``torch.cat(tensors, dim=0)`` - without the copy when the tensors are CONSECUTIVE row slices of one buffer.

    The batched stages of this package hand their per-image results out as ``torch.split`` views of one tensor
    (``Instances`` fields of image n = rows ``[off[n], off[n+1])``); the next stage of the reference's loop structure
    concatenates exactly those fields again (``PLN.inference``, ``SoftMaxClassifier.inference``, ``ROIPooler.forward``).
    Recognising the views (same storage, contiguous rows, offsets that follow on) turns that 130 MB feature copy into
    pointer arithmetic on the host; anything else falls through to ``torch.cat``."""
    tensors = list(tensors)
    if len(tensors) == 1:
        return tensors[0]
    t0 = tensors[0]
    if t0.dim() >= 1 and not t0.requires_grad and t0.is_contiguous():
        tail = t0.shape[1:]
        row = 1
        for d in tail:
            row *= int(d)
        base = t0.untyped_storage().data_ptr()
        off = t0.storage_offset()
        rows = 0
        ok = row > 0
        if ok:
            for t in tensors:
                if (t.dtype != t0.dtype or t.shape[1:] != tail or t.requires_grad or not t.is_contiguous()
                        or t.device != t0.device or t.untyped_storage().data_ptr() != base
                        or t.storage_offset() != off + rows * row):
                    ok = False
                    break
                rows += int(t.shape[0])
        if ok:
            return t0.as_strided((rows,) + tuple(tail), t0.stride(), off)
    return torch.cat(tensors, dim=0)


def flat_prefixes(begins, counts, device) -> Tuple[torch.Tensor, torch.Tensor]:
    """Rows ``[begins[n], begins[n] + counts[n])`` of every segment n, concatenated: ``(positions, begin of the row's
    segment)`` as int64 device tensors, built on the host from host-known counts (one small H2D copy) - the batched
    replacement for a Python loop of per-image slices."""
    import numpy as np
    b = np.asarray(begins, dtype=np.int64)
    c = np.asarray(counts, dtype=np.int64)
    total = int(c.sum())
    seg_first = np.cumsum(c) - c                       # output position of each segment's first row
    within = np.arange(total, dtype=np.int64) - np.repeat(seg_first, c)
    rep_b = np.repeat(b, c)
    both = torch.from_numpy(np.stack((rep_b + within, rep_b))).to(device, non_blocking=False)
    return both[0], both[1]


def boxes_view(tensor: torch.Tensor) -> "Boxes":
    """``Boxes`` around an (n, 4) fp32 tensor the caller just produced: skips the constructor's ``as_tensor`` / dtype / shape
    handling (~4 us per object - with 32 images and four batched stages the per-image result objects are a measurable part
    of the inference step).  Works for detectron2's ``Boxes`` as for the stand-in: both hold one attribute, ``tensor``."""
    b = Boxes.__new__(Boxes)
    b.tensor = tensor
    return b


def make_instances(image_size, **fields) -> "Instances":
    """``Instances(image_size)`` holding ``fields`` - equally long by construction (the batched stages cut every field with
    the same per-image counts), so the per-field length assertion of ``set`` is skipped."""
    r = Instances(image_size)
    r._fields.update(fields)
    return r
