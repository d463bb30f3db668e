"""Containers at the drop-in boundary: Boxes / Instances / FreeInstances / ImageList.

Minimal restatements of the detectron2 v0.5 structures the reference's trainer and meta-arch touch
(`pt/structures/instances.py:22-46`, `pt/engine/trainer.py:179-257,557-590`), plus one extension:
fields produced by the device pipeline may live in fixed-capacity buffers with a device-side
count (`_count`); `__len__` / `trim()` materialise the exact length lazily (one host sync), the
hot path itself never does.
"""
from typing import Any, Dict, List, Tuple

import torch


class Boxes:
    def __init__(self, tensor: torch.Tensor):
        if not isinstance(tensor, torch.Tensor):
            tensor = torch.as_tensor(tensor, dtype=torch.float32)
        tensor = tensor.to(torch.float32)
        if tensor.numel() == 0:
            tensor = tensor.reshape((-1, 4))
        assert tensor.dim() == 2 and tensor.size(-1) == 4, tensor.size()
        self.tensor = tensor

    def clone(self):
        return Boxes(self.tensor.clone())

    def to(self, *a, **k):
        return Boxes(self.tensor.to(*a, **k))

    def area(self):
        b = self.tensor
        return (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])

    def clip(self, box_size: Tuple[int, int]):
        h, w = box_size
        self.tensor = torch.stack((self.tensor[:, 0].clamp(0, w), self.tensor[:, 1].clamp(0, h),
                                   self.tensor[:, 2].clamp(0, w), self.tensor[:, 3].clamp(0, h)), dim=-1)

    def nonempty(self, threshold: float = 0.0):
        b = self.tensor
        return ((b[:, 2] - b[:, 0]) > threshold) & ((b[:, 3] - b[:, 1]) > threshold)

    def __getitem__(self, item):
        if isinstance(item, int):
            return Boxes(self.tensor[item].view(1, -1))
        return Boxes(self.tensor[item])

    def __len__(self):
        return self.tensor.shape[0]

    @property
    def device(self):
        return self.tensor.device

    @classmethod
    def cat(cls, boxes_list):
        if len(boxes_list) == 0:
            return cls(torch.empty(0))
        return cls(torch.cat([b.tensor for b in boxes_list], dim=0))

    def __repr__(self):
        return "Boxes(" + str(self.tensor) + ")"


class Instances:
    def __init__(self, image_size: Tuple[int, int], **kwargs: Any):
        self._image_size = image_size
        self._fields: Dict[str, Any] = {}
        self._count = None  # optional device int32 scalar: number of valid leading rows
        for k, v in kwargs.items():
            self.set(k, v)

    @property
    def image_size(self):
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

    def to(self, *args: Any, **kwargs: Any):
        ret = type(self)(self._image_size)
        for k, v in self._fields.items():
            if hasattr(v, "to"):
                v = v.to(*args, **kwargs)
            ret.set(k, v)
        ret._count = self._count
        return ret

    def __getitem__(self, item):
        ret = type(self)(self._image_size)
        for k, v in self._fields.items():
            ret.set(k, v[item])
        return ret

    def __len__(self) -> int:
        for v in self._fields.values():
            return v.__len__()
        raise NotImplementedError("Empty Instances does not support __len__!")

    @classmethod
    def cat(cls, instance_lists: List["Instances"]):
        assert len(instance_lists) > 0
        if len(instance_lists) == 1:
            return instance_lists[0]
        ret = cls(instance_lists[0].image_size)
        for k in instance_lists[0]._fields.keys():
            values = [i.get(k) for i in instance_lists]
            v0 = values[0]
            if isinstance(v0, torch.Tensor):
                values = torch.cat(values, dim=0)
            elif hasattr(type(v0), "cat"):
                values = type(v0).cat(values)
            else:
                raise ValueError("Unsupported type {} for concatenation".format(type(v0)))
            ret.set(k, values)
        return ret

    def __repr__(self):
        s = self.__class__.__name__ + "(image_height={}, image_width={}, fields=[{}])".format(
            self._image_size[0], self._image_size[1], ", ".join(f"{k}: {v}" for k, v in self._fields.items()))
        return s


class FreeInstances(Instances):
    """`pt/structures/instances.py:22-46`: Instances whose `set` skips the equal-length check."""

    def set(self, name: str, value: Any) -> None:
        self._fields[name] = value

    def valid_count(self):
        """Device int32 tensor with the number of valid rows (None when every row is valid)."""
        return self._count

    def trim(self):
        """Exact-length copy (host sync). Only for inspection / tests; the hot path keeps capacity."""
        if self._count is None:
            return self
        n = int(self._count.item())
        ret = FreeInstances(self._image_size)
        for k, v in self._fields.items():
            ret.set(k, v[:n])
        return ret


class ImageList:
    def __init__(self, tensor: torch.Tensor, image_sizes: List[Tuple[int, int]]):
        self.tensor = tensor
        self.image_sizes = image_sizes

    def __len__(self):
        return len(self.image_sizes)
