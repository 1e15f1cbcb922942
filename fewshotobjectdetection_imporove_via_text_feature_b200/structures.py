"""Minimal detectron2-compatible containers used at the ROI-head boundary.

When detectron2 is importable its own classes are used (true drop-in); otherwise these stand-ins provide
the subset of the v0.3 API that the head's callers touch: `Boxes`, `Instances`,
`ShapeSpec`, `ImageList`, `Registry`, `pairwise_iou`.
"""
import torch

try:  # pragma: no cover - detectron2 is absent from the build image
    from detectron2.layers import ShapeSpec
    from detectron2.structures import Boxes, ImageList, Instances, pairwise_iou
    from detectron2.utils.registry import Registry
    HAVE_DETECTRON2 = True
except Exception:  # noqa: BLE001
    HAVE_DETECTRON2 = False

    class ShapeSpec:
        def __init__(self, channels=None, height=None, width=None, stride=None):
            self.channels, self.height, self.width, self.stride = channels, height, width, stride

    class Registry:
        """name -> object map with decorator registration (`@REG.register()`)."""

        def __init__(self, name):
            self._name = name
            self._map = {}

        def _add(self, name, obj):
            if name in self._map:
                raise KeyError("'%s' already registered in '%s'" % (name, self._name))
            self._map[name] = obj

        def register(self, obj=None):
            if obj is None:
                def deco(o):
                    self._add(o.__name__, o)
                    return o
                return deco
            self._add(obj.__name__, obj)
            return obj

        def get(self, name):
            try:
                return self._map[name]
            except KeyError:
                raise KeyError("No object named '%s' found in '%s' registry!" % (name, self._name))

        def __contains__(self, name):
            return name in self._map

    class Boxes:
        """(N,4) XYXY absolute boxes, fp32."""

        def __init__(self, tensor):
            if not isinstance(tensor, torch.Tensor):
                tensor = torch.as_tensor(tensor, dtype=torch.float32)
            tensor = tensor.to(torch.float32)
            if tensor.numel() == 0:
                tensor = tensor.reshape(0, 4)
            if tensor.dim() != 2 or tensor.shape[-1] != 4:
                raise ValueError("Boxes expects (N,4), got %s" % (tuple(tensor.shape),))
            self.tensor = tensor

        def clone(self):
            return Boxes(self.tensor.clone())

        def to(self, *a, **k):
            return Boxes(self.tensor.to(*a, **k))

        def area(self):
            t = self.tensor
            return (t[:, 2] - t[:, 0]) * (t[:, 3] - t[:, 1])

        def clip(self, box_size):
            h, w = box_size
            self.tensor[:, 0::2].clamp_(min=0, max=w)
            self.tensor[:, 1::2].clamp_(min=0, max=h)

        def nonempty(self, threshold=0.0):
            t = self.tensor
            return ((t[:, 2] - t[:, 0]) > threshold) & ((t[:, 3] - t[:, 1]) > threshold)

        def __getitem__(self, item):
            if isinstance(item, int):
                return Boxes(self.tensor[item].view(1, -1))
            return Boxes(self.tensor[item])

        def __len__(self):
            return self.tensor.shape[0]

        def __repr__(self):
            return "Boxes(%r)" % (self.tensor,)

        @property
        def device(self):
            return self.tensor.device

        @classmethod
        def cat(cls, boxes_list):
            if len(boxes_list) == 0:
                return cls(torch.empty(0))
            return cls(torch.cat([b.tensor for b in boxes_list], dim=0))

    def pairwise_iou(boxes1, boxes2):
        a1, a2 = boxes1.area(), boxes2.area()
        b1, b2 = boxes1.tensor, boxes2.tensor
        wh = (torch.min(b1[:, None, 2:], b2[:, 2:]) - torch.max(b1[:, None, :2], b2[:, :2])).clamp_(min=0)
        inter = wh[..., 0] * wh[..., 1]
        return torch.where(inter > 0, inter / (a1[:, None] + a2 - inter), torch.zeros((), dtype=inter.dtype, device=inter.device))

    class Instances:
        """Per-image bag of equally long fields (`pred_boxes`, `scores`, `gt_classes`, ...)."""

        def __init__(self, image_size, **fields):
            object.__setattr__(self, "_image_size", tuple(image_size))
            object.__setattr__(self, "_fields", {})
            for k, v in fields.items():
                self.set(k, v)

        @property
        def image_size(self):
            return self._image_size

        def __setattr__(self, name, val):
            if name.startswith("_"):
                object.__setattr__(self, name, val)
            else:
                self.set(name, val)

        def __getattr__(self, name):
            fields = object.__getattribute__(self, "_fields")
            if name not in fields:
                raise AttributeError("Cannot find field '%s' in the given Instances!" % name)
            return fields[name]

        def set(self, name, value):
            n = len(value)
            if len(self._fields):
                assert len(self) == n, "Adding a field of length %d to Instances of length %d" % (n, len(self))
            self._fields[name] = value

        def has(self, name):
            return name in self._fields

        def remove(self, name):
            del self._fields[name]

        def get(self, name):
            return self._fields[name]

        def get_fields(self):
            return self._fields

        def to(self, *a, **k):
            out = Instances(self._image_size)
            for n, v in self._fields.items():
                out.set(n, v.to(*a, **k) if hasattr(v, "to") else v)
            return out

        def __getitem__(self, item):
            out = Instances(self._image_size)
            for n, v in self._fields.items():
                out.set(n, v[item])
            return out

        def __len__(self):
            for v in self._fields.values():
                return len(v)
            raise NotImplementedError("Empty Instances does not support __len__!")

        @staticmethod
        def cat(instance_lists):
            assert len(instance_lists) > 0
            out = Instances(instance_lists[0].image_size)
            for k in instance_lists[0]._fields:
                vals = [i.get(k) for i in instance_lists]
                if isinstance(vals[0], torch.Tensor):
                    vals = torch.cat(vals, dim=0)
                elif hasattr(type(vals[0]), "cat"):
                    vals = type(vals[0]).cat(vals)
                else:
                    raise ValueError("Unsupported type %s for concatenation" % type(vals[0]))
                out.set(k, vals)
            return out

    class ImageList:
        def __init__(self, tensor, image_sizes):
            self.tensor, self.image_sizes = tensor, [tuple(s) for s in image_sizes]

        def __len__(self):
            return len(self.image_sizes)

        @staticmethod
        def from_tensors(tensors, size_divisibility=0, pad_value=0.0):
            sizes = [tuple(t.shape[-2:]) for t in tensors]
            mh, mw = max(s[0] for s in sizes), max(s[1] for s in sizes)
            if size_divisibility > 1:
                mh = (mh + size_divisibility - 1) // size_divisibility * size_divisibility
                mw = (mw + size_divisibility - 1) // size_divisibility * size_divisibility
            out = tensors[0].new_full((len(tensors),) + tuple(tensors[0].shape[:-2]) + (mh, mw), pad_value)
            for i, t in enumerate(tensors):
                out[i, ..., : t.shape[-2], : t.shape[-1]].copy_(t)
            return ImageList(out, sizes)
