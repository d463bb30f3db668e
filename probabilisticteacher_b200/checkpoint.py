"""Checkpoint I/O of the B200 path in the reference's on-disk format.

The reference saves ONE file per checkpoint through fvcore's `Checkpointer.save`
(`pt/engine/trainer.py:104-111`): `{"model": state_dict of EnsembleTSModel, "optimizer": ..., "scheduler": ...,
"iteration": i}`, where `EnsembleTSModel` (`pt/modeling/meta_arch/ts_ensemble.py:20-30`) prefixes the two detectors
with `modelTeacher.` / `modelStudent.`, and writes the file name into `<save_dir>/last_checkpoint`.
`DetectionTSCheckpointer._load_model` (`pt/checkpoint/detection_checkpoint.py:25-110`) loads either the whole
ensemble or -- for a Caffe2-authored (pre-trained backbone) file -- the student only, dropping entries whose shape
does not match and reporting them. The same behaviour is provided here over the flat parameter arenas: tensors are
converted between the arena layouts and the reference's names / layouts by `ParamArena.state_dict` /
`load_state_dict`, so a checkpoint written by the reference loads into this path and vice versa.

The VGG constructor's ImageNet initialisation (`pt/modeling/backbone/vgg.py:127-152`: torchvision-style
`features.N.{weight,bias}` keys of `vgg16_caffe.pth` mapped onto `vgg_block{b}.0.conv{c}`) is `load_vgg16_caffe`.

Not restated: detectron2's `align_and_update_state_dicts` name-matching heuristics (only used for files that carry
`matching_heuristics`, none of which the reference's configs point at) and pickled Caffe2 `.pkl` model-zoo files.
"""
import os
from collections import OrderedDict, namedtuple

import torch

IncompatibleKeys = namedtuple("IncompatibleKeys", ["missing_keys", "unexpected_keys", "incorrect_shapes"])

# torchvision vgg16 `features` indices of the 13 convolutions, in order (vgg.py:129-135)
_VGG16_FEATURE_IDX = (0, 2, 5, 7, 10, 12, 14, 17, 19, 21, 24, 26, 28)
_VGG16_BLOCKS = (2, 2, 3, 3, 3)


def vgg16_caffe_key_map(prefix="backbone."):
    """{`features.N.weight|bias` -> `<prefix>vgg_block{b}.0.conv{c}.weight|bias`} (vgg.py:129-147)."""
    out = OrderedDict()
    it = iter(_VGG16_FEATURE_IDX)
    for b, n in enumerate(_VGG16_BLOCKS, start=1):
        for c in range(1, n + 1):
            idx = next(it)
            for kind in ("weight", "bias"):
                out[f"features.{idx}.{kind}"] = f"{prefix}vgg_block{b}.0.conv{c}.{kind}"
    return out


def strip_prefix_if_present(state_dict, prefix):
    """fvcore `_strip_prefix_if_present`: strips `prefix` only when EVERY key carries it (in place)."""
    keys = sorted(state_dict.keys())
    if not keys or not all(k.startswith(prefix) for k in keys):
        return state_dict
    for k in keys:
        state_dict[k[len(prefix):]] = state_dict.pop(k)
    return state_dict


def _to_tensors(state_dict):
    """`Checkpointer._convert_ndarray_to_tensor`: numpy arrays (model-zoo pickles) become tensors."""
    import numpy as np
    for k in list(state_dict.keys()):
        v = state_dict[k]
        if isinstance(v, np.ndarray):
            state_dict[k] = torch.from_numpy(v)
        elif not isinstance(v, torch.Tensor):
            raise ValueError(f"Unsupported type found in checkpoint! {k}: {type(v)}")
    return state_dict


def load_filtered(model, state_dict):
    """Non-strict load of `state_dict` into one detector with the shape filter of
    `detection_checkpoint.py:87-104` (work-around for pytorch#24139: mismatching entries are dropped and
    reported instead of raising). Returns IncompatibleKeys."""
    own = model.state_dict()
    incorrect = []
    for k in list(state_dict.keys()):
        if k in own:
            sm, sc = tuple(own[k].shape), tuple(state_dict[k].shape)
            if sm != sc:
                incorrect.append((k, sc, sm))
                state_dict.pop(k)
    inc = model.load_state_dict(state_dict, strict=False)
    return IncompatibleKeys(list(inc.missing_keys), list(inc.unexpected_keys), incorrect)


class EnsembleTSModel:
    """`pt/modeling/meta_arch/ts_ensemble.py:20-30`: holds teacher and student so that ONE file carries both under
    the `modelTeacher.` / `modelStudent.` prefixes. A DDP-style wrapper (anything with `.module`) is unwrapped."""

    def __init__(self, modelTeacher, modelStudent):
        self.modelTeacher = getattr(modelTeacher, "module", modelTeacher)
        self.modelStudent = getattr(modelStudent, "module", modelStudent)

    def state_dict(self):
        sd = OrderedDict()
        for prefix, m in (("modelTeacher.", self.modelTeacher), ("modelStudent.", self.modelStudent)):
            for k, v in m.state_dict().items():
                sd[prefix + k] = v
        return sd

    def load_state_dict(self, state_dict, strict=True):
        missing, unexpected, incorrect = [], [], []
        claimed = set()
        for prefix, m in (("modelTeacher.", self.modelTeacher), ("modelStudent.", self.modelStudent)):
            sub = {k[len(prefix):]: v for k, v in state_dict.items() if k.startswith(prefix)}
            claimed.update(prefix + k for k in sub)
            inc = load_filtered(m, sub)
            missing += [prefix + k for k in inc.missing_keys]
            unexpected += [prefix + k for k in inc.unexpected_keys]
            incorrect += [(prefix + k, a, b) for k, a, b in inc.incorrect_shapes]
        unexpected += [k for k in state_dict if k not in claimed]
        if strict and (missing or unexpected or incorrect):
            raise RuntimeError(f"Error(s) in loading state_dict for EnsembleTSModel: missing {missing}, "
                               f"unexpected {unexpected}, incorrect shapes {incorrect}")
        return IncompatibleKeys(missing, unexpected, incorrect)


def reference_trainable_order(arena):
    """Trainable parameter names in the order of the reference model's `named_parameters()` -- the order detectron2's
    `build_optimizer` hands the parameters to torch SGD (one param group each), hence the integer keys of a
    reference optimizer `state_dict()`. Recorded from the reference's own classes (oracle/make_golden_model.py):
    backbone blocks, RPN head (conv, objectness_logits, anchor_deltas: weight then bias each), anchor_0, box head,
    cls_score, bbox_pred. (Differs from the arena order, which keeps the fused head blocks together.)"""
    have = {name for name, _, _, trainable in arena.exposed_parameters(arena.momentum) if trainable}
    order = [f"{name}.{kind}" for name, _, _, trainable in arena.conv_specs if trainable for kind in ("weight", "bias")]
    for mod in ("proposal_generator.rpn_head.conv", "proposal_generator.rpn_head.objectness_logits",
                "proposal_generator.rpn_head.anchor_deltas"):
        order += [mod + ".weight", mod + ".bias"]
    order.append("proposal_generator.anchor_generator.anchor_0")
    for mod in ("roi_heads.box_head.fc1", "roi_heads.box_head.fc2", "roi_heads.box_predictor.cls_score",
                "roi_heads.box_predictor.bbox_pred"):
        order += [mod + ".weight", mod + ".bias"]
    return [n for n in order if n in have]


class ArenaSGDState:
    """Checkpointable standing in for the reference's `optimizer=` entry (`trainer.py:104-111`). Written in torch
    SGD's own `state_dict()` format -- `state[i]["momentum_buffer"]` in the reference layouts, one param group per
    parameter in the reference's parameter order, as detectron2 v0.5 `build_optimizer` creates them -- so the file
    loads into the reference's optimizer; the same buffers are stored by NAME under the extra key
    `momentum_buffers` (ignored by torch). Loading accepts both: by name when present, else by position with every
    shape checked (a reference-written file). The LR schedule is a pure function of `trainer.iter` here."""

    def __init__(self, trainer):
        self.trainer = trainer

    def state_dict(self):
        tr = self.trainer
        arena = tr.model.arena
        by_name = arena.momentum_state_dict()
        order = reference_trainable_order(arena)
        cfg = getattr(tr, "cfg", None)
        if cfg is not None:
            from .solver import lr_at_iter
            sol = cfg.SOLVER
            lr, mom, wd, base = (float(lr_at_iter(cfg, max(int(tr.iter) - 1, 0))), float(sol.MOMENTUM),
                                 float(sol.WEIGHT_DECAY), float(sol.BASE_LR))
        else:  # detectron2 defaults
            lr, mom, wd, base = 0.001, 0.9, 1e-4, 0.001
        state = {i: {"momentum_buffer": by_name[n]} for i, n in enumerate(order)} if int(tr.iter) > 0 else {}
        groups = [{"lr": lr, "momentum": mom, "dampening": 0, "nesterov": False, "weight_decay": wd,
                   "initial_lr": base, "params": [i]} for i in range(len(order))]
        return {"state": state, "param_groups": groups, "momentum_buffers": by_name, "param_names": order,
                "iter": int(tr.iter)}

    def load_state_dict(self, sd):
        arena = self.trainer.model.arena
        if "momentum_buffers" in sd:
            arena.load_momentum_state_dict(sd["momentum_buffers"])
            return
        if "state" not in sd or "param_groups" not in sd:
            raise ValueError("optimizer entry is neither torch SGD's state_dict nor this trainer's: keys "
                             f"{sorted(sd)}")
        # a file written by the reference trainer: integer keys = position in the reference's parameter order
        order = reference_trainable_order(arena)
        n_params = sum(len(g["params"]) for g in sd["param_groups"])
        if n_params != len(order):
            raise ValueError(f"optimizer state holds {n_params} parameters, this model trains {len(order)}: "
                             "refusing to guess the mapping (load the weights only, checkpointables=[])")
        ref_shapes = {k: tuple(v.shape) for k, v in arena.momentum_state_dict().items()}
        flat = [i for g in sd["param_groups"] for i in g["params"]]
        by_name = {}
        for pos, key in enumerate(flat):
            ent = sd["state"].get(key)
            if ent is None or ent.get("momentum_buffer") is None:
                continue  # torch creates the buffer lazily: zero here
            buf = ent["momentum_buffer"]
            if tuple(buf.shape) != ref_shapes[order[pos]]:
                raise ValueError(f"optimizer state {key}: shape {tuple(buf.shape)} does not match "
                                 f"{order[pos]} {ref_shapes[order[pos]]}")
            by_name[order[pos]] = buf
        arena.load_momentum_state_dict(by_name)


class SchedulerState:
    """`scheduler=` entry of the reference's checkpoints (`trainer.py:104-111`): torch LR schedulers store
    `last_epoch`; here the schedule is a pure function of the iteration, so loading only cross-checks it."""

    def __init__(self, trainer):
        self.trainer = trainer

    def state_dict(self):
        it = int(self.trainer.iter)
        return {"last_epoch": it, "_step_count": it + 1, "base_lrs": [float(self.trainer.cfg.SOLVER.BASE_LR)]}

    def load_state_dict(self, sd):
        self.last_epoch = int(sd.get("last_epoch", -1))


class DetectionTSCheckpointer:
    """fvcore `Checkpointer` surface (`save`, `load`, `resume_or_load`, `has_checkpoint`, `get_checkpoint_file`,
    `tag_last_checkpoint`) with the `_load_model` of `pt/checkpoint/detection_checkpoint.py:24-76`.

    model: an `EnsembleTSModel` (the trainer's case, `trainer.py:104-111`) or a single detector
    (`train_net.py:74`). checkpointables: objects with `state_dict()` / `load_state_dict()` saved next to it."""

    def __init__(self, model, save_dir="", *, save_to_disk=True, **checkpointables):
        self.model = getattr(model, "module", model)
        self.save_dir = save_dir
        self.save_to_disk = save_to_disk
        self.checkpointables = dict(checkpointables)

    # ------------------------------------------------------------------ save
    def save(self, name, **kwargs):
        if not self.save_dir or not self.save_to_disk:
            return None
        data = {"model": OrderedDict((k, v.detach().cpu()) for k, v in self.model.state_dict().items())}
        for key, obj in self.checkpointables.items():
            data[key] = obj.state_dict()
        data.update(kwargs)
        basename = f"{name}.pth"
        os.makedirs(self.save_dir, exist_ok=True)
        path = os.path.join(self.save_dir, basename)
        tmp = path + ".tmp"
        torch.save(data, tmp)
        os.replace(tmp, path)  # a reader never sees a half-written file
        self.tag_last_checkpoint(basename)
        return path

    def tag_last_checkpoint(self, last_filename_basename):
        with open(os.path.join(self.save_dir, "last_checkpoint"), "w") as f:
            f.write(last_filename_basename)

    def has_checkpoint(self):
        return bool(self.save_dir) and os.path.exists(os.path.join(self.save_dir, "last_checkpoint"))

    def get_checkpoint_file(self):
        try:
            with open(os.path.join(self.save_dir, "last_checkpoint")) as f:
                last = f.read().strip()
        except OSError:
            return ""
        return os.path.join(self.save_dir, last)

    # ------------------------------------------------------------------ load
    def load(self, path, checkpointables=None, trusted=False):
        """Returns what the file held besides the objects that consumed their entry (e.g. `iteration`). An empty
        path means "no checkpoint": the model keeps its initialisation (fvcore behaviour).
        Files are read with `torch.load(weights_only=True)` (tensors and plain containers only: everything this class
        and the reference's checkpointer write). trusted=True falls back to full unpickling for files that carry
        other objects (numpy arrays of model-zoo pickles) -- only for files whose origin you trust."""
        if not path:
            return {}
        if not os.path.isfile(path):
            raise AssertionError(f"Checkpoint {path} not found!")
        try:
            checkpoint = torch.load(path, map_location="cpu", weights_only=True)
        except Exception as e:  # noqa: BLE001  (pickle.UnpicklingError and friends)
            if not trusted:
                raise RuntimeError(f"{path} holds objects beyond tensors / plain containers ({type(e).__name__}); "
                                   "pass trusted=True to unpickle it if you trust its origin") from e
            checkpoint = torch.load(path, map_location="cpu", weights_only=False)
        if "model" not in checkpoint:  # a bare state dict (e.g. vgg16_caffe.pth-style files)
            checkpoint = {"model": checkpoint}
        self.last_incompatible = self._load_model(checkpoint)
        for key in (self.checkpointables if checkpointables is None else checkpointables):
            if key in self.checkpointables and key in checkpoint:
                self.checkpointables[key].load_state_dict(checkpoint.pop(key))
        return checkpoint

    def resume_or_load(self, path, *, resume=True):
        if resume and self.has_checkpoint():
            return self.load(self.get_checkpoint_file())
        return self.load(path, checkpointables=[])

    def _is_ensemble(self):
        return hasattr(self.model, "modelStudent")

    def _load_model(self, checkpoint):
        if checkpoint.get("matching_heuristics", False):
            raise NotImplementedError("detectron2's name-matching heuristics (align_and_update_state_dicts) are "
                                      "not restated; convert the file to reference key names first")
        sd = _to_tensors(checkpoint.pop("model"))
        strip_prefix_if_present(sd, "module.")
        if checkpoint.get("__author__", None) == "Caffe2" and self._is_ensemble():
            # pre-trained weights: only the student is updated (detection_checkpoint.py:26-38,78-110)
            return load_filtered(self.model.modelStudent, sd)
        if self._is_ensemble():
            return self.model.load_state_dict(sd, strict=False)
        return load_filtered(self.model, sd)


def load_vgg16_caffe(model, path_or_state_dict):
    """`vgg.py:127-152`: copies the 13 ImageNet conv layers of a torchvision-style VGG16 state dict
    (`features.N.weight|bias`, BGR / 0-255 Caffe weights as the reference expects) into the detector's backbone;
    everything else keeps its initialisation. Returns IncompatibleKeys of the partial load."""
    sd = path_or_state_dict
    if isinstance(sd, (str, os.PathLike)):
        sd = torch.load(sd, map_location="cpu", weights_only=True)  # a state dict: tensors only
    mapped = OrderedDict()
    for src, dst in vgg16_caffe_key_map().items():
        mapped[dst] = sd[src]  # KeyError on a file that is not a VGG16, as `state_dict[...]` at vgg.py:149
    return load_filtered(getattr(model, "module", model), mapped)
